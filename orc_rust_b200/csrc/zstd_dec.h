// Zstandard frames (RFC 8878) and LZO1X streams: the parsing shared by the host decoder of metadata sections
// (meta.cc) and the device decoder of data chunks (k_decompress.cu).  Replaces `zstd::Decoder` and
// `lzokay_native::decompress_all` at src/compression.rs:151-159, 174-183.  Everything here is sequential code: on
// the device it runs on one lane of the warp that owns the chunk, the copies it orders are done by the whole warp.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORCB_HD __host__ __device__ __forceinline__
#else
#define ORCB_HD inline
#endif

namespace orcb {
namespace zstd {

// small constant tables live in string literals: usable from host and device code alike
#define ORCB_ZT(lit, i) ((uint32_t)(uint8_t)(lit)[(i)])

ORCB_HD int highbit(uint32_t v) {  // index of the highest set bit, v != 0
#ifdef __CUDA_ARCH__
    return 31 - __clz((int)v);
#else
    int r = 0;
    while (v >>= 1) r++;
    return r;
#endif
}

constexpr int LL_MAX_LOG = 9, ML_MAX_LOG = 9, OF_MAX_LOG = 8, HUF_MAX_LOG = 11, WT_MAX_LOG = 6;
constexpr int LL_SYMS = 36, ML_SYMS = 53, OF_SYMS = 32, WT_SYMS = 16;

// Decoding tables of one frame.  FSE entry: symbol [0:8) | bits to read [8:16) | base of the next state [16:32).
// Huffman entry: symbol [0:8) | code length [8:16).
struct Tables {
    uint32_t ll[1 << LL_MAX_LOG], ml[1 << ML_MAX_LOG], of[1 << OF_MAX_LOG];
    uint16_t huf[1 << HUF_MAX_LOG];
    union {
        struct {
            uint32_t wt[1 << WT_MAX_LOG];  // FSE table of the Huffman weights
            uint8_t weights[256];
        };
        // once the literals of a block are decoded: baseline [0:24) | extra bits [24:32) of the literal length codes
        // (0..35) and, from 36 on, of the match length codes
        uint32_t codes[LL_SYMS + ML_SYMS];
    };
    int16_t norm[64];
    uint16_t sdesc[64];
};

// which tables a frame has built so far (Repeat_Mode / treeless literals need the previous ones)
struct FrameState {
    int ll_log, of_log, ml_log, huf_log;  // -1: never set
    uint32_t rep[3];
    ORCB_HD void reset() {
        ll_log = of_log = ml_log = huf_log = -1;
        rep[0] = 1;
        rep[1] = 4;
        rep[2] = 8;
    }
};

// ---- bit readers -------------------------------------------------------------------------------
// forward, least significant bit first (FSE table descriptions)
struct Fwd {
    const uint8_t* s;
    uint32_t n, bit;
    ORCB_HD uint32_t take(int k) {  // k <= 16
        const uint32_t byte = bit >> 3, sh = bit & 7;
        uint32_t v = 0;
        for (uint32_t i = 0; i < 4; i++) v |= (byte + i < n ? (uint32_t)s[byte + i] : 0u) << (8 * i);
        bit += k;
        return (v >> sh) & ((1u << k) - 1u);
    }
};

// backward: the stream ends with a 1 bit marking the end of the padding; the first bit read is the most significant
// of a value; bits below the start of the stream read as zeros and leave `bits` negative.  On the device a read is two
// aligned word loads and a funnel shift (the words may extend a few bytes beyond the stream: inside the arena, and
// never part of the value); the host reads byte by byte inside the stream.
struct Back {
    const uint8_t* s;
    uint32_t n;
    int32_t bits;  // unread bits: [0, bits)
#ifdef __CUDA_ARCH__
    const uint32_t* W;  // s rounded down to a word
    uint32_t boff;      // bit offset of s in W
#endif
    ORCB_HD bool init(const uint8_t* p, uint32_t len) {
        s = p;
        n = len;
        bits = 0;
#ifdef __CUDA_ARCH__
        W = (const uint32_t*)((uintptr_t)p & ~(uintptr_t)3);
        boff = (uint32_t)((uintptr_t)p & 3) * 8;
#endif
        if (!len || len > (1u << 27) || !p[len - 1]) return false;
        bits = (int32_t)(len - 1) * 8 + highbit(p[len - 1]);
        return true;
    }
    ORCB_HD uint32_t peek_slow(int k) const {
        if (k == 0) return 0;
        int32_t lo = bits - k;
        int below = 0;
        if (lo < 0) {
            below = -lo < 64 ? -lo : 64;
            if (below >= k) return 0;
            lo = 0;
        }
        const uint32_t byte = (uint32_t)lo >> 3, sh = (uint32_t)lo & 7;
        uint64_t v = 0;
        for (uint32_t i = 0; i < 5; i++) v |= (uint64_t)(byte + i < n ? s[byte + i] : 0u) << (8 * i);
        v >>= sh;
        const int have = k - below;
        v &= (1ull << have) - 1ull;
        return (uint32_t)(v << below);
    }
    ORCB_HD uint32_t peek(int k) const {  // k <= 31
#ifdef __CUDA_ARCH__
        const int32_t lo = bits - k;
        if (lo < 0) return peek_slow(k);
        const uint32_t a = (uint32_t)lo + boff;
        return __funnelshift_r(W[a >> 5], W[(a >> 5) + 1], a & 31) & ((1u << k) - 1u);
#else
        return peek_slow(k);
#endif
    }
    ORCB_HD uint32_t read(int k) {
        const uint32_t v = peek(k);
        bits -= k;
        return v;
    }
};

// ---- FSE ---------------------------------------------------------------------------------------
// normalized counts of a table description; false = corrupt.  `consumed`: bytes of the description.
ORCB_HD bool fse_read_norm(const uint8_t* p, uint32_t n, int max_log, int max_syms, int16_t* norm, int& log, int& nsym,
                           uint32_t& consumed) {
    Fwd f{p, n, 0};
    log = 5 + (int)f.take(4);
    if (log > max_log) return false;
    int remaining = 1 << log, sym = 0;
    while (remaining > 0 && sym < max_syms) {
        const int bits = highbit((uint32_t)remaining + 1u) + 1;
        uint32_t val = f.take(bits);
        const uint32_t lower_mask = (1u << (bits - 1)) - 1u;
        const uint32_t threshold = (1u << bits) - 1u - ((uint32_t)remaining + 1u);
        if ((val & lower_mask) < threshold) {
            f.bit -= 1;
            val &= lower_mask;
        } else if (val > lower_mask) {
            val -= threshold;
        }
        const int proba = (int)val - 1;  // -1: "less than one", takes one cell
        remaining -= proba < 0 ? -proba : proba;
        norm[sym++] = (int16_t)proba;
        if (proba == 0) {
            uint32_t rep = f.take(2);
            for (;;) {
                for (uint32_t i = 0; i < rep && sym < max_syms; i++) norm[sym++] = 0;
                if (rep != 3) break;
                rep = f.take(2);
            }
        }
    }
    if (remaining != 0 || f.bit > 8u * n) return false;
    nsym = sym;
    consumed = (f.bit + 7) >> 3;
    return true;
}

ORCB_HD bool fse_build(const int16_t* norm, int nsym, int log, uint32_t* table, uint16_t* sdesc) {
    const uint32_t size = 1u << log;
    uint32_t high = size;
    for (int s = 0; s < nsym; s++)
        if (norm[s] == -1) {
            table[--high] = (uint32_t)s;
            sdesc[s] = 1;
        }
    const uint32_t step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    uint32_t pos = 0;
    for (int s = 0; s < nsym; s++) {
        if (norm[s] <= 0) continue;
        sdesc[s] = (uint16_t)norm[s];
        for (int i = 0; i < norm[s]; i++) {
            table[pos] = (uint32_t)s;
            do pos = (pos + step) & mask;
            while (pos >= high);
        }
    }
    if (pos != 0) return false;
    for (uint32_t i = 0; i < size; i++) {
        const uint32_t s = table[i] & 0xffu;
        const uint32_t d = sdesc[s]++;
        const int nb = log - highbit(d);
        table[i] = s | ((uint32_t)nb << 8) | (((d << nb) - size) << 16);
    }
    return true;
}

// predefined distributions (RFC 8878 3.1.1.3.2.2), as value + 1 so that -1 fits a byte
ORCB_HD void fse_predefined(int which, int16_t* norm, int& log, int& nsym) {
    const char* d;
    if (which == 0) {  // literal lengths
        d = "\5\4\3\3\3\3\3\3\3\3\3\3\3\2\2\2\3\3\3\3\3\3\3\3\3\4\3\2\2\2\2\2\0\0\0\0";
        log = 6;
        nsym = 36;
    } else if (which == 1) {  // offsets
        d = "\2\2\2\2\2\2\3\3\3\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\0\0\0\0\0";
        log = 5;
        nsym = 29;
    } else {  // match lengths
        d = "\2\5\4\3\3\3\3\3\3\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\2\0\0\0\0\0\0\0";
        log = 6;
        nsym = 53;
    }
    for (int i = 0; i < nsym; i++) norm[i] = (int16_t)((int)ORCB_ZT(d, i) - 1);
}

// extra bits and baselines of the literal length / match length codes (RFC 8878 3.1.1.3.2.1.1)
ORCB_HD uint32_t ll_bits(uint32_t c) { return ORCB_ZT("\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\1\1\1\1\2\2\3\3\4\6\7\10\11\12\13\14\15\16\17\20", c); }
ORCB_HD uint32_t ml_bits(uint32_t c) {
    return ORCB_ZT("\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\0\1\1\1\1\2\2\3\3\4\4\5\7\10\11\12\13\14\15\16\17\20", c);
}
ORCB_HD uint32_t base3(const char* t, uint32_t c) { return ORCB_ZT(t, 3 * c) | (ORCB_ZT(t, 3 * c + 1) << 8) | (ORCB_ZT(t, 3 * c + 2) << 16); }
ORCB_HD uint32_t ll_base(uint32_t c) {
    return base3("\0\0\0\1\0\0\2\0\0\3\0\0\4\0\0\5\0\0\6\0\0\7\0\0\10\0\0\11\0\0\12\0\0\13\0\0\14\0\0\15\0\0\16\0\0\17\0\0\20\0\0\22\0\0\24\0\0\26\0\0\30\0\0\34\0\0\40\0\0\50\0\0\60\0\0\100\0\0\200\0\0\0\1\0\0\2\0\0\4\0\0\10\0\0\20\0\0\40\0\0\100\0\0\200\0\0\0\1", c);
}
ORCB_HD uint32_t ml_base(uint32_t c) {
    return base3("\3\0\0\4\0\0\5\0\0\6\0\0\7\0\0\10\0\0\11\0\0\12\0\0\13\0\0\14\0\0\15\0\0\16\0\0\17\0\0\20\0\0\21\0\0\22\0\0\23\0\0\24\0\0\25\0\0\26\0\0\27\0\0\30\0\0\31\0\0\32\0\0\33\0\0\34\0\0\35\0\0\36\0\0\37\0\0\40\0\0\41\0\0\42\0\0\43\0\0\45\0\0\47\0\0\51\0\0\53\0\0\57\0\0\63\0\0\73\0\0\103\0\0\123\0\0\143\0\0\203\0\0\3\1\0\3\2\0\3\4\0\3\10\0\3\20\0\3\40\0\3\100\0\3\200\0\3\0\1", c);
}

// fills T.codes; with `stride` callers of index `first` each writes its share (the 32 lanes of a warp, or one thread)
ORCB_HD void fill_codes(Tables& T, uint32_t first, uint32_t stride) {
    for (uint32_t c = first; c < (uint32_t)(LL_SYMS + ML_SYMS); c += stride)
        T.codes[c] = c < (uint32_t)LL_SYMS ? (ll_base(c) | (ll_bits(c) << 24)) : (ml_base(c - LL_SYMS) | (ml_bits(c - LL_SYMS) << 24));
}

// ---- Huffman -----------------------------------------------------------------------------------
// Huffman_Tree_Description -> decoding table; false = corrupt
ORCB_HD bool huf_read(const uint8_t* p, uint32_t n, Tables& T, int& huf_log, uint32_t& consumed) {
    if (n < 1) return false;
    const uint32_t hb = p[0];
    uint32_t nw = 0;
    if (hb >= 128) {
        nw = hb - 127;
        const uint32_t bytes = (nw + 1) / 2;
        if (1 + bytes > n) return false;
        for (uint32_t i = 0; i < nw; i++) T.weights[i] = (i & 1) ? (p[1 + i / 2] & 15) : (p[1 + i / 2] >> 4);
        consumed = 1 + bytes;
    } else {
        if (hb == 0 || 1 + hb > n) return false;
        int log, nsym;
        uint32_t used;
        if (!fse_read_norm(p + 1, hb, WT_MAX_LOG, WT_SYMS, T.norm, log, nsym, used)) return false;
        if (!fse_build(T.norm, nsym, log, T.wt, T.sdesc)) return false;
        if (used >= hb) return false;
        Back b;
        if (!b.init(p + 1 + used, hb - used)) return false;
        uint32_t s1 = b.read(log), s2 = b.read(log);
        if (b.bits < 0) return false;
        for (;;) {
            if (nw >= 254) return false;
            uint32_t e = T.wt[s1];
            T.weights[nw++] = (uint8_t)e;
            s1 = (e >> 16) + b.read((e >> 8) & 0xff);
            if (b.bits < 0) {
                T.weights[nw++] = (uint8_t)T.wt[s2];
                break;
            }
            e = T.wt[s2];
            T.weights[nw++] = (uint8_t)e;
            s2 = (e >> 16) + b.read((e >> 8) & 0xff);
            if (b.bits < 0) {
                T.weights[nw++] = (uint8_t)T.wt[s1];
                break;
            }
        }
        consumed = 1 + hb;
    }
    // the last weight completes the sum to a power of two
    uint32_t sum = 0;
    for (uint32_t i = 0; i < nw; i++) {
        const uint32_t w = T.weights[i];
        if (w > (uint32_t)HUF_MAX_LOG) return false;
        if (w) sum += 1u << (w - 1);
    }
    if (sum == 0) return false;
    const int maxbits = highbit(sum) + 1;
    if (maxbits > HUF_MAX_LOG) return false;
    const uint32_t left = (1u << maxbits) - sum;
    if (left & (left - 1)) return false;
    T.weights[nw++] = (uint8_t)(highbit(left) + 1);
    // cells: weight 1 symbols first (one cell each), then weight 2 (two cells each) ...; symbols ascending within a weight
    uint32_t start[HUF_MAX_LOG + 2];
    for (int w = 0; w <= HUF_MAX_LOG + 1; w++) start[w] = 0;
    for (uint32_t i = 0; i < nw; i++) start[T.weights[i]]++;
    uint32_t idx = 0;
    for (int w = 1; w <= maxbits; w++) {
        const uint32_t c = start[w];
        start[w] = idx;
        idx += c << (w - 1);
    }
    if (idx != (1u << maxbits)) return false;
    for (uint32_t s = 0; s < nw; s++) {
        const uint32_t w = T.weights[s];
        if (!w) continue;
        const uint32_t len = 1u << (w - 1), e = s | ((uint32_t)(maxbits + 1 - (int)w) << 8);
        for (uint32_t k = 0; k < len; k++) T.huf[start[w] + k] = (uint16_t)e;
        start[w] += len;
    }
    huf_log = maxbits;
    return true;
}

// one Huffman-coded stream of `count` symbols; false = corrupt
ORCB_HD bool huf_stream(const uint8_t* p, uint32_t n, const uint16_t* huf, int log, uint8_t* dst, uint32_t count) {
    Back b;
    if (!b.init(p, n)) return false;
    for (uint32_t i = 0; i < count; i++) {
        const uint32_t e = huf[b.peek(log)];
        dst[i] = (uint8_t)e;
        b.bits -= (int)(e >> 8);
    }
    return b.bits == 0;
}

// ---- block headers -----------------------------------------------------------------------------
struct LitHeader {
    uint32_t type;     // 0 raw, 1 RLE, 2 Huffman, 3 Huffman with the previous tree
    uint32_t regen;    // literal bytes of the block
    uint32_t comp;     // bytes of the section after the header (raw: regen, RLE: 1)
    uint32_t streams;  // 1 or 4
    uint32_t hdr;      // header bytes
};

ORCB_HD bool lit_header(const uint8_t* p, uint32_t n, LitHeader& h) {
    if (n < 1) return false;
    const uint32_t b0 = p[0];
    h.type = b0 & 3;
    const uint32_t sf = (b0 >> 2) & 3;
    h.streams = 1;
    if (h.type < 2) {
        if (!(sf & 1)) {
            h.hdr = 1;
            h.regen = b0 >> 3;
        } else if (sf == 1) {
            if (n < 2) return false;
            h.hdr = 2;
            h.regen = (b0 | ((uint32_t)p[1] << 8)) >> 4;
        } else {
            if (n < 3) return false;
            h.hdr = 3;
            h.regen = (b0 | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16)) >> 4;
        }
        h.comp = h.type == 0 ? h.regen : 1;
    } else {
        uint64_t v = 0;
        h.hdr = sf < 2 ? 3 : sf == 2 ? 4 : 5;
        if (n < h.hdr) return false;
        for (uint32_t i = 0; i < h.hdr; i++) v |= (uint64_t)p[i] << (8 * i);
        const int bits = sf < 2 ? 10 : sf == 2 ? 14 : 18;
        h.regen = (uint32_t)(v >> 4) & ((1u << bits) - 1u);
        h.comp = (uint32_t)(v >> (4 + bits)) & ((1u << bits) - 1u);
        h.streams = sf == 0 ? 1 : 4;
    }
    return h.hdr + (uint64_t)h.comp <= n;
}

// Sequences_Section_Header + the three table descriptions; leaves the FSE tables in T.  `consumed`: bytes up to the
// bitstream.  false = corrupt.
ORCB_HD bool seq_header(const uint8_t* p, uint32_t n, Tables& T, FrameState& fs, uint32_t& nseq, uint32_t& consumed) {
    if (n < 1) return false;
    uint32_t q = 0;
    const uint32_t b0 = p[q++];
    if (b0 == 0) {
        nseq = 0;
        consumed = 1;
        return true;
    }
    if (b0 < 128) {
        nseq = b0;
    } else if (b0 < 255) {
        if (q + 1 > n) return false;
        nseq = ((b0 - 128) << 8) + p[q++];
    } else {
        if (q + 2 > n) return false;
        nseq = (uint32_t)p[q] + ((uint32_t)p[q + 1] << 8) + 0x7f00u;
        q += 2;
    }
    if (q + 1 > n) return false;
    const uint32_t modes = p[q++];
    if (modes & 3) return false;
    for (int which = 0; which < 3; which++) {
        const uint32_t mode = (modes >> (6 - 2 * which)) & 3;
        uint32_t* table = which == 0 ? T.ll : which == 1 ? T.of : T.ml;
        int& log = which == 0 ? fs.ll_log : which == 1 ? fs.of_log : fs.ml_log;
        const int max_log = which == 0 ? LL_MAX_LOG : which == 1 ? OF_MAX_LOG : ML_MAX_LOG;
        const int max_syms = which == 0 ? LL_SYMS : which == 1 ? OF_SYMS : ML_SYMS;
        if (mode == 0) {
            int nsym, l;
            fse_predefined(which, T.norm, l, nsym);
            if (!fse_build(T.norm, nsym, l, table, T.sdesc)) return false;
            log = l;
        } else if (mode == 1) {
            if (q + 1 > n) return false;
            const uint32_t sym = p[q++];
            if (sym >= (uint32_t)max_syms) return false;
            table[0] = sym;
            log = 0;
        } else if (mode == 2) {
            int nsym, l;
            uint32_t used;
            if (!fse_read_norm(p + q, n - q, max_log, max_syms, T.norm, l, nsym, used)) return false;
            if (!fse_build(T.norm, nsym, l, table, T.sdesc)) return false;
            log = l;
            q += used;
        } else if (log < 0) {
            return false;  // Repeat_Mode without a previous table
        }
    }
    consumed = q;
    return true;
}

struct SeqReader {
    Back b;
    uint32_t ll_s, of_s, ml_s;
    ORCB_HD bool init(const uint8_t* p, uint32_t n, const FrameState& fs) {
        if (!b.init(p, n)) return false;
        ll_s = b.read(fs.ll_log);
        of_s = b.read(fs.of_log);
        ml_s = b.read(fs.ml_log);
        return b.bits >= 0;
    }
    // one sequence; `last`: no state update follows.  false = corrupt.  The three fields of a sequence are adjacent in
    // the bitstream (offset bits first = most significant), and so are the three state updates: one or two reads each.
    ORCB_HD bool next(const Tables& T, FrameState& fs, bool last, uint32_t& ll, uint32_t& ml, uint32_t& off) {
        const uint32_t le = T.ll[ll_s], oe = T.of[of_s], me = T.ml[ml_s];
        const uint32_t of_code = oe & 0xff, ml_code = me & 0xff, ll_code = le & 0xff;
        if (of_code > 31 || ml_code >= (uint32_t)ML_SYMS || ll_code >= (uint32_t)LL_SYMS) return false;
        const uint32_t mc = T.codes[LL_SYMS + ml_code], lc = T.codes[ll_code];  // fill_codes() ran for this block
        const uint32_t mb = mc >> 24, lb = lc >> 24;
        uint32_t of_x, ml_x, ll_x;
        if (of_code + mb + lb <= 31) {
            const uint32_t v = b.read((int)(of_code + mb + lb));
            ll_x = v & ((1u << lb) - 1u);
            ml_x = (v >> lb) & ((1u << mb) - 1u);
            of_x = v >> (lb + mb);
        } else {
            of_x = b.read((int)of_code);
            ml_x = b.read((int)mb);
            ll_x = b.read((int)lb);
        }
        const uint32_t of_val = (1u << of_code) + of_x;
        ml = (mc & 0xffffffu) + ml_x;
        ll = (lc & 0xffffffu) + ll_x;
        if (of_val > 3) {
            off = of_val - 3;
            fs.rep[2] = fs.rep[1];
            fs.rep[1] = fs.rep[0];
            fs.rep[0] = off;
        } else {
            const uint32_t idx = of_val - 1 + (ll == 0 ? 1u : 0u);
            if (idx == 0) {
                off = fs.rep[0];
            } else {
                off = idx == 1 ? fs.rep[1] : idx == 2 ? fs.rep[2] : fs.rep[0] - 1;
                if (off == 0) off = 1;  // libzstd forces a corrupt zero offset to 1
                if (idx != 1) fs.rep[2] = fs.rep[1];
                fs.rep[1] = fs.rep[0];
                fs.rep[0] = off;
            }
        }
        if (!last) {
            const uint32_t nl = (le >> 8) & 0xff, nm = (me >> 8) & 0xff, no = (oe >> 8) & 0xff;  // <= 9 + 9 + 8 bits
            const uint32_t v = b.read((int)(nl + nm + no));
            of_s = (oe >> 16) + (v & ((1u << no) - 1u));
            ml_s = (me >> 16) + ((v >> no) & ((1u << nm) - 1u));
            ll_s = (le >> 16) + (v >> (no + nm));
        }
        return b.bits >= 0;
    }
};

// ---- frame header ------------------------------------------------------------------------------
struct FrameHeader {
    uint32_t hdr;        // bytes of magic + header
    uint64_t content;    // Frame_Content_Size, or ~0 when absent
    uint32_t checksum;   // 1: four checksum bytes follow the last block
    uint32_t skippable;  // 1: skippable frame of `hdr` bytes in all
};

ORCB_HD bool frame_header(const uint8_t* p, uint32_t n, FrameHeader& h) {
    if (n < 4) return false;
    const uint32_t magic = p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    h.skippable = 0;
    h.content = ~0ull;
    h.checksum = 0;
    if ((magic & 0xfffffff0u) == 0x184d2a50u) {
        if (n < 8) return false;
        const uint32_t len = p[4] | ((uint32_t)p[5] << 8) | ((uint32_t)p[6] << 16) | ((uint32_t)p[7] << 24);
        if (len > n - 8) return false;
        h.skippable = 1;
        h.hdr = 8 + len;
        return true;
    }
    if (magic != 0xfd2fb528u) return false;
    if (n < 5) return false;
    const uint32_t fhd = p[4];
    if (fhd & 0x08) return false;  // reserved bit
    const uint32_t single = (fhd >> 5) & 1, did = fhd & 3, fcs = fhd >> 6;
    uint32_t q = 5;
    if (!single) {
        if (q + 1 > n) return false;
        q++;  // Window_Descriptor: the output buffer is the window
    }
    const uint32_t did_bytes = did == 3 ? 4 : did;
    if (q + did_bytes > n) return false;
    uint32_t dict = 0;
    for (uint32_t i = 0; i < did_bytes; i++) dict |= (uint32_t)p[q + i] << (8 * i);
    if (dict) return false;  // no dictionaries in ORC
    q += did_bytes;
    const uint32_t fcs_bytes = fcs == 0 ? single : fcs == 1 ? 2 : fcs == 2 ? 4 : 8;
    if (q + fcs_bytes > n) return false;
    if (fcs_bytes) {
        uint64_t v = 0;
        for (uint32_t i = 0; i < fcs_bytes; i++) v |= (uint64_t)p[q + i] << (8 * i);
        if (fcs_bytes == 2) v += 256;
        h.content = v;
    }
    q += fcs_bytes;
    h.checksum = (fhd >> 2) & 1;
    h.hdr = q;
    return true;
}

}  // namespace zstd

// ---- LZO1X -------------------------------------------------------------------------------------
namespace lzo {

// One step of an LZO1X stream: a literal run and / or a match.
struct Token {
    uint32_t lit_pos, lit_len;  // literal bytes copied first (input positions)
    uint32_t m_len, m_dist;     // then the match (m_len 0: none)
    uint32_t end;               // 1: end-of-stream marker reached
};

// `state`: literals carried by the previous instruction (0..3), 4 after a literal run; starts at 0.  `first`: the
// very first instruction has its own encoding.  false = corrupt or truncated input.
ORCB_HD bool next(const uint8_t* s, uint32_t n, uint32_t& p, uint32_t& state, bool first, Token& t) {
    t.lit_len = t.m_len = t.m_dist = t.end = 0;
    t.lit_pos = 0;
    if (p >= n) return false;
    uint32_t b = s[p++];
    if (first && b > 17) {
        // 18..21: 1..4 literals ... (byte - 17) literals; a run of 4 or more leaves state 4
        const uint32_t len = b - 17;
        if (len > n - p) return false;
        t.lit_pos = p;
        t.lit_len = len;
        p += len;
        state = len < 4 ? len : 4;
        return true;
    }
    if (b < 16) {
        if (state == 0) {
            // literal run of 3 + L bytes
            uint32_t len = b;
            if (len == 0) {
                for (;;) {
                    if (p >= n) return false;
                    const uint32_t z = s[p++];
                    if (z) {
                        len += 15 + z;
                        break;
                    }
                    len += 255;
                    if (len > 0x1000000u) return false;
                }
            }
            len += 3;
            if (len > n - p) return false;
            t.lit_pos = p;
            t.lit_len = len;
            p += len;
            state = 4;
            return true;
        }
        if (p >= n) return false;
        const uint32_t h = s[p++];
        if (state == 4) {
            t.m_dist = (h << 2) + (b >> 2) + 2049;
            t.m_len = 3;
        } else {
            t.m_dist = (h << 2) + (b >> 2) + 1;
            t.m_len = 2;
        }
        state = b & 3;
    } else if (b >= 64) {
        if (p >= n) return false;
        const uint32_t h = s[p++];
        t.m_dist = (h << 3) + ((b >> 2) & 7) + 1;
        t.m_len = (b >> 5) + 1;
        state = b & 3;
    } else {
        const bool m3 = b >= 32;
        uint32_t len = b & (m3 ? 31u : 7u);
        if (len == 0) {
            for (;;) {
                if (p >= n) return false;
                const uint32_t z = s[p++];
                if (z) {
                    len += (m3 ? 31u : 7u) + z;
                    break;
                }
                len += 255;
                if (len > 0x1000000u) return false;
            }
        }
        len += 2;
        if (n - p < 2) return false;
        const uint32_t v = s[p] | ((uint32_t)s[p + 1] << 8);
        p += 2;
        if (m3) {
            t.m_dist = (v >> 2) + 1;
        } else {
            t.m_dist = 16384 + ((b & 8) << 11) + (v >> 2);
            if (t.m_dist == 16384) {
                // end of stream: the length field says 3 (byte 0x11) in every stream a compressor writes
                t.end = 1;
                t.m_len = 0;
                t.m_dist = 0;
                return len == 3;
            }
        }
        t.m_len = len;
        state = v & 3;
    }
    // the 0..3 literals that ride on a match instruction
    if (state) {
        if (state > n - p) return false;
        t.lit_pos = p;  // NOTE: these follow the match; the caller copies match first, then these
        t.lit_len = state;
        p += state;
    }
    return true;
}

}  // namespace lzo
}  // namespace orcb
