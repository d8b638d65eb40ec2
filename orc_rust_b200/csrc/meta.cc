#include "meta.h"

#include <algorithm>
#include <cstring>
#include <memory>

#include "pb.h"
#include "job_internal.h"
#include "zstd_dec.h"

namespace orcb {

// ---------------------------------------------------------------------------------------------
// Host block decoders for metadata sections (file footer, stripe footers, ROW_INDEX streams).
// Data streams never pass through these: they are decompressed by the CUDA kernels.
// ---------------------------------------------------------------------------------------------
static void host_snappy_block(const uint8_t* s, size_t n, std::vector<uint8_t>& out) {
    size_t ip = 0;
    uint64_t want = 0;
    for (int shift = 0;; shift += 7) {
        if (ip >= n || shift > 35) fail(ORCB_BUILD_SNAPPY_DECODER, "bad snappy preamble");
        uint8_t b = s[ip++];
        want |= (uint64_t)(b & 0x7f) << shift;
        if (b < 0x80) break;
    }
    size_t base = out.size();
    // no element makes more than 64 bytes out of 3: a larger preamble cannot be met, and must not size an allocation
    if (want > 32 * (uint64_t)n + 64) fail(ORCB_BUILD_SNAPPY_DECODER, "snappy length mismatch");
    out.reserve(base + want);
    while (ip < n) {
        const uint8_t tag = s[ip++];
        size_t len, dist = 0;
        switch (tag & 3) {
            case 0: {
                len = tag >> 2;
                if (len >= 60) {
                    size_t extra = len - 59;
                    if (ip + extra > n) fail(ORCB_BUILD_SNAPPY_DECODER, "truncated literal length");
                    len = 0;
                    for (size_t k = 0; k < extra; k++) len |= (size_t)s[ip + k] << (8 * k);
                    ip += extra;
                }
                len++;
                if (ip + len > n) fail(ORCB_BUILD_SNAPPY_DECODER, "truncated literal");
                out.insert(out.end(), s + ip, s + ip + len);
                ip += len;
                continue;
            }
            case 1:
                if (ip + 1 > n) fail(ORCB_BUILD_SNAPPY_DECODER, "truncated copy");
                len = 4 + ((tag >> 2) & 7);
                dist = ((size_t)(tag >> 5) << 8) | s[ip];
                ip += 1;
                break;
            case 2:
                if (ip + 2 > n) fail(ORCB_BUILD_SNAPPY_DECODER, "truncated copy");
                len = 1 + (tag >> 2);
                dist = s[ip] | ((size_t)s[ip + 1] << 8);
                ip += 2;
                break;
            default:
                if (ip + 4 > n) fail(ORCB_BUILD_SNAPPY_DECODER, "truncated copy");
                len = 1 + (tag >> 2);
                dist = s[ip] | ((size_t)s[ip + 1] << 8) | ((size_t)s[ip + 2] << 16) | ((size_t)s[ip + 3] << 24);
                ip += 4;
                break;
        }
        size_t produced = out.size() - base;
        if (dist == 0 || dist > produced) fail(ORCB_BUILD_SNAPPY_DECODER, "copy offset out of range");
        for (size_t k = 0; k < len; k++) out.push_back(out[out.size() - dist]);
    }
    if (out.size() - base != want) fail(ORCB_BUILD_SNAPPY_DECODER, "snappy length mismatch");
}

static void host_lz4_block(const uint8_t* s, size_t n, size_t max_out, std::vector<uint8_t>& out) {
    size_t ip = 0;
    const size_t base = out.size();
    auto ext = [&](size_t v) {
        if (v == 15) {
            uint8_t b;
            do {
                if (ip >= n) fail(ORCB_BUILD_LZ4_DECODER, "truncated length");
                b = s[ip++];
                v += b;
            } while (b == 255);
        }
        return v;
    };
    while (ip < n) {
        const uint8_t token = s[ip++];
        size_t lit = ext(token >> 4);
        if (ip + lit > n || out.size() - base + lit > max_out) fail(ORCB_BUILD_LZ4_DECODER, "literal overrun");
        out.insert(out.end(), s + ip, s + ip + lit);
        ip += lit;
        if (ip >= n) break;
        if (ip + 2 > n) fail(ORCB_BUILD_LZ4_DECODER, "truncated offset");
        size_t dist = s[ip] | ((size_t)s[ip + 1] << 8);
        ip += 2;
        size_t mlen = ext(token & 15) + 4;
        if (dist == 0 || dist > out.size() - base || out.size() - base + mlen > max_out)
            fail(ORCB_BUILD_LZ4_DECODER, "match out of range");
        for (size_t k = 0; k < mlen; k++) out.push_back(out[out.size() - dist]);
    }
}

// Raw DEFLATE for metadata sections of Zlib files (footer, stripe footers, row indexes); data streams are inflated on the
// device (k_decompress.cu).  Bit-by-bit canonical decode: these sections are a few KiB.
namespace {
// The input ended inside the stream.  flate2's read decoder (the reference, src/compression.rs:142-149) does not report
// that: read_to_end returns what was decoded up to there, and so does this decoder (whatever parses the section next
// usually fails on the short bytes, with its own error).
struct InflateEof {};
struct HostBits {
    const uint8_t* s;
    size_t n, pos = 0;
    uint32_t buf = 0;
    int cnt = 0;
    uint32_t bits(int k) {
        while (cnt < k) {
            if (pos >= n) throw InflateEof{};
            buf |= (uint32_t)s[pos++] << cnt;
            cnt += 8;
        }
        const uint32_t v = buf & ((1u << k) - 1u);
        buf >>= k;
        cnt -= k;
        return v;
    }
};
struct HostHuff {
    uint16_t count[16] = {0}, symbol[288] = {0};
    void build(const uint8_t* len, int n) {
        for (int i = 0; i < 16; i++) count[i] = 0;
        for (int i = 0; i < n; i++) count[len[i]]++;
        int left = 1;
        for (int l = 1; l <= 15; l++) {
            left = (left << 1) - count[l];
            if (left < 0) fail(ORCB_IO_ERROR, "corrupt deflate stream");
        }
        uint16_t offs[16];
        offs[1] = 0;
        for (int l = 1; l < 15; l++) offs[l + 1] = offs[l] + count[l];
        for (int i = 0; i < n; i++)
            if (len[i]) symbol[offs[len[i]]++] = (uint16_t)i;
    }
    int decode(HostBits& b) const {
        int code = 0, first = 0, index = 0;
        for (int l = 1; l <= 15; l++) {
            code |= (int)b.bits(1);
            const int c = count[l];
            if (code - c < first) return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        fail(ORCB_IO_ERROR, "corrupt deflate stream");
    }
};
}  // namespace

static void host_inflate_blocks(const uint8_t* s, size_t n, std::vector<uint8_t>& out);
static void host_inflate(const uint8_t* s, size_t n, std::vector<uint8_t>& out) {
    try {
        host_inflate_blocks(s, n, out);
    } catch (const InflateEof&) {
    }
}

static void host_inflate_blocks(const uint8_t* s, size_t n, std::vector<uint8_t>& out) {
    static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t ORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    HostBits b{s, n};
    const size_t base = out.size();
    int last;
    do {
        last = (int)b.bits(1);
        const uint32_t type = b.bits(2);
        if (type == 0) {
            b.buf = 0;
            b.cnt = 0;
            if (b.pos + 4 > n) throw InflateEof{};
            const uint32_t len = s[b.pos] | (s[b.pos + 1] << 8), nlen = s[b.pos + 2] | (s[b.pos + 3] << 8);
            b.pos += 4;
            if ((len ^ nlen) != 0xffffu) fail(ORCB_IO_ERROR, "corrupt deflate stream");
            if (len > n - b.pos) {  // a stored block cut short: the bytes that are there, then the end of input
                out.insert(out.end(), s + b.pos, s + n);
                throw InflateEof{};
            }
            out.insert(out.end(), s + b.pos, s + b.pos + len);
            b.pos += len;
            continue;
        }
        if (type == 3) fail(ORCB_IO_ERROR, "corrupt deflate stream");
        HostHuff lc, dc;
        uint8_t lens[320];
        if (type == 1) {
            for (int i = 0; i < 288; i++) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
            lc.build(lens, 288);
            for (int i = 0; i < 30; i++) lens[i] = 5;
            dc.build(lens, 30);
        } else {
            const int nl = (int)b.bits(5) + 257, nd = (int)b.bits(5) + 1, nc = (int)b.bits(4) + 4;
            if (nl > 286 || nd > 30) fail(ORCB_IO_ERROR, "corrupt deflate stream");
            uint8_t cl[19] = {0};
            for (int i = 0; i < nc; i++) cl[ORDER[i]] = (uint8_t)b.bits(3);
            HostHuff cc;
            cc.build(cl, 19);
            int i = 0;
            while (i < nl + nd) {
                const int sym = cc.decode(b);
                if (sym < 16) {
                    lens[i++] = (uint8_t)sym;
                } else {
                    int rep, val = 0;
                    if (sym == 16) {
                        if (i == 0) fail(ORCB_IO_ERROR, "corrupt deflate stream");
                        val = lens[i - 1];
                        rep = 3 + (int)b.bits(2);
                    } else if (sym == 17) rep = 3 + (int)b.bits(3);
                    else rep = 11 + (int)b.bits(7);
                    if (i + rep > nl + nd) fail(ORCB_IO_ERROR, "corrupt deflate stream");
                    while (rep--) lens[i++] = (uint8_t)val;
                }
            }
            if (lens[256] == 0) fail(ORCB_IO_ERROR, "corrupt deflate stream");
            lc.build(lens, nl);
            dc.build(lens + nl, nd);
        }
        for (;;) {
            const int sym = lc.decode(b);
            if (sym < 256) out.push_back((uint8_t)sym);
            else if (sym == 256) break;
            else {
                if (sym > 285) fail(ORCB_IO_ERROR, "corrupt deflate stream");
                const size_t len = LBASE[sym - 257] + b.bits(LEXT[sym - 257]);
                const int ds = dc.decode(b);
                if (ds > 29) fail(ORCB_IO_ERROR, "corrupt deflate stream");
                const size_t dist = DBASE[ds] + b.bits(DEXT[ds]);
                if (dist > out.size() - base) fail(ORCB_IO_ERROR, "corrupt deflate stream");
                for (size_t k = 0; k < len; k++) out.push_back(out[out.size() - dist]);
            }
        }
    } while (!last);
}

// Zstandard frames and LZO1X streams of metadata sections; the parsing is shared with the device decoder (zstd_dec.h).
static void host_zstd_frames(const uint8_t* s, size_t n, std::vector<uint8_t>& out) {
    using namespace zstd;
    auto bad = [] { fail(ORCB_IO_ERROR, "corrupt zstd frame"); };
    std::unique_ptr<Tables> T(new Tables);
    size_t p = 0;
    while (p < n) {
        FrameHeader fh;
        if (!frame_header(s + p, (uint32_t)std::min<size_t>(n - p, 0xffffffffu), fh)) {
            bad();
        }
        p += fh.hdr;
        if (fh.skippable) continue;
        FrameState fs;
        fs.reset();
        const size_t base = out.size();
        for (;;) {
            if (p + 3 > n) bad();
            const uint32_t bh = s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16);
            p += 3;
            const uint32_t last = bh & 1, type = (bh >> 1) & 3, size = bh >> 3;
            if (type == 3) bad();
            if (type == 0) {
                if (size > n - p) bad();
                out.insert(out.end(), s + p, s + p + size);
                p += size;
            } else if (type == 1) {
                if (p >= n) bad();
                out.insert(out.end(), size, s[p]);
                p += 1;
            } else {
                if (size > n - p || size > (128u << 10)) bad();
                const uint8_t* b = s + p;
                LitHeader lh;
                if (!lit_header(b, size, lh)) bad();
                std::vector<uint8_t> lits(lh.regen);
                const uint8_t* lp = b + lh.hdr;
                if (lh.type == 0) {
                    std::copy(lp, lp + lh.regen, lits.begin());
                } else if (lh.type == 1) {
                    std::fill(lits.begin(), lits.end(), lp[0]);
                } else {
                    uint32_t used = 0;
                    if (lh.type == 2) {
                        if (!huf_read(lp, lh.comp, *T, fs.huf_log, used)) bad();
                    } else if (fs.huf_log < 0) {
                        bad();
                    }
                    const uint8_t* sp = lp + used;
                    const uint32_t sn = lh.comp - used;
                    if (lh.streams == 1) {
                        if (!huf_stream(sp, sn, T->huf, fs.huf_log, lits.data(), lh.regen)) bad();
                    } else {
                        if (sn < 10 || lh.regen < 6) bad();  // libzstd: jump table + one byte per stream; at least 6 literals
                        const uint32_t l1 = sp[0] | (sp[1] << 8), l2 = sp[2] | (sp[3] << 8), l3 = sp[4] | (sp[5] << 8);
                        if ((uint64_t)6 + l1 + l2 + l3 > sn) bad();
                        const uint32_t l4 = sn - 6 - l1 - l2 - l3, q = (lh.regen + 3) / 4;
                        if (3 * (uint64_t)q > lh.regen) bad();
                        const uint32_t lens[4] = {l1, l2, l3, l4};
                        const uint8_t* at = sp + 6;
                        for (int k = 0; k < 4; k++) {
                            if (!huf_stream(at, lens[k], T->huf, fs.huf_log, lits.data() + k * q, k < 3 ? q : lh.regen - 3 * q)) bad();
                            at += lens[k];
                        }
                    }
                }
                const uint8_t* sq = b + lh.hdr + lh.comp;
                const uint32_t sqn = size - lh.hdr - lh.comp;
                uint32_t nseq = 0, used = 0;
                if (!seq_header(sq, sqn, *T, fs, nseq, used)) bad();
                size_t lit_at = 0;
                if (nseq) {
                    fill_codes(*T, 0, 1);
                    SeqReader r;
                    if (!r.init(sq + used, sqn - used, fs)) bad();
                    for (uint32_t i = 0; i < nseq; i++) {
                        uint32_t ll, ml, off;
                        if (!r.next(*T, fs, i + 1 == nseq, ll, ml, off)) bad();
                        if (ll > lits.size() - lit_at) bad();
                        out.insert(out.end(), lits.begin() + lit_at, lits.begin() + lit_at + ll);
                        lit_at += ll;
                        if (off > out.size() - base) bad();
                        for (uint32_t k = 0; k < ml; k++) out.push_back(out[out.size() - off]);
                    }
                    if (r.b.bits != 0) bad();
                } else if (used != sqn) {
                    bad();
                }
                out.insert(out.end(), lits.begin() + lit_at, lits.end());
                p += size;
            }
            if (last) break;
        }
        if (fh.checksum) {
            if (p + 4 > n) bad();
            p += 4;  // content checksum: not verified
        }
        if (fh.content != ~0ull && out.size() - base != fh.content) bad();
    }
}

static void host_lzo_block(const uint8_t* s, size_t n, std::vector<uint8_t>& out) {
    const size_t base = out.size();
    uint32_t p = 0, state = 0;
    bool first = true;
    for (;;) {
        lzo::Token t;
        if (!lzo::next(s, (uint32_t)n, p, state, first, t)) fail(ORCB_BUILD_LZO_DECODER, "corrupt LZO stream");
        first = false;
        if (t.end) break;
        if (t.m_len) {
            if (t.m_dist > out.size() - base) fail(ORCB_BUILD_LZO_DECODER, "LZO match before the start of the block");
            for (uint32_t k = 0; k < t.m_len; k++) out.push_back(out[out.size() - t.m_dist]);
        }
        out.insert(out.end(), s + t.lit_pos, s + t.lit_pos + t.lit_len);
    }
    if (p != n) fail(ORCB_BUILD_LZO_DECODER, "bytes after the end of the LZO stream");
}

std::vector<uint8_t> host_decompress_section(int compression, uint64_t block_size, const uint8_t* in, size_t len) {
    std::vector<uint8_t> out;
    if (compression == C_NONE) {
        out.assign(in, in + len);
        return out;
    }
    size_t p = 0;
    while (p < len) {
        if (p + 3 > len) fail(ORCB_OUT_OF_SPEC, "truncated compression chunk header");
        uint32_t h = in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16);
        p += 3;
        uint32_t clen = h >> 1;
        if (p + clen > len) fail(ORCB_OUT_OF_SPEC, "compression chunk exceeds section");
        if (h & 1) {
            out.insert(out.end(), in + p, in + p + clen);
        } else if (compression == C_SNAPPY) {
            host_snappy_block(in + p, clen, out);
        } else if (compression == C_LZ4) {
            host_lz4_block(in + p, clen, block_size, out);
        } else if (compression == C_ZLIB) {
            host_inflate(in + p, clen, out);
        } else if (compression == C_ZSTD) {
            host_zstd_frames(in + p, clen, out);
        } else if (compression == C_LZO) {
            host_lzo_block(in + p, clen, out);
        } else {
            fail(ORCB_DECODE_PROTO, "unknown compression kind");
        }
        p += clen;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// Files behind read callbacks
// ---------------------------------------------------------------------------------------------
RangeBuf::~RangeBuf() {
    if (pinned_cap) pinned_put(p, pinned_cap);
}

std::shared_ptr<RangeBuf> FileMeta::load_range(uint64_t off, uint64_t n) const {
    if (!source) return nullptr;
    if (off > len || n > len - off) fail(ORCB_IO_ERROR, "read beyond the end of the file");  // (before off + n is formed)
    {
        std::lock_guard<std::mutex> lock(source->mu);
        auto& v = source->live;
        for (size_t i = 0; i < v.size();) {
            std::shared_ptr<RangeBuf> r = v[i].lock();
            if (!r) {
                v[i] = v.back();
                v.pop_back();
                continue;
            }
            if (r->off <= off && off + n <= r->off + r->len) return r;
            i++;
        }
    }
    auto r = std::make_shared<RangeBuf>();
    r->off = off;
    r->len = n;
    int dev = 0;
    if (n && cudaGetDeviceCount(&dev) == cudaSuccess && dev > 0) {
        try {
            r->p = (uint8_t*)pinned_get((size_t)n + 256, &r->pinned_cap);
        } catch (const OrcException&) {
            r->p = nullptr;
            r->pinned_cap = 0;
        }
    } else {
        cudaGetLastError();
    }
    if (!r->p) {
        r->heap.resize((size_t)n + 256);
        r->p = r->heap.data();
    }
    if (n) {
        const int rc = source->read_at(source->ctx, off, n, r->p);
        if (rc != 0) fail(ORCB_IO_ERROR, "read callback failed with status " + std::to_string(rc));
    }
    std::lock_guard<std::mutex> lock(source->mu);
    source->reads++;
    source->bytes_read += n;
    source->live.push_back(r);
    return r;
}

std::shared_ptr<RangeBuf> FileMeta::load_stripe(uint32_t stripe) const {
    if (!source) return nullptr;
    const StripeInfo& si = stripes.at(stripe);
    return load_range(si.offset, si.index_length + si.data_length + si.footer_length);
}

const uint8_t* FileMeta::base_for(uint64_t off, uint64_t need) const {
    if (!source) return data;
    // The loaded range must hold the whole extent.  Ranges of consecutive stripes touch (the data area of a stripe
    // without an index area starts exactly where the range of the stripe before it ends) and, in a damaged file, stripes
    // may overlap: a range that merely ends at `off`, or ends inside the extent, is not the one to read from.
    auto holds = [&](const RangeBuf& r) { return r.off <= off && off - r.off <= r.len && need <= r.len - (off - r.off); };
    if (tail && holds(*tail)) return tail->p - tail->off;
    std::lock_guard<std::mutex> lock(source->mu);
    for (auto& w : source->live) {
        std::shared_ptr<RangeBuf> r = w.lock();
        if (r && holds(*r)) return r->p - r->off;
    }
    fail(ORCB_UNEXPECTED, "file bytes at offset " + std::to_string(off) + " are not loaded");
}

// ---------------------------------------------------------------------------------------------
// File tail (src/reader/metadata.rs:180-263)
// ---------------------------------------------------------------------------------------------
static std::string pb_string(const PbField& f);  // UTF-8 checked, as prost checks `string` fields

static OrcType parse_type(const uint8_t* p, size_t n) {
    OrcType t;
    PbCursor c(p, n);
    PbField f;
    while (c.next(f)) {
        switch (f.number) {
            case 1: t.kind = (int)f.value; break;
            case 2: {
                std::vector<uint64_t> v;
                PbCursor::packed_u64(f, v);
                for (auto x : v) t.subtypes.push_back((uint32_t)x);
                break;
            }
            case 3: t.field_names.push_back(pb_string(f)); break;
            case 4: t.max_length = (uint32_t)f.value; break;
            case 5: t.precision = (uint32_t)f.value; break;
            case 6: t.scale = (uint32_t)f.value; break;
            default: break;
        }
    }
    return t;
}

void parse_file_tail(FileMeta& fm) {
    const size_t n = fm.len;
    if (n == 0) fail(ORCB_EMPTY_FILE, "Empty file");
    if (fm.source) {
        // read_metadata (src/reader/metadata.rs:180-236): the last 16 KiB first; once the postscript says how long
        // footer and metadata are, the whole tail if that was not enough
        fm.tail = fm.load_range(n - std::min<size_t>(n, 16384), std::min<size_t>(n, 16384));
        const uint8_t* t = fm.tail->p - fm.tail->off;
        const size_t ps_len = t[n - 1];
        if (n - 1 >= ps_len && ps_len <= fm.tail->len - 1) {
            uint64_t fl = 0, ml = 0;
            PbCursor c(t + n - 1 - ps_len, ps_len);
            PbField f;
            while (c.next(f)) {
                if (f.number == 1) fl = f.value;
                else if (f.number == 5) ml = f.value;
            }
            const uint64_t room = n - 1 - ps_len;
            if (fl <= room && ml <= room - fl && fl + ml + ps_len + 1 > fm.tail->len) {
                const uint64_t need = fl + ml + ps_len + 1;
                fm.tail = fm.load_range(n - need, need);
            }
        }
    }
    const uint8_t* d = fm.base_for(n - 1);
    size_t ps_len = d[n - 1];
    if (n - 1 < ps_len) fail(ORCB_OUT_OF_SPEC, "File too small for given postscript length");
    bool have_footer = false, have_meta = false;
    uint64_t footer_len = 0, meta_len = 0;
    {
        PbCursor c(d + n - 1 - ps_len, ps_len);
        PbField f;
        while (c.next(f)) {
            switch (f.number) {
                case 1: footer_len = f.value; have_footer = true; break;
                case 2: fm.compression = (int)f.value; break;
                case 3: fm.block_size = f.value; break;
                case 4: {  // repeated uint32 version, packed or not
                    std::vector<uint64_t> v;
                    PbCursor::packed_u64(f, v);
                    for (uint64_t x : v) fm.format_version += (fm.format_version.empty() ? "" : ".") + std::to_string((uint32_t)x);
                    break;
                }
                case 5: meta_len = f.value; have_meta = true; break;
                default: break;
            }
        }
    }
    if (!have_footer) fail(ORCB_OUT_OF_SPEC, "Footer length is empty");
    if (!have_meta) fail(ORCB_OUT_OF_SPEC, "Metadata length is empty");
    if (fm.compression < 0 || fm.compression > C_ZSTD) fail(ORCB_DECODE_PROTO, "unknown compression kind");
    // subtraction-style checks: the lengths come from the file and their sum may wrap
    if (footer_len > n - 1 - ps_len || meta_len > n - 1 - ps_len - footer_len) fail(ORCB_OUT_OF_SPEC, "footer exceeds file");
    size_t fend = n - 1 - ps_len;
    std::vector<uint8_t> footer = host_decompress_section(fm.compression, fm.block_size, d + fend - footer_len, footer_len);
    PbCursor c(footer.data(), footer.size());
    PbField f;
    while (c.next(f)) {
        switch (f.number) {
            case 3: {
                StripeInfo s;
                PbCursor sc(f.data, f.len);
                PbField g;
                while (sc.next(g)) {
                    if (g.number == 1) s.offset = g.value;
                    else if (g.number == 2) s.index_length = g.value;
                    else if (g.number == 3) s.data_length = g.value;
                    else if (g.number == 4) s.footer_length = g.value;
                    else if (g.number == 5) s.rows = g.value;
                }
                fm.stripes.push_back(s);
                break;
            }
            case 4: fm.types.push_back(parse_type(f.data, f.len)); break;
            case 5: {
                std::string k, v;
                PbCursor mc(f.data, f.len);
                PbField g;
                while (mc.next(g)) {
                    if (g.number == 1) k = pb_string(g);  // string name = 1; bytes value = 2
                    else if (g.number == 2) v.assign((const char*)g.data, g.len);
                }
                fm.user_metadata.emplace_back(k, v);
                break;
            }
            case 6: fm.num_rows = f.value; break;
            case 8: fm.row_index_stride = (int64_t)f.value; break;
            default: break;
        }
    }
    if (fm.types.empty()) fail(ORCB_NO_TYPES, "No types found");
    const OrcType& root = fm.types[0];
    if (root.kind != T_STRUCT) fail(ORCB_UNEXPECTED, "non-struct root type is not supported");
    if (root.subtypes.size() != root.field_names.size())
        fail(ORCB_UNEXPECTED, "Struct type for column index 0 must have matching lengths for subtypes and field names lists");
    for (size_t i = 0; i < root.subtypes.size(); i++) {
        if (root.subtypes[i] >= fm.types.size()) fail(ORCB_UNEXPECTED, "Column index out of bounds");
        fm.root_columns.emplace_back(root.field_names[i], root.subtypes[i]);
    }
    for (auto& s : fm.stripes) {
        uint64_t room = n;
        for (uint64_t part : {s.offset, s.index_length, s.data_length, s.footer_length}) {
            if (part > room) fail(ORCB_IO_ERROR, "stripe exceeds file length");
            room -= part;
        }
    }
}

// src/stripe.rs:128-182 (footer decode + running stream offsets)
StripeFooter FileMeta::read_stripe_footer(uint32_t stripe) const {
    const StripeInfo& si = stripes.at(stripe);
    uint64_t off = si.offset + si.index_length + si.data_length;
    std::vector<uint8_t> raw = host_decompress_section(compression, block_size, base_for(off, si.footer_length) + off, si.footer_length);
    StripeFooter sf;
    PbCursor c(raw.data(), raw.size());
    PbField f;
    uint64_t pos = si.offset;
    while (c.next(f)) {
        if (f.number == 1) {
            StreamInfo s;
            PbCursor sc(f.data, f.len);
            PbField g;
            while (sc.next(g)) {
                if (g.number == 1) s.kind = (int)g.value;
                else if (g.number == 2) s.column = (uint32_t)g.value;
                else if (g.number == 3) s.length = g.value;
            }
            s.offset = pos;
            if (s.length > len - pos) fail(ORCB_IO_ERROR, "stream exceeds file length");
            pos += s.length;
            sf.streams.push_back(s);
        } else if (f.number == 2) {
            ColumnEncoding e;
            PbCursor ec(f.data, f.len);
            PbField g;
            while (ec.next(g)) {
                if (g.number == 1) e.kind = (int)g.value;
                else if (g.number == 2) e.dict_size = (uint32_t)g.value;
            }
            sf.encodings.push_back(e);
        } else if (f.number == 3) {
            sf.has_tz = true;
            sf.tz = pb_string(f);
        }
    }
    return sf;
}

// src/row_index.rs:204-289 — only the positions (the reference parses but never uses them, :35-51)
std::vector<std::vector<uint64_t>> FileMeta::read_row_index(const StripeInfo& si, const StripeFooter& sf,
                                                            uint32_t column) const {
    (void)si;
    std::vector<std::vector<uint64_t>> out;
    const StreamInfo* st = sf.find(column, S_ROW_INDEX);
    if (!st || st->length == 0) return out;
    std::vector<uint8_t> raw = host_decompress_section(compression, block_size, base_for(st->offset, st->length) + st->offset, st->length);
    PbCursor c(raw.data(), raw.size());
    PbField f;
    while (c.next(f)) {
        if (f.number != 1) continue;
        std::vector<uint64_t> pos;
        pos.reserve(8);  // (a row-index entry has 1..7 positions: one allocation instead of four)
        PbCursor ec(f.data, f.len);
        PbField g;
        while (ec.next(g)) {
            if (g.number == 1) PbCursor::packed_u64(g, pos);
        }
        out.push_back(std::move(pos));
    }
    return out;
}

// prost checks `string` fields while decoding; a bad one is a DecodeProto error
static std::string pb_string(const PbField& f) {
    if (f.wire != 2) fail(ORCB_DECODE_PROTO, "invalid wire type for a string field");
    const uint8_t* p = f.data;
    size_t i = 0, n = f.len;
    while (i < n) {
        const uint8_t b = p[i];
        size_t need = 0;
        uint32_t cp = 0;
        if (b < 0x80) { i++; continue; }
        if (b >= 0xc2 && b <= 0xdf) { need = 1; cp = b & 0x1f; }
        else if (b >= 0xe0 && b <= 0xef) { need = 2; cp = b & 0x0f; }
        else if (b >= 0xf0 && b <= 0xf4) { need = 3; cp = b & 0x07; }
        else fail(ORCB_DECODE_PROTO, "invalid string value: data is not UTF-8 encoded");
        for (size_t k = 1; k <= need; k++) {
            if (i + k >= n || (p[i + k] & 0xc0) != 0x80) fail(ORCB_DECODE_PROTO, "invalid string value: data is not UTF-8 encoded");
            cp = (cp << 6) | (p[i + k] & 0x3f);
        }
        if ((need == 2 && (cp < 0x800 || (cp >= 0xd800 && cp <= 0xdfff))) || (need == 3 && (cp < 0x10000 || cp > 0x10ffff)))
            fail(ORCB_DECODE_PROTO, "invalid string value: data is not UTF-8 encoded");
        i += need + 1;
    }
    return std::string((const char*)f.data, f.len);
}
static int64_t unzigzag(uint64_t v) { return (int64_t)(v >> 1) ^ -(int64_t)(v & 1); }
static double pb_double(const PbField& f) {
    if (f.wire != 1) fail(ORCB_DECODE_PROTO, "invalid wire type for a double field");
    double d;
    memcpy(&d, &f.value, 8);
    return d;
}
static void want_wire(const PbField& f, uint32_t wire) {
    if (f.wire != wire) fail(ORCB_DECODE_PROTO, "invalid wire type in ColumnStatistics");
}

// TryFrom<&proto::ColumnStatistics> (src/statistics.rs:77-143)
static ColumnStats parse_column_stats(const uint8_t* p, size_t n) {
    ColumnStats st;
    bool has[13] = {};
    bool has_smin = false, has_smax = false;
    std::string lower, upper;
    std::vector<uint64_t> counts;
    int64_t ts_min_utc = 0, ts_max_utc = 0, date_min = 0, date_max = 0, int_min = 0, int_max = 0;
    PbCursor c(p, n);
    PbField f, g;
    while (c.next(f)) {
        if (f.number >= 2 && f.number <= 9) want_wire(f, 2);
        if (f.number < 13) has[f.number] = true;
        PbCursor sc(f.data, f.len);
        switch (f.number) {
            case 1: want_wire(f, 0); st.number_of_values = f.value; break;
            case 10: want_wire(f, 0); st.has_null = f.value != 0; break;
            case 2:
                while (sc.next(g)) {
                    if (g.number >= 1 && g.number <= 3) want_wire(g, 0);
                    if (g.number == 1) int_min = unzigzag(g.value);
                    else if (g.number == 2) int_max = unzigzag(g.value);
                }
                break;
            case 3:
                while (sc.next(g)) {
                    if (g.number == 1) st.dmin = pb_double(g);
                    else if (g.number == 2) st.dmax = pb_double(g);
                    else if (g.number == 3) (void)pb_double(g);
                }
                break;
            case 4:
                while (sc.next(g)) {
                    if (g.number == 1) { st.smin = pb_string(g); has_smin = true; }
                    else if (g.number == 2) { st.smax = pb_string(g); has_smax = true; }
                    else if (g.number == 3) want_wire(g, 0);
                    else if (g.number == 4) lower = pb_string(g);
                    else if (g.number == 5) upper = pb_string(g);
                }
                break;
            case 5:
                while (sc.next(g))
                    if (g.number == 1) PbCursor::packed_u64(g, counts);
                break;
            case 6:
                while (sc.next(g)) {
                    if (g.number == 1) st.smin = pb_string(g);
                    else if (g.number == 2) st.smax = pb_string(g);
                    else if (g.number == 3) (void)pb_string(g);
                }
                break;
            case 7:
                while (sc.next(g)) {
                    if (g.number == 1 || g.number == 2) want_wire(g, 0);
                    if (g.number == 1) date_min = (int32_t)unzigzag(g.value & 0xffffffffull);
                    else if (g.number == 2) date_max = (int32_t)unzigzag(g.value & 0xffffffffull);
                }
                break;
            case 9:
                while (sc.next(g)) {
                    if (g.number >= 1 && g.number <= 6) want_wire(g, 0);
                    if (g.number == 3) ts_min_utc = unzigzag(g.value);
                    else if (g.number == 4) ts_max_utc = unzigzag(g.value);
                }
                break;
            default: break;
        }
    }
    if (st.number_of_values == 0) return st;
    if (has[2]) { st.kind = ST_INTEGER; st.imin = int_min; st.imax = int_max; }
    else if (has[3]) st.kind = ST_DOUBLE;
    else if (has[4]) {
        st.kind = ST_STRING;
        st.exact_min = has_smin;
        st.exact_max = has_smax;
        if (!has_smin) st.smin = lower;
        if (!has_smax) st.smax = upper;
    } else if (has[5]) {
        st.kind = ST_BUCKET;
        if (counts.empty()) throw ReferencePanic("index out of bounds: the len is 0 but the index is 0 (bucket statistics without a count)");
        st.true_count = counts[0];
    } else if (has[6]) st.kind = ST_DECIMAL;
    else if (has[7]) { st.kind = ST_DATE; st.imin = date_min; st.imax = date_max; }
    else if (has[8]) st.kind = ST_BINARY;
    else if (has[9]) { st.kind = ST_TIMESTAMP; st.imin = ts_min_utc; st.imax = ts_max_utc; }
    else if (has[12]) st.kind = ST_COLLECTION;
    return st;
}

// parse_stripe_row_indexes + parse_bloom_filters for one column (src/row_index.rs:204-331)
std::vector<RowGroupEntry> FileMeta::read_row_group_entries(const StripeFooter& sf, uint32_t column, bool* present) const {
    std::vector<RowGroupEntry> out;
    const StreamInfo* st = sf.find(column, S_ROW_INDEX);
    *present = st != nullptr;
    if (!st) return out;
    {
        std::vector<uint8_t> raw = host_decompress_section(compression, block_size, base_for(st->offset, st->length) + st->offset, st->length);
        PbCursor c(raw.data(), raw.size());
        PbField f, g;
        while (c.next(f)) {
            if (f.number != 1) continue;
            want_wire(f, 2);
            RowGroupEntry e;
            PbCursor ec(f.data, f.len);
            while (ec.next(g)) {
                if (g.number == 2) {
                    want_wire(g, 2);
                    e.has_stats = true;
                    e.stats = parse_column_stats(g.data, g.len);
                }
            }
            out.push_back(std::move(e));
        }
    }
    const StreamInfo* bs = sf.find(column, S_BLOOM_FILTER);
    if (!bs) bs = sf.find(column, S_BLOOM_FILTER_UTF8);
    if (!bs) return out;
    std::vector<uint8_t> raw = host_decompress_section(compression, block_size, base_for(bs->offset, bs->length) + bs->offset, bs->length);
    std::vector<BloomBits> filters;
    PbCursor c(raw.data(), raw.size());
    PbField f, g;
    while (c.next(f)) {
        if (f.number != 1) continue;
        want_wire(f, 2);
        BloomBits b;
        uint32_t k = 0;
        bool has_utf8 = false;
        std::vector<uint64_t> from_bytes;
        PbCursor bc(f.data, f.len);
        while (bc.next(g)) {
            if (g.number == 1) {
                want_wire(g, 0);
                k = (uint32_t)g.value;
            } else if (g.number == 2) {
                if (g.wire == 1) b.bitset.push_back(g.value);
                else if (g.wire == 2) {
                    if (g.len % 8) fail(ORCB_DECODE_PROTO, "buffer underflow in a packed fixed64 field");
                    for (size_t i = 0; i < g.len; i += 8) {
                        uint64_t w;
                        memcpy(&w, g.data + i, 8);
                        b.bitset.push_back(w);
                    }
                } else fail(ORCB_DECODE_PROTO, "invalid wire type for BloomFilter.bitset");
            } else if (g.number == 3) {
                want_wire(g, 2);
                has_utf8 = true;
                from_bytes.clear();
                for (size_t i = 0; i < g.len; i += 8) {  // little-endian words, the last one zero padded
                    uint64_t w = 0;
                    memcpy(&w, g.data + i, std::min<size_t>(8, g.len - i));
                    from_bytes.push_back(w);
                }
            }
        }
        if (!b.bitset.empty() && has_utf8) throw ReferencePanic("Bloom filter proto has both bitset and utf8bitset populated");
        if (b.bitset.empty() && !has_utf8) continue;  // try_from_proto -> None, dropped by filter_map
        if (b.bitset.empty()) b.bitset = std::move(from_bytes);
        b.num_hash_functions = k ? k : 3;
        filters.push_back(std::move(b));
    }
    if (filters.size() != out.size())
        throw ReferencePanic("Bloom filter count mismatch: expected " + std::to_string(out.size()) + " but got " +
                             std::to_string(filters.size()) + " for column " + std::to_string(column));
    for (size_t i = 0; i < out.size(); i++) {
        out[i].has_bloom = true;
        out[i].bloom = std::move(filters[i]);
    }
    return out;
}

// src/compression.rs:244-275 — header walk only
std::vector<ChunkInfo> FileMeta::chunk_table(uint64_t stream_off, uint64_t stream_len) const {
    std::vector<ChunkInfo> out;
    const uint8_t* s = base_for(stream_off, stream_len) + stream_off;
    uint64_t p = 0;
    while (p < stream_len) {
        if (p + 3 > stream_len) fail(ORCB_OUT_OF_SPEC, "truncated compression chunk header");
        uint32_t h = s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16);
        ChunkInfo ci;
        ci.hdr_off = (uint32_t)p;
        ci.src_off = p + 3;
        ci.src_len = h >> 1;
        ci.original = h & 1;
        if (ci.src_off + ci.src_len > stream_len) fail(ORCB_OUT_OF_SPEC, "compression chunk exceeds stream");
        ci.dst_len = -1;
        if (ci.original) {
            ci.dst_len = ci.src_len;
        } else if (compression == C_SNAPPY) {
            // uncompressed length preamble (snap::raw::decompress_len, src/compression.rs:163-165)
            uint64_t v = 0;
            uint64_t q = ci.src_off;
            bool ok = false;
            for (int shift = 0; shift <= 35 && q < ci.src_off + ci.src_len; shift += 7) {
                uint8_t b = s[q++];
                v |= (uint64_t)(b & 0x7f) << shift;
                if (b < 0x80) { ok = true; break; }
            }
            if (!ok) fail(ORCB_BUILD_SNAPPY_DECODER, "bad snappy preamble");
            ci.dst_len = (int64_t)v;
        } else {
            std::lock_guard<std::mutex> lock(chunk_sizes->mu);
            auto it = chunk_sizes->size.find(stream_off + p);
            if (it != chunk_sizes->size.end()) ci.dst_len = it->second;
        }
        if (ci.dst_len < 0 && compression == C_ZSTD) {
            // Frame_Content_Size of the (normally only) frame of the chunk
            // ... when that frame is all there is in the chunk (block headers walked, nothing decoded)
            zstd::FrameHeader fh;
            const uint8_t* f = s + ci.src_off;
            if (zstd::frame_header(f, ci.src_len, fh) && !fh.skippable && fh.content <= block_size) {
                uint64_t q = fh.hdr;
                bool whole = false;
                while (q + 3 <= ci.src_len) {
                    const uint32_t bh = f[q] | ((uint32_t)f[q + 1] << 8) | ((uint32_t)f[q + 2] << 16);
                    q += 3 + (((bh >> 1) & 3) == 1 ? 1u : (bh >> 3));
                    if (bh & 1) {
                        whole = q + (fh.checksum ? 4 : 0) == ci.src_len;
                        break;
                    }
                }
                if (whole) {
                    ci.dst_len = (int64_t)fh.content;
                    ci.guess = true;
                }
            }
        }
        out.push_back(ci);
        p = ci.src_off + ci.src_len;
    }
    return out;
}

}  // namespace orcb
