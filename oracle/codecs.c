/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C CPU restatement of the value codecs on orc-rust's decode path (reference v0.8.0,
 * paths relative to /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The CUDA product under
 * orc_rust_b200/ never links or calls it.
 *
 * Parity pinning: checked against the reference's own known-answer vectors (tests/test_oracle_kat.py:
 * rle_v2/mod.rs:587-692, rle_v1.rs:434-466, byte.rs:343-356, boolean.rs:176-211, util.rs:769-809,
 * decimal.rs:55-139, compression.rs:353-370) and, at file level, against the 26 expected_arrow
 * feather goldens + pyarrow.orc (tests/test_oracle_files.py).
 *
 * Every function cites the reference lines it restates.  Error codes are OrcError variant ordinal + 1
 * (src/error.rs:31-174); 0 = Ok.  Reference panics (index OOB, slice underflow) are mapped to OutOfSpec.
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <zlib.h>
#include <dlfcn.h>

#define ORC_OK 0
#define ERR_IO 1
#define ERR_OUT_OF_SPEC 3
#define ERR_DECODE_TIMESTAMP 5
#define ERR_OFFSET_OVERFLOW 6
#define ERR_VARINT_TOO_LARGE 12
#define ERR_UNEXPECTED 13
#define ERR_BUILD_ZSTD 14
#define ERR_BUILD_SNAPPY 15
#define ERR_BUILD_LZO 16
#define ERR_BUILD_LZ4 17
#define ERR_ARROW 18

typedef __int128 i128;
typedef unsigned __int128 u128;

typedef struct {
    const uint8_t *p;
    size_t len;
    size_t pos;
} Rd;

static inline int rd_u8(Rd *r, uint8_t *b) {
    if (r->pos >= r->len) return ERR_IO; /* read_exact on EOF -> IoError (encoding/util.rs:26-30) */
    *b = r->p[r->pos++];
    return ORC_OK;
}

/* Truncate an i64 to an N-byte two's complement integer and sign-extend back (N::from_i64, `as` casts). */
static inline int64_t trunc_n(int64_t v, int nbytes) {
    if (nbytes >= 8) return v;
    int sh = 64 - 8 * nbytes;
    return (int64_t)((uint64_t)v << sh) >> sh;
}

/* signed_zigzag_decode in width N (integer/util.rs:536-546): (v >>> 1) ^ -(v & 1) */
static inline int64_t zigzag_n(int64_t v, int nbytes) {
    uint64_t mask = nbytes >= 8 ? ~0ull : ((1ull << (8 * nbytes)) - 1);
    uint64_t u = (uint64_t)v & mask;
    uint64_t r = (u >> 1) ^ (0 - (u & 1));
    return trunc_n((int64_t)(r & mask), nbytes);
}

/* checked add of an i64 delta to an N-typed accumulator: NInt::add_i64 / sub_i64 (integer/mod.rs:236-313) */
static inline int add_i64_n(int64_t acc, int64_t d, int nbytes, int64_t *out) {
    int64_t r;
    if (__builtin_add_overflow(acc, d, &r)) return 0;
    if (trunc_n(r, nbytes) != r) return 0;
    *out = r;
    return 1;
}
static inline int sub_i64_n(int64_t acc, int64_t d, int nbytes, int64_t *out) {
    int64_t r;
    if (__builtin_sub_overflow(acc, d, &r)) return 0;
    if (trunc_n(r, nbytes) != r) return 0;
    *out = r;
    return 1;
}

/* rle_v2_decode_bit_width (integer/util.rs:370-384) */
static int decode_bit_width(int enc) {
    static const int tail[8] = {26, 28, 30, 32, 40, 48, 56, 64};
    return enc <= 23 ? enc + 1 : tail[enc - 24];
}
/* get_closest_fixed_bits (integer/util.rs:407-421) */
static int closest_fixed_bits(int n) {
    if (n == 0) return 1;
    if (n <= 24) return n;
    if (n <= 26) return 26;
    if (n <= 28) return 28;
    if (n <= 30) return 30;
    if (n <= 32) return 32;
    if (n <= 40) return 40;
    if (n <= 48) return 48;
    if (n <= 56) return 56;
    return 64;
}

/* read_ints (integer/util.rs:44-218): MSB-first big-endian bit-packed, each call starts byte aligned
 * and consumes ceil(n*w/8) bytes.  Values are returned as raw unsigned bit patterns in int64. */
static int read_ints(Rd *r, int64_t *out, size_t n, int w) {
    uint64_t cur = 0;
    int bits_left = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t v = 0;
        int need = w;
        while (need > 0) {
            if (bits_left == 0) {
                uint8_t b;
                int e = rd_u8(r, &b);
                if (e) return e;
                cur = b;
                bits_left = 8;
            }
            int take = need < bits_left ? need : bits_left;
            v = (take == 64) ? 0 : (v << take);
            v |= (cur >> (bits_left - take)) & ((1ull << take) - 1);
            bits_left -= take;
            need -= take;
        }
        out[i] = (int64_t)v;
    }
    return ORC_OK;
}

/* read_varint::<N> (integer/util.rs:475-498).  nbits = bit width of N (16/32/64). */
static int read_varint_n(Rd *r, int nbytes, int64_t *out) {
    uint64_t num = 0;
    unsigned off = 0;
    unsigned nbits = 8u * (unsigned)nbytes;
    for (;;) {
        uint8_t b;
        int e = rd_u8(r, &b);
        if (e) return e;
        if (off >= nbits) return ERR_VARINT_TOO_LARGE; /* checked_shl fails when shift >= bit width */
        num |= (uint64_t)(b & 0x7f) << off;
        off += 7;
        if ((b & 0x80) == 0) break;
    }
    *out = trunc_n((int64_t)num, nbytes);
    return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * RLE v2 (integer/rle_v2/ *.rs).  Decodes one run into run[] (<=512), returns its length.
 * ------------------------------------------------------------------------------------------- */
static int rle2_run(Rd *r, int is_signed, int nbytes, int64_t *run, size_t *run_len) {
    uint8_t h;
    if (r->pos >= r->len) return ERR_OUT_OF_SPEC; /* "not enough values to decode in RLE v2" mod.rs:115-122 */
    h = r->p[r->pos++];
    int kind = h >> 6;
    if (kind == 0) {
        /* SHORT_REPEAT short_repeat.rs:29-63 */
        int bw = ((h >> 3) & 7) + 1;
        if (nbytes < bw) return ERR_OUT_OF_SPEC;
        size_t len = (size_t)(h & 7) + 3;
        uint64_t v = 0;
        for (int i = 0; i < bw; i++) {
            uint8_t b;
            int e = rd_u8(r, &b);
            if (e) return e;
            v = (v << 8) | b;
        }
        int64_t val = trunc_n((int64_t)v, nbytes);
        if (is_signed) val = zigzag_n(val, nbytes);
        for (size_t i = 0; i < len; i++) run[i] = val;
        *run_len = len;
        return ORC_OK;
    }
    if (kind == 1) {
        /* DIRECT direct.rs:39-65 */
        int w = decode_bit_width((h >> 1) & 0x1f);
        if (nbytes * 8 < w) return ERR_OUT_OF_SPEC;
        uint8_t b2;
        int e = rd_u8(r, &b2);
        if (e) return e;
        size_t len = (size_t)((((unsigned)h & 1) << 8) | b2) + 1;
        e = read_ints(r, run, len, w);
        if (e) return e;
        for (size_t i = 0; i < len; i++) {
            int64_t v = trunc_n(run[i], nbytes);
            run[i] = is_signed ? zigzag_n(v, nbytes) : v;
        }
        *run_len = len;
        return ORC_OK;
    }
    if (kind == 2) {
        /* PATCHED_BASE patched_base.rs:38-151 */
        int w = decode_bit_width((h >> 1) & 0x1f);
        uint8_t b2, b3, b4;
        int e;
        if ((e = rd_u8(r, &b2))) return e;
        size_t len = (size_t)((((unsigned)h & 1) << 8) | b2) + 1;
        if ((e = rd_u8(r, &b3))) return e;
        if ((e = rd_u8(r, &b4))) return e;
        int base_bw = ((b3 >> 5) & 7) + 1;
        int pw = decode_bit_width(b3 & 0x1f);
        int pgw = ((b4 >> 5) & 7) + 1;
        if (pw + pgw > 64) return ERR_OUT_OF_SPEC;
        size_t pll = b4 & 0x1f;
        uint64_t ub = 0;
        for (int i = 0; i < base_bw; i++) {
            uint8_t b;
            if ((e = rd_u8(r, &b))) return e;
            ub = (ub << 8) | b;
        }
        int64_t base = (int64_t)ub;
        if (is_signed) { /* signed_msb_decode util.rs:559-569 */
            uint64_t msb = 1ull << (base_bw * 8 - 1);
            if (ub & msb) base = (int64_t)(0 - (ub & ~msb));
            else base = (int64_t)(ub & ~msb);
        }
        base = trunc_n(base, nbytes);
        /* value width wider than N: the reference panics (byte-aligned) or silently truncates
         * (unaligned); mapped to OutOfSpec here. */
        if (nbytes * 8 < w) return ERR_OUT_OF_SPEC;
        if ((e = read_ints(r, run, len, w))) return e;
        for (size_t i = 0; i < len; i++) run[i] = trunc_n(run[i], nbytes);
        int64_t patches[32];
        if ((e = read_ints(r, patches, pll, closest_fixed_bits(pw + pgw)))) return e;
        if (pll == 0) return ERR_OUT_OF_SPEC; /* patches[0] index panic :98 */
        size_t pi = 0;
        uint64_t pmask = pw >= 64 ? 0 : ((1ull << pw) - 1);
        int64_t cur_gap = (int64_t)((uint64_t)patches[pi] >> pw);
        int64_t cur_patch = (int64_t)((uint64_t)patches[pi] & pmask);
        int64_t actual_gap = 0;
        while (cur_gap == 255 && cur_patch == 0) {
            actual_gap += 255;
            pi++;
            if (pi >= pll) return ERR_OUT_OF_SPEC; /* index panic */
            cur_gap = (int64_t)((uint64_t)patches[pi] >> pw);
            cur_patch = (int64_t)((uint64_t)patches[pi] & pmask);
        }
        actual_gap += cur_gap;
        for (size_t idx = 0; idx < len; idx++) {
            if ((int64_t)idx == actual_gap) {
                if (w >= 64) return ERR_OUT_OF_SPEC; /* checked_shl(64) -> None :112-117 */
                int64_t pbits = trunc_n((int64_t)((uint64_t)cur_patch << w), nbytes);
                int64_t pv = run[idx] | pbits;
                run[idx] = trunc_n((int64_t)((uint64_t)pv + (uint64_t)base), nbytes); /* wrapping_add :122-124 */
                pi++;
                if (pi < pll) {
                    cur_gap = (int64_t)((uint64_t)patches[pi] >> pw);
                    cur_patch = (int64_t)((uint64_t)patches[pi] & pmask);
                    actual_gap = 0;
                    while (cur_gap == 255 && cur_patch == 0) {
                        actual_gap += 255;
                        pi++;
                        if (pi >= pll) return ERR_OUT_OF_SPEC;
                        cur_gap = (int64_t)((uint64_t)patches[pi] >> pw);
                        cur_patch = (int64_t)((uint64_t)patches[pi] & pmask);
                    }
                    actual_gap += cur_gap;
                    actual_gap += (int64_t)idx;
                }
            } else {
                int64_t s;
                if (!add_i64_n(run[idx], base, nbytes, &s)) return ERR_OUT_OF_SPEC; /* checked_add :144-146 */
                run[idx] = s;
            }
        }
        *run_len = len;
        return ORC_OK;
    }
    /* DELTA delta.rs:44-116 */
    {
        int enc = (h >> 1) & 0x1f;
        int w = enc == 0 ? 0 : decode_bit_width(enc);
        uint8_t b2;
        int e;
        if ((e = rd_u8(r, &b2))) return e;
        size_t len = (size_t)((((unsigned)h & 1) << 8) | b2) + 1;
        int64_t base;
        if ((e = read_varint_n(r, nbytes, &base))) return e;
        if (is_signed) base = zigzag_n(base, nbytes);
        run[0] = base;
        int64_t d0;
        if ((e = read_varint_n(r, 8, &d0))) return e;
        d0 = zigzag_n(d0, 8);
        int positive = d0 > 0;                       /* is_positive(): false for 0 */
        int64_t mag = d0 < 0 ? (int64_t)(0 - (uint64_t)d0) : d0; /* abs(), wrapping for i64::MIN */
        size_t produced = 1;
        if (w == 0) {
            int64_t acc = base;
            for (size_t i = 1; i < len; i++) {
                int ok = positive ? add_i64_n(acc, mag, nbytes, &acc) : sub_i64_n(acc, mag, nbytes, &acc);
                if (!ok) return ERR_OUT_OF_SPEC;
                run[produced++] = acc;
            }
        } else {
            int64_t acc;
            int ok = positive ? add_i64_n(base, mag, nbytes, &acc) : sub_i64_n(base, mag, nbytes, &acc);
            if (!ok) return ERR_OUT_OF_SPEC;
            run[produced++] = acc;
            /* `length - 2` underflows (usize) when len == 1: the reference then tries to read ~2^64
             * ints and fails with IoError at EOF (or panics in debug) -> report IoError. */
            if (len < 2) return ERR_IO;
            size_t nd = len - 2;
            int64_t deltas[512];
            if ((e = read_ints(r, deltas, nd, w))) return e;
            for (size_t i = 0; i < nd; i++) {
                ok = positive ? add_i64_n(acc, deltas[i], nbytes, &acc) : sub_i64_n(acc, deltas[i], nbytes, &acc);
                if (!ok) return ERR_OUT_OF_SPEC;
                run[produced++] = acc;
            }
        }
        *run_len = produced;
        return ORC_OK;
    }
}

/* GenericRle::decode loop (encoding/rle.rs:68-105) over RleV2Decoder::decode_batch (rle_v2/mod.rs:112-146).
 * Decodes exactly n values from the start of `in`; out is int64 regardless of N. */
int orc_oracle_rle_v2(const uint8_t *in, size_t in_len, int is_signed, int nbytes, int64_t *out, size_t n,
                      size_t *consumed) {
    Rd r = {in, in_len, 0};
    int64_t run[512 + 8];
    size_t done = 0;
    while (done < n) {
        size_t rl = 0;
        int e = rle2_run(&r, is_signed, nbytes, run, &rl);
        if (e) return e;
        size_t take = rl < n - done ? rl : n - done;
        memcpy(out + done, run, take * sizeof(int64_t));
        done += take;
    }
    if (consumed) *consumed = r.pos;
    return ORC_OK;
}

/* RLE v1 (integer/rle_v1.rs:54-68, 90-159) */
int orc_oracle_rle_v1(const uint8_t *in, size_t in_len, int is_signed, int nbytes, int64_t *out, size_t n,
                      size_t *consumed) {
    Rd r = {in, in_len, 0};
    size_t done = 0;
    int64_t run[130];
    while (done < n) {
        if (r.pos >= r.len) return ERR_OUT_OF_SPEC; /* "not enough values to decode" */
        int8_t h = (int8_t)r.p[r.pos++];
        size_t rl;
        int e;
        if (h < 0) {
            rl = (size_t)(-(int)h);
            for (size_t i = 0; i < rl; i++) {
                int64_t v;
                if ((e = read_varint_n(&r, nbytes, &v))) return e;
                run[i] = is_signed ? zigzag_n(v, nbytes) : v;
            }
        } else {
            rl = (size_t)(uint8_t)h + 3;
            uint8_t db;
            if ((e = rd_u8(&r, &db))) return e;
            int8_t delta = (int8_t)db;
            int64_t base;
            if ((e = read_varint_n(&r, nbytes, &base))) return e;
            if (is_signed) base = zigzag_n(base, nbytes);
            run[0] = base;
            int64_t mag = delta < 0 ? -(int64_t)delta : (int64_t)delta;
            for (size_t i = 1; i < rl; i++) {
                int ok = delta < 0 ? sub_i64_n(base, mag, nbytes, &base) : add_i64_n(base, mag, nbytes, &base);
                if (!ok) return ERR_OUT_OF_SPEC;
                run[i] = base;
            }
        }
        size_t take = rl < n - done ? rl : n - done;
        memcpy(out + done, run, take * sizeof(int64_t));
        done += take;
    }
    if (consumed) *consumed = r.pos;
    return ORC_OK;
}

/* Byte RLE (encoding/byte.rs:228-247) */
int orc_oracle_byte_rle(const uint8_t *in, size_t in_len, uint8_t *out, size_t n, size_t *consumed) {
    Rd r = {in, in_len, 0};
    size_t done = 0;
    while (done < n) {
        uint8_t h;
        int e = rd_u8(&r, &h);
        if (e) return e;
        if (h < 0x80) {
            size_t rl = (size_t)h + 3;
            uint8_t v;
            if ((e = rd_u8(&r, &v))) return e;
            size_t take = rl < n - done ? rl : n - done;
            memset(out + done, v, take);
            done += take;
        } else {
            size_t rl = 0x100 - (size_t)h;
            if (r.pos + rl > r.len) return ERR_IO;
            size_t take = rl < n - done ? rl : n - done;
            memcpy(out + done, r.p + r.pos, take);
            r.pos += rl;
            done += take;
        }
    }
    if (consumed) *consumed = r.pos;
    return ORC_OK;
}

/* Boolean RLE (encoding/boolean.rs:101-113): byte RLE then MSB-first bits; out = one byte per value. */
int orc_oracle_bool_rle(const uint8_t *in, size_t in_len, uint8_t *out, size_t n) {
    size_t nbytes = (n + 7) / 8;
    uint8_t *tmp = (uint8_t *)malloc(nbytes ? nbytes : 1);
    int e = orc_oracle_byte_rle(in, in_len, tmp, nbytes, NULL);
    if (e) {
        free(tmp);
        return e;
    }
    for (size_t i = 0; i < n; i++) out[i] = (tmp[i >> 3] >> (7 - (i & 7))) & 1;
    free(tmp);
    return ORC_OK;
}

/* Pack bool bytes into an Arrow LSB-first bitmap (NullBuffer::from(Vec<bool>), array_decoder/mod.rs:209-213). */
void orc_oracle_pack_bits(const uint8_t *bools, size_t n, uint8_t *bitmap) {
    memset(bitmap, 0, (n + 7) / 8);
    for (size_t i = 0; i < n; i++)
        if (bools[i]) bitmap[i >> 3] |= (uint8_t)(1u << (i & 7));
}

/* Unbounded zigzag varint -> i128 (encoding/decimal.rs:46-51 + util.rs:475-527). out = n * 16 bytes LE. */
int orc_oracle_varint_i128(const uint8_t *in, size_t in_len, i128 *out, size_t n, size_t *consumed) {
    Rd r = {in, in_len, 0};
    for (size_t i = 0; i < n; i++) {
        u128 num = 0;
        unsigned off = 0;
        for (;;) {
            uint8_t b;
            int e = rd_u8(&r, &b);
            if (e) return e;
            if (off >= 128) return ERR_VARINT_TOO_LARGE;
            num |= (u128)(b & 0x7f) << off;
            off += 7;
            if ((b & 0x80) == 0) break;
        }
        u128 zz = (num >> 1) ^ (u128)(0 - (num & 1));
        i128 v = (i128)zz;
        memcpy(&out[i], &v, 16);
    }
    if (consumed) *consumed = r.pos;
    return ORC_OK;
}

/* fix_i128_scale (array_decoder/decimal.rs:138-166): release-mode wrapping mul, truncating div. */
void orc_oracle_decimal_fix_scale(i128 *vals, const int64_t *scales, size_t n, uint32_t fixed_scale) {
    for (size_t i = 0; i < n; i++) {
        uint32_t vs = (uint32_t)(int32_t)scales[i];
        i128 v;
        memcpy(&v, &vals[i], 16);
        if (fixed_scale < vs) {
            uint32_t k = vs - fixed_scale;
            u128 f = 1;
            int overflow = 0;
            for (uint32_t j = 0; j < k; j++) {
                if (f > (~(u128)0 >> 1) / 10) { overflow = 1; break; }
                f *= 10;
            }
            /* 10_i128.pow overflow panics in the reference; value/huge == 0 is the closest total result */
            v = overflow ? 0 : v / (i128)f;
        } else if (fixed_scale > vs) {
            uint32_t k = fixed_scale - vs;
            u128 f = 1;
            for (uint32_t j = 0; j < k && j < 64; j++) f *= 10;
            v = (i128)((u128)v * f);
        }
        memcpy(&vals[i], &v, 16);
    }
}

/* Timestamp combine (encoding/timestamp.rs:121-192).  unit_ns = nanoseconds per output unit.
 * Returns DecodeTimestamp on precision loss / i64 overflow. */
int orc_oracle_timestamp(const int64_t *data, const int64_t *secondary, size_t n, int64_t base, int64_t unit_ns,
                         int64_t *out) {
    for (size_t i = 0; i < n; i++) {
        uint64_t nanos = (uint64_t)secondary[i];
        unsigned zeros = nanos & 7;
        nanos >>= 3;
        if (zeros != 0) {
            uint64_t p = 1;
            for (unsigned j = 0; j < zeros + 1; j++) p *= 10;
            nanos *= p; /* wrapping in release */
        }
        int64_t secs = (int64_t)((uint64_t)data[i] + (uint64_t)base);
        if (secs < 0 && nanos > 999999) secs -= 1;
        i128 ns = (i128)secs * 1000000000 + (i128)nanos;
        if (ns % unit_ns != 0) return ERR_DECODE_TIMESTAMP;
        i128 q = ns / unit_ns;
        if (q > (i128)INT64_MAX || q < (i128)INT64_MIN) return ERR_DECODE_TIMESTAMP;
        out[i] = (int64_t)q;
    }
    return ORC_OK;
}

/* decode_timestamp_as_i128 (encoding/timestamp.rs:194-197) */
void orc_oracle_timestamp_i128(const int64_t *data, const int64_t *secondary, size_t n, int64_t base, i128 *out) {
    for (size_t i = 0; i < n; i++) {
        uint64_t nanos = (uint64_t)secondary[i];
        unsigned zeros = nanos & 7;
        nanos >>= 3;
        if (zeros != 0) {
            uint64_t p = 1;
            for (unsigned j = 0; j < zeros + 1; j++) p *= 10;
            nanos *= p;
        }
        int64_t secs = (int64_t)((uint64_t)data[i] + (uint64_t)base);
        if (secs < 0 && nanos > 999999) secs -= 1;
        i128 ns = (i128)secs * 1000000000 + (i128)nanos;
        memcpy(&out[i], &ns, 16);
    }
}

/* PrimitiveValueDecoder::decode_spaced (encoding/mod.rs:64-91): dense values -> rows, null slots = 0. */
void orc_oracle_spaced(const uint8_t *dense, const uint8_t *present_bools, size_t n_rows, size_t width,
                       uint8_t *out) {
    size_t k = 0;
    memset(out, 0, n_rows * width);
    for (size_t i = 0; i < n_rows; i++) {
        if (present_bools[i]) {
            memcpy(out + i * width, dense + k * width, width);
            k++;
        }
    }
}

/* OffsetBuffer::from_lengths per batch (array_decoder/string.rs:142-143): offsets restart at 0. */
int orc_oracle_offsets(const int64_t *lengths, size_t n, int32_t *offsets) {
    int64_t acc = 0;
    offsets[0] = 0;
    for (size_t i = 0; i < n; i++) {
        /* an unsigned length >= 2^63 arrives as a negative i64; `l as usize` then overflows
         * OffsetBuffer::from_lengths (a panic in the reference, reported as OutOfSpec here) */
        if (lengths[i] < 0) return ERR_OUT_OF_SPEC;
        acc += lengths[i];
        if (acc > INT32_MAX) return ERR_OFFSET_OVERFLOW;
        offsets[i + 1] = (int32_t)acc;
    }
    return ORC_OK;
}

/* cast(Dictionary<Int64,Utf8> -> Utf8) (array_decoder/string.rs:205-224): valid keys bounds-checked
 * (DictionaryArray::try_new -> Arrow error), null slots produce empty strings.
 * Pass 1 (out_data == NULL) returns total bytes in *total. */
int orc_oracle_dict_gather(const int64_t *keys, const uint8_t *present_bools /* may be NULL */, size_t n,
                           const int32_t *dict_offsets, size_t dict_size, const uint8_t *dict_data,
                           int32_t *out_offsets, uint8_t *out_data, int64_t *total) {
    int64_t acc = 0;
    out_offsets[0] = 0;
    for (size_t i = 0; i < n; i++) {
        if (!present_bools || present_bools[i]) {
            int64_t k = keys[i];
            if (k < 0 || (uint64_t)k >= dict_size) return ERR_ARROW;
            int32_t a = dict_offsets[k], b = dict_offsets[k + 1];
            if (out_data) memcpy(out_data + acc, dict_data + a, (size_t)(b - a));
            acc += b - a;
        }
        if (acc > INT32_MAX) return ERR_ARROW;
        out_offsets[i + 1] = (int32_t)acc;
    }
    *total = acc;
    return ORC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Chunk framing + block codecs (src/compression.rs).
 * ------------------------------------------------------------------------------------------- */

/* decode_header (compression.rs:113-123) */
void orc_oracle_chunk_header(const uint8_t b[3], uint32_t *length, int *is_original) {
    uint32_t v = (uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16);
    *is_original = (int)(v & 1);
    *length = v >> 1;
}

/* Snappy raw block (snap 1.1 `raw::Decoder::decompress`, call site compression.rs:161-172).
 * Format per the published Snappy framing description: varint preamble + literal/copy elements. */
static int64_t snappy_block(const uint8_t *s, size_t n, uint8_t *d, size_t cap) {
    size_t p = 0;
    uint64_t ulen = 0;
    unsigned sh = 0;
    for (;;) {
        if (p >= n || sh > 35) return -ERR_BUILD_SNAPPY;
        uint8_t b = s[p++];
        ulen |= (uint64_t)(b & 0x7f) << sh;
        sh += 7;
        if (!(b & 0x80)) break;
    }
    if (ulen > cap) return -ERR_BUILD_SNAPPY;
    size_t o = 0;
    while (p < n) {
        uint8_t tag = s[p++];
        unsigned t = tag & 3;
        if (t == 0) {
            size_t len = (tag >> 2);
            if (len >= 60) {
                unsigned nb = (unsigned)len - 59;
                if (p + nb > n) return -ERR_BUILD_SNAPPY;
                len = 0;
                for (unsigned i = 0; i < nb; i++) len |= (size_t)s[p + i] << (8 * i);
                p += nb;
            }
            len += 1;
            if (p + len > n || o + len > ulen) return -ERR_BUILD_SNAPPY;
            memcpy(d + o, s + p, len);
            p += len;
            o += len;
        } else {
            size_t len, off;
            if (t == 1) {
                if (p + 1 > n) return -ERR_BUILD_SNAPPY;
                len = ((tag >> 2) & 7) + 4;
                off = ((size_t)(tag >> 5) << 8) | s[p];
                p += 1;
            } else if (t == 2) {
                if (p + 2 > n) return -ERR_BUILD_SNAPPY;
                len = (tag >> 2) + 1;
                off = (size_t)s[p] | ((size_t)s[p + 1] << 8);
                p += 2;
            } else {
                if (p + 4 > n) return -ERR_BUILD_SNAPPY;
                len = (tag >> 2) + 1;
                off = (size_t)s[p] | ((size_t)s[p + 1] << 8) | ((size_t)s[p + 2] << 16) | ((size_t)s[p + 3] << 24);
                p += 4;
            }
            if (off == 0 || off > o || o + len > ulen) return -ERR_BUILD_SNAPPY;
            for (size_t i = 0; i < len; i++) d[o + i] = d[o + i - off];
            o += len;
        }
    }
    if (o != ulen) return -ERR_BUILD_SNAPPY;
    return (int64_t)o;
}

/* LZ4 block (lz4_flex 0.11 `block::decompress(src, max)`, call site compression.rs:185-195). */
static int64_t lz4_block(const uint8_t *s, size_t n, uint8_t *d, size_t cap) {
    size_t p = 0, o = 0;
    while (p < n) {
        uint8_t tok = s[p++];
        size_t ll = tok >> 4;
        if (ll == 15) {
            for (;;) {
                if (p >= n) return -ERR_BUILD_LZ4;
                uint8_t b = s[p++];
                ll += b;
                if (b != 255) break;
            }
        }
        if (p + ll > n || o + ll > cap) return -ERR_BUILD_LZ4;
        memcpy(d + o, s + p, ll);
        p += ll;
        o += ll;
        if (p >= n) break; /* last sequence: literals only */
        if (p + 2 > n) return -ERR_BUILD_LZ4;
        size_t off = (size_t)s[p] | ((size_t)s[p + 1] << 8);
        p += 2;
        size_t ml = tok & 15;
        if (ml == 15) {
            for (;;) {
                if (p >= n) return -ERR_BUILD_LZ4;
                uint8_t b = s[p++];
                ml += b;
                if (b != 255) break;
            }
        }
        ml += 4;
        if (off == 0 || off > o || o + ml > cap) return -ERR_BUILD_LZ4;
        for (size_t i = 0; i < ml; i++) d[o + i] = d[o + i - off];
        o += ml;
    }
    return (int64_t)o;
}

/* raw deflate (flate2 DeflateDecoder, compression.rs:142-149) via system zlib */
static int64_t zlib_block(const uint8_t *s, size_t n, uint8_t *d, size_t cap) {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) return -ERR_IO;
    zs.next_in = (Bytef *)s;
    zs.avail_in = (uInt)n;
    zs.next_out = d;
    zs.avail_out = (uInt)cap;
    int rc = inflate(&zs, Z_FINISH);
    int64_t produced = (int64_t)zs.total_out;
    inflateEnd(&zs);
    if (rc != Z_STREAM_END && !(rc == Z_BUF_ERROR && zs.avail_in == 0)) return -ERR_IO;
    return produced;
}

/* Zstandard (zstd 0.13 crate = libzstd, `zstd::Decoder` + read_to_end, compression.rs:151-159): the system's libzstd,
 * loaded at run time (the image ships libzstd.so.1 without headers).  ZSTD_decompress handles concatenated and
 * skippable frames like the streaming decoder does.  Decode errors surface from read_to_end: IoError. */
static int64_t zstd_block(const uint8_t *s, size_t n, uint8_t *d, size_t cap) {
    typedef size_t (*dec_fn)(void *, size_t, const void *, size_t);
    typedef unsigned (*err_fn)(size_t);
    static dec_fn dec;
    static err_fn is_err;
    static int tried;
    if (!tried) {
        void *h = dlopen("libzstd.so.1", RTLD_NOW);
        if (h) {
            dec = (dec_fn)dlsym(h, "ZSTD_decompress");
            is_err = (err_fn)dlsym(h, "ZSTD_isError");
        }
        tried = 1;
    }
    if (!dec || !is_err) return -ERR_BUILD_ZSTD; /* no libzstd on this machine */
    if (n == 0) return 0;
    size_t r = dec(d, cap, s, n);
    if (is_err(r)) return -ERR_IO;
    return (int64_t)r;
}

/* LZO1X (lzokay-native 0.1 `decompress_all`, call site compression.rs:174-183; the published LZO1X stream format:
 * first byte > 17 = literal run of byte-17; then instructions M1 (0..15, meaning depends on how many literals the
 * previous instruction copied), M2 (>= 64), M3 (32..63), M4 (16..31, distance 16384 = end of stream). */
static int64_t lzo_block(const uint8_t *s, size_t n, uint8_t *d, size_t cap) {
    size_t ip = 0, op = 0;
    unsigned state = 0; /* literals copied by the previous instruction: 0..3, or 4 after a run */
    if (n == 0) return -ERR_BUILD_LZO;
    if (s[0] > 17) {
        size_t len = (size_t)s[0] - 17;
        ip = 1;
        if (ip + len > n || op + len > cap) return -ERR_BUILD_LZO;
        memcpy(d + op, s + ip, len);
        ip += len;
        op += len;
        state = len < 4 ? (unsigned)len : 4;
    }
    for (;;) {
        if (ip >= n) return -ERR_BUILD_LZO;
        unsigned inst = s[ip++];
        size_t mlen, mdist;
        unsigned trail;
        if (inst >= 64) {
            if (ip >= n) return -ERR_BUILD_LZO;
            mlen = (inst >> 5) + 1;
            mdist = ((size_t)s[ip++] << 3) + ((inst >> 2) & 7) + 1;
            trail = inst & 3;
        } else if (inst >= 32) {
            mlen = inst & 31;
            if (mlen == 0) {
                while (ip < n && s[ip] == 0) { mlen += 255; ip++; }
                if (ip >= n) return -ERR_BUILD_LZO;
                mlen += 31 + s[ip++];
            }
            mlen += 2;
            if (ip + 2 > n) return -ERR_BUILD_LZO;
            unsigned v = s[ip] | ((unsigned)s[ip + 1] << 8);
            ip += 2;
            mdist = (v >> 2) + 1;
            trail = v & 3;
        } else if (inst >= 16) {
            mlen = inst & 7;
            if (mlen == 0) {
                while (ip < n && s[ip] == 0) { mlen += 255; ip++; }
                if (ip >= n) return -ERR_BUILD_LZO;
                mlen += 7 + s[ip++];
            }
            mlen += 2;
            if (ip + 2 > n) return -ERR_BUILD_LZO;
            unsigned v = s[ip] | ((unsigned)s[ip + 1] << 8);
            ip += 2;
            mdist = 16384 + ((size_t)(inst & 8) << 11) + (v >> 2);
            trail = v & 3;
            if (mdist == 16384) {
                if (mlen != 3) return -ERR_BUILD_LZO;
                if (ip != n) return -ERR_BUILD_LZO; /* InputNotConsumed */
                return (int64_t)op;
            }
        } else if (state == 0) {
            size_t len = inst;
            if (len == 0) {
                while (ip < n && s[ip] == 0) { len += 255; ip++; }
                if (ip >= n) return -ERR_BUILD_LZO;
                len += 15 + s[ip++];
            }
            len += 3;
            if (ip + len > n || op + len > cap) return -ERR_BUILD_LZO;
            memcpy(d + op, s + ip, len);
            ip += len;
            op += len;
            state = 4;
            continue;
        } else {
            if (ip >= n) return -ERR_BUILD_LZO;
            unsigned h = s[ip++];
            if (state == 4) { mdist = ((size_t)h << 2) + (inst >> 2) + 2049; mlen = 3; }
            else { mdist = ((size_t)h << 2) + (inst >> 2) + 1; mlen = 2; }
            trail = inst & 3;
        }
        if (mdist > op || op + mlen > cap) return -ERR_BUILD_LZO;
        for (size_t i = 0; i < mlen; i++) d[op + i] = d[op + i - mdist];
        op += mlen;
        if (ip + trail > n || op + trail > cap) return -ERR_BUILD_LZO;
        for (unsigned i = 0; i < trail; i++) d[op++] = s[ip++];
        state = trail;
    }
}

/* Decompressor (compression.rs:244-347): concatenate all chunks of a stream.
 * kind: 0 NONE 1 ZLIB 2 SNAPPY 3 LZO 4 LZ4 5 ZSTD.  Returns output length or -(error code).
 * stats (optional): [0] chunks, [1] compressed chunks. */
int64_t orc_oracle_decompress_stream(int kind, const uint8_t *in, size_t in_len, size_t block_size, uint8_t *out,
                                     size_t out_cap, int64_t *stats) {
    if (kind == 0) {
        if (in_len > out_cap) return -ERR_UNEXPECTED;
        memcpy(out, in, in_len);
        return (int64_t)in_len;
    }
    size_t p = 0, o = 0;
    while (p < in_len) {
        if (p + 3 > in_len) return -ERR_OUT_OF_SPEC; /* split_to(3) panic */
        uint32_t len;
        int orig;
        orc_oracle_chunk_header(in + p, &len, &orig);
        p += 3;
        if (p + len > in_len) return -ERR_OUT_OF_SPEC; /* split_to panic */
        if (stats) stats[0]++;
        if (orig) {
            if (o + len > out_cap) return -ERR_UNEXPECTED;
            memcpy(out + o, in + p, len);
            o += len;
        } else {
            int64_t r;
            if (stats) stats[1]++;
            size_t room = out_cap - o;
            if (kind == 2) r = snappy_block(in + p, len, out + o, room);
            else if (kind == 4) r = lz4_block(in + p, len, out + o, room < block_size ? room : block_size);
            else if (kind == 1) r = zlib_block(in + p, len, out + o, room);
            else if (kind == 3) r = lzo_block(in + p, len, out + o, room);
            else r = zstd_block(in + p, len, out + o, room);
            if (r < 0) return r;
            o += (size_t)r;
        }
        p += len;
    }
    return (int64_t)o;
}

/* Upper bound on the decompressed size of a stream (for buffer sizing by the caller). */
int64_t orc_oracle_decompress_bound(int kind, const uint8_t *in, size_t in_len, size_t block_size) {
    if (kind == 0) return (int64_t)in_len;
    size_t p = 0;
    int64_t tot = 0;
    while (p + 3 <= in_len) {
        uint32_t len;
        int orig;
        orc_oracle_chunk_header(in + p, &len, &orig);
        p += 3 + len;
        tot += orig ? len : (int64_t)block_size;
    }
    return tot + 64;
}
