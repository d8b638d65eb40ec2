// CUDA kernels of the B200 ORC stripe decoder (sm_100a).  All integer / byte work, HBM-bound:
// no tensor cores.  One warp owns one (stream, row-group) segment; lanes cooperate inside a run.
//
// Semantics follow the reference (datafusion-contrib/orc-rust v0.8.0) bit for bit; each kernel cites
// the functions it replaces.  Error words: first error per column-stripe wins (atomicCAS), value =
// OrcbStatus.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "dev.h"
#include "kernels.h"

namespace orcb {

#define FULL 0xffffffffu

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void set_err(uint32_t* err, uint32_t cs, uint32_t code) { atomicCAS(&err[cs], 0u, code); }

__device__ __forceinline__ uint32_t bswap32(uint32_t x) { return __byte_perm(x, 0, 0x0123); }

// w (1..64) bits, MSB-first big-endian, starting `bitpos` bits after p (read_ints, integer/util.rs:44-218).
// Reads aligned 32-bit words; may touch up to 11 bytes past the last needed byte (arenas are padded).
__device__ __forceinline__ uint64_t load_be_bits(const uint8_t* p, uint32_t bitpos, int w) {
    const uint8_t* a = p + (bitpos >> 3);
    const uintptr_t ai = (uintptr_t)a;
    const uint32_t* q = (const uint32_t*)(ai & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)(ai & 3) << 3) + (bitpos & 7);  // 0..31
    const uint32_t w0 = bswap32(__ldg(q));
    const uint32_t w1 = bswap32(__ldg(q + 1));
    const uint32_t w2 = (sh + (uint32_t)w > 64u) ? bswap32(__ldg(q + 2)) : 0u;
    const uint32_t hi = __funnelshift_l(w1, w0, sh);
    const uint32_t lo = __funnelshift_l(w2, w1, sh);
    const uint64_t t = ((uint64_t)hi << 32) | lo;
    return t >> (64 - w);
}

// same for w <= 32: two aligned words, 32-bit arithmetic only
__device__ __forceinline__ uint32_t load_be_bits32(const uint8_t* p, uint32_t bitpos, int w) {
    const uint8_t* a = p + (bitpos >> 3);
    const uintptr_t ai = (uintptr_t)a;
    const uint32_t* q = (const uint32_t*)(ai & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)(ai & 3) << 3) + (bitpos & 7);  // 0..31
    const uint32_t w0 = bswap32(__ldg(q));
    const uint32_t w1 = bswap32(__ldg(q + 1));
    return __funnelshift_l(w1, w0, sh) >> (32 - w);
}
// packed value of width <= 32 -> i64 with the reference's N-width zigzag semantics (see trunc_n / zigzag_n):
// for w <= 8*nbytes the 32-bit zigzag followed by sign extension gives the same bits
__device__ __forceinline__ int64_t finish32(uint32_t x, bool sg, int nb) {
    const uint32_t z = (x >> 1) ^ (0u - (x & 1));
    if (sg) return (int64_t)(int32_t)z;
    return nb >= 8 ? (int64_t)(uint64_t)x : (nb == 4 ? (int64_t)(int32_t)x : (int64_t)(int16_t)x);
}

// 32 bits of an LSB-first bitmap starting at an arbitrary bit position
__device__ __forceinline__ uint32_t load_bits32(const uint32_t* bm, uint64_t bitpos) {
    const uint64_t wi = bitpos >> 5;
    const uint32_t sh = (uint32_t)(bitpos & 31);
    const uint32_t a = bm[wi];
    const uint32_t b = sh ? bm[wi + 1] : 0u;
    return __funnelshift_r(a, b, sh);
}

__device__ __forceinline__ int64_t trunc_n(int64_t v, int nbytes) {
    if (nbytes >= 8) return v;
    const int sh = 64 - 8 * nbytes;
    return (int64_t)((uint64_t)v << sh) >> sh;
}
// signed_zigzag_decode in width N (integer/util.rs:536-546)
__device__ __forceinline__ int64_t zigzag_n(int64_t v, int nbytes) {
    const uint64_t mask = nbytes >= 8 ? ~0ull : ((1ull << (8 * nbytes)) - 1);
    const uint64_t u = (uint64_t)v & mask;
    const uint64_t r = (u >> 1) ^ (0ull - (u & 1));
    return trunc_n((int64_t)(r & mask), nbytes);
}
__device__ __forceinline__ bool in_range_n(__int128 v, int nbytes) {
    const __int128 lim = (__int128)1 << (8 * nbytes - 1);
    return v >= -lim && v < lim;
}
// rle_v2_decode_bit_width (integer/util.rs:370-384)
__device__ __forceinline__ int width_of(uint32_t code) {
    // codes 24..31 -> 26, 28, 30, 32, 40, 48, 56, 64: one byte select out of two constants
    return code <= 23 ? (int)code + 1 : (int)(__byte_perm(0x201E1C1Au, 0x40383028u, code - 24) & 0xffu);
}
// get_closest_fixed_bits (integer/util.rs:407-421)
__device__ __forceinline__ int closest_fixed_bits(int n) {
    if (n == 0) return 1;
    if (n <= 24) return n;
    if (n <= 26) return 26;
    if (n <= 28) return 28;
    if (n <= 30) return 30;
    if (n <= 32) return 32;
    return (n + 7) & ~7;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_incl_scan64(uint64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

struct SegCtx {
    const Seg* s;
    uint32_t* err;
    uint32_t* mis;
};

__device__ __forceinline__ void store_val(const SegCtx& c, uint64_t idx, int64_t v) {
    const Seg& s = *c.s;
    switch (s.out_kind) {
        case OUT_I16: ((int16_t*)s.out)[idx] = (int16_t)v; break;
        case OUT_I32: ((int32_t*)s.out)[idx] = (int32_t)v; break;
        case OUT_I64: ((int64_t*)s.out)[idx] = v; break;
        case OUT_LEN31:
            if ((uint64_t)v > 0x7fffffffull) set_err(c.err, s.colstripe, s.aux);
            ((int32_t*)s.out)[idx] = (int32_t)v;
            break;
        case OUT_SCALE:
            if ((uint32_t)(int32_t)v != s.aux) atomicOr(&c.mis[s.colstripe], 1u);
            ((int32_t*)s.out)[idx] = (int32_t)v;
            break;
        default: break;
    }
}

// value i of the current run -> output, clipped to [skip, skip + take)
#define EMIT(i, val)                                                                      \
    do {                                                                                  \
        uint32_t _i = (i);                                                                \
        if (_i >= skip && _i - skip < take) store_val(c, out_pos + (_i - skip), (val));   \
    } while (0)

// read_varint::<N> (integer/util.rs:475-498). returns 0 ok / status
__device__ __forceinline__ uint32_t parse_varint(const uint8_t* in, uint32_t& p, uint32_t len, int nbits, uint64_t& out) {
    uint64_t num = 0;
    uint32_t off = 0;
    for (;;) {
        if (p >= len) return ORCB_IO_ERROR;
        const uint32_t b = in[p++];
        if (off >= (uint32_t)nbits) return ORCB_VARINT_TOO_LARGE;
        num |= (uint64_t)(b & 0x7f) << off;
        off += 7;
        if (!(b & 0x80)) break;
    }
    out = num;
    return 0;
}


// DIRECT run of width <= 32 spread over the warp: lane handles values first, first+32, ...  Because 32*w bits
// is a whole number of 32-bit words, the bit phase of a lane never changes inside a run: per value this is
// two word loads, one funnel shift, one shift, the zigzag and a store.
template <typename OutT, bool SIGNED, bool CHECK31>
__device__ __forceinline__ void direct32_lane_loop(const uint8_t* data, int w, uint32_t first, uint32_t i_end, OutT* outp,
                                                   uint32_t* err, uint32_t colstripe, uint32_t aux) {
    if (first >= i_end) return;
    const uint32_t bit0 = first * (uint32_t)w;
    const uintptr_t ai = (uintptr_t)(data + (bit0 >> 3));
    const uint32_t* q = (const uint32_t*)(ai & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)(ai & 3) << 3) + (bit0 & 7);
    const int rs = 32 - w;
    bool bad = false;
    auto conv = [&](uint32_t lo, uint32_t hi) -> OutT {
        const uint32_t x = __funnelshift_l(bswap32(hi), bswap32(lo), sh) >> rs;
        if (CHECK31 && x > 0x7fffffffu) bad = true;
        if (SIGNED) return (OutT)(int32_t)((x >> 1) ^ (0u - (x & 1)));
        return (OutT)x;
    };
    uint32_t i = first;
    for (; i + 224 < i_end; i += 256) {
        uint32_t lo[8], hi[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            lo[u] = __ldg(q + u * w);
            hi[u] = __ldg(q + u * w + 1);
        }
        q += 8 * w;
#pragma unroll
        for (int u = 0; u < 8; u++) outp[32 * u] = conv(lo[u], hi[u]);
        outp += 256;
    }
    for (; i + 96 < i_end; i += 128) {
        const uint32_t a0 = __ldg(q), a1 = __ldg(q + 1);
        const uint32_t b0 = __ldg(q + w), b1 = __ldg(q + w + 1);
        const uint32_t c0 = __ldg(q + 2 * w), c1 = __ldg(q + 2 * w + 1);
        const uint32_t d0 = __ldg(q + 3 * w), d1 = __ldg(q + 3 * w + 1);
        q += 4 * w;
        outp[0] = conv(a0, a1);
        outp[32] = conv(b0, b1);
        outp[64] = conv(c0, c1);
        outp[96] = conv(d0, d1);
        outp += 128;
    }
    for (; i < i_end; i += 32) {
        const uint32_t a0 = __ldg(q), a1 = __ldg(q + 1);
        q += w;
        outp[0] = conv(a0, a1);
        outp += 32;
    }
    if (CHECK31 && bad) set_err(err, colstripe, aux);
}

// ------------------------------------------------------------------------------------------------
// RLE v2, ONE run decoded by all 32 lanes (integer/rle_v2/{direct,patched_base,delta}.rs).
// All arguments are warp-uniform.  Values i in [skip, skip+take) go to out_pos + (i - skip).
// ------------------------------------------------------------------------------------------------
__device__ uint32_t coop_run2(const SegCtx& c, uint32_t cur, uint32_t skip, uint32_t room, uint64_t out_pos,
                              uint32_t* patchmap, uint32_t& rl_out, uint32_t& bytes_out, uint32_t& take_out) {
    const Seg& s = *c.s;
    const uint8_t* in = (const uint8_t*)s.in;
    const uint32_t len = s.in_len;
    const int lane = threadIdx.x & 31;
    const int nb = s.nbytes;
    const bool sg = (s.flags & SEG_SIGNED) != 0;
    const uint32_t hdr = (uint32_t)load_be_bits(in + cur, 0, 32);
    const uint32_t h0 = hdr >> 24;
    const uint32_t kind = h0 >> 6;
    uint32_t rl, run_bytes, take;
    if (kind == 0) {
        // SHORT_REPEAT short_repeat.rs:29-63 (normally taken by the owning lane; kept for completeness)
        const int bw = (int)((h0 >> 3) & 7) + 1;
        if (nb < bw) return ORCB_OUT_OF_SPEC;
        rl = (h0 & 7) + 3;
        run_bytes = 1 + bw;
        if (cur + run_bytes > len) return ORCB_IO_ERROR;
        int64_t v = trunc_n((int64_t)load_be_bits(in + cur + 1, 0, bw * 8), nb);
        if (sg) v = zigzag_n(v, nb);
        take = min(rl > skip ? rl - skip : 0u, room);
        EMIT(lane, v);
    } else if (kind == 1) {
        // DIRECT direct.rs:39-65
        const int w = width_of((h0 >> 1) & 31);
        if (nb * 8 < w) return ORCB_OUT_OF_SPEC;
        if (cur + 2 > len) return ORCB_IO_ERROR;
        rl = (((h0 & 1) << 8) | ((hdr >> 16) & 255)) + 1;
        run_bytes = 2 + (rl * (uint32_t)w + 7) / 8;
        if (cur + run_bytes > len) return ORCB_IO_ERROR;
        take = min(rl > skip ? rl - skip : 0u, room);
        const uint8_t* data = in + cur + 2;
        const uint32_t i_end = min(rl, skip + take);
        // four values per lane per step, all loads issued before the first store (memory-level parallelism)
        if (w <= 32 && (sg || nb == 8) && s.out_kind <= OUT_LEN31 && !(sg && s.out_kind == OUT_LEN31)) {
            const uint32_t first = skip + lane;
            const uint64_t o = out_pos + lane;
            switch (s.out_kind) {
                case OUT_I16:
                    if (sg) direct32_lane_loop<int16_t, true, false>(data, w, first, i_end, (int16_t*)s.out + o, c.err, s.colstripe, s.aux);
                    else direct32_lane_loop<int16_t, false, false>(data, w, first, i_end, (int16_t*)s.out + o, c.err, s.colstripe, s.aux);
                    break;
                case OUT_I32:
                    if (sg) direct32_lane_loop<int32_t, true, false>(data, w, first, i_end, (int32_t*)s.out + o, c.err, s.colstripe, s.aux);
                    else direct32_lane_loop<int32_t, false, false>(data, w, first, i_end, (int32_t*)s.out + o, c.err, s.colstripe, s.aux);
                    break;
                case OUT_I64:
                    if (sg) direct32_lane_loop<int64_t, true, false>(data, w, first, i_end, (int64_t*)s.out + o, c.err, s.colstripe, s.aux);
                    else direct32_lane_loop<int64_t, false, false>(data, w, first, i_end, (int64_t*)s.out + o, c.err, s.colstripe, s.aux);
                    break;
                default:  // OUT_LEN31: unsigned lengths / keys
                    if (sg) direct32_lane_loop<int32_t, true, true>(data, w, first, i_end, (int32_t*)s.out + o, c.err, s.colstripe, s.aux);
                    else direct32_lane_loop<int32_t, false, true>(data, w, first, i_end, (int32_t*)s.out + o, c.err, s.colstripe, s.aux);
                    break;
            }
        } else {
            for (uint32_t i0 = skip + lane; i0 < i_end; i0 += 128) {
                uint64_t raw[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + 32u * u;
                    raw[u] = i < i_end ? load_be_bits(data, i * (uint32_t)w, w) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const uint32_t i = i0 + 32u * u;
                    if (i < i_end) {
                        int64_t v = trunc_n((int64_t)raw[u], nb);
                        if (sg) v = zigzag_n(v, nb);
                        store_val(c, out_pos + (i - skip), v);
                    }
                }
            }
        }
    } else if (kind == 3) {
        // DELTA delta.rs:44-116
        if (cur + 2 > len) return ORCB_IO_ERROR;
        const uint32_t code = (h0 >> 1) & 31;
        const int w = code == 0 ? 0 : width_of(code);
        rl = (((h0 & 1) << 8) | ((hdr >> 16) & 255)) + 1;
        uint32_t p = cur + 2;
        uint64_t ub, ud;
        uint32_t e = parse_varint(in, p, len, nb * 8, ub);
        if (e) return e;
        int64_t base = trunc_n((int64_t)ub, nb);
        if (sg) base = zigzag_n(base, nb);
        e = parse_varint(in, p, len, 64, ud);
        if (e) return e;
        const int64_t d0 = zigzag_n((int64_t)ud, 8);
        // d0 <= 0: base - |d0| == base + d0 (is_positive() is false for 0, delta.rs:77-82);
        // |i64::MIN| wraps to i64::MIN in the reference, so subtracting it moves by +2^63
        const __int128 step = d0 == INT64_MIN ? ((__int128)1 << 63) : (__int128)d0;
        const bool positive = d0 > 0;
        if (w == 0) {
            run_bytes = p - cur;
            const __int128 last = (__int128)base + (__int128)(rl - 1) * step;
            if (!in_range_n(last, nb)) return ORCB_OUT_OF_SPEC;
            take = min(rl > skip ? rl - skip : 0u, room);
            const uint32_t i_end = min(rl, skip + take);
            const uint64_t ustep = (uint64_t)(int64_t)step;
            for (uint32_t i = skip + lane; i < i_end; i += 32)
                store_val(c, out_pos + (i - skip), (int64_t)((uint64_t)base + (uint64_t)i * ustep));
        } else {
            if (rl < 2) return ORCB_IO_ERROR;
            const uint32_t nd = rl - 2;
            run_bytes = (p - cur) + (nd * (uint32_t)w + 7) / 8;
            if (cur + run_bytes > len) return ORCB_IO_ERROR;
            const __int128 second = (__int128)base + step;
            if (!in_range_n(second, nb)) return ORCB_OUT_OF_SPEC;
            take = min(rl > skip ? rl - skip : 0u, room);
            EMIT(lane == 0 ? 0u : 0xffffffffu, base);
            EMIT(lane == 1 ? 1u : 0xffffffffu, (int64_t)second);
            const uint8_t* data = in + p;
            if (w == 64) {
                // deltas are i64 here and may be negative: exact sequential semantics
                __int128 acc = second;
                for (uint32_t i = 0; i < nd; i++) {
                    const int64_t d = (int64_t)load_be_bits(data, i * 64u, 64);
                    acc = positive ? acc + (__int128)d : acc - (__int128)d;
                    if (!in_range_n(acc, nb)) return ORCB_OUT_OF_SPEC;
                    EMIT((i & 31) == (uint32_t)lane ? i + 2 : 0xffffffffu, (int64_t)acc);
                }
            } else {
                // monotone run: wrapping prefix sums are exact iff the final value is in range
                uint64_t carry = 0, tot_lo = 0, tot_hi = 0;
                const uint64_t sec = (uint64_t)(int64_t)second;
                for (uint32_t i0 = 0; i0 < nd; i0 += 32) {
                    const uint32_t i = i0 + lane;
                    const uint64_t d = i < nd ? load_be_bits(data, i * (uint32_t)w, w) : 0ull;
                    tot_lo += d & 0xffffffffull;
                    tot_hi += d >> 32;
                    const uint64_t pre = warp_incl_scan64(d, lane) + carry;
                    if (i < nd) {
                        const uint64_t v = positive ? sec + pre : sec - pre;
                        EMIT(i + 2, (int64_t)v);
                    }
                    carry = __shfl_sync(FULL, pre, 31);
                }
                tot_lo = warp_sum64(tot_lo);
                tot_hi = warp_sum64(tot_hi);
                const __int128 total = ((__int128)tot_hi << 32) + (__int128)tot_lo;
                const __int128 fin = positive ? second + total : second - total;
                if (!in_range_n(fin, nb)) return ORCB_OUT_OF_SPEC;
            }
        }
    } else {
        // PATCHED_BASE patched_base.rs:38-151
        if (cur + 4 > len) return ORCB_IO_ERROR;
        const int w = width_of((h0 >> 1) & 31);
        rl = (((h0 & 1) << 8) | ((hdr >> 16) & 255)) + 1;
        const uint32_t b3 = (hdr >> 8) & 255, b4 = hdr & 255;
        const int base_bw = (int)((b3 >> 5) & 7) + 1;
        const int pw = width_of(b3 & 31);
        const int pgw = (int)((b4 >> 5) & 7) + 1;
        if (pw + pgw > 64) return ORCB_OUT_OF_SPEC;
        const uint32_t pll = b4 & 31;
        const int cfb = closest_fixed_bits(pw + pgw);
        const uint32_t data_off = cur + 4 + base_bw;
        const uint32_t data_bytes = (rl * (uint32_t)w + 7) / 8;
        run_bytes = 4 + base_bw + data_bytes + (pll * (uint32_t)cfb + 7) / 8;
        if (cur + run_bytes > len) return ORCB_IO_ERROR;
        // value width wider than N: the reference panics or silently truncates; reported as OutOfSpec
        if (nb * 8 < w || pll == 0) return ORCB_OUT_OF_SPEC;
        const uint64_t ubase = load_be_bits(in + cur + 4, 0, base_bw * 8);
        int64_t base = (int64_t)ubase;
        if (sg) {  // signed_msb_decode util.rs:559-569
            const uint64_t msb = 1ull << (base_bw * 8 - 1);
            base = (ubase & msb) ? (int64_t)(0ull - (ubase & ~msb)) : (int64_t)(ubase & ~msb);
        }
        base = trunc_n(base, nb);
        const uint8_t* data = in + data_off;
        const uint8_t* pdata = data + data_bytes;
        // one lane per patch-list entry
        uint64_t pe = 0;
        if ((uint32_t)lane < pll) pe = load_be_bits(pdata, (uint32_t)lane * (uint32_t)cfb, cfb);
        const uint64_t pmask = (1ull << pw) - 1;  // pw <= 63 here
        const uint64_t gap = pe >> pw;
        const uint64_t patch = pe & pmask;
        const bool live = (uint32_t)lane < pll;
        const bool ext = live && gap == 255 && patch == 0;
        const uint32_t pos = warp_incl_scan(live ? (uint32_t)gap : 0u, lane);
        const uint32_t extmask = __ballot_sync(FULL, ext);
        const bool prev_nonext = lane > 0 && !((extmask >> (lane - 1)) & 1);
        const bool bad = live && !ext && gap == 0 && lane > 0 && prev_nonext;
        const uint32_t badmask = __ballot_sync(FULL, bad);
        const uint32_t first_bad = badmask ? (uint32_t)__ffs(badmask) - 1 : 32u;
        const bool applied = live && !ext && (uint32_t)lane < first_bad && pos < rl;
        const uint32_t appmask = __ballot_sync(FULL, applied);
        // trailing gap-extension entries index past the patch list in the reference (panic)
        if ((extmask >> (pll - 1)) & 1) {
            const uint32_t nonext = ~extmask & (pll >= 32 ? FULL : ((1u << pll) - 1));
            const bool reached = nonext == 0 || ((appmask >> (31 - __clz(nonext))) & 1);
            if (reached) return ORCB_OUT_OF_SPEC;
        }
        if (lane < 16) patchmap[lane] = 0;
        __syncwarp();
        if (applied) atomicOr(&patchmap[pos >> 5], 1u << (pos & 31));
        __syncwarp();
        take = min(rl > skip ? rl - skip : 0u, room);
        bool ovf = false;
        for (uint32_t i = lane; i < rl; i += 32) {
            if ((patchmap[i >> 5] >> (i & 31)) & 1) continue;
            const int64_t raw = trunc_n((int64_t)load_be_bits(data, i * (uint32_t)w, w), nb);
            const __int128 sum = (__int128)raw + (__int128)base;
            if (!in_range_n(sum, nb)) ovf = true;  // checked_add :144-146
            EMIT(i, (int64_t)sum);
        }
        if (applied && w >= 64) ovf = true;  // checked_shl(64) -> None :112-117
        if (__any_sync(FULL, ovf)) return ORCB_OUT_OF_SPEC;
        if (applied) {
            const int64_t raw = trunc_n((int64_t)load_be_bits(data, pos * (uint32_t)w, w), nb);
            const int64_t pbits = trunc_n((int64_t)(patch << w), nb);
            const int64_t v = trunc_n((int64_t)((uint64_t)(raw | pbits) + (uint64_t)base), nb);  // wrapping_add :122-124
            EMIT(pos, v);
        }
        __syncwarp();
    }
    rl_out = rl;
    bytes_out = run_bytes;
    take_out = take;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Integer RLE, third design: every lane of a warp owns one (stream, row-group) segment and only PARSES
// its next run header (scalar code, constant cost per run); the values of the parsed runs are then
// produced by all 32 lanes together, one value per lane per step, whichever run they belong to.
// Runs that need a scan or a patch list (DELTA with packed deltas, PATCHED_BASE) are decoded one at a
// time by the whole warp (coop_run2).
// ------------------------------------------------------------------------------------------------
enum RunClass : uint32_t { RC_NONE = 0, RC_CONST = 1, RC_DIRECT = 2, RC_COOP = 3 };
constexpr uint32_t COOP_MIN_RUN = 96;  // DIRECT runs at least this long are decoded by the whole warp
constexpr uint32_t TILE_VALUES = 512;   // values of one 32-run block staged in shared memory (fast block path)
constexpr uint32_t SELF_FILL = 10;      // runs up to this long (every SHORT_REPEAT) are expanded by their own lane

struct RunSlot {       // one per lane, in shared memory
    uint64_t base;     // RC_CONST: value at k = 0
    uint64_t step;     // RC_CONST: value(k) = base + k * step
    uint64_t data;     // RC_DIRECT: packed values
    uint64_t out;      // destination buffer
    uint64_t out_idx;  // element index of the first emitted value
    uint32_t skip;     // first k emitted
    uint32_t meta;     // w | cls << 8 | out_kind << 12 | nbytes << 16 | signed << 24
    uint32_t prefix;   // inclusive prefix sum of emitted counts over the lanes
    uint32_t colstripe;
    uint32_t aux;
    uint32_t pad;
};

// up to 8 varint bytes (little-endian in x, continuation bits ignored) -> the 7-bit groups squeezed together
__device__ __forceinline__ uint64_t squeeze7(uint64_t x) {
    x &= 0x7f7f7f7f7f7f7f7full;
    x = (x & 0x007f007f007f007full) | ((x & 0x7f007f007f007f00ull) >> 1);
    x = (x & 0x00003fff00003fffull) | ((x & 0x3fff00003fff0000ull) >> 2);
    x = (x & 0x000000000fffffffull) | ((x & 0x0fffffff00000000ull) >> 4);
    return x;
}

// Rare DELTA headers (varints longer than the 8-byte window, or arithmetic that needs 128 bits): kept out of line
// so the hot kernel stays small.
__device__ __noinline__ uint32_t parse_delta_slow(const uint8_t* in, uint32_t len, uint32_t cur, int nb, bool sg, uint32_t rl,
                                                  uint64_t* base_out, uint64_t* step_out, uint32_t* bytes_out) {
    uint32_t p = cur + 2;
    uint64_t ub, ud;
    uint32_t e = parse_varint(in, p, len, nb * 8, ub);
    if (e) return e;
    e = parse_varint(in, p, len, 64, ud);
    if (e) return e;
    int64_t base = trunc_n((int64_t)ub, nb);
    if (sg) base = zigzag_n(base, nb);
    const int64_t d0 = zigzag_n((int64_t)ud, 8);
    // d0 <= 0: base - |d0| == base + d0 (is_positive() is false for 0, delta.rs:77-82);
    // |i64::MIN| wraps to i64::MIN in the reference, so subtracting it moves by +2^63
    const __int128 step = d0 == INT64_MIN ? ((__int128)1 << 63) : (__int128)d0;
    const __int128 last = (__int128)base + (__int128)(rl - 1) * step;
    if (!in_range_n(last, nb)) return ORCB_OUT_OF_SPEC;
    *base_out = (uint64_t)base;
    *step_out = (uint64_t)(int64_t)step;
    *bytes_out = p - cur;
    return 0;
}

// Parse the run at `cur` of the lane's own segment.  No values are produced here (except RLE v1 literals).
// One 8-byte window of the stream serves the common headers without byte loops.
__device__ __forceinline__ uint32_t parse_run2(const Seg& s, uint32_t cur, RunSlot& d, uint32_t& cls, uint32_t& rl_out,
                                               uint32_t& bytes_out) {
    const uint8_t* in = (const uint8_t*)s.in;
    const uint32_t len = s.in_len;
    const int nb = s.nbytes;
    const bool sg = (s.flags & SEG_SIGNED) != 0;
    const uintptr_t ai = (uintptr_t)(in + cur);
    const uint32_t* q = (const uint32_t*)(ai & ~(uintptr_t)3);
    const uint32_t shb = (uint32_t)(ai & 3) * 8;
    const uint32_t q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    const uint32_t lo = __funnelshift_r(q0, q1, shb);  // stream bytes 0..3, byte 0 in the low bits
    const uint32_t hi = __funnelshift_r(q1, q2, shb);  // stream bytes 4..7
    const uint64_t win = ((uint64_t)hi << 32) | lo;
    const uint32_t h0 = lo & 255;
    const uint32_t kind = h0 >> 6;
    if (kind == 0) {
        // SHORT_REPEAT short_repeat.rs:29-63
        const int bw = (int)((h0 >> 3) & 7) + 1;
        if (nb < bw) return ORCB_OUT_OF_SPEC;
        rl_out = (h0 & 7) + 3;
        bytes_out = 1 + bw;
        if (cur + bytes_out > len) return ORCB_IO_ERROR;
        uint64_t raw;
        if (bw <= 7) {
            // big-endian value of stream bytes 1..bw: byte-reverse the window past the header
            const uint64_t t = win >> 8;
            const uint64_t rev = ((uint64_t)bswap32((uint32_t)t) << 32) | bswap32((uint32_t)(t >> 32));
            raw = rev >> (64 - 8 * bw);
        } else {
            raw = load_be_bits(in + cur + 1, 0, 64);
        }
        int64_t v = trunc_n((int64_t)raw, nb);
        if (sg) v = zigzag_n(v, nb);
        d.base = (uint64_t)v;
        d.step = 0;
        cls = RC_CONST;
        return 0;
    }
    if (kind == 2) {
        cls = RC_COOP;
        return 0;
    }
    if (cur + 2 > len) return ORCB_IO_ERROR;
    const uint32_t rl = (((h0 & 1) << 8) | ((lo >> 8) & 255)) + 1;
    const uint32_t code = (h0 >> 1) & 31;
    rl_out = rl;
    if (kind == 1) {
        // DIRECT direct.rs:39-65
        const int w = width_of(code);
        if (nb * 8 < w) return ORCB_OUT_OF_SPEC;
        bytes_out = 2 + (rl * (uint32_t)w + 7) / 8;
        if (cur + bytes_out > len) return ORCB_IO_ERROR;
        if (rl >= COOP_MIN_RUN) {
            cls = RC_COOP;  // long run: the constant-phase warp loop of coop_run2 is cheaper per value
            return 0;
        }
        d.data = (uint64_t)(uintptr_t)(in + cur + 2);
        d.meta = (uint32_t)w;
        cls = RC_DIRECT;
        return 0;
    }
    // DELTA delta.rs:44-116
    if (code != 0) {
        cls = RC_COOP;  // packed deltas need a prefix sum
        return 0;
    }
    // both varints inside the window (terminators among stream bytes 2..7)?
    uint64_t term = ~win & 0x8080808080800000ull;
    const int t1 = __ffsll((long long)term);
    term &= term - 1;
    const int t2 = __ffsll((long long)term);
    const uint32_t e1 = (uint32_t)(t1 >> 3), e2 = (uint32_t)(t2 >> 3);  // byte index after each varint
    bool slow = !(t1 && t2);
    int64_t base = 0, d0 = 0;
    uint32_t p = cur + e2;
    if (!slow) {
        const uint64_t x = win >> 16;  // bytes 2..7
        const uint32_t n1 = e1 - 2, n2 = e2 - e1;
        // read_varint::<N>: a byte at shift >= bit-width(N) is an error even when zero (util.rs:486-489)
        if ((n1 - 1) * 7 >= (uint32_t)nb * 8) return ORCB_VARINT_TOO_LARGE;
        if (p > len) return ORCB_IO_ERROR;
        base = trunc_n((int64_t)squeeze7(x & ((1ull << (8 * n1)) - 1)), nb);
        if (sg) base = zigzag_n(base, nb);
        d0 = zigzag_n((int64_t)squeeze7((x >> (8 * n1)) & ((1ull << (8 * n2)) - 1)), 8);
        // no i64 overflow possible below these bounds (run length <= 512): plain 64-bit arithmetic
        slow = !(d0 > -(1ll << 40) && d0 < (1ll << 40) && base > -(1ll << 62) && base < (1ll << 62));
    }
    if (slow) {
        cls = RC_CONST;
        // results come back through locals so that the caller's RunSlot can stay in registers
        uint64_t sb = 0, ss = 0;
        uint32_t sbytes = 0;
        const uint32_t st = parse_delta_slow(in, len, cur, nb, sg, rl, &sb, &ss, &sbytes);
        d.base = sb;
        d.step = ss;
        bytes_out = sbytes;
        return st;
    }
    const int64_t last = base + (int64_t)(rl - 1) * d0;
    if (trunc_n(last, nb) != last) return ORCB_OUT_OF_SPEC;
    d.step = (uint64_t)d0;
    bytes_out = p - cur;
    d.base = (uint64_t)base;
    cls = RC_CONST;
    return 0;
}

// RLE v1 (integer/rle_v1.rs:54-68, 90-159).  Runs become RC_CONST; literal groups are decoded right here
// by the owning lane (legacy format, not worth a cooperative path).
__device__ __forceinline__ uint32_t parse_run1(const SegCtx& c, uint32_t cur, uint32_t skip, uint32_t room, uint64_t out_pos,
                                               RunSlot& d, uint32_t& cls, uint32_t& rl_out, uint32_t& bytes_out,
                                               uint32_t& take_out) {
    const Seg& s = *c.s;
    const uint8_t* in = (const uint8_t*)s.in;
    const uint32_t len = s.in_len;
    const int nb = s.nbytes;
    const bool sg = (s.flags & SEG_SIGNED) != 0;
    const int8_t h = (int8_t)in[cur];
    uint32_t p = cur + 1;
    if (h < 0) {
        const uint32_t rl = (uint32_t)(-(int)h);
        const uint32_t take = min(rl > skip ? rl - skip : 0u, room);
        for (uint32_t i = 0; i < rl; i++) {
            uint64_t u;
            const uint32_t e = parse_varint(in, p, len, nb * 8, u);
            if (e) return e;
            int64_t v = trunc_n((int64_t)u, nb);
            if (sg) v = zigzag_n(v, nb);
            EMIT(i, v);
        }
        rl_out = rl;
        bytes_out = p - cur;
        take_out = take;
        cls = RC_NONE;  // already emitted
        return 0;
    }
    const uint32_t rl = (uint32_t)(uint8_t)h + 3;
    if (p >= len) return ORCB_IO_ERROR;
    const int delta = (int)(int8_t)in[p++];
    uint64_t u;
    const uint32_t e = parse_varint(in, p, len, nb * 8, u);
    if (e) return e;
    int64_t base = trunc_n((int64_t)u, nb);
    if (sg) base = zigzag_n(base, nb);
    const __int128 last = (__int128)base + (__int128)(rl - 1) * (__int128)delta;
    if (!in_range_n(last, nb)) return ORCB_OUT_OF_SPEC;
    d.base = (uint64_t)base;
    d.step = (uint64_t)(int64_t)delta;
    rl_out = rl;
    bytes_out = p - cur;
    cls = RC_CONST;
    return 0;
}


// Header-only walk: length in values and bytes of the run at `cur` (no values produced).
// KNOWN_V2: the caller has already checked the segment's RLE version (hot loop of the run index).
template <bool KNOWN_V2 = false>
__device__ __forceinline__ uint32_t measure_run(const Seg& s, uint32_t cur, uint32_t& rl_out, uint32_t& bytes_out,
                                                bool& coop) {
    coop = false;
    const uint8_t* in = (const uint8_t*)s.in;
    const uint32_t len = s.in_len;
    if (!KNOWN_V2 && !(s.flags & SEG_RLE_V2)) {
        // RLE v1 (integer/rle_v1.rs:54-68)
        const int8_t h = (int8_t)in[cur];
        uint32_t p = cur + 1;
        uint32_t nvar;
        if (h < 0) {
            rl_out = (uint32_t)(-(int)h);
            nvar = rl_out;
        } else {
            rl_out = (uint32_t)(uint8_t)h + 3;
            p += 1;  // delta byte
            nvar = 1;
        }
        for (uint32_t i = 0; i < nvar; i++) {
            for (;;) {
                if (p >= len) return ORCB_IO_ERROR;
                if (!(in[p++] & 0x80)) break;
            }
        }
        bytes_out = p - cur;
        return 0;
    }
    // One 8-byte window serves every common header: SHORT_REPEAT / DIRECT need 2 bytes, DELTA needs the
    // two varints that follow (found with a continuation-bit mask instead of a byte loop).
    const uintptr_t ai = (uintptr_t)(in + cur);
    const uint32_t* q = (const uint32_t*)(ai & ~(uintptr_t)3);
    const uint32_t shb = (uint32_t)(ai & 3) * 8;
    const uint32_t q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    const uint32_t lo = __funnelshift_r(q0, q1, shb);  // stream bytes 0..3 (little-endian lanes)
    const uint32_t hi = __funnelshift_r(q1, q2, shb);  // stream bytes 4..7
    const uint32_t h0 = lo & 255, b1 = (lo >> 8) & 255;
    const uint32_t kind = h0 >> 6;
    const uint32_t code = (h0 >> 1) & 31;
    const uint32_t rl = (((h0 & 1) << 8) | b1) + 1;
    const uint32_t w = (uint32_t)width_of(code);
    rl_out = kind == 0 ? (h0 & 7) + 3 : rl;
    bytes_out = kind == 0 ? 2 + ((h0 >> 3) & 7) : 2 + (rl * w + 7) / 8;
    coop = kind == 2 || (kind == 1 && rl >= COOP_MIN_RUN) || (kind == 3 && code != 0);
    if (kind >= 2) {
        if (kind == 2) {
            const uint32_t b3 = (lo >> 16) & 255, b4 = lo >> 24;
            const int pw = width_of(b3 & 31), pgw = (int)((b4 >> 5) & 7) + 1;
            if (pw + pgw > 64) return ORCB_OUT_OF_SPEC;
            bytes_out = 4 + ((b3 >> 5) & 7) + 1 + (rl * w + 7) / 8 + ((b4 & 31) * (uint32_t)closest_fixed_bits(pw + pgw) + 7) / 8;
        } else {
            // terminator bytes (bit 7 clear) among stream bytes 2..7
            const uint64_t win = ((uint64_t)hi << 32) | lo;
            uint64_t term = ~win & 0x8080808080800000ull;
            uint32_t p;
            const int t1 = __ffsll((long long)term);  // 1-based bit index of the first terminator's bit 7
            term &= term - 1;
            const int t2 = __ffsll((long long)term);
            if (t1 && t2) {
                p = cur + (uint32_t)(t2 >> 3);  // byte after the second varint
            } else {
                p = cur + 2;
                for (int i = 0; i < 2; i++) {
                    for (;;) {
                        if (p >= len) return ORCB_IO_ERROR;
                        if (!(in[p++] & 0x80)) break;
                    }
                }
            }
            if (code != 0) {
                if (rl < 2) return ORCB_IO_ERROR;
                p += ((rl - 2) * w + 7) / 8;
            }
            bytes_out = p - cur;
        }
    }
    if (cur + bytes_out > len) return ORCB_IO_ERROR;
    return 0;
}

constexpr uint32_t IDX_LANES = 8;

// Pre-pass ("device-built run index"): one lane per segment walks the run headers only and records where
// every run starts and where its values go.  This is the only serial chain of the integer path (a run's
// position depends on all runs before it); it is a few dozen instructions per run and produces no values,
// so the decode proper (k_int_rle) runs as one fully parallel step of one run per lane.
__global__ void __launch_bounds__(128) k_rle_index(const Seg* __restrict__ segs, uint32_t nseg,
                                                   const uint32_t* __restrict__ cnt, RunRec* __restrict__ table,
                                                   BlockRec* __restrict__ blocks, uint32_t* __restrict__ nblocks,
                                                   uint32_t pool_blocks, CoopRec* __restrict__ coop_q,
                                                   uint32_t* __restrict__ ncoop, uint32_t coop_cap, uint32_t* err) {
    // only IDX_LANES lanes of each warp own a segment: a warp advances at the pace of its slowest lane
    // (the one that misses L1 this step), so fewer streams per warp and more warps hide more latency
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    if ((gt & 31) >= IDX_LANES) return;
    const uint32_t segi = (gt >> 5) * IDX_LANES + (gt & 31);
    if (segi >= nseg) return;
    const Seg& s = segs[segi];
    const uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_values;
    uint32_t cur = s.start_byte, skip = s.run_skip, produced = 0;
    // a row-index position may skip more values than its first run holds (writers record positions while
    // values are still buffered): step over the runs that are skipped entirely
    while (skip > 0 && cur < s.in_len) {
        uint32_t rl, nbytes;
        bool cq;
        if (measure_run(s, cur, rl, nbytes, cq) || skip < rl) break;
        skip -= rl;
        cur += nbytes;
    }
    uint32_t blk = 0xffffffffu, in_blk = 0, blk_skip = skip;
    const bool v2 = (s.flags & SEG_RLE_V2) != 0;
    const uint32_t len = s.in_len;
    RunRec* slot = nullptr;
    // every instruction of this loop sits on the segment's serial chain: keep it short
    while (produced < n) {
        if (in_blk == 0) {
            blk = atomicAdd(nblocks, 1u);
            if (blk >= pool_blocks) { set_err(err, s.colstripe, ORCB_UNEXPECTED); blk = 0xffffffffu; break; }
            slot = table + (uint64_t)blk * 32;
        }
        bool stop = cur >= len;  // k_int_rle reports "not enough values" for this record
        uint32_t rl = 0, nbytes = 0;
        bool cq = false;
        if (!stop) {
            // k_int_rle re-parses a run that does not parse here and reports the error
            if (v2) stop = measure_run<true>(s, cur, rl, nbytes, cq) != 0;
            else stop = measure_run<false>(s, cur, rl, nbytes, cq) != 0;
        }
        RunRec r;
        r.byte_off = cur;
        r.out_off = produced;
        if (cq && !stop) {
            // whole-warp runs go to their own queue; the block table keeps a placeholder so positions stay aligned
            const uint32_t qi = atomicAdd(ncoop, 1u);
            if (qi < coop_cap) {
                CoopRec cr;
                cr.seg = segi;
                cr.byte_off = cur;
                cr.out_off = produced;
                cr.skip = skip;
                coop_q[qi] = cr;
                r.out_off |= RUN_QUEUED;
            }
        }
        *slot++ = r;
        in_blk++;
        if (!stop) {
            produced += min(rl - min(skip, rl), n - produced);
            skip = 0;
            cur += nbytes;
        }
        if (in_blk == 32 || stop || produced >= n) {
            BlockRec br;
            br.seg = segi;
            br.n_runs = in_blk;
            br.skip = blk_skip;
            br.pad = 0;
            blocks[blk] = br;
            in_blk = 0;
            blk_skip = 0;
            if (stop) break;
        }
    }
}

constexpr int RLE_WARPS = 4;

// Integer RLE decode proper: one warp per 32 consecutive runs of one segment (run table from k_rle_index).
// One block of up to 32 consecutive runs of one segment, one run per lane.
// FAST = true: only blocks made of constant / short direct runs are decoded (shared-memory tile, coalesced
// flush); anything else returns false and is queued for the FAST = false instantiation (general blocks:
// scan + per-value search, whole-warp runs, RLE v1, error reporting).  Two kernels keep the hot one small
// enough for the instruction cache.
template <bool FAST>
__device__ __forceinline__ bool int_rle_block(const Seg* __restrict__ segs, const BlockRec br, const RunRec* __restrict__ recs,
                                              const uint32_t* __restrict__ cnt, const uint32_t* __restrict__ dstart,
                                              uint32_t* err, uint32_t* mis, uint32_t* patchmap, RunSlot* slots,
                                              int64_t* tile, const int lane) {
    SegCtx c;
    c.s = &segs[br.seg];
    c.err = err;
    c.mis = mis;
    const Seg& s = *c.s;
    bool active = (uint32_t)lane < br.n_runs;
    const uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_values;
    const uint64_t obase = s.start_idx >= 0 ? dstart[s.start_idx] : s.out_start;
    const bool v2 = (s.flags & SEG_RLE_V2) != 0;
    if (FAST && !v2) return false;
    RunRec rec;
    rec.byte_off = 0;
    rec.out_off = 0;
    if (active) rec = recs[lane];
    const bool queued = (rec.out_off & RUN_QUEUED) != 0;  // decoded by k_coop_runs
    rec.out_off &= ~RUN_QUEUED;
    if (queued) active = false;
    const uint32_t cur = rec.byte_off;
    const uint32_t skip = lane == 0 ? br.skip : 0u;
    const uint32_t room = n - rec.out_off;
    const uint64_t out_pos = obase + rec.out_off;
    __syncwarp();
    // the fast kernel keeps its run in registers; general blocks publish theirs for the per-value search
    RunSlot my_reg;
    RunSlot& my = FAST ? my_reg : slots[lane];
    uint32_t cls = RC_NONE, rl = 0, nbytes = 0, take = 0;
    bool failed = false;
    if (active) {
        uint32_t st;
        my.meta = 0;
        if (cur >= s.in_len) st = ORCB_OUT_OF_SPEC;  // "not enough values to decode" rle_v2/mod.rs:115-122
        else if (FAST || v2) st = parse_run2(s, cur, my, cls, rl, nbytes);
        else st = parse_run1(c, cur, skip, room, out_pos, my, cls, rl, nbytes, take);
        if (st) {
            if (FAST) failed = true;  // the general kernel re-parses the block and reports the error
            else set_err(err, s.colstripe, st);
            active = false;
            cls = RC_NONE;
        }
        if (cls == RC_CONST || cls == RC_DIRECT) {
            take = min(rl > skip ? rl - skip : 0u, room);
            my.out_idx = out_pos;
            my.skip = skip;
            my.meta = (my.meta & 0xffu) | (cls << 8);
        }
    }
    const uint32_t emit = (cls == RC_CONST || cls == RC_DIRECT) ? take : 0u;
    const uint32_t incl = warp_incl_scan(emit, lane);
    if (!FAST) my.prefix = incl;
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    const int nb = s.nbytes;
    const bool sg = (s.flags & SEG_SIGNED) != 0;
    const uint32_t okind = s.out_kind;
    // ---- fast block: every run of the block is a short constant / direct run.  The 32 runs are consecutive
    //      runs of one segment, so their values are contiguous in the output: each lane expands its own run
    //      into a shared-memory tile, then the warp writes the tile with coalesced stores.
    if (FAST) {
      if (__any_sync(FULL, failed)) return false;
      if (!(total <= TILE_VALUES && __all_sync(FULL, !active || cls == RC_CONST || cls == RC_DIRECT))) return false;
      {
        const uint32_t pe = incl - emit;
        const int w = (int)(my.meta & 0xff);
        const uint8_t* data = (const uint8_t*)(uintptr_t)my.data;
        const bool narrow = w <= 32 && (nb > 2 || w <= 16);
        // ---- outputs of at most 32 bits (dictionary keys, lengths, INT / DATE / SHORT, decimal scales): the same
        //      block in 32-bit arithmetic with a 32-bit tile; range checks are done once per constant run
        if (okind != OUT_I64 && __all_sync(FULL, cls != RC_DIRECT || w <= 32)) {
            uint32_t* tile32 = (uint32_t*)tile;
            bool bad = false;
            if (cls == RC_CONST) {
                const uint64_t first = my.base + (uint64_t)skip * my.step;
                const uint64_t last = first + (uint64_t)(take ? take - 1 : 0) * my.step;
                if (okind == OUT_LEN31) bad = take && (first > 0x7fffffffull || last > 0x7fffffffull);
                else if (okind == OUT_SCALE) bad = take && ((uint32_t)first != s.aux || (take > 1 && my.step != 0));
            }
            if (take <= SELF_FILL) {
                if (cls == RC_CONST) {
                    uint32_t v = (uint32_t)my.base + skip * (uint32_t)my.step;
                    for (uint32_t j = 0; j < take; j++, v += (uint32_t)my.step) tile32[pe + j] = v;
                } else if (cls == RC_DIRECT) {
                    for (uint32_t j = 0; j < take; j++) {
                        uint32_t x = load_be_bits32(data, (skip + j) * (uint32_t)w, w);
                        if (sg) x = (x >> 1) ^ (0u - (x & 1));
                        tile32[pe + j] = x;
                    }
                }
            }
            uint32_t longmask = __ballot_sync(FULL, (cls == RC_CONST || cls == RC_DIRECT) && take > SELF_FILL);
            while (longmask) {
                const int leader = __ffs(longmask) - 1;
                longmask &= longmask - 1;
                const uint32_t lcls = __shfl_sync(FULL, cls, leader);
                const uint32_t ltake = __shfl_sync(FULL, take, leader);
                const uint32_t lpe = __shfl_sync(FULL, pe, leader);
                const uint32_t lskip = __shfl_sync(FULL, skip, leader);
                if (lcls == RC_CONST) {
                    const uint32_t lbase = __shfl_sync(FULL, (uint32_t)my.base, leader);
                    const uint32_t lstep = __shfl_sync(FULL, (uint32_t)my.step, leader);
                    for (uint32_t j = lane; j < ltake; j += 32) tile32[lpe + j] = lbase + (lskip + j) * lstep;
                } else {
                    const int lw = __shfl_sync(FULL, w, leader);
                    const uint8_t* ldata = (const uint8_t*)(uintptr_t)__shfl_sync(FULL, (uint64_t)(uintptr_t)data, leader);
                    for (uint32_t j = lane; j < ltake; j += 32) {
                        uint32_t x = load_be_bits32(ldata, (lskip + j) * (uint32_t)lw, lw);
                        if (sg) x = (x >> 1) ^ (0u - (x & 1));
                        tile32[lpe + j] = x;
                    }
                }
            }
            __syncwarp();
            // DIRECT values of LEN31 / SCALE streams are checked on the way out (constant runs were checked above);
            // a value with bit 31 set is outside [0, 2^31) whether the stream is signed (negative) or not
            const uint32_t dmask = __ballot_sync(FULL, cls == RC_DIRECT && take > 0);
            // the block's values are contiguous in the output except where a queued whole-warp run sits:
            // flush the lanes between two queued lanes as one coalesced range
            const uint32_t qmask = __ballot_sync(FULL, queued);
            uint32_t pending = __ballot_sync(FULL, emit > 0);
            while (pending) {
                const int a = __ffs(pending) - 1;
                const uint32_t after = qmask & ~((2u << a) - 1u);
                const int stop = after ? __ffs(after) - 1 : 32;
                pending &= ~((stop >= 32 ? FULL : ((1u << stop) - 1u)) & ~((1u << a) - 1u));
                const uint32_t t0 = __shfl_sync(FULL, pe, a), t1 = __shfl_sync(FULL, incl, stop - 1);
                const uint64_t o0 = __shfl_sync(FULL, out_pos, a);
                if (okind == OUT_I16) {
                    for (uint32_t p = t0 + lane; p < t1; p += 32) ((int16_t*)s.out)[o0 + (p - t0)] = (int16_t)tile32[p];
                } else {
                    for (uint32_t p = t0 + lane; p < t1; p += 32) {
                        const uint32_t v = tile32[p];
                        if (dmask) {
                            if (okind == OUT_LEN31) bad |= (v >> 31) != 0;
                            else if (okind == OUT_SCALE) bad |= v != s.aux;
                        }
                        ((int32_t*)s.out)[o0 + (p - t0)] = (int32_t)v;
                    }
                }
            }
            if (bad) {
                if (okind == OUT_LEN31) set_err(err, s.colstripe, s.aux);
                else if (okind == OUT_SCALE) atomicOr(&mis[s.colstripe], 1u);
            }
            return true;
        }
        // runs of up to SELF_FILL values are expanded by their own lane ...
        if (take <= SELF_FILL) {
            if (cls == RC_CONST) {
                uint64_t v = my.base + (uint64_t)skip * my.step;
                for (uint32_t j = 0; j < take; j++, v += my.step) tile[pe + j] = (int64_t)v;
            } else if (cls == RC_DIRECT) {
                if (narrow) {
                    for (uint32_t j = 0; j < take; j++)
                        tile[pe + j] = finish32(load_be_bits32(data, (skip + j) * (uint32_t)w, w), sg, nb);
                } else {
                    for (uint32_t j = 0; j < take; j++) {
                        int64_t v = trunc_n((int64_t)load_be_bits(data, (skip + j) * (uint32_t)w, w), nb);
                        if (sg) v = zigzag_n(v, nb);
                        tile[pe + j] = v;
                    }
                }
            }
        }
        // ... longer ones by the whole warp, one run after the other
        uint32_t longmask = __ballot_sync(FULL, (cls == RC_CONST || cls == RC_DIRECT) && take > SELF_FILL);
        while (longmask) {
            const int leader = __ffs(longmask) - 1;
            longmask &= longmask - 1;
            const uint32_t lcls = __shfl_sync(FULL, cls, leader);
            const uint32_t ltake = __shfl_sync(FULL, take, leader);
            const uint32_t lpe = __shfl_sync(FULL, pe, leader);
            const uint32_t lskip = __shfl_sync(FULL, skip, leader);
            if (lcls == RC_CONST) {
                const uint64_t lbase = __shfl_sync(FULL, my.base, leader);
                const uint64_t lstep = __shfl_sync(FULL, my.step, leader);
                for (uint32_t j = lane; j < ltake; j += 32) tile[lpe + j] = (int64_t)(lbase + (uint64_t)(lskip + j) * lstep);
            } else {
                const int lw = __shfl_sync(FULL, w, leader);
                const uint8_t* ldata = (const uint8_t*)(uintptr_t)__shfl_sync(FULL, (uint64_t)(uintptr_t)data, leader);
                const bool lnarrow = lw <= 32 && (nb > 2 || lw <= 16);
                for (uint32_t j = lane; j < ltake; j += 32) {
                    int64_t v;
                    if (lnarrow) {
                        v = finish32(load_be_bits32(ldata, (lskip + j) * (uint32_t)lw, lw), sg, nb);
                    } else {
                        v = trunc_n((int64_t)load_be_bits(ldata, (lskip + j) * (uint32_t)lw, lw), nb);
                        if (sg) v = zigzag_n(v, nb);
                    }
                    tile[lpe + j] = v;
                }
            }
        }
        __syncwarp();
        bool bad = false;
        const uint32_t qmask = __ballot_sync(FULL, queued);
        uint32_t pending = __ballot_sync(FULL, emit > 0);
        while (pending) {
            const int a = __ffs(pending) - 1;
            const uint32_t after = qmask & ~((2u << a) - 1u);
            const int stop = after ? __ffs(after) - 1 : 32;
            pending &= ~((stop >= 32 ? FULL : ((1u << stop) - 1u)) & ~((1u << a) - 1u));
            const uint32_t t0 = __shfl_sync(FULL, pe, a), t1 = __shfl_sync(FULL, incl, stop - 1);
            const uint64_t o0 = __shfl_sync(FULL, out_pos, a);
            switch (okind) {
                case OUT_I16: for (uint32_t p = t0 + lane; p < t1; p += 32) ((int16_t*)s.out)[o0 + (p - t0)] = (int16_t)tile[p]; break;
                case OUT_I32: for (uint32_t p = t0 + lane; p < t1; p += 32) ((int32_t*)s.out)[o0 + (p - t0)] = (int32_t)tile[p]; break;
                case OUT_I64: for (uint32_t p = t0 + lane; p < t1; p += 32) ((int64_t*)s.out)[o0 + (p - t0)] = tile[p]; break;
                case OUT_LEN31:
                    for (uint32_t p = t0 + lane; p < t1; p += 32) {
                        const int64_t v = tile[p];
                        if ((uint64_t)v > 0x7fffffffull) bad = true;
                        ((int32_t*)s.out)[o0 + (p - t0)] = (int32_t)v;
                    }
                    break;
                case OUT_SCALE:
                    for (uint32_t p = t0 + lane; p < t1; p += 32) {
                        const int64_t v = tile[p];
                        if ((uint32_t)(int32_t)v != s.aux) bad = true;
                        ((int32_t*)s.out)[o0 + (p - t0)] = (int32_t)v;
                    }
                    break;
                default: break;
            }
        }
        if (bad) {
            if (okind == OUT_LEN31) set_err(err, s.colstripe, s.aux);
            else if (okind == OUT_SCALE) atomicOr(&mis[s.colstripe], 1u);
        }
        return true;
      }
    }
    __syncwarp();
    // ---- general block: all lanes produce the values of all parsed runs, 4 values per lane per step, every
    //      load issued before the first store.  out kind / N / signedness are per segment, hence warp-uniform.
    for (uint32_t v0 = lane; v0 < total; v0 += 128) {
        uint32_t li[4], jj[4];
        uint64_t raw[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t v = v0 + 32u * u;
            uint32_t l = 0;
            if (v < total) {
#pragma unroll
                for (int stp = 16; stp > 0; stp >>= 1)
                    if (slots[l + stp - 1].prefix <= v) l += stp;
            }
            li[u] = l;
            const RunSlot& d = slots[l];
            jj[u] = v - (l ? slots[l - 1].prefix : 0u);
            const uint32_t meta = d.meta;
            raw[u] = 0;
            if (v < total && (meta >> 8) == RC_DIRECT) {
                const int w = (int)(meta & 0xff);
                if (w <= 32 && (nb > 2 || w <= 16))
                    raw[u] = load_be_bits32((const uint8_t*)(uintptr_t)d.data, (d.skip + jj[u]) * (uint32_t)w, w);
                else
                    raw[u] = load_be_bits((const uint8_t*)(uintptr_t)d.data, (d.skip + jj[u]) * (uint32_t)w, w);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t v = v0 + 32u * u;
            if (v >= total) continue;
            const RunSlot& d = slots[li[u]];
            const uint32_t meta = d.meta;
            int64_t val;
            if ((meta >> 8) == RC_CONST) {
                val = (int64_t)(d.base + (uint64_t)(d.skip + jj[u]) * d.step);
            } else {
                const int w = (int)(meta & 0xff);
                if (w <= 32 && (nb > 2 || w <= 16)) {
                    val = finish32((uint32_t)raw[u], sg, nb);
                } else {
                    val = trunc_n((int64_t)raw[u], nb);
                    if (sg) val = zigzag_n(val, nb);
                }
            }
            const uint64_t idx = d.out_idx + jj[u];
            switch (okind) {
                case OUT_I16: ((int16_t*)s.out)[idx] = (int16_t)val; break;
                case OUT_I32: ((int32_t*)s.out)[idx] = (int32_t)val; break;
                case OUT_I64: ((int64_t*)s.out)[idx] = val; break;
                case OUT_LEN31:
                    if ((uint64_t)val > 0x7fffffffull) set_err(err, s.colstripe, s.aux);
                    ((int32_t*)s.out)[idx] = (int32_t)val;
                    break;
                case OUT_SCALE:
                    if ((uint32_t)(int32_t)val != s.aux) atomicOr(&mis[s.colstripe], 1u);
                    ((int32_t*)s.out)[idx] = (int32_t)val;
                    break;
                default: break;
            }
        }
    }
    __syncwarp();
    // ---- runs that need the whole warp (long DIRECT, DELTA with packed deltas, PATCHED_BASE)
    uint32_t bigmask = __ballot_sync(FULL, active && cls == RC_COOP);
    while (bigmask) {
        const int leader = __ffs(bigmask) - 1;
        bigmask &= bigmask - 1;
        const uint32_t lcur = __shfl_sync(FULL, cur, leader);
        const uint32_t lskip = __shfl_sync(FULL, skip, leader);
        const uint32_t lroom = __shfl_sync(FULL, room, leader);
        const uint64_t lout = __shfl_sync(FULL, out_pos, leader);
        uint32_t crl = 0, cbytes = 0, ctake = 0;
        const uint32_t st = coop_run2(c, lcur, lskip, lroom, lout, patchmap, crl, cbytes, ctake);
        if (st && lane == leader) set_err(err, s.colstripe, st);
    }
    return true;
}

// Hot kernel: persistent warps over the run blocks; blocks it cannot take are appended to `slow_list`.
__global__ void __launch_bounds__(RLE_WARPS * 32) k_int_rle(const Seg* __restrict__ segs,
                                                            const BlockRec* __restrict__ blocks,
                                                            const uint32_t* __restrict__ nblocks_ptr,
                                                            const RunRec* __restrict__ table,
                                                            const uint32_t* __restrict__ cnt,
                                                            const uint32_t* __restrict__ dstart, uint32_t* err,
                                                            uint32_t* mis, uint32_t* slow_list, uint32_t* slow_count) {
    __shared__ int64_t tile_all[RLE_WARPS][TILE_VALUES];
    const uint32_t nblocks = *nblocks_ptr;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    RunSlot* slots = nullptr;
    // persistent warps: the number of run blocks is only known on the device
    for (uint32_t blk = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; blk < nblocks; blk += nwarps) {
        const bool done = int_rle_block<true>(segs, blocks[blk], table + (uint64_t)blk * 32, cnt, dstart, err, mis, nullptr,
                                              slots, tile_all[threadIdx.x >> 5], lane);
        if (!done && lane == 0) slow_list[atomicAdd(slow_count, 1u)] = blk;
    }
}

// General blocks queued by k_int_rle.
__global__ void __launch_bounds__(RLE_WARPS * 32) k_int_rle_general(const Seg* __restrict__ segs,
                                                                    const BlockRec* __restrict__ blocks,
                                                                    const uint32_t* __restrict__ slow_list,
                                                                    const uint32_t* __restrict__ slow_count,
                                                                    const RunRec* __restrict__ table,
                                                                    const uint32_t* __restrict__ cnt,
                                                                    const uint32_t* __restrict__ dstart, uint32_t* err,
                                                                    uint32_t* mis) {
    __shared__ uint32_t patchmap_all[RLE_WARPS][16];
    __shared__ RunSlot slots_all[RLE_WARPS][32];
    const uint32_t nslow = *slow_count;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nslow; i += nwarps) {
        const uint32_t blk = slow_list[i];
        int_rle_block<false>(segs, blocks[blk], table + (uint64_t)blk * 32, cnt, dstart, err, mis,
                             patchmap_all[threadIdx.x >> 5], slots_all[threadIdx.x >> 5], nullptr, lane);
    }
}

// Whole-warp runs queued by the pre-pass: one warp per run.
__global__ void __launch_bounds__(RLE_WARPS * 32) k_coop_runs(const Seg* __restrict__ segs, const CoopRec* __restrict__ q,
                                                              const uint32_t* __restrict__ nq_ptr, uint32_t cap,
                                                              const uint32_t* __restrict__ cnt,
                                                              const uint32_t* __restrict__ dstart, uint32_t* err,
                                                              uint32_t* mis) {
    __shared__ uint32_t patchmap_all[RLE_WARPS][16];
    const uint32_t nq = min(*nq_ptr, cap);
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < nq; i += nwarps) {
        const CoopRec r = q[i];
        SegCtx c;
        c.s = &segs[r.seg];
        c.err = err;
        c.mis = mis;
        const Seg& s = *c.s;
        const uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_values;
        const uint64_t obase = s.start_idx >= 0 ? dstart[s.start_idx] : s.out_start;
        uint32_t rl = 0, nbytes = 0, take = 0;
        const uint32_t st = coop_run2(c, r.byte_off, r.skip, n - r.out_off, obase + r.out_off, patchmap_all[threadIdx.x >> 5],
                                      rl, nbytes, take);
        if (st) set_err(err, s.colstripe, st);
        __syncwarp();
    }
}

// Warp-per-segment variant for segments made of long runs (every run decoded by all 32 lanes).
__global__ void __launch_bounds__(RLE_WARPS * 32) k_int_rle_coop(const Seg* __restrict__ segs, uint32_t nseg,
                                                                 const uint32_t* __restrict__ cnt,
                                                                 const uint32_t* __restrict__ dstart, uint32_t* err,
                                                                 uint32_t* mis) {
    __shared__ uint32_t patchmap_all[RLE_WARPS][16];
    uint32_t* patchmap = patchmap_all[threadIdx.x >> 5];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nseg) return;
    SegCtx c;
    c.s = &segs[warp];
    c.err = err;
    c.mis = mis;
    const Seg& s = *c.s;
    const uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_values;
    const uint64_t obase = s.start_idx >= 0 ? dstart[s.start_idx] : s.out_start;
    uint32_t cur = s.start_byte, skip = s.run_skip, produced = 0;
    while (produced < n) {
        if (cur >= s.in_len) { set_err(err, s.colstripe, ORCB_OUT_OF_SPEC); return; }
        // pull the bytes of the following runs into L2 while this run is decoded (one line per lane, 4 KiB)
        {
            const uint32_t pf = cur + 2048u + 128u * (threadIdx.x & 31);
            if (pf < s.in_len) asm volatile("prefetch.global.L2 [%0];" ::"l"((const uint8_t*)s.in + pf));
        }
        uint32_t rl = 0, nbytes = 0, take = 0;
        const uint32_t st = coop_run2(c, cur, skip, n - produced, obase + produced, patchmap, rl, nbytes, take);
        if (st) { set_err(err, s.colstripe, st); return; }
        if (skip >= rl) skip -= rl;
        else { produced += take; skip = 0; }
        cur += nbytes;
    }
}

// ------------------------------------------------------------------------------------------------
// Byte RLE (encoding/byte.rs:228-247): TINYINT data and the byte layer under boolean RLE.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RLE_WARPS * 32) k_byte_rle(const Seg* __restrict__ segs, uint32_t nseg,
                                                             const uint32_t* __restrict__ cnt,
                                                             const uint32_t* __restrict__ dstart, uint32_t* err) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nseg) return;
    const Seg& s = segs[warp];
    const int lane = threadIdx.x & 31;
    const uint8_t* in = (const uint8_t*)s.in;
    const uint32_t len = s.in_len;
    // for boolean streams n_values counts BITS; aux = 1 marks "bits": convert to bytes incl. the bit offset
    uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_values;
    if (s.aux & 1) n = (n + (s.aux >> 1) + 7) / 8;  // aux>>1 = bit_skip
    uint8_t* out = (uint8_t*)s.out + (s.start_idx >= 0 ? dstart[s.start_idx] : s.out_start);
    uint32_t cur = s.start_byte, skip = s.run_skip, produced = 0;
    while (produced < n) {
        if (cur >= len) { set_err(err, s.colstripe, ORCB_IO_ERROR); return; }
        const uint32_t h = in[cur];
        uint32_t rl, run_bytes;
        if (h < 0x80) {
            rl = h + 3;
            run_bytes = 2;
            if (cur + 2 > len) { set_err(err, s.colstripe, ORCB_IO_ERROR); return; }
            const uint8_t v = in[cur + 1];
            const uint32_t avail = rl > skip ? rl - skip : 0;
            const uint32_t take = min(avail, n - produced);
            for (uint32_t i = lane; i < take; i += 32) out[produced + i] = v;
            if (skip >= rl) skip -= rl;
            else { produced += take; skip = 0; }
        } else {
            rl = 0x100 - h;
            run_bytes = 1 + rl;
            if (cur + run_bytes > len) { set_err(err, s.colstripe, ORCB_IO_ERROR); return; }
            const uint32_t avail = rl > skip ? rl - skip : 0;
            const uint32_t take = min(avail, n - produced);
            for (uint32_t i = lane; i < take; i += 32) out[produced + i] = in[cur + 1 + skip + i];
            if (skip >= rl) skip -= rl;
            else { produced += take; skip = 0; }
        }
        cur += run_bytes;
    }
}

// ------------------------------------------------------------------------------------------------
// Boolean bits (encoding/boolean.rs:101-113 + NullBuffer::from, array_decoder/mod.rs:209-213):
// MSB-first bytes -> LSB-first Arrow bitmap at an arbitrary bit position, plus popcount.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t load_msb_bits32(const uint8_t* src, uint32_t sbit) {
    // 32 stream bits starting at MSB-first bit index sbit, returned LSB-first
    const uint8_t* a = src + (sbit >> 3);
    const uintptr_t ai = (uintptr_t)a;
    const uint32_t* q = (const uint32_t*)(ai & ~(uintptr_t)3);
    const uint32_t sh = ((uint32_t)(ai & 3) << 3) + (sbit & 7);
    // per-byte bit reversal keeps byte order: brev reverses everything, bswap restores byte order
    const uint32_t w0 = bswap32(__brev(q[0]));
    const uint32_t w1 = bswap32(__brev(q[1]));
    return __funnelshift_r(w0, w1, sh);
}

__global__ void __launch_bounds__(128) k_bits(const BitSeg* __restrict__ segs, uint32_t nseg, uint32_t* cnt,
                                               const uint32_t* __restrict__ dstart) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nseg) return;
    const BitSeg& s = segs[warp];
    const int lane = threadIdx.x & 31;
    const uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_bits;
    const uint64_t d0 = s.start_idx >= 0 ? (uint64_t)dstart[s.start_idx] : (uint64_t)s.dst_bit0;
    const uint8_t* src = (const uint8_t*)s.src;
    uint32_t* dst = (uint32_t*)s.dst;
    uint32_t pc = 0;
    if (n > 0) {
        const uint64_t wfirst = d0 >> 5, wlast = (d0 + n - 1) >> 5;
        for (uint64_t wi = wfirst + lane; wi <= wlast; wi += 32) {
            const uint64_t lo = max(wi << 5, d0), hi = min((wi << 5) + 32, d0 + (uint64_t)n);
            const uint32_t nbits = (uint32_t)(hi - lo);
            const uint32_t sbit = (uint32_t)(lo - d0) + s.bit_skip;
            uint32_t v = load_msb_bits32(src, sbit);
            if (nbits < 32) v &= (1u << nbits) - 1;
            pc += __popc(v);
            const uint32_t word = v << (uint32_t)(lo - (wi << 5));
            if (nbits == 32) dst[wi] = word;
            else if (word) atomicOr(&dst[wi], word);
        }
    }
    if (s.popc_out >= 0) {
        pc = (uint32_t)warp_sum64(pc);
        if (lane == 0) cnt[s.popc_out] = pc;
    }
}

// exclusive scan of per-group non-null counts (value-stream entry index of each row group)
__global__ void k_seg_scan(const ScanDesc* __restrict__ descs, uint32_t ndesc, uint32_t* cnt, uint32_t* dstart) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ndesc) return;
    const ScanDesc d = descs[warp];
    const int lane = threadIdx.x & 31;
    uint32_t carry = 0;
    for (uint32_t g0 = 0; g0 < d.n_groups; g0 += 32) {
        const uint32_t g = g0 + lane;
        const uint32_t v = g < d.n_groups ? cnt[d.base + g] : 0;
        const uint32_t inc = warp_incl_scan(v, lane);
        if (g < d.n_groups) dstart[d.base + g] = carry + inc - v;
        carry += __shfl_sync(FULL, inc, 31);
    }
    if (lane == 0) {
        cnt[d.base + d.n_groups] = carry;
        dstart[d.base + d.n_groups] = carry;
    }
}

// ------------------------------------------------------------------------------------------------
// Decimal DATA: unbounded zigzag varints -> i128 (encoding/decimal.rs:46-51, integer/util.rs:475-527).
// Terminator bytes found with ballot; the lane owning a terminator assembles its value.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_varint128(const Seg* __restrict__ segs, uint32_t nseg,
                                                   const uint32_t* __restrict__ cnt,
                                                   const uint32_t* __restrict__ dstart, uint32_t* err) {
    // 128-byte windows.  Lane l looks at bytes l, l+32, l+64, l+96 so that each ballot is a terminator
    // bitmap in byte order.  Every window starts at the first byte of a value; terminator lanes publish the
    // end position of "their" value in shared memory, then the values are assembled one per lane per round.
    // The next window restarts right after the last terminator (a value cut by the edge is read again).
    // Stream bytes travel through a 512-byte ring per warp (4 aligned 128-byte chunks, indexed by the low bits
    // of the global address): the chunk after the ones a window can touch is always in flight in a register,
    // so the window never waits for memory.
    __shared__ uint32_t ring_all[4][128];
    __shared__ uint8_t ends_all[4][128];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nseg) return;
    const Seg& s = segs[warp];
    const int lane = threadIdx.x & 31;
    uint32_t* ring = ring_all[threadIdx.x >> 5];
    const uint8_t* ringb = (const uint8_t*)ring;
    uint8_t* ends = ends_all[threadIdx.x >> 5];
    const uint32_t len = s.in_len;
    const uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_values;
    const uint64_t obase = s.start_idx >= 0 ? dstart[s.start_idx] : s.out_start;
    const uint32_t colstripe = s.colstripe;
    uint4* out = (uint4*)s.out + obase;
    uint32_t produced = 0;
    uint32_t cur = s.start_byte;
    const uint32_t lt = (1u << lane) - 1;
    if (n == 0) return;
    const uint64_t a0 = (uint64_t)(uintptr_t)s.in;
    const uint64_t a_lim = a0 + len + 128;  // chunks are fetched only below this address (inside the arena slack)
    auto fetch = [&](uint64_t chunk) -> uint32_t {
        const uint64_t a = chunk + 4u * lane;
        return a < a_lim ? __ldg((const uint32_t*)(uintptr_t)a) : 0u;
    };
    uint64_t loaded_end = (a0 + cur) & ~(uint64_t)127;
    {
        const uint32_t w0 = fetch(loaded_end), w1 = fetch(loaded_end + 128), w2 = fetch(loaded_end + 256);
        ring[((uint32_t)(loaded_end >> 2) + lane) & 127] = w0;
        ring[((uint32_t)(loaded_end >> 2) + 32 + lane) & 127] = w1;
        ring[((uint32_t)(loaded_end >> 2) + 64 + lane) & 127] = w2;
        loaded_end += 384;
    }
    uint32_t pend = fetch(loaded_end);
    while (produced < n) {
        if (cur >= len) { set_err(err, colstripe, ORCB_IO_ERROR); return; }
        const uint64_t ca = a0 + cur;
        if (loaded_end < (ca & ~(uint64_t)127) + 384) {
            ring[((uint32_t)(loaded_end >> 2) + lane) & 127] = pend;
            loaded_end += 128;
            pend = fetch(loaded_end);
        }
        __syncwarp();
        const uint32_t cb = (uint32_t)ca;  // low address bits index the ring
        uint32_t T[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            // bytes past the end of the stream count as continuation bytes
            const uint32_t b = ringb[(cb + 32u * j + lane) & 511];
            T[j] = __ballot_sync(FULL, !(b & 0x80) && cur + 32u * j + lane < len);
        }
        const uint32_t c0 = __popc(T[0]), c1 = __popc(T[1]), c2 = __popc(T[2]), c3 = __popc(T[3]);
        const uint32_t total = c0 + c1 + c2 + c3;
        if (total == 0) {
            // no terminator in 128 bytes: either >= 20 continuation bytes (shift >= 128) or end of stream
            set_err(err, colstripe, (len - cur >= 20) ? ORCB_VARINT_TOO_LARGE : ORCB_IO_ERROR);
            return;
        }
        if ((T[0] >> lane) & 1) ends[__popc(T[0] & lt)] = (uint8_t)lane;
        if ((T[1] >> lane) & 1) ends[c0 + __popc(T[1] & lt)] = (uint8_t)(32 + lane);
        if ((T[2] >> lane) & 1) ends[c0 + c1 + __popc(T[2] & lt)] = (uint8_t)(64 + lane);
        if ((T[3] >> lane) & 1) ends[c0 + c1 + c2 + __popc(T[3] & lt)] = (uint8_t)(96 + lane);
        __syncwarp();
        const uint32_t room = n - produced;
        const uint32_t todo = min(total, room);
        for (uint32_t k0 = 0; k0 < todo; k0 += 32) {
            const uint32_t k = k0 + lane;
            const bool live = k < todo;
            const uint32_t end = live ? ends[k] : 0u;
            const uint32_t start = (live && k) ? (uint32_t)ends[k - 1] + 1 : 0u;
            const uint32_t nbv = end - start + 1;
            const uint32_t sa = cb + start;  // ring byte address of the value's first byte
            const uint32_t a = sa >> 2, shb = (sa & 3) * 8;
            if (__all_sync(FULL, !live || nbv <= 4)) {
                // common case (values below 2^28): one 32-bit window per value, 32-bit squeeze, hi half = sign
                if (live) {
                    uint32_t x0 = __funnelshift_r(ring[a & 127], ring[(a + 1) & 127], shb);
                    x0 &= 0xffffffffu >> (32 - 8 * nbv);
                    const uint32_t g = (x0 & 0x7fu) | ((x0 & 0x7f00u) >> 1) | ((x0 & 0x7f0000u) >> 2) | ((x0 & 0x7f000000u) >> 3);
                    const uint32_t sgn = 0u - (g & 1);
                    out[produced + k] = make_uint4((g >> 1) ^ sgn, sgn, sgn, sgn);
                }
                continue;
            }
            if (!live) continue;
            uint64_t lo, hi = 0;
            if (nbv <= 8) {
                // 8 little-endian bytes starting at the value, 7-bit groups squeezed together
                const uint32_t w0 = ring[a & 127], w1 = ring[(a + 1) & 127], w2 = ring[(a + 2) & 127];
                uint32_t x0 = __funnelshift_r(w0, w1, shb), x1 = __funnelshift_r(w1, w2, shb);
                if (nbv < 4) x0 &= (1u << (8 * nbv)) - 1;
                if (nbv <= 4) x1 = 0;
                else if (nbv < 8) x1 &= (1u << (8 * (nbv - 4))) - 1;
                const uint32_t g0 = (x0 & 0x7fu) | ((x0 & 0x7f00u) >> 1) | ((x0 & 0x7f0000u) >> 2) | ((x0 & 0x7f000000u) >> 3);
                const uint32_t g1 = (x1 & 0x7fu) | ((x1 & 0x7f00u) >> 1) | ((x1 & 0x7f0000u) >> 2) | ((x1 & 0x7f000000u) >> 3);
                lo = (uint64_t)g0 | ((uint64_t)g1 << 28);
            } else {
                if (nbv > 19) set_err(err, colstripe, ORCB_VARINT_TOO_LARGE);  // shift >= 128
                lo = 0;
                for (uint32_t q = 0; q < nbv && q < 19; q++) {
                    const uint64_t x = ringb[(sa + q) & 511] & 0x7f;
                    const uint32_t sft = 7 * q;
                    if (sft < 64) {
                        lo |= x << sft;
                        if (sft > 57) hi |= x >> (64 - sft);
                    } else {
                        hi |= x << (sft - 64);
                    }
                }
            }
            // zigzag: (v >>> 1) ^ -(v & 1) on 128 bits
            const uint64_t sgn = 0ull - (lo & 1);
            const uint64_t rlo = ((lo >> 1) | (hi << 63)) ^ sgn;
            const uint64_t rhi = (hi >> 1) ^ sgn;
            out[produced + k] = make_uint4((uint32_t)rlo, (uint32_t)(rlo >> 32), (uint32_t)rhi, (uint32_t)(rhi >> 32));
        }
        produced += total;
        cur += (uint32_t)ends[total - 1] + 1;
        __syncwarp();
    }
}

// ---- UTF-8 validation ------------------------------------------------------------------------------
// length of the character a lead byte opens (0 = not a lead byte)
__device__ __forceinline__ uint32_t utf8_len(uint32_t b) {
    return b < 0x80u ? 1u : (b >= 0xC2u && b <= 0xDFu) ? 2u : (b >= 0xE0u && b <= 0xEFu) ? 3u : (b >= 0xF0u && b <= 0xF4u) ? 4u : 0u;
}
__device__ __forceinline__ bool utf8_cont(uint32_t b) { return (b & 0xC0u) == 0x80u; }
// is the character starting at p well formed (str::from_utf8 rules: shortest form, no surrogates, <= U+10FFFF)?
__device__ __forceinline__ bool utf8_char_ok(const uint8_t* d, uint32_t p, uint32_t len) {
    const uint32_t b0 = d[p];
    if (b0 < 0x80u) return true;
    const uint32_t n = utf8_len(b0);
    if (n == 0 || p + n > len) return false;
    const uint32_t b1 = d[p + 1];
    uint32_t lo = 0x80u, hi = 0xBFu;
    if (b0 == 0xE0u) lo = 0xA0u;
    else if (b0 == 0xEDu) hi = 0x9Fu;
    else if (b0 == 0xF0u) lo = 0x90u;
    else if (b0 == 0xF4u) hi = 0x8Fu;
    if (b1 < lo || b1 > hi) return false;
    if (n >= 3 && !utf8_cont(d[p + 2])) return false;
    if (n == 4 && !utf8_cont(d[p + 3])) return false;
    return true;
}

// smallest position in [p0, p1) where the text stops being valid UTF-8 (0xffffffff = none): a lead byte must open a
// well-formed character, a continuation byte must be claimed by a lead byte at most three positions back
__device__ __forceinline__ uint32_t utf8_first_bad(const uint8_t* d, uint32_t p0, uint32_t p1, uint32_t len) {
    for (uint32_t p = p0; p < p1; p++) {
        const uint32_t b = d[p];
        if (b < 0x80u) continue;
        bool ok;
        if (utf8_cont(b)) {
            ok = false;
            for (uint32_t k = 1; k <= 3 && k <= p; k++) {
                const uint32_t lead = d[p - k];
                if (!utf8_cont(lead)) { ok = utf8_len(lead) > k; break; }
            }
        } else {
            ok = utf8_char_ok(d, p, len);
        }
        if (!ok) return p;
    }
    return 0xffffffffu;
}

// ------------------------------------------------------------------------------------------------
// Raw byte copies: FLOAT/DOUBLE streams (encoding/float.rs:70-74), string DATA (string.rs:135-140).
// ------------------------------------------------------------------------------------------------
constexpr uint32_t COPY_TILE = 16384;  // bytes per CTA

__global__ void __launch_bounds__(256) k_copy(const CopyDesc* __restrict__ descs, const uint2* __restrict__ tiles,
                                              uint32_t ntiles, const uint32_t* __restrict__ cnt, uint32_t* err,
                                              const StrCol* __restrict__ strcols) {
    if (blockIdx.x >= ntiles) return;
    const uint2 t = tiles[blockIdx.x];  // (desc index, tile index)
    const CopyDesc& d = descs[t.x];
    const uint64_t total = d.cnt_idx >= 0 ? (uint64_t)cnt[d.cnt_idx] * d.width : d.n_bytes;
    if (total > d.src_len) {
        if (t.y == 0 && threadIdx.x == 0) set_err(err, d.colstripe, ORCB_IO_ERROR);
        return;
    }
    const uint64_t off = (uint64_t)t.y * COPY_TILE;
    if (off >= total) return;
    const uint32_t nbytes = (uint32_t)min((uint64_t)COPY_TILE, total - off);
    const uint8_t* src = (const uint8_t*)d.src + off;
    uint8_t* dst = (uint8_t*)d.dst + off;  // dst tiles are 16-byte aligned (dst base is 256-byte aligned)
    const uint32_t mis = (uint32_t)((uintptr_t)src & 3);
    // string DATA is validated as UTF-8 while it passes through (string.rs:150-151): 16-byte groups without a high
    // bit are ASCII, the others are walked byte by byte; COPY_TILE == U8_TILE, so the tile is the flag unit
    const bool u8 = d.u8_col >= 0;
    const uint32_t u8_len = (uint32_t)total;
    const uint8_t* u8_d = (const uint8_t*)d.src;
    uint32_t hi_bits = 0, bad = 0xffffffffu;
    const uint32_t n16 = nbytes >> 4;
    if (mis == 0 && (((uintptr_t)src & 15) == 0)) {
        for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) {
            const uint4 v = __ldg((const uint4*)src + i);
            ((uint4*)dst)[i] = v;
            if (u8 && ((v.x | v.y | v.z | v.w) & 0x80808080u)) {
                hi_bits = 1;
                bad = min(bad, utf8_first_bad(u8_d, (uint32_t)off + i * 16, (uint32_t)off + i * 16 + 16, u8_len));
            }
        }
    } else {
        // unaligned source: aligned 32-bit loads + byte funnel, 16-byte aligned stores
        const uint32_t* q = (const uint32_t*)((uintptr_t)src & ~(uintptr_t)3);
        const uint32_t sh = mis * 8;
        for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) {
            const uint32_t* w = q + i * 4;
            const uint32_t a = __ldg(w), b = __ldg(w + 1), c2 = __ldg(w + 2), d2 = __ldg(w + 3);
            const uint32_t e = sh ? __ldg(w + 4) : 0u;
            uint4 v;
            v.x = __funnelshift_r(a, b, sh);
            v.y = __funnelshift_r(b, c2, sh);
            v.z = __funnelshift_r(c2, d2, sh);
            v.w = __funnelshift_r(d2, e, sh);
            ((uint4*)dst)[i] = v;
            if (u8 && ((v.x | v.y | v.z | v.w) & 0x80808080u)) {
                hi_bits = 1;
                bad = min(bad, utf8_first_bad(u8_d, (uint32_t)off + i * 16, (uint32_t)off + i * 16 + 16, u8_len));
            }
        }
    }
    for (uint32_t i = (n16 << 4) + threadIdx.x; i < nbytes; i += blockDim.x) {
        const uint8_t b = src[i];
        dst[i] = b;
        if (u8 && b >= 0x80u) {
            hi_bits = 1;
            bad = min(bad, utf8_first_bad(u8_d, (uint32_t)off + i, (uint32_t)off + i + 1, u8_len));
        }
    }
    if (u8) {
        const StrCol& sc = strcols[d.u8_col];
        if (__syncthreads_or(hi_bits) && threadIdx.x == 0) {
            atomicOr((uint32_t*)sc.u8_flags + (t.y >> 5), 1u << (t.y & 31));
            ((volatile uint32_t*)sc.u8_bad)[1] = 1u;  // the column has multi-byte characters at all
        }
        if (bad != 0xffffffffu) atomicMax((uint32_t*)sc.u8_bad, ~bad);
    }
}

// ------------------------------------------------------------------------------------------------
// decode_spaced (encoding/mod.rs:64-91): dense values -> row slots, null slots zero.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_spaced(const SpacedDesc* __restrict__ descs, uint32_t ndesc,
                                                const uint32_t* __restrict__ dstart) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ndesc) return;
    const SpacedDesc& d = descs[warp];
    const int lane = threadIdx.x & 31;
    const uint32_t* valid = (const uint32_t*)d.valid;
    uint64_t rank0 = dstart[d.start_idx];
    for (uint32_t t = 0; t < d.n_rows; t += 32) {
        const uint32_t nbits = min(32u, d.n_rows - t);
        uint32_t word = load_bits32(valid, (uint64_t)d.row0 + t);
        if (nbits < 32) word &= (1u << nbits) - 1;
        const bool v = (word >> lane) & 1;
        const uint64_t rank = rank0 + __popc(word & ((1u << lane) - 1));
        const uint64_t row = (uint64_t)d.row0 + t + lane;
        const bool inr = (uint32_t)lane < nbits;
        switch (d.width) {
            case 1: if (inr) ((uint8_t*)d.dst)[row] = v ? ((const uint8_t*)d.src)[rank] : (uint8_t)0; break;
            case 2: if (inr) ((uint16_t*)d.dst)[row] = v ? ((const uint16_t*)d.src)[rank] : (uint16_t)0; break;
            case 4: if (inr) ((uint32_t*)d.dst)[row] = v ? ((const uint32_t*)d.src)[rank] : 0u; break;
            case 8: if (inr) ((uint64_t*)d.dst)[row] = v ? ((const uint64_t*)d.src)[rank] : 0ull; break;
            case 16: if (inr) ((uint4*)d.dst)[row] = v ? ((const uint4*)d.src)[rank] : make_uint4(0, 0, 0, 0); break;
            default: {
                // bit mode (boolean values): dst is a zero-initialised stripe-level bitmap
                const uint32_t* sb = (const uint32_t*)d.src;
                const bool bit = v && ((sb[rank >> 5] >> (rank & 31)) & 1);
                const uint32_t bal = __ballot_sync(FULL, bit);
                if (lane == 0 && bal) {
                    const uint64_t bp = (uint64_t)d.row0 + t;
                    const uint32_t sh = (uint32_t)(bp & 31);
                    atomicOr(&((uint32_t*)d.dst)[bp >> 5], bal << sh);
                    if (sh && (bal >> (32 - sh))) atomicOr(&((uint32_t*)d.dst)[(bp >> 5) + 1], bal >> (32 - sh));
                }
            }
        }
        rank0 += __popc(word);
    }
}

// ------------------------------------------------------------------------------------------------
// Decimal scale repair (array_decoder/decimal.rs:138-166) — only does work when a SECONDARY value
// differed from the type scale (flag raised by k_int_rle OUT_SCALE).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_decimal_fix(const DecFixDesc* __restrict__ descs,
                                                     const uint32_t* __restrict__ cnt,
                                                     const uint32_t* __restrict__ mis) {
    const DecFixDesc& d = descs[blockIdx.y];
    if (!mis[d.colstripe]) return;
    const uint32_t n = d.cnt_idx >= 0 ? cnt[d.cnt_idx] : d.n;
    __int128* vals = (__int128*)d.vals;
    const int32_t* scales = (const int32_t*)d.scales;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint32_t vs = (uint32_t)scales[i];
        if (vs == d.fixed_scale) continue;
        __int128 v = vals[i];
        if (d.fixed_scale < vs) {
            uint32_t k = vs - d.fixed_scale;
            // 10^k overflows i128 for k >= 39 (the reference panics in pow); quotient is then 0
            if (k >= 39) v = 0;
            else {
                __int128 f = 1;
                for (uint32_t j = 0; j < k; j++) f *= 10;
                v = v / f;
            }
        } else {
            uint32_t k = d.fixed_scale - vs;
            unsigned __int128 f = 1;
            for (uint32_t j = 0; j < k && j < 64; j++) f *= 10;
            v = (__int128)((unsigned __int128)v * f);
        }
        vals[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// Timestamp recombination (encoding/timestamp.rs:121-196)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_timestamp(const TsDesc* __restrict__ descs, const uint32_t* __restrict__ cnt,
                                                   uint32_t* err) {
    const TsDesc& d = descs[blockIdx.y];
    const uint32_t n = d.cnt_idx >= 0 ? cnt[d.cnt_idx] : d.n;
    const int64_t* secs = (const int64_t*)d.secs;
    const int64_t* nanos = (const int64_t*)d.nanos;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint64_t ns = (uint64_t)nanos[i];
        const uint32_t zeros = (uint32_t)(ns & 7);
        ns >>= 3;
        if (zeros) {
            uint64_t p = 100;
            for (uint32_t j = 1; j < zeros; j++) p *= 10;
            ns *= p;  // wrapping, as the release-mode reference
        }
        int64_t sec = (int64_t)((uint64_t)secs[i] + (uint64_t)d.base);
        if (sec < 0 && ns > 999999ull) sec -= 1;
        const __int128 t = (__int128)sec * 1000000000 + (__int128)ns;
        if (d.as_i128) {
            // Decimal128(38, 9): nanoseconds, no range check (timestamp.rs:194-197); writer zone as below (:316-333)
            __int128 w = t;
            if (d.tz_on) {
                const __int128 ns = 1000000000;
                __int128 q = t / ns;
                if (t % ns < 0) q -= 1;  // div_euclid
                const int64_t inst = (int64_t)q;
                const int64_t* at = (const int64_t*)d.tz_at;
                uint32_t lo = 0, hi = d.tz_n;
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (at[mid] <= inst) lo = mid + 1;
                    else hi = mid;
                }
                const int64_t off = lo == 0 ? (int64_t)d.tz_first : (int64_t)((const int32_t*)d.tz_off)[lo - 1];
                w = t + (__int128)off * ns;
            }
            ((__int128*)d.out)[i] = w;
        } else {
            const __int128 u = (__int128)d.unit_ns;
            const __int128 q = t / u;
            if (t % u != 0 || q > (__int128)INT64_MAX || q < (__int128)INT64_MIN) set_err(err, d.colstripe, ORCB_DECODE_TIMESTAMP);
            int64_t v = (int64_t)q;
            if (d.tz_on) {
                // the value is an instant; the reference re-reads its wall clock in the writer's zone as UTC
                const int64_t per_s = 1000000000 / d.unit_ns;
                const int64_t inst = v / per_s - (v % per_s < 0);  // floor to seconds
                const int64_t* at = (const int64_t*)d.tz_at;
                uint32_t lo = 0, hi = d.tz_n;  // first transition after the instant
                while (lo < hi) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (at[mid] <= inst) lo = mid + 1;
                    else hi = mid;
                }
                const int64_t off = lo == 0 ? (int64_t)d.tz_first : (int64_t)((const int32_t*)d.tz_off)[lo - 1];
                const __int128 w = (__int128)v + (__int128)off * per_s;
                // the reference turns an unrepresentable nanosecond value into a null; not reproduced
                if (w > (__int128)INT64_MAX || w < (__int128)INT64_MIN) set_err(err, d.colstripe, ORCB_NOT_IMPLEMENTED);
                v = (int64_t)w;
            }
            ((int64_t*)d.out)[i] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Strings (array_decoder/string.rs:111-153, 205-224): lengths -> per-batch i32 offsets, dictionary gather.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ const StrCol& find_strcol(const StrCol* cols, uint32_t ncols, uint32_t tile) {
    uint32_t lo = 0, hi = ncols;  // last col with tile0 <= tile
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (cols[mid].tile0 <= tile) lo = mid;
        else hi = mid;
    }
    return cols[lo];
}

// length of row r (0 for null rows; dictionary keys are bounds-checked for valid rows)
__device__ __forceinline__ uint32_t str_row_len(const StrCol& c, uint32_t r, uint32_t* err, int32_t* key_out) {
    const int32_t x = ((const int32_t*)c.lens)[r];
    if (c.mode == 0) return (uint32_t)x;
    bool valid = true;
    if (c.valid) valid = (((const uint32_t*)c.valid)[r >> 5] >> (r & 31)) & 1;
    if (!valid) { *key_out = -1; return 0; }
    if ((uint32_t)x >= c.dict_size) {  // DictionaryArray::try_new rejects out-of-range valid keys
        set_err(err, c.colstripe, ORCB_ARROW);
        *key_out = -1;
        return 0;
    }
    *key_out = x;
    return (uint32_t)((const int32_t*)c.dict_len)[x];
}

// One CTA per U8_TILE bytes, 16 bytes per thread.  Pure-ASCII groups (the common case) cost four loads and an OR;
// other groups are walked byte by byte: a lead byte must open a well-formed character, a continuation byte must be
// claimed by a lead byte at most three positions back.  Only the smallest offending position is kept: whether
// it matters is decided once the number of bytes the values really use is known (k_str_tile_scan).
__global__ void __launch_bounds__(256) k_utf8(const StrCol* __restrict__ cols, const uint2* __restrict__ tiles, uint32_t ntiles) {
    if (blockIdx.x >= ntiles) return;
    const uint2 t = tiles[blockIdx.x];  // (column, tile)
    const StrCol& c = cols[t.x];
    const uint8_t* d = (const uint8_t*)c.u8_src;
    const uint32_t len = c.u8_len;
    // threads take 16-byte groups that are aligned in memory (128-bit loads, four in flight per thread): tile k is
    // groups [G k, G k + G), G = U8_TILE / 16, counted from the stream's first byte rounded down to 16; bytes outside
    // the stream are masked
    constexpr uint32_t G = U8_TILE / 16, PER_THREAD = G / 256;
    const uintptr_t base = (uintptr_t)d;
    const uintptr_t g0 = (base & ~(uintptr_t)15) + ((uintptr_t)t.y * G + threadIdx.x) * 16u;
    uint4 v[PER_THREAD];
#pragma unroll
    for (uint32_t u = 0; u < PER_THREAD; u++) {
        const uintptr_t g = g0 + (uintptr_t)u * 256u * 16u;
        v[u] = (int64_t)g - (int64_t)base < (int64_t)len ? __ldg((const uint4*)g) : make_uint4(0, 0, 0, 0);
    }
    bool multi = false;
    uint32_t bad = 0xffffffffu;
#pragma unroll
    for (uint32_t u = 0; u < PER_THREAD; u++) {
        const int64_t rel = (int64_t)(g0 + (uintptr_t)u * 256u * 16u) - (int64_t)base;  // stream offset of the group (-15.. for the first)
        if (rel >= (int64_t)len) continue;
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        const uint32_t lo = rel < 0 ? (uint32_t)(-rel) : 0u;                                   // first byte of the group inside the stream
        const uint32_t hi = (int64_t)len - rel < 16 ? (uint32_t)((int64_t)len - rel) : 16u;   // one past the last
        uint32_t any = 0;
        if (lo == 0 && hi == 16) {
            any = w[0] | w[1] | w[2] | w[3];
        } else {
#pragma unroll
            for (uint32_t i = 0; i < 4; i++) {
                uint32_t m = 0xffffffffu;
                const uint32_t b0 = 4 * i;
                if (lo > b0) m &= lo - b0 >= 4 ? 0u : 0xffffffffu << (8 * (lo - b0));
                if (hi < b0 + 4) m &= hi <= b0 ? 0u : 0xffffffffu >> (8 * (b0 + 4 - hi));
                any |= w[i] & m;
            }
        }
        if (!(any & 0x80808080u)) continue;
        multi = true;
        bad = min(bad, utf8_first_bad(d, (uint32_t)(rel + lo), (uint32_t)(rel + hi), len));
    }
    if (__syncthreads_or(multi) && threadIdx.x == 0) {
        atomicOr((uint32_t*)c.u8_flags + (t.y >> 5), 1u << (t.y & 31));
        ((volatile uint32_t*)c.u8_bad)[1] = 1u;  // the column has multi-byte characters at all
    }
    if (bad != 0xffffffffu) atomicMax((uint32_t*)c.u8_bad, ~bad);
}

// values must be cut at character boundaries: does `pos` (< total) fall on a continuation byte?
__device__ __forceinline__ bool utf8_mid_char(const StrCol& c, uint32_t pos) {
    // k_copy cuts its tiles at stream offsets, k_utf8 at 16-byte aligned addresses: look at both candidates
    const uint32_t ta = pos / U8_TILE, tb2 = (pos + ((uint32_t)c.u8_src & 15u)) / U8_TILE;
    const uint32_t* fl = (const uint32_t*)c.u8_flags;
    if (!(((fl[ta >> 5] >> (ta & 31)) | (fl[tb2 >> 5] >> (tb2 & 31))) & 1u)) return false;
    return utf8_cont(((const uint8_t*)c.u8_src)[pos]);
}

// dictionary LENGTH -> offsets (one warp per dictionary)
__global__ void k_dict_prepare(StrCol* cols, uint32_t ncols, uint32_t* err) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ncols) return;
    const StrCol& c = cols[warp];
    if (c.mode != 1) return;
    const int lane = threadIdx.x & 31;
    const int32_t* dl = (const int32_t*)c.dict_len;
    int32_t* doff = (int32_t*)c.dict_off;
    uint64_t carry = 0;
    for (uint32_t i0 = 0; i0 < c.dict_size; i0 += 32) {
        const uint32_t i = i0 + lane;
        const uint64_t v = i < c.dict_size ? (uint64_t)(uint32_t)dl[i] : 0;
        const uint64_t inc = warp_incl_scan64(v, lane);
        if (i < c.dict_size) doff[i] = (int32_t)(carry + inc - v);
        carry += __shfl_sync(FULL, inc, 31);
    }
    // dictionaries whose entries all have the same length (flags, codes) get a scan-free fast path
    uint32_t first_len = c.dict_size ? (uint32_t)dl[0] : 0u;
    bool same = true;
    for (uint32_t i = lane; i < c.dict_size; i += 32) same &= (uint32_t)dl[i] == first_len;
    same = __all_sync(FULL, same);
    // every dictionary entry starts at a character boundary (the dictionary is a string array of its own)
    if (c.u8_src && carry <= c.dict_data_len) {
        bool mid = false;
        for (uint32_t i = lane; i < c.dict_size; i += 32) {
            const uint32_t o = (uint32_t)doff[i];
            if (o < (uint32_t)carry) mid |= utf8_mid_char(c, o);
        }
        if (mid) set_err(err, c.colstripe, ORCB_ARROW);
    }
    if (lane == 0) {
        cols[warp].data_cap = (same && c.dict_size && first_len >= 1 && first_len <= 4 && !c.valid) ? first_len : 0;
        doff[c.dict_size] = (int32_t)carry;
        // the dictionary itself is a string batch: offsets must fit i32 and its bytes must exist
        if (carry > 0x7fffffffull) set_err(err, c.colstripe, ORCB_OFFSET_OVERFLOW);
        else if (carry > c.dict_data_len) set_err(err, c.colstripe, ORCB_ARROW);
    }
}

__device__ __forceinline__ void tile_rows(const StrCol& c, uint32_t tile, uint32_t& b, uint32_t& r0, uint32_t& nr) {
    b = tile / c.tiles_per_batch;
    const uint32_t k = tile - b * c.tiles_per_batch;
    const uint32_t brow0 = b * c.batch_size;
    const uint32_t brows = min(c.batch_size, c.n_rows - brow0);
    r0 = brow0 + k * STR_TILE;
    const uint32_t off = k * STR_TILE;
    nr = off >= brows ? 0 : min(STR_TILE, brows - off);
}

__global__ void __launch_bounds__(128) k_str_tile_sum(const StrCol* __restrict__ cols, uint32_t ncols, uint32_t ntiles,
                                                      uint32_t* err) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ntiles) return;
    const StrCol& c = find_strcol(cols, ncols, warp);
    const uint32_t tile = warp - c.tile0;
    const int lane = threadIdx.x & 31;
    uint32_t b, r0, nr;
    tile_rows(c, tile, b, r0, nr);
    uint64_t sum = 0;
    if (c.mode == 1 && c.data_cap) {
        // uniform entry length, no nulls: only the key bounds need checking
        bool bad = false;
        for (uint32_t i = lane; i < nr; i += 32) bad |= (uint32_t)((const int32_t*)c.lens)[r0 + i] >= c.dict_size;
        if (bad) set_err(err, c.colstripe, ORCB_ARROW);
        if (lane == 0) ((uint64_t*)c.tile_base)[tile] = (uint64_t)nr * c.data_cap;
        return;
    }
    const int32_t* lens = (const int32_t*)c.lens + r0;
    const uint32_t mode = c.mode, dict_size = c.dict_size;
    const uint32_t* valid = (const uint32_t*)c.valid;
    const int32_t* dl = (const int32_t*)c.dict_len;
    bool bad = false;
    for (uint32_t j = 0; j < nr; j += 256) {
        // 8 independent loads per lane in flight, then the dependent dictionary lookups
        int32_t x[8];
#pragma unroll
        for (uint32_t u = 0; u < 8; u++) {
            const uint32_t i = j + u * 32 + lane;
            x[u] = i < nr ? lens[i] : (mode == 0 ? 0 : -1);
        }
        if (mode == 0) {
#pragma unroll
            for (uint32_t u = 0; u < 8; u++) sum += (uint32_t)x[u];
        } else {
            uint32_t l[8];
#pragma unroll
            for (uint32_t u = 0; u < 8; u++) {
                const uint32_t i = j + u * 32 + lane;
                bool ok = i < nr;
                if (ok && valid) ok = (valid[(r0 + i) >> 5] >> ((r0 + i) & 31)) & 1;
                l[u] = 0;
                if (ok) {
                    if ((uint32_t)x[u] >= dict_size) bad = true;  // DictionaryArray::try_new rejects out-of-range valid keys
                    else l[u] = (uint32_t)dl[x[u]];
                }
            }
#pragma unroll
            for (uint32_t u = 0; u < 8; u++) sum += l[u];
        }
    }
    if (bad) set_err(err, c.colstripe, ORCB_ARROW);
    sum = warp_sum64(sum);
    if (lane == 0) ((uint64_t*)c.tile_base)[tile] = sum;
}

// per column: exclusive scan of tile sums, batch bases, overflow checks, dictionary data allocation
__global__ void k_str_tile_scan(StrCol* cols, uint32_t ncols, uint32_t* err, JobState* st, uint64_t heap_base,
                                uint64_t heap_cap, uint64_t* ptr_table) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ncols) return;
    StrCol& c = cols[warp];
    const int lane = threadIdx.x & 31;
    uint64_t* tb = (uint64_t*)c.tile_base;
    uint64_t* bb = (uint64_t*)c.batch_base;
    uint64_t carry = 0;
    for (uint32_t t0 = 0; t0 < c.n_tiles; t0 += 32) {
        const uint32_t t = t0 + lane;
        const uint64_t v = t < c.n_tiles ? tb[t] : 0;
        const uint64_t inc = warp_incl_scan64(v, lane);
        if (t < c.n_tiles) {
            const uint64_t ex = carry + inc - v;
            tb[t] = ex;
            if (t % c.tiles_per_batch == 0) bb[t / c.tiles_per_batch] = ex;
        }
        carry += __shfl_sync(FULL, inc, 31);
    }
    __syncwarp();
    if (lane == 0) {
        tb[c.n_tiles] = carry;
        bb[c.n_batches] = carry;
    }
    __syncwarp();
    // OffsetOverflow: a batch whose bytes exceed i32::MAX (string.rs:125-133)
    for (uint32_t b = lane; b < c.n_batches; b += 32) {
        if (bb[b + 1] - bb[b] > 0x7fffffffull) set_err(err, c.colstripe, c.mode == 0 ? ORCB_OFFSET_OVERFLOW : ORCB_ARROW);
    }
    if (lane == 0) {
        if (c.mode == 1) {
            const unsigned long long need = (carry + 255ull) & ~255ull;
            const unsigned long long at = atomicAdd(&st->heap_top, need);
            if (at + need > heap_cap) {
                set_err(err, c.colstripe, ORCB_DEVICE_HEAP_OVERFLOW);
                c.data = 0;
            } else {
                c.data = heap_base + at;
            }
        } else if (carry > c.data_len) {
            // fewer DATA bytes than the lengths claim: GenericByteArray::try_new fails (Arrow error)
            set_err(err, c.colstripe, ORCB_ARROW);
        }
        if (c.u8_src) {
            // the bytes the values use: direct = sum of the lengths, dictionary = sum of the entry lengths
            const uint64_t used = c.mode == 0 ? carry : (uint64_t)(uint32_t)((const int32_t*)c.dict_off)[c.dict_size];
            if (used <= c.u8_len) {
                const uint32_t word = *(const uint32_t*)c.u8_bad;
                const uint32_t first_bad = ~word;  // 0xffffffff when nothing was found
                const uint8_t* d = (const uint8_t*)c.u8_src;
                // invalid character inside the used prefix, or a character that straddles its end
                if ((word && first_bad < used) || (used && used < c.u8_len && utf8_cont(d[used]))) set_err(err, c.colstripe, ORCB_ARROW);
            }
        }
        ptr_table[c.meta_slot] = c.data;
    }
}

constexpr uint32_t SD_ENTRIES = 256;   // dictionaries up to this many entries / bytes are staged in shared memory
constexpr uint32_t SD_BYTES = 2048;
constexpr uint32_t KEY_GROUP = 256;    // lengths / keys are fetched this many rows ahead (8 loads per lane in flight)
constexpr uint32_t STAGE_BYTES = 1024; // per-warp ring of gathered bytes, indexed by the low bits of the global address

__global__ void __launch_bounds__(128) k_str_offsets(const StrCol* __restrict__ cols, uint32_t ncols, uint32_t ntiles,
                                                     uint32_t* err) {
    __shared__ uint16_t s_doff_all[4][SD_ENTRIES + 2];
    __shared__ __align__(16) uint8_t s_ddata_all[4][SD_BYTES + 16];
    __shared__ __align__(16) uint8_t s_stage_all[4][STAGE_BYTES];
    __shared__ int32_t s_keys_all[4][KEY_GROUP];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ntiles) return;
    const StrCol& c = find_strcol(cols, ncols, warp);
    const uint32_t tile = warp - c.tile0;
    const int lane = threadIdx.x & 31;
    uint16_t* s_doff = s_doff_all[threadIdx.x >> 5];
    uint8_t* s_ddata = s_ddata_all[threadIdx.x >> 5];
    uint8_t* s_stage = s_stage_all[threadIdx.x >> 5];
    uint32_t b, r0, nr;
    tile_rows(c, tile, b, r0, nr);
    const uint64_t* tb = (const uint64_t*)c.tile_base;
    const uint64_t* bb = (const uint64_t*)c.batch_base;
    const uint64_t bbase = bb[b];
    uint64_t run = tb[tile];  // absolute byte offset of the first row of this tile
    const uint32_t batch_size = c.batch_size, mode = c.mode, dict_size = c.dict_size, colstripe = c.colstripe;
    int32_t* offs = (int32_t*)c.offsets + (uint64_t)b * (batch_size + 1) + (r0 - b * batch_size);
    const uint8_t* dict = (const uint8_t*)c.dict_data;
    const int32_t* doff = (const int32_t*)c.dict_off;
    const int32_t* lens = (const int32_t*)c.lens + r0;
    const uint32_t* valid = (const uint32_t*)c.valid;
    uint8_t* data = (uint8_t*)c.data;
    int32_t* s_keys = s_keys_all[threadIdx.x >> 5];
    // small dictionaries live in shared memory for the whole tile
    bool sdict = false;
    if (mode == 1 && nr > 0 && dict_size <= SD_ENTRIES) {
        const uint32_t dbytes = (uint32_t)doff[dict_size];
        if (dbytes <= SD_BYTES) {
            sdict = true;
            for (uint32_t i = lane; i <= dict_size; i += 32) s_doff[i] = (uint16_t)doff[i];
            for (uint32_t i = lane; i < dbytes; i += 32) s_ddata[i] = dict[i];
            __syncwarp();
        }
    }
    const bool uniform = mode == 1 && c.data_cap && data;
    const uint32_t L = (uint32_t)c.data_cap;  // uniform entry length (1..4), no nulls
    const uint32_t rel0 = (uint32_t)(run - bbase);
    uint8_t* const dp0 = data + run;
    // `fl`: bytes below this absolute address are in global memory; [fl, data + run) waits in the ring
    const uint64_t d0 = (uint64_t)(uintptr_t)data;
    uint64_t fl = d0 + run;
    constexpr uint32_t RING = STAGE_BYTES - 1;
    // lengths / keys travel one group ahead of their use: 8 loads per lane in flight while a group is processed
    int32_t tn[KEY_GROUP / 32];
#pragma unroll
    for (uint32_t u = 0; u < KEY_GROUP / 32; u++) {
        const uint32_t i = u * 32 + lane;
        tn[u] = i < nr ? lens[i] : 0;
    }
    for (uint32_t g0 = 0; g0 < nr; g0 += KEY_GROUP) {
    __syncwarp();
#pragma unroll
    for (uint32_t u = 0; u < KEY_GROUP / 32; u++) s_keys[u * 32 + lane] = tn[u];
    __syncwarp();
#pragma unroll
    for (uint32_t u = 0; u < KEY_GROUP / 32; u++) {
        const uint32_t i = g0 + KEY_GROUP + u * 32 + lane;
        tn[u] = i < nr ? lens[i] : 0;
    }
    const uint32_t gend = min(nr, g0 + KEY_GROUP);
    if (uniform) {
        // offsets are an arithmetic progression and every row copies L bytes
        for (uint32_t i = g0 + lane; i < gend; i += 32) {
            offs[i] = (int32_t)(rel0 + i * L);
            const uint32_t key = (uint32_t)s_keys[i - g0];
            if (key < dict_size) {
                const uint8_t* sp = sdict ? s_ddata + key * L : dict + key * L;
                for (uint32_t k = 0; k < L; k++) dp0[i * L + k] = sp[k];
            }
        }
    } else {
        for (uint32_t i0 = g0; i0 < gend; i0 += 32) {
            const uint32_t i = i0 + lane;
            int32_t key = -1;
            uint32_t l = 0, so = 0;
            if (i < nr) {
                const int32_t x = s_keys[i - g0];
                if (mode == 0) {
                    l = (uint32_t)x;
                } else {
                    bool ok = true;
                    if (valid) ok = (valid[(r0 + i) >> 5] >> ((r0 + i) & 31)) & 1;
                    if (ok) {
                        if ((uint32_t)x >= dict_size) {  // DictionaryArray::try_new rejects out-of-range valid keys
                            set_err(err, colstripe, ORCB_ARROW);
                        } else {
                            key = x;
                            if (sdict) {
                                so = s_doff[x];
                                l = (uint32_t)s_doff[x + 1] - so;
                            } else {
                                so = (uint32_t)doff[x];
                                l = (uint32_t)doff[x + 1] - so;
                            }
                        }
                    }
                }
            }
            // a single length is below 2^31: the prefix of 32 of them fits 32 bits unless one is huge
            uint64_t inc;
            if (__any_sync(FULL, l >> 26)) inc = warp_incl_scan64(l, lane);
            else inc = warp_incl_scan(l, lane);
            const uint64_t abs0 = run + inc - l;
            if (i < nr) offs[i] = (int32_t)(abs0 - bbase);
            const uint64_t B64 = __shfl_sync(FULL, inc, 31);
            if (mode == 1 && data) {
                if (sdict && B64 <= STAGE_BYTES - 32) {
                    if (l) {
                        // own string -> ring: bytes up to a word boundary of the destination, whole words, trailing
                        // bytes; source words are read unaligned (two aligned words + byte permute)
                        const uint32_t* sw = (const uint32_t*)s_ddata;
                        uint32_t d = (uint32_t)(d0 + abs0);
                        const uint32_t nh = min((0u - d) & 3u, l);
                        if (nh) {
                            const uint32_t hw = __byte_perm(sw[so >> 2], sw[(so >> 2) + 1], 0x3210u + 0x1111u * (so & 3));
                            s_stage[d & RING] = (uint8_t)hw;
                            if (nh > 1) s_stage[(d + 1) & RING] = (uint8_t)(hw >> 8);
                            if (nh > 2) s_stage[(d + 2) & RING] = (uint8_t)(hw >> 16);
                        }
                        d += nh;
                        const uint32_t sidx = so + nh;
                        uint32_t rem = l - nh, wi = sidx >> 2;
                        const uint32_t sel = 0x3210u + 0x1111u * (sidx & 3);
                        uint32_t w0 = sw[wi];
                        while (rem >= 4) {
                            const uint32_t w1 = sw[++wi];
                            *(uint32_t*)(s_stage + (d & RING)) = __byte_perm(w0, w1, sel);
                            w0 = w1;
                            d += 4;
                            rem -= 4;
                        }
                        if (rem) {
                            const uint32_t tw = __byte_perm(w0, sw[wi + 1], sel);
                            s_stage[d & RING] = (uint8_t)tw;
                            if (rem > 1) s_stage[(d + 1) & RING] = (uint8_t)(tw >> 8);
                            if (rem > 2) s_stage[(d + 2) & RING] = (uint8_t)(tw >> 16);
                        }
                    }
                    __syncwarp();
                    // write what is complete: leading bytes up to a 16-byte boundary (first round only), then 16-byte groups
                    const uint64_t end = d0 + run + B64;
                    if (fl & 15) {
                        uint64_t h = (fl + 15) & ~(uint64_t)15;
                        if (h > end) h = end;
                        if (fl + lane < h) *(uint8_t*)(uintptr_t)(fl + lane) = s_stage[(uint32_t)(fl + lane) & RING];
                        fl = h;
                    }
                    const uint64_t e16 = end & ~(uint64_t)15;
                    for (uint64_t g = fl + 16u * lane; g < e16; g += 512)
                        *(uint4*)(uintptr_t)g = *(const uint4*)(s_stage + ((uint32_t)g & RING));
                    if (e16 > fl) fl = e16;
                    __syncwarp();
                } else {
                    // drain the ring, then every lane copies its own string
                    const uint64_t cur_end = d0 + run;
                    for (uint64_t g = fl + lane; g < cur_end; g += 32) *(uint8_t*)(uintptr_t)g = s_stage[(uint32_t)g & RING];
                    __syncwarp();
                    if (key >= 0) {
                        const uint8_t* sp = dict + so;
                        uint8_t* dp = data + abs0;
                        for (uint32_t k = 0; k < l; k++) dp[k] = sp[k];
                    }
                    fl = cur_end + B64;
                }
            }
            run += B64;
        }
    }
    }
    if (!uniform) {
        if (mode == 1 && data) {
            const uint64_t cur_end = d0 + run;
            for (uint64_t g = fl + lane; g < cur_end; g += 32) *(uint8_t*)(uintptr_t)g = s_stage[(uint32_t)g & RING];
        }
    }
    // closing offset of the batch
    const uint32_t brow0 = b * batch_size;
    const uint32_t brows = min(batch_size, c.n_rows - brow0);
    if (lane == 0 && r0 + nr == brow0 + brows) ((int32_t*)c.offsets)[(uint64_t)b * (batch_size + 1) + brows] = (int32_t)(bb[b + 1] - bbase);
}

// Direct strings whose DATA stream has multi-byte characters: every value must start at a character boundary
// (GenericByteArray::<Utf8>::try_new).  One CTA per string column; pure-ASCII columns leave at once.
__global__ void __launch_bounds__(128) k_utf8_bounds(const StrCol* __restrict__ cols, uint32_t ncols, uint32_t* err) {
    if (blockIdx.x >= ncols) return;
    const StrCol& c = cols[blockIdx.x];
    if (c.mode != 0 || !c.u8_src || ((const uint32_t*)c.u8_bad)[1] == 0) return;
    const int lane = threadIdx.x & 31;
    const uint64_t* bb = (const uint64_t*)c.batch_base;
    const uint64_t total = min(bb[c.n_batches], (uint64_t)c.u8_len);
    bool mid = false;
    for (uint32_t tile = threadIdx.x >> 5; tile < c.n_tiles; tile += blockDim.x >> 5) {
        uint32_t b, r0, nr;
        tile_rows(c, tile, b, r0, nr);
        const uint64_t bbase = bb[b];
        const int32_t* offs = (const int32_t*)c.offsets + (uint64_t)b * (c.batch_size + 1) + (r0 - b * c.batch_size);
        for (uint32_t i = lane; i < nr; i += 32) {
            const uint64_t a = bbase + (uint32_t)offs[i];
            if (a < total) mid |= utf8_mid_char(c, (uint32_t)a);
        }
    }
    if (mid) set_err(err, c.colstripe, ORCB_ARROW);
}

// ------------------------------------------------------------------------------------------------
// Stripe-level bitmaps -> per-batch bitmaps + null counts (derive_present_vec, array_decoder/mod.rs:231-252)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_repack(const RepackDesc* __restrict__ descs, uint32_t ndesc, uint32_t nwork,
                                                uint32_t* nulls) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nwork) return;
    uint32_t lo = 0, hi = ndesc;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (descs[mid].batch0 <= warp) lo = mid;
        else hi = mid;
    }
    const RepackDesc& d = descs[lo];
    const uint32_t b = warp - d.batch0;
    const int lane = threadIdx.x & 31;
    const uint32_t row0 = b * d.batch_size;
    const uint32_t rows = min(d.batch_size, d.n_rows - row0);
    const uint32_t* src = (const uint32_t*)d.src;
    uint32_t* dst = (uint32_t*)((uint8_t*)d.dst + (uint64_t)b * d.dst_stride);
    const uint32_t nwords = (rows + 31) >> 5;
    uint32_t pc = 0;
    for (uint32_t w = lane; w < nwords; w += 32) {
        uint32_t v = load_bits32(src, (uint64_t)row0 + (uint64_t)w * 32);
        const uint32_t rem = rows - w * 32;
        if (rem < 32) v &= (1u << rem) - 1;
        dst[w] = v;
        pc += __popc(v);
    }
    if (d.null_out >= 0) {
        pc = (uint32_t)warp_sum64(pc);
        if (lane == 0) nulls[d.null_out + b] = rows - pc;
    }
}

// ------------------------------------------------------------------------------------------------
// Compression chunks (src/compression.rs:113-123, 244-275): original chunks are copied, Snappy and LZ4
// blocks are decoded by one warp per chunk: lane 0 walks the tags, all lanes move the bytes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_copy_fwd(uint8_t* dst, const uint8_t* src, uint32_t n, int lane) {
    if (n < 256) {
        for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
        return;
    }
    // long copies (original chunks, incompressible literals): whole words where dst is word aligned, the source
    // words funnel-shifted into place (reads up to 7 bytes past the source, inside the arena slack)
    const uint32_t head = (uint32_t)(0u - (uint32_t)(uintptr_t)dst) & 3u;
    if ((uint32_t)lane < head) dst[lane] = src[lane];
    const uint32_t nw = (n - head) >> 2;
    uint32_t* dw = (uint32_t*)(dst + head);
    const uintptr_t sa = (uintptr_t)(src + head);
    const uint32_t* sw = (const uint32_t*)(sa & ~(uintptr_t)3);
    const uint32_t shb = (uint32_t)(sa & 3) * 8;
#pragma unroll 4
    for (uint32_t i = lane; i < nw; i += 32) dw[i] = __funnelshift_r(__ldg(sw + i), __ldg(sw + i + 1), shb);
    const uint32_t done = head + nw * 4;
    if (done + lane < n) dst[done + lane] = src[done + lane];
}
// overlapping back-reference: dst[i] = dst[i - dist]; bytes further than `dist` ahead depend on bytes
// written earlier in this same copy, so copy in rounds of `dist` bytes when dist < 32
__device__ __forceinline__ void warp_copy_match(uint8_t* out, uint32_t o, uint32_t dist, uint32_t n, int lane) {
    if (dist >= 32) {
        for (uint32_t i0 = 0; i0 < n; i0 += 32) {
            const uint32_t i = i0 + lane;
            uint8_t v = 0;
            if (i < n) v = out[o + i - dist];
            __syncwarp();
            if (i < n) out[o + i] = v;
            __syncwarp();
        }
    } else {
        // period replication: byte i equals pattern byte (i mod dist)
        for (uint32_t i = lane; i < n; i += 32) out[o + i] = out[o - dist + (i % dist)];
        __syncwarp();
    }
}

__global__ void __launch_bounds__(128) k_decompress(const ChunkDesc* __restrict__ chunks, uint32_t nchunks, uint32_t* err,
                                                    uint32_t* out_lens) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nchunks) return;
    const ChunkDesc& c = chunks[warp];
    const int lane = threadIdx.x & 31;
    const uint8_t* s = (const uint8_t*)c.src;
    uint8_t* d = (uint8_t*)c.dst;
    const uint32_t n = c.src_len;
    uint32_t o = 0;
    uint32_t fail = 0;
    if (c.codec == 0) {
        if (n > c.dst_cap) fail = ORCB_UNEXPECTED;
        else warp_copy_fwd(d, s, n, lane);
        o = n;
    } else if (c.codec == 2) {
        // Snappy raw block (snap::raw::Decoder, compression.rs:161-172)
        uint32_t p = 0;
        uint64_t ulen = 0;
        for (uint32_t sh = 0;; sh += 7) {
            if (p >= n || sh > 35) { fail = ORCB_BUILD_SNAPPY_DECODER; break; }
            const uint32_t b = s[p++];
            ulen |= (uint64_t)(b & 0x7f) << sh;
            if (b < 0x80) break;
        }
        if (!fail && ulen > c.dst_cap) fail = ORCB_BUILD_SNAPPY_DECODER;
        while (!fail && p < n) {
            const uint32_t tag = s[p++];
            const uint32_t t = tag & 3;
            if (t == 0) {
                uint32_t l = tag >> 2;
                if (l >= 60) {
                    const uint32_t extra = l - 59;
                    if (p + extra > n) { fail = ORCB_BUILD_SNAPPY_DECODER; break; }
                    l = 0;
                    for (uint32_t k = 0; k < extra; k++) l |= (uint32_t)s[p + k] << (8 * k);
                    p += extra;
                }
                l += 1;
                if (p + l > n || (uint64_t)o + l > ulen) { fail = ORCB_BUILD_SNAPPY_DECODER; break; }
                warp_copy_fwd(d + o, s + p, l, lane);
                __syncwarp();
                p += l;
                o += l;
            } else {
                uint32_t l, dist;
                if (t == 1) {
                    if (p + 1 > n) { fail = ORCB_BUILD_SNAPPY_DECODER; break; }
                    l = 4 + ((tag >> 2) & 7);
                    dist = ((tag >> 5) << 8) | s[p];
                    p += 1;
                } else if (t == 2) {
                    if (p + 2 > n) { fail = ORCB_BUILD_SNAPPY_DECODER; break; }
                    l = 1 + (tag >> 2);
                    dist = s[p] | ((uint32_t)s[p + 1] << 8);
                    p += 2;
                } else {
                    if (p + 4 > n) { fail = ORCB_BUILD_SNAPPY_DECODER; break; }
                    l = 1 + (tag >> 2);
                    dist = s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16) | ((uint32_t)s[p + 3] << 24);
                    p += 4;
                }
                if (dist == 0 || dist > o || (uint64_t)o + l > ulen) { fail = ORCB_BUILD_SNAPPY_DECODER; break; }
                warp_copy_match(d, o, dist, l, lane);
                o += l;
            }
        }
        if (!fail && o != ulen) fail = ORCB_BUILD_SNAPPY_DECODER;
    } else {
        // LZ4 block (lz4_flex::block::decompress(src, max), compression.rs:185-195)
        uint32_t p = 0;
        while (!fail && p < n) {
            const uint32_t tok = s[p++];
            uint32_t ll = tok >> 4;
            if (ll == 15) {
                for (;;) {
                    if (p >= n) { fail = ORCB_BUILD_LZ4_DECODER; break; }
                    const uint32_t b = s[p++];
                    ll += b;
                    if (b != 255) break;
                }
                if (fail) break;
            }
            if (p + ll > n || o + ll > c.dst_cap) { fail = ORCB_BUILD_LZ4_DECODER; break; }
            warp_copy_fwd(d + o, s + p, ll, lane);
            __syncwarp();
            p += ll;
            o += ll;
            if (p >= n) break;
            if (p + 2 > n) { fail = ORCB_BUILD_LZ4_DECODER; break; }
            const uint32_t dist = s[p] | ((uint32_t)s[p + 1] << 8);
            p += 2;
            uint32_t ml = tok & 15;
            if (ml == 15) {
                for (;;) {
                    if (p >= n) { fail = ORCB_BUILD_LZ4_DECODER; break; }
                    const uint32_t b = s[p++];
                    ml += b;
                    if (b != 255) break;
                }
                if (fail) break;
            }
            ml += 4;
            if (dist == 0 || dist > o || o + ml > c.dst_cap) { fail = ORCB_BUILD_LZ4_DECODER; break; }
            warp_copy_match(d, o, dist, ml, lane);
            o += ml;
        }
    }
    if (!fail && c.expect_len >= 0 && o != (uint32_t)c.expect_len) fail = ORCB_UNEXPECTED;
    if (lane == 0) {
        if (fail) set_err(err, c.colstripe, fail);
        if (out_lens) out_lens[warp] = o;
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launch wrappers
// ------------------------------------------------------------------------------------------------
static inline uint32_t blocks_for_warps(uint32_t nwarps, uint32_t warps_per_block) {
    return (nwarps + warps_per_block - 1) / warps_per_block;
}

#define LAUNCH_CHECK()                           \
    do {                                         \
        cudaError_t _e = cudaGetLastError();     \
        if (_e != cudaSuccess) return (int)_e;   \
    } while (0)

int launch_rle_index(const Seg* segs, uint32_t n, const uint32_t* cnt, RunRec* table, BlockRec* blocks, uint32_t* nblocks,
                     uint32_t pool_blocks, CoopRec* coop_q, uint32_t* ncoop, uint32_t coop_cap, uint32_t* err, cudaStream_t st) {
    if (!n) return 0;
    const uint32_t nwarps = (n + IDX_LANES - 1) / IDX_LANES;
    k_rle_index<<<(nwarps + 3) / 4, 128, 0, st>>>(segs, n, cnt, table, blocks, nblocks, pool_blocks, coop_q, ncoop, coop_cap, err);
    LAUNCH_CHECK();
    return 0;
}
int launch_int_rle(const Seg* segs, const BlockRec* blocks, const uint32_t* nblocks, uint32_t pool_blocks, const RunRec* table,
                   const uint32_t* cnt, const uint32_t* dstart, uint32_t* err, uint32_t* mis, uint32_t* slow_list,
                   uint32_t* slow_count, const CoopRec* coop_q, const uint32_t* ncoop, uint32_t coop_cap, cudaStream_t st) {
    if (!pool_blocks) return 0;
    // persistent grids: enough CTAs to fill every SM, never more warps than blocks could exist
    static int ctas_fast = 0, ctas_gen = 0;
    if (!ctas_fast) {
        int dev = 0, sms = 148, per_sm = 8;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_int_rle, RLE_WARPS * 32, 0);
        ctas_fast = sms * (per_sm > 0 ? per_sm : 1);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_int_rle_general, RLE_WARPS * 32, 0);
        ctas_gen = sms * (per_sm > 0 ? per_sm : 1);
    }
    const uint64_t need = ((uint64_t)pool_blocks + RLE_WARPS - 1) / RLE_WARPS;
    k_int_rle<<<(uint32_t)std::min<uint64_t>(ctas_fast, need), RLE_WARPS * 32, 0, st>>>(segs, blocks, nblocks, table, cnt, dstart,
                                                                                         err, mis, slow_list, slow_count);
    LAUNCH_CHECK();
    k_int_rle_general<<<(uint32_t)std::min<uint64_t>(ctas_gen, need), RLE_WARPS * 32, 0, st>>>(segs, blocks, slow_list, slow_count,
                                                                                                table, cnt, dstart, err, mis);
    LAUNCH_CHECK();
    if (coop_cap) {
        const uint64_t needq = ((uint64_t)coop_cap + RLE_WARPS - 1) / RLE_WARPS;
        k_coop_runs<<<(uint32_t)std::min<uint64_t>(ctas_gen, needq), RLE_WARPS * 32, 0, st>>>(segs, coop_q, ncoop, coop_cap, cnt, dstart,
                                                                                               err, mis);
        LAUNCH_CHECK();
    }
    return 0;
}
int launch_int_rle_coop(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err,
                        uint32_t* mis, cudaStream_t st) {
    if (!n) return 0;
    k_int_rle_coop<<<blocks_for_warps(n, RLE_WARPS), RLE_WARPS * 32, 0, st>>>(segs, n, cnt, dstart, err, mis);
    LAUNCH_CHECK();
    return 0;
}
int launch_byte_rle(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err,
                    cudaStream_t st) {
    if (!n) return 0;
    k_byte_rle<<<blocks_for_warps(n, RLE_WARPS), RLE_WARPS * 32, 0, st>>>(segs, n, cnt, dstart, err);
    LAUNCH_CHECK();
    return 0;
}
int launch_bits(const BitSeg* segs, uint32_t n, uint32_t* cnt, const uint32_t* dstart, cudaStream_t st) {
    if (!n) return 0;
    k_bits<<<blocks_for_warps(n, 4), 128, 0, st>>>(segs, n, cnt, dstart);
    LAUNCH_CHECK();
    return 0;
}
int launch_seg_scan(const ScanDesc* d, uint32_t n, uint32_t* cnt, uint32_t* dstart, cudaStream_t st) {
    if (!n) return 0;
    k_seg_scan<<<blocks_for_warps(n, 4), 128, 0, st>>>(d, n, cnt, dstart);
    LAUNCH_CHECK();
    return 0;
}
int launch_varint128(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err,
                     cudaStream_t st) {
    if (!n) return 0;
    k_varint128<<<blocks_for_warps(n, 4), 128, 0, st>>>(segs, n, cnt, dstart, err);
    LAUNCH_CHECK();
    return 0;
}
int launch_copy(const CopyDesc* d, const uint2* tiles, uint32_t ntiles, const uint32_t* cnt, uint32_t* err,
                const StrCol* strcols, cudaStream_t st) {
    static_assert(COPY_TILE == U8_TILE, "k_copy sets one UTF-8 flag bit per tile");
    if (!ntiles) return 0;
    k_copy<<<ntiles, 256, 0, st>>>(d, tiles, ntiles, cnt, err, strcols);
    LAUNCH_CHECK();
    return 0;
}
int launch_spaced(const SpacedDesc* d, uint32_t n, const uint32_t* dstart, cudaStream_t st) {
    if (!n) return 0;
    k_spaced<<<blocks_for_warps(n, 4), 128, 0, st>>>(d, n, dstart);
    LAUNCH_CHECK();
    return 0;
}
int launch_decimal_fix(const DecFixDesc* d, uint32_t n, const uint32_t* cnt, const uint32_t* mis, cudaStream_t st) {
    if (!n) return 0;
    k_decimal_fix<<<dim3(64, n), 256, 0, st>>>(d, cnt, mis);
    LAUNCH_CHECK();
    return 0;
}
int launch_timestamp(const TsDesc* d, uint32_t n, const uint32_t* cnt, uint32_t* err, cudaStream_t st) {
    if (!n) return 0;
    k_timestamp<<<dim3(64, n), 256, 0, st>>>(d, cnt, err);
    LAUNCH_CHECK();
    return 0;
}
int launch_utf8(const StrCol* cols, const uint2* tiles, uint32_t ntiles, cudaStream_t st) {
    if (!ntiles) return 0;
    k_utf8<<<ntiles, 256, 0, st>>>(cols, tiles, ntiles);
    LAUNCH_CHECK();
    return 0;
}
int launch_strings(StrCol* cols, uint32_t ncols, uint32_t ntiles, uint32_t* err, JobState* state, uint64_t heap_base,
                   uint64_t heap_cap, uint64_t* ptr_table, cudaStream_t st) {
    if (!ncols) return 0;
    k_dict_prepare<<<blocks_for_warps(ncols, 4), 128, 0, st>>>(cols, ncols, err);
    LAUNCH_CHECK();
    k_str_tile_sum<<<blocks_for_warps(ntiles, 4), 128, 0, st>>>(cols, ncols, ntiles, err);
    LAUNCH_CHECK();
    k_str_tile_scan<<<blocks_for_warps(ncols, 4), 128, 0, st>>>(cols, ncols, err, state, heap_base, heap_cap, ptr_table);
    LAUNCH_CHECK();
    k_str_offsets<<<blocks_for_warps(ntiles, 4), 128, 0, st>>>(cols, ncols, ntiles, err);
    LAUNCH_CHECK();
    k_utf8_bounds<<<ncols, 128, 0, st>>>(cols, ncols, err);
    LAUNCH_CHECK();
    return 0;
}
int launch_repack(const RepackDesc* d, uint32_t ndesc, uint32_t nwork, uint32_t* nulls, cudaStream_t st) {
    if (!nwork) return 0;
    k_repack<<<blocks_for_warps(nwork, 4), 128, 0, st>>>(d, ndesc, nwork, nulls);
    LAUNCH_CHECK();
    return 0;
}
int launch_decompress(const ChunkDesc* c, uint32_t n, uint32_t* err, uint32_t* out_lens, cudaStream_t st) {
    if (!n) return 0;
    k_decompress<<<blocks_for_warps(n, 4), 128, 0, st>>>(c, n, err, out_lens);
    LAUNCH_CHECK();
    return 0;
}

}  // namespace orcb
