// Shared device helpers of the decode kernels (sm_100a): error words, bit loads, warp scans, the per-segment
// context and value stores, the constant-phase DIRECT unpack loop, UTF-8 checks, launch helpers.
// Included by the k_*.cu translation units; everything here is inline.
#pragma once
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
    // decimal SECONDARY streams are almost always one constant (the type's scale): their long constant runs are
    // only compared, not written, unless a value differed somewhere in the column-stripe (second pass of the kernel)
    bool scale_lazy = false;
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


constexpr int RLE_WARPS = 4;  // warps per CTA of the run-length kernels

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
// host-side launch helpers
// ------------------------------------------------------------------------------------------------
static inline uint32_t blocks_for_warps(uint32_t nwarps, uint32_t warps_per_block) {
    return (nwarps + warps_per_block - 1) / warps_per_block;
}

#define LAUNCH_CHECK()                           \
    do {                                         \
        cudaError_t _e = cudaGetLastError();     \
        if (_e != cudaSuccess) return (int)_e;   \
    } while (0)

}  // namespace orcb
