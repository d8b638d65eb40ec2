// Compression chunks: original copies, Snappy and LZ4 block decoding.
#include "kernel_util.cuh"
#include "zstd_dec.h"

namespace orcb {

// ------------------------------------------------------------------------------------------------
// Compression chunks (src/compression.rs:113-123, 244-275): original chunks are copied, Snappy and LZ4
// blocks are decoded by one warp per chunk.
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

// tag byte -> header bytes [0:3] | literal [3] | length is in the following bytes [4] | bytes of offset / length
// that follow [5:8] | high offset bits of a 1-byte-offset copy [8:11] | length [16:]
__device__ __forceinline__ uint32_t snappy_tag_entry(uint32_t tag) {
    const uint32_t t = tag & 3u, l = tag >> 2;
    if (t == 0) return l < 60 ? (1u | 8u | ((l + 1u) << 16)) : ((1u + (l - 59u)) | 8u | 16u | ((l - 59u) << 5));
    if (t == 1) return 2u | (1u << 5) | ((tag >> 5) << 8) | ((4u + (l & 7u)) << 16);
    if (t == 2) return 3u | (2u << 5) | ((l + 1u) << 16);
    return 5u | (4u << 5) | ((l + 1u) << 16);
}

// uncompressed length preamble (snap::raw::decompress_len, compression.rs:163-165)
__device__ __forceinline__ uint32_t snappy_preamble(const uint8_t* __restrict__ s, uint32_t n, uint32_t dst_cap, uint32_t& p, uint64_t& ulen) {
    p = 0;
    ulen = 0;
    for (uint32_t sh = 0;; sh += 7) {
        if (p >= n || sh > 35) return ORCB_BUILD_SNAPPY_DECODER;
        const uint32_t b = s[p++];
        ulen |= (uint64_t)(b & 0x7f) << sh;
        if (b < 0x80) break;
    }
    return ulen > dst_cap ? (uint32_t)ORCB_BUILD_SNAPPY_DECODER : 0u;
}

__device__ __forceinline__ void chunk_done(const ChunkDesc& c, uint32_t fail, uint32_t o, uint32_t* err, uint32_t* out_lens,
                                           uint32_t* retry, int lane) {
    bool again = false;
    if (!fail && c.expect_len >= 0 && o != (uint32_t)c.expect_len) {
        if (c.assumed && retry) again = true;  // the host laid the stream out on a guess: it re-plans with the real sizes
        else fail = ORCB_UNEXPECTED;
    }
    if (lane == 0) {
        if (fail) set_err(err, c.colstripe, fail);
        if (again) atomicOr(retry, 1u);
        if (out_lens) out_lens[c.id] = fail ? 0xffffffffu : o;
    }
}

// ------------------------------------------------------------------------------------------------
// Tile decoder (Snappy and LZ4): lane-parallel parsing and lane-per-element copying.
//
// A compressed chunk is one chain of elements (Snappy: literal | copy; LZ4: sequence = literals + match), each two or
// three input bytes on average, so any scheme that spends a warp instruction per element and phase is issue-bound at
// a few hundred MB/s per warp.  Here a warp takes LZ_TILE (2 KiB) of input at a time:
//   1. the tile is staged in shared memory, one padded row per 64-byte sub-block (rows overlap by 12 bytes so that a
//      header is always read inside one row; 19-word rows keep the 32 lanes on different banks);
//   2. SPECULATIVE WALK: every lane walks the elements of its own sub-block from the sub-block's first byte, which is
//      usually not an element start, and remembers which positions it visited.  LZ streams re-synchronise within a
//      few elements, so when the lane before hands over the true entry position it is almost always one the lane
//      has visited: everything from there on (and the lane's exit position) is then already right.  The hand-over
//      runs lane by lane (32 short steps); a lane whose walk never met the true chain re-walks alone;
//   3. every lane walks its now certain elements twice more: once to count elements and output bytes (warp scan ->
//      rank and output position of each lane's first element), once to write one 32-bit descriptor per element;
//   4. the descriptors are consumed 32 at a time, ONE ELEMENT PER LANE: a lane copies its own element (<= 32 bytes)
//      from the staged input (literal) or from a 4 KiB ring of recent output (older bytes from global memory);
//      elements that read the output of their own group, and elements of 33..64 bytes, follow in order, copied by
//      the whole warp; the group's output range is then written out from the ring in whole words.
// Elements longer than 64 bytes, literals that run out of the staged tile and anything irregular are decoded by the
// element-by-element loop (one element, or - for damaged input - the rest of the chunk, so errors are reported
// exactly as by the serial decoder).
// ------------------------------------------------------------------------------------------------
#ifndef ORCB_LZ_SB
#define ORCB_LZ_SB 64   // sub-block bytes per lane: 64 (2 KiB tiles) or 32 (1 KiB tiles, less shared memory, more warps)
#endif
#ifndef ORCB_LZ_CTAS
#define ORCB_LZ_CTAS 5  // resident CTAs per SM the register allocation aims for
#endif
constexpr uint32_t LZ_SB = ORCB_LZ_SB, LZ_SHIFT = LZ_SB == 64 ? 6 : 5, LZ_TILE = 32 * LZ_SB;
// rows overlap by 12 bytes; 19 (or 11) words per row keep the 32 lanes' rows on different banks
constexpr uint32_t LZ_ROW_WORDS = LZ_SB / 4 + 3, LZ_HIST = 4096, LZ_DESC_CAP = 16 * LZ_SB;
static_assert(LZ_SB == 64 || LZ_SB == 32, "sub-blocks of 32 or 64 bytes");
constexpr uint32_t LZ_INLINE = 32, LZ_MAX_ELEM = 64, LZ_MAX_MATCH = 1024;
constexpr uint32_t LZ_INVALID = 0xffffffffu;
constexpr uint32_t LZD_LIT = 1u << 14;  // descriptor: len [0:14) | literal [14] | src [16:32) (literal: tile position; match: distance)

struct LzWarp {
    uint32_t in[32 * LZ_ROW_WORDS];
    uint32_t desc[LZ_DESC_CAP];
    uint8_t hist[LZ_HIST];
};

// one element (Snappy) / one sequence (LZ4) as the walks see it
struct LzElem {
    uint32_t adv;      // input bytes to the next element (LZ_INVALID: cannot be handled here)
    uint32_t lit_len;  // literal bytes (0: none)
    uint32_t lit_pos;  // absolute input position of the first literal byte
    uint32_t m_len;    // match bytes (0: none)
    uint32_t m_dist;
};

// 8 input bytes at tile position x (x < LZ_TILE + 4) out of the padded rows
__device__ __forceinline__ uint64_t lz_window(const uint32_t* in, uint32_t x) {
    const uint32_t row = min(x >> LZ_SHIFT, 31u), col = x - (row << LZ_SHIFT);  // col < row bytes - 8
    const uint32_t* r = in + row * LZ_ROW_WORDS + (col >> 2);
    const uint32_t sh = (col & 3) * 8;
    const uint32_t w0 = r[0], w1 = r[1], w2 = r[2];
    return ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
}
__device__ __forceinline__ uint32_t lz_byte(const uint32_t* in, uint32_t x) {
    const uint32_t row = min(x >> LZ_SHIFT, 31u), col = x - (row << LZ_SHIFT);
    return (in[row * LZ_ROW_WORDS + (col >> 2)] >> ((col & 3) * 8)) & 0xffu;
}
// 8 input bytes at absolute position q: from the staged tile when they lie in it, else from global memory
__device__ __forceinline__ uint64_t lz_fetch(const uint32_t* in, const uint8_t* __restrict__ s, uint32_t tb, uint32_t q) {
    if (q >= tb && q - tb < LZ_TILE + 4) return lz_window(in, q - tb);
    uint64_t w = 0;
    for (int k = 0; k < 8; k++) w |= (uint64_t)s[q + k] << (8 * k);  // arenas are padded: reading past the chunk is safe
    return w;
}

// DIST = false: lengths and `adv` only (the walks that do not write descriptors); most elements then need no more
// than their first byte
template <int CODEC, bool DIST = true>
__device__ __forceinline__ LzElem lz_decode(const uint32_t* in, const uint8_t* __restrict__ s, uint32_t n, uint32_t tb, uint32_t q,
                                            const uint32_t* lut) {
    LzElem e;
    e.lit_len = e.m_len = e.m_dist = 0;
    e.lit_pos = 0;
    if (!DIST) {
        const uint32_t b0 = lz_byte(in, q - tb);
        if (CODEC == 2) {
            const uint32_t t = lut[b0];
            if (!((t >> 4) & 1u)) {  // not a literal with length bytes
                const uint32_t hdr = t & 7u;
                if ((t >> 3) & 1u) {
                    e.lit_len = t >> 16;
                    e.lit_pos = q + hdr;
                    e.adv = hdr + e.lit_len;
                } else {
                    e.m_len = t >> 16;
                    e.adv = hdr;
                }
                if (e.adv > n - q) e.adv = LZ_INVALID;
                return e;
            }
        } else if ((b0 >> 4) != 15u && (b0 & 15u) != 15u) {
            const uint32_t ll = b0 >> 4;
            if (ll > n - (q + 1)) { e.adv = LZ_INVALID; return e; }
            e.lit_len = ll;
            e.lit_pos = q + 1;
            if (q + 1 + ll >= n) { e.adv = 1 + ll; return e; }  // last sequence: literals only
            if (q + 1 + ll + 2 > n) { e.adv = LZ_INVALID; return e; }
            e.m_len = (b0 & 15u) + 4u;
            e.adv = 1 + ll + 2;
            return e;
        }
    }
    const uint64_t w = lz_window(in, q - tb);
    if (CODEC == 2) {
        const uint32_t lo = (uint32_t)w, b4 = (uint32_t)(w >> 32) & 0xffu;
        const uint32_t t = lut[lo & 0xffu];
        const uint32_t hdr = t & 7u, nbytes = (t >> 5) & 7u;
        const bool lit = (t >> 3) & 1u;
        const uint32_t raw = (lo >> 8) | (b4 << 24);
        const uint32_t field = raw & __funnelshift_rc(0xffffffffu, 0u, 32u - 8u * nbytes);
        if (lit) {
            const bool longl = (t >> 4) & 1u;
            if (longl && field >= 0x40000000u) { e.adv = LZ_INVALID; return e; }
            e.lit_len = (t >> 16) + (longl ? field + 1u : 0u);
            e.lit_pos = q + hdr;
            e.adv = hdr + e.lit_len;
        } else {
            e.m_len = t >> 16;
            e.m_dist = field | (((t >> 8) & 7u) << 8);
            e.adv = hdr;
        }
        if (e.adv > n - q) e.adv = LZ_INVALID;  // runs past the end of the chunk
        return e;
    }
    // LZ4: token, literal-length bytes, literals, 2 offset bytes, match-length bytes
    const uint32_t tok = (uint32_t)w & 0xffu;
    uint32_t ll = tok >> 4, p = q + 1;
    if (ll == 15) {
        uint32_t k = 1;
        for (;;) {
            if (k >= 7) { e.adv = LZ_INVALID; return e; }  // length bytes beyond the window: a very long literal
            const uint32_t b = (uint32_t)(w >> (8 * k)) & 0xffu;
            k++;
            ll += b;
            if (b != 255) break;
        }
        p = q + k;
    }
    if (p > n || ll > n - p) { e.adv = LZ_INVALID; return e; }
    e.lit_len = ll;
    e.lit_pos = p;
    p += ll;
    if (p >= n) {  // last sequence: literals only
        e.adv = p - q;
        return e;
    }
    if (p + 2 > n) { e.adv = LZ_INVALID; return e; }
    const uint64_t m = lz_fetch(in, s, tb, p);
    e.m_dist = (uint32_t)m & 0xffffu;
    uint32_t ml = tok & 15u;
    p += 2;
    if (ml == 15) {
        uint32_t k = 2;
        for (;;) {
            if (k >= 8 || p >= n) { e.adv = LZ_INVALID; return e; }
            const uint32_t b = (uint32_t)(m >> (8 * k)) & 0xffu;
            k++;
            p++;
            ml += b;
            if (b != 255) break;
        }
    }
    e.m_len = ml + 4;
    e.adv = p - q;
    return e;
}

// refill the ring with the last bytes of output (after an element was written to global memory only)
__device__ __forceinline__ void lz_hist_reload(uint8_t* hist, const uint8_t* d, uint32_t o, int lane) {
    const uint32_t from = o > LZ_HIST ? o - LZ_HIST : 0u;
    if ((((uintptr_t)d) & 3) == 0) {
        // whole words (a word that reaches behind o brings bytes the ring does not hold yet: they are written before read)
        for (uint32_t a = (from & ~3u) + 4u * lane; a < o; a += 128)
            *(uint32_t*)(hist + (a & (LZ_HIST - 1))) = __ldcg((const uint32_t*)(d + a));
    } else {
        for (uint32_t a = from + lane; a < o; a += 32) hist[a & (LZ_HIST - 1)] = __ldcg(d + a);
    }
    __syncwarp();
}

// One element / sequence by the whole warp, straight to global memory: what the tiles leave to it (long elements,
// literals that run out of the staged input).  Returns a status; p and o advance.
template <int CODEC>
__device__ __forceinline__ uint32_t lz_serial_step(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint64_t ulen, uint32_t& p,
                                                   uint32_t& o, int lane) {
    const uint32_t bad = CODEC == 2 ? (uint32_t)ORCB_BUILD_SNAPPY_DECODER : (uint32_t)ORCB_BUILD_LZ4_DECODER;
    if (CODEC == 2) {
        const uint32_t tag = s[p++];
        const uint32_t t = tag & 3;
        if (t == 0) {
            uint64_t l = tag >> 2;
            if (l >= 60) {
                const uint32_t extra = (uint32_t)l - 59;
                if (extra > n - p) return bad;
                l = 0;
                for (uint32_t k = 0; k < extra; k++) l |= (uint64_t)s[p + k] << (8 * k);
                p += extra;
            }
            l += 1;
            if (l > n - p || (uint64_t)o + l > ulen) return bad;
            warp_copy_fwd(d + o, s + p, (uint32_t)l, lane);
            __syncwarp();
            p += (uint32_t)l;
            o += (uint32_t)l;
            return 0;
        }
        uint32_t l, dist;
        if (t == 1) {
            if (p + 1 > n) return bad;
            l = 4 + ((tag >> 2) & 7);
            dist = ((tag >> 5) << 8) | s[p];
            p += 1;
        } else if (t == 2) {
            if (p + 2 > n) return bad;
            l = 1 + (tag >> 2);
            dist = s[p] | ((uint32_t)s[p + 1] << 8);
            p += 2;
        } else {
            if (p + 4 > n) return bad;
            l = 1 + (tag >> 2);
            dist = s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16) | ((uint32_t)s[p + 3] << 24);
            p += 4;
        }
        if (dist == 0 || dist > o || (uint64_t)o + l > ulen) return bad;
        warp_copy_match(d, o, dist, l, lane);
        o += l;
        return 0;
    }
    const uint32_t tok = s[p++];
    uint64_t ll = tok >> 4;
    if (ll == 15) {
        for (;;) {
            if (p >= n) return bad;
            const uint32_t b = s[p++];
            ll += b;
            if (b != 255) break;
        }
    }
    if (ll > n - p || (uint64_t)o + ll > ulen) return bad;
    warp_copy_fwd(d + o, s + p, (uint32_t)ll, lane);
    __syncwarp();
    p += (uint32_t)ll;
    o += (uint32_t)ll;
    if (p >= n) return 0;
    if (p + 2 > n) return bad;
    const uint32_t dist = s[p] | ((uint32_t)s[p + 1] << 8);
    p += 2;
    uint64_t ml = tok & 15;
    if (ml == 15) {
        for (;;) {
            if (p >= n) return bad;
            const uint32_t b = s[p++];
            ml += b;
            if (b != 255) break;
        }
    }
    ml += 4;
    if (dist == 0 || dist > o || (uint64_t)o + ml > ulen) return bad;
    warp_copy_match(d, o, dist, (uint32_t)ml, lane);
    o += (uint32_t)ml;
    return 0;
}

// the rest of a chunk, element by element (damaged input is reported from here)
template <int CODEC>
__device__ __forceinline__ uint32_t lz_serial_rest(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint64_t ulen, uint32_t p,
                                                   uint32_t& o, int lane) {
    while (p < n) {
        const uint32_t st = lz_serial_step<CODEC>(s, n, d, ulen, p, o, lane);
        if (st) return st;
    }
    if (CODEC == 2 && o != ulen) return ORCB_BUILD_SNAPPY_DECODER;
    return 0;
}

enum : uint32_t { LZT_OK = 0, LZT_CUT = 1, LZT_IRREGULAR = 2 };

// One tile starting at input position p / output position o.  On LZT_OK / LZT_CUT p and o are behind the last element
// the tile decoded (CUT: the element at p is for lz_serial_step); on LZT_IRREGULAR nothing was decoded.
template <int CODEC>
__device__ __forceinline__ uint32_t lz_tile(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint64_t ulen, uint32_t& p,
                                            uint32_t& o, LzWarp& sm, const uint32_t* lut, int lane, uint32_t& cut_lane_out) {
    const uint32_t tb = p;
    // ---- 1. stage the tile: row r = input bytes [tb + 64 r, tb + 64 r + 76)
    {
        const uintptr_t a0 = (uintptr_t)(s + tb);
        const uint32_t* g = (const uint32_t*)(a0 & ~(uintptr_t)3);
        const uint32_t sh = (uint32_t)(a0 & 3) * 8;
#pragma unroll
        for (uint32_t it = 0; it < LZ_ROW_WORDS; it++) {
            const uint32_t k = it * 32 + (uint32_t)lane;
            const uint32_t row = k / LZ_ROW_WORDS, j = k - row * LZ_ROW_WORDS;
            const uint32_t wi = row * (LZ_SB / 4) + j;
            // nothing is read more than a few bytes behind the chunk (the arenas are padded, but not by a tile)
            sm.in[k] = tb + 4u * wi < n + 8u ? __funnelshift_r(__ldg(g + wi), __ldg(g + wi + 1), sh) : 0u;
        }
    }
    __syncwarp();
    const uint32_t s_i = tb + (uint32_t)lane * LZ_SB;
    const uint32_t end_i = min(s_i + LZ_SB, n);
    // ---- 2. speculative walk: positions visited from the sub-block's first byte
    uint64_t visited = 0;
    uint32_t exit_i = s_i;
    if (s_i < n) {
        uint32_t q = s_i;
        while (q < end_i) {
            visited |= 1ull << (q - s_i);
            const LzElem e = lz_decode<CODEC, false>(sm.in, s, n, tb, q, lut);
            if (e.adv == LZ_INVALID) { q = LZ_INVALID; break; }
            q += e.adv;
        }
        exit_i = q;
    }
    // ---- hand-over of the true entry positions, lane by lane.  A lane whose walk never met the position it is handed
    //      walks again from there; all such lanes do so at once, then the hand-over is repeated (the first of them had
    //      the right entry for certain, so every round settles at least one lane; usually one round settles all).
    uint32_t my_entry = LZ_INVALID, final_exit = tb;
    for (int round = 0; round < 34; round++) {
        uint32_t cur = tb;
        bool flagged = false;
        for (int i = 0; i < 32; i++) {
            uint32_t my_exit = cur;
            if (lane == i) {
                my_entry = cur;
                if (cur != LZ_INVALID && cur < end_i) {
                    my_exit = exit_i;
                    flagged = !((visited >> (cur - s_i)) & 1ull);
                }
            }
            cur = __shfl_sync(FULL, my_exit, i);
        }
        final_exit = cur;
        if (!__any_sync(FULL, flagged)) break;
        if (flagged) {
            uint32_t q = my_entry;
            uint64_t mine = 0;
            while (q < end_i) {
                if ((visited >> (q - s_i)) & 1ull) break;
                mine |= 1ull << (q - s_i);
                const LzElem e = lz_decode<CODEC, false>(sm.in, s, n, tb, q, lut);
                if (e.adv == LZ_INVALID) { q = LZ_INVALID; break; }
                q += e.adv;
            }
            if (q != LZ_INVALID && q < end_i) {  // merged: the old walk is right from here on
                visited = mine | (visited & ~((1ull << (q - s_i)) - 1ull));
            } else {
                visited = mine;
                exit_i = q;
            }
        }
    }
    // ---- 3a. count: elements and output bytes of every lane; a lane stops in front of what the tile cannot take
    uint32_t cnt = 0, bytes = 0, stop = LZ_INVALID;  // stop: position of the element the lane stopped at
    bool irregular = my_entry == LZ_INVALID;
    if (!irregular && my_entry < end_i) {
        uint32_t q = my_entry;
        while (q < end_i) {
            const LzElem e = lz_decode<CODEC, false>(sm.in, s, n, tb, q, lut);
            if (e.adv == LZ_INVALID || e.m_len > LZ_MAX_MATCH || (e.lit_len && e.lit_pos + e.lit_len > tb + LZ_TILE + 8)) {
                stop = q;
                break;
            }
            // long literals: pieces of LZ_INLINE bytes; long matches: pieces of LZ_MAX_ELEM bytes at the same distance
            cnt += (e.lit_len + LZ_INLINE - 1) / LZ_INLINE + (e.m_len + LZ_MAX_ELEM - 1) / LZ_MAX_ELEM;
            bytes += e.lit_len + e.m_len;
            q += e.adv;
        }
    }
    // the tile ends in front of the first stop; lanes behind it are dropped
    const uint32_t stopmask = __ballot_sync(FULL, stop != LZ_INVALID || irregular);
    int cut_lane = stopmask ? __ffs(stopmask) - 1 : 32;
    if (lane > cut_lane) { cnt = 0; bytes = 0; }
    const bool cut_irregular = cut_lane < 32 && __shfl_sync(FULL, (int)irregular, cut_lane & 31);
    uint32_t cnt_incl = warp_incl_scan(cnt, lane), bytes_incl = warp_incl_scan(bytes, lane);
    uint32_t total_cnt = __shfl_sync(FULL, cnt_incl, 31), total_bytes = __shfl_sync(FULL, bytes_incl, 31);
    if (total_cnt > LZ_DESC_CAP) {
        // more pieces than the descriptor table holds (LZ4 sequences of long matches): the tile ends in front of the lane
        // that would overflow it
        const int lo = __ffs(__ballot_sync(FULL, cnt_incl > LZ_DESC_CAP)) - 1;
        if (lane >= lo) { cnt = 0; bytes = 0; }
        if (lane == lo) stop = my_entry;
        cut_lane = lo;
        cnt_incl = warp_incl_scan(cnt, lane);
        bytes_incl = warp_incl_scan(bytes, lane);
        total_cnt = __shfl_sync(FULL, cnt_incl, 31);
        total_bytes = __shfl_sync(FULL, bytes_incl, 31);
    }
    cut_lane_out = (uint32_t)cut_lane;
    if (cut_irregular && total_cnt == 0) return LZT_IRREGULAR;
    if ((uint64_t)o + total_bytes > ulen) return LZT_IRREGULAR;  // output past the announced length: the serial loop reports it
    // ---- 3b. emit the descriptors
    bool bad = false;
    if (cnt) {
        uint32_t q = my_entry, r = cnt_incl - cnt, op = o + (bytes_incl - bytes);
        const uint32_t lim = stop != LZ_INVALID ? stop : end_i;
        while (q < lim) {
            const LzElem e = lz_decode<CODEC>(sm.in, s, n, tb, q, lut);
            for (uint32_t done = 0; done < e.lit_len; done += LZ_INLINE)
                sm.desc[r++] = min(e.lit_len - done, LZ_INLINE) | LZD_LIT | ((e.lit_pos + done - tb) << 16);
            op += e.lit_len;
            if (e.m_len) {
                if (e.m_dist == 0 || e.m_dist > op || e.m_dist > 0xffffu) bad = true;
                for (uint32_t done = 0; done < e.m_len; done += LZ_MAX_ELEM) sm.desc[r++] = min(e.m_len - done, LZ_MAX_ELEM) | (e.m_dist << 16);
                op += e.m_len;
            }
            q += e.adv;
        }
    }
    if (__any_sync(FULL, bad)) return LZT_IRREGULAR;
    __syncwarp();
    // where the next tile (or the serial step) starts
    uint32_t next_p;
    if (cut_lane < 32) next_p = __shfl_sync(FULL, stop, cut_lane);
    else next_p = final_exit;
    // ---- 4. copy, 32 elements at a time
    const bool words_ok = (((uintptr_t)d) & 3) == 0;
    uint32_t gO = o;
    for (uint32_t g0 = 0; g0 < total_cnt; g0 += 32) {
        const uint32_t di = g0 + (uint32_t)lane;
        const uint32_t desc = di < total_cnt ? sm.desc[di] : 0u;
        const uint32_t len = desc & 0x3fffu, src = desc >> 16;
        const bool lit = (desc & LZD_LIT) != 0;
        const uint32_t incl = warp_incl_scan(len, lane);
        const uint32_t T = __shfl_sync(FULL, incl, 31);
        const uint32_t op = gO + incl - len, gEnd = gO + T;
        // Rounds: every element whose source lies in front of the first unfinished element is copied by its own lane, all
        // of them at once; an element that needs the whole warp (33..64 bytes, or a match that overlaps its own output)
        // is copied when it is the first unfinished one.  Every round finishes at least that first element.
        const bool big = len > LZ_INLINE || (!lit && src < len);
        const uint32_t need = (len && !lit) ? op - src + min(len, src) : 0u;  // end of the source range
        uint32_t rem = __ballot_sync(FULL, len != 0);
        while (rem) {
            const int first = __ffs(rem) - 1;
            const uint32_t resolved = __shfl_sync(FULL, op, first);
            const bool ready = ((rem >> lane) & 1u) && !big && need <= resolved;
            const uint32_t rmask = __ballot_sync(FULL, ready);
            // (with only a few ready elements the whole-warp copy of the first one is cheaper than a round)
            if (__popc(rmask) >= 4) {
                rem &= ~rmask;
                const uint32_t maxlen = __reduce_max_sync(FULL, ready ? len : 0u);
                if (ready) {
                    // All source bytes exist: aligned words are loaded back to back, shifted into place and stored.
                    // No load waits for a store, so a lane's bytes cost one memory latency.
                    uint32_t w[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                    uint32_t sh;
                    if (lit) {
                        sh = (src & 3) * 8;
#pragma unroll
                        for (int j = 0; j < 9; j++) {
                            if (4u * j >= maxlen + 3u) break;
                            const uint32_t x = (src & ~3u) + 4u * j;
                            const uint32_t rr = min(x >> LZ_SHIFT, 31u), cc = x - (rr << LZ_SHIFT);
                            w[j] = 4u * j < len + (src & 3) ? sm.in[rr * LZ_ROW_WORDS + (cc >> 2)] : 0u;
                        }
                    } else {
                        const uint32_t a0 = op - src;
                        sh = (a0 & 3) * 8;
                        if (gEnd - a0 <= LZ_HIST) {
                            const uint32_t* hw = (const uint32_t*)sm.hist;
#pragma unroll
                            for (int j = 0; j < 9; j++) {
                                if (4u * j >= maxlen + 3u) break;
                                w[j] = 4u * j < len + (a0 & 3) ? hw[((a0 >> 2) + j) & (LZ_HIST / 4 - 1)] : 0u;
                            }
                        } else if (words_ok) {
                            const uint32_t* gw = (const uint32_t*)(d + (a0 & ~3u));
#pragma unroll
                            for (int j = 0; j < 9; j++) {
                                if (4u * j >= maxlen + 3u) break;
                                w[j] = 4u * j < len + (a0 & 3) ? __ldcg(gw + j) : 0u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 9; j++) {
                                uint32_t v = 0;
                                if (4u * j < len + (a0 & 3))
                                    for (int b = 0; b < 4; b++) v |= (uint32_t)__ldcg(d + (a0 & ~3u) + 4 * j + b) << (8 * b);
                                w[j] = v;
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; j++) {
                        if (4u * j >= maxlen) break;
                        const uint32_t v = __funnelshift_r(w[j], w[j + 1], sh);
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (4u * j + b < len) sm.hist[(op + 4u * j + b) & (LZ_HIST - 1)] = (uint8_t)(v >> (8 * b));
                    }
                }
                __syncwarp();
                continue;
            }
            // the first unfinished element needs the whole warp
            rem &= ~(1u << first);
            const uint32_t e_len = __shfl_sync(FULL, len, first), e_src = __shfl_sync(FULL, src, first), e_op = resolved;
            const bool e_lit = __shfl_sync(FULL, (int)lit, first);
            if (e_lit) {
                for (uint32_t k = lane; k < e_len; k += 32) sm.hist[(e_op + k) & (LZ_HIST - 1)] = (uint8_t)lz_byte(sm.in, e_src + k);
            } else if (e_src >= e_len || e_src >= 32) {
                // source and destination do not overlap within one step of 32 bytes
                for (uint32_t k0 = 0; k0 < e_len; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    uint8_t b = 0;
                    if (k < e_len) {
                        const uint32_t a = e_op - e_src + k;
                        b = (gEnd - a <= LZ_HIST) ? sm.hist[a & (LZ_HIST - 1)] : __ldcg(d + a);
                    }
                    __syncwarp();
                    if (k < e_len) sm.hist[(e_op + k) & (LZ_HIST - 1)] = b;
                    __syncwarp();
                }
            } else {
                // short period: byte k repeats byte k mod dist of the source
                const uint32_t a0 = e_op - e_src;
                uint8_t pat = 0;
                if ((uint32_t)lane < e_src) {
                    const uint32_t a = a0 + lane;
                    pat = (gEnd - a <= LZ_HIST) ? sm.hist[a & (LZ_HIST - 1)] : __ldcg(d + a);
                }
                __syncwarp();
                for (uint32_t k0 = 0; k0 < e_len; k0 += 32) {
                    const uint32_t k = k0 + lane;
                    const uint8_t b = (uint8_t)__shfl_sync(FULL, (uint32_t)pat, (int)(k % e_src));
                    if (k < e_len) sm.hist[(e_op + k) & (LZ_HIST - 1)] = b;
                }
            }
            __syncwarp();
        }
        // write the group's bytes out of the ring
        if (words_ok) {
            const uint32_t a_head = min((gO + 3u) & ~3u, gEnd), a_tail = max(gEnd & ~3u, a_head);
            if (gO + lane < a_head) d[gO + lane] = sm.hist[(gO + lane) & (LZ_HIST - 1)];
            for (uint32_t a = a_head + 4u * lane; a < a_tail; a += 128)
                *(uint32_t*)(d + a) = *(const uint32_t*)(sm.hist + (a & (LZ_HIST - 1)));
            if (a_tail + lane < gEnd) d[a_tail + lane] = sm.hist[(a_tail + lane) & (LZ_HIST - 1)];
        } else {
            for (uint32_t a = gO + lane; a < gEnd; a += 32) d[a] = sm.hist[a & (LZ_HIST - 1)];
        }
        __syncwarp();
        gO = gEnd;
    }
    o = gO;
    p = next_p;
    return cut_lane < 32 ? LZT_CUT : LZT_OK;
}

template <int CODEC>
__device__ __forceinline__ uint32_t lz_chunk(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint64_t ulen, uint32_t p, uint32_t& o,
                                             LzWarp& sm, const uint32_t* lut, int lane) {
    uint32_t serial_budget = 0, early_cuts = 0;
    while (p < n) {
        if (serial_budget) {
            const uint32_t st = lz_serial_step<CODEC>(s, n, d, ulen, p, o, lane);
            if (st) return st;
            if (--serial_budget == 0) lz_hist_reload(sm.hist, d, o, lane);
            continue;
        }
        uint32_t cut_lane = 32;
        const uint32_t r = lz_tile<CODEC>(s, n, d, ulen, p, o, sm, lut, lane, cut_lane);
        if (r == LZT_IRREGULAR) return lz_serial_rest<CODEC>(s, n, d, ulen, p, o, lane);
        if (r == LZT_CUT) {
            // data made of long elements: the more often tiles end early, the longer the serial loop keeps the chunk
            early_cuts = cut_lane < 8 ? min(early_cuts + 1, 6u) : 0u;
            serial_budget = 1u << early_cuts;
        } else {
            early_cuts = 0;
        }
    }
    if (CODEC == 2 && o != ulen) return ORCB_BUILD_SNAPPY_DECODER;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Zlib chunks: raw DEFLATE (src/compression.rs:142-149, flate2::read::DeflateDecoder).
//
// Huffman decoding is a serial bit chain; one lane walks it and queues up to 32 tokens (a literal byte or a match),
// then the warp places them: a scan of the token lengths gives every token its output position, all literals are
// written at once, matches follow in order, copied by the whole warp.  The tables of a block are built by the whole
// warp: canonical codes come from per-length counters (rank among the symbols of the same length through
// __match_any_sync), every lane fills the lookup entries of its own symbols.  Codes longer than the lookup index are
// resolved bit by bit against the canonical first-code table.
// ------------------------------------------------------------------------------------------------
constexpr int INF_LBITS = 10, INF_DBITS = 8;
struct InfWarp {
    uint16_t llut[1 << INF_LBITS];   // (symbol << 4) | code length; 0 = longer than INF_LBITS or unused
    uint16_t dlut[1 << INF_DBITS];
    uint16_t lsorted[288], dsorted[32];  // symbols ordered by (length, symbol): canonical decode of long codes
    uint16_t lfirst[16], dfirst[16];     // first canonical code of each length
    uint16_t loffs[16], doffs[16];       // index of that code's symbol in *sorted
    uint16_t lcount[16], dcount[16];
    uint8_t lens[320];                   // literal/length code lengths at 0.., distance code lengths at 288..
    uint8_t stage[352];                  // dynamic blocks: 19 code-length lengths, then (from 32 on) the lengths as they are read
    uint32_t tok[32];                    // literal: byte | 1 << 31; match: len | dist << 9  (len <= 258, dist <= 32768)
};
static_assert(sizeof(InfWarp) <= sizeof(LzWarp), "the inflate tables reuse the LZ decoder's shared memory");

__device__ __forceinline__ uint32_t brev_n(uint32_t v, int n) { return __brev(v) >> (32 - n); }

// canonical Huffman tables of `nsym` symbols with the given code lengths; false = over-subscribed
__device__ __forceinline__ bool inf_build(const uint8_t* lens, int nsym, uint16_t* lut, int lut_bits, uint16_t* sorted, uint16_t* first,
                                          uint16_t* offs, uint16_t* count, int lane) {
    for (int i = lane; i < (1 << lut_bits); i += 32) lut[i] = 0;
    if (lane < 16) count[lane] = 0;
    __syncwarp();
    for (int b = 0; b < nsym; b += 32) {
        const int sym = b + lane;
        const int L = sym < nsym ? lens[sym] : 0;
        const uint32_t same = __match_any_sync(FULL, L);
        if (L && lane == __ffs(same) - 1) count[L] += (uint16_t)__popc(same);
        __syncwarp();
    }
    // first codes / offsets (15 serial steps), over-subscription check
    bool ok = true;
    if (lane == 0) {
        uint32_t code = 0, off = 0;
        int left = 1;
        for (int L = 1; L <= 15; L++) {
            left = (left << 1) - (int)count[L];
            if (left < 0) ok = false;
            first[L] = (uint16_t)code;
            offs[L] = (uint16_t)off;
            code = (code + count[L]) << 1;
            off += count[L];
        }
        first[0] = offs[0] = 0;
    }
    ok = __shfl_sync(FULL, (int)ok, 0);
    if (!ok) return false;
    if (lane < 16) count[lane] = 0;  // reused as running counters
    __syncwarp();
    for (int b = 0; b < nsym; b += 32) {
        const int sym = b + lane;
        const int L = sym < nsym ? lens[sym] : 0;
        const uint32_t same = __match_any_sync(FULL, L);
        const uint32_t rank = __popc(same & ((1u << lane) - 1u));
        if (L) {
            const uint32_t idx = count[L] + rank;  // among the symbols of length L, in symbol order
            const uint32_t code = first[L] + idx;
            sorted[offs[L] + idx] = (uint16_t)sym;
            if (L <= lut_bits) {
                const uint32_t rev = brev_n(code, L);
                for (uint32_t k = rev; k < (1u << lut_bits); k += 1u << L) lut[k] = (uint16_t)((sym << 4) | L);
            }
        }
        __syncwarp();
        if (L && lane == __ffs(same) - 1) count[L] += (uint16_t)__popc(same);
        __syncwarp();
    }
    return true;
}

// LSB-first bit reader: a read is two aligned word loads and a funnel shift (the words may extend a few bytes beyond the
// chunk: inside the arena, never part of a value)
struct InfBits {
    const uint32_t* W;  // chunk start rounded down to a word
    uint32_t boff;      // bit offset of the chunk in W
    uint32_t n;         // chunk bytes
    uint32_t pos;       // next bit
    __device__ __forceinline__ void init(const uint8_t* s, uint32_t len) {
        W = (const uint32_t*)((uintptr_t)s & ~(uintptr_t)3);
        boff = (uint32_t)((uintptr_t)s & 3) * 8;
        n = len;
        pos = 0;
    }
    __device__ __forceinline__ uint32_t peek32() const {
        const uint32_t a = pos + boff;
        return __funnelshift_r(W[a >> 5], W[(a >> 5) + 1], a & 31);
    }
    __device__ __forceinline__ uint32_t take(int k) {  // k <= 16
        const uint32_t v = peek32() & ((1u << k) - 1u);
        pos += k;
        return v;
    }
    __device__ __forceinline__ bool overrun() const { return pos > n * 8u; }
};

// one symbol; -1 = invalid code
__device__ __forceinline__ int inf_symbol(InfBits& b, const uint16_t* lut, int lut_bits, const uint16_t* sorted, const uint16_t* first,
                                          const uint16_t* offs, const uint16_t* count) {
    const uint32_t e = lut[b.peek32() & ((1u << lut_bits) - 1u)];
    if (e) {
        b.pos += e & 15;
        return (int)(e >> 4);
    }
    // long code: canonical decode, one bit at a time (count[] holds the number of codes per length after the build)
    uint32_t code = 0;
    for (int L = 1; L <= 15; L++) {
        code = (code << 1) | b.take(1);
        const uint32_t c = count[L];
        if (code - first[L] < c && code >= first[L]) return sorted[offs[L] + (code - first[L])];
    }
    return -1;
}

__device__ uint32_t inflate_chunk(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint32_t cap, uint32_t& o_out, InfWarp& w, int lane) {
    static const uint16_t LBASE[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint8_t LEXT[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t DBASE[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t DEXT[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    static const uint8_t CLORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    const uint32_t BAD = ORCB_IO_ERROR;
    InfBits b;
    b.init(s, n);
    if (n > (1u << 28)) return BAD;
    uint32_t o = 0;
    int last = 0;
    while (!last) {
        // ---- block header (lane 0 reads, everyone follows)
        uint32_t type = 0;
        if (lane == 0) {
            last = (int)b.take(1);
            type = b.take(2);
            if (b.overrun()) type = 3;
        }
        last = __shfl_sync(FULL, last, 0);
        type = __shfl_sync(FULL, type, 0);
        if (type == 3) return BAD;
        if (type == 0) {
            // stored block: to the next byte boundary, LEN, ~LEN, bytes
            uint32_t len = 0, from = 0, bad = 0;
            if (lane == 0) {
                b.pos = (b.pos + 7u) & ~7u;
                len = b.take(16);
                const uint32_t nlen = b.take(16);
                if ((len ^ nlen) != 0xffffu) bad = 1;
                from = b.pos >> 3;  // first byte not yet consumed
                if (from > n || len > n - from || len > cap - o) bad = 1;
                if (!bad) b.pos = (from + len) * 8u;
            }
            if (__shfl_sync(FULL, bad, 0)) return BAD;
            len = __shfl_sync(FULL, len, 0);
            from = __shfl_sync(FULL, from, 0);
            warp_copy_fwd(d + o, s + from, len, lane);
            __syncwarp();
            o += len;
            continue;
        }
        int nlen = 288, ndist = 30;
        if (type == 1) {
            // fixed code
            for (int i = lane; i < 288; i += 32) w.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
            w.lens[288 + lane] = lane < 30 ? 5 : 0;
        } else {
            // dynamic code: code-length code, then the literal/length and distance code lengths
            uint32_t bad = 0;
            if (lane == 0) {
                nlen = (int)b.take(5) + 257;
                ndist = (int)b.take(5) + 1;
                const int ncode = (int)b.take(4) + 4;
                if (nlen > 286 || ndist > 30) bad = 1;
                for (int i = 0; i < 19; i++) w.stage[i] = 0;
                for (int i = 0; i < ncode && !bad; i++) w.stage[CLORDER[i]] = (uint8_t)b.take(3);
            }
            if (__shfl_sync(FULL, bad, 0)) return BAD;
            nlen = __shfl_sync(FULL, nlen, 0);
            ndist = __shfl_sync(FULL, ndist, 0);
            __syncwarp();
            // the code-length code borrows the distance tables (built again below)
            if (!inf_build(w.stage, 19, w.dlut, 7, w.dsorted, w.dfirst, w.doffs, w.dcount, lane)) return BAD;
            __syncwarp();
            if (lane == 0) {
                int i = 0, prev = 0;
                uint8_t* out = w.stage + 32;
                while (i < nlen + ndist && !bad) {
                    if (b.overrun()) { bad = 1; break; }
                    const int sym = inf_symbol(b, w.dlut, 7, w.dsorted, w.dfirst, w.doffs, w.dcount);
                    if (sym < 0) { bad = 1; break; }
                    if (sym < 16) {
                        out[i++] = (uint8_t)sym;
                        prev = sym;
                    } else {
                        int rep, val = 0;
                        if (sym == 16) {
                            if (i == 0) { bad = 1; break; }
                            val = prev;
                            rep = 3 + (int)b.take(2);
                        } else if (sym == 17) {
                            rep = 3 + (int)b.take(3);
                            prev = 0;
                        } else {
                            rep = 11 + (int)b.take(7);
                            prev = 0;
                        }
                        if (i + rep > nlen + ndist) { bad = 1; break; }
                        while (rep--) out[i++] = (uint8_t)val;
                    }
                }
                if (!bad && out[256] == 0) bad = 1;  // no end-of-block code
                if (b.overrun()) bad = 1;
            }
            if (__shfl_sync(FULL, bad, 0)) return BAD;
            __syncwarp();
            for (int i = lane; i < 288; i += 32) w.lens[i] = i < nlen ? w.stage[32 + i] : 0;
            w.lens[288 + lane] = lane < ndist ? w.stage[32 + nlen + lane] : 0;
        }
        __syncwarp();
        if (!inf_build(w.lens, 288, w.llut, INF_LBITS, w.lsorted, w.lfirst, w.loffs, w.lcount, lane)) return BAD;
        if (!inf_build(w.lens + 288, 32, w.dlut, INF_DBITS, w.dsorted, w.dfirst, w.doffs, w.dcount, lane)) return BAD;
        __syncwarp();
        // ---- symbols: lane 0 queues up to 32 tokens, the warp writes them
        bool end = false;
        while (!end) {
            uint32_t nt = 0, bad = 0;
            if (lane == 0) {
                while (nt < 32) {
                    const int sym = inf_symbol(b, w.llut, INF_LBITS, w.lsorted, w.lfirst, w.loffs, w.lcount);
                    if (sym < 0) { bad = 1; break; }
                    if (sym < 256) {
                        w.tok[nt++] = (uint32_t)sym | (1u << 31);
                    } else if (sym == 256) {
                        end = true;
                        break;
                    } else {
                        if (sym > 285) { bad = 1; break; }
                        const uint32_t len = LBASE[sym - 257] + b.take(LEXT[sym - 257]);
                        const int ds = inf_symbol(b, w.dlut, INF_DBITS, w.dsorted, w.dfirst, w.doffs, w.dcount);
                        if (ds < 0 || ds > 29) { bad = 1; break; }
                        const uint32_t dist = DBASE[ds] + b.take(DEXT[ds]);
                        w.tok[nt++] = len | (dist << 9);
                    }
                }
                if (b.overrun()) bad = 1;
            }
            if (__shfl_sync(FULL, bad, 0)) return BAD;
            nt = __shfl_sync(FULL, nt, 0);
            end = __shfl_sync(FULL, (int)end, 0);
            __syncwarp();
            const uint32_t t = (uint32_t)lane < nt ? w.tok[lane] : 0u;
            const bool lit = (t >> 31) != 0;
            const uint32_t len = (uint32_t)lane < nt ? (lit ? 1u : (t & 511u)) : 0u;
            const uint32_t incl = warp_incl_scan(len, lane);
            const uint32_t total = __shfl_sync(FULL, incl, 31);
            if (total > cap - o) return BAD;
            const uint32_t op = o + incl - len;
            if (lit) d[op] = (uint8_t)t;
            uint32_t mm = __ballot_sync(FULL, len != 0 && !lit);
            __syncwarp();
            while (mm) {
                const int l = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t e_len = __shfl_sync(FULL, len, l), e_dist = __shfl_sync(FULL, t >> 9, l), e_op = __shfl_sync(FULL, op, l);
                if (e_dist == 0 || e_dist > e_op) return BAD;
                warp_copy_match(d, e_op, e_dist, e_len, lane);
                __syncwarp();
            }
            o += total;
        }
    }
    o_out = o;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Zstandard frames and LZO1X streams (src/compression.rs:151-159, 174-183).  Both are decoded the way the inflate
// path is: lane 0 runs the sequential part (zstd_dec.h: headers, FSE / Huffman tables, the sequence bitstream, the LZO
// instruction bytes) and queues up to 32 (literals, match) tokens, the warp then places them.  Huffman-coded
// literals of a Zstandard block are decoded first, one lane per stream (a block has 1 or 4), into the END of the
// chunk's output range: the literals still waiting there always lie at or above the write position (what separates
// the two is the match bytes the block has yet to produce), so executing the sequences front to back never
// overwrites a literal before it has been copied down.
// ------------------------------------------------------------------------------------------------
struct ZstdWarp {
    zstd::Tables T;
    uint32_t tok[32 * 4];  // literal bytes | literal source offset | match bytes | match distance
};
static_assert(sizeof(ZstdWarp) <= sizeof(LzWarp), "the Zstandard tables reuse the LZ decoder's shared memory");

// literals: the source is the input, or the tail of the output at or above dst: read a round, then write it
__device__ __forceinline__ void warp_copy_lit(uint8_t* dst, const uint8_t* src, uint32_t n, int lane) {
    for (uint32_t i0 = 0; i0 < n; i0 += 128) {
        uint8_t v[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t i = i0 + u * 32 + lane;
            v[u] = i < n ? src[i] : (uint8_t)0;
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t i = i0 + u * 32 + lane;
            if (i < n) dst[i] = v[u];
        }
        __syncwarp();
    }
}

// Places `cnt` queued tokens at d + o.  `reserve`: bytes at the end of the output range that must stay untouched
// (Zstandard literals not yet consumed after this batch).  `win_base`: matches may not reach below this offset.
// rle >= 0: literals are this byte.  Returns non-zero when the output does not fit or a match is out of range.
// Short literals, and short matches whose source lies wholly before the batch, do not depend on anything the batch
// writes: every lane copies its own.  What is left (long copies, matches that read the batch's own output) follows
// in token order, copied by the whole warp.
__device__ __forceinline__ uint32_t lz_exec_tokens(uint8_t* d, uint32_t& o, uint32_t cap, uint32_t reserve, uint32_t win_base,
                                                   const uint32_t* tok, uint32_t cnt, const uint8_t* lit_base, int rle, int lane) {
    constexpr uint32_t OWN_MAX = 48;  // longest copy a lane does alone
    const bool on = (uint32_t)lane < cnt;
    const uint32_t ll = on ? tok[4 * lane] : 0u, lo = on ? tok[4 * lane + 1] : 0u;
    const uint32_t ml = on ? tok[4 * lane + 2] : 0u, dist = on ? tok[4 * lane + 3] : 0u;
    const uint32_t len = ll + ml;
    const uint32_t incl = warp_incl_scan(len, lane);
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    if (total > cap - o || reserve > cap - o - total) return 1;
    const uint32_t op = o + incl - len, at = op + ll;
    if (__any_sync(FULL, ml && (dist == 0 || dist > at - win_base))) return 1;
    // literals that wait in the tail of the output range must not be overwritten before they are read: lanes copy on
    // their own only when everything the batch writes lies below the first literal it reads
    const uint64_t lit0 = (uint64_t)(uintptr_t)(lit_base + __shfl_sync(FULL, lo, 0));
    const uint64_t d0 = (uint64_t)(uintptr_t)d;
    const bool lit_apart = lit0 < d0 || lit0 >= d0 + cap || lit0 >= d0 + o + total;
    const bool lit_own = ll && (rle >= 0 || (lit_apart && ll <= OWN_MAX));
    if (lit_own) {
        if (rle >= 0) {
            for (uint32_t i = 0; i < ll; i++) d[op + i] = (uint8_t)rle;
        } else {
            const uint8_t* sp = lit_base + lo;
            for (uint32_t i = 0; i < ll; i++) d[op + i] = sp[i];
        }
    }
    uint32_t lm = __ballot_sync(FULL, ll && !lit_own);
    const bool m_own = ml && ml <= OWN_MAX && at - dist + ml <= o && !lm;
    // (with cooperative literals pending, matches wait: the literal phase below must stay in token order with them)
    if (m_own) {
        const uint8_t* sp = d + (at - dist);
        for (uint32_t i = 0; i < ml; i++) d[at + i] = sp[i];
    }
    uint32_t mm = __ballot_sync(FULL, ml && !m_own);
    __syncwarp();
    uint32_t rest = lm | mm;
    while (rest) {
        const int l = __ffs(rest) - 1;
        rest &= rest - 1;
        if ((lm >> l) & 1u) {
            const uint32_t e_ll = __shfl_sync(FULL, ll, l), e_lo = __shfl_sync(FULL, lo, l), e_op = __shfl_sync(FULL, op, l);
            warp_copy_lit(d + e_op, lit_base + e_lo, e_ll, lane);
        }
        if ((mm >> l) & 1u) {
            const uint32_t e_ml = __shfl_sync(FULL, ml, l), e_dist = __shfl_sync(FULL, dist, l), e_at = __shfl_sync(FULL, at, l);
            warp_copy_match(d, e_at, e_dist, e_ml, lane);
            __syncwarp();
        }
    }
    o += total;
    return 0;
}

// one Compressed_Block of `size` bytes at b
__device__ uint32_t zstd_block(const uint8_t* __restrict__ b, uint32_t size, uint8_t* d, uint32_t cap, uint32_t& o, uint32_t win_base,
                               zstd::FrameState& fs, ZstdWarp& w, int lane) {
    using namespace zstd;
    LitHeader lh;
    uint32_t bad = 0;
    if (lane == 0) bad = !lit_header(b, size, lh);
    if (__shfl_sync(FULL, bad, 0)) return 1;
    lh.type = __shfl_sync(FULL, lh.type, 0);
    lh.regen = __shfl_sync(FULL, lh.regen, 0);
    lh.comp = __shfl_sync(FULL, lh.comp, 0);
    lh.streams = __shfl_sync(FULL, lh.streams, 0);
    lh.hdr = __shfl_sync(FULL, lh.hdr, 0);
    if (lh.regen > cap - o) return 1;
    const uint8_t* lp = b + lh.hdr;
    const uint8_t* lit_base = lp;
    int rle = -1;
    if (lh.type == 1) {
        rle = lp[0];
    } else if (lh.type >= 2) {
        uint32_t used = 0;
        if (lane == 0) {
            if (lh.type == 2) bad = !huf_read(lp, lh.comp, w.T, fs.huf_log, used);
            else bad = fs.huf_log < 0;
        }
        if (__shfl_sync(FULL, bad, 0)) return 1;
        used = __shfl_sync(FULL, used, 0);
        fs.huf_log = __shfl_sync(FULL, fs.huf_log, 0);
        __syncwarp();
        const uint8_t* sp = lp + used;
        const uint32_t sn = lh.comp - used;
        uint8_t* area = d + (cap - lh.regen);
        bool ok = true;
        if (lh.streams == 1) {
            if (lane == 0) ok = huf_stream(sp, sn, w.T.huf, fs.huf_log, area, lh.regen);
        } else {
            if (sn < 10 || lh.regen < 6) return 1;  // libzstd: jump table + one byte per stream; at least 6 literals
            const uint32_t l1 = sp[0] | (sp[1] << 8), l2 = sp[2] | (sp[3] << 8), l3 = sp[4] | (sp[5] << 8);
            if (6u + l1 + l2 + l3 > sn) return 1;
            const uint32_t q = (lh.regen + 3) / 4;
            if (3ull * q > lh.regen) return 1;
            if (lane < 4) {
                const uint32_t at = 6 + (lane > 0 ? l1 : 0u) + (lane > 1 ? l2 : 0u) + (lane > 2 ? l3 : 0u);
                const uint32_t sl = lane == 0 ? l1 : lane == 1 ? l2 : lane == 2 ? l3 : sn - 6 - l1 - l2 - l3;
                ok = huf_stream(sp + at, sl, w.T.huf, fs.huf_log, area + lane * q, lane < 3 ? q : lh.regen - 3 * q);
            }
        }
        if (!__all_sync(FULL, ok)) return 1;
        __syncwarp();
        lit_base = area;
    }
    // ---- sequences
    const uint8_t* sq = b + lh.hdr + lh.comp;
    const uint32_t sqn = size - lh.hdr - lh.comp;
    uint32_t nseq = 0, used = 0;
    SeqReader r;
    if (lane == 0) {
        bad = !seq_header(sq, sqn, w.T, fs, nseq, used);
        if (!bad && nseq) bad = !r.init(sq + used, sqn - used, fs);
        if (!bad && !nseq && used != sqn) bad = 1;
    }
    if (__shfl_sync(FULL, bad, 0)) return 1;
    nseq = __shfl_sync(FULL, nseq, 0);
    fill_codes(w.T, (uint32_t)lane, 32);  // over the Huffman weights, which are not needed any more
    __syncwarp();
    uint32_t lit_at = 0;  // lane 0 keeps it
    for (uint32_t s0 = 0; s0 < nseq; s0 += 32) {
        const uint32_t cnt = min(32u, nseq - s0);
        if (lane == 0) {
            for (uint32_t i = 0; i < cnt && !bad; i++) {
                uint32_t ll, ml, off;
                if (!r.next(w.T, fs, s0 + i + 1 == nseq, ll, ml, off) || ll > lh.regen - lit_at) {
                    bad = 1;
                    break;
                }
                w.tok[4 * i] = ll;
                w.tok[4 * i + 1] = lit_at;
                w.tok[4 * i + 2] = ml;
                w.tok[4 * i + 3] = off;
                lit_at += ll;
            }
            if (!bad && s0 + cnt == nseq && r.b.bits != 0) bad = 1;
        }
        if (__shfl_sync(FULL, bad, 0)) return 1;
        const uint32_t reserve = lh.regen - __shfl_sync(FULL, lit_at, 0);
        __syncwarp();
        if (lz_exec_tokens(d, o, cap, reserve, win_base, w.tok, cnt, lit_base, rle, lane)) return 1;
        __syncwarp();
    }
    // literals after the last sequence
    lit_at = __shfl_sync(FULL, lit_at, 0);
    const uint32_t rest = lh.regen - lit_at;
    if (rest > cap - o) return 1;
    if (rest) {
        if (rle >= 0) {
            for (uint32_t i = lane; i < rest; i += 32) d[o + i] = (uint8_t)rle;
            __syncwarp();
        } else {
            warp_copy_lit(d + o, lit_base + lit_at, rest, lane);
        }
        o += rest;
    }
    return 0;
}

__device__ uint32_t zstd_chunk(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint32_t cap, uint32_t& o_out, ZstdWarp& w, int lane) {
    using namespace zstd;
    const uint32_t BAD = ORCB_IO_ERROR;
    uint32_t p = 0, o = 0;
    FrameState fs;
    while (p < n) {
        FrameHeader fh;
        uint32_t bad = 0;
        if (lane == 0) bad = !frame_header(s + p, n - p, fh);
        if (__shfl_sync(FULL, bad, 0)) return BAD;
        p += __shfl_sync(FULL, fh.hdr, 0);
        if (__shfl_sync(FULL, fh.skippable, 0)) continue;
        const uint32_t csum = __shfl_sync(FULL, fh.checksum, 0);
        const uint32_t c_lo = __shfl_sync(FULL, (uint32_t)fh.content, 0), c_hi = __shfl_sync(FULL, (uint32_t)(fh.content >> 32), 0);
        fs.reset();
        const uint32_t base = o;
        for (;;) {
            if (n - p < 3) return BAD;
            const uint32_t bh = s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16);
            p += 3;
            const uint32_t last = bh & 1, type = (bh >> 1) & 3, size = bh >> 3;
            if (type == 3) return BAD;
            if (type == 0) {
                if (size > n - p || size > cap - o) return BAD;
                warp_copy_fwd(d + o, s + p, size, lane);
                __syncwarp();
                o += size;
                p += size;
            } else if (type == 1) {
                if (p >= n || size > cap - o) return BAD;
                const uint8_t v = s[p];
                for (uint32_t i = lane; i < size; i += 32) d[o + i] = v;
                __syncwarp();
                o += size;
                p += 1;
            } else {
                if (size > n - p || size > (128u << 10)) return BAD;
                if (zstd_block(s + p, size, d, cap, o, base, fs, w, lane)) return BAD;
                p += size;
            }
            if (last) break;
        }
        if (csum) {
            if (n - p < 4) return BAD;
            p += 4;  // content checksum: not verified
        }
        if ((c_lo & c_hi) != 0xffffffffu && (c_hi != 0 || o - base != c_lo)) return BAD;
    }
    o_out = o;
    return 0;
}

__device__ uint32_t lzo_chunk(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint32_t cap, uint32_t& o_out, uint32_t* tok, int lane) {
    const uint32_t BAD = ORCB_BUILD_LZO_DECODER;
    uint32_t p = 0, state = 0, o = 0;
    uint32_t carry_pos = 0, carry_len = 0;  // literals waiting for the match they precede (lane 0)
    bool first = true, done = false;
    while (!done) {
        uint32_t cnt = 0, bad = 0;
        if (lane == 0) {
            while (cnt < 32) {
                lzo::Token t;
                if (!lzo::next(s, n, p, state, first, t)) {
                    bad = 1;
                    break;
                }
                first = false;
                if (t.end || t.m_len == 0) {
                    // a literal instruction: flush literals already waiting (crafted input only), then wait for the match
                    if (carry_len && (t.end || t.lit_len)) {
                        tok[4 * cnt] = carry_len;
                        tok[4 * cnt + 1] = carry_pos;
                        tok[4 * cnt + 2] = 0;
                        tok[4 * cnt + 3] = 0;
                        cnt++;
                        carry_len = 0;
                    }
                    if (t.end) {
                        done = true;
                        if (p != n) bad = 1;  // bytes after the end marker
                        break;
                    }
                    carry_pos = t.lit_pos;
                    carry_len = t.lit_len;
                    continue;
                }
                tok[4 * cnt] = carry_len;
                tok[4 * cnt + 1] = carry_pos;
                tok[4 * cnt + 2] = t.m_len;
                tok[4 * cnt + 3] = t.m_dist;
                cnt++;
                carry_pos = t.lit_pos;
                carry_len = t.lit_len;
            }
        }
        if (__shfl_sync(FULL, bad, 0)) return BAD;
        cnt = __shfl_sync(FULL, cnt, 0);
        done = __shfl_sync(FULL, (int)done, 0);
        __syncwarp();
        if (lz_exec_tokens(d, o, cap, 0, 0, tok, cnt, s, -1, lane)) return BAD;
        __syncwarp();
    }
    o_out = o;
    return 0;
}

// Persistent warps over the chunk list (most expensive chunks first, see plan.cc); one chunk per warp at a time.
// Three kernels, so that each keeps its register allocation and its code stays small (the Snappy and the LZ4 tile
// decoder in one kernel are 175 KB of SASS: the LZ4 path then waits for instruction fetches a third of the time):
// k_decompress<2> takes Snappy chunks, k_decompress<4> LZ4 and stored chunks, k_decompress_bits the codecs whose chunk is
// one serial bit / instruction chain (Zlib, Zstandard, LZO).  Every kernel copies stored chunks it is handed.
template <int CODEC>
__global__ void __launch_bounds__(128, ORCB_LZ_CTAS) k_decompress(const ChunkDesc* __restrict__ chunks, uint32_t nchunks, uint32_t* err,
                                                       uint32_t* out_lens, uint32_t* counter, uint32_t* retry) {
    __shared__ LzWarp warp_sm[4];
    __shared__ uint32_t lut[256];
    if (CODEC == 2) {
        lut[threadIdx.x] = snappy_tag_entry(threadIdx.x);
        lut[threadIdx.x + 128] = snappy_tag_entry(threadIdx.x + 128);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    LzWarp& sm = warp_sm[threadIdx.x >> 5];
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(counter, 1u);
        ci = __shfl_sync(FULL, ci, 0);
        if (ci >= nchunks) return;
        const ChunkDesc& c = chunks[ci];
        const uint8_t* s = (const uint8_t*)c.src;
        uint8_t* d = (uint8_t*)c.dst;
        const uint32_t n = c.src_len;
        uint32_t o = 0, fail = 0;
        if (c.codec == 0) {
            if (n > c.dst_cap) fail = ORCB_UNEXPECTED;
            else warp_copy_fwd(d, s, n, lane);
            o = n;
        } else if (CODEC == 2 && c.codec == 2) {
            uint32_t p;
            uint64_t ulen;
            fail = snappy_preamble(s, n, c.dst_cap, p, ulen);
            if (!fail) fail = lz_chunk<2>(s, n, d, ulen, p, o, sm, lut, lane);
        } else if (CODEC == 4 && c.codec == 4) {
            fail = lz_chunk<4>(s, n, d, c.dst_cap, 0, o, sm, lut, lane);
            // the size of a stream's last LZ4 chunk is only known here: what the layout reserved beyond it reads as zeros
            if (c.expect_len < 0 && !fail)
                for (uint32_t a = o + lane; a < c.dst_cap; a += 32) d[a] = 0;
        } else {
            fail = ORCB_UNEXPECTED;  // not this kernel's (launch_decompress splits the list)
        }
        chunk_done(c, fail, o, err, out_lens, retry, lane);
        __syncwarp();
    }
}

// one kernel per serial-chain codec: each has only its own tables in shared memory, so the small ones (inflate: 4.5 KiB
// per warp, LZO: 512 bytes) keep more chains in flight per SM than the Zstandard tables would allow
template <int CODEC> struct BitsSmem;
template <> struct BitsSmem<1> { InfWarp w; };
template <> struct BitsSmem<3> { uint32_t tok[32 * 4]; };
template <> struct BitsSmem<5> { ZstdWarp w; };
#ifndef ORCB_BITS_CTAS
#define ORCB_BITS_CTAS 5
#endif
#ifndef ORCB_INF_CTAS
#define ORCB_INF_CTAS 10
#endif

template <int CODEC>
__global__ void __launch_bounds__(128, CODEC == 5 ? ORCB_BITS_CTAS : ORCB_INF_CTAS)
k_decompress_bits(const ChunkDesc* __restrict__ chunks, uint32_t nchunks, uint32_t* err, uint32_t* out_lens, uint32_t* counter,
                  uint32_t* retry, int copy_stored) {
    __shared__ BitsSmem<CODEC> warp_sm[4];
    const int lane = threadIdx.x & 31;
    BitsSmem<CODEC>& sm = warp_sm[threadIdx.x >> 5];
    for (;;) {
        uint32_t ci = 0;
        if (lane == 0) ci = atomicAdd(counter, 1u);
        ci = __shfl_sync(FULL, ci, 0);
        if (ci >= nchunks) return;
        const ChunkDesc& c = chunks[ci];
        if (c.codec != CODEC && !(c.codec == 0 && copy_stored)) continue;  // another launch takes it
        const uint8_t* s = (const uint8_t*)c.src;
        uint8_t* d = (uint8_t*)c.dst;
        const uint32_t n = c.src_len;
        uint32_t o = 0, fail = 0;
        if (c.codec == 0) {
            if (n > c.dst_cap) fail = ORCB_UNEXPECTED;
            else warp_copy_fwd(d, s, n, lane);
            o = n;
        } else if constexpr (CODEC == 1) {
            fail = inflate_chunk(s, n, d, c.dst_cap, o, sm.w, lane);
        } else if constexpr (CODEC == 5) {
            fail = zstd_chunk(s, n, d, c.dst_cap, o, sm.w, lane);
        } else {
            fail = lzo_chunk(s, n, d, c.dst_cap, o, sm.tok, lane);
        }
        // the size of a stream's last chunk is only known here: what the layout reserved beyond it reads as zeros
        if (c.codec != 0 && c.expect_len < 0 && !fail)
            for (uint32_t a = o + lane; a < c.dst_cap; a += 32) d[a] = 0;
        chunk_done(c, fail, o, err, out_lens, retry, lane);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// host-side launch wrappers
// ------------------------------------------------------------------------------------------------
template <class K>
static uint32_t resident_ctas(K kernel, int& cache) {
    if (!cache) {
        int dev = 0, sms = 148, per_sm = 4;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
        cache = sms * (per_sm > 0 ? per_sm : 1);
    }
    return (uint32_t)cache;
}

template <int CODEC>
static int launch_bits(const ChunkDesc* c, uint32_t n, uint32_t* err, uint32_t* out_lens, uint32_t* counter, uint32_t* retry, int copy_stored,
                       cudaStream_t st) {
    static int ctas = 0;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(resident_ctas(k_decompress_bits<CODEC>, ctas), ((uint64_t)n + 3) / 4);
    k_decompress_bits<CODEC><<<grid, 128, 0, st>>>(c, n, err, out_lens, counter, retry, copy_stored);
    LAUNCH_CHECK();
    return 0;
}

// The first n_bits chunks of the list go to the k_decompress_bits kernels (one launch per codec in `bits_codecs`, a
// bit mask of 1 << codec; each takes its own chunks of the range), the next n_snappy to k_decompress<2>, the rest
// (LZ4, stored) to k_decompress<4>.  counter: eight zeroed words.
int launch_decompress(const ChunkDesc* c, uint32_t n, uint32_t n_bits, uint32_t bits_codecs, uint32_t n_snappy, uint32_t* err, uint32_t* out_lens,
                      uint32_t* counter, uint32_t* retry, cudaStream_t st) {
    if (!n) return 0;
    static int ctas_snappy = 0, ctas_lz4 = 0;
    if (n_bits > n) n_bits = n;
    if (n_snappy > n - n_bits) n_snappy = n - n_bits;
    if (n_bits) {
        int copy_stored = 1, rc = 0;  // stored chunks of the range (stream-level entry point only) go with the first launch
        if (bits_codecs & (1u << 1)) { rc = launch_bits<1>(c, n_bits, err, out_lens, counter + 1, retry, copy_stored, st); copy_stored = 0; }
        if (rc) return rc;
        if (bits_codecs & (1u << 5)) { rc = launch_bits<5>(c, n_bits, err, out_lens, counter + 3, retry, copy_stored, st); copy_stored = 0; }
        if (rc) return rc;
        if ((bits_codecs & (1u << 3)) || copy_stored) rc = launch_bits<3>(c, n_bits, err, out_lens, counter + 4, retry, copy_stored, st);
        if (rc) return rc;
    }
    if (n_snappy) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>(resident_ctas(k_decompress<2>, ctas_snappy), ((uint64_t)n_snappy + 3) / 4);
        k_decompress<2><<<grid, 128, 0, st>>>(c + n_bits, n_snappy, err, out_lens, counter, retry);
        LAUNCH_CHECK();
    }
    const uint32_t rest = n - n_bits - n_snappy;
    if (rest) {
        const uint32_t grid = (uint32_t)std::min<uint64_t>(resident_ctas(k_decompress<4>, ctas_lz4), ((uint64_t)rest + 3) / 4);
        k_decompress<4><<<grid, 128, 0, st>>>(c + n_bits + n_snappy, rest, err, out_lens, counter + 2, retry);
        LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace orcb
