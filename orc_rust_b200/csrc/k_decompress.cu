// Compression chunks: original copies, Snappy and LZ4 block decoding.
#include "kernel_util.cuh"

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

// ------------------------------------------------------------------------------------------------
// Snappy, 32 input positions at a time.  An element (a literal or a back-reference) is 2 or 3 input bytes on
// average, so a serial tag loop spends a whole warp on ~6 bytes of output per step.  Here every lane decodes the
// element that WOULD start at its byte of a 32-byte input window; the elements that really start there are the
// chain 0 -> next[0] -> next[next[0]] .., found by pointer jumping over the lanes (4 rounds, at most 16 elements of
// >= 2 bytes fit).  A warp scan of their output lengths places them, and the bytes are then produced 32 at a time,
// one output byte per lane, whichever element it belongs to:
//   * literal bytes come from the input, bytes of a back-reference whose source lies before this window's output
//     come from the output already written (a source shorter than the copy repeats, i % dist);
//   * a back-reference into this window's own output waits for the pass above, then is copied by the whole warp,
//     in order.
// A chunk is one serial chain of windows, so its time is windows x latency of one window, and a kernel of a few
// thousand chunks is as slow as its slowest chunk.  The round trips through global memory are therefore kept off the
// chain: every output byte also goes into a ring of the last SW_HIST bytes in shared memory, where nearly all
// back-references (and all that point into the current window) find their source.
// Anything out of the ordinary (a header or literal that crosses the end of the input, a distance of zero or
// beyond the output so far, output past the announced length) leaves the window untouched and returns false: the
// serial loop takes over from the same position and reports the error exactly as it always did.
// ------------------------------------------------------------------------------------------------
// tag byte -> header bytes [0:3] | literal [3] | length is in the following bytes [4] | bytes of offset / length
// that follow [5:8] | high offset bits of a 1-byte-offset copy [8:11] | length [16:]
__device__ __forceinline__ uint32_t snappy_tag_entry(uint32_t tag) {
    const uint32_t t = tag & 3u, l = tag >> 2;
    if (t == 0) return l < 60 ? (1u | 8u | ((l + 1u) << 16)) : ((1u + (l - 59u)) | 8u | 16u | ((l - 59u) << 5));
    if (t == 1) return 2u | (1u << 5) | ((tag >> 5) << 8) | ((4u + (l & 7u)) << 16);
    if (t == 2) return 3u | (2u << 5) | ((l + 1u) << 16);
    return 5u | (4u << 5) | ((l + 1u) << 16);
}

struct SnappyWin {
    uint32_t out_off[32];  // where the element's output starts, relative to the window's first output byte
    uint32_t src[32];      // literal: input offset of its first byte; back-reference: distance
    uint32_t info[32];     // output length | literal << 31 | into-this-window << 30
};
constexpr uint32_t SW_LITERAL = 1u << 31, SW_DEPENDENT = 1u << 30, SW_LEN = SW_DEPENDENT - 1;
constexpr uint32_t SW_LONG_LITERAL = 128;  // a literal this long ends its window and is copied word-wise
constexpr uint32_t SW_HIST = 4096;         // bytes of recent output each warp keeps in shared memory (power of two)

__device__ __forceinline__ bool snappy_window(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint64_t ulen, uint32_t& p,
                                              uint32_t& o, SnappyWin& w, uint8_t* hist, const uint32_t* lut, int lane) {
    // bytes q .. q+4 of the input for q = p + lane, out of ten aligned words
    const uintptr_t a0 = (uintptr_t)(s + p);
    const uint32_t* wp = (const uint32_t*)(a0 & ~(uintptr_t)3);
    const uint32_t word = lane < 10 ? wp[lane] : 0u;
    const uint32_t b = (uint32_t)(a0 & 3) + (uint32_t)lane;
    const uint32_t w0 = __shfl_sync(FULL, word, b >> 2), w1 = __shfl_sync(FULL, word, (b >> 2) + 1);
    const uint32_t lo = __funnelshift_r(w0, w1, (b & 3) * 8);  // bytes 0..3
    const uint32_t b4 = (w1 >> ((b & 3) * 8)) & 0xffu;          // byte 4
    // what the tag says comes out of a 256-entry table (snappy_tag_entry): header bytes, literal or not, the length
    // when the tag holds it, and how many of the following bytes are an offset or a length
    const uint32_t e = lut[lo & 0xffu];
    const uint32_t hdr = e & 7u, nbytes = (e >> 5) & 7u;
    const bool lit = (e >> 3) & 1u;
    const uint32_t raw = (lo >> 8) | (b4 << 24);
    const uint32_t field = raw & __funnelshift_rc(0xffffffffu, 0u, 32u - 8u * nbytes);  // the low `nbytes` bytes
    const uint32_t len = (e >> 16) + (((e >> 4) & 1u) ? field + 1u : 0u);             // long literals: length bytes + 1
    const uint32_t q = p + (uint32_t)lane;
    const uint32_t src = lit ? q + hdr : (field | (((e >> 8) & 7u) << 8));
    // input bytes the element takes; a lane past the end of the input starts nothing.  A literal longer than the
    // input is not sane, so the 32-bit sums below cannot wrap for the lanes that matter
    const bool inside = q < n;
    const uint32_t adv = hdr + (lit ? len : 0u);
    const bool sane = inside && len != 0 && (!lit || len <= n) && adv <= n - q;
    const uint32_t nxt = sane ? min((uint32_t)lane + adv, 32u) : 32u;
    // the chain of real element starts
    uint32_t reach = 1u, jump = nxt;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        reach |= __reduce_or_sync(FULL, (((reach >> lane) & 1u) && jump < 32u) ? (1u << jump) : 0u);
        const uint32_t j2 = __shfl_sync(FULL, jump, jump & 31u);
        jump = jump < 32u ? j2 : 32u;
    }
    const bool mine = ((reach >> lane) & 1u) && inside;
    const uint32_t mask = __ballot_sync(FULL, mine);
    // place the output
    const uint32_t olen = mine ? len : 0u;
    const uint32_t incl = warp_incl_scan(mine && sane ? olen : 0u, lane);
    const uint32_t excl = incl - (mine && sane ? olen : 0u);
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    const uint64_t oo = (uint64_t)o + excl;
    bool bad = mine && (!sane || oo + len > ulen || (!lit && (src == 0 || (uint64_t)src > oo)));
    if (__any_sync(FULL, bad)) return false;
    // a back-reference reads [oo - dist, oo - dist + min(len, dist)); it depends on this window when that ends past o
    const bool dep = mine && !lit && excl + min(len, src) > src;
    const int last = 31 - __clz(mask);
    const bool long_lit = lit && len >= SW_LONG_LITERAL;  // only possible for the window's last element
    const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
    if (mine) {
        w.out_off[rank] = excl;
        w.src[rank] = src;
        w.info[rank] = len | (lit ? SW_LITERAL : 0u) | ((dep || long_lit) ? SW_DEPENDENT : 0u);
    }
    __syncwarp();
    const uint32_t depmask = __ballot_sync(FULL, dep);
    const uint32_t long_len = __shfl_sync(FULL, long_lit && mine ? len : 0u, last);
    uint8_t* const dw = d + o;
    // pass 1: one output byte per lane
    const uint32_t body = total - long_len;
    for (uint32_t v0 = 0; v0 < body; v0 += 32) {
        const uint32_t starts = __reduce_or_sync(FULL, (mine && excl >= v0 && excl < v0 + 32u) ? (1u << (excl - v0)) : 0u);
        const uint32_t before = __popc(__ballot_sync(FULL, mine && excl < v0));
        const uint32_t v = v0 + (uint32_t)lane;
        if (v < body) {
            const uint32_t r = before + __popc(starts & (0xffffffffu >> (31 - lane))) - 1u;
            const uint32_t inf = w.info[r];
            if (!(inf & SW_DEPENDENT)) {
                const uint32_t k = v - w.out_off[r], sv = w.src[r];
                uint8_t byte;
                if (inf & SW_LITERAL) {
                    byte = s[sv + k];
                } else {
                    const uint32_t kk = k < sv ? k : k % sv;
                    const uint32_t a = o + v - k - sv + kk;  // output position of the source byte, < o
                    // the ring still holds it unless this window's own output has come round to its slot
                    byte = (o + total - a <= SW_HIST) ? hist[a & (SW_HIST - 1)] : d[a];
                }
                hist[(o + v) & (SW_HIST - 1)] = byte;
                dw[v] = byte;
            }
        }
    }
    __syncwarp();
    // pass 2: what reads this window's own output, in order
    uint32_t dm = depmask;
    while (dm) {
        const int l = __ffs(dm) - 1;
        dm &= dm - 1;
        const uint32_t e_off = __shfl_sync(FULL, excl, l), e_len = __shfl_sync(FULL, len, l), e_dist = __shfl_sync(FULL, src, l);
        // its source starts less than 64 bytes before the window: all of it is in the ring, and none of it is
        // written by this copy (a source shorter than the copy repeats)
        const uint32_t eo = o + e_off;
        for (uint32_t i = lane; i < e_len; i += 32) {
            const uint32_t kk = i < e_dist ? i : i % e_dist;
            const uint8_t byte = hist[(eo - e_dist + kk) & (SW_HIST - 1)];
            hist[(eo + i) & (SW_HIST - 1)] = byte;
            d[eo + i] = byte;
        }
        __syncwarp();
    }
    if (long_len) {
        const uint32_t e_off = __shfl_sync(FULL, excl, last), e_src = __shfl_sync(FULL, src, last);
        warp_copy_fwd(dw + e_off, s + e_src, long_len, lane);
        for (uint32_t i = (long_len > SW_HIST ? long_len - SW_HIST : 0u) + lane; i < long_len; i += 32)
            hist[(o + e_off + i) & (SW_HIST - 1)] = s[e_src + i];
        __syncwarp();
    }
    p += __shfl_sync(FULL, (uint32_t)lane + adv, last);
    o += total;
    return true;
}

__global__ void __launch_bounds__(128) k_decompress(const ChunkDesc* __restrict__ chunks, uint32_t nchunks, uint32_t* err,
                                                    uint32_t* out_lens) {
    __shared__ SnappyWin win_all[4];
    __shared__ uint8_t hist_all[4][SW_HIST];
    __shared__ uint32_t lut[256];
    lut[threadIdx.x] = snappy_tag_entry(threadIdx.x);
    lut[threadIdx.x + 128] = snappy_tag_entry(threadIdx.x + 128);
    __syncthreads();
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
        // windows of 32 input bytes while all is well; the serial loop below finishes (and reports) the rest
        if (!fail) {
            SnappyWin& w = win_all[threadIdx.x >> 5];
            while (p < n && snappy_window(s, n, d, ulen, p, o, w, hist_all[threadIdx.x >> 5], lut, lane)) __syncwarp();
        }
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
int launch_decompress(const ChunkDesc* c, uint32_t n, uint32_t* err, uint32_t* out_lens, cudaStream_t st) {
    if (!n) return 0;
    k_decompress<<<blocks_for_warps(n, 4), 128, 0, st>>>(c, n, err, out_lens);
    LAUNCH_CHECK();
    return 0;
}

}  // namespace orcb
