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
// Snappy, 32 input positions at a time.
//
// An element (a literal or a back-reference) is 2 or 3 input bytes on average, so a serial tag loop spends a whole
// warp on a few bytes of output per step; and a chunk is one serial chain, so a launch of a few thousand chunks is as
// slow as its slowest chunk (decimal varints and dictionary keys: up to 120 000 elements of 4 bytes in 256 KiB).
//   * Every lane decodes the element that WOULD start at its byte of a 32-byte input window (snappy_parse); the
//     elements that really start there are the chain 0 -> next[0] -> next[next[0]] .., found by pointer jumping over
//     the lanes (4 rounds, at most 16 elements of >= 2 bytes fit).  A warp scan of their output lengths places them.
//   * The bytes are then produced 32 at a time, one output byte per lane, whichever element it belongs to
//     (snappy_copy).  Literal bytes come from the input, bytes of a back-reference from a ring of the last SW_HIST
//     output bytes in shared memory (older ones from global memory); a back-reference into its own window's output
//     waits for the rest of the window and is then copied by the whole warp, in order.
//   * Where elements start and where their output goes depends on the input only, never on the output.  Large
//     launches (more chunks than the GPU holds warps) run both steps in one warp per chunk, k_decompress: they are
//     bound by instruction issue and a second warp would only take a slot away.  Small launches (a reader's group of
//     stripes) are bound by the latency of the slowest chunk; there k_decompress_pair gives each chunk a PARSER
//     warp that runs ahead and leaves one descriptor per window in a small shared-memory queue, and a COPIER warp
//     that turns descriptors into bytes, which shortens the chain per window by about a third.
// Anything out of the ordinary (a header or literal that crosses the end of the input, a distance of zero or
// beyond the output so far, output past the announced length) is not parsed as a window: the element-by-element
// loop takes over from that position (in the pair, after the copier has drained the queue) and reports the error
// exactly as it always did.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t SW_LITERAL = 1u << 31, SW_DEPENDENT = 1u << 30, SW_LONG = 1u << 29, SW_LEN = SW_LONG - 1;
constexpr uint32_t SW_LONG_LITERAL = 128;  // a literal this long ends its window and is copied word-wise
constexpr uint32_t SW_HIST = 4096;         // bytes of recent output the copier keeps in shared memory (power of two)
constexpr uint32_t SW_QUEUE = 4;           // windows the parser may be ahead
enum : uint32_t { SW_WINDOW = 1, SW_END = 2, SW_FALLBACK = 3 };

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
    uint32_t info[32];     // output length | SW_LITERAL | SW_DEPENDENT (reads this window's output) | SW_LONG
    uint32_t state, count, total, o, p;  // o: output position of the window; p: input position (FALLBACK)
    uint32_t pad[3];
};
struct SnappyPair {
    SnappyWin win[SW_QUEUE];
    uint32_t head, tail;  // windows published / consumed
    uint8_t hist[SW_HIST];
};

// parser: one window at input position p / output position o into `w`; false = not a regular window
__device__ __forceinline__ bool snappy_parse(const uint8_t* __restrict__ s, uint32_t n, const uint8_t* d, uint64_t ulen, uint32_t& p,
                                             uint32_t& o, SnappyWin& w, const uint32_t* lut, int lane) {
    // bytes q .. q+4 of the input for q = p + lane, out of ten aligned words
    const uintptr_t a0 = (uintptr_t)(s + p);
    const uint32_t* wp = (const uint32_t*)(a0 & ~(uintptr_t)3);
    const uint32_t word = lane < 10 ? wp[lane] : 0u;
    if (lane == 10) asm volatile("prefetch.global.L1 [%0];" ::"l"(s + p + 160));
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
    const uint32_t olen = mine && sane ? len : 0u;
    const uint32_t incl = warp_incl_scan(olen, lane);
    const uint32_t excl = incl - olen;
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    const uint64_t oo = (uint64_t)o + excl;
    const bool bad = mine && (!sane || oo + len > ulen || (!lit && (src == 0 || (uint64_t)src > oo)));
    if (__any_sync(FULL, bad)) return false;
    // a back-reference reads [oo - dist, oo - dist + min(len, dist)); it depends on this window when that ends past o
    const bool dep = mine && !lit && excl + min(len, src) > src;
    const bool long_lit = mine && lit && len >= SW_LONG_LITERAL;  // only possible for the window's last element
    if (mine && !lit && !dep && src > SW_HIST / 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(d + (oo - src)));
    const int last = 31 - __clz(mask);
    const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
    if (mine) {
        w.out_off[rank] = excl;
        w.src[rank] = src;
        w.info[rank] = len | (lit ? SW_LITERAL : 0u) | (dep ? SW_DEPENDENT : 0u) | (long_lit ? SW_LONG : 0u);
    }
    if (lane == 0) {
        w.count = __popc(mask);
        w.total = total;
        w.o = o;
    }
    p += __shfl_sync(FULL, (uint32_t)lane + adv, last);
    o += total;
    return true;
}

// copier: the bytes of one window
__device__ __forceinline__ void snappy_copy(const uint8_t* __restrict__ s, uint8_t* d, const SnappyWin& w, uint8_t* hist, int lane) {
    const uint32_t count = w.count, total = w.total, o = w.o;
    const bool mine = (uint32_t)lane < count;  // lane r holds element r
    const uint32_t excl = mine ? w.out_off[lane] : 0u, info = mine ? w.info[lane] : 0u, src = mine ? w.src[lane] : 0u;
    const uint32_t len = info & SW_LEN;
    const uint32_t depmask = __ballot_sync(FULL, (info & SW_DEPENDENT) != 0);
    const uint32_t long_len = __shfl_sync(FULL, (info & SW_LONG) ? len : 0u, (count - 1u) & 31u);
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
            if (!(inf & (SW_DEPENDENT | SW_LONG))) {
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
        const uint32_t e_off = __shfl_sync(FULL, excl, (count - 1u) & 31u), e_src = __shfl_sync(FULL, src, (count - 1u) & 31u);
        warp_copy_fwd(dw + e_off, s + e_src, long_len, lane);
        for (uint32_t i = (long_len > SW_HIST ? long_len - SW_HIST : 0u) + lane; i < long_len; i += 32)
            hist[(o + e_off + i) & (SW_HIST - 1)] = s[e_src + i];
        __syncwarp();
    }
}

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p) { return *(const volatile uint32_t*)p; }
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v) { *(volatile uint32_t*)p = v; }

// what the windows leave over: element by element (damaged input ends up here and is reported)
__device__ __forceinline__ uint32_t snappy_serial(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint64_t ulen, uint32_t p,
                                                  uint32_t& o, int lane) {
    while (p < n) {
        const uint32_t tag = s[p++];
        const uint32_t t = tag & 3;
        if (t == 0) {
            uint32_t l = tag >> 2;
            if (l >= 60) {
                const uint32_t extra = l - 59;
                if (p + extra > n) return ORCB_BUILD_SNAPPY_DECODER;
                l = 0;
                for (uint32_t k = 0; k < extra; k++) l |= (uint32_t)s[p + k] << (8 * k);
                p += extra;
            }
            l += 1;
            if (p + l > n || (uint64_t)o + l > ulen) return ORCB_BUILD_SNAPPY_DECODER;
            warp_copy_fwd(d + o, s + p, l, lane);
            __syncwarp();
            p += l;
            o += l;
        } else {
            uint32_t l, dist;
            if (t == 1) {
                if (p + 1 > n) return ORCB_BUILD_SNAPPY_DECODER;
                l = 4 + ((tag >> 2) & 7);
                dist = ((tag >> 5) << 8) | s[p];
                p += 1;
            } else if (t == 2) {
                if (p + 2 > n) return ORCB_BUILD_SNAPPY_DECODER;
                l = 1 + (tag >> 2);
                dist = s[p] | ((uint32_t)s[p + 1] << 8);
                p += 2;
            } else {
                if (p + 4 > n) return ORCB_BUILD_SNAPPY_DECODER;
                l = 1 + (tag >> 2);
                dist = s[p] | ((uint32_t)s[p + 1] << 8) | ((uint32_t)s[p + 2] << 16) | ((uint32_t)s[p + 3] << 24);
                p += 4;
            }
            if (dist == 0 || dist > o || (uint64_t)o + l > ulen) return ORCB_BUILD_SNAPPY_DECODER;
            warp_copy_match(d, o, dist, l, lane);
            o += l;
        }
    }
    return o != ulen ? (uint32_t)ORCB_BUILD_SNAPPY_DECODER : 0u;
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

// LZ4 block (lz4_flex::block::decompress(src, max), compression.rs:185-195)
__device__ __forceinline__ uint32_t lz4_block(const uint8_t* __restrict__ s, uint32_t n, uint8_t* d, uint32_t dst_cap, uint32_t& o, int lane) {
    uint32_t p = 0;
    while (p < n) {
        const uint32_t tok = s[p++];
        uint32_t ll = tok >> 4;
        if (ll == 15) {
            for (;;) {
                if (p >= n) return ORCB_BUILD_LZ4_DECODER;
                const uint32_t b = s[p++];
                ll += b;
                if (b != 255) break;
            }
        }
        if (p + ll > n || o + ll > dst_cap) return ORCB_BUILD_LZ4_DECODER;
        warp_copy_fwd(d + o, s + p, ll, lane);
        __syncwarp();
        p += ll;
        o += ll;
        if (p >= n) break;
        if (p + 2 > n) return ORCB_BUILD_LZ4_DECODER;
        const uint32_t dist = s[p] | ((uint32_t)s[p + 1] << 8);
        p += 2;
        uint32_t ml = tok & 15;
        if (ml == 15) {
            for (;;) {
                if (p >= n) return ORCB_BUILD_LZ4_DECODER;
                const uint32_t b = s[p++];
                ml += b;
                if (b != 255) break;
            }
        }
        ml += 4;
        if (dist == 0 || dist > o || o + ml > dst_cap) return ORCB_BUILD_LZ4_DECODER;
        warp_copy_match(d, o, dist, ml, lane);
        o += ml;
    }
    return 0;
}

__device__ __forceinline__ void chunk_done(const ChunkDesc& c, uint32_t ci, uint32_t fail, uint32_t o, uint32_t* err, uint32_t* out_lens,
                                           int lane) {
    if (!fail && c.expect_len >= 0 && o != (uint32_t)c.expect_len) fail = ORCB_UNEXPECTED;
    if (lane == 0) {
        if (fail) set_err(err, c.colstripe, fail);
        if (out_lens) out_lens[ci] = o;
    }
}

// One chunk per warp: launches with more chunks than the GPU holds warps
__global__ void __launch_bounds__(128) k_decompress(const ChunkDesc* __restrict__ chunks, uint32_t nchunks, uint32_t* err,
                                                    uint32_t* out_lens) {
    __shared__ SnappyWin win_all[4];
    __shared__ uint8_t hist_all[4][SW_HIST];
    __shared__ uint32_t lut[256];
    lut[threadIdx.x] = snappy_tag_entry(threadIdx.x);
    lut[threadIdx.x + 128] = snappy_tag_entry(threadIdx.x + 128);
    __syncthreads();
    const uint32_t ci = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (ci >= nchunks) return;
    const ChunkDesc& c = chunks[ci];
    const int lane = threadIdx.x & 31;
    const uint8_t* s = (const uint8_t*)c.src;
    uint8_t* d = (uint8_t*)c.dst;
    const uint32_t n = c.src_len;
    uint32_t o = 0, fail = 0;
    if (c.codec == 0) {
        if (n > c.dst_cap) fail = ORCB_UNEXPECTED;
        else warp_copy_fwd(d, s, n, lane);
        o = n;
    } else if (c.codec == 2) {
        uint32_t p;
        uint64_t ulen;
        fail = snappy_preamble(s, n, c.dst_cap, p, ulen);
        if (!fail) {
            SnappyWin& w = win_all[threadIdx.x >> 5];
            uint8_t* hist = hist_all[threadIdx.x >> 5];
            while (p < n && snappy_parse(s, n, d, ulen, p, o, w, lut, lane)) {
                __syncwarp();
                snappy_copy(s, d, w, hist, lane);
                __syncwarp();
            }
            fail = snappy_serial(s, n, d, ulen, p, o, lane);
        }
    } else {
        fail = lz4_block(s, n, d, c.dst_cap, o, lane);
    }
    chunk_done(c, ci, fail, o, err, out_lens, lane);
}

// One chunk per pair of warps (two pairs per block): launches the GPU has warps to spare for
__global__ void __launch_bounds__(128) k_decompress_pair(const ChunkDesc* __restrict__ chunks, uint32_t nchunks, uint32_t* err,
                                                         uint32_t* out_lens) {
    __shared__ SnappyPair pairs[2];
    __shared__ uint32_t lut[256];
    lut[threadIdx.x] = snappy_tag_entry(threadIdx.x);
    lut[threadIdx.x + 128] = snappy_tag_entry(threadIdx.x + 128);
    if (threadIdx.x < 2) {
        pairs[threadIdx.x].head = 0;
        pairs[threadIdx.x].tail = 0;
    }
    __syncthreads();
    const uint32_t pair_in_block = threadIdx.x >> 6;
    const uint32_t ci = blockIdx.x * 2 + pair_in_block;
    if (ci >= nchunks) return;
    const bool parser = ((threadIdx.x >> 5) & 1u) == 0;
    const ChunkDesc& c = chunks[ci];
    const int lane = threadIdx.x & 31;
    const uint8_t* s = (const uint8_t*)c.src;
    uint8_t* d = (uint8_t*)c.dst;
    const uint32_t n = c.src_len;
    SnappyPair& sp = pairs[pair_in_block];
    uint32_t o = 0, fail = 0;
    if (c.codec == 0) {
        // both warps copy, half each
        const uint32_t half = (n / 2) & ~127u;
        if (n > c.dst_cap) fail = ORCB_UNEXPECTED;
        else if (parser) warp_copy_fwd(d, s, half, lane);
        else warp_copy_fwd(d + half, s + half, n - half, lane);
        if (parser) return;
        o = n;
    } else if (c.codec == 2) {
        uint32_t p;
        uint64_t ulen;
        fail = snappy_preamble(s, n, c.dst_cap, p, ulen);  // both warps read it; the copier reports
        if (parser) {
            // ---- parser: descriptors of regular windows, then END or FALLBACK
            if (fail) return;
            uint32_t head = 0;
            for (;;) {
                while (head - ld_volatile(&sp.tail) >= SW_QUEUE) {}
                __threadfence_block();
                SnappyWin& w = sp.win[head % SW_QUEUE];
                uint32_t state = SW_WINDOW;
                if (p >= n) state = SW_END;
                else if (!snappy_parse(s, n, d, ulen, p, o, w, lut, lane)) state = SW_FALLBACK;
                if (lane == 0) {
                    w.state = state;
                    if (state != SW_WINDOW) {
                        w.o = o;
                        w.p = p;
                    }
                }
                __syncwarp();
                __threadfence_block();
                head++;
                if (lane == 0) st_volatile(&sp.head, head);
                if (state != SW_WINDOW) return;
            }
        }
        // ---- copier
        if (!fail) {
            uint32_t tail = 0;
            for (;;) {
                while (ld_volatile(&sp.head) == tail) {}
                __threadfence_block();
                const SnappyWin& w = sp.win[tail % SW_QUEUE];
                const uint32_t state = w.state;
                if (state != SW_WINDOW) {
                    o = w.o;
                    p = state == SW_END ? n : w.p;
                    break;
                }
                snappy_copy(s, d, w, sp.hist, lane);
                __syncwarp();
                __threadfence_block();
                tail++;
                if (lane == 0) st_volatile(&sp.tail, tail);
            }
            fail = snappy_serial(s, n, d, ulen, p, o, lane);
        }
    } else {
        if (parser) return;
        fail = lz4_block(s, n, d, c.dst_cap, o, lane);
    }
    chunk_done(c, ci, fail, o, err, out_lens, lane);
}

// ------------------------------------------------------------------------------------------------
// host-side launch wrappers
// ------------------------------------------------------------------------------------------------
int launch_decompress(const ChunkDesc* c, uint32_t n, uint32_t* err, uint32_t* out_lens, cudaStream_t st) {
    if (!n) return 0;
    // the pair kernel while all its warps are resident at once: 16 blocks of two pairs per SM
    // (ORCB_DECOMP_PAIR=0/1 forces either, for measurements)
    static int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0) sm_count = 148;
    }
    bool pair = n <= (uint32_t)sm_count * 32u;
    if (const char* e = getenv("ORCB_DECOMP_PAIR")) pair = e[0] == '1';
    if (pair) k_decompress_pair<<<(n + 1) / 2, 128, 0, st>>>(c, n, err, out_lens);
    else k_decompress<<<blocks_for_warps(n, 4), 128, 0, st>>>(c, n, err, out_lens);
    LAUNCH_CHECK();
    return 0;
}

}  // namespace orcb
