// Compression chunks: original copies, Snappy and LZ4 block decoding.
#include "kernel_util.cuh"

namespace orcb {

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
int launch_decompress(const ChunkDesc* c, uint32_t n, uint32_t* err, uint32_t* out_lens, cudaStream_t st) {
    if (!n) return 0;
    k_decompress<<<blocks_for_warps(n, 4), 128, 0, st>>>(c, n, err, out_lens);
    LAUNCH_CHECK();
    return 0;
}

}  // namespace orcb
