// Byte / boolean RLE, PRESENT scans, decimal varints, raw copies (+ UTF-8 check), spaced placement, decimal scale fix,
// timestamps, per-batch bitmaps.
#include "kernel_util.cuh"

namespace orcb {

// ------------------------------------------------------------------------------------------------
// Byte RLE (encoding/byte.rs:228-247): TINYINT data and the byte layer under boolean RLE.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RLE_WARPS * 32) k_byte_rle(const Seg* __restrict__ segs, uint32_t nseg,
                                                             const uint32_t* __restrict__ cnt,
                                                             const uint32_t* __restrict__ dstart, uint32_t* err,
                                                             SegCheck* __restrict__ chk) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nseg) return;
    const Seg& s = segs[warp];
    const int lane = threadIdx.x & 31;
    const uint8_t* in = (const uint8_t*)s.in;
    const uint32_t len = s.in_len;
    // for boolean streams n_values counts BITS; aux = 1 marks "bits": convert to bytes incl. the bit offset
    uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_values;
    const uint32_t bit0 = (s.aux & 1) ? (s.aux >> 1) : 0u;  // bit_skip
    // bytes used up completely: where the next row group of the stream starts (a partly used byte is its first)
    const uint32_t full = (s.aux & 1) ? (n + bit0) / 8 : n;
    const uint32_t end_bit = (s.aux & 1) ? (n + bit0) & 7u : 0u;
    if (s.aux & 1) n = (n + bit0 + 7) / 8;
    uint8_t* out = (uint8_t*)s.out + (s.start_idx >= 0 ? dstart[s.start_idx] : s.out_start);
    uint32_t cur = s.start_byte, skip = s.run_skip, produced = 0;
    const bool checking = chk && s.chk;
    SegCheck ck;
    if (checking) {
        // canonical start: step over the runs the skip covers entirely
        while (cur < len) {
            const uint32_t h = in[cur];
            const uint32_t rl = h < 0x80 ? h + 3 : 0x100 - h;
            if (skip < rl) break;
            skip -= rl;
            cur += h < 0x80 ? 2u : 1u + rl;
        }
        ck.start_byte = ck.end_byte = cur;
        ck.start_cons = skip | (bit0 << 16);
        ck.end_cons = skip | (end_bit << 16);
    }
    while (produced < n) {
        if (cur >= len) { set_err(err, s.colstripe, ORCB_IO_ERROR); return; }
        const uint32_t h = in[cur];
        uint32_t rl, run_bytes, take;
        const uint32_t skip0 = skip;
        if (h < 0x80) {
            rl = h + 3;
            run_bytes = 2;
            if (cur + 2 > len) { set_err(err, s.colstripe, ORCB_IO_ERROR); return; }
            const uint8_t v = in[cur + 1];
            const uint32_t avail = rl > skip ? rl - skip : 0;
            take = min(avail, n - produced);
            for (uint32_t i = lane; i < take; i += 32) out[produced + i] = v;
        } else {
            rl = 0x100 - h;
            run_bytes = 1 + rl;
            if (cur + run_bytes > len) { set_err(err, s.colstripe, ORCB_IO_ERROR); return; }
            const uint32_t avail = rl > skip ? rl - skip : 0;
            take = min(avail, n - produced);
            for (uint32_t i = lane; i < take; i += 32) out[produced + i] = in[cur + 1 + skip + i];
        }
        if (skip >= rl) {
            skip -= rl;
        } else {
            if (checking && full > produced && full <= produced + take) {
                // the byte position `full` bytes into the segment falls in this run
                const uint32_t cons = skip0 + (full - produced);
                ck.end_byte = cons >= rl ? cur + run_bytes : cur;
                ck.end_cons = (cons >= rl ? 0u : cons) | (end_bit << 16);
            }
            produced += take;
            skip = 0;
        }
        cur += run_bytes;
    }
    if (checking && lane == 0) chk[s.chk - 1] = ck;
}

// Row-index consistency: a segment must stop where the next segment of its stream starts (SegCheck, dev.h)
__global__ void k_seg_check(const uint2* __restrict__ pairs, uint32_t npairs, const SegCheck* __restrict__ chk, uint32_t* retry) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npairs) return;
    const SegCheck a = chk[pairs[i].x], b = chk[pairs[i].y];
    if (a.end_byte == 0xffffffffu) return;  // the segment failed: its error is reported
    if (a.end_byte != b.start_byte || a.end_cons != b.start_cons) atomicOr(retry, 2u);
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

// popcount of a bitmap (validity handed down by a parent column)
__global__ void __launch_bounds__(128) k_popc(const PopcDesc* __restrict__ descs, uint32_t ndesc, uint32_t* cnt) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ndesc) return;
    const PopcDesc d = descs[warp];
    const int lane = threadIdx.x & 31;
    const uint32_t* bm = (const uint32_t*)d.bits;
    uint32_t pc = 0;
    const uint32_t nw = (d.n_bits + 31) / 32;
    for (uint32_t w = lane; w < nw; w += 32) {
        uint32_t v = bm[w];
        if (w == nw - 1 && (d.n_bits & 31)) v &= (1u << (d.n_bits & 31)) - 1u;
        pc += __popc(v);
    }
    pc = (uint32_t)warp_sum64(pc);
    if (lane == 0) cnt[d.out] = pc;
}

// validity of the children of a sparse union (union.rs:83-113)
__global__ void __launch_bounds__(128) k_union_valid(const UnionDesc* __restrict__ descs, uint32_t ndesc, uint32_t* meta) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= ndesc) return;
    const UnionDesc d = descs[warp];
    const int lane = threadIdx.x & 31;
    const int8_t* tags = (const int8_t*)d.tags;
    const uint32_t* valid = (const uint32_t*)d.valid;
    for (uint32_t c = 0; c < d.n_children; c++) {
        uint32_t* out = (uint32_t*)(d.bits + (uint64_t)c * d.stride);
        uint32_t pc = 0;
        for (uint32_t r0 = 0; r0 < d.n; r0 += 32) {
            const uint32_t r = r0 + lane;
            bool on = r < d.n && (uint32_t)(int32_t)tags[r] == c;
            if (on && c == 0 && valid) on = (valid[r >> 5] >> (r & 31)) & 1u;
            const uint32_t word = __ballot_sync(FULL, on);
            if (lane == 0) out[r0 >> 5] = word;
            pc += __popc(word);
        }
        if (lane == 0) meta[d.counts + c] = pc;
    }
}

// ------------------------------------------------------------------------------------------------
// Decimal DATA: unbounded zigzag varints -> i128 (encoding/decimal.rs:46-51, integer/util.rs:475-527).
// Terminator bytes found with ballot; the lane owning a terminator assembles its value.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_varint128(const Seg* __restrict__ segs, uint32_t nseg,
                                                   const uint32_t* __restrict__ cnt,
                                                   const uint32_t* __restrict__ dstart, uint32_t* err,
                                                   SegCheck* __restrict__ chk) {
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
    const bool checking = chk && s.chk;
    if (n == 0) {
        if (checking && lane == 0) chk[s.chk - 1] = SegCheck{cur, 0u, cur, 0u};
        return;
    }
    const uint32_t start_byte = cur;
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
        // (the segment ends behind its last value, wherever the window's last terminator lies)
        if (checking && total >= room && lane == 0) chk[s.chk - 1] = SegCheck{start_byte, 0u, cur + (uint32_t)ends[room - 1] + 1, 0u};
        produced += total;
        cur += (uint32_t)ends[total - 1] + 1;
        __syncwarp();
    }
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
                if (w > (__int128)INT64_MAX || w < (__int128)INT64_MIN) {
                    // out of the unit's range: the reference makes the value a null (timestamp.rs:277-283)
                    if (d.tznull) atomicOr((uint32_t*)d.tznull + (i >> 5), 1u << (i & 31));
                    else set_err(err, d.colstripe, ORCB_NOT_IMPLEMENTED);
                    v = 0;
                } else {
                    v = (int64_t)w;
                }
            }
            ((int64_t*)d.out)[i] = v;
        }
    }
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
    const uint32_t* mask = (const uint32_t*)d.mask;
    uint32_t* dst = (uint32_t*)((uint8_t*)d.dst + (uint64_t)b * d.dst_stride);
    const uint32_t nwords = (rows + 31) >> 5;
    uint32_t pc = 0;
    for (uint32_t w = lane; w < nwords; w += 32) {
        uint32_t v = src ? load_bits32(src, (uint64_t)row0 + (uint64_t)w * 32) : 0xffffffffu;
        if (mask) v &= ~load_bits32(mask, (uint64_t)row0 + (uint64_t)w * 32);
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
// host-side launch wrappers
// ------------------------------------------------------------------------------------------------
int launch_byte_rle(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err, SegCheck* chk,
                    cudaStream_t st) {
    if (!n) return 0;
    k_byte_rle<<<blocks_for_warps(n, RLE_WARPS), RLE_WARPS * 32, 0, st>>>(segs, n, cnt, dstart, err, chk);
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
int launch_varint128(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err, SegCheck* chk,
                     cudaStream_t st) {
    if (!n) return 0;
    k_varint128<<<blocks_for_warps(n, 4), 128, 0, st>>>(segs, n, cnt, dstart, err, chk);
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
int launch_repack(const RepackDesc* d, uint32_t ndesc, uint32_t nwork, uint32_t* nulls, cudaStream_t st) {
    if (!nwork) return 0;
    k_repack<<<blocks_for_warps(nwork, 4), 128, 0, st>>>(d, ndesc, nwork, nulls);
    LAUNCH_CHECK();
    return 0;
}

int launch_seg_check(const uint2* pairs, uint32_t n, const SegCheck* chk, uint32_t* retry, cudaStream_t st) {
    if (!n) return 0;
    k_seg_check<<<(n + 255) / 256, 256, 0, st>>>(pairs, n, chk, retry);
    LAUNCH_CHECK();
    return 0;
}
int launch_popc(const PopcDesc* d, uint32_t n, uint32_t* cnt, cudaStream_t st) {
    if (!n) return 0;
    k_popc<<<blocks_for_warps(n, 4), 128, 0, st>>>(d, n, cnt);
    LAUNCH_CHECK();
    return 0;
}
int launch_union_valid(const UnionDesc* d, uint32_t n, uint32_t* meta, cudaStream_t st) {
    if (!n) return 0;
    k_union_valid<<<blocks_for_warps(n, 4), 128, 0, st>>>(d, n, meta);
    LAUNCH_CHECK();
    return 0;
}

}  // namespace orcb
