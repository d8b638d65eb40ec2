// Integer RLE v1 / v2 (src/encoding/integer/**): whole-warp run decode, run parsing, the run index, the block decoders.
// CUDA kernels of the B200 ORC stripe decoder (sm_100a).  All integer / byte work, HBM-bound:
// no tensor cores.  One warp owns one (stream, row-group) segment; lanes cooperate inside a run.
//
// Semantics follow the reference (datafusion-contrib/orc-rust v0.8.0) bit for bit; each kernel cites
// the functions it replaces.  Error words: first error per column-stripe wins (atomicCAS), value =
// OrcbStatus.
#include "kernel_util.cuh"

namespace orcb {

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
            if (c.scale_lazy && s.out_kind == OUT_SCALE) {
                // the whole run is the type's scale: nothing to write (k_decimal_fix will not look); otherwise raise
                // the flag, the second pass writes every value of the column-stripe
                const uint64_t first = (uint64_t)base + (uint64_t)skip * ustep;
                if (take && ((uint32_t)first != s.aux || (take > 1 && ustep != 0)) && lane == 0) atomicOr(&c.mis[s.colstripe], 1u);
            } else {
                for (uint32_t i = skip + lane; i < i_end; i += 32)
                    store_val(c, out_pos + (i - skip), (int64_t)((uint64_t)base + (uint64_t)i * ustep));
            }
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
                                                   uint32_t* __restrict__ ncoop, uint32_t coop_cap, uint32_t* err,
                                                   uint32_t idx_lanes, SegCheck* __restrict__ chk) {
    // only `idx_lanes` lanes of each warp own a segment: a warp advances at the pace of its slowest lane
    // (the one that misses L1 this step), so fewer streams per warp and more warps hide more latency
    const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
    if ((gt & 31) >= idx_lanes) return;
    const uint32_t segi = (gt >> 5) * idx_lanes + (gt & 31);
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
    // row-index consistency: canonical start now, canonical end after the walk (from the last run's bookkeeping)
    SegCheck ck;
    ck.start_byte = ck.end_byte = cur;
    ck.start_cons = ck.end_cons = skip;
    const uint32_t skip0 = skip;
    bool broken = false;
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
            if (stop) { broken = true; break; }
        }
    }
    if (chk && s.chk) {
        if (slot && !broken) {
            // The last run again (nothing of this sits on the walk's chain): its record says where it starts and how
            // many values came before it; only a segment's first run is entered with values to skip.  Was it used up?
            const RunRec last = slot[-1];
            const uint32_t last_cur = last.byte_off, last_produced = last.out_off & ~RUN_QUEUED;
            const uint32_t last_skip = last_produced == 0 ? skip0 : 0u;
            uint32_t rl = 0, nbytes = 0;
            bool cq;
            if (measure_run(s, last_cur, rl, nbytes, cq) == 0) {
                const uint32_t cons = min(last_skip, rl) + min(rl - min(last_skip, rl), n - last_produced);
                ck.end_byte = cons >= rl ? last_cur + nbytes : last_cur;
                ck.end_cons = cons >= rl ? 0u : cons;
            }
        } else if (broken) {
            ck.end_byte = ck.end_cons = 0xffffffffu;  // the stream failed here: the error is what gets reported
        }
        chk[s.chk - 1] = ck;
    }
}


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
                                                            uint32_t* mis, uint32_t* slow_list, uint32_t* slow_count,
                                                            const uint32_t* __restrict__ first) {
    __shared__ int64_t tile_all[RLE_WARPS][TILE_VALUES];
    const uint32_t nblocks = *nblocks_ptr;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    RunSlot* slots = nullptr;
    // persistent warps: the number of run blocks is only known on the device.  `first`: (blocks, general blocks, queued
    // runs) an earlier phase of this launch has already decoded (the segments that feed the string kernels go first)
    for (uint32_t blk = (first ? first[0] : 0u) + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); blk < nblocks; blk += nwarps) {
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
                                                                    uint32_t* mis, const uint32_t* __restrict__ first) {
    __shared__ uint32_t patchmap_all[RLE_WARPS][16];
    __shared__ RunSlot slots_all[RLE_WARPS][32];
    const uint32_t nslow = *slow_count;
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t i = (first ? first[1] : 0u) + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < nslow; i += nwarps) {
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
                                                              uint32_t* mis, const uint32_t* __restrict__ first) {
    __shared__ uint32_t patchmap_all[RLE_WARPS][16];
    const uint32_t nq = min(*nq_ptr, cap);
    const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t i = (first ? min(first[2], cap) : 0u) + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < nq; i += nwarps) {
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
                                                                 uint32_t* mis, int second_pass,
                                                                 SegCheck* __restrict__ chk) {
    __shared__ uint32_t patchmap_all[RLE_WARPS][16];
    uint32_t* patchmap = patchmap_all[threadIdx.x >> 5];
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nseg) return;
    SegCtx c;
    c.s = &segs[warp];
    c.err = err;
    c.mis = mis;
    const Seg& s = *c.s;
    // second pass: only decimal scale segments of column-stripes where some scale differed, now with every value written
    if (second_pass && (s.out_kind != OUT_SCALE || mis[s.colstripe] == 0)) return;
    c.scale_lazy = !second_pass;
    const uint32_t n = s.cnt_idx >= 0 ? cnt[s.cnt_idx] : s.n_values;
    const uint64_t obase = s.start_idx >= 0 ? dstart[s.start_idx] : s.out_start;
    uint32_t cur = s.start_byte, skip = s.run_skip, produced = 0;
    const bool checking = chk && s.chk && !second_pass;
    SegCheck ck;
    bool started = false;
    ck.start_byte = ck.end_byte = cur;
    ck.start_cons = ck.end_cons = skip;
    // (n == 0: the start is put into canonical form by stepping over the runs the skip covers)
    if (checking && n == 0) {
        while (skip > 0 && cur < s.in_len) {
            uint32_t rl, nbytes;
            bool cq;
            if (measure_run(s, cur, rl, nbytes, cq) || skip < rl) break;
            skip -= rl;
            cur += nbytes;
        }
        ck.start_byte = ck.end_byte = cur;
        ck.start_cons = ck.end_cons = skip;
    }
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
        else {
            if (!started) {  // the first run that yields a value: canonical start
                started = true;
                ck.start_byte = cur;
                ck.start_cons = skip;
            }
            const uint32_t cons = skip + take;
            ck.end_byte = cons >= rl ? cur + nbytes : cur;
            ck.end_cons = cons >= rl ? 0u : cons;
            produced += take;
            skip = 0;
        }
        cur += nbytes;
    }
    if (checking && (threadIdx.x & 31) == 0) chk[s.chk - 1] = ck;
}

// ------------------------------------------------------------------------------------------------
// host-side launch wrappers
// ------------------------------------------------------------------------------------------------
int launch_rle_index(const Seg* segs, uint32_t n, const uint32_t* cnt, RunRec* table, BlockRec* blocks, uint32_t* nblocks,
                     uint32_t pool_blocks, CoopRec* coop_q, uint32_t* ncoop, uint32_t coop_cap, uint32_t* err, SegCheck* chk,
                     cudaStream_t st) {
    if (!n) return 0;
    static uint32_t lanes = 0;
    if (!lanes) {
        lanes = IDX_LANES;
        if (const char* e = getenv("ORCB_IDX_LANES")) {
            const int v = atoi(e);
            if (v >= 1 && v <= 32) lanes = (uint32_t)v;
        }
    }
    const uint32_t nwarps = (n + lanes - 1) / lanes;
    k_rle_index<<<(nwarps + 3) / 4, 128, 0, st>>>(segs, n, cnt, table, blocks, nblocks, pool_blocks, coop_q, ncoop, coop_cap, err, lanes, chk);
    LAUNCH_CHECK();
    return 0;
}
int launch_int_rle(const Seg* segs, const BlockRec* blocks, const uint32_t* nblocks, uint32_t pool_blocks, const RunRec* table,
                   const uint32_t* cnt, const uint32_t* dstart, uint32_t* err, uint32_t* mis, uint32_t* slow_list,
                   uint32_t* slow_count, const CoopRec* coop_q, const uint32_t* ncoop, uint32_t coop_cap, const uint32_t* first,
                   cudaStream_t st) {
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
                                                                                         err, mis, slow_list, slow_count, first);
    LAUNCH_CHECK();
    k_int_rle_general<<<(uint32_t)std::min<uint64_t>(ctas_gen, need), RLE_WARPS * 32, 0, st>>>(segs, blocks, slow_list, slow_count,
                                                                                                table, cnt, dstart, err, mis, first);
    LAUNCH_CHECK();
    if (coop_cap) {
        const uint64_t needq = ((uint64_t)coop_cap + RLE_WARPS - 1) / RLE_WARPS;
        k_coop_runs<<<(uint32_t)std::min<uint64_t>(ctas_gen, needq), RLE_WARPS * 32, 0, st>>>(segs, coop_q, ncoop, coop_cap, cnt, dstart,
                                                                                               err, mis, first);
        LAUNCH_CHECK();
    }
    return 0;
}
// (blocks, general blocks, queued runs) so far -> dst[0..2]: where the next phase of the integer path starts
__global__ void k_snapshot3(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst) {
    if (threadIdx.x < 3) dst[threadIdx.x] = src[threadIdx.x];
}
int launch_snapshot3(const uint32_t* src, uint32_t* dst, cudaStream_t st) {
    k_snapshot3<<<1, 32, 0, st>>>(src, dst);
    LAUNCH_CHECK();
    return 0;
}
int launch_int_rle_coop(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err,
                        uint32_t* mis, int second_pass, SegCheck* chk, cudaStream_t st) {
    if (!n) return 0;
    k_int_rle_coop<<<blocks_for_warps(n, RLE_WARPS), RLE_WARPS * 32, 0, st>>>(segs, n, cnt, dstart, err, mis, second_pass, chk);
    LAUNCH_CHECK();
    return 0;
}

}  // namespace orcb
