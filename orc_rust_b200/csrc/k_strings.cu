// String columns: dictionary preparation, tile sums and scan, per-batch offsets and dictionary gather, UTF-8 passes.
#include "kernel_util.cuh"

namespace orcb {

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
// host-side launch wrappers
// ------------------------------------------------------------------------------------------------
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

}  // namespace orcb
