// Launch wrappers implemented in the k_*.cu translation units (each returns 0 or a cudaError_t value).
#pragma once
#include <cuda_runtime.h>

#include "../../include/orc_b200.h"
#include "dev.h"

namespace orcb {

// short-run integer path: header-walk pre-pass into the run table, then one warp per 32 runs
int launch_rle_index(const Seg* segs, uint32_t n, const uint32_t* cnt, RunRec* table, BlockRec* blocks, uint32_t* nblocks,
                     uint32_t pool_blocks, CoopRec* coop_q, uint32_t* ncoop, uint32_t coop_cap, uint32_t* err, SegCheck* chk,
                     cudaStream_t st);
int launch_int_rle(const Seg* segs, const BlockRec* blocks, const uint32_t* nblocks, uint32_t pool_blocks, const RunRec* table,
                   const uint32_t* cnt, const uint32_t* dstart, uint32_t* err, uint32_t* mis, uint32_t* slow_list,
                   uint32_t* slow_count, const CoopRec* coop_q, const uint32_t* ncoop, uint32_t coop_cap, const uint32_t* first,
                   cudaStream_t st);
// first (may be NULL): three device words = (blocks, general blocks, queued runs) an earlier phase already decoded;
// launch_snapshot3 copies the three counters there between the phases
int launch_snapshot3(const uint32_t* src, uint32_t* dst, cudaStream_t st);
// second_pass = 1: only the decimal scale segments of column-stripes whose mismatch flag is set, every value written
int launch_int_rle_coop(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err,
                        uint32_t* mis, int second_pass, SegCheck* chk, cudaStream_t st);
// chk (may be NULL): position-check table, see SegCheck in dev.h
int launch_byte_rle(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err, SegCheck* chk,
                    cudaStream_t st);
int launch_seg_check(const uint2* pairs, uint32_t n, const SegCheck* chk, uint32_t* retry, cudaStream_t st);
int launch_bits(const BitSeg* segs, uint32_t n, uint32_t* cnt, const uint32_t* dstart, cudaStream_t st);
int launch_seg_scan(const ScanDesc* d, uint32_t n, uint32_t* cnt, uint32_t* dstart, cudaStream_t st);
int launch_varint128(const Seg* segs, uint32_t n, const uint32_t* cnt, const uint32_t* dstart, uint32_t* err, SegCheck* chk,
                     cudaStream_t st);
int launch_copy(const CopyDesc* d, const uint2* tiles, uint32_t ntiles, const uint32_t* cnt, uint32_t* err,
                const StrCol* strcols, cudaStream_t st);
int launch_spaced(const SpacedDesc* d, uint32_t n, const uint32_t* dstart, cudaStream_t st);
int launch_popc(const PopcDesc* d, uint32_t n, uint32_t* cnt, cudaStream_t st);
int launch_union_valid(const UnionDesc* d, uint32_t n, uint32_t* meta, cudaStream_t st);
int launch_decimal_fix(const DecFixDesc* d, uint32_t n, const uint32_t* cnt, const uint32_t* mis, cudaStream_t st);
int launch_timestamp(const TsDesc* d, uint32_t n, const uint32_t* cnt, uint32_t* err, cudaStream_t st);
int launch_utf8(const StrCol* cols, const uint2* tiles, uint32_t ntiles, cudaStream_t st);
int launch_strings(StrCol* cols, uint32_t ncols, uint32_t ntiles, uint32_t* err, JobState* state, uint64_t heap_base,
                   uint64_t heap_cap, uint64_t* ptr_table, cudaStream_t st);
int launch_repack(const RepackDesc* d, uint32_t ndesc, uint32_t nwork, uint32_t* nulls, cudaStream_t st);
// The chunk list is ordered: first n_bits Zlib / Zstandard / LZO chunks (k_decompress_bits<codec>, one launch per codec
// of the bit mask `bits_codecs`), then n_snappy Snappy chunks (k_decompress<2>), then LZ4 and stored chunks
// (k_decompress<4>).  counter: eight zeroed words (the persistent warps of each kernel draw chunk indices from their own).
// out_lens[c.id] receives every chunk's decompressed size; *retry is raised when a chunk whose size the planner only
// assumed turned out different (either may be NULL)
int launch_decompress(const ChunkDesc* c, uint32_t n, uint32_t n_bits, uint32_t bits_codecs, uint32_t n_snappy, uint32_t* err, uint32_t* out_lens,
                      uint32_t* counter, uint32_t* retry, cudaStream_t st);

constexpr uint32_t COPY_TILE_BYTES = 16384;

}  // namespace orcb
