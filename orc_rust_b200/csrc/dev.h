// POD descriptor tables shared by the host planner (plan.cc) and the CUDA kernels (k_*.cu).
// One decode job = flat arrays of these, uploaded once; every kernel indexes them by warp.
//
// Pointer-typed fields (uint64_t) hold an arena-tagged offset while planning on the host
// (top 4 bits = arena id, low 60 bits = byte offset) and a device address after relocation.
#pragma once
#include <cstdint>

namespace orcb {

// AR_ZERO: temporaries that must be zero before every launch (atomicOr targets)
// AR_ABS: not an arena - the offset is an absolute device address (buffers of an earlier job, e.g. a parent's validity)
enum Arena : uint64_t { AR_NULL = 0, AR_IN = 1, AR_DEC = 2, AR_OUT = 3, AR_TMP = 4, AR_HEAP = 5, AR_ZERO = 6, AR_ABS = 7 };
static inline uint64_t aref(Arena a, uint64_t off) { return ((uint64_t)a << 60) | off; }

// how an integer-RLE segment stores its values
enum OutKind : uint8_t {
    OUT_I16 = 0,
    OUT_I32 = 1,
    OUT_I64 = 2,
    OUT_LEN31 = 3,   // i32, error `aux` if value outside [0, 2^31)  (string lengths, dictionary keys)
    OUT_SCALE = 4,   // i32, raises the colstripe's scale-mismatch flag when value != aux (decimal SECONDARY)
    OUT_I8 = 5       // byte RLE payload
};

enum SegFlags : uint8_t { SEG_SIGNED = 1, SEG_RLE_V2 = 2 };

// One (stream, row-group) unit for k_int_rle / k_byte_rle / k_varint128.
struct Seg {
    uint64_t in;         // stream base
    uint64_t out;        // element 0 of the destination buffer
    uint32_t in_len;     // stream bytes
    uint32_t start_byte; // entry point (row-index position or 0)
    uint32_t run_skip;   // values already consumed from the run at start_byte
    uint32_t n_values;   // used when cnt_idx < 0
    int32_t cnt_idx;     // >= 0: n_values = cnt[cnt_idx]
    int32_t start_idx;   // >= 0: out_start = dstart[start_idx]
    uint32_t out_start;  // element offset when start_idx < 0
    uint32_t colstripe;  // error word index
    uint8_t flags;
    uint8_t nbytes;      // N of the reference's NInt (2/4/8)
    uint8_t out_kind;
    uint8_t pad0;
    uint32_t aux;
    uint32_t chk;        // > 0: slot + 1 in the position-check table (row-index consistency, see SegCheck)
    uint32_t run_cap;
};

// Row-index consistency.  With the row index in use every (stream, row group) segment starts from the position the
// writer recorded.  The reference decodes sequentially and never looks at those positions, so a damaged stream (or a
// wrong index) would make the two differ.  Every segment therefore leaves where it started and where it stopped, both
// in canonical form: (byte of the run header, values of that run already consumed < run length [| bit << 16 for
// boolean streams]); a segment must stop exactly where the next one of its stream starts, else the job is decoded again
// without the row index (Job::finish, IndexRetry).
struct SegCheck {
    uint32_t start_byte, start_cons, end_byte, end_cons;
};

// A run that needs the whole warp (long DIRECT, DELTA with packed deltas, PATCHED_BASE), queued by the pre-pass
// for k_coop_runs.  Its RunRec in the block table carries RUN_QUEUED in out_off.
struct CoopRec {
    uint32_t seg;
    uint32_t byte_off;
    uint32_t out_off;
    uint32_t skip;
};
constexpr uint32_t RUN_QUEUED = 0x80000000u;

// 32 consecutive runs of one segment: the unit of work of one warp of k_int_rle.  Allocated from a pool by
// the k_rle_index pre-pass (one atomicAdd per 32 runs).
struct BlockRec {
    uint32_t seg;        // index into the short-run segment table
    uint32_t n_runs;     // 1..32
    uint32_t skip;       // values to skip in the block's first run (only the first block of a segment)
    uint32_t pad;
};

// One run found by the k_rle_index header walk: where it starts and where its first emitted value goes.
// The unit of work of one lane of k_int_rle.
struct RunRec {
    uint32_t byte_off;   // offset of the run header inside the stream
    uint32_t out_off;    // element offset from the segment's first output element
};

// MSB-first boolean bytes -> LSB-first bitmap (+ popcount)
struct BitSeg {
    uint64_t src;        // decoded byte-RLE bytes of this segment
    uint64_t dst;        // bitmap base (32-bit words), zero-initialised
    uint32_t bit_skip;   // bits to skip in src
    uint32_t n_bits;     // when cnt_idx < 0
    int32_t cnt_idx;
    int32_t start_idx;   // >= 0: dst_bit0 = dstart[start_idx]
    uint32_t dst_bit0;
    int32_t popc_out;    // >= 0: cnt[popc_out] = number of set bits
};

// exclusive scan of per-group non-null counts of one column-stripe
struct ScanDesc {
    uint32_t base;       // first index in cnt[] / dstart[]; entry base + n_groups receives the total
    uint32_t n_groups;
};

struct CopyDesc {
    uint64_t src;
    uint64_t dst;
    uint64_t n_bytes;    // when cnt_idx < 0
    int32_t cnt_idx;     // >= 0: n_bytes = cnt[cnt_idx] * width
    uint32_t width;
    uint32_t src_len;    // available source bytes (bound; IoError if short)
    uint32_t colstripe;
    int32_t u8_col;      // >= 0: string DATA, validated as UTF-8 on the way through (index of its StrCol)
    uint32_t pad;
};

// dense values -> row slots of one row group (decode_spaced)
struct SpacedDesc {
    uint64_t src;        // dense buffer (element 0 of the colstripe)
    uint64_t dst;        // row-domain buffer
    uint64_t valid;      // stripe-level LSB-first validity bitmap
    uint32_t row0;
    uint32_t n_rows;
    int32_t start_idx;   // dense start of this group = dstart[start_idx]
    uint32_t width;      // bytes per element; 0 = bit mode (boolean values)
};

// elementwise epilogues over the dense domain of one column-stripe
struct DecFixDesc {
    uint64_t vals;       // i128 dense
    uint64_t scales;     // i32 dense
    uint32_t n;          // when cnt_idx < 0
    int32_t cnt_idx;
    uint32_t fixed_scale;
    uint32_t colstripe;  // flag word index (scale mismatch) == error word index
};

struct TsDesc {
    uint64_t secs;       // i64 dense
    uint64_t nanos;      // i64 dense
    uint64_t out;        // i64 (or i128 when as_i128)
    int64_t base;
    int64_t unit_ns;
    uint32_t n;
    int32_t cnt_idx;
    uint32_t colstripe;
    uint32_t as_i128;
    // writer time zone (array_decoder/timestamp.rs:242-286): out = instant + UTC offset in force at the instant
    uint64_t tz_at;      // i64[tz_n] transition instants (UTC seconds), ascending; 0 = no conversion
    uint64_t tz_off;     // i32[tz_n] offset in force from tz_at[i] on
    uint32_t tz_n;
    int32_t tz_first;    // offset before the first transition
    uint32_t tz_on;
    uint32_t pad;
    // A value the zone move pushes out of i64 becomes NULL (try_unary -> unary_opt, array_decoder/timestamp.rs:277-283):
    // bit i of this zero-initialised bitmap (dense domain) is raised and 0 stored; 0 = report an error instead
    uint64_t tznull;
};

// string column of one stripe
struct StrCol {
    uint64_t lens;       // i32 row-domain lengths (direct) or keys (dictionary)
    uint64_t valid;      // stripe-level validity bitmap or 0
    uint64_t dict_len;   // i32[dict_size]      (dictionary mode)
    uint64_t dict_off;   // i32[dict_size + 1]  (dictionary mode; written by k_dict_prepare)
    uint64_t dict_data;  // dictionary bytes
    uint64_t offsets;    // out: i32, batch b at b * (batch_size + 1)
    uint64_t tile_base;  // tmp: i64[n_tiles + 1]
    uint64_t batch_base; // meta: i64[n_batches + 1] absolute byte offset of each batch start
    uint64_t data;       // out: string bytes (direct: static; dictionary: bump-allocated by k_tile_scan)
    uint64_t data_cap;   // capacity behind `data` (direct) / unused
    uint32_t n_rows;
    uint32_t batch_size;
    uint32_t n_batches;
    uint32_t tiles_per_batch;
    uint32_t n_tiles;
    uint32_t dict_size;
    uint32_t dict_data_len;
    uint32_t mode;       // 0 direct, 1 dictionary
    uint32_t colstripe;
    uint32_t tile0;      // first global tile index of this column (for the tile -> column map)
    uint32_t data_len;   // direct: bytes available in the DATA stream
    uint32_t meta_slot;  // index of this column's data pointer in the per-job pointer table
    // UTF-8 validation (GenericByteArray::<Utf8>::try_new, string.rs:150-151): the bytes every value is cut
    // from (DATA stream / dictionary bytes) are checked tile by tile before their real length is known
    uint64_t u8_src;     // bytes to validate, 0 = BINARY column
    uint64_t u8_bad;     // u32[2] in the zeroed arena: [0] max over invalid characters of ~position (0 = none), [1] any non-ASCII byte
    uint64_t u8_flags;   // u32 bitmap in the zeroed arena: U8_TILE-byte tiles that hold non-ASCII bytes
    uint32_t u8_len;     // bytes behind u8_src
    uint32_t u8_pad;
};
constexpr uint32_t U8_TILE = 16384;

// stripe-level bitmap -> per-batch bitmaps (+ null counts)
struct RepackDesc {
    uint64_t src;        // stripe-level bitmap; 0 = every row valid
    uint64_t mask;       // optional: rows whose bit is set here are nulls whatever `src` says (zone-move overflow)
    uint64_t dst;        // batch b at dst + b * dst_stride
    uint32_t dst_stride; // bytes
    uint32_t n_rows;
    uint32_t batch_size;
    uint32_t n_batches;
    int32_t null_out;    // >= 0: nulls[null_out + b] = rows_in_batch - popcount
    uint32_t batch0;     // first global (desc, batch) work index
};

// one compression chunk (src/compression.rs framing, resolved on the host)
struct ChunkDesc {
    uint64_t src;
    uint64_t dst;
    uint32_t src_len;
    uint32_t dst_cap;    // bytes this chunk may produce (exact when known, else block size)
    int32_t expect_len;  // >= 0: decompressed size must equal this (layout was planned on it)
    uint8_t codec;       // 0 original (copy), 1 zlib, 2 snappy, 3 lzo, 4 lz4, 5 zstd
    uint8_t assumed;     // expect_len is the planner's assumption (a non-final chunk fills the block), not a known size:
                         // a different size asks for a re-plan with the sizes found instead of being an error
    uint8_t pad[2];
    uint32_t colstripe;
    uint32_t id;         // slot of this chunk in the table of decompressed sizes
};

// popcount of a validity bitmap into cnt[] (a column whose validity comes from its parent, merged or as is)
struct PopcDesc {
    uint64_t bits;
    uint32_t n_bits;
    uint32_t out;        // cnt[out] = number of set bits
};

// sparse union (array_decoder/union.rs:83-113): the validity every child decodes under.  Child i is valid where
// tag == i; child 0 additionally only where the union itself is valid (null slots carry tag 0).
struct UnionDesc {
    uint64_t tags;       // i8 per slot (null slots 0)
    uint64_t valid;      // the union's own validity bitmap or 0
    uint64_t bits;       // out: n_children bitmaps, `stride` bytes apart, zero-initialised
    uint32_t n;
    uint32_t n_children;
    uint32_t stride;
    uint32_t counts;     // meta: nulls[counts + i] = set bits of child i's bitmap
};

// job-wide mutable device state
struct JobState {
    unsigned long long heap_top;
    unsigned long long reserved;
};

// tile size (rows) of the string offset scan
constexpr uint32_t STR_TILE = 1024;

}  // namespace orcb
