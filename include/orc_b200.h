/*
 * orc_b200.h — C ABI of the B200-native ORC stripe decoder (drop-in boundary for orc-rust's decode path).
 *
 * This is exactly what a Rust `orc-rust-cuda` crate's `extern "C"` block would bind.
 * Plain pointers and sizes only; Arrow data crosses as the Arrow C Data / C Device Data interface.
 * Reference citations are file:line in datafusion-contrib/orc-rust v0.8.0.
 *
 * Threading: handles are not thread-safe; distinct handles may be used from distinct threads.
 * Errors: every call returns 0 (ORCB_OK) or an OrcbStatus; details via orcb_last_error().  Nothing
 * panics or aborts across this boundary (reference panics are reported as a status: ORCB_OUT_OF_SPEC on the decode
 * path, ORCB_UNEXPECTED with the panic's message for with_predicate / row-selection combinations).
 */
#ifndef ORC_B200_H
#define ORC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status = orc-rust `OrcError` variant ordinal + 1 (src/error.rs:31-174), then device-path additions. */
typedef enum OrcbStatus {
    ORCB_OK = 0,
    ORCB_IO_ERROR = 1,
    ORCB_EMPTY_FILE = 2,
    ORCB_OUT_OF_SPEC = 3,
    ORCB_DECODE_FLOAT = 4,
    ORCB_DECODE_TIMESTAMP = 5,
    ORCB_OFFSET_OVERFLOW = 6,
    ORCB_DECODE_PROTO = 7,
    ORCB_NO_TYPES = 8,
    ORCB_UNSUPPORTED_TYPE_VARIANT = 9,
    ORCB_MISMATCHED_SCHEMA = 10,
    ORCB_CONVERT_RECORD_BATCH = 11,
    ORCB_VARINT_TOO_LARGE = 12,
    ORCB_UNEXPECTED = 13,
    ORCB_BUILD_ZSTD_DECODER = 14,
    ORCB_BUILD_SNAPPY_DECODER = 15,
    ORCB_BUILD_LZO_DECODER = 16,
    ORCB_BUILD_LZ4_DECODER = 17,
    ORCB_ARROW = 18,
    /* device-path additions.  UNSUPPORTED_DEVICE_CODEC is no longer produced: every compression kind is decoded on the
     * device (kept so that the numbering stays stable) */
    ORCB_UNSUPPORTED_DEVICE_CODEC = 19,
    ORCB_CUDA = 20,
    ORCB_INVALID_ARGUMENT = 21,
    ORCB_NOT_IMPLEMENTED = 22,
    ORCB_DEVICE_HEAP_OVERFLOW = 23
} OrcbStatus;

/* ---- Arrow C Data Interface (ABI-stable structs from the Arrow specification) ---- */
#ifndef ARROW_C_DATA_INTERFACE
#define ARROW_C_DATA_INTERFACE
struct ArrowSchema {
    const char* format;
    const char* name;
    const char* metadata;
    int64_t flags;
    int64_t n_children;
    struct ArrowSchema** children;
    struct ArrowSchema* dictionary;
    void (*release)(struct ArrowSchema*);
    void* private_data;
};
struct ArrowArray {
    int64_t length;
    int64_t null_count;
    int64_t offset;
    int64_t n_buffers;
    int64_t n_children;
    const void** buffers;
    struct ArrowArray** children;
    struct ArrowArray* dictionary;
    void (*release)(struct ArrowArray*);
    void* private_data;
};
#endif
#ifndef ARROW_C_DEVICE_DATA_INTERFACE
#define ARROW_C_DEVICE_DATA_INTERFACE
typedef int32_t ArrowDeviceType;
#define ARROW_DEVICE_CPU 1
#define ARROW_DEVICE_CUDA 2
struct ArrowDeviceArray {
    struct ArrowArray array;
    int64_t device_id;
    ArrowDeviceType device_type;
    void* sync_event; /* cudaEvent_t* or NULL */
    int64_t reserved[3];
};
#endif

typedef struct OrcbFile OrcbFile;     /* parsed file tail (FileMetadata, src/reader/metadata.rs:63-178) */
typedef struct OrcbReader OrcbReader; /* ArrowReader (src/arrow_reader.rs:233-347) with a device */
typedef struct OrcbJob OrcbJob;       /* one device launch plan over many stripes (new; no reference twin) */

/* ---- file open: replaces ArrowReaderBuilder::try_new -> read_metadata
 *      (src/arrow_reader.rs:202-205, src/reader/metadata.rs:180-263) ---- */
/* The library borrows `data` for the life of the OrcbFile (the ChunkReader for Bytes, src/reader/mod.rs:64-76). */
int orcb_open_memory(const uint8_t* data, size_t len, OrcbFile** out);
/* Reads the whole file into (pinned, if a device is present) host memory owned by the handle
 * (ChunkReader for File, src/reader/mod.rs:48-62). */
int orcb_open_path(const char* path, OrcbFile** out);
/* The file behind a read callback: ChunkReader (src/reader/mod.rs:27-46) as a C function.  `read_at` fills dst with
 * `length` bytes from `offset` and returns 0, or a non-zero status that surfaces as ORCB_IO_ERROR; it may be called from
 * any thread that uses the handle (readers built from one handle call it one at a time per reader).  Open reads the
 * tail (the last 16 KiB, then the exact footer range if that was short: read_metadata, src/reader/metadata.rs:180-236);
 * decoding reads each stripe with ONE call (index, data and footer together - the reference issues one get_bytes per
 * stream, src/stripe.rs:161) straight into pinned memory the stripe is staged from; with_predicate reads only the index
 * area and the footer of a stripe before deciding whether any of it is needed.  `ctx` must outlive the handle. */
typedef int (*OrcbReadAt)(void* ctx, uint64_t offset, uint64_t length, uint8_t* dst);
int orcb_open_callbacks(uint64_t len, OrcbReadAt read_at, void* ctx, OrcbFile** out);
/* out[0] = read_at calls so far, out[1] = bytes they returned (0, 0 for files opened from memory or a path) */
int orcb_file_io_stats(const OrcbFile* f, uint64_t out[2]);
/* A second handle on the same bytes (they stay with `f`, which must outlive the clone).  A bulk job stages every
 * handle's stripes separately, so clones let one host copy of a file stand in for several files of a larger set. */
int orcb_file_clone(const OrcbFile* f, OrcbFile** out);
void orcb_file_free(OrcbFile* f);

/* FileMetadata accessors (src/reader/metadata.rs:141-178) */
uint64_t orcb_file_num_rows(const OrcbFile* f);
uint32_t orcb_file_num_stripes(const OrcbFile* f);
int32_t orcb_file_compression(const OrcbFile* f);          /* proto CompressionKind */
uint64_t orcb_file_compression_block_size(const OrcbFile* f);
int64_t orcb_file_row_index_stride(const OrcbFile* f);     /* -1 when absent */
uint32_t orcb_file_num_root_columns(const OrcbFile* f);
const char* orcb_file_root_column_name(const OrcbFile* f, uint32_t i);
/* ORC column index of root column i (what ProjectionMask::roots takes, src/projection.rs:37-50); 0 when out of range */
uint32_t orcb_file_root_column_id(const OrcbFile* f, uint32_t i);
/* FileMetadata::file_format_version (src/reader/metadata.rs:119-127, 175-177): "0.12", "" when the postscript has none */
const char* orcb_file_format_version(const OrcbFile* f);
/* FileMetadata::user_custom_metadata (src/reader/metadata.rs:112-117, 161-163): entry i of *orcb_file_num_user_metadata*,
 * in file order (the same pairs travel as the metadata of the Arrow schema).  `key` is NUL-terminated; the value is
 * `*value_len` bytes and need not be text.  Both stay valid for the life of the handle. */
uint32_t orcb_file_num_user_metadata(const OrcbFile* f);
int orcb_file_user_metadata(const OrcbFile* f, uint32_t i, const char** key, const uint8_t** value, size_t* value_len);
/* StripeMetadata (src/stripe.rs:38-81): out[0..5) = offset, index_length, data_length, footer_length, rows */
int orcb_file_stripe_info(const OrcbFile* f, uint32_t stripe, uint64_t out[5]);

/* ---- reader options: the ArrowReaderBuilder setters (src/arrow_reader.rs:39-198) + `device` ---- */
typedef struct OrcbReadOptions {
    int32_t device;            /* CUDA ordinal (with_device; new) */
    uint32_t batch_size;       /* with_batch_size; 0 = default 8192 (src/arrow_reader.rs:37) */
    const char* const* projection_names; /* with_projection(ProjectionMask::named_roots) or NULL = all */
    uint32_t n_projection;
    uint64_t range_start;      /* with_file_byte_range: stripes whose offset is in [start, end) */
    uint64_t range_end;        /* 0,0 = no range */
    int32_t timestamp_unit;    /* with_timestamp_precision: 0 = nanosecond (default), 1 = microsecond;
                                  2 = millisecond, 3 = second (with_schema overrides) */
    int32_t use_row_index;     /* 1 (default when 0 passed with flags==0): split streams at row-index positions */
    int32_t device_resident;   /* 0: batches come back in host memory; 1: ArrowDeviceArray in HBM */
    uint32_t max_stripes_per_launch; /* 0 = library default */
    void* cuda_stream;         /* cudaStream_t to enqueue on, or NULL for the reader's own stream */
    uint32_t flags;            /* bit0: no_row_index */
    uint32_t stripe_shard_index;  /* multi-GPU stripe sharding: keep stripes with */
    uint32_t stripe_shard_count;  /*   (ordinal % count) == index; count 0 or 1 = all.  A reader counts the stripes
                                     of its file (after the byte range), a bulk job the stripes of all its files in order */
    uint32_t waves;            /* bulk jobs: stripe waves kept in flight on separate streams (0 = automatic) */
    uint32_t reserved0;
} OrcbReadOptions;

/* ArrowReaderBuilder::schema (src/arrow_reader.rs:182-198) for the given options */
int orcb_schema(const OrcbFile* f, const OrcbReadOptions* opt, struct ArrowSchema* out);

/* ArrowReaderBuilder::build (src/arrow_reader.rs:207-230) */
int orcb_reader_new(OrcbFile* f, const OrcbReadOptions* opt, OrcbReader** out);
/* ArrowReaderBuilder::with_row_selection (src/arrow_reader.rs:113-116; RowSelector src/row_selection.rs:32-57):
 * runs of rows to skip / to read, counted over the stripes the reader visits.  Batches follow
 * NaiveStripeDecoder::next_with_row_selection (src/array_decoder/mod.rs:313-364) exactly; they are views into
 * the decoded stripe (Arrow `offset` != 0), stripes without a selected row are never staged. */
typedef struct OrcbRowSelector {
    uint64_t row_count;
    int32_t skip;      /* 1 = skip these rows, 0 = read them */
    int32_t reserved;
} OrcbRowSelector;
int orcb_reader_new_with_selection(OrcbFile* f, const OrcbReadOptions* opt, const OrcbRowSelector* selectors,
                                   uint32_t n_selectors, OrcbReader** out);
/* The builder with everything set: selection as above when has_selection != 0, and ArrowReaderBuilder::with_schema
 * (src/arrow_reader.rs:80-83) when `schema` is not NULL: one field per projected column; every type must be the
 * default one except timestamps, which may be asked for in any unit (TIMESTAMP: no zone; TIMESTAMP WITH LOCAL TIME
 * ZONE: "UTC") or as Decimal128(38, 9) nanoseconds (src/array_decoder/timestamp.rs:150-232); anything else is
 * MismatchedSchema / UnsupportedTypeVariant as in array_decoder_factory (src/array_decoder/mod.rs:390-511).
 * The schema is only read; the caller keeps and releases it. */
int orcb_reader_new_ex(OrcbFile* f, const OrcbReadOptions* opt, const OrcbRowSelector* selectors, uint32_t n_selectors,
                       int has_selection, const struct ArrowSchema* schema, OrcbReader** out);
/* Host-only: the batches a selection yields, as (stripe ordinal, first row, rows) triples in output order, for stripes
 * of the given row counts; applies[s] = 0 where the stripe is read whole because the selection was used up before it
 * (src/arrow_reader.rs:296-309).  *n_triples is the number of triples (also when it exceeds cap_triples). */
int orcb_selection_plan(const OrcbRowSelector* selectors, uint32_t n_selectors, const uint64_t* stripe_rows, uint32_t n_stripes,
                        uint64_t batch_size, int32_t* applies, uint64_t* triples, size_t cap_triples, size_t* n_triples);
/* ArrowReaderBuilder::with_predicate (src/arrow_reader.rs:140-176): a predicate over the projected top-level columns,
 * evaluated per stripe against the row groups' statistics and Bloom filters (src/row_group_filter.rs); the row groups
 * that cannot match are neither staged nor decoded.  The tree (src/predicate.rs:28-104) is passed in pre-order: a
 * node, then the subtrees of its n_children children. */
enum OrcbPredicateKind { ORCB_PRED_COMPARISON = 0, ORCB_PRED_IS_NULL = 1, ORCB_PRED_IS_NOT_NULL = 2, ORCB_PRED_AND = 3,
                         ORCB_PRED_OR = 4, ORCB_PRED_NOT = 5 };
enum OrcbComparisonOp { ORCB_OP_EQ = 0, ORCB_OP_NE = 1, ORCB_OP_LT = 2, ORCB_OP_LE = 3, ORCB_OP_GT = 4, ORCB_OP_GE = 5 };
enum OrcbPredicateValueType { ORCB_VAL_BOOLEAN = 0, ORCB_VAL_INT8 = 1, ORCB_VAL_INT16 = 2, ORCB_VAL_INT32 = 3,
                              ORCB_VAL_INT64 = 4, ORCB_VAL_FLOAT32 = 5, ORCB_VAL_FLOAT64 = 6, ORCB_VAL_UTF8 = 7 };
typedef struct OrcbPredicateNode {
    int32_t kind;          /* OrcbPredicateKind */
    int32_t op;            /* OrcbComparisonOp (comparisons) */
    int32_t value_type;    /* OrcbPredicateValueType (comparisons) */
    int32_t value_is_null; /* PredicateValue::X(None) */
    int64_t i64;           /* Boolean (0 / 1), Int8 .. Int64 */
    double f64;            /* Float32 (widened, as the reference does before comparing) and Float64 */
    const char* column;    /* comparisons, IS NULL, IS NOT NULL: NUL-terminated column name */
    const uint8_t* str;    /* Utf8: the value's bytes */
    uint64_t str_len;
    uint32_t n_children;   /* AND / OR: any number; NOT: 1; leaves: 0 */
    uint32_t reserved;
} OrcbPredicateNode;
/* Everything ArrowReaderBuilder can be given, in one call; zero / NULL members mean "not set". */
typedef struct OrcbReaderBuild {
    const OrcbReadOptions* options;
    const OrcbRowSelector* selectors; /* with_row_selection */
    uint32_t n_selectors;
    int32_t has_selection;
    const struct ArrowSchema* schema; /* with_schema */
    const OrcbPredicateNode* predicate; /* with_predicate */
    uint32_t n_predicate_nodes;
    uint32_t reserved;
} OrcbReaderBuild;
int orcb_reader_build(OrcbFile* f, const OrcbReaderBuild* build, OrcbReader** out);
/* Host-only: the verdict of a predicate on one stripe.  *evaluated = 0 when the reference would fall back to reading
 * the whole stripe (no row index, column not projected, value of the wrong type ...); else keep[g] = 1 for the row
 * groups that may hold matching rows, *n_groups of them (also when that exceeds cap). */
int orcb_predicate_row_groups(OrcbFile* f, uint32_t stripe, const OrcbReadOptions* opt, const OrcbPredicateNode* predicate,
                              uint32_t n_predicate_nodes, uint8_t* keep, size_t cap, size_t* n_groups, int* evaluated);
/* Host-only: what a reader built with a selection and / or a predicate will yield, in the form of
 * orcb_selection_plan: triples (ordinal among the stripes the reader visits, first row, rows); applies[s] = 0 for a
 * stripe that is read whole.  *n_stripes / *n_triples are the full counts, also when they exceed the capacities. */
int orcb_reader_plan(OrcbReader* r, int32_t* applies, size_t cap_stripes, size_t* n_stripes, uint64_t* triples,
                     size_t cap_triples, size_t* n_triples);
/* Host-only: BloomFilter::hash_long / hash_bytes (src/bloom_filter.rs:136-149, 182-230) */
uint64_t orcb_bloom_hash_long(int64_t value);
uint64_t orcb_bloom_hash_bytes(const uint8_t* bytes, size_t len);
/* Work done so far: out[0] = (stream, row-group) segments planned, out[1] = stripe tasks staged.  With a selection only
 * the row groups that hold selected rows are decoded, so both shrink with it. */
int orcb_reader_counters(const OrcbReader* r, uint64_t out[2]);
void orcb_reader_free(OrcbReader* r);
/* ArrowReader::total_row_count (src/arrow_reader.rs:243-247) */
uint64_t orcb_reader_total_row_count(const OrcbReader* r);
/* Iterator::next (src/arrow_reader.rs:333-346): one RecordBatch as a struct ArrowArray in host memory.
 * *eos = 1 and `out` untouched at end of stream. */
int orcb_reader_next(OrcbReader* r, struct ArrowArray* out, int* eos);
/* ArrowStreamReader::poll_next (src/async_arrow_reader.rs:283-290, 252-277) as a completion callback: starts producing
 * the next batch on a thread of the library and returns at once; `done(ctx, status, eos, error)` is called from that
 * thread when `out` is filled (status 0, eos 0), the stream has ended (eos 1) or failed (status != 0; the reader stays in
 * the error state, StreamState::Error).  One request per reader at a time; `out` and `ctx` must stay valid until `done`. */
typedef void (*OrcbBatchCallback)(void* ctx, int status, int eos, const char* error);
int orcb_reader_next_async(OrcbReader* r, struct ArrowArray* out, OrcbBatchCallback done, void* ctx);
/* Same, batch buffers stay in HBM (device_type = ARROW_DEVICE_CUDA).  Requires device_resident = 1. */
int orcb_reader_next_device(OrcbReader* r, struct ArrowDeviceArray* out, int* eos);

/* Drains the reader inside the library: every remaining batch is produced as orcb_reader_next (host-resident readers,
 * device-to-host copy included) or orcb_reader_next_device would, and released at once.  out[0] = batches,
 * out[1] = rows.  The `for batch in reader {}` of benches/arrow_reader.rs:53-59 without per-batch FFI calls. */
int orcb_reader_drain(OrcbReader* r, uint64_t out[2]);

/* ---- bulk job API: NaiveStripeDecoder::new_with_selection + drain (src/array_decoder/mod.rs:570-594,
 *      371-387) for many stripes in one launch plan.  Used by the reader internally and by bench.py. ---- */
int orcb_job_new(OrcbFile* const* files, uint32_t n_files, const OrcbReadOptions* opt, OrcbJob** out);
void orcb_job_free(OrcbJob* j);
/* Host planning only (no CUDA calls): builds every descriptor table. */
int orcb_job_plan(OrcbJob* j);
/* Allocates device arenas and copies compressed stripe bytes + descriptor tables H2D (async on the stream). */
int orcb_job_stage(OrcbJob* j);
/* Enqueues every decode kernel on the stream; returns without synchronising. */
int orcb_job_launch(OrcbJob* j);
/* Copies per-batch metadata + error words D2H, synchronises, maps device error words to OrcbStatus. */
int orcb_job_finish(OrcbJob* j);
/* Plan statistics: see OrcbJobStats. */
typedef struct OrcbJobStats {
    uint64_t n_stripes, n_rows, n_columns;
    uint64_t input_bytes;       /* Σ stored bytes of projected non-index streams */
    uint64_t staged_bytes;      /* bytes copied H2D per stage() */
    uint64_t output_bytes;      /* Σ logical Arrow buffer bytes (valid after finish()) */
    uint64_t device_bytes;      /* arena capacity allocated in HBM */
    uint64_t n_segments;        /* (stream, row-group) work units */
    uint64_t n_kernel_launches; /* kernels enqueued by one launch() */
    uint64_t n_batches;
    uint64_t d2h_meta_bytes;    /* per-batch metadata + error words read back in finish() */
    uint64_t aliased_output_bytes; /* part of output_bytes that no kernel writes: values buffers of direct string
                                      columns are views of the staged / decompressed DATA stream (valid after finish()) */
    uint64_t n_waves;
} OrcbJobStats;
int orcb_job_stats(const OrcbJob* j, OrcbJobStats* out);
/* Re-copies the compressed stripe bytes H2D into the already allocated arenas (end-to-end timing loops). */
int orcb_job_restage(OrcbJob* j);
/* Per-kernel device time of the last launch() (CUDA events recorded on the job's stream around every
 * kernel) and the algorithmic bytes each kernel has to move: stream bytes it consumes + buffer bytes it
 * produces.  Valid after finish(). */
typedef struct OrcbKernelStat {
    char name[32];
    double ms;
    uint64_t alg_bytes;
    uint64_t work_items; /* warps / CTAs launched */
} OrcbKernelStat;
int orcb_job_kernel_stats(const OrcbJob* j, OrcbKernelStat* out, uint32_t cap, uint32_t* n);
uint64_t orcb_job_num_batches(const OrcbJob* j);
/* Export batch `i` (stripe-major order, reference batch boundaries).  Host copy happens lazily, per job. */
int orcb_job_export_batch(OrcbJob* j, uint64_t i, struct ArrowArray* out);
int orcb_job_export_batch_device(OrcbJob* j, uint64_t i, struct ArrowDeviceArray* out);

/* ---- stream-level entry points (one kernel each; used by the parity tests against the reference's
 *      unit-test vectors, src/encoding/ ** /tests) ---- */
/* kind: 0 RLEv1, 1 RLEv2 (src/encoding/integer); is_signed per EncodingSign; nbytes 2/4/8 = N. */
int orcb_decode_int_rle(int device, const uint8_t* in, size_t in_len, int version, int is_signed, int nbytes,
                        int64_t* out, size_t n_values);
int orcb_decode_byte_rle(int device, const uint8_t* in, size_t in_len, uint8_t* out, size_t n_values);
/* boolean RLE -> LSB-first Arrow bitmap of n_values bits (src/encoding/boolean.rs:101-113) */
int orcb_decode_bool_rle(int device, const uint8_t* in, size_t in_len, uint8_t* out_bitmap, size_t n_values);
/* decimal DATA: zigzag varint -> i128 little-endian, 16 bytes each (src/encoding/decimal.rs:46-51) */
int orcb_decode_varint128(int device, const uint8_t* in, size_t in_len, uint8_t* out16, size_t n_values);
/* chunk framing + block decompression of one whole stream (src/compression.rs:244-347);
 * *out_len receives the decompressed size; out_cap must be >= chunks * block_size. */
int orcb_decompress_stream(int device, int compression_kind, const uint8_t* in, size_t in_len, size_t block_size,
                           uint8_t* out, size_t out_cap, size_t* out_len);

/* The host decoder of METADATA sections (postscript-described footer, stripe footers, row indexes: parsed on the host
 * like the reference does, src/reader/metadata.rs:238-263, src/stripe.rs:215-244), exposed so that it can be tested
 * without a GPU.  Data streams never go through it.  *out_len receives the size; ORCB_INVALID_ARGUMENT when out_cap is
 * too small. */
int orcb_host_decompress_section(int compression_kind, const uint8_t* in, size_t in_len, size_t block_size, uint8_t* out,
                                 size_t out_cap, size_t* out_len);

/* Host-only: the writer-zone table the device searches when a TIMESTAMP column was written in a zone other than UTC
 * (src/array_decoder/timestamp.rs:128-147, 242-286; the reference asks chrono-tz): `at[i]` = UTC instant of transition i
 * (ascending), `off[i]` = UTC offset in seconds east in force from at[i] on, *first_off = the offset before at[0].
 * *n = number of transitions (also when it exceeds cap); *orc_epoch = 2015-01-01 00:00:00 on the zone's wall clock as
 * seconds since the UNIX epoch (the base of the column's DATA stream).  ORCB_NOT_IMPLEMENTED when the host has no TZif
 * file for `name`. */
int orcb_zone_table(const char* name, int64_t* at, int32_t* off, size_t cap, size_t* n, int32_t* first_off, int64_t* orc_epoch);

/* Jobs of this process that were decoded a second time without the row index because a (stream, row group) segment did
 * not end where the index says the next one starts (damaged stream or index).  0 on well-formed files. */
uint64_t orcb_index_retries(void);
/* Jobs of this process that were planned and decoded a second time because a Zlib / LZ4 / LZO / unsized Zstandard chunk
 * in the middle of a stream did not fill its block (the writers of the reference's fixtures do fill them in 32 files out
 * of 33; the sizes found are remembered per open file, so a file pays at most once per group of stripes). */
uint64_t orcb_layout_retries(void);
/* Error detail of the last failing call on this thread. */
const char* orcb_last_error(void);
/* Build identification: "sm_100a" etc. */
const char* orcb_build_info(void);
/* 1 if a CUDA device is usable in this process. */
int orcb_device_available(void);

#ifdef __cplusplus
}
#endif
#endif /* ORC_B200_H */
