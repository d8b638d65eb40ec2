// Decode job: host planner + device execution for a list of (file, stripe) tasks.
#pragma once
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <array>
#include <map>
#include <vector>

#include "common.h"
#include "dev.h"
#include "meta.h"

namespace orcb {

struct ReadOptions {
    int device = 0;
    uint32_t batch_size = 8192;
    bool project_all = true;
    std::vector<std::string> projection;
    uint64_t range_start = 0, range_end = 0;
    int timestamp_unit = 0;  // 0 ns, 1 us, 2 ms, 3 s
    bool use_row_index = true;
    bool device_resident = false;
    uint32_t max_stripes_per_launch = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    uint32_t shard_index = 0, shard_count = 1;
    uint32_t waves = 0;      // bulk jobs: number of stripe waves in flight (0 = automatic)
    // with_schema (src/arrow_reader.rs:80-83): per projected column, in output order; -1 = the default mapping,
    // 0..3 = timestamp unit ns/us/ms/s, 4 = Decimal128(38, 9) nanoseconds (array_decoder/timestamp.rs:150-190)
    std::vector<int> ts_hint;
    static ReadOptions from_c(const OrcbReadOptions* o);
};

// projected root column after schema mapping (src/schema.rs:503-577)
struct OutColumn {
    std::string name;
    uint32_t col_id;
    int kind;            // ORC TypeKind
    uint32_t precision = 0, scale = 0;
    std::string format;  // Arrow C format string
    uint32_t width = 0;  // bytes per value (0 for bool / strings)
    int ts_unit = 0;     // timestamps: 0 ns, 1 us, 2 ms, 3 s
    bool ts_decimal = false;  // timestamps read as Decimal128(38, 9)
    std::vector<uint32_t> child_ids;  // struct / list / map / union: the children's column ids
};

std::vector<OutColumn> project_columns(const FileMeta& fm, const ReadOptions& opt);
OutColumn column_info(const FileMeta& fm, uint32_t col_id, const std::string& name, const ReadOptions& opt, int hint);
bool has_nested_columns(const std::vector<OutColumn>& cols);
void export_schema(const FileMeta& fm, const std::vector<OutColumn>& cols, const ReadOptions& opt, ArrowSchema* out);
// with_schema: checks the caller's Arrow schema against the file (array_decoder_factory, src/array_decoder/mod.rs:390-511)
// and records the timestamp variants it asks for in opt.ts_hint
void apply_schema_hints(const FileMeta& fm, ReadOptions& opt, const ArrowSchema* schema);

// A column decoded by a job over `n_slots` slots: a projected root column over the stripe's rows, or - in the jobs
// that follow for nested types - a child column over the slots its parent gives it.
struct RootSpec {
    uint32_t col_id = 0;
    uint64_t n_slots = 0;
    bool has_parent = false;    // decodes under a parent's validity (struct and union children)
    uint64_t parent_bits = 0;   // device address of that validity bitmap over the slots, LSB first
    uint64_t parent_count = 0;  // set bits in it = entries of this column's own PRESENT stream (merge_parent_present)
};

struct StripeTask {
    const FileMeta* file;
    uint32_t stripe;
    std::vector<RootSpec> roots;  // empty: the projected root columns over the stripe's rows
    // a later nesting level reads the stripe bytes its parent level staged (absolute device address of the data area)
    bool has_staged = false;
    uint64_t staged_abs = 0;
    // row selection (src/array_decoder/mod.rs:313-364): the batches of this stripe are these row ranges, in order,
    // instead of consecutive batch_size slices.  The stripe is decoded once; the ranges are exported as views.
    bool has_views = false;
    std::vector<std::pair<uint32_t, uint32_t>> views;  // (first row, rows), rows counted from the start of the stripe
    // partial decode: only row groups [g_begin, g_end) of the stripe are decoded (the views lie inside them).  Needs
    // the row index; a column whose index entries are unusable is decoded whole instead.
    bool has_window = false;
    uint32_t g_begin = 0, g_end = 0;
};

// RowSelector (src/row_selection.rs:32-57)
struct RowSelector {
    uint64_t row_count;
    bool skip;
};
// The batches each stripe yields under a row selection: restates RowSelection::from(Vec) (:466-482), split_off
// (:278-320), ArrowReader::try_advance_stripe (src/arrow_reader.rs:296-309) and next_with_row_selection.
// out[i].first = false: the stripe is read without a selection (the selection was used up before it).
// `predicate`: per stripe, the selection with_predicate derived (predicate.h); combined as `mine.and_then(predicate's)`.
std::vector<std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>>> selection_views(
    std::vector<RowSelector> selectors, const std::vector<uint64_t>& stripe_rows, uint64_t batch_size,
    const std::vector<std::vector<RowSelector>>* predicate = nullptr, bool has_selection = true);

// The same, one stripe at a time: with_predicate evaluates a stripe's row-group statistics only when the reader gets
// to that stripe (ArrowReader::try_advance_stripe, src/arrow_reader.rs:256-309), so a reader that is dropped early - or a
// ChunkReader behind slow storage - never touches the index areas of the stripes it does not reach.
class SelectionCursor {
public:
    SelectionCursor() = default;
    SelectionCursor(std::vector<RowSelector> selectors, bool has_selection);
    std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>> next_stripe(uint64_t rows, uint64_t batch_size,
                                                                            const std::vector<RowSelector>* predicate);

private:
    std::vector<RowSelector> sel_;
    bool has_selection_ = false;
};

// where one column of one stripe lands
struct ColStripePlan {
    uint32_t row_base = 0;   // stripe row that row 0 of this column's buffers holds (> 0: windowed partial decode)
    uint32_t task = 0, col = 0;       // indices into tasks / out columns
    uint32_t n_rows = 0, n_batches = 0;
    bool has_present = false;
    uint32_t nulls_idx = 0;           // first entry in nulls[] (has_present)
    uint64_t validity = 0;            // AR_OUT offset, batch b at + b * validity_stride
    uint32_t validity_stride = 0;
    uint64_t values = 0;              // AR_OUT offset (prims: n_rows*width; bool: per-batch bitmaps)
    uint32_t values_stride = 0;       // bool only
    uint64_t offsets = 0;             // AR_OUT offset, strings
    uint64_t str_data = 0;            // direct strings: arena-tagged start of the DATA bytes the rows use (no copy is made)
    uint64_t str_data_len = 0;        //   and their number; dictionary: the heap pointer comes from ptr_table
    uint64_t str_host_off = 0;        //   offset of their host copy in HostOutput::strs
    int32_t str_slot = -1;            // index into StrCol table / ptr_table
    uint64_t batch_base_off = 0;      // byte offset inside the meta blob of i64[n_batches+1]
    // nested types (struct / list / map / union): the children are decoded by the next-level job
    bool is_list = false;             // list / map: `offsets` are element offsets, batch_base[1] the children's slot count
    std::vector<uint32_t> kids;       // children's column ids
    uint64_t valid_bits = 0;          // arena-tagged stripe-level validity bitmap of this column (0 = none)
    uint64_t union_bits = 0;          // union: arena-tagged child validity bitmaps, union_stride bytes apart
    uint32_t union_stride = 0, union_counts = 0;  // nulls[union_counts + i] = valid slots of child i
};

struct HostOutput;  // D2H copy of the output arenas, shared by exported batches

// Thrown by Job::finish when chunks of unknown decompressed size (Zlib, LZ4) turned out not to fill their blocks: the
// sizes found are in the files' ChunkSizeCache by then, and the same tasks planned again get an exact layout.
struct LayoutRetry {};
// Thrown by Job::finish when a (stream, row group) segment did not stop where the row index says the next one starts
// (damaged stream or index): the same tasks are decoded again without the row index, as the reference decodes.
struct IndexRetry {};

class Job {
  public:
    Job(std::vector<StripeTask> tasks, const ReadOptions& opt);
    // a fresh job over the same tasks and options (after LayoutRetry)
    std::unique_ptr<Job> rebuild(bool without_row_index = false) const {
        ReadOptions o = orig_opt_;
        if (without_row_index) o.use_row_index = false;
        return std::make_unique<Job>(tasks_, o);
    }
    ~Job();
    void plan();
    void stage();
    void launch();
    void finish();
    void restage();
    void stats(OrcbJobStats* out) const;
    uint32_t kernel_stats(OrcbKernelStat* out, uint32_t cap) const;
    uint64_t num_batches() const { return batch_task_.size(); }
    // nested columns: the roots the next-level job decodes for task `t` (valid after finish())
    std::vector<RootSpec> next_roots(uint32_t t) const;
    cudaStream_t stream() const { return stream_; }
    void export_batch(uint64_t i, ArrowArray* out);
    void export_batch_device(uint64_t i, ArrowDeviceArray* out);
    const std::vector<OutColumn>& columns() const { return cols_; }

  private:
    void run_next_level();
    // one column (any depth) of batch rows [slot0, slot0 + n): `top` = a root column of the batch (carries the view offset)
    void export_node(uint32_t cs, int64_t slot0, int64_t n, bool top, bool device, ArrowArray* a);
    uint64_t alloc(Arena a, uint64_t bytes, uint64_t align = 256);
    uint64_t reloc(uint64_t tagged) const;
    void plan_stripe(uint32_t task_idx);
    void ensure_host_output();

    std::vector<StripeTask> tasks_;
    ReadOptions opt_, orig_opt_;
    std::vector<OutColumn> cols_;
    bool planned_ = false, staged_ = false, launched_ = false, finished_ = false;

    // arena sizes (bytes) and device bases
    uint64_t size_[8] = {0};
    uint8_t* base_[8] = {nullptr};

    // descriptor tables (host copies; pointer fields relocated at stage())
    std::vector<Seg> present_byte_segs_, data_byte_segs_, int_segs_, int_big_segs_, var_segs_;
    std::vector<BitSeg> present_bit_segs_, data_bit_segs_;
    std::vector<ScanDesc> scans_;
    std::vector<CopyDesc> copies_;
    std::vector<uint2> copy_tiles_;
    std::vector<uint8_t> tz_blob_;                  // zone tables: i64 instants then i32 offsets, 16-byte aligned pieces
    std::map<std::string, std::array<uint64_t, 4>> tz_tables_;  // zone -> {offset of instants, offset of offsets, n, first}
    std::map<std::pair<const void*, uint32_t>, uint64_t> staged_stripes_;  // (file, stripe) -> offset of its data in the IN arena
    std::vector<uint2> u8_tiles_;   // (string column, U8_TILE-byte tile) units of the UTF-8 check
    std::vector<SpacedDesc> spaced_, spaced_late_, merge_spaced_;
    std::vector<PopcDesc> popcs_;
    std::vector<uint2> chk_pairs_;         // (slot, slot of the next segment of the same stream)
    uint32_t n_chk_ = 0;
    uint64_t chk_table_ = 0, o_chkpair_ = 0;  // AR_TMP offset of SegCheck[n_chk_]
    std::vector<UnionDesc> unions_;
    bool nested_ = false;                  // some column is (or descends from) a struct / list / map / union
    std::unique_ptr<Job> next_level_;      // decodes the children of this job's nested columns
    std::vector<DecFixDesc> decfix_;
    std::vector<TsDesc> ts_;
    std::vector<StrCol> strcols_;
    std::vector<RepackDesc> repacks_;
    std::vector<ChunkDesc> chunks_;
    uint32_t str_tiles_ = 0, repack_work_ = 0, pool_blocks_ = 0;  // capacity of the run-block pool
    uint64_t run_table_ = 0, block_recs_ = 0, slow_list_ = 0, coop_q_ = 0;
    uint32_t coop_cap_ = 0;
    uint64_t small_values_ = 0;  // values decoded by the short-run path (sizes the whole-warp run queue)  // AR_TMP offsets: RunRec table (32 per block), BlockRec table

    // stage copies: (file ptr, file offset, AR_IN offset, bytes)
    struct StageCopy {
        const uint8_t* src;
        uint64_t dst_off, bytes;
    };
    std::vector<StageCopy> stage_copies_;
    std::vector<uint64_t> task_in_off_;  // per task: tagged address of the stripe's staged data area
    std::vector<std::shared_ptr<RangeBuf>> range_keep_;  // stripes read through callbacks, held while this job stages from them

    // device blobs
    uint8_t* d_desc_ = nullptr;
    uint64_t desc_bytes_ = 0;
    // offsets of each table inside the descriptor blob
    uint64_t o_pbyte_ = 0, o_dbyte_ = 0, o_int_ = 0, o_intbig_ = 0, o_var_ = 0, o_pbit_ = 0, o_dbit_ = 0, o_scan_ = 0, o_copy_ = 0,
             o_ctile_ = 0, o_sp_ = 0, o_sp2_ = 0, o_spm_ = 0, o_popc_ = 0, o_union_ = 0, o_dec_ = 0, o_ts_ = 0, o_str_ = 0, o_rep_ = 0, o_chunk_ = 0, o_u8tile_ = 0, o_tz_ = 0;
    std::vector<uint8_t> desc_blob_;

    // state blob (device only, zeroed per launch): cnt[], dstart[], mis[], JobState
    uint32_t n_cnt_ = 0, n_colstripes_ = 0;
    uint8_t* d_state_ = nullptr;
    uint32_t n_str_int_segs_ = 0;  // leading int_segs_ that feed the string kernels (string lengths, dictionary keys)
    uint32_t bits_codecs_ = 0;  // 1 << codec over the serial-chain chunks
    uint32_t n_bits_chunks_ = 0, n_snappy_chunks_ = 0;  // chunks_ is ordered: serial-chain codecs, Snappy, then LZ4 / stored
    uint64_t state_bytes_ = 0, o_cnt_ = 0, o_dstart_ = 0, o_mis_ = 0, o_jobstate_ = 0, o_nblocks_ = 0;
    // meta blob (device, zeroed per launch, copied to host at finish): err[], nulls[], ptr_table[], batch_base[]
    uint8_t* d_meta_ = nullptr;
    uint8_t* h_meta_ = nullptr;  // pinned, from the process-wide cache of small pinned buffers (job.cc)
    size_t h_meta_cap_ = 0;
    uint64_t meta_bytes_ = 0, o_err_ = 0, o_nulls_ = 0, o_ptrs_ = 0, o_bbase_ = 0, o_clens_ = 0, o_retry_ = 0;
    std::vector<std::pair<const FileMeta*, uint64_t>> chunk_keys_;  // chunk id -> (file, offset of the chunk header)
    uint32_t n_nulls_ = 0;

    std::vector<ColStripePlan> colstripes_;
    // batch i -> (task, batch-in-stripe)
    std::vector<uint32_t> batch_task_, batch_idx_;
    std::vector<uint32_t> batch_row0_, batch_rows_;  // views (row selection): first row inside the stripe / rows
    uint32_t user_batch_size_ = 8192;
    bool view_mode_ = false;                          // every stripe is one internal batch, user batches are views
    std::vector<uint32_t> task_first_cs_;

    // stats
    uint64_t input_bytes_ = 0, n_rows_ = 0, n_segments_ = 0, n_launches_ = 0, output_bytes_ = 0, aliased_bytes_ = 0;

    struct KStat {
        std::string name;
        uint64_t alg_bytes = 0, work = 0;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        double ms = 0;
        bool ran = false;
    };
    std::vector<KStat> kstats_;
    KStat& kstat(const char* name);
    // algorithmic byte counters accumulated by the planner (per kernel)
    uint64_t ab_utf8_ = 0, ab_decomp_ = 0, ab_present_ = 0, ab_byte_ = 0, ab_bits_ = 0, ab_int_ = 0, ab_intbig_ = 0, ab_var_ = 0, ab_copy_ = 0,
             ab_spaced_ = 0, ab_dec_ = 0, ab_ts_ = 0, ab_str_ = 0, ab_repack_ = 0;

    cudaStream_t aux_stream_ = nullptr;  // latency-bound pre-pass + short-run integer decode overlap the rest
    cudaEvent_t ev_fork_ = nullptr, ev_join_ = nullptr, ev_str_ = nullptr;
    cudaStream_t stream_ = nullptr;
    bool own_stream_ = false;
    cudaEvent_t done_ = nullptr;
    std::shared_ptr<HostOutput> host_out_;
    std::shared_ptr<void> dev_keepalive_;
    friend struct DeviceArenas;
};

}  // namespace orcb
