// Host planner + executor.  The planner restates, as table construction, what the reference does
// stream by stream at run time:
//   Stripe::new stream offsets            src/stripe.rs:128-182
//   array_decoder_factory type dispatch   src/array_decoder/mod.rs:390-511
//   Column::rle_version / dictionary_size src/column.rs:40-59
//   new_timestamp_decoder base selection  src/array_decoder/timestamp.rs:128-147
//   NaiveStripeDecoder batch boundaries   src/array_decoder/mod.rs:371-387, 514-564
// and adds what the reference never uses: row-index positions (src/row_index.rs:37-51) as parallel
// entry points, one segment per (stream, row group).
#include "job.h"
#include "tz.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace orcb {

#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) fail(ORCB_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

static const int64_t ORC_EPOCH_UTC = 1420070400;  // array_decoder/timestamp.rs:51
static const uint64_t ARENA_PAD = 256;            // slack so word-granular kernel loads may overrun streams

ReadOptions ReadOptions::from_c(const OrcbReadOptions* o) {
    ReadOptions r;
    if (!o) return r;
    r.device = o->device;
    r.batch_size = o->batch_size ? o->batch_size : 8192;
    if (o->projection_names) {
        r.project_all = false;
        for (uint32_t i = 0; i < o->n_projection; i++) r.projection.emplace_back(o->projection_names[i]);
    }
    r.range_start = o->range_start;
    r.range_end = o->range_end;
    r.timestamp_unit = o->timestamp_unit;
    r.use_row_index = !(o->flags & 1u);
    r.device_resident = o->device_resident != 0;
    r.max_stripes_per_launch = o->max_stripes_per_launch;
    r.stream = (cudaStream_t)o->cuda_stream;
    r.own_stream = o->cuda_stream == nullptr;
    r.shard_index = o->stripe_shard_index;
    r.shard_count = o->stripe_shard_count ? o->stripe_shard_count : 1;
    return r;
}

// ------------------------------------------------------------------------------------------------
// schema (src/schema.rs:503-577, src/arrow_reader.rs:182-198)
// ------------------------------------------------------------------------------------------------
std::vector<OutColumn> project_columns(const FileMeta& fm, const ReadOptions& opt) {
    std::vector<OutColumn> out;
    static const char* ts_fmt[4] = {"tsn:", "tsu:", "tsm:", "tss:"};
    if (opt.timestamp_unit < 0 || opt.timestamp_unit > 3) fail(ORCB_INVALID_ARGUMENT, "timestamp_unit must be 0..3");
    for (auto& rc : fm.root_columns) {
        if (!opt.project_all &&
            std::find(opt.projection.begin(), opt.projection.end(), rc.first) == opt.projection.end())
            continue;
        const OrcType& t = fm.types[rc.second];
        OutColumn c;
        c.name = rc.first;
        c.col_id = rc.second;
        c.kind = t.kind;
        switch (t.kind) {
            case T_BOOLEAN: c.format = "b"; c.width = 0; break;
            case T_BYTE: c.format = "c"; c.width = 1; break;
            case T_SHORT: c.format = "s"; c.width = 2; break;
            case T_INT: c.format = "i"; c.width = 4; break;
            case T_LONG: c.format = "l"; c.width = 8; break;
            case T_FLOAT: c.format = "f"; c.width = 4; break;
            case T_DOUBLE: c.format = "g"; c.width = 8; break;
            case T_STRING: case T_VARCHAR: case T_CHAR: c.format = "u"; break;
            case T_BINARY: c.format = "z"; break;
            case T_DATE: c.format = "tdD"; c.width = 4; break;
            case T_DECIMAL:
                c.precision = t.precision;
                c.scale = t.scale;
                // arrow validates Decimal128 precision/scale (array_decoder/decimal.rs:99)
                if (t.precision == 0 || t.precision > 38 || t.scale > t.precision)
                    fail(ORCB_ARROW, "invalid Decimal128 precision/scale for column " + c.name);
                c.format = "d:" + std::to_string(t.precision) + "," + std::to_string(t.scale);
                c.width = 16;
                break;
            case T_TIMESTAMP: case T_TIMESTAMP_INSTANT: {
                const int hint = out.size() < opt.ts_hint.size() ? opt.ts_hint[out.size()] : -1;
                c.ts_unit = hint >= 0 && hint <= 3 ? hint : opt.timestamp_unit;
                c.ts_decimal = hint == 4;
                if (c.ts_decimal) {
                    c.format = "d:38,9";
                    c.width = 16;
                    c.ts_unit = 0;
                } else {
                    c.format = std::string(ts_fmt[c.ts_unit]) + (t.kind == T_TIMESTAMP_INSTANT ? "UTC" : "");
                    c.width = 8;
                }
                break;
            }
            default:
                fail(ORCB_NOT_IMPLEMENTED, "nested ORC types (struct/list/map/union) are not on the device path yet: column " + c.name);
        }
        out.push_back(std::move(c));
    }
    return out;
}

void apply_schema_hints(const FileMeta& fm, ReadOptions& opt, const ArrowSchema* schema) {
    if (!schema) return;
    opt.ts_hint.clear();
    const std::vector<OutColumn> cols = project_columns(fm, opt);
    if (!schema->format || strcmp(schema->format, "+s") != 0) fail(ORCB_INVALID_ARGUMENT, "with_schema: the schema must be a struct");
    if ((size_t)schema->n_children != cols.size())
        fail(ORCB_MISMATCHED_SCHEMA, "with_schema: " + std::to_string(schema->n_children) + " fields for " +
                                         std::to_string(cols.size()) + " projected columns");
    std::vector<int> hints(cols.size(), -1);
    for (size_t i = 0; i < cols.size(); i++) {
        const std::string fmt = schema->children[i]->format ? schema->children[i]->format : "";
        const OutColumn& c = cols[i];
        auto mismatch = [&]() {
            fail(ORCB_MISMATCHED_SCHEMA, "column '" + c.name + "' (ORC type kind " + std::to_string(c.kind) + ") cannot be read as Arrow '" + fmt + "'");
        };
        if (c.kind == T_TIMESTAMP || c.kind == T_TIMESTAMP_INSTANT) {
            if (fmt == "d:38,9") { hints[i] = 4; continue; }
            static const char units[4] = {'n', 'u', 'm', 's'};
            int unit = -1;
            if (fmt.size() >= 4 && fmt[0] == 't' && fmt[1] == 's' && fmt[3] == ':')
                for (int u = 0; u < 4; u++)
                    if (fmt[2] == units[u]) unit = u;
            if (unit < 0) mismatch();
            const std::string tz = fmt.substr(4);
            if (c.kind == T_TIMESTAMP) {
                if (!tz.empty()) mismatch();  // new_timestamp_decoder only takes Timestamp(_, None)
            } else {
                if (tz.empty()) mismatch();
                if (tz != "UTC") fail(ORCB_UNSUPPORTED_TYPE_VARIANT, "Non-UTC Arrow timestamps");  // timestamp.rs:214-217
            }
            hints[i] = unit;
        } else if (fmt != c.format) {
            mismatch();
        }
    }
    opt.ts_hint = hints;
}

namespace {
struct SchemaPriv {
    std::string format, name, metadata;
    std::vector<ArrowSchema> child_store;
    std::vector<ArrowSchema*> child_ptrs;
};
void release_schema(ArrowSchema* s) {
    if (!s || !s->release) return;
    for (int64_t i = 0; i < s->n_children; i++)
        if (s->children[i]->release) s->children[i]->release(s->children[i]);
    delete (SchemaPriv*)s->private_data;
    s->release = nullptr;
}
void fill_schema(ArrowSchema* s, const std::string& fmt, const std::string& name, int64_t flags) {
    auto* p = new SchemaPriv();
    p->format = fmt;
    p->name = name;
    s->format = p->format.c_str();
    s->name = p->name.c_str();
    s->metadata = nullptr;
    s->flags = flags;
    s->n_children = 0;
    s->children = nullptr;
    s->dictionary = nullptr;
    s->release = release_schema;
    s->private_data = p;
}
}  // namespace

void export_schema(const FileMeta& fm, const std::vector<OutColumn>& cols, ArrowSchema* out) {
    fill_schema(out, "+s", "", 0);
    auto* p = (SchemaPriv*)out->private_data;
    if (!fm.user_metadata.empty()) {
        std::string& m = p->metadata;
        auto put32 = [&](int32_t v) { m.append((const char*)&v, 4); };
        put32((int32_t)fm.user_metadata.size());
        for (auto& kv : fm.user_metadata) {
            put32((int32_t)kv.first.size());
            m.append(kv.first);
            put32((int32_t)kv.second.size());
            m.append(kv.second);
        }
        out->metadata = p->metadata.data();
    }
    p->child_store.resize(cols.size());
    p->child_ptrs.resize(cols.size());
    for (size_t i = 0; i < cols.size(); i++) {
        fill_schema(&p->child_store[i], cols[i].format, cols[i].name, 2 /* ARROW_FLAG_NULLABLE: src/schema.rs:131 */);
        p->child_ptrs[i] = &p->child_store[i];
    }
    out->n_children = (int64_t)cols.size();
    out->children = p->child_ptrs.data();
}

// ------------------------------------------------------------------------------------------------
// device memory owned jointly by the job and by exported device batches
// ------------------------------------------------------------------------------------------------
struct DeviceArenas {
    int device = 0;
    std::vector<void*> ptrs;
    ~DeviceArenas() {
        for (void* p : ptrs)
            if (p) cudaFree(p);
    }
};

struct HostOutput {
    uint8_t* out = nullptr;   // pinned copy of AR_OUT
    uint8_t* heap = nullptr;  // pinned copy of the used part of AR_HEAP
    ~HostOutput() {
        if (out) cudaFreeHost(out);
        if (heap) cudaFreeHost(heap);
    }
};

Job::Job(std::vector<StripeTask> tasks, const ReadOptions& opt) : tasks_(std::move(tasks)), opt_(opt) {
    if (!tasks_.empty()) cols_ = project_columns(*tasks_[0].file, opt_);
}

Job::KStat& Job::kstat(const char* name) {
    for (auto& k : kstats_)
        if (k.name == name) return k;
    kstats_.emplace_back();
    KStat& k = kstats_.back();
    k.name = name;
    cudaEventCreate(&k.e0);
    cudaEventCreate(&k.e1);
    return k;
}

uint32_t Job::kernel_stats(OrcbKernelStat* out, uint32_t cap) const {
    uint32_t n = 0;
    for (auto& k : kstats_) {
        if (!k.ran || n >= cap) continue;
        memset(&out[n], 0, sizeof(out[n]));
        strncpy(out[n].name, k.name.c_str(), sizeof(out[n].name) - 1);
        out[n].ms = k.ms;
        out[n].alg_bytes = k.alg_bytes;
        out[n].work_items = k.work;
        n++;
    }
    return n;
}

void Job::restage() {
    if (!staged_) { stage(); return; }
    CUDA_OK(cudaSetDevice(opt_.device));
    CUDA_OK(cudaMemcpyAsync(d_desc_, desc_blob_.data(), desc_blob_.size(), cudaMemcpyHostToDevice, stream_));
    for (auto& sc : stage_copies_)
        CUDA_OK(cudaMemcpyAsync(base_[AR_IN] + sc.dst_off, sc.src, sc.bytes, cudaMemcpyHostToDevice, stream_));
}

Job::~Job() {
    for (auto& k : kstats_) {
        if (k.e0) cudaEventDestroy(k.e0);
        if (k.e1) cudaEventDestroy(k.e1);
    }
    if (h_meta_) cudaFreeHost(h_meta_);
    if (done_) cudaEventDestroy(done_);
    if (ev_fork_) cudaEventDestroy(ev_fork_);
    if (ev_join_) cudaEventDestroy(ev_join_);
    if (aux_stream_) cudaStreamDestroy(aux_stream_);
    if (own_stream_ && stream_) cudaStreamDestroy(stream_);
    // device arenas are released by dev_keepalive_ (shared with exported device batches)
}

uint64_t Job::alloc(Arena a, uint64_t bytes, uint64_t align) {
    uint64_t off = align_up(size_[a], align);
    size_[a] = off + bytes;
    return aref(a, off);
}

uint64_t Job::reloc(uint64_t tagged) const {
    const uint64_t a = tagged >> 60;
    if (a == AR_NULL) return 0;
    return (uint64_t)(uintptr_t)base_[a] + (tagged & ((1ull << 60) - 1));
}

// ------------------------------------------------------------------------------------------------
// planning
// ------------------------------------------------------------------------------------------------
namespace {

enum PosKind { PK_RAW = 0, PK_INT_RLE = 1, PK_BYTE_RLE = 1, PK_BOOL = 2 };  // extra positions after the byte offset

// A stream as the kernels will see it: a contiguous (possibly still to be decompressed) byte range.
struct StreamRef {
    bool present = false;
    uint64_t ptr = 0;   // arena-tagged
    uint32_t len = 0;   // bytes available (exact, or an upper bound for LZ4 tails)
    std::vector<ChunkInfo> chunks;     // compressed files
    std::vector<uint64_t> chunk_dst;   // decompressed offset of each chunk
};

struct Entry {
    uint32_t byte = 0, skip = 0, bit = 0;
};

bool is_utc_zone(const std::string& z) {
    static const char* names[] = {"UTC", "GMT", "Etc/UTC", "Etc/GMT", "Etc/UCT", "UCT", "Etc/Universal", "Universal",
                                  "Etc/Zulu", "Zulu", "Etc/GMT+0", "Etc/GMT-0", "Etc/GMT0", "GMT+0", "GMT-0", "GMT0",
                                  "Etc/Greenwich", "Greenwich"};
    for (auto n : names)
        if (z == n) return true;
    return false;
}

// Scheduling hint only: do the first few RLE v2 runs at `pos` all hold more than 64 values?  (Header walk,
// no values decoded.)  Such segments go to the warp-per-segment kernel.
bool rle2_opens_with_long_runs(const uint8_t* s, uint32_t len, uint32_t pos) {
    static const int W[32] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 26, 28, 30, 32, 40, 48, 56, 64};
    for (int r = 0; r < 8; r++) {
        if (pos + 4 > len) return r > 0;
        const uint32_t h = s[pos], kind = h >> 6;
        if (kind == 0) return false;
        const uint32_t rl = (((h & 1) << 8) | s[pos + 1]) + 1;
        if (rl <= 64) return false;
        const uint32_t code = (h >> 1) & 31;
        if (kind == 1) {
            pos += 2 + (rl * W[code] + 7) / 8;
        } else if (kind == 2) {
            const uint32_t b3 = s[pos + 2], b4 = s[pos + 3];
            const int pw = W[b3 & 31], pgw = ((b4 >> 5) & 7) + 1;
            const int t = pw + pgw;
            const int cfb = t <= 24 ? t : t <= 26 ? 26 : t <= 28 ? 28 : t <= 30 ? 30 : t <= 32 ? 32 : (t + 7) / 8 * 8;
            pos += 4 + ((b3 >> 5) & 7) + 1 + (rl * W[code] + 7) / 8 + ((b4 & 31) * cfb + 7) / 8;
        } else {
            uint32_t p = pos + 2;
            for (int v = 0; v < 2; v++) {
                while (p < len && (s[p] & 0x80)) p++;
                p++;
            }
            if (code) p += ((rl - 2) * W[code] + 7) / 8;
            pos = p;
        }
    }
    return true;
}

}  // namespace

void Job::plan() {
    if (planned_) return;
    if (opt_.batch_size == 0) fail(ORCB_INVALID_ARGUMENT, "batch_size must be > 0");
    task_first_cs_.clear();
    staged_stripes_.clear();
    for (auto& t : tasks_) view_mode_ |= t.has_views;
    user_batch_size_ = opt_.batch_size;
    if (view_mode_) {
        // row selection: one internal batch per stripe (stripe-level offsets and bitmaps); the user's batches are
        // views into it, exported through the Arrow `offset` field
        uint64_t mx = 1;
        for (auto& t : tasks_) mx = std::max<uint64_t>(mx, t.file->stripes[t.stripe].rows);
        opt_.batch_size = (uint32_t)std::min<uint64_t>(mx, 0xfffffff0ull);
    }
    for (uint32_t t = 0; t < tasks_.size(); t++) plan_stripe(t);

    if (pool_blocks_) {
        run_table_ = alloc(AR_TMP, (uint64_t)pool_blocks_ * 32 * sizeof(RunRec));
        block_recs_ = alloc(AR_TMP, (uint64_t)pool_blocks_ * sizeof(BlockRec));
        slow_list_ = alloc(AR_TMP, (uint64_t)pool_blocks_ * 4 + 16);
        // queue of whole-warp runs found by the pre-pass; on overflow the run simply stays in its block
        coop_cap_ = (uint32_t)std::min<uint64_t>(small_values_ / 64 + 65536, 0x7fffffffu);
        coop_q_ = alloc(AR_TMP, (uint64_t)coop_cap_ * sizeof(CoopRec));
    }

    // ---- descriptor blob layout
    auto place = [&](uint64_t& off, size_t bytes) {
        off = align_up(desc_bytes_, 256);
        desc_bytes_ = off + bytes;
    };
    place(o_pbyte_, present_byte_segs_.size() * sizeof(Seg));
    place(o_dbyte_, data_byte_segs_.size() * sizeof(Seg));
    place(o_int_, int_segs_.size() * sizeof(Seg));
    place(o_intbig_, int_big_segs_.size() * sizeof(Seg));
    place(o_var_, var_segs_.size() * sizeof(Seg));
    place(o_pbit_, present_bit_segs_.size() * sizeof(BitSeg));
    place(o_dbit_, data_bit_segs_.size() * sizeof(BitSeg));
    place(o_scan_, scans_.size() * sizeof(ScanDesc));
    place(o_copy_, copies_.size() * sizeof(CopyDesc));
    place(o_ctile_, copy_tiles_.size() * sizeof(uint2));
    place(o_u8tile_, u8_tiles_.size() * sizeof(uint2));
    place(o_sp_, spaced_.size() * sizeof(SpacedDesc));
    place(o_sp2_, spaced_late_.size() * sizeof(SpacedDesc));
    place(o_dec_, decfix_.size() * sizeof(DecFixDesc));
    place(o_ts_, ts_.size() * sizeof(TsDesc));
    place(o_str_, strcols_.size() * sizeof(StrCol));
    place(o_rep_, repacks_.size() * sizeof(RepackDesc));
    place(o_chunk_, chunks_.size() * sizeof(ChunkDesc));
    place(o_tz_, tz_blob_.size());
    desc_bytes_ = align_up(desc_bytes_, 256);

    // ---- state blob
    n_colstripes_ = (uint32_t)colstripes_.size();
    state_bytes_ = 0;
    auto splace = [&](uint64_t& off, size_t bytes) {
        off = align_up(state_bytes_, 256);
        state_bytes_ = off + bytes;
    };
    splace(o_cnt_, (size_t)(n_cnt_ + 1) * 4);
    splace(o_dstart_, (size_t)(n_cnt_ + 1) * 4);
    splace(o_mis_, (size_t)(n_colstripes_ + 1) * 4);
    splace(o_jobstate_, sizeof(JobState));
    splace(o_nblocks_, 16);
    state_bytes_ = align_up(state_bytes_, 256);

    // ---- meta blob: err | nulls | ptr table | batch bases (batch_base_off were assigned relative to o_bbase_)
    uint64_t bbase_bytes = meta_bytes_;  // accumulated by plan_stripe as running batch-base bytes
    meta_bytes_ = 0;
    auto mplace = [&](uint64_t& off, size_t bytes) {
        off = align_up(meta_bytes_, 256);
        meta_bytes_ = off + bytes;
    };
    mplace(o_err_, (size_t)(n_colstripes_ + 1) * 4);
    mplace(o_nulls_, (size_t)(n_nulls_ + 1) * 4);
    mplace(o_ptrs_, (strcols_.size() + 1) * 8);
    mplace(o_bbase_, bbase_bytes + 8);
    meta_bytes_ = align_up(meta_bytes_, 256);

    // heap capacity: dictionary string bytes are bump-allocated on the device
    size_[AR_HEAP] = align_up(size_[AR_HEAP], 256);
    for (int a = 1; a < 8; a++) size_[a] = align_up(size_[a] + ARENA_PAD, 256);

    // batches
    batch_task_.clear();
    batch_idx_.clear();
    batch_row0_.clear();
    batch_rows_.clear();
    for (uint32_t t = 0; t < tasks_.size(); t++) {
        const StripeInfo& si = tasks_[t].file->stripes[tasks_[t].stripe];
        if (view_mode_) {
            std::vector<std::pair<uint32_t, uint32_t>> whole;
            const auto* views = &tasks_[t].views;
            if (!tasks_[t].has_views) {  // stripe without a selection inside a job that has one: consecutive batches
                for (uint64_t r = 0; r < si.rows; r += user_batch_size_)
                    whole.emplace_back((uint32_t)r, (uint32_t)std::min<uint64_t>(user_batch_size_, si.rows - r));
                views = &whole;
            }
            for (auto& v : *views) {
                batch_task_.push_back(t);
                batch_idx_.push_back(0);
                batch_row0_.push_back(v.first);
                batch_rows_.push_back(v.second);
            }
            continue;
        }
        uint32_t nb = (uint32_t)((si.rows + opt_.batch_size - 1) / opt_.batch_size);
        for (uint32_t b = 0; b < nb; b++) {
            batch_task_.push_back(t);
            batch_idx_.push_back(b);
            batch_row0_.push_back(0);
            batch_rows_.push_back((uint32_t)std::min<uint64_t>(opt_.batch_size, si.rows - (uint64_t)b * opt_.batch_size));
        }
    }
    planned_ = true;
}

void Job::plan_stripe(uint32_t task_idx) {
    const FileMeta& fm = *tasks_[task_idx].file;
    const uint32_t stripe = tasks_[task_idx].stripe;
    const StripeInfo& si = fm.stripes[stripe];
    const StripeFooter sf = fm.read_stripe_footer(stripe);
    task_first_cs_.push_back((uint32_t)colstripes_.size());
    if (si.rows > 0xfffffff0ull) fail(ORCB_NOT_IMPLEMENTED, "stripes with more than 2^32 rows");
    const uint32_t stripe_rows = (uint32_t)si.rows;
    n_rows_ += stripe_rows;
    const uint32_t bs = opt_.batch_size;
    const bool compressed = fm.compression != C_NONE;
    const uint64_t data_start = si.offset + si.index_length;

    // stage the stripe's data area once (tasks that decode different row-group windows of one stripe share it)
    uint64_t in_off = 0;
    if (si.data_length) {
        auto key = std::make_pair((const void*)&fm, stripe);
        auto it = staged_stripes_.find(key);
        if (it != staged_stripes_.end()) {
            in_off = it->second;
        } else {
            in_off = alloc(AR_IN, si.data_length);
            stage_copies_.push_back({fm.data + data_start, in_off & ((1ull << 60) - 1), si.data_length});
            staged_stripes_[key] = in_off;
        }
    }

    // row-index stride usable for this stripe?
    uint32_t stride = stripe_rows ? stripe_rows : 1;
    bool want_index = opt_.use_row_index && fm.row_index_stride > 0 && stripe_rows > 0;
    if (want_index) stride = (uint32_t)std::min<uint64_t>((uint64_t)fm.row_index_stride, 0xffffffffull);
    const uint32_t idx_groups = stripe_rows ? (stripe_rows + stride - 1) / stride : 0;
    // partial decode (row selection): row groups [wg0, wg1) only
    const StripeTask& task = tasks_[task_idx];
    const bool windowed = task.has_window && want_index && idx_groups > 1 && task.g_begin < task.g_end && task.g_end <= idx_groups;
    const uint32_t wg0 = windowed ? task.g_begin : 0, wg1 = windowed ? task.g_end : idx_groups;

    for (uint32_t ci = 0; ci < cols_.size(); ci++) {
        const OutColumn& oc = cols_[ci];
        const uint32_t cid = oc.col_id;
        const uint32_t cs = (uint32_t)colstripes_.size();
        uint32_t n_rows = stripe_rows;  // rows this column decodes: the stripe's, or the window's once the index is known good
        uint32_t n_batches = (n_rows + bs - 1) / bs;
        ColStripePlan cp;
        cp.task = task_idx;
        cp.col = ci;
        cp.n_rows = n_rows;
        cp.n_batches = n_batches;
        if (n_rows == 0) {
            colstripes_.push_back(cp);
            continue;
        }
        ColumnEncoding enc;
        if (cid < sf.encodings.size()) enc = sf.encodings[cid];
        const bool v2 = enc.kind == E_DIRECT_V2 || enc.kind == E_DICTIONARY_V2;  // src/column.rs:52-59
        const bool dict_enc = enc.kind == E_DICTIONARY || enc.kind == E_DICTIONARY_V2;

        // ---- resolve streams (StreamMap::get: a missing stream decodes as empty, src/stripe.rs:319-326)
        auto resolve = [&](int kind) -> StreamRef {
            StreamRef r;
            const StreamInfo* st = sf.find(cid, kind);
            if (!st) return r;
            r.present = true;
            if (st->offset < data_start || st->offset + st->length > data_start + si.data_length)
                fail(ORCB_OUT_OF_SPEC, "data stream outside the stripe's data area");
            input_bytes_ += st->length;
            const uint64_t rel = st->offset - data_start;
            if (!compressed) {
                r.ptr = in_off + rel;
                if (st->length > 0xffffff00ull) fail(ORCB_NOT_IMPLEMENTED, "streams larger than 4 GiB");
                r.len = (uint32_t)st->length;
                return r;
            }
            r.chunks = fm.chunk_table(st->offset, st->length);
            uint64_t total = 0;
            for (size_t i = 0; i < r.chunks.size(); i++) {
                r.chunk_dst.push_back(total);
                const ChunkInfo& c = r.chunks[i];
                total += c.dst_len >= 0 ? (uint64_t)c.dst_len : fm.block_size;
            }
            if (total > 0xffffff00ull) fail(ORCB_NOT_IMPLEMENTED, "streams larger than 4 GiB");
            r.len = (uint32_t)total;
            r.ptr = alloc(AR_DEC, total + 64);
            for (size_t i = 0; i < r.chunks.size(); i++) {
                const ChunkInfo& c = r.chunks[i];
                ChunkDesc d{};
                d.src = in_off + rel + c.src_off;
                d.dst = r.ptr + r.chunk_dst[i];
                d.src_len = c.src_len;
                d.codec = c.original ? 0 : (uint8_t)fm.compression;
                d.colstripe = cs;
                if (c.dst_len >= 0) {
                    d.dst_cap = (uint32_t)c.dst_len;
                    d.expect_len = (int32_t)c.dst_len;
                } else {
                    d.dst_cap = (uint32_t)fm.block_size;
                    // layout assumes every non-final chunk fills the block; verified on the device
                    d.expect_len = i + 1 < r.chunks.size() ? (int32_t)fm.block_size : -1;
                }
                chunks_.push_back(d);
                ab_decomp_ += c.src_len + (uint64_t)d.dst_cap;
            }
            return r;
        };

        StreamRef s_present = resolve(S_PRESENT);
        StreamRef s_data = resolve(S_DATA);
        StreamRef s_length, s_secondary, s_dict;
        const int k = oc.kind;
        const bool is_str = k == T_STRING || k == T_VARCHAR || k == T_CHAR || k == T_BINARY;
        const bool use_dict = is_str && k != T_BINARY && dict_enc;  // string.rs:51-84 (binary is always direct)
        if (is_str) s_length = resolve(S_LENGTH);
        if (use_dict) s_dict = resolve(S_DICTIONARY_DATA);
        if (k == T_DECIMAL || k == T_TIMESTAMP || k == T_TIMESTAMP_INSTANT) s_secondary = resolve(S_SECONDARY);
        const bool has_present = s_present.present;
        cp.has_present = has_present;

        // ---- row-index entries -> per-stream entry points
        // stream order inside a RowIndexEntry.positions list follows the writers' (Java/C++) recording order
        struct PosSpec {
            StreamRef* sr;
            int extra;
        };
        std::vector<PosSpec> specs;
        if (has_present) specs.push_back({&s_present, 2});
        switch (k) {
            case T_BOOLEAN: specs.push_back({&s_data, 2}); break;
            case T_BYTE: specs.push_back({&s_data, 1}); break;
            case T_SHORT: case T_INT: case T_LONG: case T_DATE: specs.push_back({&s_data, 1}); break;
            case T_FLOAT: case T_DOUBLE: specs.push_back({&s_data, 0}); break;
            case T_STRING: case T_VARCHAR: case T_CHAR: case T_BINARY:
                if (use_dict) specs.push_back({&s_data, 1});
                else { specs.push_back({&s_data, 0}); specs.push_back({&s_length, 1}); }
                break;
            case T_DECIMAL: specs.push_back({&s_data, 0}); specs.push_back({&s_secondary, 1}); break;
            case T_TIMESTAMP: case T_TIMESTAMP_INSTANT:
                specs.push_back({&s_data, 1}); specs.push_back({&s_secondary, 1}); break;
            default: break;
        }
        uint32_t n_groups = 1;
        uint32_t gstride = n_rows;
        std::vector<std::vector<Entry>> entries(specs.size());  // [spec][group]
        bool indexed = false;
        if (want_index && idx_groups > 1) {
            std::vector<std::vector<uint64_t>> ri = fm.read_row_index(si, sf, cid);
            size_t expect = 0;
            for (auto& sp : specs) expect += (compressed ? 2 : 1) + sp.extra;
            bool ok = ri.size() == idx_groups;
            size_t lead = 0;
            if (ok) {
                // writers drop the PRESENT positions together with a suppressed PRESENT stream; tolerate
                // files that kept them
                const size_t with_present = expect + (has_present ? 0 : (compressed ? 2 : 1) + 2);
                for (auto& e : ri) {
                    if (e.size() == expect) continue;
                    if (!has_present && e.size() == with_present) { lead = with_present - expect; continue; }
                    ok = false;
                    break;
                }
            }
            if (ok) {
                for (size_t sidx = 0; sidx < specs.size(); sidx++) entries[sidx].resize(idx_groups);
                for (uint32_t g = 0; g < idx_groups && ok; g++) {
                    const std::vector<uint64_t>& p = ri[g];
                    size_t q = p.size() == expect ? 0 : lead;
                    for (size_t sidx = 0; sidx < specs.size() && ok; sidx++) {
                        StreamRef& sr = *specs[sidx].sr;
                        Entry e;
                        uint64_t byte;
                        if (compressed) {
                            const uint64_t cstart = p[q++], inchunk = p[q++];
                            byte = 0;
                            bool found = false;
                            for (size_t ch = 0; ch < sr.chunks.size(); ch++) {
                                if (sr.chunks[ch].hdr_off == cstart) {
                                    byte = sr.chunk_dst[ch] + inchunk;
                                    found = true;
                                    break;
                                }
                            }
                            // a position at the very end of the stream (empty tail group) names no chunk
                            if (!found) {
                                if (!sr.present || cstart >= (sr.chunks.empty() ? 0 : sr.chunks.back().hdr_off + 3ull + sr.chunks.back().src_len))
                                    byte = sr.len;
                                else ok = false;
                            }
                        } else {
                            byte = p[q++];
                        }
                        if (specs[sidx].extra >= 1) e.skip = (uint32_t)p[q++];
                        if (specs[sidx].extra >= 2) e.bit = (uint32_t)p[q++];
                        if (byte > sr.len || e.bit > 7) ok = false;
                        e.byte = (uint32_t)byte;
                        entries[sidx][g] = e;
                    }
                }
            }
            if (ok) {
                indexed = true;
                n_groups = idx_groups;
                gstride = stride;
            }
        }
        if (!indexed) {
            for (size_t sidx = 0; sidx < specs.size(); sidx++) entries[sidx].assign(1, Entry{});
        }
        // where each positioned stream stops being needed (raw-copied streams are cut there)
        std::vector<uint32_t> win_end(specs.size());
        for (size_t sidx = 0; sidx < specs.size(); sidx++) win_end[sidx] = specs[sidx].sr->len;
        if (windowed && indexed) {
            for (size_t sidx = 0; sidx < specs.size(); sidx++) {
                if (wg1 < idx_groups) win_end[sidx] = entries[sidx][wg1].byte;
                entries[sidx] = std::vector<Entry>(entries[sidx].begin() + wg0, entries[sidx].begin() + wg1);
            }
            cp.row_base = wg0 * stride;
            n_rows = std::min<uint64_t>(stripe_rows, (uint64_t)wg1 * stride) - cp.row_base;
            n_groups = wg1 - wg0;
            n_batches = (n_rows + bs - 1) / bs;
            cp.n_rows = n_rows;
            cp.n_batches = n_batches;
        }
        auto win_end_of = [&](StreamRef* sr) -> uint32_t {
            for (size_t sidx = 0; sidx < specs.size(); sidx++)
                if (specs[sidx].sr == sr) return win_end[sidx];
            return sr->len;
        };
        auto spec_of = [&](StreamRef* sr) -> const std::vector<Entry>& {
            for (size_t sidx = 0; sidx < specs.size(); sidx++)
                if (specs[sidx].sr == sr) return entries[sidx];
            static const std::vector<Entry> one(1);
            return one;
        };
        auto rows_in_group = [&](uint32_t g) { return std::min(gstride, n_rows - g * gstride); };

        // ---- PRESENT -> stripe-level validity bitmap, per-group non-null counts, per-batch bitmaps
        uint32_t cnt_base = 0;
        uint64_t valid_raw = 0;
        if (has_present) {
            cnt_base = n_cnt_;
            n_cnt_ += n_groups + 1;
            const uint32_t slot = (gstride + 7) / 8 + 8;
            const uint64_t raw = alloc(AR_TMP, (uint64_t)slot * n_groups + 16);
            valid_raw = alloc(AR_ZERO, ((uint64_t)n_rows + 31) / 32 * 4 + 16);
            const std::vector<Entry>& pe = spec_of(&s_present);
            for (uint32_t g = 0; g < n_groups; g++) {
                const uint32_t rows = rows_in_group(g);
                Seg sg{};
                sg.in = s_present.ptr;
                sg.in_len = s_present.len;
                sg.out = raw;
                sg.start_byte = pe[g].byte;
                sg.run_skip = pe[g].skip;
                sg.n_values = rows;
                sg.cnt_idx = -1;
                sg.start_idx = -1;
                sg.out_start = g * slot;
                sg.colstripe = cs;
                sg.out_kind = OUT_I8;
                sg.aux = 1u | (pe[g].bit << 1);
                                present_byte_segs_.push_back(sg);
                BitSeg b{};
                b.src = raw + (uint64_t)g * slot;
                b.dst = valid_raw;
                b.bit_skip = pe[g].bit;
                b.n_bits = rows;
                b.cnt_idx = -1;
                b.start_idx = -1;
                b.dst_bit0 = g * gstride;
                b.popc_out = (int32_t)(cnt_base + g);
                present_bit_segs_.push_back(b);
            }
            scans_.push_back({cnt_base, n_groups});
            cp.validity_stride = (uint32_t)align_up((bs + 7) / 8, 64);
            cp.validity = alloc(AR_OUT, (uint64_t)cp.validity_stride * n_batches);
            cp.nulls_idx = n_nulls_;
            n_nulls_ += n_batches;
            RepackDesc rp{};
            rp.src = valid_raw;
            rp.dst = cp.validity;
            rp.dst_stride = cp.validity_stride;
            rp.n_rows = n_rows;
            rp.batch_size = bs;
            rp.n_batches = n_batches;
            rp.null_out = (int32_t)cp.nulls_idx;
            rp.batch0 = repack_work_;
            repack_work_ += n_batches;
            repacks_.push_back(rp);
            n_segments_ += n_groups;
            ab_present_ += s_present.len + (uint64_t)n_rows / 8;
            ab_repack_ += (uint64_t)n_rows / 4;
        }
        const int32_t total_idx = has_present ? (int32_t)(cnt_base + n_groups) : -1;

        // run-table slots of a short-run segment: every run is at least two bytes long and (bar corrupt
        // row-index entries) emits at least one value
        auto assign_run_slots = [&](Seg& sg, uint32_t n_bound, uint32_t span_bytes) {
            const uint32_t cap = std::min(span_bytes / 2 + 3, n_bound + 2);
            sg.run_cap = cap;
            pool_blocks_ += (cap + 31) / 32;
            small_values_ += n_bound;
        };
        // helper: integer-RLE segments of one stream into `dst` (dense domain if has_present)
        auto add_int_segs = [&](StreamRef& sr, uint64_t dst, bool is_signed, int nbytes, OutKind okind, uint32_t aux,
                                bool per_group_counts) {
            const std::vector<Entry>& en = spec_of(&sr);
            const uint32_t ng = indexed ? n_groups : 1;
            for (uint32_t g = 0; g < ng; g++) {
                Seg sg{};
                sg.in = sr.ptr;
                sg.in_len = sr.len;
                sg.out = dst;
                sg.start_byte = en[g].byte;
                sg.run_skip = en[g].skip;
                sg.colstripe = cs;
                sg.flags = (is_signed ? SEG_SIGNED : 0) | (v2 ? SEG_RLE_V2 : 0);
                sg.nbytes = (uint8_t)nbytes;
                sg.out_kind = okind;
                sg.aux = aux;
                                if (has_present && per_group_counts) {
                    sg.cnt_idx = (int32_t)(cnt_base + g);
                    sg.start_idx = (int32_t)(cnt_base + g);
                } else {
                    sg.cnt_idx = -1;
                    sg.start_idx = -1;
                    sg.n_values = rows_in_group(g);
                    sg.out_start = g * gstride;
                }
                // Peek at the run header at the entry point (host has the bytes when the file is not
                // compressed): segments that open with a long run go to the warp-per-segment kernel,
                // everything else to the lane-per-segment kernel.  Purely a scheduling hint.
                bool long_runs = false;
                if (v2 && sr.present) {
                    const uint8_t* sp = fm.data + (sf.find(cid, &sr == &s_data ? S_DATA : (&sr == &s_length ? S_LENGTH : S_SECONDARY))->offset);
                    if (!compressed) {
                        long_runs = rle2_opens_with_long_runs(sp, sr.len, sg.start_byte);
                    } else if (!sr.chunks.empty()) {
                        // compressed file: the bytes are readable where the entry point lies in a chunk that was stored
                        // as is (incompressible integer streams usually are)
                        const size_t ci = (size_t)(std::upper_bound(sr.chunk_dst.begin(), sr.chunk_dst.end(), (uint64_t)sg.start_byte) - sr.chunk_dst.begin()) - 1;
                        if (ci < sr.chunks.size() && sr.chunks[ci].original)
                            long_runs = rle2_opens_with_long_runs(sp + sr.chunks[ci].src_off, sr.chunks[ci].src_len,
                                                                  (uint32_t)(sg.start_byte - sr.chunk_dst[ci]));
                    }
                }
                if (!long_runs) {
                    const uint32_t bound = (has_present && per_group_counts) ? rows_in_group(g) : sg.n_values;
                    const uint32_t span = (g + 1 < ng ? en[g + 1].byte : sr.len) - std::min(en[g].byte, sr.len);
                    assign_run_slots(sg, bound, span);
                }
                (long_runs ? int_big_segs_ : int_segs_).push_back(sg);
                static const uint32_t ow1[6] = {2, 4, 8, 4, 4, 1};
                (long_runs ? ab_intbig_ : ab_int_) += (uint64_t)(sr.len / ng) + (uint64_t)rows_in_group(g) * ow1[okind];
            }
            n_segments_ += ng;
        };
        // helper: dense -> rows
        auto add_spaced = [&](uint64_t src, uint64_t dst, uint32_t width, bool late) {
            for (uint32_t g = 0; g < n_groups; g++) {
                SpacedDesc d{};
                d.src = src;
                d.dst = dst;
                d.valid = valid_raw;
                d.row0 = g * gstride;
                d.n_rows = rows_in_group(g);
                d.start_idx = (int32_t)(cnt_base + g);
                d.width = width;
                (late ? spaced_late_ : spaced_).push_back(d);
            }
            ab_spaced_ += (uint64_t)n_rows * (2 * std::max(width, 1u)) + n_rows / 8;
        };
        auto add_copy = [&](uint64_t src, uint32_t src_len, uint64_t dst, uint64_t nbytes, int32_t cnt_idx, uint32_t width,
                            uint64_t max_bytes) {
            CopyDesc d{};
            d.src = src;
            d.dst = dst;
            d.n_bytes = nbytes;
            d.cnt_idx = cnt_idx;
            d.width = width;
            d.src_len = src_len;
            d.colstripe = cs;
            d.u8_col = -1;
            const uint32_t di = (uint32_t)copies_.size();
            copies_.push_back(d);
            ab_copy_ += 2 * max_bytes;
            const uint32_t nt = (uint32_t)((max_bytes + COPY_TILE_BYTES - 1) / COPY_TILE_BYTES);
            for (uint32_t t = 0; t < std::max(nt, 1u); t++) copy_tiles_.push_back(make_uint2(di, t));
        };

        const uint32_t w = oc.width;
        switch (k) {
            case T_SHORT: case T_INT: case T_LONG: case T_DATE: {
                cp.values = alloc(AR_OUT, (uint64_t)n_rows * w);
                uint64_t dst = cp.values;
                if (has_present) dst = alloc(AR_TMP, (uint64_t)n_rows * w);
                const int nb = k == T_SHORT ? 2 : (k == T_LONG ? 8 : 4);
                add_int_segs(s_data, dst, true, nb, k == T_SHORT ? OUT_I16 : (k == T_LONG ? OUT_I64 : OUT_I32), 0, true);
                if (has_present) add_spaced(dst, cp.values, w, false);
                break;
            }
            case T_BYTE: {
                cp.values = alloc(AR_OUT, (uint64_t)n_rows);
                uint64_t dst = cp.values;
                if (has_present) dst = alloc(AR_TMP, (uint64_t)n_rows);
                const std::vector<Entry>& en = spec_of(&s_data);
                const uint32_t ng = indexed ? n_groups : 1;
                for (uint32_t g = 0; g < ng; g++) {
                    Seg sg{};
                    sg.in = s_data.ptr;
                    sg.in_len = s_data.len;
                    sg.out = dst;
                    sg.start_byte = en[g].byte;
                    sg.run_skip = en[g].skip;
                    sg.colstripe = cs;
                    sg.out_kind = OUT_I8;
                                        if (has_present) {
                        sg.cnt_idx = (int32_t)(cnt_base + g);
                        sg.start_idx = (int32_t)(cnt_base + g);
                    } else {
                        sg.cnt_idx = -1;
                        sg.start_idx = -1;
                        sg.n_values = rows_in_group(g);
                        sg.out_start = g * gstride;
                    }
                    data_byte_segs_.push_back(sg);
                }
                n_segments_ += ng;
                if (has_present) add_spaced(dst, cp.values, 1, false);
                break;
            }
            case T_BOOLEAN: {
                const uint32_t ng = indexed ? n_groups : 1;
                const uint32_t slot = (gstride + 7) / 8 + 8;
                const uint64_t raw = alloc(AR_TMP, (uint64_t)slot * ng + 16);
                const uint64_t bm_bytes = ((uint64_t)n_rows + 31) / 32 * 4 + 16;
                const uint64_t rows_bits = alloc(AR_ZERO, bm_bytes);
                const uint64_t dense_bits = has_present ? alloc(AR_ZERO, bm_bytes) : rows_bits;
                const std::vector<Entry>& en = spec_of(&s_data);
                for (uint32_t g = 0; g < ng; g++) {
                    Seg sg{};
                    sg.in = s_data.ptr;
                    sg.in_len = s_data.len;
                    sg.out = raw;
                    sg.start_byte = en[g].byte;
                    sg.run_skip = en[g].skip;
                    sg.out_start = g * slot;
                    sg.start_idx = -1;
                    sg.colstripe = cs;
                    sg.out_kind = OUT_I8;
                    sg.aux = 1u | (en[g].bit << 1);
                                        BitSeg b{};
                    b.src = raw + (uint64_t)g * slot;
                    b.dst = dense_bits;
                    b.bit_skip = en[g].bit;
                    b.popc_out = -1;
                    if (has_present) {
                        sg.cnt_idx = (int32_t)(cnt_base + g);
                        b.cnt_idx = (int32_t)(cnt_base + g);
                        b.start_idx = (int32_t)(cnt_base + g);
                    } else {
                        sg.cnt_idx = -1;
                        sg.n_values = rows_in_group(g);
                        b.cnt_idx = -1;
                        b.n_bits = rows_in_group(g);
                        b.start_idx = -1;
                        b.dst_bit0 = g * gstride;
                    }
                    data_byte_segs_.push_back(sg);
                    data_bit_segs_.push_back(b);
                }
                n_segments_ += ng;
                if (has_present) add_spaced(dense_bits, rows_bits, 0, false);
                cp.values_stride = (uint32_t)align_up((bs + 7) / 8, 64);
                cp.values = alloc(AR_OUT, (uint64_t)cp.values_stride * n_batches);
                RepackDesc rp{};
                rp.src = rows_bits;
                rp.dst = cp.values;
                rp.dst_stride = cp.values_stride;
                rp.n_rows = n_rows;
                rp.batch_size = bs;
                rp.n_batches = n_batches;
                rp.null_out = -1;
                rp.batch0 = repack_work_;
                repack_work_ += n_batches;
                repacks_.push_back(rp);
                break;
            }
            case T_FLOAT: case T_DOUBLE: {
                cp.values = alloc(AR_OUT, (uint64_t)n_rows * w);
                // the window's values start at the first group's recorded byte
                const uint32_t f0 = std::min(spec_of(&s_data)[0].byte, s_data.len);
                if (has_present) {
                    const uint64_t dense = alloc(AR_TMP, (uint64_t)n_rows * w);
                    add_copy(s_data.ptr + f0, s_data.len - f0, dense, 0, total_idx, w, (uint64_t)n_rows * w);
                    add_spaced(dense, cp.values, w, false);
                } else {
                    add_copy(s_data.ptr + f0, s_data.len - f0, cp.values, (uint64_t)n_rows * w, -1, w, (uint64_t)n_rows * w);
                }
                n_segments_ += 1;
                break;
            }
            case T_STRING: case T_VARCHAR: case T_CHAR: case T_BINARY: {
                uint32_t str_d0 = 0, str_dn = 0;  // direct strings: first byte / number of bytes of DATA the decoded rows use
                StrCol sc{};
                sc.n_rows = n_rows;
                sc.batch_size = bs;
                sc.n_batches = n_batches;
                sc.tiles_per_batch = (std::min(bs, n_rows) + STR_TILE - 1) / STR_TILE;
                if (sc.tiles_per_batch == 0) sc.tiles_per_batch = 1;
                sc.n_tiles = sc.tiles_per_batch * n_batches;
                sc.colstripe = cs;
                sc.tile0 = str_tiles_;
                str_tiles_ += sc.n_tiles;
                sc.meta_slot = (uint32_t)strcols_.size();
                cp.str_slot = (int32_t)strcols_.size();
                cp.offsets = alloc(AR_OUT, (uint64_t)n_batches * (bs + 1) * 4);
                sc.offsets = cp.offsets;
                sc.tile_base = alloc(AR_TMP, ((uint64_t)sc.n_tiles + 1) * 8);
                cp.batch_base_off = meta_bytes_;  // running byte cursor inside the batch-base region
                meta_bytes_ += ((uint64_t)n_batches + 1) * 8;
                sc.batch_base = cp.batch_base_off;  // rebased to the meta blob at stage()
                const uint64_t rows_i32 = alloc(AR_TMP, (uint64_t)n_rows * 4 + 16);
                const uint64_t dense_i32 = has_present ? alloc(AR_TMP, (uint64_t)n_rows * 4 + 16) : rows_i32;
                sc.lens = rows_i32;
                if (use_dict) {
                    sc.mode = 1;
                    sc.valid = has_present ? valid_raw : 0;
                    sc.dict_size = enc.dict_size;
                    sc.dict_len = alloc(AR_TMP, (uint64_t)enc.dict_size * 4 + 16);
                    sc.dict_off = alloc(AR_TMP, ((uint64_t)enc.dict_size + 1) * 4 + 16);
                    sc.dict_data = s_dict.ptr;
                    sc.dict_data_len = s_dict.len;
                    // dictionary LENGTH: one segment of dictionary_size unsigned values (string.rs:65-74)
                    if (enc.dict_size) {
                        Seg sg{};
                        sg.in = s_length.ptr;
                        sg.in_len = s_length.len;
                        sg.out = sc.dict_len;
                        sg.n_values = enc.dict_size;
                        sg.cnt_idx = -1;
                        sg.start_idx = -1;
                        sg.colstripe = cs;
                        sg.flags = v2 ? SEG_RLE_V2 : 0;
                        sg.nbytes = 8;
                        sg.out_kind = OUT_LEN31;
                        sg.aux = ORCB_OFFSET_OVERFLOW;
                        assign_run_slots(sg, enc.dict_size, s_length.len);
                        int_segs_.push_back(sg);
                        n_segments_ += 1;
                    }
                    add_int_segs(s_data, dense_i32, false, 8, OUT_LEN31, ORCB_ARROW, true);
                    // heap budget: every row could reference the longest entry; cap by a generous multiple
                    const uint64_t bound = std::min<uint64_t>((uint64_t)n_rows * std::max<uint32_t>(s_dict.len, 1u),
                                                              std::max<uint64_t>((uint64_t)s_dict.len * 8 + (uint64_t)n_rows * 32, 1 << 20));
                    size_[AR_HEAP] += align_up(bound, 256) + 256;
                } else {
                    sc.mode = 0;
                    add_int_segs(s_length, dense_i32, false, 8, OUT_LEN31, ORCB_OFFSET_OVERFLOW, true);
                    // the bytes of the decoded rows: the whole DATA stream, or its part between the window's positions
                    str_d0 = std::min(spec_of(&s_data)[0].byte, s_data.len);
                    str_dn = std::max(std::min(win_end_of(&s_data), s_data.len), str_d0) - str_d0;
                    cp.str_data = alloc(AR_OUT, (uint64_t)str_dn + 16);
                    sc.data = cp.str_data;
                    sc.data_len = str_dn;
                    if (str_dn) {
                        add_copy(s_data.ptr + str_d0, str_dn, cp.str_data, str_dn, -1, 1, str_dn);
                        if (k != T_BINARY) copies_.back().u8_col = (int32_t)strcols_.size();  // validated on the way through
                    }
                    n_segments_ += 1;
                }
                if (has_present) add_spaced(dense_i32, rows_i32, 4, false);
                // lengths/keys read twice (tile sums, offsets) + offsets written; gathered bytes added in finish()
                ab_str_ += (uint64_t)n_rows * 12;
                if (k != T_BINARY) {
                    // Utf8 arrays are validated (string.rs:150-151): direct = the DATA stream, dictionary = its bytes
                    const StreamRef& u8 = use_dict ? s_dict : s_data;
                    const uint32_t u8_off = use_dict ? 0 : str_d0, u8_n = use_dict ? u8.len : str_dn;  // direct: the window's bytes
                    if (u8.present && u8_n) {
                        sc.u8_src = u8.ptr + u8_off;
                        sc.u8_len = u8_n;
                        sc.u8_bad = alloc(AR_ZERO, 16);
                        const uint32_t nt = (uint32_t)(((uint64_t)u8_n + 15) / U8_TILE + 1);  // tiles are cut at aligned addresses
                        sc.u8_flags = alloc(AR_ZERO, ((uint64_t)nt + 32) / 32 * 4 + 16);
                        if (use_dict) {
                            // dictionary bytes get their own pass; direct DATA is checked by the copy kernel
                            for (uint32_t t = 0; t < nt; t++) u8_tiles_.push_back(make_uint2((uint32_t)strcols_.size(), t));
                            ab_utf8_ += u8_n;
                        }
                    }
                }
                strcols_.push_back(sc);
                break;
            }
            case T_DECIMAL: {
                cp.values = alloc(AR_OUT, (uint64_t)n_rows * 16);
                uint64_t dst = cp.values;
                if (has_present) dst = alloc(AR_TMP, (uint64_t)n_rows * 16);
                const uint64_t scales = alloc(AR_TMP, (uint64_t)n_rows * 4 + 16);
                const std::vector<Entry>& en = spec_of(&s_data);
                const uint32_t ng = indexed ? n_groups : 1;
                for (uint32_t g = 0; g < ng; g++) {
                    Seg sg{};
                    sg.in = s_data.ptr;
                    sg.in_len = s_data.len;
                    sg.out = dst;
                    sg.start_byte = en[g].byte;
                    sg.colstripe = cs;
                                        if (has_present) {
                        sg.cnt_idx = (int32_t)(cnt_base + g);
                        sg.start_idx = (int32_t)(cnt_base + g);
                    } else {
                        sg.cnt_idx = -1;
                        sg.start_idx = -1;
                        sg.n_values = rows_in_group(g);
                        sg.out_start = g * gstride;
                    }
                    var_segs_.push_back(sg);
                }
                n_segments_ += ng;
                ab_var_ += s_data.len + (uint64_t)n_rows * 16;
                add_int_segs(s_secondary, scales, true, 4, OUT_SCALE, oc.scale, true);
                DecFixDesc df{};
                df.vals = dst;
                df.scales = scales;
                df.n = n_rows;
                df.cnt_idx = total_idx;
                df.fixed_scale = oc.scale;
                df.colstripe = cs;
                decfix_.push_back(df);
                if (has_present) add_spaced(dst, cp.values, 16, true);
                break;
            }
            case T_TIMESTAMP: case T_TIMESTAMP_INSTANT: {
                const uint32_t tw = oc.ts_decimal ? 16 : 8;  // Decimal128(38, 9) on request (with_schema)
                cp.values = alloc(AR_OUT, (uint64_t)n_rows * tw);
                uint64_t dst = cp.values;
                if (has_present) dst = alloc(AR_TMP, (uint64_t)n_rows * tw);
                const uint64_t secs = alloc(AR_TMP, (uint64_t)n_rows * 8);
                const uint64_t nanos = alloc(AR_TMP, (uint64_t)n_rows * 8);
                add_int_segs(s_data, secs, true, 8, OUT_I64, 0, true);
                add_int_segs(s_secondary, nanos, false, 8, OUT_I64, 0, true);
                int64_t base = ORC_EPOCH_UTC;
                bool tz_on = false;
                std::array<uint64_t, 4> tzt{};
                if (k == T_TIMESTAMP && sf.has_tz && !is_utc_zone(sf.tz)) {
                    // the ORC epoch is 2015-01-01 00:00 on the writer's wall clock (timestamp.rs:128-147); values are
                    // moved from the writer's zone to UTC after decoding (:242-286)
                    auto it = tz_tables_.find(sf.tz);
                    if (it == tz_tables_.end()) {
                        ZoneTable z;
                        std::string why;
                        if (!load_zone_table(sf.tz, z, why)) fail(ORCB_NOT_IMPLEMENTED, "TIMESTAMP column written in a zone this host has no table for: " + why);
                        int64_t b = 0;
                        if (!zone_local_to_utc(z, ORC_EPOCH_UTC, b)) fail(ORCB_UNEXPECTED, "2015-01-01 00:00 is not a unique instant in zone " + sf.tz);
                        auto append = [&](const void* p, size_t n) {
                            tz_blob_.resize((tz_blob_.size() + 15) / 16 * 16);
                            const uint64_t at = tz_blob_.size();
                            tz_blob_.insert(tz_blob_.end(), (const uint8_t*)p, (const uint8_t*)p + n);
                            return at;
                        };
                        std::array<uint64_t, 4> e{};
                        e[0] = append(z.at.data(), z.at.size() * 8);
                        e[1] = append(z.off.data(), z.off.size() * 4);
                        e[2] = z.at.size();
                        e[3] = ((uint64_t)(uint32_t)z.first_off << 32) | (uint64_t)(uint32_t)(int32_t)(b - ORC_EPOCH_UTC);
                        it = tz_tables_.emplace(sf.tz, e).first;
                    }
                    tzt = it->second;
                    base = ORC_EPOCH_UTC + (int64_t)(int32_t)(uint32_t)(tzt[3] & 0xffffffffu);
                    tz_on = true;
                }
                static const int64_t unit_ns[4] = {1, 1000, 1000000, 1000000000};
                TsDesc td{};
                td.secs = secs;
                td.nanos = nanos;
                td.out = dst;
                td.base = base;
                td.unit_ns = unit_ns[oc.ts_unit];
                td.n = n_rows;
                td.cnt_idx = total_idx;
                td.colstripe = cs;
                td.as_i128 = oc.ts_decimal ? 1 : 0;
                if (tz_on) {
                    td.tz_on = 1;
                    td.tz_at = tzt[0];   // offsets inside the zone-table blob until stage()
                    td.tz_off = tzt[1];
                    td.tz_n = (uint32_t)tzt[2];
                    td.tz_first = (int32_t)(uint32_t)(tzt[3] >> 32);
                }
                ts_.push_back(td);
                ab_ts_ += (uint64_t)n_rows * (16 + tw);
                if (has_present) add_spaced(dst, cp.values, tw, true);
                break;
            }
            default: fail(ORCB_NOT_IMPLEMENTED, "unsupported column type on the device path");
        }
        colstripes_.push_back(cp);
    }
}

// ------------------------------------------------------------------------------------------------
// staging: allocate, relocate descriptor pointers, upload
// ------------------------------------------------------------------------------------------------
void Job::stage() {
    if (!planned_) plan();
    if (staged_) return;
    CUDA_OK(cudaSetDevice(opt_.device));
    if (opt_.own_stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        own_stream_ = true;
    } else {
        stream_ = opt_.stream;
    }
    CUDA_OK(cudaEventCreateWithFlags(&done_, cudaEventDisableTiming));
    CUDA_OK(cudaStreamCreateWithFlags(&aux_stream_, cudaStreamNonBlocking));
    CUDA_OK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
    auto arenas = std::make_shared<DeviceArenas>();
    arenas->device = opt_.device;
    for (int a = 1; a < 8; a++) {
        if (size_[a] <= ARENA_PAD && a != AR_TMP) {
            // still give kernels a valid base for empty arenas
        }
        void* p = nullptr;
        CUDA_OK(cudaMalloc(&p, std::max<uint64_t>(size_[a], 256)));
        arenas->ptrs.push_back(p);
        base_[a] = (uint8_t*)p;
    }
    void* p = nullptr;
    CUDA_OK(cudaMalloc(&p, std::max<uint64_t>(desc_bytes_, 256)));
    arenas->ptrs.push_back(p);
    d_desc_ = (uint8_t*)p;
    CUDA_OK(cudaMalloc(&p, std::max<uint64_t>(state_bytes_, 256)));
    arenas->ptrs.push_back(p);
    d_state_ = (uint8_t*)p;
    CUDA_OK(cudaMalloc(&p, std::max<uint64_t>(meta_bytes_, 256)));
    arenas->ptrs.push_back(p);
    d_meta_ = (uint8_t*)p;
    CUDA_OK(cudaHostAlloc((void**)&h_meta_, std::max<uint64_t>(meta_bytes_, 256), cudaHostAllocDefault));
    dev_keepalive_ = arenas;

    // relocate
    auto R = [&](uint64_t& v) { v = reloc(v); };
    auto rseg = [&](std::vector<Seg>& v) { for (auto& s : v) { R(s.in); R(s.out); } };
    rseg(present_byte_segs_);
    rseg(data_byte_segs_);
    rseg(int_segs_);
    rseg(int_big_segs_);
    rseg(var_segs_);
    for (auto* v : {&present_bit_segs_, &data_bit_segs_})
        for (auto& b : *v) { R(b.src); R(b.dst); }
    for (auto& c : copies_) { R(c.src); R(c.dst); }
    for (auto* v : {&spaced_, &spaced_late_})
        for (auto& d : *v) { R(d.src); R(d.dst); R(d.valid); }
    for (auto& d : decfix_) { R(d.vals); R(d.scales); }
    for (auto& d : ts_) {
        R(d.secs); R(d.nanos); R(d.out);
        if (d.tz_on) {
            d.tz_at = (uint64_t)(uintptr_t)(d_desc_ + o_tz_ + d.tz_at);
            d.tz_off = (uint64_t)(uintptr_t)(d_desc_ + o_tz_ + d.tz_off);
        }
    }
    for (auto& c : strcols_) {
        R(c.lens); R(c.valid); R(c.dict_len); R(c.dict_off); R(c.dict_data); R(c.offsets); R(c.tile_base); R(c.data);
        R(c.u8_src); R(c.u8_bad); R(c.u8_flags);
        c.batch_base = (uint64_t)(uintptr_t)(d_meta_ + o_bbase_ + c.batch_base);
    }
    for (auto& d : repacks_) { R(d.src); R(d.dst); }
    for (auto& c : chunks_) { R(c.src); R(c.dst); }

    desc_blob_.assign(std::max<uint64_t>(desc_bytes_, 256), 0);
    auto put = [&](uint64_t off, const void* src, size_t bytes) {
        if (bytes) memcpy(desc_blob_.data() + off, src, bytes);
    };
    put(o_pbyte_, present_byte_segs_.data(), present_byte_segs_.size() * sizeof(Seg));
    put(o_dbyte_, data_byte_segs_.data(), data_byte_segs_.size() * sizeof(Seg));
    put(o_int_, int_segs_.data(), int_segs_.size() * sizeof(Seg));
    put(o_intbig_, int_big_segs_.data(), int_big_segs_.size() * sizeof(Seg));
    put(o_var_, var_segs_.data(), var_segs_.size() * sizeof(Seg));
    put(o_pbit_, present_bit_segs_.data(), present_bit_segs_.size() * sizeof(BitSeg));
    put(o_dbit_, data_bit_segs_.data(), data_bit_segs_.size() * sizeof(BitSeg));
    put(o_scan_, scans_.data(), scans_.size() * sizeof(ScanDesc));
    put(o_copy_, copies_.data(), copies_.size() * sizeof(CopyDesc));
    put(o_ctile_, copy_tiles_.data(), copy_tiles_.size() * sizeof(uint2));
    put(o_u8tile_, u8_tiles_.data(), u8_tiles_.size() * sizeof(uint2));
    put(o_sp_, spaced_.data(), spaced_.size() * sizeof(SpacedDesc));
    put(o_sp2_, spaced_late_.data(), spaced_late_.size() * sizeof(SpacedDesc));
    put(o_dec_, decfix_.data(), decfix_.size() * sizeof(DecFixDesc));
    put(o_ts_, ts_.data(), ts_.size() * sizeof(TsDesc));
    put(o_str_, strcols_.data(), strcols_.size() * sizeof(StrCol));
    put(o_rep_, repacks_.data(), repacks_.size() * sizeof(RepackDesc));
    put(o_chunk_, chunks_.data(), chunks_.size() * sizeof(ChunkDesc));
    put(o_tz_, tz_blob_.data(), tz_blob_.size());
    CUDA_OK(cudaMemcpyAsync(d_desc_, desc_blob_.data(), desc_blob_.size(), cudaMemcpyHostToDevice, stream_));
    for (auto& sc : stage_copies_)
        CUDA_OK(cudaMemcpyAsync(base_[AR_IN] + sc.dst_off, sc.src, sc.bytes, cudaMemcpyHostToDevice, stream_));
    staged_ = true;
}

// ------------------------------------------------------------------------------------------------
// launch: the fixed kernel sequence over the whole plan
// ------------------------------------------------------------------------------------------------
void Job::launch() {
    if (!staged_) stage();
    CUDA_OK(cudaSetDevice(opt_.device));
    cudaStream_t st = stream_;
    // StrCol rows are mutated on the device (dictionary data pointers): refresh them for a re-launch
    if (launched_ && !strcols_.empty())
        CUDA_OK(cudaMemcpyAsync(d_desc_ + o_str_, desc_blob_.data() + o_str_, strcols_.size() * sizeof(StrCol),
                                cudaMemcpyHostToDevice, st));
    CUDA_OK(cudaMemsetAsync(d_state_, 0, state_bytes_, st));
    CUDA_OK(cudaMemsetAsync(d_meta_, 0, meta_bytes_, st));
    if (size_[AR_ZERO] > 0) CUDA_OK(cudaMemsetAsync(base_[AR_ZERO], 0, size_[AR_ZERO], st));
    const uint64_t heap_cap = size_[AR_HEAP] > ARENA_PAD ? size_[AR_HEAP] - ARENA_PAD : 0;

    uint32_t* cnt = (uint32_t*)(d_state_ + o_cnt_);
    uint32_t* dstart = (uint32_t*)(d_state_ + o_dstart_);
    uint32_t* mis = (uint32_t*)(d_state_ + o_mis_);
    JobState* dstate = (JobState*)(d_state_ + o_jobstate_);
    uint32_t* err = (uint32_t*)(d_meta_ + o_err_);
    uint32_t* nulls = (uint32_t*)(d_meta_ + o_nulls_);
    uint64_t* ptrs = (uint64_t*)(d_meta_ + o_ptrs_);
    uint64_t launches = 0;
    auto chk = [&](int rc, const char* what) {
        if (rc) fail(ORCB_CUDA, std::string("launch ") + what + ": " + cudaGetErrorString((cudaError_t)rc));
    };
#define N(v) ((uint32_t)(v).size())
    for (auto& k : kstats_) k.ran = false;
    cudaStream_t cur_st = st;
    auto run = [&](const char* name, uint64_t alg_bytes, uint64_t work, int nk, auto&& fn) {
        KStat& k = kstat(name);
        k.alg_bytes = alg_bytes;
        k.work = work;
        k.ran = true;
        CUDA_OK(cudaEventRecord(k.e0, cur_st));
        chk(fn(), name);
        CUDA_OK(cudaEventRecord(k.e1, cur_st));
        launches += nk;
    };
    if (N(chunks_))
        run("k_decompress", ab_decomp_, N(chunks_), 1, [&] { return launch_decompress((ChunkDesc*)(d_desc_ + o_chunk_), N(chunks_), err, nullptr, st); });
    if (N(present_byte_segs_)) {
        run("k_byte_rle(present)", ab_present_, N(present_byte_segs_), 1, [&] { return launch_byte_rle((Seg*)(d_desc_ + o_pbyte_), N(present_byte_segs_), cnt, dstart, err, st); });
        run("k_bits(present)", ab_present_, N(present_bit_segs_), 1, [&] { return launch_bits((BitSeg*)(d_desc_ + o_pbit_), N(present_bit_segs_), cnt, dstart, st); });
        run("k_seg_scan", 0, N(scans_), 1, [&] { return launch_seg_scan((ScanDesc*)(d_desc_ + o_scan_), N(scans_), cnt, dstart, st); });
    }
    if (N(data_byte_segs_))
        run("k_byte_rle", ab_byte_, N(data_byte_segs_), 1, [&] { return launch_byte_rle((Seg*)(d_desc_ + o_dbyte_), N(data_byte_segs_), cnt, dstart, err, st); });
    if (N(data_bit_segs_))
        run("k_bits", ab_bits_, N(data_bit_segs_), 1, [&] { return launch_bits((BitSeg*)(d_desc_ + o_dbit_), N(data_bit_segs_), cnt, dstart, st); });
    // fork: the header-walk pre-pass is a long dependent chain on few warps, so it runs beside the
    // bandwidth-heavy kernels on a second stream and joins before the epilogues
    // ORCB_SERIAL=1 keeps everything on one stream (clean per-kernel timings when profiling)
    const bool serial_env = getenv("ORCB_SERIAL") != nullptr;  // read at every launch: bench.py times one serial pass
    const bool forked = N(int_segs_) > 0;
    cudaStream_t aux = serial_env ? st : aux_stream_;
    if (forked) {
        if (!serial_env) {
            CUDA_OK(cudaEventRecord(ev_fork_, st));
            CUDA_OK(cudaStreamWaitEvent(aux, ev_fork_, 0));
        }
        cur_st = aux;
        RunRec* rtab = (RunRec*)(uintptr_t)reloc(run_table_);
        BlockRec* brec = (BlockRec*)(uintptr_t)reloc(block_recs_);
        uint32_t* nblk = (uint32_t*)(d_state_ + o_nblocks_);
        run("k_rle_index", 0, N(int_segs_), 1, [&] { return launch_rle_index((Seg*)(d_desc_ + o_int_), N(int_segs_), cnt, rtab, brec, nblk, pool_blocks_, (CoopRec*)(uintptr_t)reloc(coop_q_), nblk + 2, coop_cap_, err, aux); });
        run("k_int_rle(+general,+coop_runs)", ab_int_, pool_blocks_, 3, [&] { return launch_int_rle((Seg*)(d_desc_ + o_int_), brec, nblk, pool_blocks_, rtab, cnt, dstart, err, mis, (uint32_t*)(uintptr_t)reloc(slow_list_), nblk + 1, (CoopRec*)(uintptr_t)reloc(coop_q_), nblk + 2, coop_cap_, aux); });
        if (!serial_env) CUDA_OK(cudaEventRecord(ev_join_, aux));
        cur_st = st;
    }
    // main-stream kernels that run beside the short-run integer path: the copy first (bandwidth-bound, it leaves
    // the issue slots to the latency-bound header walk), then the issue-bound decoders
    if (N(u8_tiles_))
        run("k_utf8", ab_utf8_, N(u8_tiles_), 1, [&] { return launch_utf8((StrCol*)(d_desc_ + o_str_), (uint2*)(d_desc_ + o_u8tile_), N(u8_tiles_), st); });
    static const char* order_env = getenv("ORCB_MAIN_ORDER");
    const char* order = order_env ? order_env : "kcv";
    for (const char* o = order; *o; o++) {
        if (*o == 'c' && N(int_big_segs_))
            run("k_int_rle_coop", ab_intbig_, N(int_big_segs_), 1, [&] { return launch_int_rle_coop((Seg*)(d_desc_ + o_intbig_), N(int_big_segs_), cnt, dstart, err, mis, st); });
        if (*o == 'v' && N(var_segs_))
            run("k_varint128", ab_var_, N(var_segs_), 1, [&] { return launch_varint128((Seg*)(d_desc_ + o_var_), N(var_segs_), cnt, dstart, err, st); });
        if (*o == 'k' && N(copy_tiles_))
            run("k_copy", ab_copy_, N(copy_tiles_), 1, [&] { return launch_copy((CopyDesc*)(d_desc_ + o_copy_), (uint2*)(d_desc_ + o_ctile_), N(copy_tiles_), cnt, err, (StrCol*)(d_desc_ + o_str_), st); });
    }
    if (forked && !serial_env) CUDA_OK(cudaStreamWaitEvent(st, ev_join_, 0));
    if (N(decfix_))
        run("k_decimal_fix", ab_dec_, N(decfix_), 1, [&] { return launch_decimal_fix((DecFixDesc*)(d_desc_ + o_dec_), N(decfix_), cnt, mis, st); });
    if (N(ts_))
        run("k_timestamp", ab_ts_, N(ts_), 1, [&] { return launch_timestamp((TsDesc*)(d_desc_ + o_ts_), N(ts_), cnt, err, st); });
    if (N(spaced_))
        run("k_spaced", ab_spaced_, N(spaced_), 1, [&] { return launch_spaced((SpacedDesc*)(d_desc_ + o_sp_), N(spaced_), dstart, st); });
    if (N(spaced_late_))
        run("k_spaced(late)", ab_spaced_, N(spaced_late_), 1, [&] { return launch_spaced((SpacedDesc*)(d_desc_ + o_sp2_), N(spaced_late_), dstart, st); });
    if (N(strcols_))
        run("k_strings(5 kernels)", ab_str_, str_tiles_, 5, [&] { return launch_strings((StrCol*)(d_desc_ + o_str_), N(strcols_), str_tiles_, err, dstate, (uint64_t)(uintptr_t)base_[AR_HEAP], heap_cap, ptrs, st); });
    if (repack_work_)
        run("k_repack", ab_repack_, repack_work_, 1, [&] { return launch_repack((RepackDesc*)(d_desc_ + o_rep_), N(repacks_), repack_work_, nulls, st); });
#undef N
    n_launches_ = launches;
    launched_ = true;
    finished_ = false;
    host_out_.reset();
}

void Job::finish() {
    if (!launched_) launch();
    if (finished_) return;
    CUDA_OK(cudaSetDevice(opt_.device));
    CUDA_OK(cudaMemcpyAsync(h_meta_, d_meta_, meta_bytes_, cudaMemcpyDeviceToHost, stream_));
    CUDA_OK(cudaEventRecord(done_, stream_));
    CUDA_OK(cudaStreamSynchronize(stream_));
    for (auto& k : kstats_) {
        if (!k.ran) continue;
        float ms = 0;
        if (cudaEventElapsedTime(&ms, k.e0, k.e1) == cudaSuccess) k.ms = ms;
        else cudaGetLastError();
    }
    const uint32_t* err = (const uint32_t*)(h_meta_ + o_err_);
    for (uint32_t i = 0; i < n_colstripes_; i++) {
        if (err[i]) {
            const ColStripePlan& cp = colstripes_[i];
            fail((int)err[i], "device decode error in column '" + cols_[cp.col].name + "' of stripe " +
                                  std::to_string(tasks_[cp.task].stripe));
        }
    }
    // logical output bytes: values + offsets + string bytes + validity where emitted
    output_bytes_ = 0;
    const uint32_t* nulls = (const uint32_t*)(h_meta_ + o_nulls_);
    for (auto& cp : colstripes_) {
        const OutColumn& oc = cols_[cp.col];
        const uint32_t bs = opt_.batch_size;
        for (uint32_t b = 0; b < cp.n_batches; b++) {
            const uint32_t rows = std::min(bs, cp.n_rows - b * bs);
            if (cp.has_present && nulls[cp.nulls_idx + b]) output_bytes_ += (rows + 7) / 8;
            if (oc.kind == T_BOOLEAN) output_bytes_ += (rows + 7) / 8;
            else if (cp.str_slot >= 0) {
                const int64_t* bb = (const int64_t*)(h_meta_ + o_bbase_ + cp.batch_base_off);
                output_bytes_ += 4ull * (rows + 1) + (uint64_t)(bb[b + 1] - bb[b]);
            } else output_bytes_ += (uint64_t)rows * oc.width;
        }
    }
    finished_ = true;
}

void Job::stats(OrcbJobStats* out) const {
    memset(out, 0, sizeof(*out));
    out->n_stripes = tasks_.size();
    out->n_rows = n_rows_;
    out->n_columns = cols_.size();
    out->input_bytes = input_bytes_;
    for (auto& sc : stage_copies_) out->staged_bytes += sc.bytes;
    out->staged_bytes += desc_bytes_;
    out->output_bytes = output_bytes_;
    for (int a = 1; a < 8; a++) out->device_bytes += size_[a];
    out->device_bytes += desc_bytes_ + state_bytes_ + meta_bytes_;
    out->n_segments = n_segments_;
    out->n_kernel_launches = n_launches_;
    out->n_batches = batch_task_.size();
    out->d2h_meta_bytes = meta_bytes_;
}

// ------------------------------------------------------------------------------------------------
// Arrow export
// ------------------------------------------------------------------------------------------------
namespace {
struct ArrayPriv {
    std::shared_ptr<void> keep1, keep2;
    std::vector<const void*> buffers;
    std::vector<ArrowArray> child_store;
    std::vector<ArrowArray*> child_ptrs;
};
void release_array(ArrowArray* a) {
    if (!a || !a->release) return;
    for (int64_t i = 0; i < a->n_children; i++)
        if (a->children[i]->release) a->children[i]->release(a->children[i]);
    delete (ArrayPriv*)a->private_data;
    a->release = nullptr;
}
void init_array(ArrowArray* a, int64_t length, int64_t null_count, size_t n_buffers) {
    auto* p = new ArrayPriv();
    p->buffers.assign(n_buffers, nullptr);
    a->length = length;
    a->null_count = null_count;
    a->offset = 0;
    a->n_buffers = (int64_t)n_buffers;
    a->n_children = 0;
    a->buffers = p->buffers.data();
    a->children = nullptr;
    a->dictionary = nullptr;
    a->release = release_array;
    a->private_data = p;
}
}  // namespace

void Job::ensure_host_output() {
    if (host_out_) return;
    if (!finished_) finish();
    auto ho = std::make_shared<HostOutput>();
    CUDA_OK(cudaSetDevice(opt_.device));
    const uint64_t out_bytes = std::max<uint64_t>(size_[AR_OUT], 256);
    CUDA_OK(cudaHostAlloc((void**)&ho->out, out_bytes, cudaHostAllocDefault));
    CUDA_OK(cudaMemcpyAsync(ho->out, base_[AR_OUT], out_bytes, cudaMemcpyDeviceToHost, stream_));
    // used part of the heap: max over dictionary columns of (ptr - heap_base + total)
    uint64_t heap_used = 0;
    const uint64_t* ptrs = (const uint64_t*)(h_meta_ + o_ptrs_);
    for (auto& cp : colstripes_) {
        if (cp.str_slot < 0 || strcols_[cp.str_slot].mode != 1) continue;
        const int64_t* bb = (const int64_t*)(h_meta_ + o_bbase_ + cp.batch_base_off);
        const uint64_t p = ptrs[cp.str_slot];
        if (p) heap_used = std::max<uint64_t>(heap_used, p - (uint64_t)(uintptr_t)base_[AR_HEAP] + (uint64_t)bb[cp.n_batches]);
    }
    if (heap_used) {
        CUDA_OK(cudaHostAlloc((void**)&ho->heap, heap_used, cudaHostAllocDefault));
        CUDA_OK(cudaMemcpyAsync(ho->heap, base_[AR_HEAP], heap_used, cudaMemcpyDeviceToHost, stream_));
    }
    CUDA_OK(cudaStreamSynchronize(stream_));
    host_out_ = ho;
}

static void build_batch(const Job* /*unused*/, ArrowArray* out, int64_t rows, size_t ncols) {
    init_array(out, rows, 0, 1);
    auto* p = (ArrayPriv*)out->private_data;
    p->child_store.resize(ncols);
    p->child_ptrs.resize(ncols);
    for (size_t i = 0; i < ncols; i++) p->child_ptrs[i] = &p->child_store[i];
    out->n_children = (int64_t)ncols;
    out->children = p->child_ptrs.data();
}

void Job::export_batch(uint64_t i, ArrowArray* out) {
    if (i >= batch_task_.size()) fail(ORCB_INVALID_ARGUMENT, "batch index out of range");
    ensure_host_output();
    const uint32_t t = batch_task_[i], b = batch_idx_[i];
    const uint32_t bs = opt_.batch_size;
    const uint32_t* nulls = (const uint32_t*)(h_meta_ + o_nulls_);
    const uint64_t* ptrs = (const uint64_t*)(h_meta_ + o_ptrs_);
    const uint32_t cs0 = task_first_cs_[t];
    // (the stripe's row count also sizes the batches of an empty projection, mod.rs:538-549)
    const uint32_t rows = batch_rows_[i];
    const int64_t vrow0 = view_mode_ ? (int64_t)batch_row0_[i] : 0;  // view into the stripe-wide internal batch
    build_batch(this, out, rows, cols_.size());
    auto* tp = (ArrayPriv*)out->private_data;
    tp->keep1 = host_out_;
    const uint64_t omask = (1ull << 60) - 1;
    for (size_t c = 0; c < cols_.size(); c++) {
        const ColStripePlan& cp = colstripes_[cs0 + c];
        const OutColumn& oc = cols_[c];
        ArrowArray* a = &tp->child_store[c];
        const bool is_str = cp.str_slot >= 0;
        const int64_t voff = view_mode_ ? vrow0 - (int64_t)cp.row_base : 0;  // the column may hold a row-group window only
        int64_t nc = 0;
        const void* vbuf = nullptr;
        if (cp.has_present) {
            const uint8_t* bits = host_out_->out + (cp.validity & omask) + (uint64_t)b * cp.validity_stride;
            if (view_mode_) {
                // nulls inside the view: counted here, the device only knows the stripe-wide figure
                int64_t valid = 0;
                for (int64_t r = voff; r < voff + rows; r++) valid += (bits[r >> 3] >> (r & 7)) & 1;
                nc = rows - valid;
            } else {
                nc = nulls[cp.nulls_idx + b];
            }
            if (nc) vbuf = bits;
        }
        init_array(a, rows, nc, is_str ? 3 : 2);
        a->offset = voff;
        auto* ap = (ArrayPriv*)a->private_data;
        ap->keep1 = host_out_;
        ap->buffers[0] = vbuf;
        if (is_str) {
            const int64_t* bb = (const int64_t*)(h_meta_ + o_bbase_ + cp.batch_base_off);
            ap->buffers[1] = host_out_->out + (cp.offsets & omask) + (uint64_t)b * (bs + 1) * 4;
            const uint8_t* dbase;
            if (strcols_[cp.str_slot].mode == 1) {
                const uint64_t p = ptrs[cp.str_slot];
                dbase = host_out_->heap ? host_out_->heap + (p - (uint64_t)(uintptr_t)base_[AR_HEAP]) : host_out_->out;
            } else {
                dbase = host_out_->out + (cp.str_data & omask);
            }
            ap->buffers[2] = dbase + bb[b];
        } else if (oc.kind == T_BOOLEAN) {
            ap->buffers[1] = host_out_->out + (cp.values & omask) + (uint64_t)b * cp.values_stride;
        } else {
            ap->buffers[1] = host_out_->out + (cp.values & omask) + (uint64_t)b * bs * oc.width;
        }
    }
}

void Job::export_batch_device(uint64_t i, ArrowDeviceArray* out) {
    if (i >= batch_task_.size()) fail(ORCB_INVALID_ARGUMENT, "batch index out of range");
    if (!finished_) finish();
    const uint32_t t = batch_task_[i], b = batch_idx_[i];
    const uint32_t bs = opt_.batch_size;
    const uint32_t* nulls = (const uint32_t*)(h_meta_ + o_nulls_);
    const uint64_t* ptrs = (const uint64_t*)(h_meta_ + o_ptrs_);
    const uint32_t cs0 = task_first_cs_[t];
    const uint32_t rows = batch_rows_[i];
    const int64_t vrow0 = view_mode_ ? (int64_t)batch_row0_[i] : 0;
    memset(out, 0, sizeof(*out));
    build_batch(this, &out->array, rows, cols_.size());
    out->device_id = opt_.device;
    out->device_type = ARROW_DEVICE_CUDA;
    out->sync_event = nullptr;  // finish() already synchronised the stream
    auto* tp = (ArrayPriv*)out->array.private_data;
    tp->keep1 = dev_keepalive_;
    const uint64_t omask = (1ull << 60) - 1;
    for (size_t c = 0; c < cols_.size(); c++) {
        const ColStripePlan& cp = colstripes_[cs0 + c];
        const OutColumn& oc = cols_[c];
        ArrowArray* a = &tp->child_store[c];
        const bool is_str = cp.str_slot >= 0;
        const int64_t voff = view_mode_ ? vrow0 - (int64_t)cp.row_base : 0;
        int64_t nc = 0;
        const void* vbuf = nullptr;
        if (cp.has_present) {
            nc = view_mode_ ? (nulls[cp.nulls_idx + b] ? -1 : 0) : (int64_t)nulls[cp.nulls_idx + b];  // -1: not counted for a view
            if (nc) vbuf = base_[AR_OUT] + (cp.validity & omask) + (uint64_t)b * cp.validity_stride;
        }
        init_array(a, rows, nc, is_str ? 3 : 2);
        a->offset = voff;
        auto* ap = (ArrayPriv*)a->private_data;
        ap->keep1 = dev_keepalive_;
        ap->buffers[0] = vbuf;
        if (is_str) {
            const int64_t* bb = (const int64_t*)(h_meta_ + o_bbase_ + cp.batch_base_off);
            ap->buffers[1] = base_[AR_OUT] + (cp.offsets & omask) + (uint64_t)b * (bs + 1) * 4;
            const uint8_t* dbase = strcols_[cp.str_slot].mode == 1 ? (const uint8_t*)(uintptr_t)ptrs[cp.str_slot]
                                                                   : base_[AR_OUT] + (cp.str_data & omask);
            ap->buffers[2] = dbase + bb[b];
        } else if (oc.kind == T_BOOLEAN) {
            ap->buffers[1] = base_[AR_OUT] + (cp.values & omask) + (uint64_t)b * cp.values_stride;
        } else {
            ap->buffers[1] = base_[AR_OUT] + (cp.values & omask) + (uint64_t)b * bs * oc.width;
        }
    }
}


std::vector<std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>>> selection_views(
    std::vector<RowSelector> raw, const std::vector<uint64_t>& stripe_rows, uint64_t batch_size) {
    // RowSelection::from(Vec<RowSelector>): empty selectors dropped, neighbours of the same kind merged
    std::vector<RowSelector> sel;
    for (auto& r : raw) {
        if (r.row_count == 0) continue;
        if (!sel.empty() && sel.back().skip == r.skip) sel.back().row_count += r.row_count;
        else sel.push_back(r);
    }
    auto total = [](const std::vector<RowSelector>& v) {
        uint64_t n = 0;
        for (auto& x : v) n += x.row_count;
        return n;
    };
    // RowSelection::split_off: the first `n` rows leave `self`
    auto split_off = [](std::vector<RowSelector>& self, uint64_t n) {
        uint64_t acc = 0;
        size_t idx = self.size();
        for (size_t i = 0; i < self.size(); i++) {
            acc += self[i].row_count;
            if (acc > n) { idx = i; break; }
        }
        if (idx == self.size()) {
            std::vector<RowSelector> all;
            all.swap(self);
            return all;
        }
        std::vector<RowSelector> head(self.begin(), self.begin() + idx), rest(self.begin() + idx, self.end());
        const uint64_t overflow = acc - n;
        if (rest.front().row_count != overflow) head.push_back({rest.front().row_count - overflow, rest.front().skip});
        rest.front().row_count = overflow;
        self.swap(rest);
        return head;
    };
    std::vector<std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>>> out;
    for (uint64_t rows : stripe_rows) {
        std::vector<std::pair<uint32_t, uint32_t>> views;
        if (total(sel) == 0) {  // arrow_reader.rs:298: a used-up selection no longer restricts anything
            out.emplace_back(false, views);
            continue;
        }
        const std::vector<RowSelector> s = split_off(sel, rows);
        // NaiveStripeDecoder::next / next_with_row_selection.  A selector is left behind only once a single step has
        // covered its whole row_count, so a select longer than the batch size keeps yielding batches (kept as is)
        uint64_t index = 0;
        size_t si = 0;
        while (index < rows && si < s.size()) {
            const uint64_t remaining = rows - index;
            if (s[si].skip) {
                const uint64_t k = std::min(s[si].row_count, remaining);
                if (k == 0) { si++; continue; }
                index += k;
                if (k >= s[si].row_count) si++;
            } else {
                const uint64_t k = std::min(std::min(s[si].row_count, batch_size), remaining);
                if (k == 0) { si++; continue; }
                views.emplace_back((uint32_t)index, (uint32_t)k);
                index += k;
                if (k >= s[si].row_count) si++;
            }
        }
        out.emplace_back(true, views);
    }
    return out;
}

}  // namespace orcb
