// Host planner + executor.  The planner restates, as table construction, what the reference does
// stream by stream at run time:
//   Stripe::new stream offsets            src/stripe.rs:128-182
//   array_decoder_factory type dispatch   src/array_decoder/mod.rs:390-511
//   Column::rle_version / dictionary_size src/column.rs:40-59
//   new_timestamp_decoder base selection  src/array_decoder/timestamp.rs:128-147
//   NaiveStripeDecoder batch boundaries   src/array_decoder/mod.rs:371-387, 514-564
// and adds what the reference never uses: row-index positions (src/row_index.rs:37-51) as parallel
// entry points, one segment per (stream, row group).
#include "job_internal.h"
#include "tz.h"

namespace orcb {

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

// Scheduling hint only: do the first few (4; ORCB_PEEK_RUNS) RLE v2 runs at `pos` all hold more than 64 values?  Every
// run looked at is a cache miss in a large file; 4 against 8 runs: plan 9.0 -> 6.8 ms for a 9-stripe file, SF10 decode
// 6.45 -> 6.37 ms (8 % more segments go to the warp-per-segment kernel).  (Header walk,
// no values decoded.)  Such segments go to the warp-per-segment kernel.
bool rle2_opens_with_long_runs(const uint8_t* s, uint32_t len, uint32_t pos) {
    static const int W[32] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 26, 28, 30, 32, 40, 48, 56, 64};
    static const int depth = getenv("ORCB_PEEK_RUNS") ? std::max(1, atoi(getenv("ORCB_PEEK_RUNS"))) : 4;
    for (int r = 0; r < depth; r++) {
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
    task_in_off_.clear();
    staged_stripes_.clear();
    for (auto& t : tasks_) view_mode_ |= t.has_views;
    // nested columns: children have as many slots as their parents' data says, batches of a list's children do not
    // fall on fixed boundaries - everything is decoded as one internal batch and exported as views
    nested_ = has_nested_columns(cols_);
    for (auto& t : tasks_) nested_ |= !t.roots.empty();
    view_mode_ |= nested_;
    user_batch_size_ = opt_.batch_size;
    if (view_mode_) {
        // row selection: one internal batch per stripe (stripe-level offsets and bitmaps); the user's batches are
        // views into it, exported through the Arrow `offset` field
        uint64_t mx = 1;
        for (auto& t : tasks_) {
            mx = std::max<uint64_t>(mx, t.file->stripes[t.stripe].rows);
            for (auto& r : t.roots) mx = std::max<uint64_t>(mx, r.n_slots);
        }
        if (mx > 0xfffffff0ull) fail(ORCB_NOT_IMPLEMENTED, "more than 2^32 slots in one nested column of a stripe");
        opt_.batch_size = (uint32_t)mx;
    }
    for (uint32_t t = 0; t < tasks_.size(); t++) plan_stripe(t);

    // The header walk gives every segment one lane for as long as the segment has runs: the longest chains go first
    // (they bound the kernel's duration) and lanes of one warp get chains of similar length.  Nothing refers to a
    // short-run segment by index before the walk assigns it blocks on the device.
    {
        static const char* e = getenv("ORCB_SORT_SEGS");
        n_str_int_segs_ = 0;
        if (!(e && e[0] == '0')) {
            // ... and the segments the string kernels wait for (lengths, dictionary keys) in front of all others: they are
            // walked and decoded as a first phase, so that the string kernels start while the rest is still being decoded
            auto key = [](const Seg& a) { return ((uint64_t)(a.out_kind == OUT_LEN31) << 32) | a.run_cap; };
            std::stable_sort(int_segs_.begin(), int_segs_.end(), [&](const Seg& a, const Seg& b) { return key(a) > key(b); });
            for (auto& sg : int_segs_) n_str_int_segs_ += sg.out_kind == OUT_LEN31;
        }
    }

    // chunks that take longest first (compressed before stored, long before short): the persistent decompression
    // warps draw them in this order, so the launch does not end on one late, expensive chunk
    // (Zlib / Zstandard / LZO chunks in front, then Snappy, then LZ4 and stored: one kernel each, launch_decompress)
    auto chunk_class = [](const ChunkDesc& c) -> uint64_t { return c.codec == 1 || c.codec == 3 || c.codec == 5 ? 2 : c.codec == 2 ? 1 : 0; };
    auto chunk_key = [&](const ChunkDesc& c) { return (chunk_class(c) << 40) | ((uint64_t)(c.codec != 0) << 32) | c.src_len; };
    std::stable_sort(chunks_.begin(), chunks_.end(), [&](const ChunkDesc& a, const ChunkDesc& b) { return chunk_key(a) > chunk_key(b); });
    n_bits_chunks_ = n_snappy_chunks_ = bits_codecs_ = 0;
    for (auto& c : chunks_) {
        if (chunk_class(c) == 2) bits_codecs_ |= 1u << c.codec;
        n_bits_chunks_ += chunk_class(c) == 2;
        n_snappy_chunks_ += chunk_class(c) == 1;
    }

    if (n_chk_) chk_table_ = alloc(AR_TMP, (uint64_t)n_chk_ * sizeof(SegCheck));
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
    place(o_spm_, merge_spaced_.size() * sizeof(SpacedDesc));
    place(o_popc_, popcs_.size() * sizeof(PopcDesc));
    place(o_union_, unions_.size() * sizeof(UnionDesc));
    place(o_chkpair_, chk_pairs_.size() * sizeof(uint2));
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
    splace(o_nblocks_, 48);
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
    mplace(o_clens_, (chunk_keys_.size() + 1) * 4);
    mplace(o_retry_, 16);
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
    if (std::shared_ptr<RangeBuf> held = fm.load_stripe(stripe)) range_keep_.push_back(held);  // callback files: one read per stripe
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
    if (tasks_[task_idx].has_staged) {
        in_off = aref(AR_ABS, tasks_[task_idx].staged_abs);
    } else if (si.data_length) {
        auto key = std::make_pair((const void*)&fm, stripe);
        auto it = staged_stripes_.find(key);
        if (it != staged_stripes_.end()) {
            in_off = it->second;
        } else {
            in_off = alloc(AR_IN, si.data_length);
            stage_copies_.push_back({fm.base_for(data_start, si.data_length) + data_start, in_off & ((1ull << 60) - 1), si.data_length});
            staged_stripes_[key] = in_off;
        }
    }

    task_in_off_.push_back(in_off);

    // row-index stride usable for this stripe?
    uint32_t stride = stripe_rows ? stripe_rows : 1;
    bool want_index = opt_.use_row_index && fm.row_index_stride > 0 && stripe_rows > 0;
    if (want_index) stride = (uint32_t)std::min<uint64_t>((uint64_t)fm.row_index_stride, 0xffffffffull);
    const uint32_t idx_groups = stripe_rows ? (stripe_rows + stride - 1) / stride : 0;
    // partial decode (row selection): row groups [wg0, wg1) only
    const StripeTask& task = tasks_[task_idx];
    const bool windowed = !nested_ && task.has_window && want_index && idx_groups > 1 && task.g_begin < task.g_end && task.g_end <= idx_groups;
    const uint32_t wg0 = windowed ? task.g_begin : 0, wg1 = windowed ? task.g_end : idx_groups;

    for (uint32_t ci = 0; ci < cols_.size(); ci++) {
        const OutColumn& oc = cols_[ci];
        const uint32_t cid = oc.col_id;
        const uint32_t cs = (uint32_t)colstripes_.size();
        // a column of a later nesting level decodes the slots its parent gave it, possibly under the parent's validity
        const RootSpec* rs = task.roots.empty() ? nullptr : &task.roots.at(ci);
        const bool has_parent = rs && rs->has_parent;
        if (rs && rs->n_slots > 0xfffffff0ull) fail(ORCB_NOT_IMPLEMENTED, "more than 2^32 slots in one nested column of a stripe");
        uint32_t n_rows = rs ? (uint32_t)rs->n_slots : stripe_rows;  // rows this column decodes: the stripe's, or the window's once the index is known good
        uint32_t n_batches = (n_rows + bs - 1) / bs;
        ColStripePlan cp;
        cp.task = task_idx;
        cp.col = ci;
        cp.n_rows = n_rows;
        cp.n_batches = n_batches;
        if (n_rows == 0) {
            cp.kids = oc.child_ids;
            cp.is_list = oc.kind == T_LIST || oc.kind == T_MAP;
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
            if (st->offset < data_start || st->offset - data_start > si.data_length || st->length > si.data_length - (st->offset - data_start))
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
                d.id = (uint32_t)chunk_keys_.size();
                chunk_keys_.emplace_back(&fm, st->offset + c.hdr_off);
                if (c.dst_len >= 0) {
                    d.dst_cap = (uint32_t)c.dst_len;
                    d.expect_len = (int32_t)c.dst_len;
                    d.assumed = c.guess ? 1 : 0;
                } else {
                    d.dst_cap = (uint32_t)fm.block_size;
                    // The layout assumes that every non-final chunk fills the block (writers cut a chunk when their
                    // buffer is full).  The device checks; where it does not hold (Java writers also cut chunks at
                    // other points) the job learns the real sizes and is planned again (Job::finish, LayoutRetry).
                    d.expect_len = i + 1 < r.chunks.size() ? (int32_t)fm.block_size : -1;
                    d.assumed = 1;
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
        if (is_str || k == T_LIST || k == T_MAP) s_length = resolve(S_LENGTH);
        if (use_dict) s_dict = resolve(S_DICTIONARY_DATA);
        if (k == T_DECIMAL || k == T_TIMESTAMP || k == T_TIMESTAMP_INSTANT) s_secondary = resolve(S_SECONDARY);
        const bool own_present = s_present.present;
        const bool has_present = own_present || has_parent;  // the column has a validity bitmap of its own slots
        cp.has_present = has_present;
        cp.kids = oc.child_ids;

        // ---- row-index entries -> per-stream entry points
        // stream order inside a RowIndexEntry.positions list follows the writers' (Java/C++) recording order
        struct PosSpec {
            StreamRef* sr;
            int extra;
        };
        std::vector<PosSpec> specs;
        if (own_present) specs.push_back({&s_present, 2});
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
            case T_LIST: case T_MAP: specs.push_back({&s_length, 1}); break;
            case T_UNION: specs.push_back({&s_data, 1}); break;
            default: break;
        }
        uint32_t n_groups = 1;
        uint32_t gstride = n_rows;
        std::vector<std::vector<Entry>> entries(specs.size());  // [spec][group]
        bool indexed = false;
        if (want_index && idx_groups > 1 && !rs && oc.child_ids.empty()) {
            // (the reference never reads ROW_INDEX streams on this path: one that does not even parse must not fail the
            // decode - the column is decoded sequentially, like any column whose positions do not fit)
            std::vector<std::vector<uint64_t>> ri;
            bool readable = true;
            try {
                ri = fm.read_row_index(si, sf, cid);
            } catch (const OrcException&) {
                readable = false;
            }
            size_t expect = 0;
            for (auto& sp : specs) expect += (compressed ? 2 : 1) + sp.extra;
            bool ok = readable && ri.size() == idx_groups;
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
        // row-index consistency (SegCheck): consecutive row groups of one stream are linked; `prev` is per stream
        auto link = [&](Seg& sg, uint32_t g, uint32_t& prev) {
            if (!indexed) return;
            const uint32_t slot = n_chk_++;
            sg.chk = slot + 1;
            if (g > 0) chk_pairs_.push_back(make_uint2(prev, slot));
            prev = slot;
        };

        // ---- PRESENT -> stripe-level validity bitmap, per-group non-null counts, per-batch bitmaps
        uint32_t cnt_base = 0;
        uint64_t valid_raw = 0;
        uint32_t chk_prev_present = 0;
        int32_t validity_repack = -1;  // index of the RepackDesc that builds this column's per-batch validity
        if (has_present) {
            cnt_base = n_cnt_;
            n_cnt_ += n_groups + 1;
            if (!has_parent) {
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
                    link(sg, g, chk_prev_present);
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
            } else {
                // Validity handed down by the parent (struct and union children; one group).  The column's own PRESENT
                // stream has one entry per valid slot of the parent and is scattered into those slots
                // (merge_parent_present, array_decoder/mod.rs:216-229); without one the parent's bitmap is the column's.
                const uint64_t parent_bits = aref(AR_ABS, rs->parent_bits);
                if (own_present) {
                    if (rs->parent_count > n_rows) fail(ORCB_UNEXPECTED, "parent validity count exceeds the slots");
                    const uint32_t pc = (uint32_t)rs->parent_count;
                    const uint64_t raw = alloc(AR_TMP, (uint64_t)pc / 8 + 32);
                    const uint64_t own_bits = alloc(AR_ZERO, ((uint64_t)pc + 31) / 32 * 4 + 16);
                    Seg sg{};
                    sg.in = s_present.ptr;
                    sg.in_len = s_present.len;
                    sg.out = raw;
                    sg.n_values = pc;
                    sg.cnt_idx = -1;
                    sg.start_idx = -1;
                    sg.colstripe = cs;
                    sg.out_kind = OUT_I8;
                    sg.aux = 1u;
                    present_byte_segs_.push_back(sg);
                    BitSeg b{};
                    b.src = raw;
                    b.dst = own_bits;
                    b.n_bits = pc;
                    b.cnt_idx = -1;
                    b.start_idx = -1;
                    b.popc_out = -1;
                    present_bit_segs_.push_back(b);
                    valid_raw = alloc(AR_ZERO, ((uint64_t)n_rows + 31) / 32 * 4 + 16);
                    SpacedDesc d{};
                    d.src = own_bits;
                    d.dst = valid_raw;
                    d.valid = parent_bits;
                    d.n_rows = n_rows;
                    d.start_idx = (int32_t)n_cnt_++;  // a slot nothing writes: its dense start stays 0
                    d.width = 0;
                    merge_spaced_.push_back(d);
                    ab_present_ += s_present.len;
                } else {
                    valid_raw = parent_bits;
                }
                popcs_.push_back({valid_raw, n_rows, cnt_base});
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
            validity_repack = (int32_t)repacks_.size();
            repacks_.push_back(rp);
            n_segments_ += n_groups;
            ab_present_ += (has_parent ? 0 : s_present.len) + (uint64_t)n_rows / 8;
            ab_repack_ += (uint64_t)n_rows / 4;
            cp.valid_bits = valid_raw;
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
            uint32_t chk_prev = 0;
            // the stream's bytes as the host sees them (for the scheduling hint below)
            const uint8_t* sp = nullptr;
            if (v2 && sr.present)
                sp = fm.base_for(data_start, si.data_length) + (sf.find(cid, &sr == &s_data ? S_DATA : (&sr == &s_length ? S_LENGTH : S_SECONDARY))->offset);
            // (every look is a cache miss in a file of hundreds of MB: have them all in flight before the loop needs them)
            if (sp && !compressed)
                for (uint32_t g = 0; g < ng; g++) __builtin_prefetch(sp + std::min(en[g].byte, sr.len));
            for (uint32_t g = 0; g < ng; g++) {
                Seg sg{};
                sg.in = sr.ptr;
                sg.in_len = sr.len;
                sg.out = dst;
                link(sg, g, chk_prev);
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
                if (sp) {
                    if (!compressed) {
                        long_runs = rle2_opens_with_long_runs(sp, sr.len, sg.start_byte);
                    } else if (!sr.chunks.empty()) {
                        // compressed file: the bytes are readable where the entry point lies in a chunk that was stored
                        // as is (incompressible integer streams usually are)
                        const size_t chi = (size_t)(std::upper_bound(sr.chunk_dst.begin(), sr.chunk_dst.end(), (uint64_t)sg.start_byte) - sr.chunk_dst.begin()) - 1;
                        if (chi < sr.chunks.size() && sr.chunks[chi].original)
                            long_runs = rle2_opens_with_long_runs(sp + sr.chunks[chi].src_off, sr.chunks[chi].src_len,
                                                                  (uint32_t)(sg.start_byte - sr.chunk_dst[chi]));
                        else
                            // not readable here: decimal scales are constant runs in practice, and so is any stream that
                            // spends less than a bit per value (sr.len is exact for Snappy and sized Zstandard frames,
                            // an upper bound otherwise)
                            long_runs = okind == OUT_SCALE || (uint64_t)sr.len * 8 < (uint64_t)n_rows;
                    }
                }
                if (!long_runs) {
                    const uint32_t bound = (has_present && per_group_counts) ? rows_in_group(g) : sg.n_values;
                    const uint32_t span = (g + 1 < ng ? en[g + 1].byte : sr.len) - std::min(en[g].byte, sr.len);
                    assign_run_slots(sg, bound, span);
                }
                (long_runs ? int_big_segs_ : int_segs_).push_back(sg);
                static const uint32_t ow1[6] = {2, 4, 8, 4, 4, 1};
                // (long constant runs of decimal scales are compared, not written: only their stream bytes count)
                (long_runs ? ab_intbig_ : ab_int_) += (uint64_t)(sr.len / ng) + (long_runs && okind == OUT_SCALE ? 0ull : (uint64_t)rows_in_group(g) * ow1[okind]);
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
                uint32_t chk_prev = 0;
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
                    link(sg, g, chk_prev);
                    data_byte_segs_.push_back(sg);
                }
                n_segments_ += ng;
                if (has_present) add_spaced(dst, cp.values, 1, false);
                break;
            }
            case T_BOOLEAN: {
                uint32_t chk_prev_bool = 0;
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
                    link(sg, g, chk_prev_bool);
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
                    // no copy: the values buffer of a batch IS its byte range of the staged (or decompressed) DATA stream,
                    // where the reference reads the bytes into a new Vec (string.rs:135-140)
                    cp.str_data = s_data.ptr + str_d0;
                    cp.str_data_len = str_dn;
                    sc.data = cp.str_data;
                    sc.data_len = str_dn;
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
                        for (uint32_t t = 0; t < nt; t++) u8_tiles_.push_back(make_uint2((uint32_t)strcols_.size(), t));
                        ab_utf8_ += u8_n;
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
                uint32_t chk_prev_var = 0;
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
                    link(sg, g, chk_prev_var);
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
                if (tz_on && !oc.ts_decimal) {
                    // values the zone move pushes out of range become nulls: marked in the dense domain, carried to the
                    // rows like any dense value, and taken out of the validity when it is cut into batches
                    const uint64_t bm = ((uint64_t)n_rows + 31) / 32 * 4 + 16;
                    td.tznull = alloc(AR_ZERO, bm);
                    uint64_t null_rows = td.tznull;
                    if (has_present) {
                        null_rows = alloc(AR_ZERO, bm);
                        add_spaced(td.tznull, null_rows, 0, true);
                        repacks_[validity_repack].mask = null_rows;
                    } else {
                        cp.has_present = true;  // (for export: a validity buffer exists, attached where a batch has nulls)
                        cp.validity_stride = (uint32_t)align_up((bs + 7) / 8, 64);
                        cp.validity = alloc(AR_OUT, (uint64_t)cp.validity_stride * n_batches);
                        cp.nulls_idx = n_nulls_;
                        n_nulls_ += n_batches;
                        RepackDesc rp{};
                        rp.src = 0;
                        rp.mask = null_rows;
                        rp.dst = cp.validity;
                        rp.dst_stride = cp.validity_stride;
                        rp.n_rows = n_rows;
                        rp.batch_size = bs;
                        rp.n_batches = n_batches;
                        rp.null_out = (int32_t)cp.nulls_idx;
                        rp.batch0 = repack_work_;
                        repack_work_ += n_batches;
                        repacks_.push_back(rp);
                    }
                }
                ts_.push_back(td);
                ab_ts_ += (uint64_t)n_rows * (16 + tw);
                if (has_present) add_spaced(dst, cp.values, tw, true);
                break;
            }
            case T_STRUCT:
                // struct_decoder.rs:59-78: nothing of its own but the validity; the children decode under it (next level)
                break;
            case T_LIST: case T_MAP: {
                // list.rs:63-87 / map.rs:74-104: LENGTH -> element offsets, exactly the offsets of a direct string column
                // without bytes behind them; their total is the slot count of the children (next level)
                StrCol sc{};
                sc.n_rows = n_rows;
                sc.batch_size = bs;
                sc.n_batches = n_batches;
                sc.tiles_per_batch = std::max<uint32_t>(1, (std::min(bs, n_rows) + STR_TILE - 1) / STR_TILE);
                sc.n_tiles = sc.tiles_per_batch * n_batches;
                sc.colstripe = cs;
                sc.tile0 = str_tiles_;
                str_tiles_ += sc.n_tiles;
                sc.meta_slot = (uint32_t)strcols_.size();
                cp.str_slot = (int32_t)strcols_.size();
                cp.is_list = true;
                cp.offsets = alloc(AR_OUT, (uint64_t)n_batches * (bs + 1) * 4);
                sc.offsets = cp.offsets;
                sc.tile_base = alloc(AR_TMP, ((uint64_t)sc.n_tiles + 1) * 8);
                cp.batch_base_off = meta_bytes_;
                meta_bytes_ += ((uint64_t)n_batches + 1) * 8;
                sc.batch_base = cp.batch_base_off;
                const uint64_t rows_i32 = alloc(AR_TMP, (uint64_t)n_rows * 4 + 16);
                const uint64_t dense_i32 = has_present ? alloc(AR_TMP, (uint64_t)n_rows * 4 + 16) : rows_i32;
                sc.lens = rows_i32;
                sc.mode = 0;
                sc.data = 0;
                sc.data_len = 0xffffffffu;
                add_int_segs(s_length, dense_i32, false, 8, OUT_LEN31, ORCB_OFFSET_OVERFLOW, true);
                if (has_present) add_spaced(dense_i32, rows_i32, 4, false);
                ab_str_ += (uint64_t)n_rows * 12;
                strcols_.push_back(sc);
                break;
            }
            case T_UNION: {
                // union.rs:69-136: tags are a byte-RLE stream over the valid slots (null slots 0) = the type ids of a sparse
                // union; each child decodes every slot under the validity `tag == child` (next level)
                cp.values = alloc(AR_OUT, (uint64_t)n_rows);
                uint64_t dst = cp.values;
                if (has_present) dst = alloc(AR_TMP, (uint64_t)n_rows);
                Seg sg{};
                sg.in = s_data.ptr;
                sg.in_len = s_data.len;
                sg.out = dst;
                sg.colstripe = cs;
                sg.out_kind = OUT_I8;
                if (has_present) {
                    sg.cnt_idx = (int32_t)cnt_base;
                    sg.start_idx = (int32_t)cnt_base;
                } else {
                    sg.cnt_idx = -1;
                    sg.start_idx = -1;
                    sg.n_values = n_rows;
                }
                data_byte_segs_.push_back(sg);
                n_segments_ += 1;
                if (has_present) add_spaced(dst, cp.values, 1, false);
                UnionDesc ud{};
                ud.tags = cp.values;
                ud.valid = has_present ? valid_raw : 0;
                ud.n = n_rows;
                ud.n_children = (uint32_t)oc.child_ids.size();
                ud.stride = (uint32_t)align_up(((uint64_t)n_rows + 31) / 32 * 4 + 16, 16);
                ud.bits = alloc(AR_ZERO, (uint64_t)ud.stride * ud.n_children);
                ud.counts = n_nulls_;
                n_nulls_ += ud.n_children;
                cp.union_bits = ud.bits;
                cp.union_stride = ud.stride;
                cp.union_counts = ud.counts;
                unions_.push_back(ud);
                break;
            }
            default: fail(ORCB_NOT_IMPLEMENTED, "unsupported column type on the device path");
        }
        colstripes_.push_back(cp);
    }
}

}  // namespace orcb
