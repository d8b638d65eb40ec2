// Arrow C Data / C Device Data export of decoded batches (views into the arenas, kept alive by shared ownership).
#include "job_internal.h"

namespace orcb {

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
    ho->out = (uint8_t*)pinned_get(out_bytes, &ho->out_cap);
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
        ho->heap = (uint8_t*)pinned_get(heap_used, &ho->heap_cap);
        CUDA_OK(cudaMemcpyAsync(ho->heap, base_[AR_HEAP], heap_used, cudaMemcpyDeviceToHost, stream_));
    }
    // direct strings alias the staged / decompressed DATA streams: bring those byte ranges back as well
    uint64_t strs_bytes = 0;
    for (auto& cp : colstripes_) {
        if (cp.str_slot < 0 || strcols_[cp.str_slot].mode != 0) continue;
        cp.str_host_off = strs_bytes;
        strs_bytes += (cp.str_data_len + 63) / 64 * 64;
    }
    if (strs_bytes) {
        ho->strs = (uint8_t*)pinned_get(strs_bytes, &ho->strs_cap);
        for (auto& cp : colstripes_) {
            if (cp.str_slot < 0 || strcols_[cp.str_slot].mode != 0 || !cp.str_data_len) continue;
            CUDA_OK(cudaMemcpyAsync(ho->strs + cp.str_host_off, (const void*)(uintptr_t)reloc(cp.str_data), cp.str_data_len,
                                    cudaMemcpyDeviceToHost, stream_));
        }
    }
    CUDA_OK(cudaStreamSynchronize(stream_));
    host_out_ = ho;
}

static void add_children(ArrowArray* out, size_t n) {
    auto* p = (ArrayPriv*)out->private_data;
    p->child_store.resize(n);
    p->child_ptrs.resize(n);
    for (size_t i = 0; i < n; i++) {
        p->child_ptrs[i] = &p->child_store[i];
        p->child_store[i].release = nullptr;
    }
    out->n_children = (int64_t)n;
    out->children = p->child_ptrs.data();
}

static void build_batch(const Job* /*unused*/, ArrowArray* out, int64_t rows, size_t ncols) {
    init_array(out, rows, 0, 1);
    add_children(out, ncols);
}

// One column of internal batch `b`, slots [off, off + len) of it, as an Arrow array.  Root columns of a batch carry the
// view's offset; the children of nested columns are exported whole (offset 0, every slot) - a struct's or union's
// offset applies to its children, a list's offsets address its child absolutely - so nothing is copied or rebased.
void Job::export_node(uint32_t csi, int64_t off, int64_t len, bool top, bool device, ArrowArray* a) {
    (void)top;
    const ColStripePlan& cp = colstripes_[csi];
    const OutColumn& oc = cols_[cp.col];
    if (!device) ensure_host_output();
    else if (!finished_) finish();
    const uint32_t bs = opt_.batch_size;
    const uint32_t b = view_mode_ ? 0u : (uint32_t)(off / bs);  // internal batch (views: one per stripe)
    const int64_t voff = view_mode_ ? off : 0;
    const uint32_t* nulls = (const uint32_t*)(h_meta_ + o_nulls_);
    const uint64_t* ptrs = (const uint64_t*)(h_meta_ + o_ptrs_);
    const uint64_t omask = (1ull << 60) - 1;
    const uint8_t* out_base = device ? base_[AR_OUT] : host_out_->out;
    std::shared_ptr<void> keep = device ? dev_keepalive_ : std::shared_ptr<void>(host_out_);
    static const int32_t zero_offsets[4] = {0, 0, 0, 0};
    const bool is_str = oc.kind == T_STRING || oc.kind == T_VARCHAR || oc.kind == T_CHAR || oc.kind == T_BINARY;
    const bool nested = !oc.child_ids.empty();
    // ---- validity: attached only when the slots exported hold a null (derive_present_vec, mod.rs:247-251)
    int64_t nc = 0;
    const void* vbuf = nullptr;
    if (cp.has_present && cp.n_rows) {
        const uint8_t* bits = out_base + (cp.validity & omask) + (uint64_t)b * cp.validity_stride;
        if (view_mode_) {
            if (device) {
                uint64_t any = 0;
                for (uint32_t k = 0; k < cp.n_batches; k++) any += nulls[cp.nulls_idx + k];
                nc = any ? -1 : 0;  // not counted for a view in device memory
            } else {
                int64_t valid = 0;
                for (int64_t r = voff; r < voff + len; r++) valid += (bits[r >> 3] >> (r & 7)) & 1;
                nc = len - valid;
            }
        } else {
            nc = nulls[cp.nulls_idx + b];
        }
        if (nc) vbuf = bits;
    }
    const size_t nbuf = oc.kind == T_STRUCT || oc.kind == T_UNION ? 1 : (is_str ? 3 : 2);
    init_array(a, len, nc, nbuf);
    a->offset = voff;
    auto* ap = (ArrayPriv*)a->private_data;
    ap->keep1 = keep;
    ap->buffers[0] = vbuf;
    if (cp.n_rows == 0) {
        // nothing was decoded: empty array (offsets need one readable entry)
        for (size_t k = 1; k < nbuf; k++) ap->buffers[k] = device ? (const void*)base_[AR_ZERO] : (const void*)zero_offsets;
        if (oc.kind == T_UNION) ap->buffers[0] = device ? (const void*)base_[AR_ZERO] : (const void*)zero_offsets;
    } else if (is_str) {
        const int64_t* bb = (const int64_t*)(h_meta_ + o_bbase_ + cp.batch_base_off);
        ap->buffers[1] = out_base + (cp.offsets & omask) + (uint64_t)b * (bs + 1) * 4;
        const uint8_t* dbase;
        if (strcols_[cp.str_slot].mode == 1) {
            const uint64_t p = ptrs[cp.str_slot];
            if (device) dbase = (const uint8_t*)(uintptr_t)p;
            else dbase = host_out_->heap ? host_out_->heap + (p - (uint64_t)(uintptr_t)base_[AR_HEAP]) : host_out_->out;
        } else {
            if (device) dbase = (const uint8_t*)(uintptr_t)reloc(cp.str_data);
            else dbase = host_out_->strs ? host_out_->strs + cp.str_host_off : host_out_->out;
        }
        ap->buffers[2] = dbase + bb[b];
    } else if (cp.is_list) {
        ap->buffers[1] = out_base + (cp.offsets & omask) + (uint64_t)b * (bs + 1) * 4;
    } else if (oc.kind == T_BOOLEAN) {
        ap->buffers[1] = out_base + (cp.values & omask) + (uint64_t)b * cp.values_stride;
    } else if (oc.kind == T_UNION) {
        ap->buffers[0] = out_base + (cp.values & omask);  // type ids; a union has no validity of its own
        a->null_count = 0;
    } else if (oc.kind != T_STRUCT) {
        ap->buffers[1] = out_base + (cp.values & omask) + (uint64_t)b * bs * oc.width;
    }
    if (!nested) return;
    // ---- children: decoded by the next-level job, in the order this level listed them
    if (!next_level_) fail(ORCB_UNEXPECTED, "nested column without its children's job");
    uint32_t kid0 = 0;
    for (uint32_t c = task_first_cs_[cp.task]; c < csi; c++) kid0 += (uint32_t)colstripes_[c].kids.size();
    const uint32_t child_cs0 = next_level_->task_first_cs_[cp.task] + kid0;
    auto child = [&](size_t i, ArrowArray* out) {
        const ColStripePlan& cc = next_level_->colstripes_[child_cs0 + i];
        next_level_->export_node(child_cs0 + (uint32_t)i, 0, cc.n_rows, false, device, out);
    };
    if (oc.kind == T_MAP) {
        // Map = List<entries: Struct<keys, values>> (map.rs:88-100); the entries struct has no nulls
        add_children(a, 1);
        ArrowArray* e = &ap->child_store[0];
        const ColStripePlan& kc = next_level_->colstripes_[child_cs0];
        init_array(e, kc.n_rows, 0, 1);
        ((ArrayPriv*)e->private_data)->keep1 = keep;
        add_children(e, 2);
        auto* ep = (ArrayPriv*)e->private_data;
        child(0, &ep->child_store[0]);
        child(1, &ep->child_store[1]);
    } else {
        add_children(a, oc.child_ids.size());
        for (size_t i = 0; i < oc.child_ids.size(); i++) child(i, &ap->child_store[i]);
    }
}

void Job::export_batch(uint64_t i, ArrowArray* out) {
    if (i >= batch_task_.size()) fail(ORCB_INVALID_ARGUMENT, "batch index out of range");
    ensure_host_output();
    const uint32_t t = batch_task_[i], b = batch_idx_[i];
    const uint32_t cs0 = task_first_cs_[t];
    // (the stripe's row count also sizes the batches of an empty projection, mod.rs:538-549)
    const uint32_t rows = batch_rows_[i];
    build_batch(this, out, rows, cols_.size());
    auto* tp = (ArrayPriv*)out->private_data;
    tp->keep1 = host_out_;
    for (size_t c = 0; c < cols_.size(); c++) {
        const ColStripePlan& cp = colstripes_[cs0 + c];
        // views: the column may hold a row-group window only
        const int64_t off = view_mode_ ? (int64_t)batch_row0_[i] - (int64_t)cp.row_base : (int64_t)b * opt_.batch_size;
        export_node(cs0 + (uint32_t)c, off, rows, true, false, &tp->child_store[c]);
    }
}

void Job::export_batch_device(uint64_t i, ArrowDeviceArray* out) {
    if (i >= batch_task_.size()) fail(ORCB_INVALID_ARGUMENT, "batch index out of range");
    if (!finished_) finish();
    const uint32_t t = batch_task_[i], b = batch_idx_[i];
    const uint32_t cs0 = task_first_cs_[t];
    const uint32_t rows = batch_rows_[i];
    memset(out, 0, sizeof(*out));
    build_batch(this, &out->array, rows, cols_.size());
    out->device_id = opt_.device;
    out->device_type = ARROW_DEVICE_CUDA;
    out->sync_event = nullptr;  // finish() already synchronised the stream
    auto* tp = (ArrayPriv*)out->array.private_data;
    tp->keep1 = dev_keepalive_;
    for (size_t c = 0; c < cols_.size(); c++) {
        const ColStripePlan& cp = colstripes_[cs0 + c];
        const int64_t off = view_mode_ ? (int64_t)batch_row0_[i] - (int64_t)cp.row_base : (int64_t)b * opt_.batch_size;
        export_node(cs0 + (uint32_t)c, off, rows, true, true, &tp->child_store[c]);
    }
}

}  // namespace orcb
