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
                dbase = host_out_->strs ? host_out_->strs + cp.str_host_off : host_out_->out;
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
                                                                   : (const uint8_t*)(uintptr_t)reloc(cp.str_data);
            ap->buffers[2] = dbase + bb[b];
        } else if (oc.kind == T_BOOLEAN) {
            ap->buffers[1] = base_[AR_OUT] + (cp.values & omask) + (uint64_t)b * cp.values_stride;
        } else {
            ap->buffers[1] = base_[AR_OUT] + (cp.values & omask) + (uint64_t)b * bs * oc.width;
        }
    }
}

}  // namespace orcb
