// Job life cycle and execution: arenas, staging (H2D), the kernel launch sequence on two streams, completion and
// error mapping, statistics.  Planning is in plan.cc, Arrow export in export.cc.
#include <chrono>
#include <cstdio>
#include <map>
#include <mutex>

#include "job_internal.h"

namespace orcb {

// pinned host buffers, kept between jobs (job_internal.h).  Sizes are rounded up to a power of two below 1 MiB and to a
// multiple of 1/8 of the next lower power of two above, so a request finds a spare buffer at most 12.5 % larger
namespace {
struct PinnedCache {
    std::mutex mu;
    std::multimap<size_t, void*> spare;
    size_t held = 0, limit = 4096ull << 20;
    PinnedCache() {
        if (const char* e = getenv("ORCB_PINNED_CACHE_MB")) limit = (size_t)strtoull(e, nullptr, 10) << 20;
    }
    static size_t size_class(size_t want) {
        size_t c = 4096;
        while (c < want && c < (1u << 20)) c <<= 1;
        if (c >= want) return c;
        size_t p2 = 1u << 20;
        while (p2 * 2 <= want) p2 <<= 1;
        const size_t step = p2 / 8;
        return (want + step - 1) / step * step;
    }
    void* get(size_t want, size_t* cap) {
        const size_t c = size_class(want);
        *cap = c;
        {
            std::lock_guard<std::mutex> lock(mu);
            auto it = spare.find(c);
            if (it != spare.end()) {
                void* p = it->second;
                spare.erase(it);
                held -= c;
                return p;
            }
        }
        void* p = nullptr;
        if (cudaHostAlloc(&p, c, cudaHostAllocDefault) != cudaSuccess) {
            // host memory is short: give back what is held and try once more
            cudaGetLastError();
            drop_all();
            CUDA_OK(cudaHostAlloc(&p, c, cudaHostAllocDefault));
        }
        return p;
    }
    void put(void* p, size_t cap) {
        {
            std::lock_guard<std::mutex> lock(mu);
            if (held + cap <= limit) {
                spare.emplace(cap, p);
                held += cap;
                return;
            }
        }
        cudaFreeHost(p);
    }
    void drop_all() {
        std::multimap<size_t, void*> old;
        {
            std::lock_guard<std::mutex> lock(mu);
            old.swap(spare);
            held = 0;
        }
        for (auto& kv : old) cudaFreeHost(kv.second);
    }
};
PinnedCache& pinned_cache() {
    static PinnedCache* c = new PinnedCache();  // never destroyed: buffers may come back while the process exits
    return *c;
}
}  // namespace

void* pinned_get(size_t bytes, size_t* capacity) { return pinned_cache().get(bytes, capacity); }
void pinned_put(void* p, size_t capacity) { pinned_cache().put(p, capacity); }

bool DeviceArenas::use_pool(int device) {
    static std::mutex mu;
    static std::map<int, bool> ready;
    std::lock_guard<std::mutex> lock(mu);
    auto it = ready.find(device);
    if (it != ready.end()) return it->second;
    bool ok = false;
    const char* e = getenv("ORCB_SYNC_ALLOC");
    if (!(e && e[0] == '1')) {
        int supported = 0;
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetAttribute(&supported, cudaDevAttrMemoryPoolsSupported, device) == cudaSuccess && supported &&
            cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;  // freed memory stays in the pool instead of going back to the driver at every sync
            ok = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep) == cudaSuccess;
        }
        cudaGetLastError();
    }
    ready[device] = ok;
    return ok;
}

void* DeviceArenas::alloc(size_t bytes, cudaStream_t st) {
    void* p = nullptr;
    if (pooled) CUDA_OK(cudaMallocAsync(&p, bytes, st));
    else CUDA_OK(cudaMalloc(&p, bytes));
    ptrs.push_back(p);
    return p;
}

DeviceArenas::~DeviceArenas() {
    int cur = 0;
    const bool switched = cudaGetDevice(&cur) == cudaSuccess && cur != device && cudaSetDevice(device) == cudaSuccess;
    for (void* p : ptrs) {
        if (!p) continue;
        if (pooled) cudaFreeAsync(p, nullptr);
        else cudaFree(p);
    }
    if (switched) cudaSetDevice(cur);
}


Job::Job(std::vector<StripeTask> tasks, const ReadOptions& opt) : tasks_(std::move(tasks)), opt_(opt), orig_opt_(opt) {
    if (tasks_.empty()) return;
    if (tasks_[0].roots.empty()) {
        cols_ = project_columns(*tasks_[0].file, opt_);
    } else {
        // a later nesting level: the columns are the children the level before handed over (same ids for every task)
        for (auto& r : tasks_[0].roots) cols_.push_back(column_info(*tasks_[0].file, r.col_id, "", opt_, -1));
        for (auto& t : tasks_)
            if (t.roots.size() != cols_.size()) fail(ORCB_UNEXPECTED, "stripes of one job disagree on their nested columns");
    }
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
    next_level_.reset();  // the levels below run on this job's stream: they go first
    // pooled memory does not wait for pending work when it is freed (cudaFree did): a job dropped between launch()
    // and finish(), or while an end-to-end restage is in flight, drains its streams first
    if (staged_) {
        if (stream_) cudaStreamSynchronize(stream_);
        if (aux_stream_) cudaStreamSynchronize(aux_stream_);
    }
    for (auto& k : kstats_) {
        if (k.e0) cudaEventDestroy(k.e0);
        if (k.e1) cudaEventDestroy(k.e1);
    }
    if (h_meta_) pinned_put(h_meta_, h_meta_cap_);
    if (done_) cudaEventDestroy(done_);
    if (ev_fork_) cudaEventDestroy(ev_fork_);
    if (ev_join_) cudaEventDestroy(ev_join_);
    if (ev_str_) cudaEventDestroy(ev_str_);
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
    if (a == AR_ABS) return tagged & ((1ull << 60) - 1);
    return (uint64_t)(uintptr_t)base_[a] + (tagged & ((1ull << 60) - 1));
}

// ------------------------------------------------------------------------------------------------
// staging: allocate, relocate descriptor pointers, upload
// ------------------------------------------------------------------------------------------------
void Job::stage() {
    if (!planned_) plan();
    if (staged_) return;
    const bool timing = getenv("ORCB_READER_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto lap = [&](std::chrono::steady_clock::time_point& t) {
        const auto n = now();
        const double d = std::chrono::duration<double, std::milli>(n - t).count();
        t = n;
        return d;
    };
    auto tl = now();
    double t_streams = 0, t_alloc = 0, t_host = 0, t_reloc = 0, t_copy = 0;
    CUDA_OK(cudaSetDevice(opt_.device));
    if (opt_.own_stream) {
        CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        own_stream_ = true;
    } else {
        stream_ = opt_.stream;
    }
    CUDA_OK(cudaEventCreateWithFlags(&done_, cudaEventDisableTiming));
    {
        // The header walk is a latency-bound chain at the head of the integer path: on a stream of higher priority its
        // few warps (and the run decode that follows) are placed as soon as they are launched, whatever else is queued:
        // 6.98 -> 6.68 ms at SF10 (ORCB_AUX_PRIO=0: no priority).  Tried and dropped: the walk alone at high priority
        // (7.14 ms), the decimal varint decoder on a third, low-priority stream (7.10 vs 6.91 ms).
        static const bool prio = !(getenv("ORCB_AUX_PRIO") && getenv("ORCB_AUX_PRIO")[0] == '0');
        int least = 0, greatest = 0;
        if (prio && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess && greatest != least)
            CUDA_OK(cudaStreamCreateWithPriority(&aux_stream_, cudaStreamNonBlocking, greatest));
        else
            CUDA_OK(cudaStreamCreateWithFlags(&aux_stream_, cudaStreamNonBlocking));
    }
    CUDA_OK(cudaEventCreateWithFlags(&ev_fork_, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_join_, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&ev_str_, cudaEventDisableTiming));
    t_streams = lap(tl);
    auto arenas = std::make_shared<DeviceArenas>();
    arenas->device = opt_.device;
    arenas->pooled = DeviceArenas::use_pool(opt_.device);
    for (int a = 1; a < 7; a++) base_[a] = (uint8_t*)arenas->alloc(std::max<uint64_t>(size_[a], 256), stream_);
    d_desc_ = (uint8_t*)arenas->alloc(std::max<uint64_t>(desc_bytes_, 256), stream_);
    d_state_ = (uint8_t*)arenas->alloc(std::max<uint64_t>(state_bytes_, 256), stream_);
    d_meta_ = (uint8_t*)arenas->alloc(std::max<uint64_t>(meta_bytes_, 256), stream_);
    t_alloc = lap(tl);
    h_meta_ = (uint8_t*)pinned_get(std::max<uint64_t>(meta_bytes_, 256), &h_meta_cap_);
    dev_keepalive_ = arenas;
    t_host = lap(tl);

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
    for (auto* v : {&spaced_, &spaced_late_, &merge_spaced_})
        for (auto& d : *v) { R(d.src); R(d.dst); R(d.valid); }
    for (auto& d : popcs_) R(d.bits);
    for (auto& d : unions_) { R(d.tags); R(d.valid); R(d.bits); }
    for (auto& d : decfix_) { R(d.vals); R(d.scales); }
    for (auto& d : ts_) {
        R(d.secs); R(d.nanos); R(d.out); R(d.tznull);
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
    for (auto& d : repacks_) { R(d.src); R(d.mask); R(d.dst); }
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
    put(o_spm_, merge_spaced_.data(), merge_spaced_.size() * sizeof(SpacedDesc));
    put(o_popc_, popcs_.data(), popcs_.size() * sizeof(PopcDesc));
    put(o_union_, unions_.data(), unions_.size() * sizeof(UnionDesc));
    put(o_chkpair_, chk_pairs_.data(), chk_pairs_.size() * sizeof(uint2));
    put(o_dec_, decfix_.data(), decfix_.size() * sizeof(DecFixDesc));
    put(o_ts_, ts_.data(), ts_.size() * sizeof(TsDesc));
    put(o_str_, strcols_.data(), strcols_.size() * sizeof(StrCol));
    put(o_rep_, repacks_.data(), repacks_.size() * sizeof(RepackDesc));
    put(o_chunk_, chunks_.data(), chunks_.size() * sizeof(ChunkDesc));
    put(o_tz_, tz_blob_.data(), tz_blob_.size());
    t_reloc = lap(tl);
    CUDA_OK(cudaMemcpyAsync(d_desc_, desc_blob_.data(), desc_blob_.size(), cudaMemcpyHostToDevice, stream_));
    for (auto& sc : stage_copies_)
        CUDA_OK(cudaMemcpyAsync(base_[AR_IN] + sc.dst_off, sc.src, sc.bytes, cudaMemcpyHostToDevice, stream_));
    t_copy = lap(tl);
    if (timing)
        fprintf(stderr, "orcb stage: streams/events %.2f ms, device memory %.2f ms, pinned meta %.2f ms, descriptors %.2f ms (%zu KiB), "
                        "copies issued %.2f ms\n", t_streams, t_alloc, t_host, t_reloc, desc_blob_.size() >> 10, t_copy);
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
    SegCheck* segchk = n_chk_ ? (SegCheck*)(uintptr_t)reloc(chk_table_) : nullptr;
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
        run("k_decompress", ab_decomp_, N(chunks_), (int)__builtin_popcount(bits_codecs_) + (n_snappy_chunks_ ? 1 : 0) + (N(chunks_) > n_bits_chunks_ + n_snappy_chunks_ ? 1 : 0), [&] { return launch_decompress((ChunkDesc*)(d_desc_ + o_chunk_), N(chunks_), n_bits_chunks_, bits_codecs_, n_snappy_chunks_, err, (uint32_t*)(d_meta_ + o_clens_), (uint32_t*)(d_state_ + o_nblocks_) + 3, (uint32_t*)(d_meta_ + o_retry_), st); });
    if (N(present_byte_segs_)) {
        run("k_byte_rle(present)", ab_present_, N(present_byte_segs_), 1, [&] { return launch_byte_rle((Seg*)(d_desc_ + o_pbyte_), N(present_byte_segs_), cnt, dstart, err, segchk, st); });
        run("k_bits(present)", ab_present_, N(present_bit_segs_), 1, [&] { return launch_bits((BitSeg*)(d_desc_ + o_pbit_), N(present_bit_segs_), cnt, dstart, st); });
    }
    // validity handed down by a parent column: own PRESENT bits scattered into the parent's valid slots, then counted
    if (N(merge_spaced_))
        run("k_spaced(validity)", 0, N(merge_spaced_), 1, [&] { return launch_spaced((SpacedDesc*)(d_desc_ + o_spm_), N(merge_spaced_), dstart, st); });
    if (N(popcs_))
        run("k_popc", 0, N(popcs_), 1, [&] { return launch_popc((PopcDesc*)(d_desc_ + o_popc_), N(popcs_), cnt, st); });
    if (N(scans_))
        run("k_seg_scan", 0, N(scans_), 1, [&] { return launch_seg_scan((ScanDesc*)(d_desc_ + o_scan_), N(scans_), cnt, dstart, st); });
    if (N(data_byte_segs_))
        run("k_byte_rle", ab_byte_, N(data_byte_segs_), 1, [&] { return launch_byte_rle((Seg*)(d_desc_ + o_dbyte_), N(data_byte_segs_), cnt, dstart, err, segchk, st); });
    if (N(data_bit_segs_))
        run("k_bits", ab_bits_, N(data_bit_segs_), 1, [&] { return launch_bits((BitSeg*)(d_desc_ + o_dbit_), N(data_bit_segs_), cnt, dstart, st); });
    // fork: the header-walk pre-pass is a long dependent chain on few warps, so it runs beside the
    // bandwidth-heavy kernels on a second stream and joins before the epilogues
    // ORCB_SERIAL=1 keeps everything on one stream (clean per-kernel timings when profiling)
    const bool serial_env = getenv("ORCB_SERIAL") != nullptr;  // read at every launch: bench.py times one serial pass
    const bool forked = N(int_segs_) > 0;
    // two phases of the short-run integer path when string columns wait for part of it and nothing sits between the
    // integer decode and the string kernels (no dense -> rows expansion, no unions); ORCB_SPLIT_INT=0 disables
    static const bool split_env = !(getenv("ORCB_SPLIT_INT") && getenv("ORCB_SPLIT_INT")[0] == '0');
    const bool split_int = forked && !serial_env && split_env && n_str_int_segs_ > 0 && n_str_int_segs_ < N(int_segs_) &&
                           N(strcols_) > 0 && N(spaced_) == 0 && N(unions_) == 0;
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
        Seg* isegs = (Seg*)(d_desc_ + o_int_);
        CoopRec* coopq = (CoopRec*)(uintptr_t)reloc(coop_q_);
        uint32_t* slow = (uint32_t*)(uintptr_t)reloc(slow_list_);
        if (split_int) {
            // phase 1: the segments the string kernels wait for; phase 2 (the rest) appends to the same block pool and
            // starts where the snapshot of the three counters says
            const uint32_t na = n_str_int_segs_, nb = N(int_segs_) - n_str_int_segs_;
            uint32_t* snap = nblk + 8;
            run("k_rle_index(strings)", 0, na, 1, [&] { return launch_rle_index(isegs, na, cnt, rtab, brec, nblk, pool_blocks_, coopq, nblk + 2, coop_cap_, err, segchk, aux); });
            run("k_int_rle(strings)", 0, pool_blocks_, 3, [&] { return launch_int_rle(isegs, brec, nblk, pool_blocks_, rtab, cnt, dstart, err, mis, slow, nblk + 1, coopq, nblk + 2, coop_cap_, nullptr, aux); });
            CUDA_OK(cudaEventRecord(ev_str_, aux));
            chk(launch_snapshot3(nblk, snap, aux), "k_snapshot3");
            launches += 1;
            run("k_rle_index", 0, nb, 1, [&] { return launch_rle_index(isegs + na, nb, cnt, rtab, brec, nblk, pool_blocks_, coopq, nblk + 2, coop_cap_, err, segchk, aux); });
            run("k_int_rle(+general,+coop_runs)", ab_int_, pool_blocks_, 3, [&] { return launch_int_rle(isegs + na, brec, nblk, pool_blocks_, rtab, cnt, dstart, err, mis, slow, nblk + 1, coopq, nblk + 2, coop_cap_, snap, aux); });
        } else {
            run("k_rle_index", 0, N(int_segs_), 1, [&] { return launch_rle_index(isegs, N(int_segs_), cnt, rtab, brec, nblk, pool_blocks_, coopq, nblk + 2, coop_cap_, err, segchk, aux); });
            run("k_int_rle(+general,+coop_runs)", ab_int_, pool_blocks_, 3, [&] { return launch_int_rle(isegs, brec, nblk, pool_blocks_, rtab, cnt, dstart, err, mis, slow, nblk + 1, coopq, nblk + 2, coop_cap_, nullptr, aux); });
        }
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
            run("k_int_rle_coop", ab_intbig_, N(int_big_segs_), 1, [&] { return launch_int_rle_coop((Seg*)(d_desc_ + o_intbig_), N(int_big_segs_), cnt, dstart, err, mis, 0, segchk, st); });
        if (*o == 'v' && N(var_segs_))
            run("k_varint128", ab_var_, N(var_segs_), 1, [&] { return launch_varint128((Seg*)(d_desc_ + o_var_), N(var_segs_), cnt, dstart, err, segchk, st); });
        if (*o == 'k' && N(copy_tiles_))
            run("k_copy", ab_copy_, N(copy_tiles_), 1, [&] { return launch_copy((CopyDesc*)(d_desc_ + o_copy_), (uint2*)(d_desc_ + o_ctile_), N(copy_tiles_), cnt, err, (StrCol*)(d_desc_ + o_str_), st); });
    }
    auto strings = [&] {
        run("k_strings(5 kernels)", ab_str_, str_tiles_, 5, [&] { return launch_strings((StrCol*)(d_desc_ + o_str_), N(strcols_), str_tiles_, err, dstate, (uint64_t)(uintptr_t)base_[AR_HEAP], heap_cap, ptrs, st); });
    };
    if (split_int) {
        CUDA_OK(cudaStreamWaitEvent(st, ev_str_, 0));
        strings();
    }
    if (forked && !serial_env) CUDA_OK(cudaStreamWaitEvent(st, ev_join_, 0));
    // strings first: they close the longer chain
    if (N(spaced_))
        run("k_spaced", ab_spaced_, N(spaced_), 1, [&] { return launch_spaced((SpacedDesc*)(d_desc_ + o_sp_), N(spaced_), dstart, st); });
    if (N(unions_))
        run("k_union_valid", 0, N(unions_), 1, [&] { return launch_union_valid((UnionDesc*)(d_desc_ + o_union_), N(unions_), nulls, st); });
    if (N(strcols_) && !split_int) strings();
    if (N(chk_pairs_))  // every segment has left its start and end position: do they join up?
        run("k_seg_check", 0, N(chk_pairs_), 1, [&] { return launch_seg_check((uint2*)(d_desc_ + o_chkpair_), N(chk_pairs_), segchk, (uint32_t*)(d_meta_ + o_retry_), st); });
    if (N(decfix_) && N(int_big_segs_))  // scales that were only compared so far are written where one of them differed
        run("k_int_rle_coop(scales)", 0, N(int_big_segs_), 1, [&] { return launch_int_rle_coop((Seg*)(d_desc_ + o_intbig_), N(int_big_segs_), cnt, dstart, err, mis, 1, nullptr, st); });
    if (N(decfix_))
        run("k_decimal_fix", ab_dec_, N(decfix_), 1, [&] { return launch_decimal_fix((DecFixDesc*)(d_desc_ + o_dec_), N(decfix_), cnt, mis, st); });
    if (N(ts_))
        run("k_timestamp", ab_ts_, N(ts_), 1, [&] { return launch_timestamp((TsDesc*)(d_desc_ + o_ts_), N(ts_), cnt, err, st); });
    if (N(spaced_late_))
        run("k_spaced(late)", ab_spaced_, N(spaced_late_), 1, [&] { return launch_spaced((SpacedDesc*)(d_desc_ + o_sp2_), N(spaced_late_), dstart, st); });
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
    const uint32_t retry = *(const uint32_t*)(h_meta_ + o_retry_);
    if ((retry & 2u) && !(retry & 1u)) throw IndexRetry{};
    if (retry & 1u) {
        // some chunk did not have the size the layout assumed: remember every size the device found and ask for a re-plan
        const uint32_t* cl = (const uint32_t*)(h_meta_ + o_clens_);
        for (size_t i = 0; i < chunk_keys_.size(); i++) {
            if (cl[i] == 0xffffffffu) continue;
            ChunkSizeCache& c = *chunk_keys_[i].first->chunk_sizes;
            std::lock_guard<std::mutex> lock(c.mu);
            c.size[chunk_keys_[i].second] = cl[i];
        }
        throw LayoutRetry{};
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
    aliased_bytes_ = 0;
    const uint32_t* nulls = (const uint32_t*)(h_meta_ + o_nulls_);
    for (auto& cp : colstripes_) {
        const OutColumn& oc = cols_[cp.col];
        const uint32_t bs = opt_.batch_size;
        for (uint32_t b = 0; b < cp.n_batches; b++) {
            const uint32_t rows = std::min(bs, cp.n_rows - b * bs);
            if (cp.has_present && nulls[cp.nulls_idx + b]) output_bytes_ += (rows + 7) / 8;
            if (oc.kind == T_BOOLEAN) output_bytes_ += (rows + 7) / 8;
            else if (cp.is_list) output_bytes_ += 4ull * (rows + 1);
            else if (oc.kind == T_STRUCT) {}
            else if (oc.kind == T_UNION) output_bytes_ += rows;
            else if (cp.str_slot >= 0) {
                const int64_t* bb = (const int64_t*)(h_meta_ + o_bbase_ + cp.batch_base_off);
                output_bytes_ += 4ull * (rows + 1) + (uint64_t)(bb[b + 1] - bb[b]);
                if (strcols_[cp.str_slot].mode == 0) aliased_bytes_ += (uint64_t)(bb[b + 1] - bb[b]);
            } else output_bytes_ += (uint64_t)rows * oc.width;
        }
    }
    finished_ = true;
    run_next_level();
}

// the roots of the next nesting level for task t: the children of this level's struct / list / map / union columns
std::vector<RootSpec> Job::next_roots(uint32_t t) const {
    std::vector<RootSpec> out;
    const uint32_t* nulls = (const uint32_t*)(h_meta_ + o_nulls_);
    const uint32_t cs0 = task_first_cs_[t];
    for (uint32_t c = 0; c < cols_.size(); c++) {
        const ColStripePlan& cp = colstripes_[cs0 + c];
        const int kind = cols_[c].kind;
        for (size_t i = 0; i < cp.kids.size(); i++) {
            RootSpec r;
            r.col_id = cp.kids[i];
            if (cp.n_rows == 0) {
                out.push_back(r);
                continue;
            }
            if (kind == T_LIST || kind == T_MAP) {
                const int64_t* bb = (const int64_t*)(h_meta_ + o_bbase_ + cp.batch_base_off);
                r.n_slots = (uint64_t)bb[cp.n_batches];
            } else if (kind == T_UNION) {
                r.n_slots = cp.n_rows;
                r.has_parent = true;
                r.parent_bits = reloc(cp.union_bits) + (uint64_t)i * cp.union_stride;
                r.parent_count = nulls[cp.union_counts + i];
            } else {  // struct
                r.n_slots = cp.n_rows;
                if (cp.has_present) {
                    uint64_t n_null = 0;
                    for (uint32_t b = 0; b < cp.n_batches; b++) n_null += nulls[cp.nulls_idx + b];
                    r.has_parent = true;
                    r.parent_bits = reloc(cp.valid_bits);
                    r.parent_count = cp.n_rows - n_null;
                }
            }
            out.push_back(r);
        }
    }
    return out;
}

void Job::run_next_level() {
    bool any = false;
    for (auto& cp : colstripes_) any |= !cp.kids.empty();
    if (!any) return;
    std::vector<StripeTask> tasks;
    for (uint32_t t = 0; t < tasks_.size(); t++) {
        StripeTask nt{tasks_[t].file, tasks_[t].stripe};
        nt.roots = next_roots(t);
        if (tasks_[t].file->stripes[tasks_[t].stripe].data_length) {
            // the level below decodes from the stripe bytes this level staged (this job outlives it)
            nt.has_staged = true;
            nt.staged_abs = reloc(task_in_off_[t]);
        }
        tasks.push_back(std::move(nt));
    }
    ReadOptions o = orig_opt_;
    o.stream = stream_;       // same stream: the children read this level's validity bitmaps
    o.own_stream = false;
    next_level_ = std::make_unique<Job>(std::move(tasks), o);
    next_level_->plan();
    next_level_->stage();
    next_level_->launch();
    next_level_->finish();    // (recursively runs the levels below)
    output_bytes_ += next_level_->output_bytes_;
    aliased_bytes_ += next_level_->aliased_bytes_;
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
    out->aliased_output_bytes = aliased_bytes_;
    out->n_waves = 1;
    for (int a = 1; a < 8; a++) out->device_bytes += size_[a];
    out->device_bytes += desc_bytes_ + state_bytes_ + meta_bytes_;
    out->n_segments = n_segments_;
    out->n_kernel_launches = n_launches_;
    out->n_batches = batch_task_.size();
    out->d2h_meta_bytes = meta_bytes_;
}

}  // namespace orcb
