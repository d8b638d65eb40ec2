// Shared by the translation units that implement Job (job.cc, plan.cc, export.cc, schema.cc).
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "job.h"
#include "kernels.h"

namespace orcb {

#define CUDA_OK(expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) fail(ORCB_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

static const int64_t ORC_EPOCH_UTC = 1420070400;  // array_decoder/timestamp.rs:51
static const uint64_t ARENA_PAD = 256;            // slack so word-granular kernel loads may overrun streams

// ------------------------------------------------------------------------------------------------
// device memory owned jointly by the job and by exported device batches
// ------------------------------------------------------------------------------------------------
// Device memory of one job.  Allocated from the device's stream-ordered pool (cudaMallocAsync) with the pool told to
// keep what is freed, so that a reader that builds one job per group of stripes does not pay cudaMalloc / cudaFree
// (and the device-wide synchronisation of cudaFree) for every group: 3 ms -> 0.3 ms per job.  As with any
// stream-ordered allocator, memory goes back to the pool when the last holder lets go (the job, or the last
// device-resident batch exported from it), and the holder's work on it must be complete by then.
// ORCB_SYNC_ALLOC=1 selects plain cudaMalloc / cudaFree.
struct DeviceArenas {
    int device = 0;
    bool pooled = false;
    std::vector<void*> ptrs;
    static bool use_pool(int device);
    void* alloc(size_t bytes, cudaStream_t st);
    ~DeviceArenas();
};

// Pinned host memory is expensive to get and to give back (cudaHostAlloc pins pages: 3-14 ms for a few KiB on a busy
// host, ~0.4 ms per MiB for the copies of a job's output), far more than copying and decoding the stripe it is for.
// Buffers are therefore kept when their holder lets go and handed to the next job that asks for about that size
// (job.cc).  At most ORCB_PINNED_CACHE_MB (default 4096) MiB are held; ORCB_PINNED_CACHE_MB=0 turns the cache off.
void* pinned_get(size_t bytes, size_t* capacity);
void pinned_put(void* p, size_t capacity);

struct HostOutput {
    uint8_t* out = nullptr;   // pinned copy of AR_OUT
    uint8_t* heap = nullptr;  // pinned copy of the used part of AR_HEAP
    uint8_t* strs = nullptr;  // pinned copy of the DATA byte ranges that direct string columns point into
    size_t out_cap = 0, heap_cap = 0, strs_cap = 0;
    ~HostOutput() {
        if (out) pinned_put(out, out_cap);
        if (heap) pinned_put(heap, heap_cap);
        if (strs) pinned_put(strs, strs_cap);
    }
};

}  // namespace orcb
