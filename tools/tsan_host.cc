// ThreadSanitizer driver for the host threads of the library that run without a GPU (tools/tsan_host.sh): the parallel
// file read of orcb_open_path, orcb_reader_next_async (a library thread per request, completion callback, the join in the
// next request and in orcb_reader_free), several readers on clones of one file from several threads, and a callback-fed
// file (orcb_open_callbacks) planned from two threads at once.  Without a device every decode ends in a status
// (ORCB_CUDA); the threading around it is what is checked.
//   tsan_host <big.orc (> 16 MiB, so the read is split)> <small.orc>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../include/orc_b200.h"

static std::atomic<int> g_done{0}, g_fail{0};
static void on_batch(void* ctx, int status, int eos, const char* error) {
    (void)eos;
    (void)error;
    if (status) g_fail++;
    g_done++;
    *(std::atomic<int>*)ctx = 1;
}

struct Src {
    std::vector<uint8_t> bytes;
    std::atomic<uint64_t> calls{0};
};
static int read_at(void* ctx, uint64_t off, uint64_t len, uint8_t* dst) {
    Src* s = (Src*)ctx;
    if (off > s->bytes.size() || len > s->bytes.size() - off) return 1;
    memcpy(dst, s->bytes.data() + off, len);
    s->calls++;
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    OrcbReadOptions opt;
    memset(&opt, 0, sizeof opt);
    opt.use_row_index = 1;

    // 1. parallel read + async requests, from several threads on their own handles
    std::vector<std::thread> th;
    for (int t = 0; t < 4; t++)
        th.emplace_back([&, t] {
            OrcbFile* f = nullptr;
            if (orcb_open_path(argv[1 + (t & 1)], &f) != 0) { g_fail++; return; }
            OrcbFile* c = nullptr;
            orcb_file_clone(f, &c);
            for (OrcbFile* h : {f, c}) {
                OrcbReader* r = nullptr;
                if (orcb_reader_new(h, &opt, &r) != 0) continue;
                for (int k = 0; k < 3; k++) {
                    struct ArrowArray a;
                    std::atomic<int> flag{0};
                    if (orcb_reader_next_async(r, &a, on_batch, &flag) != 0) break;
                    if (k == 1) while (!flag.load()) std::this_thread::yield();  // sometimes wait, sometimes let the next call join
                    else while (!flag.load()) std::this_thread::yield();
                }
                orcb_reader_free(r);
            }
            orcb_file_free(c);
            orcb_file_free(f);
        });
    for (auto& x : th) x.join();
    th.clear();

    // 2. one callback-fed file, planned by two threads at once (stripe reads go through the handle's lock)
    Src src;
    {
        FILE* fp = fopen(argv[2], "rb");
        fseek(fp, 0, SEEK_END);
        long n = ftell(fp);
        fseek(fp, 0, SEEK_SET);
        src.bytes.resize((size_t)n);
        if (fread(src.bytes.data(), 1, (size_t)n, fp) != (size_t)n) return 2;
        fclose(fp);
    }
    OrcbFile* cf = nullptr;
    if (orcb_open_callbacks(src.bytes.size(), read_at, &src, &cf) != 0) return 3;
    for (int t = 0; t < 2; t++)
        th.emplace_back([&] {
            for (int k = 0; k < 5; k++) {
                OrcbJob* j = nullptr;
                OrcbFile* files[1] = {cf};
                if (orcb_job_new(files, 1, &opt, &j) == 0) {
                    orcb_job_plan(j);
                    orcb_job_free(j);
                }
            }
        });
    for (auto& x : th) x.join();
    uint64_t io[2];
    orcb_file_io_stats(cf, io);
    orcb_file_free(cf);
    printf("tsan_host: %d async completions (%d with a status), %llu callback reads\n", g_done.load(), g_fail.load(),
           (unsigned long long)io[0]);
    return 0;
}
