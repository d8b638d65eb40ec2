// Host-side fuzz driver for an AddressSanitizer / UBSan build of the library (tools/fuzz_host.sh builds and runs it).
// No GPU needed: it exercises what the host does with file bytes - tail and footer parsing, stripe footers, row
// indexes, the host section decoders (inflate, Snappy, LZ4, LZO, Zstandard; the last two are the code the device runs
// too, csrc/zstd_dec.h), schema export, the planner.  Every input lives in an exactly-sized heap block, so a read one
// byte outside it is reported.  A call may fail with a status; it may not crash, hang or trip a sanitizer.
//
//   fuzz_host <iterations-per-seed> <seed> file.orc... section.sec...
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../include/orc_b200.h"

static uint64_t rng_state = 88172645463325252ull;
static uint64_t rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}
static uint64_t below(uint64_t n) { return n ? rnd() % n : 0; }

static std::vector<uint8_t> slurp(const char* path) {
    std::vector<uint8_t> v;
    FILE* f = fopen(path, "rb");
    if (!f) return v;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    v.resize(n);
    if (n && fread(v.data(), 1, n, f) != (size_t)n) v.clear();
    fclose(f);
    return v;
}

static void mutate(std::vector<uint8_t>& b, uint64_t span, int it) {
    const size_t n = b.size();
    if (!n) return;
    const int k = 1 + (int)below(3);
    for (int i = 0; i < k; i++) {
        const size_t at = n - 1 - below(span < n ? span : n);
        switch (below(4)) {
            case 0: b[at] = (uint8_t)rnd(); break;
            case 1: b[at] ^= (uint8_t)(1u << below(8)); break;
            case 2: b[at] = (it & 1) ? 0xff : 0x00; break;
            default: b[at] = (uint8_t)(b[at] + 1 - 2 * below(2)); break;
        }
    }
    if (below(6) == 0) {  // a huge varint (lengths and offsets near 2^64 / 2^32) written over whatever was there
        const size_t len = 4 + below(7), at = n - 1 - below(span < n ? span : n);
        for (size_t i = 0; i < len && at + i < n; i++) b[at + i] = (i + 1 == len) ? (uint8_t)(below(2) ? 0x01 : 0x7f) : 0xff;
    }
    if (below(16) == 0) b.resize(n - below(n < 64 ? n : 64));  // truncated tail
}

static long n_ok = 0, n_err = 0;

struct CbCtx {
    const uint8_t* p;
    size_t n;
};
static int read_at(void* ctx, uint64_t off, uint64_t len, uint8_t* dst) {
    const CbCtx* c = (const CbCtx*)ctx;
    if (off > c->n || len > c->n - off) return 5;  // a reader that refuses ranges outside the file
    memcpy(dst, c->p + off, len);
    return 0;
}

// The same bytes behind a read callback (ChunkReader): tail reads at open, one read per stripe at plan time, and a
// reader with a row selection whose batches are planned on the host.
static void one_file_callbacks(const uint8_t* heap, size_t n) {
    CbCtx ctx{heap, n};
    OrcbFile* f = nullptr;
    if (orcb_open_callbacks(n, read_at, &ctx, &f) != 0) return;
    OrcbReadOptions opt;
    memset(&opt, 0, sizeof opt);
    opt.use_row_index = 1;
    opt.batch_size = 1 + (uint32_t)below(5000);
    OrcbRowSelector sel[6];
    for (auto& s : sel) s.row_count = below(3000), s.skip = (int32_t)below(2), s.reserved = 0;
    OrcbReader* r = nullptr;
    if (orcb_reader_new_with_selection(f, &opt, sel, 6, &r) == 0) {
        int32_t applies[16];
        uint64_t triples[3 * 64];
        size_t ns = 0, nt = 0;
        orcb_reader_plan(r, applies, 16, &ns, triples, 64, &nt);
        orcb_reader_free(r);
    }
    // with_predicate on the callback-fed file: index areas and stripe footers are read on their own, then whole stripes
    if (const char* name = orcb_file_root_column_name(f, 0)) {
        OrcbPredicateNode node;
        memset(&node, 0, sizeof node);
        node.kind = ORCB_PRED_COMPARISON, node.op = ORCB_OP_GE, node.value_type = ORCB_VAL_INT64, node.i64 = 0, node.column = name;
        OrcbReaderBuild rb;
        memset(&rb, 0, sizeof rb);
        rb.options = &opt;
        rb.predicate = &node;
        rb.n_predicate_nodes = 1;
        OrcbReader* pr = nullptr;
        if (orcb_reader_build(f, &rb, &pr) == 0) {
            int32_t applies[16];
            uint64_t triples[3 * 64];
            size_t ns = 0, nt = 0;
            orcb_reader_plan(pr, applies, 16, &ns, triples, 64, &nt);
            orcb_reader_free(pr);
        }
    }
    OrcbJob* j = nullptr;
    OrcbFile* files[1] = {f};
    if (orcb_job_new(files, 1, &opt, &j) == 0) {
        orcb_job_plan(j);
        orcb_job_free(j);
    }
    uint64_t io[2];
    orcb_file_io_stats(f, io);
    orcb_file_free(f);
}

static void one_file(const std::vector<uint8_t>& bytes) {
    // exactly-sized heap copy: ASan's redzones sit right behind the last byte
    uint8_t* heap = (uint8_t*)malloc(bytes.size() ? bytes.size() : 1);
    memcpy(heap, bytes.data(), bytes.size());
    OrcbFile* f = nullptr;
    if (orcb_open_memory(heap, bytes.size(), &f) != 0) {
        n_err++;
        free(heap);
        return;
    }
    OrcbReadOptions opt;
    memset(&opt, 0, sizeof opt);
    opt.use_row_index = 1;
    struct ArrowSchema sch;
    memset(&sch, 0, sizeof sch);
    bool ok = true;
    if (orcb_schema(f, &opt, &sch) == 0) {
        if (sch.release) sch.release(&sch);
    } else {
        ok = false;
    }
    for (uint32_t s = 0, ns = orcb_file_num_stripes(f); s < ns && s < 8; s++) {
        uint64_t info[5];
        orcb_file_stripe_info(f, s, info);
    }
    for (int variant = 0; variant < 2; variant++) {
        opt.flags = variant;  // with and without the row index
        OrcbJob* j = nullptr;
        OrcbFile* files[1] = {f};
        if (orcb_job_new(files, 1, &opt, &j) == 0) {
            if (orcb_job_plan(j) != 0) ok = false;
            orcb_job_free(j);
        } else {
            ok = false;
        }
    }
    // with_predicate: row-group statistics and Bloom filters of the first stripes, for the first root columns
    opt.flags = 0;
    for (uint32_t c = 0, nc = orcb_file_num_root_columns(f); c < nc && c < 6; c++) {
        const char* name = orcb_file_root_column_name(f, c);
        if (!name) continue;
        OrcbPredicateNode nodes[5];
        memset(nodes, 0, sizeof nodes);
        nodes[0].kind = ORCB_PRED_OR;
        nodes[0].n_children = 4;
        nodes[1].kind = ORCB_PRED_COMPARISON, nodes[1].op = ORCB_OP_EQ, nodes[1].value_type = ORCB_VAL_INT64, nodes[1].i64 = 5, nodes[1].column = name;
        nodes[2].kind = ORCB_PRED_COMPARISON, nodes[2].op = ORCB_OP_EQ, nodes[2].value_type = ORCB_VAL_UTF8, nodes[2].str = (const uint8_t*)"abc", nodes[2].str_len = 3, nodes[2].column = name;
        nodes[3].kind = ORCB_PRED_COMPARISON, nodes[3].op = ORCB_OP_LT, nodes[3].value_type = ORCB_VAL_FLOAT64, nodes[3].f64 = 1.5, nodes[3].column = name;
        nodes[4].kind = ORCB_PRED_IS_NULL, nodes[4].column = name;
        for (uint32_t st = 0, ns = orcb_file_num_stripes(f); st < ns && st < 2; st++) {
            uint8_t keep[64];
            size_t ng = 0;
            int evaluated = 0;
            orcb_predicate_row_groups(f, st, &opt, nodes, 5, keep, sizeof keep, &ng, &evaluated);
        }
    }
    OrcbReader* r = nullptr;
    if (orcb_reader_new(f, &opt, &r) == 0) {
        orcb_reader_total_row_count(r);
        orcb_reader_free(r);
    } else {
        ok = false;
    }
    orcb_file_free(f);
    static const bool always_cb = getenv("ORCB_FUZZ_CALLBACKS") != nullptr;  // every input through the callback feed as well
    if (always_cb || below(4) == 0) one_file_callbacks(heap, bytes.size());
    free(heap);
    (ok ? n_ok : n_err)++;
}

static void one_section(int kind, const std::vector<uint8_t>& bytes, size_t block) {
    uint8_t* in = (uint8_t*)malloc(bytes.size() ? bytes.size() : 1);
    memcpy(in, bytes.data(), bytes.size());
    const size_t cap = 1u << 20;
    uint8_t* out = (uint8_t*)malloc(cap);
    size_t n = 0;
    (orcb_host_decompress_section(kind, in, bytes.size(), block, out, cap, &n) == 0 ? n_ok : n_err)++;
    free(out);
    free(in);
}

int main(int argc, char** argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: fuzz_host <iterations-per-seed> <seed> inputs...\n");
        return 2;
    }
    const int iters = atoi(argv[1]);
    rng_state ^= (uint64_t)atoll(argv[2]) * 0x9e3779b97f4a7c15ull;
    for (int a = 3; a < argc; a++) {
        const std::string path = argv[a];
        const std::vector<uint8_t> seed = slurp(argv[a]);
        if (seed.empty()) continue;
        const bool section = path.size() > 4 && path.compare(path.size() - 4, 4, ".sec") == 0;
        if (section) {
            const size_t slash = path.find_last_of('/');
            const int kind = atoi(path.c_str() + (slash == std::string::npos ? 0 : slash + 1));
            one_section(kind, seed, 1u << 18);
            for (int it = 0; it < iters; it++) {
                std::vector<uint8_t> b = seed;
                mutate(b, (it % 3 == 0) ? 16 : b.size(), it);
                one_section(kind, b, (it % 5 == 0) ? 4096 : (1u << 18));
            }
        } else {
            one_file(seed);
            for (int it = 0; it < iters; it++) {
                std::vector<uint8_t> b = seed;
                const uint64_t spans[3] = {400, 4000, b.size()};
                mutate(b, spans[it % 3], it);
                one_file(b);
            }
        }
    }
    printf("fuzz_host: %ld inputs accepted, %ld rejected with a status\n", n_ok, n_err);
    return 0;
}
