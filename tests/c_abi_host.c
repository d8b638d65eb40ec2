/* The C ABI from plain C (C99): what a cgo / JNI / Rust `extern "C"` binding sees.  Host-only calls, so it runs without
 * a GPU: open a file from memory, read its metadata and Arrow schema, plan a bulk job, evaluate a predicate, map an
 * error.  Compiled and run by tests/test_cabi_host.py::test_c_program_against_the_header.
 *   c_abi_host <file.orc> */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "orc_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc_ = (call);                                                            \
        if (rc_ != ORCB_OK) {                                                        \
            fprintf(stderr, "%s -> %d: %s\n", #call, rc_, orcb_last_error());        \
            return 1;                                                                \
        }                                                                            \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* fp = fopen(argv[1], "rb");
    if (!fp) return 2;
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    uint8_t* data = (uint8_t*)malloc((size_t)n);
    if (fread(data, 1, (size_t)n, fp) != (size_t)n) return 2;
    fclose(fp);

    OrcbFile* f = NULL;
    CHECK(orcb_open_memory(data, (size_t)n, &f));
    printf("rows=%llu stripes=%u compression=%d version=%s columns=%u\n", (unsigned long long)orcb_file_num_rows(f),
           orcb_file_num_stripes(f), (int)orcb_file_compression(f), orcb_file_format_version(f), orcb_file_num_root_columns(f));

    OrcbReadOptions opt;
    memset(&opt, 0, sizeof opt);
    opt.use_row_index = 1;
    struct ArrowSchema schema;
    CHECK(orcb_schema(f, &opt, &schema));
    printf("schema=%s children=%lld first=%s:%s\n", schema.format, (long long)schema.n_children, schema.children[0]->name,
           schema.children[0]->format);
    schema.release(&schema);

    OrcbJob* job = NULL;
    OrcbFile* files[1];
    files[0] = f;
    CHECK(orcb_job_new(files, 1, &opt, &job));
    CHECK(orcb_job_plan(job));
    OrcbJobStats st;
    CHECK(orcb_job_stats(job, &st));
    printf("planned stripes=%llu rows=%llu batches=%llu segments=%llu\n", (unsigned long long)st.n_stripes,
           (unsigned long long)st.n_rows, (unsigned long long)st.n_batches, (unsigned long long)st.n_segments);
    orcb_job_free(job);

    /* without a device the decode entry points report a status, they do not fall back to the CPU */
    if (!orcb_device_available()) {
        OrcbReader* r = NULL;
        CHECK(orcb_reader_new(f, &opt, &r));
        struct ArrowArray batch;
        int eos = 0;
        int rc = orcb_reader_next(r, &batch, &eos);
        printf("next_without_device=%d\n", rc);
        orcb_reader_free(r);
    }

    /* errors are statuses with a message, never aborts */
    OrcbFile* bad = NULL;
    int rc = orcb_open_memory(data, 3, &bad);
    printf("open_truncated=%d (%s)\n", rc, orcb_last_error());
    orcb_file_free(f);
    free(data);
    return 0;
}
