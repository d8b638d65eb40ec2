// extern "C" boundary (include/orc_b200.h).  No exception, panic or abort crosses it.
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <exception>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <atomic>
#include <memory>
#include <thread>

#include "job.h"
#include "job_internal.h"
#include "predicate.h"
#include "tz.h"
#include "kernels.h"
#include "meta.h"

using namespace orcb;

#define ORCB_STR2(x) #x
#define ORCB_STR(x) ORCB_STR2(x)

static thread_local std::string g_last_error;

struct OrcbFile {
    FileMeta meta;
    size_t pinned_cap = 0;  // capacity of meta.pinned (from the pinned-buffer cache)
};

// Bulk job = one or more "waves": contiguous slices of the stripe list, each its own Job on its own pair of streams.
// The integer path of a job is a chain (header walk -> run decode -> strings) whose links are latency- or
// issue-bound rather than bandwidth-bound; several waves in flight let the links of one wave run under those of another.
struct OrcbJob {
    std::vector<std::unique_ptr<Job>> waves;
    int device = 0;
    cudaStream_t user_stream = nullptr;  // the caller's stream (timing, ordering); waves fork from / join to it
    bool has_user_stream = false;
    cudaEvent_t ev_fork = nullptr;
    std::vector<cudaEvent_t> ev_join;
    ~OrcbJob() {
        waves.clear();
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto e : ev_join) cudaEventDestroy(e);
    }
    // batch i of the job -> (wave, batch of that wave)
    std::pair<Job*, uint64_t> locate(uint64_t i) const {
        for (auto& w : waves) {
            if (i < w->num_batches()) return {w.get(), i};
            i -= w->num_batches();
        }
        fail(ORCB_INVALID_ARGUMENT, "batch index out of range");
    }
};

// ArrowReader: walks the selected stripes in groups, one Job per group (src/arrow_reader.rs:233-347)
struct OrcbReader {
    OrcbFile* file;
    ReadOptions opt;
    std::vector<uint32_t> stripes;  // after byte-range / shard filtering (Cursor::get_stripe_metadatas :358-372)
    size_t next_stripe = 0;
    std::unique_ptr<Job> job;
    uint64_t next_batch = 0;
    bool failed = false;
    uint64_t segments_planned = 0, stripes_staged = 0;  // summed over the jobs run so far (orcb_reader_counters)
    // with_row_selection: per entry of `stripes`, whether a selection applies and the row ranges it yields
    bool has_selection = false;
    std::vector<std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>>> views;
    // with_predicate: a stripe's verdict is worked out when the reader gets to the stripe, like the reference does while
    // iterating (`views` grows stripe by stripe; `cursor` holds what is left of the caller's selection)
    bool lazy = false;
    SelectionCursor cursor;
    std::unique_ptr<Job> ahead;        // the next group of stripes, already launched
    std::exception_ptr ahead_error;    // what starting it ran into; reported when its turn comes
    bool has_predicate = false;
    Predicate predicate;
    std::vector<RowSelector> selectors;
    // stripes (indices into `stripes`) of the group started last, of r->job and of r->ahead; `single`: one stripe per
    // group from here on (a group failed: its good stripes are yielded first, the error surfaces at the bad one)
    size_t grp_first = 0, grp_end = 0, job_first = 0, job_end = 0;
    bool single = false;
    std::thread worker;  // orcb_reader_next_async: at most one request in flight
    ~OrcbReader() {
        if (worker.joinable()) worker.join();
    }
};

static std::atomic<uint64_t> g_index_retries{0}, g_layout_retries{0};

// finish() with the re-plan a LayoutRetry asks for (at most a few rounds: every round fixes the sizes of all chunks the
// device could decode)
static void finish_job(std::unique_ptr<Job>& job) {
    for (int round = 0;; round++) {
        try {
            job->finish();
            return;
        } catch (const IndexRetry&) {
            // a segment did not end where the row index says the next one starts: decode as the reference does,
            // sequentially (the damaged stream is then reported - or not - exactly as there)
            g_index_retries++;
            std::unique_ptr<Job> again = job->rebuild(true);
            again->plan();
            again->stage();
            again->launch();
            job = std::move(again);
        } catch (const LayoutRetry&) {
            if (round >= 3) fail(ORCB_UNEXPECTED, "compressed chunk sizes keep changing between decode passes");
            g_layout_retries++;
            std::unique_ptr<Job> again = job->rebuild();
            again->plan();
            again->stage();
            again->launch();
            job = std::move(again);
        }
    }
}

template <typename F>
static int guarded(F&& f) {
    try {
        f();
        return ORCB_OK;
    } catch (const OrcException& e) {
        g_last_error = e.msg;
        return e.code;
    } catch (const LayoutRetry&) {
        g_last_error = "compressed chunk sizes differ from the planned layout (finish the job before exporting from it)";
        return ORCB_UNEXPECTED;
    } catch (const IndexRetry&) {
        g_last_error = "row-index positions do not agree with the streams (finish the job before exporting from it)";
        return ORCB_UNEXPECTED;
    } catch (const std::bad_alloc&) {
        g_last_error = "out of host memory";
        return ORCB_UNEXPECTED;
    } catch (const std::exception& e) {
        g_last_error = e.what();
        return ORCB_UNEXPECTED;
    } catch (...) {
        g_last_error = "unknown error";
        return ORCB_UNEXPECTED;
    }
}

extern "C" {

const char* orcb_last_error(void) { return g_last_error.c_str(); }
uint64_t orcb_index_retries(void) { return g_index_retries.load(); }
uint64_t orcb_layout_retries(void) { return g_layout_retries.load(); }
const char* orcb_build_info(void) { return "orc_b200 " __DATE__ " sm_100a cuda " ORCB_STR(CUDART_VERSION); }

int orcb_device_available(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n > 0;
}

int orcb_open_memory(const uint8_t* data, size_t len, OrcbFile** out) {
    return guarded([&] {
        if (!out) fail(ORCB_INVALID_ARGUMENT, "out is NULL");
        auto f = std::make_unique<OrcbFile>();
        f->meta.data = data;
        f->meta.len = len;
        parse_file_tail(f->meta);
        *out = f.release();
    });
}

// The whole file into `buf`: slices read in parallel (page cache -> pinned memory runs at one core's copy speed, about
// 8 GB/s, so a 100 MB file costs more than its H2D copy and decode together when a single thread reads it)
static bool read_file_parallel(int fd, uint8_t* buf, size_t n) {
    static const unsigned max_threads = [] {
        unsigned t = 6;
        if (const char* e = getenv("ORCB_READ_THREADS")) t = (unsigned)std::max(1, atoi(e));
        return t;
    }();
    const size_t slice = 8u << 20;
    const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(max_threads, n / slice));
    std::vector<std::thread> th;
    std::vector<int> ok(nt, 1);
    const size_t per = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; t++) {
        const size_t a = std::min(n, t * per), b = std::min(n, (t + 1) * per);
        auto work = [fd, buf, a, b, &ok, t] {
            size_t p = a;
            while (p < b) {
                const ssize_t got = pread(fd, buf + p, b - p, (off_t)p);
                if (got <= 0) { ok[t] = 0; return; }
                p += (size_t)got;
            }
        };
        if (t + 1 == nt) work();
        else th.emplace_back(work);
    }
    for (auto& x : th) x.join();
    for (int v : ok)
        if (!v) return false;
    return true;
}

int orcb_open_path(const char* path, OrcbFile** out) {
    return guarded([&] {
        if (!out || !path) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        const int fd = open(path, O_RDONLY);
        if (fd < 0) fail(ORCB_IO_ERROR, std::string("cannot open ") + path);
        struct stat sb;
        if (fstat(fd, &sb) != 0) {
            close(fd);
            fail(ORCB_IO_ERROR, std::string("cannot stat ") + path);
        }
        const long n = (long)sb.st_size;
        auto f = std::make_unique<OrcbFile>();
        uint8_t* buf = nullptr;
        // with a device present the file is read into pinned memory (asynchronous H2D copies); the buffer comes from
        // the cache of pinned buffers, so opening file after file does not pin and unpin 100 MB each time
        if (n > 0 && orcb_device_available()) {
            try {
                buf = (uint8_t*)pinned_get((size_t)n, &f->pinned_cap);
                f->meta.pinned = buf;
            } catch (const OrcException&) {
                buf = nullptr;
            }
        }
        if (!buf) {
            cudaGetLastError();
            f->meta.owned.resize((size_t)std::max<long>(n, 0));
            buf = f->meta.owned.data();
        }
        const bool ok = n > 0 ? read_file_parallel(fd, buf, (size_t)n) : true;
        close(fd);
        if (!ok) {
            if (f->meta.pinned) pinned_put(f->meta.pinned, f->pinned_cap);
            fail(ORCB_IO_ERROR, std::string("short read on ") + path);
        }
        f->meta.data = buf;
        f->meta.len = (size_t)std::max<long>(n, 0);
        try {
            parse_file_tail(f->meta);
        } catch (...) {
            if (f->meta.pinned) pinned_put(f->meta.pinned, f->pinned_cap);
            throw;
        }
        *out = f.release();
    });
}

int orcb_open_callbacks(uint64_t len, OrcbReadAt read_at, void* ctx, OrcbFile** out) {
    return guarded([&] {
        if (!out || !read_at) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        auto f = std::make_unique<OrcbFile>();
        f->meta.len = (size_t)len;
        f->meta.source = std::make_shared<RangeSource>();
        f->meta.source->read_at = read_at;
        f->meta.source->ctx = ctx;
        parse_file_tail(f->meta);
        *out = f.release();
    });
}

int orcb_file_io_stats(const OrcbFile* f, uint64_t out[2]) {
    return guarded([&] {
        if (!f || !out) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        out[0] = out[1] = 0;
        if (f->meta.source) {
            std::lock_guard<std::mutex> lock(f->meta.source->mu);
            out[0] = f->meta.source->reads;
            out[1] = f->meta.source->bytes_read;
        }
    });
}

int orcb_file_clone(const OrcbFile* f, OrcbFile** out) {
    return guarded([&] {
        if (!f || !out) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        auto c = std::make_unique<OrcbFile>();
        c->meta = f->meta;  // parsed tail; the bytes stay with the original
        c->meta.owned.clear();
        c->meta.pinned = nullptr;
        c->meta.data = f->meta.data;
        *out = c.release();
    });
}

void orcb_file_free(OrcbFile* f) {
    if (!f) return;
    if (f->meta.pinned) pinned_put(f->meta.pinned, f->pinned_cap);
    delete f;
}

uint64_t orcb_file_num_rows(const OrcbFile* f) { return f->meta.num_rows; }
uint32_t orcb_file_num_stripes(const OrcbFile* f) { return (uint32_t)f->meta.stripes.size(); }
int32_t orcb_file_compression(const OrcbFile* f) { return f->meta.compression; }
uint64_t orcb_file_compression_block_size(const OrcbFile* f) { return f->meta.block_size; }
int64_t orcb_file_row_index_stride(const OrcbFile* f) { return f->meta.row_index_stride; }
uint32_t orcb_file_num_root_columns(const OrcbFile* f) { return (uint32_t)f->meta.root_columns.size(); }
const char* orcb_file_root_column_name(const OrcbFile* f, uint32_t i) {
    return i < f->meta.root_columns.size() ? f->meta.root_columns[i].first.c_str() : nullptr;
}
uint32_t orcb_file_root_column_id(const OrcbFile* f, uint32_t i) { return i < f->meta.root_columns.size() ? f->meta.root_columns[i].second : 0u; }
const char* orcb_file_format_version(const OrcbFile* f) { return f->meta.format_version.c_str(); }
uint32_t orcb_file_num_user_metadata(const OrcbFile* f) { return (uint32_t)f->meta.user_metadata.size(); }
int orcb_file_user_metadata(const OrcbFile* f, uint32_t i, const char** key, const uint8_t** value, size_t* value_len) {
    return guarded([&] {
        if (!f || !key || !value || !value_len) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        if (i >= f->meta.user_metadata.size()) fail(ORCB_INVALID_ARGUMENT, "metadata index out of range");
        *key = f->meta.user_metadata[i].first.c_str();
        *value = (const uint8_t*)f->meta.user_metadata[i].second.data();
        *value_len = f->meta.user_metadata[i].second.size();
    });
}
int orcb_file_stripe_info(const OrcbFile* f, uint32_t stripe, uint64_t out[5]) {
    return guarded([&] {
        if (stripe >= f->meta.stripes.size()) fail(ORCB_INVALID_ARGUMENT, "stripe index out of range");
        const StripeInfo& s = f->meta.stripes[stripe];
        out[0] = s.offset;
        out[1] = s.index_length;
        out[2] = s.data_length;
        out[3] = s.footer_length;
        out[4] = s.rows;
    });
}

int orcb_schema(const OrcbFile* f, const OrcbReadOptions* opt, struct ArrowSchema* out) {
    return guarded([&] {
        ReadOptions o = ReadOptions::from_c(opt);
        auto cols = project_columns(f->meta, o);
        export_schema(f->meta, cols, o, out);
    });
}

static std::vector<uint32_t> select_stripes(const FileMeta& fm, const ReadOptions& o) {
    std::vector<uint32_t> out;
    uint32_t ordinal = 0;
    for (uint32_t i = 0; i < fm.stripes.size(); i++) {
        if (o.range_start || o.range_end) {
            const uint64_t off = fm.stripes[i].offset;
            if (off < o.range_start || off >= o.range_end) continue;
        }
        if (o.shard_count > 1 && (ordinal % o.shard_count) != o.shard_index) {
            ordinal++;
            continue;
        }
        ordinal++;
        out.push_back(i);
    }
    return out;
}

int orcb_reader_new(OrcbFile* f, const OrcbReadOptions* opt, OrcbReader** out) {
    return guarded([&] {
        if (!f || !out) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        auto r = std::make_unique<OrcbReader>();
        r->file = f;
        r->opt = ReadOptions::from_c(opt);
        (void)project_columns(f->meta, r->opt);  // surfaces schema errors at build time
        r->stripes = select_stripes(f->meta, r->opt);
        *out = r.release();
    });
}

int orcb_reader_new_ex(OrcbFile* f, const OrcbReadOptions* opt, const OrcbRowSelector* selectors, uint32_t n_selectors,
                       int has_selection, const struct ArrowSchema* schema, OrcbReader** out) {
    return guarded([&] {
        if (!f || !out || (!selectors && n_selectors)) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        auto r = std::make_unique<OrcbReader>();
        r->file = f;
        r->opt = ReadOptions::from_c(opt);
        apply_schema_hints(f->meta, r->opt, schema);
        (void)project_columns(f->meta, r->opt);
        r->stripes = select_stripes(f->meta, r->opt);
        if (has_selection) {
            std::vector<RowSelector> sel;
            for (uint32_t i = 0; i < n_selectors; i++) sel.push_back({selectors[i].row_count, selectors[i].skip != 0});
            std::vector<uint64_t> rows;
            for (uint32_t s : r->stripes) rows.push_back(f->meta.stripes[s].rows);
            r->views = selection_views(std::move(sel), rows, r->opt.batch_size);
            r->has_selection = true;
        }
        *out = r.release();
    });
}

int orcb_reader_build(OrcbFile* f, const OrcbReaderBuild* b, OrcbReader** out) {
    return guarded([&] {
        if (!f || !b || !out || (!b->selectors && b->n_selectors)) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        auto r = std::make_unique<OrcbReader>();
        r->file = f;
        r->opt = ReadOptions::from_c(b->options);
        apply_schema_hints(f->meta, r->opt, b->schema);
        (void)project_columns(f->meta, r->opt);
        r->stripes = select_stripes(f->meta, r->opt);
        if (b->predicate || b->n_predicate_nodes) {
            r->predicate = predicate_from_c(b->predicate, b->n_predicate_nodes);
            r->has_predicate = true;
        }
        if (b->has_selection)
            for (uint32_t i = 0; i < b->n_selectors; i++) r->selectors.push_back({b->selectors[i].row_count, b->selectors[i].skip != 0});
        r->has_selection = b->has_selection != 0;
        if (!r->has_predicate && r->has_selection) {
            std::vector<uint64_t> rows;
            for (uint32_t s : r->stripes) rows.push_back(f->meta.stripes[s].rows);
            r->views = selection_views(r->selectors, rows, r->opt.batch_size);
        }
        if (r->has_predicate) {
            r->lazy = true;
            r->cursor = SelectionCursor(r->selectors, r->has_selection);
            r->has_selection = true;  // `views` (filled as the reader advances) says what every stripe yields
        }
        *out = r.release();
    });
}

int orcb_predicate_row_groups(OrcbFile* f, uint32_t stripe, const OrcbReadOptions* opt, const OrcbPredicateNode* predicate,
                              uint32_t n_predicate_nodes, uint8_t* keep, size_t cap, size_t* n_groups, int* evaluated) {
    return guarded([&] {
        if (!f || !n_groups || !evaluated) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        if (stripe >= f->meta.stripes.size()) fail(ORCB_INVALID_ARGUMENT, "stripe out of range");
        const ReadOptions o = ReadOptions::from_c(opt);
        const auto cols = project_columns(f->meta, o);
        const Predicate p = predicate_from_c(predicate, n_predicate_nodes);
        std::vector<uint8_t> filter;
        bool ok = false;
        (void)predicate_selection(f->meta, stripe, cols, p, &filter, &ok);
        *evaluated = ok ? 1 : 0;
        *n_groups = filter.size();
        for (size_t i = 0; keep && i < filter.size() && i < cap; i++) keep[i] = filter[i];
    });
}

int orcb_zone_table(const char* name, int64_t* at, int32_t* off, size_t cap, size_t* n, int32_t* first_off, int64_t* orc_epoch) {
    return guarded([&] {
        if (!name || !n || (cap && (!at || !off))) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        ZoneTable z;
        std::string why;
        if (!load_zone_table(name, z, why)) fail(ORCB_NOT_IMPLEMENTED, "no table for zone " + std::string(name) + ": " + why);
        *n = z.at.size();
        for (size_t i = 0; i < z.at.size() && i < cap; i++) at[i] = z.at[i], off[i] = z.off[i];
        if (first_off) *first_off = z.first_off;
        if (orc_epoch) {
            const int64_t local = ORC_EPOCH_UTC;  // 2015-01-01 00:00:00
            if (!zone_local_to_utc(z, local, *orc_epoch)) fail(ORCB_UNEXPECTED, "2015-01-01 00:00 is not a unique instant in zone " + std::string(name));
        }
    });
}

uint64_t orcb_bloom_hash_long(int64_t value) { return bloom_hash_long(value); }
uint64_t orcb_bloom_hash_bytes(const uint8_t* bytes, size_t len) { return bloom_hash_bytes(bytes, len); }

int orcb_reader_new_with_selection(OrcbFile* f, const OrcbReadOptions* opt, const OrcbRowSelector* selectors, uint32_t n_selectors,
                                   OrcbReader** out) {
    return orcb_reader_new_ex(f, opt, selectors, n_selectors, 1, nullptr, out);
}

int orcb_selection_plan(const OrcbRowSelector* selectors, uint32_t n_selectors, const uint64_t* stripe_rows, uint32_t n_stripes,
                        uint64_t batch_size, int32_t* applies, uint64_t* triples, size_t cap_triples, size_t* n_triples) {
    return guarded([&] {
        if ((!selectors && n_selectors) || (!stripe_rows && n_stripes) || !n_triples) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        std::vector<RowSelector> sel;
        for (uint32_t i = 0; i < n_selectors; i++) sel.push_back({selectors[i].row_count, selectors[i].skip != 0});
        const auto plan = selection_views(std::move(sel), std::vector<uint64_t>(stripe_rows, stripe_rows + n_stripes), batch_size ? batch_size : 8192);
        size_t k = 0;
        for (uint32_t s = 0; s < n_stripes; s++) {
            if (applies) applies[s] = plan[s].first ? 1 : 0;
            for (auto& v : plan[s].second) {
                if (triples && k < cap_triples) {
                    triples[3 * k] = s;
                    triples[3 * k + 1] = v.first;
                    triples[3 * k + 2] = v.second;
                }
                k++;
            }
        }
        *n_triples = k;
    });
}

int orcb_reader_counters(const OrcbReader* r, uint64_t out[2]) {
    return guarded([&] {
        if (!r || !out) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        out[0] = r->segments_planned;
        out[1] = r->stripes_staged;
    });
}

void orcb_reader_free(OrcbReader* r) { delete r; }

uint64_t orcb_reader_total_row_count(const OrcbReader* r) { return r->file->meta.num_rows; }

// with_predicate: the verdict of stripe `idx` (and of every stripe before it), combined with the caller's selection
// (ArrowReader::try_advance_stripe)
static void ensure_views(OrcbReader* r, size_t idx) {
    if (!r->lazy) return;
    while (r->views.size() <= idx && r->views.size() < r->stripes.size()) {
        const uint32_t s = r->stripes[r->views.size()];
        const auto cols = project_columns(r->file->meta, r->opt);
        const std::vector<RowSelector> pred = predicate_selection(r->file->meta, s, cols, r->predicate, nullptr, nullptr);
        r->views.push_back(r->cursor.next_stripe(r->file->meta.stripes[s].rows, r->opt.batch_size, &pred));
    }
}

int orcb_reader_plan(OrcbReader* r, int32_t* applies, size_t cap_stripes, size_t* n_stripes, uint64_t* triples,
                     size_t cap_triples, size_t* n_triples) {
    return guarded([&] {
        if (!r || !n_triples || !n_stripes) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        if (!r->stripes.empty()) ensure_views(r, r->stripes.size() - 1);
        *n_stripes = r->stripes.size();
        size_t k = 0;
        for (size_t s = 0; s < r->stripes.size(); s++) {
            const bool restricted = r->has_selection && r->views[s].first;
            if (applies && s < cap_stripes) applies[s] = restricted ? 1 : 0;
            if (!restricted) continue;
            for (auto& v : r->views[s].second) {
                if (triples && k < cap_triples) {
                    triples[3 * k] = s;
                    triples[3 * k + 1] = v.first;
                    triples[3 * k + 2] = v.second;
                }
                k++;
            }
        }
        *n_triples = k;
    });
}

// The next group of stripes (at most max_stripes_per_launch, ~1 GiB of input) as a job that is planned, staged and
// launched but not waited for; nullptr when no stripe is left.
static std::unique_ptr<Job> reader_start_group(OrcbReader* r) {
    for (;;) {
        if (r->next_stripe >= r->stripes.size()) return nullptr;
        r->grp_first = r->grp_end = r->next_stripe;
        const uint32_t group = r->single ? 1u : (r->opt.max_stripes_per_launch ? r->opt.max_stripes_per_launch : 16);
        std::vector<StripeTask> tasks;
        uint64_t bytes = 0;
        while (r->next_stripe < r->stripes.size() && tasks.size() < group) {
            const StripeInfo& si = r->file->meta.stripes[r->stripes[r->next_stripe]];
            if (!tasks.empty() && bytes + si.data_length > (1ull << 30)) break;  // bound one launch plan to ~1 GiB in
            StripeTask task{&r->file->meta, r->stripes[r->next_stripe]};
            ensure_views(r, r->next_stripe);
            if (r->has_selection && r->views[r->next_stripe].first) {
                const auto& views = r->views[r->next_stripe].second;
                if (views.empty()) {  // nothing selected in this stripe: not even staged
                    r->next_stripe++;
                    continue;
                }
                task.has_views = true;
                const uint64_t stride = r->file->meta.row_index_stride > 0 ? (uint64_t)r->file->meta.row_index_stride : 0;
                if (r->opt.use_row_index && stride && si.rows > stride) {
                    // partial decode: one task per run of consecutive row groups the selected ranges touch
                    bytes += si.data_length;
                    for (size_t i = 0; i < views.size();) {
                        uint64_t g0 = views[i].first / stride, g1 = ((uint64_t)views[i].first + views[i].second - 1) / stride + 1;
                        size_t j = i + 1;
                        while (j < views.size() && views[j].first / stride <= g1) {
                            g1 = std::max<uint64_t>(g1, ((uint64_t)views[j].first + views[j].second - 1) / stride + 1);
                            j++;
                        }
                        StripeTask w = task;
                        w.has_window = true;
                        w.g_begin = (uint32_t)g0;
                        w.g_end = (uint32_t)g1;
                        w.views.assign(views.begin() + i, views.begin() + j);
                        tasks.push_back(std::move(w));
                        i = j;
                    }
                    r->next_stripe++;
                    continue;
                }
                task.views = views;
            }
            bytes += si.data_length;
            tasks.push_back(std::move(task));
            r->next_stripe++;
        }
        r->grp_end = r->next_stripe;
        if (tasks.empty()) continue;  // only unselected stripes were left in this round
        const bool timing = getenv("ORCB_READER_TIMING") != nullptr;  // host phases to stderr (tools/reader_probe.py)
        auto now = [] { return std::chrono::steady_clock::now(); };
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        const auto t0 = now();
        auto job = std::make_unique<Job>(std::move(tasks), r->opt);
        job->plan();
        {
            OrcbJobStats st{};
            job->stats(&st);
            r->segments_planned += st.n_segments;
            r->stripes_staged += st.n_stripes;
        }
        const auto t1 = now();
        job->stage();
        const auto t2 = now();
        job->launch();
        if (timing)
            fprintf(stderr, "orcb reader: plan %.2f ms, stage %.2f ms, launch %.2f ms\n", ms(t0, t1), ms(t1, t2), ms(t2, now()));
        return job;
    }
}

// Makes r->job hold the group that r->next_batch indexes; false at end of stream.  While the caller consumes one
// group, the next one is already copied and decoded on its own streams (ORCB_NO_PREFETCH=1 turns that off).  What goes
// wrong while starting a group ahead of time is reported when that group's turn comes, after every batch before it,
// as the reference reports a bad stripe only when it gets there.
static bool reader_advance(OrcbReader* r) {
    static const bool prefetch = !(getenv("ORCB_NO_PREFETCH") && getenv("ORCB_NO_PREFETCH")[0] == '1');
    // A group of several stripes that fails (while it is planned, or on the device) is started again one stripe at a
    // time: every batch of the stripes before the bad one is yielded first, as the reference does
    // (src/arrow_reader.rs:296-316 decodes stripe by stripe).
    auto retry_single = [&](size_t first, size_t end) {
        if (r->single || end - first <= 1) return false;
        r->single = true;
        r->next_stripe = first;
        r->job.reset();
        r->ahead.reset();
        r->ahead_error = nullptr;
        return true;
    };
    while (!r->job || r->next_batch >= r->job->num_batches()) {
        r->job.reset();
        if (r->ahead_error) {
            std::exception_ptr e = r->ahead_error;
            r->ahead_error = nullptr;
            if (retry_single(r->grp_first, r->grp_end > r->grp_first ? r->grp_end : r->stripes.size())) continue;
            std::rethrow_exception(e);
        }
        if (!r->ahead) {
            try {
                r->ahead = reader_start_group(r);
            } catch (const OrcException&) {
                if (retry_single(r->grp_first, r->grp_end > r->grp_first ? r->grp_end : r->stripes.size())) continue;
                throw;
            }
        }
        if (!r->ahead) return false;
        r->job = std::move(r->ahead);
        r->job_first = r->grp_first;
        r->job_end = r->grp_end;
        r->next_batch = 0;
        if (prefetch) {  // planned and launched on the host while the device still works on r->job
            try {
                r->ahead = reader_start_group(r);
            } catch (...) {
                r->ahead_error = std::current_exception();
            }
        }
        try {
            finish_job(r->job);
        } catch (const OrcException&) {
            if (retry_single(r->job_first, r->job_end)) continue;
            throw;
        }
    }
    return true;
}

int orcb_reader_next(OrcbReader* r, struct ArrowArray* out, int* eos) {
    return guarded([&] {
        if (!r || !out || !eos) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        *eos = 0;
        if (r->failed) fail(ORCB_UNEXPECTED, "reader is in the error state");
        try {
            if (!reader_advance(r)) {
                *eos = 1;
                return;
            }
            r->job->export_batch(r->next_batch++, out);
        } catch (...) {
            r->failed = true;  // StreamState::Error analogue (src/async_arrow_reader.rs:262-277)
            throw;
        }
    });
}

// poll_next of ArrowStreamReader (src/async_arrow_reader.rs:283-290) in callback form: the batch is produced on a
// thread of the library, the caller's thread returns at once
int orcb_reader_next_async(OrcbReader* r, struct ArrowArray* out, OrcbBatchCallback done, void* ctx) {
    return guarded([&] {
        if (!r || !out || !done) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        if (r->worker.joinable()) r->worker.join();  // the request before this one has completed or is completing
        r->worker = std::thread([r, out, done, ctx] {
            int eos = 0;
            const int rc = orcb_reader_next(r, out, &eos);
            done(ctx, rc, eos, rc ? orcb_last_error() : "");
        });
    });
}

int orcb_reader_next_device(OrcbReader* r, struct ArrowDeviceArray* out, int* eos) {
    return guarded([&] {
        if (!r || !out || !eos) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        *eos = 0;
        if (r->failed) fail(ORCB_UNEXPECTED, "reader is in the error state");
        try {
            if (!reader_advance(r)) {
                *eos = 1;
                return;
            }
            r->job->export_batch_device(r->next_batch++, out);
        } catch (...) {
            r->failed = true;
            throw;
        }
    });
}

// `for batch in reader {}` without leaving the library: every batch is produced exactly as orcb_reader_next /
// orcb_reader_next_device would hand it out (host copy included for host-resident readers) and released at once.
int orcb_reader_drain(OrcbReader* r, uint64_t out[2]) {
    return guarded([&] {
        if (!r || !out) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        out[0] = out[1] = 0;
        if (r->failed) fail(ORCB_UNEXPECTED, "reader is in the error state");
        try {
            while (reader_advance(r)) {
                if (r->opt.device_resident) {
                    ArrowDeviceArray a;
                    r->job->export_batch_device(r->next_batch++, &a);
                    out[1] += (uint64_t)a.array.length;
                    a.array.release(&a.array);
                } else {
                    ArrowArray a;
                    r->job->export_batch(r->next_batch++, &a);
                    out[1] += (uint64_t)a.length;
                    a.release(&a);
                }
                out[0]++;
            }
        } catch (...) {
            r->failed = true;
            throw;
        }
    });
}

// stripes of all files in order; multi-GPU sharding counts over this list (stripe i -> rank i % count), not per file
static std::vector<StripeTask> job_tasks(OrcbFile* const* files, uint32_t n_files, const ReadOptions& o) {
    ReadOptions per_file = o;
    per_file.shard_index = 0;
    per_file.shard_count = 1;
    std::vector<StripeTask> tasks;
    uint64_t ordinal = 0;
    for (uint32_t i = 0; i < n_files; i++)
        for (uint32_t s : select_stripes(files[i]->meta, per_file)) {
            if (o.shard_count <= 1 || ordinal % o.shard_count == o.shard_index) tasks.push_back({&files[i]->meta, s});
            ordinal++;
        }
    return tasks;
}

int orcb_job_new(OrcbFile* const* files, uint32_t n_files, const OrcbReadOptions* opt, OrcbJob** out) {
    return guarded([&] {
        if (!files || !n_files || !out) fail(ORCB_INVALID_ARGUMENT, "NULL argument");
        ReadOptions o = ReadOptions::from_c(opt);
        std::vector<StripeTask> tasks = job_tasks(files, n_files, o);
        auto j = std::make_unique<OrcbJob>();
        j->device = o.device;
        // waves: the caller's number, else ORCB_WAVES, else one per ~48 stripes up to 2 (more measured slower)
        uint32_t waves = o.waves;
        if (!waves)
            if (const char* e = getenv("ORCB_WAVES")) waves = (uint32_t)atoi(e);
        if (!waves) waves = (uint32_t)std::min<size_t>(2, (tasks.size() + 47) / 48);
        waves = std::max<uint32_t>(1, std::min<uint32_t>(waves, (uint32_t)std::max<size_t>(tasks.size(), 1)));
        if (waves > 1) {
            j->has_user_stream = !o.own_stream;
            j->user_stream = o.stream;
            o.own_stream = true;
            o.stream = nullptr;
        }
        const size_t per = (tasks.size() + waves - 1) / waves;
        for (uint32_t w = 0; w < waves; w++) {
            const size_t a = std::min(tasks.size(), w * per), b = std::min(tasks.size(), (w + 1) * per);
            if (w > 0 && a == b) break;
            j->waves.push_back(std::make_unique<Job>(std::vector<StripeTask>(tasks.begin() + a, tasks.begin() + b), o));
        }
        *out = j.release();
    });
}
void orcb_job_free(OrcbJob* j) { delete j; }
int orcb_job_plan(OrcbJob* j) {
    return guarded([&] { for (auto& w : j->waves) w->plan(); });
}
int orcb_job_stage(OrcbJob* j) {
    return guarded([&] { for (auto& w : j->waves) w->stage(); });
}
int orcb_job_launch(OrcbJob* j) {
    return guarded([&] {
        const bool serial = getenv("ORCB_SERIAL") != nullptr;  // one kernel at a time: clean per-kernel timings
        if (j->waves.size() > 1 && j->has_user_stream) {
            for (auto& w : j->waves) w->stage();
            CUDA_OK(cudaSetDevice(j->device));
            if (!j->ev_fork) CUDA_OK(cudaEventCreateWithFlags(&j->ev_fork, cudaEventDisableTiming));
            while (j->ev_join.size() < j->waves.size()) {
                cudaEvent_t e = nullptr;
                CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                j->ev_join.push_back(e);
            }
            CUDA_OK(cudaEventRecord(j->ev_fork, j->user_stream));
            for (auto& w : j->waves) CUDA_OK(cudaStreamWaitEvent(w->stream(), j->ev_fork, 0));
        }
        for (size_t i = 0; i < j->waves.size(); i++) {
            j->waves[i]->launch();
            if (serial && j->waves.size() > 1) CUDA_OK(cudaStreamSynchronize(j->waves[i]->stream()));
        }
        if (j->waves.size() > 1 && j->has_user_stream) {
            for (size_t i = 0; i < j->waves.size(); i++) {
                CUDA_OK(cudaEventRecord(j->ev_join[i], j->waves[i]->stream()));
                CUDA_OK(cudaStreamWaitEvent(j->user_stream, j->ev_join[i], 0));
            }
        }
    });
}
int orcb_job_finish(OrcbJob* j) {
    return guarded([&] { for (auto& w : j->waves) finish_job(w); });
}
int orcb_job_stats(const OrcbJob* j, OrcbJobStats* out) {
    return guarded([&] {
        memset(out, 0, sizeof(*out));
        for (auto& w : j->waves) {
            OrcbJobStats s{};
            w->stats(&s);
            out->n_stripes += s.n_stripes;
            out->n_rows += s.n_rows;
            out->n_columns = s.n_columns;
            out->input_bytes += s.input_bytes;
            out->staged_bytes += s.staged_bytes;
            out->output_bytes += s.output_bytes;
            out->device_bytes += s.device_bytes;
            out->n_segments += s.n_segments;
            out->n_kernel_launches += s.n_kernel_launches;
            out->n_batches += s.n_batches;
            out->d2h_meta_bytes += s.d2h_meta_bytes;
            out->aliased_output_bytes += s.aliased_output_bytes;
        }
        out->n_waves = j->waves.size();
    });
}
int orcb_job_restage(OrcbJob* j) {
    return guarded([&] { for (auto& w : j->waves) w->restage(); });
}
int orcb_job_kernel_stats(const OrcbJob* j, OrcbKernelStat* out, uint32_t cap, uint32_t* n) {
    return guarded([&] {
        // summed over the waves by kernel name (waves overlap: the sum of a kernel's times can exceed the step)
        uint32_t total = 0;
        std::vector<OrcbKernelStat> tmp(64);
        for (auto& w : j->waves) {
            const uint32_t k = w->kernel_stats(tmp.data(), (uint32_t)tmp.size());
            for (uint32_t i = 0; i < k; i++) {
                uint32_t at = 0;
                while (at < total && strncmp(out[at].name, tmp[i].name, sizeof(out[at].name)) != 0) at++;
                if (at == total) {
                    if (total >= cap) continue;
                    out[total++] = tmp[i];
                } else {
                    out[at].ms += tmp[i].ms;
                    out[at].alg_bytes += tmp[i].alg_bytes;
                    out[at].work_items += tmp[i].work_items;
                }
            }
        }
        *n = total;
    });
}
uint64_t orcb_job_num_batches(const OrcbJob* j) {
    uint64_t n = 0;
    for (auto& w : j->waves) n += w->num_batches();
    return n;
}
int orcb_job_export_batch(OrcbJob* j, uint64_t i, struct ArrowArray* out) {
    return guarded([&] {
        for (auto& w : j->waves) finish_job(w);
        auto at = j->locate(i);
        at.first->export_batch(at.second, out);
    });
}
int orcb_job_export_batch_device(OrcbJob* j, uint64_t i, struct ArrowDeviceArray* out) {
    return guarded([&] {
        for (auto& w : j->waves) finish_job(w);
        auto at = j->locate(i);
        at.first->export_batch_device(at.second, out);
    });
}

// ---------------------------------------------------------------------------------------------
// stream-level entry points: one segment, one kernel
// ---------------------------------------------------------------------------------------------
#define CU(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) fail(ORCB_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)

namespace {
struct DevBuf {
    void* p = nullptr;
    explicit DevBuf(size_t n) { CU(cudaMalloc(&p, n ? n : 16)); }
    ~DevBuf() { cudaFree(p); }
    DevBuf(const DevBuf&) = delete;
};
struct StreamRun {
    DevBuf in, err;
    size_t in_len;
    StreamRun(int device, const uint8_t* src, size_t n) : in(((cudaSetDevice(device)), n + 256)), err(16), in_len(n) {
        CU(cudaMemset(in.p, 0, n + 256));
        if (n) CU(cudaMemcpy(in.p, src, n, cudaMemcpyHostToDevice));
        CU(cudaMemset(err.p, 0, 16));
    }
    void check() {
        CU(cudaDeviceSynchronize());
        uint32_t e[2] = {0, 0};
        CU(cudaMemcpy(e, err.p, 8, cudaMemcpyDeviceToHost));
        if (e[0]) fail((int)e[0], "device decode error");
    }
};
}  // namespace

int orcb_decode_int_rle(int device, const uint8_t* in, size_t in_len, int version, int is_signed, int nbytes,
                        int64_t* out, size_t n_values) {
    return guarded([&] {
        if (nbytes != 2 && nbytes != 4 && nbytes != 8) fail(ORCB_INVALID_ARGUMENT, "nbytes must be 2, 4 or 8");
        StreamRun sr(device, in, in_len);
        DevBuf dout(n_values * 8 + 16), dseg(sizeof(Seg)), mis(16);
        CU(cudaMemset(mis.p, 0, 16));
        Seg s{};
        s.in = (uint64_t)(uintptr_t)sr.in.p;
        s.out = (uint64_t)(uintptr_t)dout.p;
        s.in_len = (uint32_t)in_len;
        s.n_values = (uint32_t)n_values;
        s.cnt_idx = -1;
        s.start_idx = -1;
        s.flags = (is_signed ? SEG_SIGNED : 0) | (version == 1 ? SEG_RLE_V2 : 0);
        s.nbytes = (uint8_t)nbytes;
        s.out_kind = OUT_I64;
        // same two kernels as the job path: header-walk pre-pass, then one warp per 32 runs
        const uint32_t cap = (uint32_t)std::min<size_t>(in_len / 2 + 3, n_values + 2);
        const uint32_t pool = (cap + 31) / 32 + 1;
        DevBuf dtab((size_t)pool * 32 * sizeof(RunRec)), dblk((size_t)pool * sizeof(BlockRec)), dcnt(16), dslow((size_t)pool * 4 + 16);
        const uint32_t qcap = (uint32_t)(n_values / 64 + 1024);
        DevBuf dq((size_t)qcap * sizeof(CoopRec));
        CU(cudaMemset(dcnt.p, 0, 16));
        s.run_cap = cap;
        CU(cudaMemcpy(dseg.p, &s, sizeof(s), cudaMemcpyHostToDevice));
        int rc = launch_rle_index((Seg*)dseg.p, 1, nullptr, (RunRec*)dtab.p, (BlockRec*)dblk.p, (uint32_t*)dcnt.p, pool,
                                  (CoopRec*)dq.p, (uint32_t*)dcnt.p + 2, qcap, (uint32_t*)sr.err.p, nullptr, 0);
        if (rc) fail(ORCB_CUDA, cudaGetErrorString((cudaError_t)rc));
        rc = launch_int_rle((Seg*)dseg.p, (BlockRec*)dblk.p, (uint32_t*)dcnt.p, pool, (RunRec*)dtab.p, nullptr, nullptr,
                            (uint32_t*)sr.err.p, (uint32_t*)mis.p, (uint32_t*)dslow.p, (uint32_t*)dcnt.p + 1,
                            (CoopRec*)dq.p, (uint32_t*)dcnt.p + 2, qcap, nullptr, 0);
        if (rc) fail(ORCB_CUDA, cudaGetErrorString((cudaError_t)rc));
        sr.check();
        if (n_values) CU(cudaMemcpy(out, dout.p, n_values * 8, cudaMemcpyDeviceToHost));
    });
}

int orcb_decode_byte_rle(int device, const uint8_t* in, size_t in_len, uint8_t* out, size_t n_values) {
    return guarded([&] {
        StreamRun sr(device, in, in_len);
        DevBuf dout(n_values + 16), dseg(sizeof(Seg));
        Seg s{};
        s.in = (uint64_t)(uintptr_t)sr.in.p;
        s.out = (uint64_t)(uintptr_t)dout.p;
        s.in_len = (uint32_t)in_len;
        s.n_values = (uint32_t)n_values;
        s.cnt_idx = -1;
        s.start_idx = -1;
        s.out_kind = OUT_I8;
        CU(cudaMemcpy(dseg.p, &s, sizeof(s), cudaMemcpyHostToDevice));
        int rc = launch_byte_rle((Seg*)dseg.p, 1, nullptr, nullptr, (uint32_t*)sr.err.p, nullptr, 0);
        if (rc) fail(ORCB_CUDA, cudaGetErrorString((cudaError_t)rc));
        sr.check();
        if (n_values) CU(cudaMemcpy(out, dout.p, n_values, cudaMemcpyDeviceToHost));
    });
}

int orcb_decode_bool_rle(int device, const uint8_t* in, size_t in_len, uint8_t* out_bitmap, size_t n_values) {
    return guarded([&] {
        StreamRun sr(device, in, in_len);
        const size_t raw_bytes = (n_values + 7) / 8;
        const size_t bm_bytes = (n_values + 31) / 32 * 4 + 16;
        DevBuf raw(raw_bytes + 32), bm(bm_bytes), dseg(sizeof(Seg)), dbit(sizeof(BitSeg)), cnt(64);
        CU(cudaMemset(bm.p, 0, bm_bytes));
        CU(cudaMemset(raw.p, 0, raw_bytes + 32));
        Seg s{};
        s.in = (uint64_t)(uintptr_t)sr.in.p;
        s.out = (uint64_t)(uintptr_t)raw.p;
        s.in_len = (uint32_t)in_len;
        s.n_values = (uint32_t)n_values;
        s.cnt_idx = -1;
        s.start_idx = -1;
        s.out_kind = OUT_I8;
        s.aux = 1;  // bit mode, bit_skip 0
        CU(cudaMemcpy(dseg.p, &s, sizeof(s), cudaMemcpyHostToDevice));
        BitSeg b{};
        b.src = (uint64_t)(uintptr_t)raw.p;
        b.dst = (uint64_t)(uintptr_t)bm.p;
        b.n_bits = (uint32_t)n_values;
        b.cnt_idx = -1;
        b.start_idx = -1;
        b.popc_out = 0;
        CU(cudaMemcpy(dbit.p, &b, sizeof(b), cudaMemcpyHostToDevice));
        int rc = launch_byte_rle((Seg*)dseg.p, 1, nullptr, nullptr, (uint32_t*)sr.err.p, nullptr, 0);
        if (rc) fail(ORCB_CUDA, cudaGetErrorString((cudaError_t)rc));
        rc = launch_bits((BitSeg*)dbit.p, 1, (uint32_t*)cnt.p, nullptr, 0);
        if (rc) fail(ORCB_CUDA, cudaGetErrorString((cudaError_t)rc));
        sr.check();
        if (n_values) CU(cudaMemcpy(out_bitmap, bm.p, raw_bytes, cudaMemcpyDeviceToHost));
    });
}

int orcb_decode_varint128(int device, const uint8_t* in, size_t in_len, uint8_t* out16, size_t n_values) {
    return guarded([&] {
        StreamRun sr(device, in, in_len);
        DevBuf dout(n_values * 16 + 16), dseg(sizeof(Seg));
        Seg s{};
        s.in = (uint64_t)(uintptr_t)sr.in.p;
        s.out = (uint64_t)(uintptr_t)dout.p;
        s.in_len = (uint32_t)in_len;
        s.n_values = (uint32_t)n_values;
        s.cnt_idx = -1;
        s.start_idx = -1;
        CU(cudaMemcpy(dseg.p, &s, sizeof(s), cudaMemcpyHostToDevice));
        int rc = launch_varint128((Seg*)dseg.p, 1, nullptr, nullptr, (uint32_t*)sr.err.p, nullptr, 0);
        if (rc) fail(ORCB_CUDA, cudaGetErrorString((cudaError_t)rc));
        sr.check();
        if (n_values) CU(cudaMemcpy(out16, dout.p, n_values * 16, cudaMemcpyDeviceToHost));
    });
}

int orcb_host_decompress_section(int compression_kind, const uint8_t* in, size_t in_len, size_t block_size, uint8_t* out,
                                 size_t out_cap, size_t* out_len) {
    return guarded([&] {
        std::vector<uint8_t> v = host_decompress_section(compression_kind, block_size, in, in_len);
        *out_len = v.size();
        if (v.size() > out_cap) fail(ORCB_INVALID_ARGUMENT, "output too small");
        if (!v.empty()) memcpy(out, v.data(), v.size());
    });
}

int orcb_decompress_stream(int device, int compression_kind, const uint8_t* in, size_t in_len, size_t block_size,
                           uint8_t* out, size_t out_cap, size_t* out_len) {
    return guarded([&] {
        if (compression_kind < C_NONE || compression_kind > C_ZSTD) fail(ORCB_INVALID_ARGUMENT, "unknown compression kind");
        if (compression_kind == C_NONE) {
            if (in_len > out_cap) fail(ORCB_INVALID_ARGUMENT, "output too small");
            memcpy(out, in, in_len);
            *out_len = in_len;
            return;
        }
        FileMeta fm;
        fm.data = in;
        fm.len = in_len;
        fm.compression = compression_kind;
        fm.block_size = block_size;
        std::vector<ChunkInfo> chunks = fm.chunk_table(0, in_len);
        StreamRun sr(device, in, in_len);
        const size_t cap = chunks.size() * block_size + 256;
        if (chunks.size() * block_size > out_cap) fail(ORCB_INVALID_ARGUMENT, "out_cap must be >= chunks * block_size");
        DevBuf dout(cap), ddesc(chunks.size() * sizeof(ChunkDesc) + 16), dlens(chunks.size() * 4 + 16), dctr(32);
        CU(cudaMemset(dctr.p, 0, 32));
        std::vector<ChunkDesc> descs(chunks.size());
        for (size_t i = 0; i < chunks.size(); i++) {
            ChunkDesc& d = descs[i];
            memset(&d, 0, sizeof(d));
            d.src = (uint64_t)(uintptr_t)sr.in.p + chunks[i].src_off;
            d.dst = (uint64_t)(uintptr_t)dout.p + i * block_size;
            d.src_len = chunks[i].src_len;
            d.dst_cap = (uint32_t)block_size;
            d.expect_len = -1;
            d.codec = chunks[i].original ? 0 : (uint8_t)compression_kind;
            d.id = (uint32_t)i;
        }
        if (!chunks.empty()) CU(cudaMemcpy(ddesc.p, descs.data(), descs.size() * sizeof(ChunkDesc), cudaMemcpyHostToDevice));
        const bool timing = getenv("ORCB_STREAM_TIMING") != nullptr;  // kernel time to stderr (tools/snappy_probe.py)
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (timing) {
            CU(cudaEventCreate(&e0));
            CU(cudaEventCreate(&e1));
            CU(cudaEventRecord(e0, 0));
        }
        const bool bits = compression_kind == C_ZLIB || compression_kind == C_ZSTD || compression_kind == C_LZO;
        const uint32_t nc = (uint32_t)chunks.size();
        int rc = launch_decompress((ChunkDesc*)ddesc.p, nc, bits ? nc : 0u, bits ? 1u << compression_kind : 0u, compression_kind == C_SNAPPY ? nc : 0u,
                                   (uint32_t*)sr.err.p, (uint32_t*)dlens.p, (uint32_t*)dctr.p, nullptr, 0);
        if (rc) fail(ORCB_CUDA, cudaGetErrorString((cudaError_t)rc));
        if (timing) {
            float ms = 0;
            CU(cudaEventRecord(e1, 0));
            CU(cudaEventSynchronize(e1));
            CU(cudaEventElapsedTime(&ms, e0, e1));
            fprintf(stderr, "orcb_decompress_stream: %zu chunks, kernel %.3f ms\n", chunks.size(), ms);
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
        sr.check();
        std::vector<uint32_t> lens(chunks.size());
        if (!chunks.empty()) CU(cudaMemcpy(lens.data(), dlens.p, chunks.size() * 4, cudaMemcpyDeviceToHost));
        size_t o = 0;
        for (size_t i = 0; i < chunks.size(); i++) {
            if (lens[i]) CU(cudaMemcpy(out + o, (uint8_t*)dout.p + i * block_size, lens[i], cudaMemcpyDeviceToHost));
            o += lens[i];
        }
        *out_len = o;
    });
}

}  // extern "C"
