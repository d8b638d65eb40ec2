// File tail / stripe footer / row-index metadata, parsed on the host.
// Restates (as a planner input, not as decoders): src/reader/metadata.rs:63-263, src/stripe.rs:38-182,
// src/column.rs:40-59, src/schema.rs:390-495, src/row_index.rs:204-289 of the reference.
#pragma once
#include <memory>
#include <mutex>
#include <unordered_map>
#include <string>
#include <utility>
#include <vector>

#include "common.h"

namespace orcb {

struct OrcType {
    int kind = T_BOOLEAN;
    std::vector<uint32_t> subtypes;
    std::vector<std::string> field_names;
    uint32_t max_length = 0, precision = 0, scale = 0;
};

struct StripeInfo {
    uint64_t offset = 0, index_length = 0, data_length = 0, footer_length = 0, rows = 0;
};

struct StreamInfo {
    int kind = 0;
    uint32_t column = 0;
    uint64_t length = 0;
    uint64_t offset = 0;  // absolute file offset (running sum, src/stripe.rs:154-165)
};

struct ColumnEncoding {
    int kind = E_DIRECT;
    uint32_t dict_size = 0;
};

struct StripeFooter {
    std::vector<StreamInfo> streams;
    std::vector<ColumnEncoding> encodings;
    bool has_tz = false;
    std::string tz;
    const StreamInfo* find(uint32_t column, int kind) const {
        for (auto& s : streams)
            if (s.column == column && s.kind == kind) return &s;
        return nullptr;
    }
};

// One compression chunk of a stream (src/compression.rs:113-123, 244-275), located by the host so the
// device kernels start from a flat table instead of walking headers serially.
struct ChunkInfo {
    uint64_t src_off;   // offset of the payload (after the 3-byte header) relative to stream start
    uint32_t src_len;   // payload bytes
    uint32_t hdr_off;   // offset of the 3-byte header relative to stream start
    bool original;
    int64_t dst_len;    // decompressed bytes if known on the host (original: = src_len; snappy: preamble), else -1
    bool guess = false; // dst_len is what a Zstandard frame header announces: planned on, checked on the device
};

// ColumnStatistics of one row group (src/statistics.rs:21-143); `kind` is the TypeStatistics variant the reference
// would build: none when number_of_values == 0, else the first sub-message present in its order of tests
enum StatsKind : int { ST_NONE = 0, ST_INTEGER, ST_DOUBLE, ST_STRING, ST_BUCKET, ST_DECIMAL, ST_DATE, ST_BINARY, ST_TIMESTAMP, ST_COLLECTION };
struct ColumnStats {
    uint64_t number_of_values = 0;
    bool has_null = false;
    int kind = ST_NONE;
    int64_t imin = 0, imax = 0;  // Integer, Date; Timestamp: minimumUtc / maximumUtc
    double dmin = 0, dmax = 0;
    std::string smin, smax;      // String: minimum / maximum, else lowerBound / upperBound; Decimal: minimum / maximum
    bool exact_min = false, exact_max = false;
    uint64_t true_count = 0;     // Bucket: count[0]
};
// BloomFilter (src/bloom_filter.rs:28-75)
struct BloomBits {
    uint32_t num_hash_functions = 3;
    std::vector<uint64_t> bitset;
};
// RowGroupEntry (src/row_index.rs:33-61)
struct RowGroupEntry {
    bool has_stats = false;
    ColumnStats stats;
    bool has_bloom = false;
    BloomBits bloom;
};
// A place where the reference panics instead of returning an error (asserts, slice indexing); surfaces as
// ORCB_UNEXPECTED with the panic message
struct ReferencePanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// Decompressed sizes of chunks whose size only the device can find (Zlib and LZ4 carry none in their framing), keyed by
// the file offset of the chunk header.  Filled by the first job that meets a chunk that does not fill its block; shared
// by the clones of a file handle.
struct ChunkSizeCache {
    std::mutex mu;
    std::unordered_map<uint64_t, uint32_t> size;
};

// A byte range of a file opened through read callbacks (ChunkReader::get_bytes, src/reader/mod.rs:27-46): one read per
// stripe (index + data + footer together, where the reference reads stream by stream, src/stripe.rs:161), or the file
// tail.  Lives as long as someone holds it: the job that stages from it, the reader that plans with it.
struct RangeBuf {
    uint64_t off = 0, len = 0;
    uint8_t* p = nullptr;
    size_t pinned_cap = 0;            // > 0: from the pinned-buffer cache (asynchronous H2D copies)
    std::vector<uint8_t> heap;        // without a device
    ~RangeBuf();
};
typedef int (*OrcbReadAtFn)(void* ctx, uint64_t offset, uint64_t length, uint8_t* dst);
struct RangeSource {
    OrcbReadAtFn read_at = nullptr;
    void* ctx = nullptr;
    std::mutex mu;
    std::vector<std::weak_ptr<RangeBuf>> live;
    uint64_t reads = 0, bytes_read = 0;
};

struct FileMeta {
    std::shared_ptr<ChunkSizeCache> chunk_sizes = std::make_shared<ChunkSizeCache>();
    // callback-backed files: `data` is null, bytes come through `source` (shared by the clones of a handle)
    std::shared_ptr<RangeSource> source;
    std::shared_ptr<RangeBuf> tail;   // footer + metadata + postscript
    // Pointer q such that q[off] is the file's byte at offset `off`.  Memory files: `data`.  Callback files: the range
    // that holds `off` must be alive (load_stripe / the tail), else Unexpected.
    // base pointer (indexable by file offset) of loaded bytes that hold [off, off + need)
    const uint8_t* base_for(uint64_t off, uint64_t need = 1) const;
    // the stripe's bytes (index, data, footer) in one read; null for memory files
    std::shared_ptr<RangeBuf> load_stripe(uint32_t stripe) const;
    std::shared_ptr<RangeBuf> load_range(uint64_t off, uint64_t len) const;
    const uint8_t* data = nullptr;
    size_t len = 0;
    std::vector<uint8_t> owned;   // when opened from a path without pinned memory
    void* pinned = nullptr;       // cudaHostAlloc'd copy when opened from a path with a device present
    int compression = C_NONE;
    uint64_t block_size = 256 * 1024;  // src/compression.rs:31
    std::vector<OrcType> types;
    std::vector<StripeInfo> stripes;
    std::vector<std::pair<std::string, std::string>> user_metadata;
    std::vector<std::pair<std::string, uint32_t>> root_columns;  // (name, column id)
    uint64_t num_rows = 0;
    int64_t row_index_stride = -1;
    std::string format_version;  // PostScript.version joined with '.', "" when absent (src/reader/metadata.rs:119-127)

    StripeFooter read_stripe_footer(uint32_t stripe) const;
    // positions of every row-index entry of `column` in `stripe` (empty if no ROW_INDEX stream)
    std::vector<std::vector<uint64_t>> read_row_index(const StripeInfo& si, const StripeFooter& sf,
                                                      uint32_t column) const;
    // statistics and Bloom filter of every row-index entry of `column`; *present = false when the stripe has no
    // ROW_INDEX stream for it (src/row_index.rs:204-289).  Throws what the reference returns as an error.
    std::vector<RowGroupEntry> read_row_group_entries(const StripeFooter& sf, uint32_t column, bool* present) const;
    std::vector<ChunkInfo> chunk_table(uint64_t stream_off, uint64_t stream_len) const;
};

// Parses PostScript + Footer. Throws OrcException.  The sections are decompressed on the host whatever the file's
// compression kind (host_decompress_section); data streams never are.
void parse_file_tail(FileMeta& fm);

// Host-side chunk decompression for METADATA sections only (footer, stripe footer, row index).
std::vector<uint8_t> host_decompress_section(int compression, uint64_t block_size, const uint8_t* in, size_t len);

}  // namespace orcb
