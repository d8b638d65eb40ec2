// File tail / stripe footer / row-index metadata, parsed on the host.
// Restates (as a planner input, not as decoders): src/reader/metadata.rs:63-263, src/stripe.rs:38-182,
// src/column.rs:40-59, src/schema.rs:390-495, src/row_index.rs:204-289 of the reference.
#pragma once
#include <memory>
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
};

struct FileMeta {
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

    StripeFooter read_stripe_footer(uint32_t stripe) const;
    // positions of every row-index entry of `column` in `stripe` (empty if no ROW_INDEX stream)
    std::vector<std::vector<uint64_t>> read_row_index(const StripeInfo& si, const StripeFooter& sf,
                                                      uint32_t column) const;
    std::vector<ChunkInfo> chunk_table(uint64_t stream_off, uint64_t stream_len) const;
};

// Parses PostScript + Footer. Throws OrcException.  Zlib/Zstd/LZO files are rejected here with
// ORCB_UNSUPPORTED_DEVICE_CODEC (their footers are never inflated on the host).
void parse_file_tail(FileMeta& fm);

// Host-side chunk decompression for METADATA sections only (footer, stripe footer, row index).
std::vector<uint8_t> host_decompress_section(int compression, uint64_t block_size, const uint8_t* in, size_t len);

}  // namespace orcb
