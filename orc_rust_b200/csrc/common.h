// Shared host-side definitions for the B200 ORC stripe decoder.
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/orc_b200.h"

namespace orcb {

// Status carried through the host code; converted to an integer + message at the C ABI.
struct OrcException : public std::exception {
    int code;
    std::string msg;
    OrcException(int c, std::string m) : code(c), msg(std::move(m)) {}
    const char* what() const noexcept override { return msg.c_str(); }
};

[[noreturn]] inline void fail(int code, const std::string& msg) { throw OrcException(code, msg); }

// ORC protobuf enums (format/orc_proto.proto in the reference: stream kinds :125-142, column
// encodings :150-155, type kinds :200-220, compression kinds :383-390).
enum TypeKind : int {
    T_BOOLEAN = 0, T_BYTE, T_SHORT, T_INT, T_LONG, T_FLOAT, T_DOUBLE, T_STRING, T_BINARY, T_TIMESTAMP,
    T_LIST, T_MAP, T_STRUCT, T_UNION, T_DECIMAL, T_DATE, T_VARCHAR, T_CHAR, T_TIMESTAMP_INSTANT
};
enum StreamKind : int {
    S_PRESENT = 0, S_DATA = 1, S_LENGTH = 2, S_DICTIONARY_DATA = 3, S_DICTIONARY_COUNT = 4, S_SECONDARY = 5,
    S_ROW_INDEX = 6, S_BLOOM_FILTER = 7, S_BLOOM_FILTER_UTF8 = 8
};
enum EncodingKind : int { E_DIRECT = 0, E_DICTIONARY = 1, E_DIRECT_V2 = 2, E_DICTIONARY_V2 = 3 };
enum CompressionKind : int { C_NONE = 0, C_ZLIB = 1, C_SNAPPY = 2, C_LZO = 3, C_LZ4 = 4, C_ZSTD = 5 };

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace orcb
