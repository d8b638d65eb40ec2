// Minimal protobuf wire-format cursor (the reference uses prost-generated src/proto.rs; only the
// handful of messages the decode planner needs are walked here, by field number).
#pragma once
#include "common.h"

namespace orcb {

struct PbField {
    uint32_t number;
    uint32_t wire;        // 0 varint, 1 fixed64, 2 bytes, 5 fixed32
    uint64_t value;       // varint / fixed value
    const uint8_t* data;  // bytes payload
    size_t len;
};

class PbCursor {
  public:
    PbCursor(const uint8_t* p, size_t n) : p_(p), end_(p + n) {}
    bool next(PbField& f) {
        if (p_ >= end_) return false;
        uint64_t key = varint();
        if ((key >> 32) != 0 || (key >> 3) == 0) fail(ORCB_DECODE_PROTO, "invalid protobuf key");  // prost: key > u32, tag 0
        f.number = (uint32_t)(key >> 3);
        f.wire = (uint32_t)(key & 7);
        f.data = nullptr;
        f.len = 0;
        f.value = 0;
        switch (f.wire) {
            case 0: f.value = varint(); break;
            case 1:
                need(8);
                for (int i = 0; i < 8; i++) f.value |= (uint64_t)p_[i] << (8 * i);
                p_ += 8;
                break;
            case 2: {
                uint64_t n = varint();
                need(n);
                f.data = p_;
                f.len = (size_t)n;
                p_ += n;
                break;
            }
            case 5:
                need(4);
                for (int i = 0; i < 4; i++) f.value |= (uint64_t)p_[i] << (8 * i);
                p_ += 4;
                break;
            default: fail(ORCB_DECODE_PROTO, "unsupported protobuf wire type");
        }
        return true;
    }
    // repeated scalar that may be packed (wire 2) or not (wire 0)
    static void packed_u64(const PbField& f, std::vector<uint64_t>& out) {
        if (f.wire == 0) {
            out.push_back(f.value);
            return;
        }
        PbCursor c(f.data, f.len);
        while (c.p_ < c.end_) out.push_back(c.varint());
    }

  private:
    void need(uint64_t n) {
        if ((uint64_t)(end_ - p_) < n) fail(ORCB_DECODE_PROTO, "truncated protobuf message");
    }
    uint64_t varint() {
        uint64_t r = 0;
        int s = 0;
        for (;;) {
            need(1);
            uint8_t b = *p_++;
            if (s < 64) r |= (uint64_t)(b & 0x7f) << s;
            s += 7;
            if (!(b & 0x80)) return r;
            if (s > 70) fail(ORCB_DECODE_PROTO, "varint too long");
        }
    }
    const uint8_t* p_;
    const uint8_t* end_;
};

}  // namespace orcb
