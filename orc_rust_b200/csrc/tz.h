// Writer time zones: TZif (RFC 8536) tables flattened for the device.  See tz.cc.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace orcb {

struct ZoneTable {
    std::vector<int64_t> at;   // UTC instants of the transitions, ascending
    std::vector<int32_t> off;  // UTC offset (seconds east) in force from at[i] on
    int32_t first_off = 0;     // offset before at[0]
};

// `name` as written in the stripe footer (IANA name or one of the old link names).
bool load_zone_table(const std::string& name, ZoneTable& out, std::string& why);
int32_t zone_offset_at(const ZoneTable& z, int64_t utc);
// wall clock -> instant; false when the wall clock is skipped or repeated (the reference unwraps a Single)
bool zone_local_to_utc(const ZoneTable& z, int64_t local, int64_t& utc);

}  // namespace orcb
