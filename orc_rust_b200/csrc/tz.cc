// Writer time zones (src/array_decoder/timestamp.rs:128-147, 242-286).  The reference converts with chrono-tz's
// compiled-in IANA database; here the tables come from the system's TZif files (RFC 8536) and are flattened
// into (instant, UTC offset after it) pairs so that the device only needs a binary search.
#include "tz.h"

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>

namespace orcb {
namespace {

// the "backward" links most tz installations leave out
const char* const kAliases[][2] = {
    {"US/Pacific", "America/Los_Angeles"},   {"US/Eastern", "America/New_York"},     {"US/Central", "America/Chicago"},
    {"US/Mountain", "America/Denver"},       {"US/Alaska", "America/Anchorage"},     {"US/Hawaii", "Pacific/Honolulu"},
    {"US/Arizona", "America/Phoenix"},       {"US/Michigan", "America/Detroit"},     {"US/Aleutian", "America/Adak"},
    {"US/East-Indiana", "America/Indiana/Indianapolis"},                               {"US/Samoa", "Pacific/Pago_Pago"},
    {"US/Indiana-Starke", "America/Indiana/Knox"},                                     {"Canada/Pacific", "America/Vancouver"},
    {"Canada/Mountain", "America/Edmonton"}, {"Canada/Central", "America/Winnipeg"}, {"Canada/Eastern", "America/Toronto"},
    {"Canada/Atlantic", "America/Halifax"},  {"Canada/Newfoundland", "America/St_Johns"},
    {"Asia/Calcutta", "Asia/Kolkata"},       {"Asia/Saigon", "Asia/Ho_Chi_Minh"},    {"Asia/Katmandu", "Asia/Kathmandu"},
    {"Asia/Rangoon", "Asia/Yangon"},         {"Europe/Kiev", "Europe/Kyiv"},         {"Australia/ACT", "Australia/Sydney"},
    {"Australia/NSW", "Australia/Sydney"},   {"Australia/Victoria", "Australia/Melbourne"},
    {"Australia/Queensland", "Australia/Brisbane"},                                    {"Australia/West", "Australia/Perth"},
    {"Australia/South", "Australia/Adelaide"},                                         {"Australia/North", "Australia/Darwin"},
    {"Australia/Tasmania", "Australia/Hobart"},                                        {"Brazil/East", "America/Sao_Paulo"},
    {"Mexico/General", "America/Mexico_City"}, {"Chile/Continental", "America/Santiago"}, {"Cuba", "America/Havana"},
    {"Egypt", "Africa/Cairo"},               {"Eire", "Europe/Dublin"},              {"GB", "Europe/London"},
    {"Hongkong", "Asia/Hong_Kong"},          {"Iceland", "Atlantic/Reykjavik"},      {"Iran", "Asia/Tehran"},
    {"Israel", "Asia/Jerusalem"},            {"Jamaica", "America/Jamaica"},         {"Japan", "Asia/Tokyo"},
    {"NZ", "Pacific/Auckland"},              {"PRC", "Asia/Shanghai"},               {"Poland", "Europe/Warsaw"},
    {"Portugal", "Europe/Lisbon"},           {"ROC", "Asia/Taipei"},                 {"ROK", "Asia/Seoul"},
    {"Singapore", "Asia/Singapore"},         {"Turkey", "Europe/Istanbul"},          {"W-SU", "Europe/Moscow"},
};

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    out.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
    return true;
}

bool find_zone_file(const std::string& name, std::vector<uint8_t>& out) {
    if (name.empty() || name[0] == '/' || name.find("..") != std::string::npos) return false;
    std::vector<std::string> dirs;
    if (const char* d = getenv("TZDIR")) dirs.push_back(d);
    for (const char* d : {"/usr/share/zoneinfo", "/usr/lib/zoneinfo", "/usr/share/lib/zoneinfo", "/etc/zoneinfo"}) dirs.push_back(d);
    std::vector<std::string> names{name};
    for (auto& a : kAliases)
        if (name == a[0]) names.push_back(a[1]);
    for (auto& n : names)
        for (auto& d : dirs)
            if (read_file(d + "/" + n, out) && out.size() > 44 && memcmp(out.data(), "TZif", 4) == 0) return true;
    return false;
}

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
int64_t be64(const uint8_t* p) { return (int64_t)(((uint64_t)be32(p) << 32) | be32(p + 4)); }

// days since 1970-01-01 of a proleptic Gregorian date
int64_t days_from_civil(int64_t y, int m, int d) {
    y -= m <= 2;
    const int64_t era = (y >= 0 ? y : y - 399) / 400;
    const int64_t yoe = y - era * 400;
    const int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}
int year_of(int64_t unix_secs) {
    int64_t z = unix_secs / 86400 - (unix_secs % 86400 < 0) + 719468;
    const int64_t era = (z >= 0 ? z : z - 146096) / 146097;
    const int64_t doe = z - era * 146097;
    const int64_t yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    const int64_t doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const int64_t mp = (5 * doy + 2) / 153;
    return (int)(yoe + era * 400 + (mp >= 10));
}

// ---- POSIX TZ footer: std offset dst [offset] , Mm.w.d[/time] , Mm.w.d[/time]
struct Rule {
    int m = 0, w = 0, d = 0;
    int64_t time = 7200;
};
bool parse_name(const char*& p) {
    if (*p == '<') {
        while (*p && *p != '>') p++;
        if (*p != '>') return false;
        p++;
        return true;
    }
    const char* s = p;
    while ((*p >= 'A' && *p <= 'Z') || (*p >= 'a' && *p <= 'z')) p++;
    return p - s >= 3;
}
bool parse_hms(const char*& p, int64_t& secs) {
    int sign = 1;
    if (*p == '+') p++;
    else if (*p == '-') { sign = -1; p++; }
    if (*p < '0' || *p > '9') return false;
    int64_t v[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) {
        while (*p >= '0' && *p <= '9') v[i] = v[i] * 10 + (*p++ - '0');
        if (*p != ':') break;
        p++;
    }
    secs = sign * (v[0] * 3600 + v[1] * 60 + v[2]);
    return true;
}
bool parse_rule(const char*& p, Rule& r) {
    if (*p != 'M') return false;  // Jn / n day-of-year forms are not used by the IANA database
    p++;
    auto num = [&](int& out) {
        if (*p < '0' || *p > '9') return false;
        out = 0;
        while (*p >= '0' && *p <= '9') out = out * 10 + (*p++ - '0');
        return true;
    };
    if (!num(r.m) || *p++ != '.' || !num(r.w) || *p++ != '.' || !num(r.d)) return false;
    if (*p == '/') {
        p++;
        if (!parse_hms(p, r.time)) return false;
    }
    return r.m >= 1 && r.m <= 12 && r.w >= 1 && r.w <= 5 && r.d >= 0 && r.d <= 6;
}
// local seconds since the epoch (wall clock of the rule) of the rule's instant in year y
int64_t rule_local(const Rule& r, int y) {
    const int64_t first = days_from_civil(y, r.m, 1);
    const int wd = (int)(((first % 7) + 11) % 7);  // 1970-01-01 was a Thursday (4)
    int64_t day = first + ((r.d - wd + 7) % 7) + 7 * (r.w - 1);
    const int64_t next = r.m == 12 ? days_from_civil(y + 1, 1, 1) : days_from_civil(y, r.m + 1, 1);
    while (day >= next) day -= 7;  // week 5 = the last one
    return day * 86400 + r.time;
}

}  // namespace

bool load_zone_table(const std::string& name, ZoneTable& out, std::string& why) {
    std::vector<uint8_t> f;
    if (!find_zone_file(name, f)) {
        why = "no TZif file for zone '" + name + "' under /usr/share/zoneinfo (or $TZDIR)";
        return false;
    }
    auto bad = [&](const char* m) {
        why = std::string("zone '") + name + "': " + m;
        return false;
    };
    size_t pos = 0;
    auto header = [&](uint32_t cnt[6]) {
        if (pos + 44 > f.size() || memcmp(&f[pos], "TZif", 4) != 0) return false;
        for (int i = 0; i < 6; i++) cnt[i] = be32(&f[pos + 20 + 4 * i]);
        pos += 44;
        return true;
    };
    uint32_t c[6];  // isutcnt, isstdcnt, leapcnt, timecnt, typecnt, charcnt
    if (!header(c)) return bad("not a TZif file");
    const uint8_t version = f[4];
    size_t tsz = 4;
    if (version >= '2') {
        pos += (size_t)c[3] * 5 + (size_t)c[4] * 6 + c[5] + (size_t)c[2] * 8 + c[1] + c[0];
        if (!header(c)) return bad("truncated TZif v2 block");
        tsz = 8;
    }
    const uint32_t timecnt = c[3], typecnt = c[4];
    if (typecnt == 0) return bad("no local time types");
    const size_t need = (size_t)timecnt * (tsz + 1) + (size_t)typecnt * 6 + c[5] + (size_t)c[2] * (tsz + 4) + c[1] + c[0];
    if (pos + need > f.size()) return bad("truncated TZif data");
    const uint8_t* times = &f[pos];
    const uint8_t* idx = times + (size_t)timecnt * tsz;
    const uint8_t* types = idx + timecnt;
    auto utoff = [&](uint32_t t) { return (int32_t)be32(types + 6 * t); };
    auto isdst = [&](uint32_t t) { return types[6 * t + 4] != 0; };
    out.at.clear();
    out.off.clear();
    // before the first transition: the first standard-time type (as CPython's zoneinfo), else type 0
    out.first_off = utoff(0);
    for (uint32_t t = 0; t < typecnt; t++)
        if (!isdst(t)) { out.first_off = utoff(t); break; }
    for (uint32_t i = 0; i < timecnt; i++) {
        if (idx[i] >= typecnt) return bad("transition type out of range");
        out.at.push_back(tsz == 8 ? be64(times + 8 * i) : (int64_t)(int32_t)be32(times + 4 * i));
        out.off.push_back(utoff(idx[i]));
    }
    pos += need;
    // footer rule: transitions after the table, generated up to the year 2500
    if (tsz == 8 && pos < f.size() && f[pos] == '\n') {
        const char* b = (const char*)&f[pos + 1];
        const char* e = (const char*)memchr(b, '\n', f.size() - pos - 1);
        if (e && e > b) {
            const std::string tzs(b, e);
            const char* p = tzs.c_str();
            int64_t std_west = 0, dst_west = 0;
            Rule rs, re;
            if (parse_name(p) && parse_hms(p, std_west)) {
                const int32_t std_off = (int32_t)-std_west;
                if (*p == 0) {
                    // fixed offset from the last transition on
                    if (out.at.empty()) out.first_off = std_off;
                } else if (parse_name(p)) {
                    dst_west = std_west - 3600;
                    if (*p != ',' && !parse_hms(p, dst_west)) return bad("unparsable TZ footer");
                    const int32_t dst_off = (int32_t)-dst_west;
                    if (*p == ',' && parse_rule(++p, rs) && *p == ',' && parse_rule(++p, re)) {
                        const bool none = out.at.empty();
                        const int64_t last = none ? INT64_MIN : out.at.back();
                        auto add = [&](int64_t at, int32_t off) {
                            if (at > last) { out.at.push_back(at); out.off.push_back(off); }
                        };
                        for (int y = none ? 1970 : year_of(last); y <= 2500; y++) {
                            const int64_t s = rule_local(rs, y) - std_off;  // DST starts: rule is in standard time
                            const int64_t t = rule_local(re, y) - dst_off;  // DST ends: rule is in daylight time
                            if (s < t) { add(s, dst_off); add(t, std_off); }
                            else { add(t, std_off); add(s, dst_off); }
                        }
                    }
                }
            }
        }
    }
    for (size_t i = 1; i < out.at.size(); i++)
        if (out.at[i] < out.at[i - 1]) return bad("transitions out of order");
    return true;
}

int32_t zone_offset_at(const ZoneTable& z, int64_t utc) {
    const size_t k = std::upper_bound(z.at.begin(), z.at.end(), utc) - z.at.begin();
    return k == 0 ? z.first_off : z.off[k - 1];
}

bool zone_local_to_utc(const ZoneTable& z, int64_t local, int64_t& utc) {
    // candidates: every distinct offset that makes local - off map back to itself
    int found = 0;
    int32_t offs[3] = {zone_offset_at(z, local - 86400), zone_offset_at(z, local), zone_offset_at(z, local + 86400)};
    for (int i = 0; i < 3; i++) {
        bool dup = false;
        for (int j = 0; j < i; j++) dup |= offs[j] == offs[i];
        if (dup) continue;
        const int64_t u = local - offs[i];
        if (zone_offset_at(z, u) == offs[i]) {
            utc = u;
            found++;
        }
    }
    return found == 1;  // chrono's LocalResult::Single; the reference unwraps
}

}  // namespace orcb
