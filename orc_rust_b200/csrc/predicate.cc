#include "predicate.h"

#include <cmath>
#include <cstring>
#include <map>

namespace orcb {

// ---------------------------------------------------------------------------------------------
// Bloom filters (src/bloom_filter.rs)
// ---------------------------------------------------------------------------------------------
// Thomas Wang's 64-bit mix, as ORC's Java and C++ writers use it for integers (:136-149)
uint64_t bloom_hash_long(int64_t v) {
    uint64_t key = (uint64_t)v;
    key = ~key + (key << 21);
    key ^= (uint64_t)((int64_t)key >> 24);  // the reference shifts an i64: arithmetic
    key = key + (key << 3) + (key << 8);
    key ^= (uint64_t)((int64_t)key >> 14);
    key = key + (key << 2) + (key << 4);
    key ^= (uint64_t)((int64_t)key >> 28);
    key = key + (key << 31);
    return key;
}

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// Murmur3 x64 128-bit, first half, seed 104729 (:182-230)
uint64_t bloom_hash_bytes(const uint8_t* p, size_t n) {
    const uint64_t C1 = 0x87c37b91114253d5ull, C2 = 0x4cf5ad432745937full;
    uint64_t h1 = 104729;
    const size_t nblocks = n / 8;
    for (size_t i = 0; i < nblocks; i++) {
        uint64_t k1;
        memcpy(&k1, p + 8 * i, 8);
        k1 *= C1;
        k1 = rotl64(k1, 31);
        k1 *= C2;
        h1 ^= k1;
        h1 = rotl64(h1, 27);
        h1 = h1 * 5 + 1390208809ull;
    }
    const uint8_t* tail = p + 8 * nblocks;
    const size_t t = n - 8 * nblocks;
    if (t) {
        uint64_t k1 = 0;
        for (size_t j = 0; j < t; j++) k1 ^= (uint64_t)tail[j] << (8 * j);
        k1 *= C1;
        k1 = rotl64(k1, 31);
        k1 *= C2;
        h1 ^= k1;
    }
    h1 ^= (uint64_t)n;
    h1 ^= h1 >> 33;
    h1 *= 0xff51afd7ed558ccdull;
    h1 ^= h1 >> 33;
    h1 *= 0xc4ceb9fe1a85ec53ull;
    h1 ^= h1 >> 33;
    return h1;
}

// double hashing over the two 32-bit halves, a negative combination flipped with `!` (:109-133)
bool bloom_test_hash(const BloomBits& b, uint64_t hash64) {
    const uint64_t bit_count = (uint64_t)b.bitset.size() * 64;
    if (bit_count == 0) return true;
    const uint32_t h1 = (uint32_t)hash64, h2 = (uint32_t)(hash64 >> 32);
    // 64-bit counter: `1..=k` ends for k = u32::MAX (a 32-bit one would wrap and never leave the loop)
    for (uint64_t i = 1; i <= b.num_hash_functions; i++) {
        uint32_t combined = h1 + (uint32_t)i * h2;  // i32 wrapping arithmetic
        if ((int32_t)combined < 0) combined = ~combined;
        const uint64_t bit = (uint64_t)combined % bit_count;
        if (!((b.bitset[bit / 64] >> (bit % 64)) & 1)) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// the C form of a predicate
// ---------------------------------------------------------------------------------------------
static Predicate node_at(const OrcbPredicateNode* nodes, uint32_t n, uint32_t& pos, int depth) {
    if (pos >= n) fail(ORCB_INVALID_ARGUMENT, "predicate nodes end inside a tree");
    if (depth > 256) fail(ORCB_INVALID_ARGUMENT, "predicate nested too deeply");
    const OrcbPredicateNode& c = nodes[pos++];
    Predicate p;
    p.kind = c.kind;
    switch (c.kind) {
        case ORCB_PRED_COMPARISON:
            if (c.op < ORCB_OP_EQ || c.op > ORCB_OP_GE) fail(ORCB_INVALID_ARGUMENT, "unknown comparison operator");
            if (c.value_type < ORCB_VAL_BOOLEAN || c.value_type > ORCB_VAL_UTF8) fail(ORCB_INVALID_ARGUMENT, "unknown predicate value type");
            p.op = c.op;
            p.vtype = c.value_type;
            p.vnull = c.value_is_null != 0;
            p.i = c.i64;
            p.f = c.f64;
            if (c.value_type == ORCB_VAL_UTF8 && !p.vnull) {
                if (!c.str && c.str_len) fail(ORCB_INVALID_ARGUMENT, "NULL string value");
                p.s.assign((const char*)c.str, (size_t)c.str_len);
            }
            // fallthrough: the column name
        case ORCB_PRED_IS_NULL:
        case ORCB_PRED_IS_NOT_NULL:
            if (!c.column) fail(ORCB_INVALID_ARGUMENT, "predicate without a column name");
            p.column = c.column;
            if (c.n_children) fail(ORCB_INVALID_ARGUMENT, "a leaf predicate has no children");
            break;
        case ORCB_PRED_NOT:
            if (c.n_children != 1) fail(ORCB_INVALID_ARGUMENT, "NOT takes one predicate");
            // fallthrough
        case ORCB_PRED_AND:
        case ORCB_PRED_OR:
            for (uint32_t k = 0; k < c.n_children; k++) p.children.push_back(node_at(nodes, n, pos, depth + 1));
            break;
        default: fail(ORCB_INVALID_ARGUMENT, "unknown predicate kind");
    }
    return p;
}

Predicate predicate_from_c(const OrcbPredicateNode* nodes, uint32_t n_nodes) {
    if (!nodes || !n_nodes) fail(ORCB_INVALID_ARGUMENT, "empty predicate");
    uint32_t pos = 0;
    Predicate p = node_at(nodes, n_nodes, pos, 0);
    if (pos != n_nodes) fail(ORCB_INVALID_ARGUMENT, "nodes left over after the predicate tree");
    return p;
}

// ---------------------------------------------------------------------------------------------
// evaluation against one stripe's row index (src/row_group_filter.rs)
// ---------------------------------------------------------------------------------------------
namespace {

struct StripeIndex {
    std::map<uint32_t, std::vector<RowGroupEntry>> columns;  // StripeRowIndex.columns
    size_t num_row_groups = 0;                                // ceil(total_rows / rows_per_group), 0 for stride 0
};

bool integer_cmp(int64_t lo, int64_t hi, int op, int64_t v) {  // :438-466
    switch (op) {
        case ORCB_OP_EQ: return lo <= v && v <= hi;
        case ORCB_OP_NE: return !(lo == v && hi == v);
        case ORCB_OP_LT: return lo < v;
        case ORCB_OP_LE: return lo <= v;
        case ORCB_OP_GT: return hi > v;
        default: return hi >= v;
    }
}

bool float_cmp(double lo, double hi, int op, double v) {  // :468-498
    const double EPS = 1e-9;
    switch (op) {
        case ORCB_OP_EQ: return (lo - EPS) <= v && v <= (hi + EPS);
        case ORCB_OP_NE: return !(std::fabs(lo - v) < EPS && std::fabs(hi - v) < EPS);
        case ORCB_OP_LT: return lo < v;
        case ORCB_OP_LE: return lo <= v;
        case ORCB_OP_GT: return hi > v;
        default: return hi >= v;
    }
}

// Rust `str` ordering is byte-wise, as std::string's (unsigned char traits)
bool string_cmp(const std::string& lo, const std::string& hi, int op, const std::string& v, bool exact_min, bool exact_max) {  // :500-530
    const bool min_le = lo < v || (lo == v && exact_min);
    const bool max_ge = hi > v || (hi == v && exact_max);
    switch (op) {
        case ORCB_OP_EQ: return min_le && max_ge;
        case ORCB_OP_LE: return min_le;
        case ORCB_OP_GE: return max_ge;
        case ORCB_OP_LT: return lo < v;
        case ORCB_OP_GT: return hi > v;
        default: return !(lo == hi && exact_min && exact_max && lo == v);
    }
}

bool is_int_value(const Predicate& p, int from) { return !p.vnull && p.vtype >= from && p.vtype <= ORCB_VAL_INT64; }

// evaluate_comparison_with_stats (:170-360)
bool stats_match(const ColumnStats& st, int op, const Predicate& p) {
    switch (st.kind) {
        case ST_NONE: fail(ORCB_UNEXPECTED, "Statistics missing type-specific information");
        case ST_INTEGER:
            if (!is_int_value(p, ORCB_VAL_INT8)) fail(ORCB_UNEXPECTED, "Type mismatch: expected integer value");
            return integer_cmp(st.imin, st.imax, op, p.i);
        case ST_DOUBLE:
            if (p.vnull || (p.vtype != ORCB_VAL_FLOAT32 && p.vtype != ORCB_VAL_FLOAT64))
                fail(ORCB_UNEXPECTED, "Type mismatch: expected float value");
            return float_cmp(st.dmin, st.dmax, op, p.f);
        case ST_STRING:
            if (p.vnull || p.vtype != ORCB_VAL_UTF8) fail(ORCB_UNEXPECTED, "Type mismatch: expected string value");
            return string_cmp(st.smin, st.smax, op, p.s, st.exact_min, st.exact_max);
        case ST_DATE:
            if (!is_int_value(p, ORCB_VAL_INT32)) fail(ORCB_UNEXPECTED, "Type mismatch: expected integer value for date");
            return integer_cmp(st.imin, st.imax, op, p.i);
        case ST_TIMESTAMP:
            if (!is_int_value(p, ORCB_VAL_INT64)) fail(ORCB_UNEXPECTED, "Type mismatch: expected integer value for timestamp");
            return integer_cmp(st.imin, st.imax, op, p.i);
        case ST_DECIMAL:
            if (p.vnull || p.vtype != ORCB_VAL_UTF8) fail(ORCB_UNEXPECTED, "Type mismatch: expected string value for decimal");
            return string_cmp(st.smin, st.smax, op, p.s, true, true);  // compared as text, like the reference
        case ST_BUCKET: {
            if (p.vnull || p.vtype != ORCB_VAL_BOOLEAN) fail(ORCB_UNEXPECTED, "Type mismatch: expected boolean value");
            const uint64_t false_count = st.number_of_values - st.true_count;
            const bool v = p.i != 0;
            if (op == ORCB_OP_EQ) return v ? st.true_count > 0 : false_count > 0;
            if (op == ORCB_OP_NE) return v ? false_count > 0 : st.true_count > 0;
            return true;
        }
        default: return true;  // Binary, Collection
    }
}

// row_group_might_match_bloom + bloom_value_hash64 (:362-406)
bool bloom_match(const RowGroupEntry& e, int op, const Predicate& p) {
    if (op != ORCB_OP_EQ || !e.has_bloom || p.vnull) return true;
    uint64_t h;
    if (p.vtype == ORCB_VAL_UTF8) h = bloom_hash_bytes((const uint8_t*)p.s.data(), p.s.size());
    else if (p.vtype == ORCB_VAL_FLOAT32 || p.vtype == ORCB_VAL_FLOAT64) {
        int64_t bits;
        memcpy(&bits, &p.f, 8);
        h = bloom_hash_long(bits);
    } else if (p.vtype == ORCB_VAL_BOOLEAN) h = bloom_hash_long(p.i ? 1 : 0);
    else h = bloom_hash_long(p.i);
    return bloom_test_hash(e.bloom, h);
}

struct Evaluator {
    const StripeIndex& idx;
    const std::vector<OutColumn>& cols;

    const std::vector<RowGroupEntry>& column_index(const std::string& name) const {  // find_column_index + row_index.column
        for (auto& c : cols) {
            if (c.name != name) continue;
            auto it = idx.columns.find(c.col_id);
            if (it == idx.columns.end()) fail(ORCB_UNEXPECTED, "Row index not found for column '" + name + "'");
            return it->second;
        }
        fail(ORCB_UNEXPECTED, "Column '" + name + "' not found in schema");
    }

    void comparison(const Predicate& p, int op, std::vector<uint8_t>& result) const {  // :128-168
        const auto& entries = column_index(p.column);
        for (size_t g = 0; g < result.size() && g < entries.size(); g++) {
            const RowGroupEntry& e = entries[g];
            if (e.has_stats && !stats_match(e.stats, op, p)) {
                result[g] = 0;
                continue;
            }
            result[g] = bloom_match(e, op, p);
        }
    }
    void null_test(const Predicate& p, bool want_null, std::vector<uint8_t>& result) const {  // :532-574
        const auto& entries = column_index(p.column);
        for (size_t g = 0; g < result.size() && g < entries.size(); g++) {
            const RowGroupEntry& e = entries[g];
            result[g] = !e.has_stats ? 1 : want_null ? e.stats.has_null : e.stats.number_of_values > 0;
        }
    }
    static int negate(int op) {  // ComparisonOp::negate (src/predicate.rs:66-77)
        switch (op) {
            case ORCB_OP_EQ: return ORCB_OP_NE;
            case ORCB_OP_NE: return ORCB_OP_EQ;
            case ORCB_OP_LT: return ORCB_OP_GE;
            case ORCB_OP_LE: return ORCB_OP_GT;
            case ORCB_OP_GT: return ORCB_OP_LE;
            default: return ORCB_OP_LT;
        }
    }
    // evaluate_predicate_recursive (:38-114); `negated` folds its De Morgan rewriting of NOT
    void eval(const Predicate& p, bool negated, std::vector<uint8_t>& result) const {
        switch (p.kind) {
            case ORCB_PRED_COMPARISON: comparison(p, negated ? negate(p.op) : p.op, result); break;
            case ORCB_PRED_IS_NULL: null_test(p, !negated, result); break;
            case ORCB_PRED_IS_NOT_NULL: null_test(p, negated, result); break;
            case ORCB_PRED_NOT: eval(p.children[0], !negated, result); break;
            default: {
                const bool conj = (p.kind == ORCB_PRED_AND) != negated;  // NOT(AND) = OR(NOT..), NOT(OR) = AND(NOT..)
                std::vector<std::vector<uint8_t>> temps;
                for (auto& c : p.children) {
                    temps.emplace_back(result.size(), 1);
                    eval(c, negated, temps.back());
                }
                for (size_t g = 0; g < result.size(); g++) {
                    if (conj) {
                        for (auto& t : temps) result[g] = result[g] && t[g];
                    } else {
                        bool any = false;
                        for (auto& t : temps) any = any || t[g];
                        result[g] = any;
                    }
                }
            }
        }
    }
};

}  // namespace

std::vector<RowSelector> predicate_selection(const FileMeta& fm, uint32_t stripe, const std::vector<OutColumn>& cols,
                                             const Predicate& pred, std::vector<uint8_t>* filter, bool* evaluated) {
    const StripeInfo& si = fm.stripes[stripe];
    const uint64_t rows_per_group = fm.row_index_stride >= 0 ? (uint64_t)fm.row_index_stride : 10000;  // src/stripe.rs:300
    // (the decoder takes stripes of up to 2^32 rows; without this a damaged row count sizes the verdict vector below)
    if (si.rows > 0xfffffff0ull) fail(ORCB_NOT_IMPLEMENTED, "stripes with more than 2^32 rows");
    std::vector<uint8_t> result;
    bool ok = true;
    // callback files: the index area and the stripe footer, not the data in between (a pruned stripe is never read)
    std::shared_ptr<RangeBuf> held_index, held_footer;
    try {
        if (fm.source) {
            held_index = fm.load_range(si.offset, si.index_length);
            held_footer = fm.load_range(si.offset + si.index_length + si.data_length, si.footer_length);
        }
        StripeIndex idx;
        const StripeFooter sf = fm.read_stripe_footer(stripe);
        for (auto& c : cols) {
            bool present = false;
            auto entries = fm.read_row_group_entries(sf, c.col_id, &present);
            if (present) idx.columns[c.col_id] = std::move(entries);
        }
        idx.num_row_groups = rows_per_group ? (size_t)((si.rows + rows_per_group - 1) / rows_per_group) : 0;
        result.assign(idx.num_row_groups, 1);
        Evaluator{idx, cols}.eval(pred, false, result);
    } catch (const OrcException&) {
        ok = false;  // "Keep all rows (maybe)", src/arrow_reader.rs:281-291
    }
    if (evaluated) *evaluated = ok;
    std::vector<RowSelector> sel;
    if (!ok) {
        if (filter) filter->clear();
        sel.push_back({si.rows, false});  // RowSelection::select_all
        return sel;
    }
    if (filter) *filter = result;
    // RowSelection::from_row_group_filter (src/row_selection.rs:348-390): every kept group selects a full stride, also
    // the last, shorter one; rows past the groups are skipped
    if (result.empty()) {
        sel.push_back({si.rows, true});
        return sel;
    }
    for (uint8_t keep : result) {
        if (!sel.empty() && sel.back().skip == !keep) sel.back().row_count += rows_per_group;
        else sel.push_back({rows_per_group, !keep});
    }
    const uint64_t covered = (uint64_t)result.size() * rows_per_group;
    if (covered < si.rows) {
        if (sel.back().skip) sel.back().row_count += si.rows - covered;
        else sel.push_back({si.rows - covered, true});
    }
    return sel;
}

}  // namespace orcb
