// ArrowReaderBuilder::with_predicate: row groups pruned by their statistics and Bloom filters
// (src/predicate.rs, src/row_group_filter.rs, src/bloom_filter.rs).  Host-only; the result is a row selection that
// goes down the same partial-decode path as with_row_selection.
#pragma once
#include "job.h"

namespace orcb {

// Predicate / PredicateValue / ComparisonOp (src/predicate.rs:28-104); the numeric codes are those of
// OrcbPredicateNode in include/orc_b200.h
struct Predicate {
    int kind = 0;        // ORCB_PRED_*
    int op = 0;          // ORCB_OP_*
    int vtype = 0;       // ORCB_VAL_*
    bool vnull = false;  // PredicateValue::X(None)
    int64_t i = 0;
    double f = 0;
    std::string column, s;
    std::vector<Predicate> children;
};

Predicate predicate_from_c(const OrcbPredicateNode* nodes, uint32_t n_nodes);

// What ArrowReader::try_advance_stripe (src/arrow_reader.rs:256-293) derives for one stripe: the row-group filter
// turned into a selection, or select_all(stripe rows) when the index cannot be read or the predicate cannot be
// evaluated.  `filter` (optional) receives the row-group verdicts, *evaluated says which of the two happened.
std::vector<RowSelector> predicate_selection(const FileMeta& fm, uint32_t stripe, const std::vector<OutColumn>& cols,
                                             const Predicate& pred, std::vector<uint8_t>* filter, bool* evaluated);

// BloomFilter::hash_long / hash_bytes / test_hash (src/bloom_filter.rs:96-167, 182-230)
uint64_t bloom_hash_long(int64_t v);
uint64_t bloom_hash_bytes(const uint8_t* p, size_t n);
bool bloom_test_hash(const BloomBits& b, uint64_t hash64);

}  // namespace orcb
