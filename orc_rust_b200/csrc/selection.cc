// Row selection: the batches each stripe yields under ArrowReaderBuilder::with_row_selection.
#include "job_internal.h"

namespace orcb {

namespace {
uint64_t total(const std::vector<RowSelector>& v) {
    uint64_t n = 0;
    for (auto& x : v) n += x.row_count;
    return n;
}
// RowSelection::split_off: the first `n` rows leave `self`
std::vector<RowSelector> split_off(std::vector<RowSelector>& self, uint64_t n) {
    uint64_t acc = 0;
    size_t idx = self.size();
    for (size_t i = 0; i < self.size(); i++) {
        acc += self[i].row_count;
        if (acc > n) { idx = i; break; }
    }
    if (idx == self.size()) {
        std::vector<RowSelector> all;
        all.swap(self);
        return all;
    }
    std::vector<RowSelector> head(self.begin(), self.begin() + idx), rest(self.begin() + idx, self.end());
    const uint64_t overflow = acc - n;
    if (rest.front().row_count != overflow) head.push_back({rest.front().row_count - overflow, rest.front().skip});
    rest.front().row_count = overflow;
    self.swap(rest);
    return head;
}
// RowSelection::and_then (src/row_selection.rs:401-463): `second` picks among the rows `first` selects
std::vector<RowSelector> and_then(std::vector<RowSelector> first, std::vector<RowSelector> second) {
    std::vector<RowSelector> outsel;
    size_t a = 0, b = 0;
    uint64_t to_skip = 0;
    while (b < second.size()) {
        if (a >= first.size()) throw ReferencePanic("selection exceeds the number of selected rows");
        if (second[b].row_count == 0) { b++; continue; }
        if (first[a].row_count == 0) { a++; continue; }
        if (first[a].skip) {
            to_skip += first[a].row_count;
            a++;
            continue;
        }
        const uint64_t k = std::min(first[a].row_count, second[b].row_count);
        first[a].row_count -= k;
        second[b].row_count -= k;
        if (second[b].skip) {
            to_skip += k;
        } else {
            if (to_skip) outsel.push_back({to_skip, true});
            to_skip = 0;
            outsel.push_back({k, false});
        }
    }
    for (; a < first.size(); a++) {
        if (first[a].row_count == 0) continue;
        if (!first[a].skip) throw ReferencePanic("selection contains less than the number of selected rows");
        to_skip += first[a].row_count;
    }
    if (to_skip) outsel.push_back({to_skip, true});
    return outsel;
}
}  // namespace

SelectionCursor::SelectionCursor(std::vector<RowSelector> raw, bool has_selection) : has_selection_(has_selection) {
    // RowSelection::from(Vec<RowSelector>): empty selectors dropped, neighbours of the same kind merged
    for (auto& r : raw) {
        if (r.row_count == 0) continue;
        if (!sel_.empty() && sel_.back().skip == r.skip) sel_.back().row_count += r.row_count;
        else sel_.push_back(r);
    }
}

std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>> SelectionCursor::next_stripe(uint64_t rows, uint64_t batch_size,
                                                                                         const std::vector<RowSelector>* predicate) {
    std::vector<std::pair<uint32_t, uint32_t>> views;
    // ArrowReader::try_advance_stripe (src/arrow_reader.rs:256-309): the predicate's selection for this stripe,
    // then the caller's while it still has rows (a used-up selection no longer restricts anything)
    bool applies = predicate != nullptr;
    std::vector<RowSelector> s;
    if (predicate) s = *predicate;
    if (has_selection_ && total(sel_) > 0) {
        std::vector<RowSelector> mine = split_off(sel_, rows);
        s = predicate ? and_then(std::move(mine), std::move(s)) : std::move(mine);
        applies = true;
    }
    if (!applies) return {false, views};
    // NaiveStripeDecoder::next / next_with_row_selection.  A selector is left behind only once a single step has
    // covered its whole row_count, so a select longer than the batch size keeps yielding batches (kept as is)
    uint64_t index = 0;
    size_t si = 0;
    while (index < rows && si < s.size()) {
        const uint64_t remaining = rows - index;
        if (s[si].skip) {
            const uint64_t k = std::min(s[si].row_count, remaining);
            if (k == 0) { si++; continue; }
            index += k;
            if (k >= s[si].row_count) si++;
        } else {
            const uint64_t k = std::min(std::min(s[si].row_count, batch_size), remaining);
            if (k == 0) { si++; continue; }
            views.emplace_back((uint32_t)index, (uint32_t)k);
            index += k;
            if (k >= s[si].row_count) si++;
        }
    }
    return {true, views};
}

std::vector<std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>>> selection_views(
    std::vector<RowSelector> raw, const std::vector<uint64_t>& stripe_rows, uint64_t batch_size,
    const std::vector<std::vector<RowSelector>>* predicate, bool has_selection) {
    SelectionCursor cur(std::move(raw), has_selection);
    std::vector<std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>>> out;
    for (size_t sidx = 0; sidx < stripe_rows.size(); sidx++)
        out.push_back(cur.next_stripe(stripe_rows[sidx], batch_size, predicate ? &(*predicate)[sidx] : nullptr));
    return out;
}

}  // namespace orcb
