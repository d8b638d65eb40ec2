// Row selection: the batches each stripe yields under ArrowReaderBuilder::with_row_selection.
#include "job_internal.h"

namespace orcb {

std::vector<std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>>> selection_views(
    std::vector<RowSelector> raw, const std::vector<uint64_t>& stripe_rows, uint64_t batch_size) {
    // RowSelection::from(Vec<RowSelector>): empty selectors dropped, neighbours of the same kind merged
    std::vector<RowSelector> sel;
    for (auto& r : raw) {
        if (r.row_count == 0) continue;
        if (!sel.empty() && sel.back().skip == r.skip) sel.back().row_count += r.row_count;
        else sel.push_back(r);
    }
    auto total = [](const std::vector<RowSelector>& v) {
        uint64_t n = 0;
        for (auto& x : v) n += x.row_count;
        return n;
    };
    // RowSelection::split_off: the first `n` rows leave `self`
    auto split_off = [](std::vector<RowSelector>& self, uint64_t n) {
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
    };
    std::vector<std::pair<bool, std::vector<std::pair<uint32_t, uint32_t>>>> out;
    for (uint64_t rows : stripe_rows) {
        std::vector<std::pair<uint32_t, uint32_t>> views;
        if (total(sel) == 0) {  // arrow_reader.rs:298: a used-up selection no longer restricts anything
            out.emplace_back(false, views);
            continue;
        }
        const std::vector<RowSelector> s = split_off(sel, rows);
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
        out.emplace_back(true, views);
    }
    return out;
}

}  // namespace orcb
