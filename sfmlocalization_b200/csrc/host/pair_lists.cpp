#include "pair_lists.h"

#include <algorithm>
#include <numeric>

namespace hulo {

void generateAllPairs(const Views &views, std::vector<Pair> &pairs) {
    std::vector<std::size_t> ids;
    for (const auto &kv : views) ids.push_back(kv.second.id_view);
    for (std::size_t a = 0; a < ids.size(); ++a)
        for (std::size_t b = a + 1; b < ids.size(); ++b) pairs.push_back(Pair(ids[a], ids[b]));
}

void generateVideoMatchPairs(const Views &views, std::vector<Pair> &pairs, int frame) {
    std::vector<std::size_t> ids;
    for (const auto &kv : views) ids.push_back(kv.second.id_view);
    for (std::size_t a = 0; a < ids.size(); ++a)
        for (std::size_t b = a + 1; b < ids.size() && b <= a + (std::size_t)std::max(frame, 0); ++b)
            pairs.push_back(Pair(ids[a], ids[b]));
}

void orderPair(Pair &p) {
    if (p.first > p.second) std::swap(p.first, p.second);
}

void removeDupPairs(std::vector<Pair> &pairs) {
    // The reference scans from the back, normalises earlier entries in place (orderPair on
    // pairs[j]) and erases an entry when an earlier one names the same unordered pair.  Net
    // effect: first occurrences survive in order; every survivor that was compared against a
    // later entry has been normalised to (min, max) -- i.e. all but the last element.
    std::vector<Pair> out;
    std::vector<Pair> seen;
    const std::size_t n = pairs.size();
    for (std::size_t k = 0; k < n; ++k) {
        Pair key = pairs[k];
        orderPair(key);
        if (std::find(seen.begin(), seen.end(), key) != seen.end()) continue;
        seen.push_back(key);
        Pair keep = pairs[k];
        if (k + 1 < n) orderPair(keep);
        out.push_back(keep);
    }
    pairs.swap(out);
}

std::vector<std::size_t> partitionPairs(const std::vector<Pair> &pairs,
                                        const std::map<std::size_t, std::size_t> &rows_of_view, int rank,
                                        int world) {
    std::vector<std::size_t> mine;
    if (world <= 1) {
        mine.resize(pairs.size());
        std::iota(mine.begin(), mine.end(), (std::size_t)0);
        return mine;
    }
    auto rows = [&](std::size_t v) {
        auto it = rows_of_view.find(v);
        return it == rows_of_view.end() ? (std::size_t)0 : it->second;
    };
    std::vector<std::size_t> order(pairs.size());
    std::iota(order.begin(), order.end(), (std::size_t)0);
    std::vector<unsigned long long> cost(pairs.size());
    for (std::size_t k = 0; k < pairs.size(); ++k)
        cost[k] = (unsigned long long)rows(pairs[k].first) * (unsigned long long)rows(pairs[k].second);
    std::stable_sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return cost[a] > cost[b]; });
    std::vector<unsigned long long> load((std::size_t)world, 0);
    for (std::size_t k : order) {
        const std::size_t r = (std::size_t)(std::min_element(load.begin(), load.end()) - load.begin());
        load[r] += cost[k] + 1;
        if ((int)r == rank) mine.push_back(k);
    }
    std::sort(mine.begin(), mine.end());
    return mine;
}

}  // namespace hulo
