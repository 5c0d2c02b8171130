// pair_lists.h -- pair-list generators of VisionLocalizeCommon/src/SfMDataUtils.cpp:128-207 and
// the multi-GPU partition of a pair list (independent units, no collective).
#pragma once
#include <vector>

#include "hulo_types.h"

namespace hulo {

// (i, j), i < j, over the views in map order, pairs of view ids  (SfMDataUtils.cpp:128-141)
void generateAllPairs(const Views &views, std::vector<Pair> &pairs);
// (i, j), i < j <= i + frame  (SfMDataUtils.cpp:144-157)
void generateVideoMatchPairs(const Views &views, std::vector<Pair> &pairs, int frame);
void orderPair(Pair &p);
// drops later duplicates of an unordered pair, keeps first occurrences in order (:168-190)
void removeDupPairs(std::vector<Pair> &pairs);

// Shard a pair list over `world` ranks balanced by the work n_I * n_J of each pair (longest
// processing time first onto the least loaded rank).  rows_of_view gives the descriptor count
// of a view id.  Returns, for `rank`, the positions (ascending) of its pairs in `pairs`; the
// union over ranks is every position exactly once, so concatenating the ranks' results in
// position order restores the single-GPU output.
std::vector<std::size_t> partitionPairs(const std::vector<Pair> &pairs,
                                        const std::map<std::size_t, std::size_t> &rows_of_view, int rank,
                                        int world);

}  // namespace hulo
