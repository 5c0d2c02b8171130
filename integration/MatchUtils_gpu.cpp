// MatchUtils_gpu.cpp -- drop-in replacement for the three putative matchers of the reference's
// VisionLocalizeCommon/src/MatchUtils.cpp (lines 73-152, 156-277, 283-367).  Compile this file
// INSTEAD of those three function bodies (keep hulo::geometricMatch from the original file) and
// link libhulo_host.so + libhulo_gpu.so.  Signatures are the reference's own
// (VisionLocalizeCommon/src/MatchUtils.h:39-61), so every caller -- LocalizeEngine.cc:423,
// localization.cpp:398, computeFeaturesAndMatches.cpp:156/187 -- is untouched.
//
// This file needs OpenMVG headers and therefore cannot be compiled in the development image;
// it only converts containers.  The logic it forwards to (csrc/host/match_utils_gpu.cpp) is
// built and tested there.
#include "MatchUtils.h"                      // the reference's header, unchanged

#include "match_utils_gpu.h"                 // sfmlocalization_b200/csrc/host

namespace {

// SfM_Data -> the view list the matchers read (view id + image path)
hulo_plain::Views toViews(const openMVG::sfm::SfM_Data &sfm_data) {
    hulo_plain::Views v;
    for (const auto &kv : sfm_data.views)
        v[kv.first] = hulo_plain::View{kv.second->id_view, kv.second->s_Img_path};
    return v;
}

void toOpenMVG(const hulo_plain::PairWiseMatches &in, openMVG::matching::PairWiseMatches &out) {
    for (const auto &kv : in) {
        openMVG::matching::IndMatches &dst = out[openMVG::Pair(kv.first.first, kv.first.second)];
        for (const hulo_plain::IndMatch &m : kv.second) dst.push_back(openMVG::matching::IndMatch(m.i_, m.j_));
    }
}

}  // namespace

// Build note: compile csrc/host with -Dhulo=hulo_plain (or wrap its headers in a namespace
// alias) so the plain-container functions do not collide with the names defined here.

void hulo::matchAKAZE(const openMVG::sfm::SfM_Data &sfm_data, const std::string &sMatchesDir,
                      const std::vector<std::pair<size_t, size_t>> &pairs, const float fDistRatio,
                      openMVG::matching::PairWiseMatches &matches) {
    hulo_plain::PairWiseMatches m;
    hulo_plain::matchAKAZE(toViews(sfm_data), sMatchesDir, pairs, fDistRatio, m);
    toOpenMVG(m, matches);
}

void hulo::trackAKAZE(const openMVG::sfm::SfM_Data &sfm_data, const std::string &sMatchesDir,
                      const size_t maxFrameDist, const float fDistRatio,
                      openMVG::matching::PairWiseMatches &matches) {
    hulo_plain::PairWiseMatches m;
    hulo_plain::trackAKAZE(toViews(sfm_data), sMatchesDir, maxFrameDist, fDistRatio, m);
    toOpenMVG(m, matches);
}

void hulo::matchAKAZEToQuery(const openMVG::sfm::SfM_Data &sfm_data, const std::string &sMatchesDir,
                             const std::string &sQueryMatchesDir, const std::vector<size_t> &pairs,
                             const size_t queryInd, const float fDistRatio,
                             openMVG::matching::PairWiseMatches &matches,
                             std::map<std::pair<size_t, size_t>, std::map<size_t, int>> &featDist) {
    hulo_plain::PairWiseMatches m;
    hulo_plain::FeatDistMap fd;
    hulo_plain::matchAKAZEToQuery(toViews(sfm_data), sMatchesDir, sQueryMatchesDir, pairs, queryInd, fDistRatio, m, fd);
    toOpenMVG(m, matches);
    for (const auto &kv : fd) featDist[kv.first] = kv.second;
}
