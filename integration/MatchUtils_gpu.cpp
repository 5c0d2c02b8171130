// MatchUtils_gpu.cpp -- drop-in replacement for the three putative matchers of the reference's
// VisionLocalizeCommon/src/MatchUtils.cpp (lines 73-152, 156-277, 283-367).  Compile this file
// INSTEAD of those three function bodies (and, optionally, of hulo::geometricMatch: see the end) and
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

// ---------------------------------------------------------------------------------------------
// hulo::geometricMatch (MatchUtils.cpp:372-420, decl MatchUtils.h:66-72) on the GPU: compile this
// INSTEAD of the original body to move the F-matrix AC-RANSAC (and, with bGuided_matching, the
// guided matching) of every pair into one launch each (hulo_geometric_filter / hulo_guided_match).
// The Regions_Provider gives feature positions (through the camera's undistortion, like
// MatchesPairToMat does); the descriptors for guided matching are read from sMatchesDir, which the
// callers have at hand (LocalizeEngine::mMatchDir, the CLIs' sMatchesDir) -- pass it through a
// file-scope setter if the signature must stay exactly as it is.
namespace hulo { std::string g_geometricMatchesDir; }     // set by the caller before guided matching

void hulo::geometricMatch(openMVG::sfm::SfM_Data &sfm_dataFull,
                          std::shared_ptr<openMVG::sfm::Regions_Provider> regions_provider,
                          openMVG::matching::PairWiseMatches &map_putativeMatches,
                          openMVG::matching::PairWiseMatches &map_geometricMatches, int ransacRound, double geomPrec,
                          bool bGuided_matching) {
    hulo_plain::Views views;
    hulo_plain::RegionsProvider regions;
    auto add_view = [&](openMVG::IndexT id) {
        if (views.count(id)) return;
        const auto &v = sfm_dataFull.views.at(id);
        hulo_plain::View pv{v->id_view, v->s_Img_path};
        pv.ui_width = v->ui_width;
        pv.ui_height = v->ui_height;
        views[id] = pv;
        const auto it = sfm_dataFull.GetIntrinsics().find(v->id_intrinsic);
        const openMVG::cameras::IntrinsicBase *cam = it == sfm_dataFull.GetIntrinsics().end() ? nullptr : it->second.get();
        const auto &reg = *regions_provider->regions_per_view.at(id);
        hulo_plain::FeatureLocations &f = regions[id];
        for (size_t i = 0; i < reg.RegionCount(); ++i) {
            const openMVG::Vec2 x = reg.GetRegionPosition(i);
            const openMVG::Vec2 u = (cam && cam->have_disto()) ? cam->get_ud_pixel(x) : x;
            f.push_back(std::make_pair(u(0), u(1)));
        }
    };
    hulo_plain::PairWiseMatches put, geo;
    for (const auto &kv : map_putativeMatches) {
        add_view(kv.first.first);
        add_view(kv.first.second);
        hulo_plain::IndMatches &dst = put[hulo_plain::Pair(kv.first.first, kv.first.second)];
        for (const auto &m : kv.second) dst.push_back(hulo_plain::IndMatch(m.i_, m.j_));
    }
    hulo_plain::geometricMatch(hulo_plain::defaultSession(), views, regions, hulo::g_geometricMatchesDir, put, geo,
                               ransacRound, geomPrec, bGuided_matching);
    map_geometricMatches.clear();
    toOpenMVG(geo, map_geometricMatches);
}
