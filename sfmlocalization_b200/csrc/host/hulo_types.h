// hulo_types.h -- plain C++ stand-ins for the OpenMVG containers that appear in the signatures
// of the reference's matching entry points (VisionLocalizeCommon/src/MatchUtils.h:39-72), so the
// host layer builds without OpenMVG.  Layout-compatible in spirit: PairWiseMatches is a
// std::map<Pair, IndMatches> exactly like openMVG::matching::PairWiseMatches, IndMatch has the
// i_ / j_ members of OpenMVG >= 1.0.  integration/MatchUtils_gpu.cpp shows the conversion from
// the real openMVG::sfm::SfM_Data.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace hulo {

typedef std::pair<std::size_t, std::size_t> Pair;

struct IndMatch {
    uint32_t i_, j_;
    IndMatch(uint32_t i = 0, uint32_t j = 0) : i_(i), j_(j) {}
    bool operator==(const IndMatch &o) const { return i_ == o.i_ && j_ == o.j_; }
};
typedef std::vector<IndMatch> IndMatches;
typedef std::map<Pair, IndMatches> PairWiseMatches;
// (viewID, viewID) -> query feature -> distance to its nearest neighbour (MatchUtils.h:60)
typedef std::map<Pair, std::map<std::size_t, int>> FeatDistMap;

// What the matchers read from openMVG::sfm::SfM_Data: views in ascending id with their image path
// and, for the geometric filter, the image size (View::ui_width / ui_height).
struct View {
    std::size_t id_view;
    std::string s_Img_path;
    std::size_t ui_width = 0, ui_height = 0;
    std::size_t id_intrinsic = 0, id_pose = 0;
};
typedef std::map<std::size_t, View> Views;

// What hulo::geometricMatch reads from the Regions_Provider: the (x, y) position of every
// feature of a view, in feature order (openMVG::features::PointFeature).
typedef std::vector<std::pair<double, double>> FeatureLocations;
typedef std::map<std::size_t, FeatureLocations> RegionsProvider;

}  // namespace hulo
