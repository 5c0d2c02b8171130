// match_utils_gpu.h -- the reference's putative-matching entry points
// (VisionLocalizeCommon/src/MatchUtils.h:39-61) on top of the C-ABI of libhulo_gpu.so.
// Same names, argument order, argument meaning and result containers; SfM_Data is replaced by
// the view list the functions actually read from it (hulo_types.h).  Error behaviour follows
// the reference: the functions return void, unreadable descriptor files behave as empty
// images; a CUDA / library failure throws std::runtime_error (the reference has no equivalent
// because its matcher cannot fail).
#pragma once
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "hulo_types.h"

struct hulo_gpu;
struct hulo_db;

namespace hulo {

// Owns the device context and the descriptor tables kept resident between calls -- the state
// the reference lacks: it re-reads every view's .desc file on every query
// (MatchUtils.cpp:328-332).  Copyable handle (shared state), so it can live inside a
// LocalizeEngine that is stored by value (localizeImage.cc:100).
class GpuSession {
public:
    explicit GpuSession(int device = 0);
    hulo_gpu *gpu() const;
    // Descriptor table of the given views read from sMatchesDir (cached per directory + view set).
    struct Table {
        hulo_db *db = nullptr;
        std::vector<std::size_t> view_ids;          // segment s holds view view_ids[s]
        std::map<std::size_t, uint32_t> seg_of_view;
        std::map<std::size_t, std::size_t> rows_of_view;
    };
    std::shared_ptr<Table> table(const Views &views, const std::string &sMatchesDir,
                                 const std::vector<std::size_t> &view_ids);
    void clearCache();
    struct State;
    std::map<std::string, std::shared_ptr<struct BowTable>> &bowCache();
private:
    std::shared_ptr<State> st_;
};
GpuSession &defaultSession();

// hulo::matchAKAZE, MatchUtils.cpp:73-152
void matchAKAZE(const Views &views, const std::string &sMatchesDir, const std::vector<Pair> &pairs,
                const float fDistRatio, PairWiseMatches &matches);
// hulo::trackAKAZE, MatchUtils.cpp:156-277
void trackAKAZE(const Views &views, const std::string &sMatchesDir, const std::size_t maxFrameDist,
                const float fDistRatio, PairWiseMatches &matches);
// hulo::matchAKAZEToQuery, MatchUtils.cpp:283-367
void matchAKAZEToQuery(const Views &views, const std::string &sMatchesDir, const std::string &sQueryMatchesDir,
                       const std::vector<std::size_t> &pairs, const std::size_t queryInd, const float fDistRatio,
                       PairWiseMatches &matches, FeatDistMap &featDist);

// hulo::geometricMatch, MatchUtils.cpp:372-420 (decl MatchUtils.h:66-72): OpenMVG's
// GeometricFilter_FMatrix_AC(geomPrec, ransacRound) on every pair of map_putativeMatches; pairs
// whose robust estimation fails get no key in map_geometricMatches; surviving pairs hold the
// inliers in ACRANSAC's order.  SfM_Data is replaced by the views (image sizes) and the
// Regions_Provider by the feature positions (+ the directory of the .desc files, which the
// reference's Regions_Provider holds in memory).  With bGuided_matching the inliers of every
// surviving pair are replaced by OpenMVG's Geometry_guided_matching over ALL features of the two
// images (hulo_guided_match: epipolar gate at the robust precision, descriptor ratio 0.6^2), in
// ascending i; the overloads without sMatchesDir have no descriptors and throw
// std::invalid_argument when it is requested.
void geometricMatch(const Views &views, const RegionsProvider &regions_provider,
                    const PairWiseMatches &map_putativeMatches, PairWiseMatches &map_geometricMatches,
                    int ransacRound, double geomPrec, bool bGuided_matching);
void geometricMatch(GpuSession &s, const Views &views, const RegionsProvider &regions_provider,
                    const PairWiseMatches &map_putativeMatches, PairWiseMatches &map_geometricMatches,
                    int ransacRound, double geomPrec, bool bGuided_matching);
void geometricMatch(GpuSession &s, const Views &views, const RegionsProvider &regions_provider,
                    const std::string &sMatchesDir, const PairWiseMatches &map_putativeMatches,
                    PairWiseMatches &map_geometricMatches, int ransacRound, double geomPrec, bool bGuided_matching);
// sampler seed of the filter (pair (I, J) draws from a stream derived from it and from I, J, so
// the result of a pair does not depend on which other pairs are filtered with it)
extern uint64_t g_geometricSeed;

// hulo::selectViewByBoF, BoWCommon/src/BoFUtils.cpp:27-68: the knn views of viewList whose
// bag-of-features vector (<matchDir>/<basename>.bow, readMatBin) is nearest to the query's `bow`
// under L2.  The vectors of all views are read once per matchDir and kept on the device; the search
// is exact where the reference's per-query FLANN KD-tree (4 trees, 64 checks) is approximate.
// Throws std::invalid_argument unless knn < viewList.size() (CV_Assert at :30).
void selectViewByBoF(GpuSession &s, const std::vector<float> &bow, const std::string &matchDir,
                     const std::set<std::size_t> &viewList, const Views &views, int knn,
                     std::set<std::size_t> &selectedViewList);

// the same three against an explicit session (several GPUs, tests)
void matchAKAZE(GpuSession &s, const Views &views, const std::string &sMatchesDir, const std::vector<Pair> &pairs,
                const float fDistRatio, PairWiseMatches &matches);
void trackAKAZE(GpuSession &s, const Views &views, const std::string &sMatchesDir, const std::size_t maxFrameDist,
                const float fDistRatio, PairWiseMatches &matches);
void matchAKAZEToQuery(GpuSession &s, const Views &views, const std::string &sMatchesDir,
                       const std::string &sQueryMatchesDir, const std::vector<std::size_t> &pairs,
                       const std::size_t queryInd, const float fDistRatio, PairWiseMatches &matches,
                       FeatDistMap &featDist);

// Track propagation half of trackAKAZE (MatchUtils.cpp:239-276): extends consecutive-frame
// matches (f, f+1) to pairs (f, f+2 .. f+maxFrameDist-1).  n_frames = number of views,
// feat_number[f] = descriptor count of frame f (f < n_frames - 1).
void propagateTracks(std::size_t n_frames, std::size_t maxFrameDist, const std::vector<int> &feat_number,
                     PairWiseMatches &matches);

}  // namespace hulo
