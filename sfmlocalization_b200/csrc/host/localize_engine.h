// localize_engine.h -- the reference's LocalizeEngine class
// (VisionLocalizeServer/src/LocalizeEngine.h:46-82) on top of the C-ABI of libhulo_gpu.so.
//
// Same constructor arguments and the same localize() result ([t_out(3), R row-major(9)] or an
// empty vector) as the reference.  Differences, all forced by scope:
//   * localize() takes the query image's extracted AKAZE regions (descriptor rows + feature
//     positions) and its size instead of a cv::Mat: extraction (LocalizeEngine.cc:200-260,
//     334-352) is upstream of the accelerated path and stays with the caller;
//   * BoW view pre-selection (bowKnnNum, LocalizeEngine.cc:334-362) works from the views' .bow
//     files and the query's bag-of-features vector handed to localize(); computing that vector
//     from the image (dense features + vocabulary) stays with the caller.  iBeacon pre-selection
//     (beaconKnnNum, beaconStr) is out of scope: accepted, a non-zero value prints a note.
// guidedMatching = true re-matches all features of every pair that passed the F-matrix filter
// behind the epipolar gate (hulo_engine_set_guided_matching), like bGuided_matching of
// hulo::geometricMatch (LocalizeEngine.cc:458).
// The object is a copyable handle (shared state), because the reference stores engines by
// value in a std::map (localizeImage.cc:100).  Not re-entrant, like the reference (:71-74).
#pragma once
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "hulo_types.h"
#include "sfm_data_io.h"

struct hulo_gpu;
struct hulo_engine;

namespace hulo {

class LocalizeEngine {
public:
    LocalizeEngine();
    LocalizeEngine(const std::string sfmDataDir, const std::string matchDir, const std::string AmatFile,
                   double secondTestRatio, int ransacRound, double ransacPrecision, bool guidedMatching,
                   int beaconKnnNum = 0, int bowKnnNum = 0, int device = 0);

    // desc: n rows of `stride` bytes (61..64), qFeatLoc: n feature positions (distorted pixels,
    // as extractAKAZESingleImg returns them), imageWidth/Height: size of the query image.
    // points2D (N x 2), points3D (N x 3) and pointsInlier are filled when bReturnKeypoints;
    // times gets the reference's six slots (beacon, bow, extract, putative, geometric, PnP) in
    // seconds when bReturnTime (the first three are 0 here).
    std::vector<double> localize(const uint8_t *desc, std::size_t n, std::size_t stride,
                                 const FeatureLocations &qFeatLoc, std::size_t imageWidth, std::size_t imageHeight,
                                 const std::string &beaconStr, bool bReturnKeypoints, std::vector<double> &points2D,
                                 std::vector<double> &points3D, std::vector<int> &pointsInlier, bool bReturnTime,
                                 std::vector<double> &times, const std::vector<double> &center = std::vector<double>(),
                                 double radius = -1.0, uint64_t seed = 1, const std::vector<float> *queryBow = nullptr);

    // what the CLI prints into <outDir>/<basename>.json (localization.cpp:100-144)
    struct LastResult {
        bool localized = false;
        double K[9], R[9], t_out[3];
        std::vector<std::pair<std::size_t, std::size_t>> inlier_pairs;   // (query feature, landmark id)
    };
    const LastResult &last() const;
    const SfMScene &scene() const;
    // views restricted to a sphere around a location (hulo::getLocalViews, SfMDataUtils.cpp:210-227,
    // used by the CLI with -x -y -z -d): squared distance compared with `radius`, as the reference does
    void setLocalViews(const std::vector<double> &center, double radius);

    struct State;
private:
    std::shared_ptr<State> st_;
};

}  // namespace hulo
