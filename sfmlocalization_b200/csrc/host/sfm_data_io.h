// sfm_data_io.h -- what the localisation engine reads from an OpenMVG 1.x cereal sfm_data.json
// (openMVG::sfm::Load(..., VIEWS | INTRINSICS | EXTRINSICS | STRUCTURE), called at
// VisionLocalizeServer/src/LocalizeEngine.cc:94-100 and OpenMVGLocalization_AKAZE/src/
// localization.cpp:236-244), and the global-coordinate matrix "A" of an OpenCV YAML file
// (LocalizeEngine.cc:113-119).
#pragma once
#include <map>
#include <string>
#include <utility>
#include <vector>

#include "hulo_types.h"

namespace hulo {

// openMVG::cameras::Pinhole_Intrinsic and its radial-distortion subclasses (k1 / k3).
struct Intrinsic {
    std::string type = "pinhole";      // cereal polymorphic_name: pinhole, pinhole_radial_k1, pinhole_radial_k3
    std::size_t width = 0, height = 0;
    double focal = 1.0, ppx = 0.0, ppy = 0.0;
    std::vector<double> disto;         // k1 [, k2, k3]
    void K(double out[9]) const;
    // IntrinsicBase::get_ud_pixel: undistorted pixel of a (distorted) image point
    std::pair<double, double> get_ud_pixel(double x, double y) const;
};

struct Pose {
    double R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};   // row-major rotation
    double center[3] = {0, 0, 0};
};

struct Observation { std::size_t id_view, id_feat; double x[2] = {0, 0}; };   // x: the observed image point
struct Landmark {
    std::size_t id = 0;
    double X[3] = {0, 0, 0};
    std::vector<Observation> obs;
};

struct SfMScene {
    std::string root_path;
    Views views;                                  // ascending id_view
    std::map<std::size_t, Intrinsic> intrinsics;
    std::map<std::size_t, Pose> poses;            // extrinsics, by id_pose
    std::vector<Landmark> landmarks;              // ascending id (the std::map order of Landmarks)
};

bool loadSfMData(const std::string &sfm_data_json, SfMScene &scene);

// openMVG::sfm::Save after the poses changed (adjust_sfm_data.cpp:152-155): `in_json` written to
// `out_json` with the entries of "extrinsics" replaced by `poses` (ascending id_pose); every other
// member, and every number of it, is written back as it was read.
bool saveSfMDataPoses(const std::string &in_json, const std::string &out_json, const std::map<std::size_t, Pose> &poses);

// cv::FileStorage(file, READ)[name] >> Mat for a numeric matrix: reads rows, cols and data of the
// "!!opencv-matrix" node `name`.  Returns false when the file or the node is missing.
bool readOpenCVMatrix(const std::string &yaml_file, const std::string &name, int &rows, int &cols,
                      std::vector<double> &data);

}  // namespace hulo
