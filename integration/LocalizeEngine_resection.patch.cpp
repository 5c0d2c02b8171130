// LocalizeEngine_resection.patch.cpp -- how the resection call of the reference
//   VisionLocalizeServer/src/LocalizeEngine.cc:525-532
//   OpenMVGLocalization_AKAZE/src/localization.cpp:503-509
//   OpenMVG_BA/src/adjust_sfm_data.cpp:118-137
// is redirected to the GPU.  The block below replaces the single line
//     bResection = sfm::SfM_Localizer::Localize(make_pair(imageHeight, imageWidth), cam_I, resection_data, pose);
// and fills the same openMVG::sfm::Image_Localizer_Match_Data fields the callers read afterwards
// (projection_matrix, vec_inliers, error_max).  Needs OpenMVG/Eigen: not compiled in the
// development image.
#include <openMVG/sfm/pipelines/localization/SfM_Localizer.hpp>

#include "hulo_gpu.h"

// `gpu` is the hulo_gpu* the engine created in its constructor (LocalizeEngine.cc:84-198) and
// destroys in its destructor; hold it through a std::shared_ptr because LocalizeEngine objects
// are copied into a std::map (localizeImage.cc:100).
static bool LocalizeOnGpu(hulo_gpu *gpu, const openMVG::cameras::Pinhole_Intrinsic *cam_I,
                          openMVG::sfm::Image_Localizer_Match_Data &d, openMVG::geometry::Pose3 &pose,
                          uint64_t seed) {
    const size_t N = (size_t)d.pt2D.cols();
    // Eigen matrices are column-major: pt2D (2 x N) and pt3D (3 x N) are already the
    // "N points, coordinates contiguous" layout the C-ABI takes.
    const openMVG::Mat3 Kc = cam_I->K();
    const double K[9] = {Kc(0, 0), Kc(0, 1), Kc(0, 2), Kc(1, 0), Kc(1, 1), Kc(1, 2), Kc(2, 0), Kc(2, 1), Kc(2, 2)};
    double P[12], err_max = 0.0;
    std::vector<int32_t> inl(N ? N : 1);
    size_t n_inl = 0;
    int found = 0;
    if (hulo_resect_acransac(gpu, d.pt2D.data(), d.pt3D.data(), N, K, d.max_iteration, seed, P, inl.data(), &n_inl,
                             &err_max, &found) != HULO_OK)
        return false;
    d.vec_inliers.assign(inl.begin(), inl.begin() + n_inl);
    d.error_max = err_max;
    if (!found) return false;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c) d.projection_matrix(r, c) = P[4 * r + c];
    openMVG::Mat3 K_, R_;
    openMVG::Vec3 t_;
    openMVG::KRt_From_P(d.projection_matrix, &K_, &R_, &t_);
    pose = openMVG::geometry::Pose3(R_, -R_.transpose() * t_);
    return true;
}
