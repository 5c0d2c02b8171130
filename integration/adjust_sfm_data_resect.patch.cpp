// adjust_sfm_data_resect.patch.cpp -- the view loop of the reference's OpenMVG_BA tool
//   OpenMVG_BA/src/adjust_sfm_data.cpp:88-150   (#pragma omp parallel for over the views, one
//   SfM_Localizer::Localize per view)
// as ONE batched call.  Replace the loop by ResectAllViewsOnGpu(gpu, sfm_data, seed); the save
// (:152-155) and the Ceres stage (:157-244) stay as they are.  Needs OpenMVG/Eigen: not compiled in
// the development image; sfmlocalization_b200/csrc/host/resect_views.cpp is the same loop over
// plain containers and is built and tested there (tests/test_ba_resect_cli_gpu.py).
#include <openMVG/sfm/sfm.hpp>

#include <vector>

#include "hulo_gpu.h"

#define MINIMUM_VIEW_NUM_TO_ESTIMATAE_CAMERA_POSE 10      // adjust_sfm_data.cpp:39

// Returns false when a view had too few observations (the reference prints a warning, :148-150).
static bool ResectAllViewsOnGpu(hulo_gpu *gpu, openMVG::sfm::SfM_Data &sfm_data, uint64_t seed) {
    using namespace openMVG;
    using namespace openMVG::sfm;
    // 2D-3D pairs of every view in one pass over the structure (the reference scans the whole
    // structure once per view, :100-107)
    std::map<IndexT, std::vector<double>> x2d_of, X3d_of;
    for (const auto &lm : sfm_data.structure)
        for (const auto &ob : lm.second.obs) {
            std::vector<double> &a = x2d_of[ob.first], &b = X3d_of[ob.first];
            a.push_back(ob.second.x(0)); a.push_back(ob.second.x(1));        // raw observation, as :109-115
            b.push_back(lm.second.X(0)); b.push_back(lm.second.X(1)); b.push_back(lm.second.X(2));
        }
    std::vector<const View *> views;
    std::vector<uint64_t> offsets(1, 0), seeds;
    std::vector<double> x2d, X3d, K;
    bool all_views_used = true;
    for (const auto &kv : sfm_data.views) {
        const View *v = kv.second.get();
        const std::vector<double> &a = x2d_of[v->id_view];
        if (a.size() / 2 <= MINIMUM_VIEW_NUM_TO_ESTIMATAE_CAMERA_POSE) { all_views_used = false; continue; }
        const auto it = sfm_data.GetIntrinsics().find(v->id_intrinsic);
        const cameras::Pinhole_Intrinsic *cam = dynamic_cast<const cameras::Pinhole_Intrinsic *>(it->second.get());
        const Mat3 Kc = cam->K();
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) K.push_back(Kc(r, c));
        x2d.insert(x2d.end(), a.begin(), a.end());
        const std::vector<double> &b = X3d_of[v->id_view];
        X3d.insert(X3d.end(), b.begin(), b.end());
        offsets.push_back(offsets.back() + a.size() / 2);
        seeds.push_back(seed + 1000003ull * v->id_view);
        views.push_back(v);
    }
    const size_t n = views.size();
    if (n == 0) return all_views_used;
    std::vector<double> P(12 * n), error_max(n);
    std::vector<int32_t> inliers(offsets.back()), found(n);
    std::vector<uint64_t> n_inliers(n);
    if (hulo_resect_acransac_batch(gpu, n, offsets.data(), x2d.data(), X3d.data(), K.data(), 4096 /* max_iteration */,
                                   seed, seeds.data(), P.data(), inliers.data(), n_inliers.data(), error_max.data(),
                                   found.data()) != HULO_OK)
        throw std::runtime_error(hulo_last_error());
    for (size_t k = 0; k < n; ++k) {
        if (!found[k]) continue;                       // the reference would decompose an unset matrix here
        double R[9], C[3];
        hulo_pose_from_projection(&P[12 * k], nullptr, R, C);                 // KRt_From_P, -R^T t  (:138-142)
        Mat3 R_;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) R_(r, c) = R[3 * r + c];
        sfm_data.poses[views[k]->id_pose] = geometry::Pose3(R_, Vec3(C[0], C[1], C[2]));
    }
    return all_views_used;
}
