// resect_views.cpp -- see resect_views.h
#include "resect_views.h"

#include <chrono>
#include <map>

namespace hulo {

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int resectViews(hulo_gpu *h, SfMScene &scene, std::size_t max_iter, uint64_t seed, ResectViewsReport *report) {
    ResectViewsReport rep;
    rep.views = scene.views.size();
    const double t0 = now_ms();
    // 2D-3D pairs of every view in ascending landmark id: one pass over the structure (the reference
    // scans the whole structure once per view, adjust_sfm_data.cpp:100-107)
    std::map<std::size_t, std::size_t> slot_of_view;
    std::vector<const View *> view_at;
    for (const auto &kv : scene.views) {
        slot_of_view[kv.second.id_view] = view_at.size();
        view_at.push_back(&kv.second);
    }
    std::vector<std::size_t> count(view_at.size(), 0);
    for (const Landmark &lm : scene.landmarks)
        for (const Observation &o : lm.obs) {
            const auto it = slot_of_view.find(o.id_view);
            if (it != slot_of_view.end()) ++count[it->second];
        }
    // problems: views with more than MINIMUM_VIEW_NUM_TO_ESTIMATAE_CAMERA_POSE (10) pairs and a pinhole intrinsic
    std::vector<long> problem_of(view_at.size(), -1);
    std::vector<uint64_t> offsets(1, 0);
    std::vector<double> K;
    for (std::size_t s = 0; s < view_at.size(); ++s) {
        if (count[s] <= 10) { rep.too_few_warning = true; continue; }
        const auto in = scene.intrinsics.find(view_at[s]->id_intrinsic);
        if (in == scene.intrinsics.end()) continue;
        problem_of[s] = (long)rep.view_ids.size();
        rep.view_ids.push_back(view_at[s]->id_view);
        rep.n_points.push_back(count[s]);
        offsets.push_back(offsets.back() + count[s]);
        double k9[9];
        in->second.K(k9);
        K.insert(K.end(), k9, k9 + 9);
    }
    const std::size_t n = rep.view_ids.size();
    rep.attempted = n;
    std::vector<double> x2d(2 * offsets.back()), X3d(3 * offsets.back());
    std::vector<uint64_t> fill(offsets.begin(), offsets.end() - 1);
    for (const Landmark &lm : scene.landmarks)
        for (const Observation &o : lm.obs) {
            const auto it = slot_of_view.find(o.id_view);
            if (it == slot_of_view.end() || problem_of[it->second] < 0) continue;
            const uint64_t at = fill[(std::size_t)problem_of[it->second]]++;
            x2d[2 * at] = o.x[0];
            x2d[2 * at + 1] = o.x[1];
            for (int c = 0; c < 3; ++c) X3d[3 * at + c] = lm.X[c];
        }
    const double t1 = now_ms();
    rep.ms_gather = t1 - t0;

    std::vector<double> P(12 * std::max<std::size_t>(n, 1)), emax(std::max<std::size_t>(n, 1));
    std::vector<int32_t> inliers(std::max<uint64_t>(offsets.back(), 1)), found(std::max<std::size_t>(n, 1));
    std::vector<uint64_t> n_inl(std::max<std::size_t>(n, 1));
    // the seed of a view hangs on its id, not on its position in the batch
    std::vector<uint64_t> seeds(n);
    for (std::size_t p = 0; p < n; ++p) seeds[p] = seed + 1000003ull * (uint64_t)rep.view_ids[p];
    if (n > 0) {
        const int rc = hulo_resect_acransac_batch(h, n, offsets.data(), x2d.data(), X3d.data(), K.data(), max_iter, seed,
                                                  seeds.data(), P.data(), inliers.data(), n_inl.data(), emax.data(),
                                                  found.data());
        if (rc != HULO_OK) return rc;
    }
    for (std::size_t s = 0; s < view_at.size(); ++s) {
        if (problem_of[s] < 0) continue;
        const std::size_t p = (std::size_t)problem_of[s];
        rep.n_inliers.push_back((std::size_t)n_inl[p]);
        rep.error_max.push_back(emax[p]);
        if (!found[p]) continue;
        Pose pose;
        hulo_pose_from_projection(P.data() + 12 * p, nullptr, pose.R, pose.center);   // adjust_sfm_data.cpp:138-142
        scene.poses[view_at[s]->id_pose] = pose;
        ++rep.resected;
    }
    rep.ms_resect = now_ms() - t1;
    if (report) *report = rep;
    return HULO_OK;
}

}  // namespace hulo
