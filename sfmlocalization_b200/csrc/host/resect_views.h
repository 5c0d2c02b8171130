// resect_views.h -- the first stage of the reference's OpenMVG_BA tool
// (OpenMVG_BA/src/adjust_sfm_data.cpp:91-146): every view of a reconstruction is re-resected
// against the structure it observes -- SfM_Localizer::Localize per view inside an omp loop there,
// ONE hulo_resect_acransac_batch call here -- and its pose overwritten.  The Ceres bundle
// adjustment that follows in that tool (:182-244) is outside the accelerated path.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "../../../include/hulo_gpu.h"
#include "sfm_data_io.h"

namespace hulo {

struct ResectViewsReport {
    std::size_t views = 0;            // views of the scene
    std::size_t attempted = 0;        // with more than 10 observations (adjust_sfm_data.cpp:39, :118)
    std::size_t resected = 0;         // SfM_Localizer::Localize returned true; pose overwritten
    bool too_few_warning = false;     // "there is/are frames with too few matches" (:148-150)
    std::vector<std::size_t> view_ids, n_points, n_inliers;   // per attempted view
    std::vector<double> error_max;                             // px
    double ms_gather = 0, ms_resect = 0;
};

// Re-resects the views of `scene` in place (scene.poses[view.id_pose]).  max_iter: the reference
// runs Image_Localizer_Match_Data's default 4096.  Observations are used as stored -- the
// reference fills resection_data.pt2D with the raw ob.x (adjust_sfm_data.cpp:109-115; the
// undistorted copy made at :124-130 is never used).  A view whose resection fails keeps its pose
// (the reference would decompose an unset projection matrix there).
int resectViews(hulo_gpu *h, SfMScene &scene, std::size_t max_iter, uint64_t seed, ResectViewsReport *report);

}  // namespace hulo
