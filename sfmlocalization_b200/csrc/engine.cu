// engine.cu -- the per-query pipeline of LocalizeEngine::localize around the two kernels
// (VisionLocalizeServer/src/LocalizeEngine.cc:423-602; CLI twin
// OpenMVGLocalization_AKAZE/src/localization.cpp:398-547): putative matching (K1), the
// "< 16 matches" view filter, the 2D-3D assembly of hulo::matchProviderToMatchSet
// (VisionLocalizeCommon/src/SfMDataUtils.cpp:59-125), AC-RANSAC resection (K2) and the pose
// extraction.  Host side C++; the map state lives on the device for the life of the engine.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "context.cuh"
#include "guided.cuh"

struct hulo_engine {
    hulo_gpu *h = nullptr;
    hulo_db *map = nullptr;
    size_t n_views = 0;
    // (view, feat) -> landmark as a dense table over the map rows: lm_of_row[seg[view] + feat], -1 = none
    std::vector<uint64_t> seg;
    std::vector<int32_t> lm_of_row;
    std::vector<double> X;     // n_landmarks x 3
    double K[9];
    float ratio = 0.6f;        // secondTestRatio, localizeImage.cc:46-59
    int min_putative = 16, min_points = 8, min_inliers = 10;
    size_t max_iter = 4096;
    int resect_schedule = HULO_RESECT_BATCHED;
    // geometric filter (hulo::geometricMatch, LocalizeEngine.cc:458): keypoint positions of the map
    // features (row order of the descriptor table), image size per view and of the query camera
    std::vector<double> map_xy;
    std::vector<int32_t> view_wh;
    int32_t query_wh[2] = {0, 0};
    bool geo_enabled = false;
    size_t geo_rounds = 25;        // mRansacRound (LocalizeParam.py:35)
    double geo_precision = 4.0;    // mRansacPrecision
    std::vector<double> g_xI, g_xJ;
    std::vector<uint64_t> g_off;
    std::vector<int32_t> g_sizes, g_valid, g_inl;
    std::vector<uint32_t> g_ninl;
    std::vector<size_t> g_view;    // position in views[] of each filtered pair
    // guided matching after the filter (bGuided_matching of hulo::geometricMatch, MatchUtils.cpp:407-416):
    // map keypoints resident on the device, position groups of the map views cached across queries
    bool guided = false;
    hulo::DevBuf d_map_xy, d_q_xy;
    hulo::GuidedGroups map_groups;
    std::vector<double> g_F, g_err, gm_F, gm_thr;
    std::vector<uint32_t> gm_pairs, gm_i, gm_j;
    std::vector<uint64_t> gm_off;
    std::vector<int64_t> gm_of_pair;
    // per-query scratch
    std::vector<uint32_t> m_view, m_i, m_j, view_counts;
    std::vector<int32_t> m_d0;
    std::vector<int32_t> fd, fd_stamp, best_d;
    std::vector<int64_t> best_lm;
    std::vector<double> x2d, X3d;
    std::vector<int32_t> inl;
};

namespace hulo {
namespace {

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

void mat3_mul(const double *A, const double *B, double *O) {
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
    memcpy(O, T, sizeof T);
}

// KRt_From_P (openMVG multiview/projection.hpp): RQ decomposition of the left 3x3 block with
// three Givens rotations, positive diagonal for K, det(R) = +1, K(2,2) = 1.
void krt_from_p(const double *P, double *K, double *R, double *t) {
    double Kk[9] = {P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10]};
    double Q[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    {   // zero (2,1)
        double c = -Kk[8], s = Kk[7];
        const double l = std::sqrt(c * c + s * s);
        c /= l; s /= l;
        const double G[9] = {1, 0, 0, 0, c, -s, 0, s, c};
        mat3_mul(Kk, G, Kk); mat3_mul(Q, G, Q);
    }
    {   // zero (2,0)
        double c = Kk[8], s = Kk[6];
        const double l = std::sqrt(c * c + s * s);
        c /= l; s /= l;
        const double G[9] = {c, 0, s, 0, 1, 0, -s, 0, c};
        mat3_mul(Kk, G, Kk); mat3_mul(Q, G, Q);
    }
    {   // zero (1,0)
        double c = -Kk[4], s = Kk[3];
        const double l = std::sqrt(c * c + s * s);
        c /= l; s /= l;
        const double G[9] = {c, -s, 0, s, c, 0, 0, 0, 1};
        mat3_mul(Kk, G, Kk); mat3_mul(Q, G, Q);
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = Q[3 * j + i];
    for (int a = 0; a < 3; ++a)
        if (Kk[4 * a] < 0) {
            for (int r = 0; r < 3; ++r) Kk[3 * r + a] = -Kk[3 * r + a];
            for (int c = 0; c < 3; ++c) R[3 * a + c] = -R[3 * a + c];
        }
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) +
                       R[2] * (R[3] * R[7] - R[4] * R[6]);
    double p4[3] = {P[3], P[7], P[11]};
    if (det < 0) {
        for (int k = 0; k < 9; ++k) R[k] = -R[k];
        for (int k = 0; k < 3; ++k) p4[k] = -p4[k];
    }
    // t = K^-1 p4, K upper triangular
    t[2] = p4[2] / Kk[8];
    t[1] = (p4[1] - Kk[5] * t[2]) / Kk[4];
    t[0] = (p4[0] - Kk[1] * t[1] - Kk[2] * t[2]) / Kk[0];
    const double sc = Kk[8];
    for (int k = 0; k < 9; ++k) K[k] = Kk[k] / sc;
}

}  // namespace
}  // namespace hulo

using namespace hulo;

extern "C" {

int hulo_pose_from_projection(const double *P, double *K, double *R, double *center) {
    if (P == nullptr || R == nullptr || center == nullptr) return HULO_ERR_ARG;
    double Kd[9], t[3];
    krt_from_p(P, Kd, R, t);
    // camera centre = -R^T t  (LocalizeEngine.cc:582-585, adjust_sfm_data.cpp:139-142)
    for (int c = 0; c < 3; ++c) center[c] = -(R[c] * t[0] + R[3 + c] * t[1] + R[6 + c] * t[2]);
    if (K != nullptr) memcpy(K, Kd, sizeof Kd);
    return HULO_OK;
}

int hulo_engine_create(hulo_gpu *h, const uint8_t *rows, size_t n, size_t stride, const uint64_t *seg_offsets,
                       size_t n_views, const uint32_t *obs_view, const uint32_t *obs_feat,
                       const uint32_t *obs_landmark, size_t n_obs, const double *landmark_X, size_t n_landmarks,
                       const double *K, hulo_engine **out) {
    HULO_ARG(h != nullptr && out != nullptr && K != nullptr, "null argument");
    *out = nullptr;
    HULO_ARG(seg_offsets != nullptr && n_views >= 1, "the map needs a view table");
    HULO_ARG(n_obs == 0 || (obs_view && obs_feat && obs_landmark), "null observation table");
    HULO_ARG(n_landmarks == 0 || landmark_X != nullptr, "null landmark positions");
    HULO_ARG(n_landmarks < (size_t)INT32_MAX, "more than 2^31 - 1 landmarks (ids are int32 on the device)");
    for (size_t k = 0; k < n_obs; ++k) {
        HULO_ARG(obs_view[k] < n_views, "observation refers to a view that does not exist");
        HULO_ARG(obs_landmark[k] < n_landmarks, "observation refers to a landmark that does not exist");
    }
    hulo_engine *e = new (std::nothrow) hulo_engine();
    HULO_ARG(e != nullptr, "out of host memory");
    e->h = h;
    int rc = hulo_db_upload(h, rows, n, stride, seg_offsets, n_views, &e->map);
    if (rc != HULO_OK) { delete e; return rc; }
    e->n_views = n_views;
    e->seg.assign(seg_offsets, seg_offsets + n_views + 1);
    e->lm_of_row.assign(std::max<size_t>(n, 1), -1);
    for (size_t k = 0; k < n_obs; ++k) {
        const uint64_t row = e->seg[obs_view[k]] + obs_feat[k];
        if (row >= e->seg[obs_view[k] + 1]) {
            set_error("hulo_engine_create: observation %zu refers to feature %u of view %u, which has %llu rows", k,
                      obs_feat[k], obs_view[k], (unsigned long long)(e->seg[obs_view[k] + 1] - e->seg[obs_view[k]]));
            hulo_db_free(e->map);
            delete e;
            return HULO_ERR_ARG;
        }
        // a (view, feature) observed by several landmarks keeps the last one written, like the
        // map[view][feat] = landmark assignment of SfMDataUtils.cpp:42
        e->lm_of_row[row] = (int32_t)obs_landmark[k];
    }
    e->X.assign(landmark_X, landmark_X + 3 * n_landmarks);
    memcpy(e->K, K, sizeof e->K);
    *out = e;
    return HULO_OK;
}

void hulo_engine_destroy(hulo_engine *e) {
    if (!e) return;
    hulo_db_free(e->map);
    e->d_map_xy.release();
    e->d_q_xy.release();
    delete e;
}

int hulo_engine_configure(hulo_engine *e, float ratio, int min_putative, int min_points, int min_inliers,
                          size_t max_iter) {
    HULO_ARG(e != nullptr, "null engine");
    HULO_ARG(ratio > 0.0f && min_putative >= 0 && min_points >= 3 && min_inliers >= 0 && max_iter >= 1, "bad threshold");
    e->ratio = ratio;
    e->min_putative = min_putative;
    e->min_points = min_points;
    e->min_inliers = min_inliers;
    e->max_iter = max_iter;
    return HULO_OK;
}

int hulo_engine_set_resection_schedule(hulo_engine *e, int schedule) {
    HULO_ARG(e != nullptr, "null engine");
    HULO_ARG(schedule == HULO_RESECT_BATCHED || schedule == HULO_RESECT_SEQUENTIAL, "unknown schedule");
    e->resect_schedule = schedule;
    return HULO_OK;
}

int hulo_engine_set_keypoints(hulo_engine *e, const double *map_xy, const int32_t *view_wh, int query_w, int query_h) {
    HULO_ARG(e != nullptr && map_xy != nullptr && view_wh != nullptr, "null argument");
    HULO_ARG(query_w > 0 && query_h > 0, "the query image size must be positive");
    for (size_t v = 0; v < 2 * e->n_views; ++v) HULO_ARG(view_wh[v] > 0, "view image sizes must be positive");
    const size_t n = (size_t)e->seg[e->n_views];
    e->map_xy.assign(map_xy, map_xy + 2 * n);
    e->view_wh.assign(view_wh, view_wh + 2 * e->n_views);
    e->query_wh[0] = query_w;
    e->query_wh[1] = query_h;
    // resident copy for guided matching; positions changed, so the cached position groups go
    HULO_CUDA(cudaSetDevice(e->h->device));
    HULO_CUDA(e->d_map_xy.reserve(std::max<size_t>(n, 1) * 2 * sizeof(double)));
    if (n) HULO_CUDA(cudaMemcpyAsync(e->d_map_xy.ptr, e->map_xy.data(), n * 2 * sizeof(double), cudaMemcpyHostToDevice, e->h->stream));
    HULO_CUDA(cudaStreamSynchronize(e->h->stream));
    e->map_groups.segs.clear();
    return HULO_OK;
}

int hulo_engine_set_guided_matching(hulo_engine *e, int enabled) {
    HULO_ARG(e != nullptr, "null engine");
    if (enabled) HULO_ARG(e->geo_enabled, "guided matching refines the geometric filter: enable hulo_engine_configure_geometric first");
    e->guided = enabled != 0;
    return HULO_OK;
}

int hulo_engine_set_query_size(hulo_engine *e, int query_w, int query_h) {
    HULO_ARG(e != nullptr, "null engine");
    HULO_ARG(query_w > 0 && query_h > 0, "the query image size must be positive");
    e->query_wh[0] = query_w;
    e->query_wh[1] = query_h;
    return HULO_OK;
}

int hulo_engine_configure_geometric(hulo_engine *e, int enabled, size_t ransac_round, double precision_px) {
    HULO_ARG(e != nullptr, "null engine");
    if (enabled) {
        HULO_ARG(!e->view_wh.empty(), "hulo_engine_set_keypoints must be called before enabling the geometric filter");
        HULO_ARG(ransac_round >= 1 && precision_px > 0.0, "bad geometric filter parameter");
        e->geo_rounds = ransac_round;
        e->geo_precision = precision_px;
    }
    e->geo_enabled = enabled != 0;
    if (!e->geo_enabled) e->guided = false;
    return HULO_OK;
}

// 2D-3D pairs of the queries of a batch, gathered for one batched resection (a query with too few
// pairs owns an empty range).
struct ResectionBatch {
    std::vector<double> x2d, X3d;
    std::vector<uint64_t> offsets = std::vector<uint64_t>(1, 0);
};

// Stages after the putative matching, shared by the single and the batched entry points:
// view filter, 2D-3D assembly, resection, pose.  m_* are the matches of this query grouped by
// view position (view_counts[v] entries each, emission order).
static int assemble_and_resect(hulo_engine *e, size_t nq, const double *qxy, const uint32_t *views, size_t n_views,
                               const uint32_t *m_i, const uint32_t *m_j, const int32_t *m_d0,
                               const uint32_t *view_counts, uint64_t seed, double *pose12, int *localized,
                               uint32_t *corr_qfeat, uint32_t *corr_landmark, size_t *n_corr, int32_t *inliers,
                               size_t *n_inliers, double *t_assembly, double *t_pnp, double *t_geo,
                               ResectionBatch *defer = nullptr, size_t q_row0 = 0) {
    double t1 = now_ms();
    // ---- 2D-3D assembly, hulo::matchProviderToMatchSet (SfMDataUtils.cpp:59-125).
    // The reference walks a std::map keyed by (view id, query id): ascending view id, and inside
    // a view the matches in emission order.  featDist[(v,q)][j] is the distance of the LAST
    // match of view v onto query feature j (MatchUtils.cpp:351); every candidate of that view for
    // j is weighed with it, and a candidate replaces the current one only if strictly closer.
    std::vector<size_t> start(n_views + 1, 0);
    for (size_t v = 0; v < n_views; ++v) start[v + 1] = start[v] + view_counts[v];
    std::vector<size_t> order(n_views);
    for (size_t v = 0; v < n_views; ++v) order[v] = v;
    if (views)
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return views[a] < views[b]; });
    e->fd.assign(nq, 0);
    e->fd_stamp.assign(nq, -1);
    e->best_lm.assign(nq, -1);
    e->best_d.assign(nq, 0);

    // ---- geometric filter, hulo::geometricMatch (LocalizeEngine.cc:458): every surviving
    // (view, query) pair goes through the F-matrix AC-RANSAC in one launch; the assembly below
    // then walks the geometric inliers of the pairs that stayed valid, in ACRANSAC's inlier order
    // (map_geometricMatches), while featDist keeps coming from the putative matches.
    std::vector<int64_t> geo_pair(n_views, -1);       // pair number of views[v], -1 = not filtered
    if (e->geo_enabled) {
        e->g_xI.clear(); e->g_xJ.clear(); e->g_sizes.clear(); e->g_view.clear();
        e->g_off.assign(1, 0);
        int32_t prev = -1;
        for (size_t oi = 0; oi < n_views; ++oi) {
            const size_t v = order[oi];
            const uint32_t view_id = views ? views[v] : (uint32_t)v;
            if ((int32_t)view_id == prev) continue;
            prev = (int32_t)view_id;
            if ((int)view_counts[v] < e->min_putative) continue;
            geo_pair[v] = (int64_t)e->g_view.size();
            e->g_view.push_back(v);
            for (size_t k = start[v]; k < start[v + 1]; ++k) {
                const size_t row = e->seg[view_id] + m_i[k];
                e->g_xI.push_back(e->map_xy[2 * row]);
                e->g_xI.push_back(e->map_xy[2 * row + 1]);
                e->g_xJ.push_back(qxy[2 * (size_t)m_j[k]]);
                e->g_xJ.push_back(qxy[2 * (size_t)m_j[k] + 1]);
            }
            e->g_off.push_back(e->g_off.back() + view_counts[v]);
            e->g_sizes.push_back(e->view_wh[2 * view_id]);
            e->g_sizes.push_back(e->view_wh[2 * view_id + 1]);
            e->g_sizes.push_back(e->query_wh[0]);
            e->g_sizes.push_back(e->query_wh[1]);
        }
        const size_t P = e->g_view.size();
        e->g_valid.assign(std::max<size_t>(P, 1), 0);
        e->g_ninl.assign(std::max<size_t>(P, 1), 0);
        e->g_inl.resize(std::max<size_t>((size_t)e->g_off.back(), 1));
        e->g_F.resize(9 * std::max<size_t>(P, 1));
        e->g_err.resize(std::max<size_t>(P, 1));
        if (P) {
            int rc = hulo_geometric_filter(e->h, e->g_xI.data(), e->g_xJ.data(), e->g_off.data(), P, e->g_sizes.data(),
                                           e->geo_precision, e->geo_rounds, seed + 77, nullptr, e->g_valid.data(),
                                           e->g_ninl.data(), e->g_inl.data(), e->g_F.data(), e->g_err.data(), nullptr);
            if (rc != HULO_OK) return rc;
        }
        if (e->guided) {
            // Geometry_guided_matching: ALL features of the view against ALL features of the query
            // behind the epipolar gate of the pair's F, descriptor ratio 0.6^2 (MatchUtils.cpp:410-414);
            // the result replaces the pair's inlier list.  I side: the resident map; J side: the
            // query rows as hulo_match_to_query(ies) left them (folded) in the staging buffer.
            e->gm_pairs.clear(); e->gm_F.clear(); e->gm_thr.clear();
            e->gm_of_pair.assign(std::max<size_t>(P, 1), -1);
            for (size_t p = 0; p < P; ++p) {
                if (!e->g_valid[p]) continue;
                const size_t v = e->g_view[p];
                e->gm_of_pair[p] = (int64_t)(e->gm_pairs.size() / 2);
                e->gm_pairs.push_back(views ? views[v] : (uint32_t)v);
                e->gm_pairs.push_back(0);
                e->gm_F.insert(e->gm_F.end(), e->g_F.begin() + 9 * p, e->g_F.begin() + 9 * p + 9);
                e->gm_thr.push_back(e->g_err[p] * e->g_err[p]);               // Square(m_dPrecision_robust)
            }
            const size_t G = e->gm_pairs.size() / 2;
            e->gm_off.assign(G + 1, 0);
            if (G) {
                HULO_CUDA(e->d_q_xy.reserve(std::max<size_t>(nq, 1) * 2 * sizeof(double)));
                HULO_CUDA(cudaMemcpyAsync(e->d_q_xy.ptr, qxy, nq * 2 * sizeof(double), cudaMemcpyHostToDevice, e->h->stream));
                const uint64_t q_seg[2] = {0, (uint64_t)nq};
                hulo::GuidedSide SI, SJ;
                SI.rows = e->map->rows; SI.seg = e->seg.data(); SI.n_seg = e->n_views;
                SI.h_xy = e->map_xy.data(); SI.d_xy = e->d_map_xy.as<double2>(); SI.groups = &e->map_groups;
                SJ.rows = e->h->stageB.as<uint4>() + 4 * q_row0; SJ.seg = q_seg; SJ.n_seg = 1;
                SJ.h_xy = qxy; SJ.d_xy = e->d_q_xy.as<double2>();
                size_t cap = std::max<size_t>(e->gm_i.size(), 4096), n_g = 0;
                for (;;) {
                    e->gm_i.resize(cap);
                    e->gm_j.resize(cap);
                    int rc = hulo::guided_match_sides(e->h, SI, SJ, e->gm_pairs.data(), G, e->gm_F.data(), e->gm_thr.data(),
                                                      0.6 * 0.6, 1, e->gm_off.data(), e->gm_i.data(), e->gm_j.data(), cap, &n_g);
                    if (rc == HULO_ERR_CAPACITY) { cap = n_g + n_g / 8; continue; }
                    if (rc != HULO_OK) return rc;
                    break;
                }
            }
        }
        const double tg = now_ms();
        if (t_geo) *t_geo += tg - t1;
        t1 = tg;
    }

    int32_t prev_view_id = -1;
    for (size_t oi = 0; oi < n_views; ++oi) {
        const size_t v = order[oi];
        const uint32_t view_id = views ? views[v] : (uint32_t)v;
        if ((int32_t)view_id == prev_view_id) continue;              // a view listed twice: one map key
        prev_view_id = (int32_t)view_id;
        // views with fewer putative matches than the threshold are dropped (LocalizeEngine.cc:428-434)
        if ((int)view_counts[v] < e->min_putative) continue;
        size_t n_cand = start[v + 1] - start[v];
        const int32_t *cand = nullptr;                               // nullptr: all putative matches, emission order
        if (e->geo_enabled) {
            const size_t p = (size_t)geo_pair[v];
            if (!e->g_valid[p]) continue;                            // pair not in map_geometricMatches
            cand = e->g_inl.data() + e->g_off[p];
            n_cand = e->g_ninl[p];
        }
        for (size_t k = start[v]; k < start[v + 1]; ++k) {
            const uint32_t j = m_j[k];
            e->fd[j] = m_d0[k];
            e->fd_stamp[j] = (int32_t)oi;
        }
        const uint32_t *g_i = nullptr, *g_j = nullptr;               // guided matches of the pair: (feature, query feature)
        if (e->geo_enabled && e->guided) {
            const size_t g = (size_t)e->gm_of_pair[(size_t)geo_pair[v]];
            g_i = e->gm_i.data() + e->gm_off[g];
            g_j = e->gm_j.data() + e->gm_off[g];
            n_cand = (size_t)(e->gm_off[g + 1] - e->gm_off[g]);
        }
        for (size_t c = 0; c < n_cand; ++c) {
            uint32_t fi, j;
            if (g_i) {
                fi = g_i[c]; j = g_j[c];
            } else {
                const size_t k = start[v] + (cand ? (size_t)cand[c] : c);
                fi = m_i[k]; j = m_j[k];
            }
            const int32_t lmi = e->lm_of_row[e->seg[view_id] + fi];
            if (lmi < 0) continue;                                     // feature has no landmark
            const uint32_t lm = (uint32_t)lmi;
            if (e->fd_stamp[j] != (int32_t)oi) continue;
            const int32_t d = e->fd[j];
            if (e->best_lm[j] < 0 || (float)e->best_d[j] > (float)d) {
                e->best_lm[j] = lm;
                e->best_d[j] = d;
            }
        }
    }
    e->x2d.clear();
    e->X3d.clear();
    size_t N = 0;
    for (size_t j = 0; j < nq; ++j) {
        if (e->best_lm[j] < 0) continue;
        if (corr_qfeat) corr_qfeat[N] = (uint32_t)j;
        if (corr_landmark) corr_landmark[N] = (uint32_t)e->best_lm[j];
        e->x2d.push_back(qxy[2 * j]);
        e->x2d.push_back(qxy[2 * j + 1]);
        const double *X = &e->X[3 * (size_t)e->best_lm[j]];
        e->X3d.insert(e->X3d.end(), X, X + 3);
        ++N;
    }
    if (n_corr) *n_corr = N;
    const double t2 = now_ms();

    if (defer) {
        // batched server: the resections of all queries run as one hulo_resect_acransac_batch
        if ((int)N > e->min_points) {
            defer->x2d.insert(defer->x2d.end(), e->x2d.begin(), e->x2d.end());
            defer->X3d.insert(defer->X3d.end(), e->X3d.begin(), e->X3d.end());
        }
        defer->offsets.push_back(defer->x2d.size() / 2);
        if (t_assembly) *t_assembly += t2 - t1;
        return HULO_OK;
    }
    // ---- resection, SfM_Localizer::Localize (LocalizeEngine.cc:529-532) and acceptance (:560)
    if ((int)N > e->min_points) {
        double P[12];
        e->inl.resize(N);
        size_t n_inl = 0;
        double err_max = 0.0;
        int found = 0;
        auto resect = e->resect_schedule == HULO_RESECT_SEQUENTIAL ? hulo_resect_acransac_sequential : hulo_resect_acransac;
        int rc = resect(e->h, e->x2d.data(), e->X3d.data(), N, e->K, e->max_iter, seed, P, e->inl.data(), &n_inl, &err_max,
                        &found);
        if (rc != HULO_OK) return rc;
        if (inliers) memcpy(inliers, e->inl.data(), n_inl * sizeof(int32_t));
        if (n_inliers) *n_inliers = n_inl;
        if (found && (int)n_inl > e->min_inliers) {
            double Kd[9], R[9], t[3];
            krt_from_p(P, Kd, R, t);
            // t_out = -R^T t  (LocalizeEngine.cc:582-585), result = [t_out, R row-major] (:593-602)
            for (int c = 0; c < 3; ++c) pose12[c] = -(R[c] * t[0] + R[3 + c] * t[1] + R[6 + c] * t[2]);
            memcpy(pose12 + 3, R, 9 * sizeof(double));
            *localized = 1;
        }
    }
    const double t3 = now_ms();
    if (t_assembly) *t_assembly += t2 - t1;
    if (t_pnp) *t_pnp += t3 - t2;
    return HULO_OK;
}

int hulo_engine_localize(hulo_engine *e, const uint8_t *qdesc, size_t nq, size_t q_stride, const double *qxy,
                         const uint32_t *views, size_t n_views, uint64_t seed, double *pose12, int *localized,
                         uint32_t *corr_qfeat, uint32_t *corr_landmark, size_t *n_corr, int32_t *inliers,
                         size_t *n_inliers, double *times_ms) {
    HULO_ARG(e != nullptr && pose12 != nullptr && localized != nullptr, "null argument");
    HULO_ARG(nq == 0 || (qdesc != nullptr && qxy != nullptr), "null query");
    *localized = 0;
    if (n_corr) *n_corr = 0;
    if (n_inliers) *n_inliers = 0;
    if (times_ms) times_ms[0] = times_ms[1] = times_ms[2] = times_ms[3] = 0.0;
    if (views == nullptr) n_views = e->n_views;
    const double t0 = now_ms();

    // ---- putative matching, hulo::matchAKAZEToQuery (LocalizeEngine.cc:423)
    const size_t cap = std::max<size_t>(hulo_db_rows(e->map), 1);
    e->m_view.resize(cap);
    e->m_i.resize(cap);
    e->m_j.resize(cap);
    e->m_d0.resize(cap);
    e->view_counts.assign(std::max<size_t>(n_views, 1), 0);
    size_t n_m = 0;
    int rc = hulo_match_to_query(e->h, e->map, views, n_views, qdesc, nq, q_stride, e->ratio, e->m_view.data(),
                                 e->m_i.data(), e->m_j.data(), e->m_d0.data(), cap, &n_m, e->view_counts.data());
    if (rc != HULO_OK) return rc;
    if (times_ms) times_ms[0] = now_ms() - t0;
    return assemble_and_resect(e, nq, qxy, views, n_views, e->m_i.data(), e->m_j.data(), e->m_d0.data(),
                               e->view_counts.data(), seed, pose12, localized, corr_qfeat, corr_landmark, n_corr,
                               inliers, n_inliers, times_ms ? times_ms + 1 : nullptr, times_ms ? times_ms + 2 : nullptr,
                               times_ms ? times_ms + 3 : nullptr);
}

int hulo_partition_views(const uint64_t *rows_per_view, size_t n_views, int world, uint64_t *bounds) {
    HULO_ARG(world >= 1 && bounds != nullptr && (n_views == 0 || rows_per_view != nullptr), "bad argument");
    uint64_t total = 0;
    for (size_t v = 0; v < n_views; ++v) total += rows_per_view[v];
    // contiguous ranges with about total / world rows each: range r starts at the first view whose
    // prefix of rows reaches r * total / world
    bounds[0] = 0;
    size_t v = 0;
    uint64_t prefix = 0;
    for (int r = 1; r < world; ++r) {
        const uint64_t want = (uint64_t)(((unsigned __int128)total * (unsigned)r) / (unsigned)world);
        while (v < n_views && prefix + rows_per_view[v] / 2 < want) prefix += rows_per_view[v++];
        bounds[r] = v;
    }
    bounds[world] = n_views;
    return HULO_OK;
}

int hulo_engine_localize_sharded(hulo_engine *e, const uint8_t *qdesc, size_t nq, size_t q_stride, const double *qxy,
                                 const uint32_t *views, size_t n_views, uint64_t seed, double *pose12, int *localized,
                                 uint32_t *corr_qfeat, uint32_t *corr_landmark, size_t *n_corr, int32_t *inliers,
                                 size_t *n_inliers, double *times_ms) {
    HULO_ARG(e != nullptr && pose12 != nullptr && localized != nullptr, "null argument");
    HULO_ARG(nq == 0 || (qdesc != nullptr && qxy != nullptr), "null query");
    const int world = hulo_comm_world(e->h), rank = hulo_comm_rank(e->h);
    if (world == 1)
        return hulo_engine_localize(e, qdesc, nq, q_stride, qxy, views, n_views, seed, pose12, localized, corr_qfeat,
                                    corr_landmark, n_corr, inliers, n_inliers, times_ms);
    *localized = 0;
    if (n_corr) *n_corr = 0;
    if (n_inliers) *n_inliers = 0;
    if (times_ms) times_ms[0] = times_ms[1] = times_ms[2] = times_ms[3] = 0.0;
    if (views == nullptr) n_views = e->n_views;
    const double t0 = now_ms();

    // ---- this rank's share of the views: a contiguous range of the list, balanced by rows
    std::vector<uint64_t> rows(n_views), bounds((size_t)world + 1);
    for (size_t v = 0; v < n_views; ++v) {
        const size_t s = views ? views[v] : v;
        HULO_ARG(s < e->n_views, "view index out of range");
        rows[v] = e->seg[s + 1] - e->seg[s];
    }
    hulo_partition_views(rows.data(), n_views, world, bounds.data());
    const size_t v0 = (size_t)bounds[(size_t)rank], nv = (size_t)(bounds[(size_t)rank + 1] - bounds[(size_t)rank]);
    size_t max_nv = 0;
    for (int r = 0; r < world; ++r) max_nv = std::max<size_t>(max_nv, (size_t)(bounds[(size_t)r + 1] - bounds[(size_t)r]));
    std::vector<uint32_t> my_views(std::max<size_t>(nv, 1));
    for (size_t k = 0; k < nv; ++k) my_views[k] = views ? views[v0 + k] : (uint32_t)(v0 + k);

    // ---- putative matching of the local views (the query goes to every rank: the broadcast)
    uint64_t local_rows = 0;
    for (size_t k = 0; k < nv; ++k) local_rows += rows[v0 + k];
    const size_t cap = std::max<size_t>((size_t)local_rows, 1);
    e->m_view.resize(cap); e->m_i.resize(cap); e->m_j.resize(cap); e->m_d0.resize(cap);
    std::vector<uint32_t> local_counts(std::max<size_t>(nv, 1), 0);
    size_t n_m = 0;
    int rc;

    // ---- exchange.  Block of a rank: {n_matches, counts[max_nv], records[slots] of (i, j, d0)}.  One
    // all-gather when every rank's matches fit the slots, else a second one sized by the largest.
    // sized for the usual yield (a few matches per query descriptor over all views, shared by the
    // ranks) with a factor of two to spare
    size_t slots = std::max<size_t>(256, 4 * nq / (size_t)world);
    std::vector<uint32_t> block, all;
    std::vector<uint64_t> rank_n((size_t)world);
    // Default: the survivors never leave the device on their own -- the compaction's output goes
    // straight into every rank's peer-mapped exchange buffer (stores over NVLink, flags), and one
    // copy brings all ranks' blocks to the host: one synchronisation per query.
    bool on_device = false;
    if (strcmp(hulo_comm_exchange_kind(e->h), "peer-store") == 0) {
        hulo::QueryMatchesDev dm;
        rc = hulo::match_to_query_device(e->h, e->map, my_views.data(), nv, qdesc, nq, q_stride, e->ratio, &dm);
        if (rc != HULO_OK) return rc;
        for (int round = 0; round < 2; ++round) {
            const size_t words = 1 + max_nv + 3 * slots;
            all.resize(words * (size_t)world);
            rc = hulo::gather_query_matches(e->h, dm, nv, max_nv, slots, all.data());
            if (rc == hulo::kNoPeerExchange) break;
            if (rc != HULO_OK) return rc;
            on_device = true;
            size_t largest = 0;
            for (int r = 0; r < world; ++r) {
                rank_n[(size_t)r] = all[words * (size_t)r];
                largest = std::max<size_t>(largest, (size_t)rank_n[(size_t)r]);
            }
            if (largest <= slots) break;
            slots = largest;                   // the same decision on every rank
        }
    }
    if (!on_device) {
    // peers not mappable (or HULO_EXCHANGE=nccl): matches to the host, all-gather staged through host buffers
    rc = hulo_match_to_query(e->h, e->map, my_views.data(), nv, qdesc, nq, q_stride, e->ratio, e->m_view.data(),
                             e->m_i.data(), e->m_j.data(), e->m_d0.data(), cap, &n_m, local_counts.data());
    if (rc != HULO_OK) return rc;
    for (int round = 0; round < 2; ++round) {
        const size_t words = 1 + max_nv + 3 * slots;
        block.assign(words, 0);
        block[0] = (uint32_t)n_m;
        for (size_t k = 0; k < nv; ++k) block[1 + k] = local_counts[k];
        const size_t fit = std::min(n_m, slots);
        for (size_t k = 0; k < fit; ++k) {
            block[1 + max_nv + 3 * k] = e->m_i[k];
            block[1 + max_nv + 3 * k + 1] = e->m_j[k];
            block[1 + max_nv + 3 * k + 2] = (uint32_t)e->m_d0[k];
        }
        all.resize(words * (size_t)world);
        rc = hulo_comm_allgather(e->h, block.data(), words * sizeof(uint32_t), all.data());
        if (rc != HULO_OK) return rc;
        size_t largest = 0;
        for (int r = 0; r < world; ++r) {
            rank_n[(size_t)r] = all[words * (size_t)r];
            largest = std::max<size_t>(largest, (size_t)rank_n[(size_t)r]);
        }
        if (largest <= slots) break;
        slots = largest;                       // the same decision on every rank
    }
    }
    // the matches of all views in list order: exactly what one GPU emits for the whole list
    const size_t words = 1 + max_nv + 3 * slots;
    size_t total = 0;
    for (int r = 0; r < world; ++r) total += (size_t)rank_n[(size_t)r];
    e->m_i.resize(std::max<size_t>(total, 1)); e->m_j.resize(std::max<size_t>(total, 1)); e->m_d0.resize(std::max<size_t>(total, 1));
    e->view_counts.assign(std::max<size_t>(n_views, 1), 0);
    size_t at = 0;
    for (int r = 0; r < world; ++r) {
        const uint32_t *b = all.data() + words * (size_t)r;
        const size_t rv0 = (size_t)bounds[(size_t)r], rnv = (size_t)(bounds[(size_t)r + 1] - bounds[(size_t)r]);
        for (size_t k = 0; k < rnv; ++k) e->view_counts[rv0 + k] = b[1 + k];
        for (size_t k = 0; k < (size_t)rank_n[(size_t)r]; ++k, ++at) {
            e->m_i[at] = b[1 + max_nv + 3 * k];
            e->m_j[at] = b[1 + max_nv + 3 * k + 1];
            e->m_d0[at] = (int32_t)b[1 + max_nv + 3 * k + 2];
        }
    }
    if (times_ms) times_ms[0] = now_ms() - t0;
    // ---- every rank finishes the query on its own: same matches, same seed, same pose
    return assemble_and_resect(e, nq, qxy, views, n_views, e->m_i.data(), e->m_j.data(), e->m_d0.data(),
                               e->view_counts.data(), seed, pose12, localized, corr_qfeat, corr_landmark, n_corr,
                               inliers, n_inliers, times_ms ? times_ms + 1 : nullptr, times_ms ? times_ms + 2 : nullptr,
                               times_ms ? times_ms + 3 : nullptr);
}

int hulo_engine_localize_batch(hulo_engine *e, size_t n_queries, const uint8_t *qdesc, size_t q_stride,
                               const uint64_t *q_offsets, const double *qxy, const uint32_t *views, size_t n_views,
                               uint64_t seed, double *pose12, int *localized, uint32_t *n_corr, uint32_t *n_inliers,
                               double *times_ms) {
    HULO_ARG(e != nullptr, "null engine");
    HULO_ARG(n_queries == 0 || (q_offsets != nullptr && pose12 != nullptr && localized != nullptr), "null argument");
    if (times_ms) times_ms[0] = times_ms[1] = times_ms[2] = times_ms[3] = 0.0;
    if (n_queries == 0) return HULO_OK;
    HULO_ARG(q_offsets[n_queries] == 0 || (qdesc != nullptr && qxy != nullptr), "null query");
    if (views == nullptr) n_views = e->n_views;
    const double t0 = now_ms();
    // ---- putative matching of every query image in one pass
    std::vector<uint32_t> counts(n_queries * std::max<size_t>(n_views, 1), 0);
    // room for four matches per query descriptor; a retry (which repeats the whole pass) only
    // happens beyond that
    size_t cap = std::max<size_t>(std::max<size_t>(e->m_i.size(), 1 << 20), 4 * (size_t)q_offsets[n_queries]), n_m = 0;
    for (;;) {
        e->m_i.resize(cap);
        e->m_j.resize(cap);
        e->m_d0.resize(cap);
        int rc = hulo_match_to_queries(e->h, e->map, views, n_views, qdesc, q_stride, q_offsets, n_queries, e->ratio,
                                       nullptr, nullptr, e->m_i.data(), e->m_j.data(), e->m_d0.data(), cap, &n_m,
                                       counts.data());
        if (rc == HULO_ERR_CAPACITY) { cap = n_m + n_m / 8; continue; }
        if (rc != HULO_OK) return rc;
        break;
    }
    if (times_ms) times_ms[0] = now_ms() - t0;
    size_t k0 = 0;
    ResectionBatch rb;
    for (size_t q = 0; q < n_queries; ++q) {
        const size_t nq = (size_t)(q_offsets[q + 1] - q_offsets[q]);
        const uint32_t *vc = counts.data() + q * n_views;
        size_t n_q_matches = 0;
        for (size_t v = 0; v < n_views; ++v) n_q_matches += vc[v];
        localized[q] = 0;
        size_t nc = 0, ni = 0;
        int rc = assemble_and_resect(e, nq, qxy + 2 * q_offsets[q], views, n_views, e->m_i.data() + k0,
                                     e->m_j.data() + k0, e->m_d0.data() + k0, vc, seed + q, pose12 + 12 * q,
                                     localized + q, nullptr, nullptr, &nc, nullptr, &ni,
                                     times_ms ? times_ms + 1 : nullptr, times_ms ? times_ms + 2 : nullptr,
                                     times_ms ? times_ms + 3 : nullptr, &rb, (size_t)q_offsets[q]);
        if (rc != HULO_OK) return rc;
        if (n_corr) n_corr[q] = (uint32_t)nc;
        if (n_inliers) n_inliers[q] = 0;
        k0 += n_q_matches;
    }
    // ---- resection of every query at once (SfM_Localizer::Localize, LocalizeEngine.cc:529-532),
    // query q with the seed the one-query entry point would give it, and acceptance (:560)
    const double t_r0 = now_ms();
    std::vector<double> P(12 * n_queries), emax(n_queries), Ks(9 * n_queries);
    std::vector<uint64_t> n_inl(n_queries), seeds(n_queries);
    std::vector<int32_t> found(n_queries), inl(std::max<size_t>(rb.x2d.size() / 2, 1));
    for (size_t q = 0; q < n_queries; ++q) {
        seeds[q] = seed + q;
        memcpy(&Ks[9 * q], e->K, 9 * sizeof(double));
    }
    int rc = hulo_resect_acransac_batch(e->h, n_queries, rb.offsets.data(), rb.x2d.data(), rb.X3d.data(), Ks.data(),
                                        e->max_iter, seed, seeds.data(), P.data(), inl.data(), n_inl.data(),
                                        emax.data(), found.data());
    if (rc != HULO_OK) return rc;
    for (size_t q = 0; q < n_queries; ++q) {
        if (n_inliers) n_inliers[q] = (uint32_t)n_inl[q];
        if (found[q] && (int)n_inl[q] > e->min_inliers) {
            double Kd[9], R[9], t[3];
            krt_from_p(&P[12 * q], Kd, R, t);
            double *pose = pose12 + 12 * q;
            for (int c = 0; c < 3; ++c) pose[c] = -(R[c] * t[0] + R[3 + c] * t[1] + R[6 + c] * t[2]);
            memcpy(pose + 3, R, 9 * sizeof(double));
            localized[q] = 1;
        }
    }
    if (times_ms) times_ms[2] += now_ms() - t_r0;
    return HULO_OK;
}

}  // extern "C"
