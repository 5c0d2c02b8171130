#include "localize_engine.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <iostream>
#include <stdexcept>

#include "../../../include/hulo_gpu.h"
#include "desc_files.h"

namespace hulo {

namespace {
void must(int status, const char *what) {
    if (status != HULO_OK) throw std::runtime_error(std::string(what) + ": " + hulo_last_error());
}
}  // namespace

struct LocalizeEngine::State {
    hulo_gpu *gpu = nullptr;
    hulo_engine *eng = nullptr;
    hulo_bow *bow = nullptr;                           // the views' bag-of-features vectors (bowKnnNum > 0)
    std::map<std::size_t, uint32_t> bow_row_of_view;
    std::vector<std::size_t> bow_view_of_row;
    std::size_t bow_dim = 0;
    int bowKnnNum = 0;
    SfMScene scene;
    std::string sfmDataDir, matchDir;
    double ratio = 0.6;
    int ransacRound = 25;
    double ransacPrecision = 4.0;
    bool guidedMatching = false;
    bool hasA = false;
    double A[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};   // row-major 4 x 4 (3 x 4 padded)
    std::vector<std::size_t> seg_view;                 // segment s of the device table = view seg_view[s]
    std::map<std::size_t, uint32_t> seg_of_view;
    std::vector<std::size_t> landmark_id;              // dense landmark index -> Landmark::id
    std::set<std::size_t> cli_local_views;             // setLocalViews (CLI flavour); empty = unrestricted
    bool cli_restricted = false;
    LastResult last;
    ~State() {
        if (bow) hulo_bow_destroy(bow);
        if (eng) hulo_engine_destroy(eng);
        if (gpu) hulo_gpu_destroy(gpu);
    }
};

LocalizeEngine::LocalizeEngine() : st_(std::make_shared<State>()) {}

LocalizeEngine::LocalizeEngine(const std::string sfmDataDir, const std::string matchDir, const std::string AmatFile,
                               double secondTestRatio, int ransacRound, double ransacPrecision, bool guidedMatching,
                               int beaconKnnNum, int bowKnnNum, int device)
    : st_(std::make_shared<State>()) {
    State &s = *st_;
    s.guidedMatching = guidedMatching;
    s.sfmDataDir = sfmDataDir;
    s.matchDir = matchDir;
    s.ratio = secondTestRatio;
    s.ransacRound = ransacRound;
    s.ransacPrecision = ransacPrecision;
    const std::string sSfM_data = sfmDataDir + (sfmDataDir.empty() || sfmDataDir.back() == '/' ? "" : "/") + "sfm_data.json";
    std::cout << "Reading sfm_data.json file : " << sSfM_data << std::endl;
    if (!loadSfMData(sSfM_data, s.scene))
        throw std::runtime_error("The input sfm_data.json file \"" + sSfM_data + "\" cannot be read.");
    if (beaconKnnNum > 0) std::cout << "iBeacon view selection is outside this library: beaconKnnNum ignored" << std::endl;

    // the global-coordinate matrix A (LocalizeEngine.cc:113-119); landmarks and camera centres are
    // moved into global coordinates once (TRANSFORM_SFM_DATA_BEFORE_LOCALIZE, :61, :122-144)
    if (!AmatFile.empty()) {
        int r = 0, c = 0;
        std::vector<double> a;
        std::cout << "Reading A mat file : " << AmatFile << std::endl;
        if (readOpenCVMatrix(AmatFile, "A", r, c, a) && c == 4 && (r == 3 || r == 4)) {
            memcpy(s.A, a.data(), sizeof(double) * 4 * r);
            s.hasA = true;
        } else {
            std::cerr << "Cannot find A mat file" << std::endl;
        }
    }
    if (s.hasA) {
        const double *A = s.A;
        for (Landmark &lm : s.scene.landmarks) {
            const double x = lm.X[0], y = lm.X[1], z = lm.X[2];
            for (int i = 0; i < 3; ++i) lm.X[i] = A[4 * i] * x + A[4 * i + 1] * y + A[4 * i + 2] * z + A[4 * i + 3];
        }
        for (auto &kv : s.scene.poses) {
            Pose &p = kv.second;
            const double x = p.center[0], y = p.center[1], z = p.center[2];
            for (int i = 0; i < 3; ++i) p.center[i] = A[4 * i] * x + A[4 * i + 1] * y + A[4 * i + 2] * z + A[4 * i + 3];
            double Rn[9];   // rot * A(0:3,0:3)^T
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) Rn[3 * i + j] = p.R[3 * i] * A[4 * j] + p.R[3 * i + 1] * A[4 * j + 1] + p.R[3 * i + 2] * A[4 * j + 2];
            memcpy(p.R, Rn, sizeof Rn);
        }
    }

    // descriptor rows and feature positions of every view (HuloSfMRegionsProvider::load, :103-108)
    std::vector<uint8_t> all, rows;
    std::vector<double> map_xy;
    std::vector<uint64_t> off(1, 0);
    std::vector<int32_t> view_wh;
    for (const auto &kv : s.scene.views) {
        std::size_t n = 0;
        FeatureLocations feats;
        readAKAZEBin(descPath(matchDir, kv.second.s_Img_path, true), rows, n);
        if (!readFeatFile(featPath(matchDir, kv.second.s_Img_path), feats) || feats.size() != n) {
            std::cerr << "Cannot construct regions providers: view " << kv.first << std::endl;
            n = std::min(n, feats.size());
            rows.resize(n * HULO_ROW_BYTES);
            feats.resize(n);
        }
        const Intrinsic *cam = nullptr;
        auto ci = s.scene.intrinsics.find(kv.second.id_intrinsic);
        if (ci != s.scene.intrinsics.end()) cam = &ci->second;
        all.insert(all.end(), rows.begin(), rows.end());
        for (const auto &f : feats) {
            // MatchesPairToMat feeds the geometric filter with cam->get_ud_pixel(feature)
            const std::pair<double, double> ud = cam ? cam->get_ud_pixel(f.first, f.second) : f;
            map_xy.push_back(ud.first);
            map_xy.push_back(ud.second);
        }
        off.push_back(off.back() + n);
        s.seg_of_view[kv.first] = (uint32_t)s.seg_view.size();
        s.seg_view.push_back(kv.first);
        view_wh.push_back((int32_t)std::max<std::size_t>(kv.second.ui_width, 1));
        view_wh.push_back((int32_t)std::max<std::size_t>(kv.second.ui_height, 1));
    }
    // (view, feature) -> landmark (hulo::structureToMapViewFeatTo3D, SfMDataUtils.cpp:33-46)
    std::vector<uint32_t> ov, of, ol;
    std::vector<double> X;
    for (std::size_t k = 0; k < s.scene.landmarks.size(); ++k) {
        const Landmark &lm = s.scene.landmarks[k];
        s.landmark_id.push_back(lm.id);
        X.insert(X.end(), lm.X, lm.X + 3);
        for (const Observation &o : lm.obs) {
            auto sv = s.seg_of_view.find(o.id_view);
            if (sv == s.seg_of_view.end()) continue;
            if (o.id_feat >= off[sv->second + 1] - off[sv->second]) continue;   // feature file shorter than the track
            ov.push_back(sv->second); of.push_back((uint32_t)o.id_feat); ol.push_back((uint32_t)k);
        }
    }
    // intrinsics of the query camera: intrinsic 0 (LocalizeEngine.cc:509-510)
    double K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    auto i0 = s.scene.intrinsics.find(0);
    if (i0 == s.scene.intrinsics.end()) throw std::runtime_error("sfm_data.json has no intrinsic 0");
    i0->second.K(K);

    must(hulo_gpu_create(device, &s.gpu), "hulo_gpu_create");
    static const uint8_t none = 0;
    static const double zero3[3] = {0, 0, 0};
    must(hulo_engine_create(s.gpu, all.empty() ? &none : all.data(), (std::size_t)off.back(), HULO_ROW_BYTES, off.data(),
                            s.seg_view.size(), ov.data(), of.data(), ol.data(), ov.size(), X.empty() ? zero3 : X.data(),
                            s.scene.landmarks.size(), K, &s.eng),
         "hulo_engine_create");
    // MINUM_NUMBER_OF_POINT_PUTATIVE_MATCH 16, _RESECTION 8, _INLIER_RESECTION 10 (LocalizeEngine.cc:63-65)
    must(hulo_engine_configure(s.eng, (float)secondTestRatio, 16, 8, 10, 4096), "hulo_engine_configure");
    s.last.localized = false;
    memcpy(s.last.K, K, sizeof K);
    // BoW model of the views (LocalizeEngine.cc:146-179): the .bow vectors, resident on the device
    if (bowKnnNum > 0) {
        std::vector<float> all;
        std::vector<double> vec;
        for (const auto &kv : s.scene.views) {
            int r = 0, c = 0;
            if (!readMatBin(bowPath(matchDir, kv.second.s_Img_path), r, c, vec) || vec.empty()) continue;
            if (s.bow_dim == 0) s.bow_dim = vec.size();
            if (vec.size() != s.bow_dim) continue;
            s.bow_row_of_view[kv.first] = (uint32_t)s.bow_view_of_row.size();
            s.bow_view_of_row.push_back(kv.first);
            for (double v : vec) all.push_back((float)v);
        }
        if (s.bow_view_of_row.empty()) {
            std::cout << "cannot find BOW vectors of the views, localize without using BOW model" << std::endl;
        } else {
            must(hulo_bow_create(s.gpu, all.data(), s.bow_view_of_row.size(), s.bow_dim, &s.bow), "hulo_bow_create");
            s.bowKnnNum = bowKnnNum;
        }
    }
    // keypoints are set now; the query image size is only known per call (hulo_engine_set_query_size)
    std::vector<double> xy1(2, 0.0);
    must(hulo_engine_set_keypoints(s.eng, map_xy.empty() ? xy1.data() : map_xy.data(), view_wh.data(), 1, 1),
         "hulo_engine_set_keypoints");
}

const LocalizeEngine::LastResult &LocalizeEngine::last() const { return st_->last; }
const SfMScene &LocalizeEngine::scene() const { return st_->scene; }

void LocalizeEngine::setLocalViews(const std::vector<double> &center, double radius) {
    State &s = *st_;
    s.cli_local_views.clear();
    s.cli_restricted = center.size() == 3 && radius > 0;
    if (!s.cli_restricted) return;
    for (const auto &kv : s.scene.views) {
        auto p = s.scene.poses.find(kv.second.id_pose);
        if (p == s.scene.poses.end()) continue;
        double d2 = 0;
        for (int i = 0; i < 3; ++i) d2 += (p->second.center[i] - center[i]) * (p->second.center[i] - center[i]);
        if (d2 <= radius) s.cli_local_views.insert(kv.first);      // squaredNorm() <= radius, SfMDataUtils.cpp:218-219
    }
}

std::vector<double> LocalizeEngine::localize(const uint8_t *desc, std::size_t n, std::size_t stride,
                                             const FeatureLocations &qFeatLoc, std::size_t imageWidth,
                                             std::size_t imageHeight, const std::string &beaconStr,
                                             bool bReturnKeypoints, std::vector<double> &points2D,
                                             std::vector<double> &points3D, std::vector<int> &pointsInlier,
                                             bool bReturnTime, std::vector<double> &times,
                                             const std::vector<double> &center, double radius, uint64_t seed,
                                             const std::vector<float> *queryBow) {
    State &s = *st_;
    std::vector<double> result;
    s.last.localized = false;
    s.last.inlier_pairs.clear();
    if (!s.eng) return result;
    if (qFeatLoc.size() != n) throw std::invalid_argument("LocalizeEngine::localize: one feature position per descriptor row");

    // ---- view selection (LocalizeEngine.cc:272-293): views with a pose, optionally within `radius`
    // of `center`.  The reference's engine multiplies the (already transformed) centre by A once
    // more (:262-283 after :135-143); reproduced as it is.
    std::vector<uint32_t> views;
    const bool restricted = center.size() == 3 && radius > 0;
    for (const auto &kv : s.scene.views) {
        auto p = s.scene.poses.find(kv.second.id_pose);
        if (p == s.scene.poses.end()) continue;
        if (restricted) {
            const double *c = p->second.center, *A = s.A;
            double d2 = 0;
            for (int i = 0; i < 3; ++i) {
                const double g = A[4 * i] * c[0] + A[4 * i + 1] * c[1] + A[4 * i + 2] * c[2] + A[4 * i + 3];
                d2 += (g - center[i]) * (g - center[i]);
            }
            if (std::sqrt(d2) > radius) continue;
        }
        if (s.cli_restricted && !s.cli_local_views.count(kv.first)) continue;
        views.push_back(s.seg_of_view.at(kv.first));
    }
    if (restricted) {
        std::cout << "number of selected local views by center location : " << views.size() << std::endl;
        if (views.empty()) return result;
    }
    // ---- BoW pre-selection (LocalizeEngine.cc:334-362): the knn views nearest to the query in
    // bag-of-features space, when there are more candidates than knn
    if (s.bow && s.bowKnnNum > 0 && queryBow && queryBow->size() == s.bow_dim && views.size() > (std::size_t)s.bowKnnNum) {
        std::vector<uint32_t> rows;
        for (uint32_t seg : views) {
            auto it = s.bow_row_of_view.find(s.seg_view[seg]);
            if (it != s.bow_row_of_view.end()) rows.push_back(it->second);
        }
        if (rows.size() > (std::size_t)s.bowKnnNum) {
            std::vector<int32_t> idx((std::size_t)s.bowKnnNum);
            must(hulo_bow_knn(s.bow, queryBow->data(), rows.data(), rows.size(), (std::size_t)s.bowKnnNum, idx.data(), nullptr),
                 "hulo_bow_knn");
            std::set<std::size_t> chosen;
            for (int32_t r : idx) chosen.insert(s.bow_view_of_row[(std::size_t)r]);
            views.clear();
            for (std::size_t v : chosen) views.push_back(s.seg_of_view.at(v));     // ascending view id, like the std::set
            std::cout << "number of selected local views by bow : " << views.size() << std::endl;
        }
    }
    if (views.empty()) {                       // no pair to match: map_putativeMatches stays empty (:439-454)
        std::cout << "Not enough putative matches" << std::endl;
        return result;
    }

    // ---- query regions: undistorted positions for the geometric filter and the resection (:519-523)
    const Intrinsic &cam = s.scene.intrinsics.at(0);
    std::vector<double> qxy(2 * std::max<std::size_t>(n, 1));
    for (std::size_t k = 0; k < n; ++k) {
        const std::pair<double, double> ud = cam.get_ud_pixel(qFeatLoc[k].first, qFeatLoc[k].second);
        qxy[2 * k] = ud.first;
        qxy[2 * k + 1] = ud.second;
    }
    must(hulo_engine_set_query_size(s.eng, (int)std::max<std::size_t>(imageWidth, 1), (int)std::max<std::size_t>(imageHeight, 1)),
         "hulo_engine_set_query_size");
    must(hulo_engine_configure_geometric(s.eng, 1, (std::size_t)std::max(s.ransacRound, 1), s.ransacPrecision),
         "hulo_engine_configure_geometric");
    must(hulo_engine_set_guided_matching(s.eng, s.guidedMatching ? 1 : 0), "hulo_engine_set_guided_matching");

    double pose12[12];
    int localized = 0;
    std::vector<uint32_t> cq(std::max<std::size_t>(n, 1)), cl(std::max<std::size_t>(n, 1));
    std::vector<int32_t> inl(std::max<std::size_t>(n, 1));
    std::size_t n_corr = 0, n_inl = 0;
    double tm[4] = {0, 0, 0, 0};
    static const uint8_t none = 0;
    must(hulo_engine_localize(s.eng, n ? desc : &none, n, n ? stride : HULO_ROW_BYTES, qxy.data(), views.data(),
                              views.size(), seed, pose12, &localized, cq.data(), cl.data(), &n_corr, inl.data(), &n_inl,
                              tm),
         "hulo_engine_localize");
    if (bReturnTime) {
        times.clear();
        times.push_back(0.0);                       // select views beacon
        times.push_back(0.0);                       // select views bow
        times.push_back(0.0);                       // extract feature (the caller's)
        times.push_back(tm[0] * 1e-3);              // putative matching
        times.push_back(tm[3] * 1e-3);              // geometric matching
        times.push_back((tm[1] + tm[2]) * 1e-3);    // PnP (2D-3D assembly + resection)
    }
    if (bReturnKeypoints) {                          // :534-557
        points2D.clear(); points3D.clear(); pointsInlier.clear();
        for (std::size_t k = 0; k < n_corr; ++k) {
            points2D.push_back(qxy[2 * cq[k]]);
            points2D.push_back(qxy[2 * cq[k] + 1]);
            const Landmark &lm = s.scene.landmarks[cl[k]];
            points3D.insert(points3D.end(), lm.X, lm.X + 3);
        }
        for (std::size_t k = 0; k < n_inl; ++k) pointsInlier.push_back((int)inl[k]);
    }
    if (!localized) {
        std::cout << "Fail to estimate camera matrix" << std::endl;
        return result;
    }
    result.assign(pose12, pose12 + 12);              // [t_out, R row-major], :593-602
    s.last.localized = true;
    memcpy(s.last.t_out, pose12, 3 * sizeof(double));
    memcpy(s.last.R, pose12 + 3, 9 * sizeof(double));
    for (std::size_t k = 0; k < n_inl; ++k)
        s.last.inlier_pairs.push_back(std::make_pair((std::size_t)cq[inl[k]], s.landmark_id[cl[inl[k]]]));
    return result;
}

}  // namespace hulo
