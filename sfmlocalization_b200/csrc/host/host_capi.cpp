// host_capi.cpp -- a small extern "C" surface over the C++ host layer so the CPU test-suite can
// exercise the parts that need no GPU (file formats, pair lists, track propagation) through
// ctypes.  Not part of the drop-in boundary (that is include/hulo_gpu.h).
#include <cstdio>
#include <cstring>
#include <set>

#include "desc_files.h"
#include "match_utils_gpu.h"
#include "pair_lists.h"
#include "sfm_data_io.h"

using namespace hulo;

extern "C" {

// .desc round trip: returns row count or -1
long long hulo_host_read_desc(const char *path, unsigned char *rows64, unsigned long long cap_rows) {
    std::vector<uint8_t> r;
    std::size_t n = 0;
    if (!readAKAZEBin(path, r, n)) return -1;
    if (rows64 && n <= cap_rows) memcpy(rows64, r.data(), r.size());
    return (long long)n;
}
int hulo_host_write_desc(const char *path, const unsigned char *rows, unsigned long long n, unsigned long long width) {
    return saveAKAZEBin(path, rows, (std::size_t)n, (std::size_t)width) ? 0 : 1;
}

// pair generators over views with ids 0..n-1 (or arbitrary ids given): out holds 2 * cap entries
static Views make_views(const unsigned long long *ids, unsigned long long n) {
    Views v;
    for (unsigned long long k = 0; k < n; ++k) v[(std::size_t)ids[k]] = View{(std::size_t)ids[k], ""};
    return v;
}
static unsigned long long emit(const std::vector<Pair> &p, unsigned long long *out, unsigned long long cap) {
    for (std::size_t k = 0; k < p.size() && k < cap; ++k) { out[2 * k] = p[k].first; out[2 * k + 1] = p[k].second; }
    return p.size();
}
unsigned long long hulo_host_all_pairs(const unsigned long long *ids, unsigned long long n, unsigned long long *out,
                                       unsigned long long cap) {
    std::vector<Pair> p;
    generateAllPairs(make_views(ids, n), p);
    return emit(p, out, cap);
}
unsigned long long hulo_host_video_pairs(const unsigned long long *ids, unsigned long long n, int frame,
                                         unsigned long long *out, unsigned long long cap) {
    std::vector<Pair> p;
    generateVideoMatchPairs(make_views(ids, n), p, frame);
    return emit(p, out, cap);
}
unsigned long long hulo_host_remove_dup_pairs(unsigned long long *pairs, unsigned long long n) {
    std::vector<Pair> p;
    for (unsigned long long k = 0; k < n; ++k) p.push_back(Pair(pairs[2 * k], pairs[2 * k + 1]));
    removeDupPairs(p);
    return emit(p, pairs, n);
}
// positions of `rank`'s pairs; rows[v] = descriptor count of view v (ids 0..n_views-1)
unsigned long long hulo_host_partition_pairs(const unsigned long long *pairs, unsigned long long n,
                                             const unsigned long long *rows, unsigned long long n_views, int rank,
                                             int world, unsigned long long *out_pos) {
    std::vector<Pair> p;
    for (unsigned long long k = 0; k < n; ++k) p.push_back(Pair(pairs[2 * k], pairs[2 * k + 1]));
    std::map<std::size_t, std::size_t> r;
    for (unsigned long long v = 0; v < n_views; ++v) r[(std::size_t)v] = (std::size_t)rows[v];
    const std::vector<std::size_t> mine = partitionPairs(p, r, rank, world);
    for (std::size_t k = 0; k < mine.size(); ++k) out_pos[k] = mine[k];
    return mine.size();
}

// track propagation: consecutive matches in (m_off, m_i, m_j) form -> appended (f, to, i, j)
unsigned long long hulo_host_propagate_tracks(unsigned long long n_frames, unsigned long long max_frame_dist,
                                              const int *feat_number, const long long *m_off, const int *m_i,
                                              const int *m_j, int *out4, unsigned long long cap) {
    PairWiseMatches m;
    std::vector<int> fn(feat_number, feat_number + (n_frames ? n_frames - 1 : 0));
    for (unsigned long long f = 0; f + 1 < n_frames; ++f)
        for (long long k = m_off[f]; k < m_off[f + 1]; ++k)
            m[Pair(f, f + 1)].push_back(IndMatch((uint32_t)m_i[k], (uint32_t)m_j[k]));
    propagateTracks((std::size_t)n_frames, (std::size_t)max_frame_dist, fn, m);
    unsigned long long n = 0;
    for (const auto &kv : m) {
        if (kv.first.second == kv.first.first + 1) continue;
        for (const IndMatch &im : kv.second) {
            if (n < cap) {
                out4[4 * n] = (int)kv.first.first; out4[4 * n + 1] = (int)kv.first.second;
                out4[4 * n + 2] = (int)im.i_; out4[4 * n + 3] = (int)im.j_;
            }
            ++n;
        }
    }
    return n;
}

// match file round trip through the C++ writer / reader: returns number of pairs read back
long long hulo_host_matches_roundtrip(const char *in_path, const char *out_path) {
    PairWiseMatches m;
    if (!importPairWiseMatches(in_path, m)) return -1;
    if (!exportPairWiseMatches(m, out_path)) return -2;
    return (long long)m.size();
}

long long hulo_host_views_from_sfm_data(const char *path, unsigned long long *ids, char *names, unsigned long long cap,
                                        unsigned long long name_stride) {
    Views v;
    if (!readViewsFromSfmData(path, v)) return -1;
    unsigned long long k = 0;
    for (const auto &kv : v) {
        if (k < cap) {
            ids[k] = kv.first;
            strncpy(names + k * name_stride, kv.second.s_Img_path.c_str(), name_stride - 1);
            names[k * name_stride + name_stride - 1] = 0;
        }
        ++k;
    }
    return (long long)k;
}

// sfm_data.json reader: counts = {views, intrinsics, poses, landmarks, observations};
// intrinsic0 = {focal, ppx, ppy, width, height, k1, k2, k3}; view_wh: 2 per view (ascending id)
long long hulo_host_load_sfm_data(const char *path, unsigned long long *counts, double *first_X, double *intrinsic0,
                                  unsigned long long *view_wh, unsigned long long cap_views) {
    SfMScene sc;
    if (!loadSfMData(path, sc)) return -1;
    unsigned long long n_obs = 0;
    for (const Landmark &lm : sc.landmarks) n_obs += lm.obs.size();
    counts[0] = sc.views.size(); counts[1] = sc.intrinsics.size(); counts[2] = sc.poses.size();
    counts[3] = sc.landmarks.size(); counts[4] = n_obs;
    if (first_X && !sc.landmarks.empty()) memcpy(first_X, sc.landmarks[0].X, 3 * sizeof(double));
    if (intrinsic0 && sc.intrinsics.count(0)) {
        const Intrinsic &in = sc.intrinsics.at(0);
        intrinsic0[0] = in.focal; intrinsic0[1] = in.ppx; intrinsic0[2] = in.ppy;
        intrinsic0[3] = (double)in.width; intrinsic0[4] = (double)in.height;
        for (int k = 0; k < 3; ++k) intrinsic0[5 + k] = (size_t)k < in.disto.size() ? in.disto[k] : 0.0;
    }
    unsigned long long k = 0;
    for (const auto &kv : sc.views) {
        if (view_wh && k < cap_views) { view_wh[2 * k] = kv.second.ui_width; view_wh[2 * k + 1] = kv.second.ui_height; }
        ++k;
    }
    return (long long)sc.views.size();
}

// sums of the observed image points (checks that Observation::x is read)
int hulo_host_sfm_observation_sums(const char *path, double *sums2) {
    SfMScene sc;
    if (!loadSfMData(path, sc)) return 1;
    sums2[0] = sums2[1] = 0.0;
    for (const Landmark &lm : sc.landmarks)
        for (const Observation &o : lm.obs) { sums2[0] += o.x[0]; sums2[1] += o.x[1]; }
    return 0;
}

int hulo_host_save_sfm_poses(const char *in_json, const char *out_json, unsigned long long n,
                             const unsigned long long *ids, const double *R, const double *center) {
    std::map<std::size_t, Pose> poses;
    for (unsigned long long k = 0; k < n; ++k) {
        Pose p;
        memcpy(p.R, R + 9 * k, sizeof p.R);
        memcpy(p.center, center + 3 * k, sizeof p.center);
        poses[(std::size_t)ids[k]] = p;
    }
    return saveSfMDataPoses(in_json, out_json, poses) ? 0 : 1;
}

void hulo_host_undistort(double focal, double ppx, double ppy, double k1, double k2, double k3, double x, double y,
                         double *out2) {
    Intrinsic in;
    in.focal = focal; in.ppx = ppx; in.ppy = ppy;
    in.disto = {k1, k2, k3};
    const std::pair<double, double> ud = in.get_ud_pixel(x, y);
    out2[0] = ud.first; out2[1] = ud.second;
}

int hulo_host_read_cv_matrix(const char *path, const char *name, double *out, int cap, int *rows, int *cols) {
    std::vector<double> d;
    if (!readOpenCVMatrix(path, name, *rows, *cols, d)) return 1;
    for (int k = 0; k < cap && k < (int)d.size(); ++k) out[k] = d[k];
    return 0;
}

long long hulo_host_read_feat(const char *path, double *xy, unsigned long long cap) {
    FeatureLocations f;
    if (!readFeatFile(path, f)) return -1;
    for (unsigned long long k = 0; k < f.size() && k < cap; ++k) { xy[2 * k] = f[k].first; xy[2 * k + 1] = f[k].second; }
    return (long long)f.size();
}

// .bow / readMatBin round trip: values as doubles, returns rows * cols or -1
long long hulo_host_read_mat_bin(const char *path, int *rows, int *cols, double *values, unsigned long long cap) {
    std::vector<double> v;
    if (!readMatBin(path, *rows, *cols, v)) return -1;
    for (unsigned long long k = 0; k < v.size() && k < cap; ++k) values[k] = v[k];
    return (long long)v.size();
}
int hulo_host_save_mat_bin(const char *path, int rows, int cols, int cv_type, const double *values) {
    return saveMatBin(path, rows, cols, cv_type, values) ? 0 : 1;
}

// hulo::selectViewByBoF over views 0..n_views-1 named frame%04d.jpg in match_dir (needs a GPU):
// returns the number of selected views written to out (ascending, the std::set order) or -1
long long hulo_host_select_view_by_bof(const char *match_dir, unsigned long long n_views, const float *bow,
                                       unsigned long long dim, const unsigned long long *view_list,
                                       unsigned long long n_list, int knn, unsigned long long *out) {
    try {
        Views views;
        char name[64];
        for (unsigned long long v = 0; v < n_views; ++v) {
            snprintf(name, sizeof name, "frame%04llu.jpg", v);
            views[(std::size_t)v] = View{(std::size_t)v, name};
        }
        std::set<std::size_t> list(view_list, view_list + n_list), sel;
        selectViewByBoF(defaultSession(), std::vector<float>(bow, bow + dim), match_dir, list, views, knn, sel);
        unsigned long long k = 0;
        for (std::size_t v : sel) out[k++] = v;
        return (long long)k;
    } catch (const std::exception &) {
        return -1;
    }
}

}  // extern "C"
