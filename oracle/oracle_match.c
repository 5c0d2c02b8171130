/*
 * oracle_match.c -- CPU restatement of the matching half of the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sfmlocalization_b200/ may include, link or
 * call this file.  It is used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker and the timed CPU baseline.
 *
 * What it restates (paths relative to /root/reference):
 *   - the k=2 nearest-neighbour search of VisionLocalizeCommon/src/MatchUtils.cpp:105-108,
 *     191-194, 339-340.  The reference calls OpenCV 3.0 cv::flann::Index (LSH 2/20/2,
 *     checks=2), a third-party dependency that is not vendored; per BASELINE.json the
 *     target is the EXACT 2-NN, so this is the exact search with the result order both
 *     cv::BFMatcher and FLANN's KNNUniqueResultSet produce: (distance asc, index asc).
 *     Missing neighbours are idx -1 / dist INT_MAX (the FLANN convention guarded at
 *     MatchUtils.cpp:115, 203, 349).
 *   - the ratio test of MatchUtils.cpp:113-116, 202-205, 347-349
 *   - the post filters of matchAKAZE (MatchUtils.cpp:111-150), trackAKAZE (:200-276)
 *     and matchAKAZEToQuery (:346-355)
 *   - the 2D-3D assembly of SfMDataUtils.cpp:59-125
 *   - the 61 -> 64 byte row padding of FileUtils.cpp:77-103
 *
 * Parity pin: the reference ships no golden vectors for this path.
 * The knn2 restatement is pinned against OpenCV's exact matchers (cv2.BFMatcher and
 * cv2.flann_Index LINEAR) through tests/golden/ -- see tests/golden/make_golden.py.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_NONE (-1)

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Thread count of the parallel loops from now on, whatever OMP_NUM_THREADS said at start-up
 * (launchers such as torchrun export OMP_NUM_THREADS=1 to every rank). */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* Hamming distance of two rows of `len` bytes (len <= 64). */
static inline int row_hamming(const uint8_t *a, const uint8_t *b, size_t len) {
    int d = 0;
    size_t k = 0;
    for (; k + 8 <= len; k += 8) {
        uint64_t x, y;
        memcpy(&x, a + k, 8);
        memcpy(&y, b + k, 8);
        d += __builtin_popcountll(x ^ y);
    }
    for (; k < len; ++k) d += __builtin_popcount((unsigned)(a[k] ^ b[k]));
    return d;
}

/* FileUtils.cpp:77-92 : rows narrower than 64 bytes are zero padded to 64. */
void orc_pad_rows(const uint8_t *src, size_t n, size_t stride, uint8_t *dst64) {
    size_t w = stride < 64 ? stride : 64;
    for (size_t i = 0; i < n; ++i) {
        memcpy(dst64 + 64 * i, src + stride * i, w);
        if (w < 64) memset(dst64 + 64 * i + w, 0, 64 - w);
    }
}

/*
 * Exact 2-NN of every row of A among the rows of B.
 * idx2/dist2 are nA x 2 int32, row-major: [i0, i1], [d0, d1].
 * Order (distance asc, index asc); fewer than two rows in B leaves -1 / INT_MAX.
 * Only the first min(strideA, strideB, 64) bytes of a row are compared; callers pass
 * equally wide rows.
 */
void orc_knn2_hamming(const uint8_t *A, size_t nA, size_t strideA, const uint8_t *B, size_t nB,
                      size_t strideB, int32_t *idx2, int32_t *dist2) {
    size_t len = strideA < strideB ? strideA : strideB;
    if (len > 64) len = 64;
    /* Eight searcher rows share one pass over B so that each database row is fetched once per
     * group (the scan is otherwise memory bound on large tables).  Per row the candidates are
     * still visited in ascending j, so the tie order is unchanged. */
    enum { TA = 8 };
    const long long n_groups = (long long)((nA + TA - 1) / TA);
#pragma omp parallel for schedule(dynamic, 1)
    for (long long g = 0; g < n_groups; ++g) {
        const size_t a0 = (size_t)g * TA;
        const size_t na = nA - a0 < TA ? nA - a0 : TA;
        uint64_t q[TA][8];
        int d0[TA], d1[TA];
        int32_t i0[TA], i1[TA];
        memset(q, 0, sizeof q);
        for (size_t r = 0; r < TA; ++r) {
            if (r < na) memcpy(q[r], A + strideA * (a0 + r), len);
            d0[r] = d1[r] = INT_MAX;
            i0[r] = i1[r] = ORC_NONE;
        }
        for (size_t j = 0; j < nB; ++j) {
            uint64_t b[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            memcpy(b, B + strideB * j, len);
            for (size_t r = 0; r < TA; ++r) {
                int d = 0;
                for (int k = 0; k < 8; ++k) d += __builtin_popcountll(q[r][k] ^ b[k]);
                /* j ascends, so a tie never displaces an earlier (lower) index */
                if (d < d1[r]) {
                    if (d < d0[r]) {
                        d1[r] = d0[r]; i1[r] = i0[r];
                        d0[r] = d;     i0[r] = (int32_t)j;
                    } else {
                        d1[r] = d;     i1[r] = (int32_t)j;
                    }
                }
            }
        }
        for (size_t r = 0; r < na; ++r) {
            idx2[2 * (a0 + r)] = i0[r];  idx2[2 * (a0 + r) + 1] = i1[r];
            dist2[2 * (a0 + r)] = d0[r]; dist2[2 * (a0 + r) + 1] = d1[r];
        }
    }
}

/*
 * Ratio test, MatchUtils.cpp:347-349 (and :113-116, :202-205):
 *   (0.0f + d0) / d1 < ratio   evaluated in float32, then   d1 < INT_MAX.
 * 0/0 is NaN and compares false.
 */
int orc_ratio_pass(int32_t d0, int32_t d1, float ratio) {
    volatile float q = (0.0f + (float)d0) / (float)d1;
    return (q < ratio) && (d1 < INT_MAX);
}

/*
 * matchAKAZEToQuery for one view, MatchUtils.cpp:339-355.
 * A = the view's rows, B = the query image's rows.  Emits (i, j, d0) for every row i
 * ascending that passes the ratio test; returns the count.  out_* hold up to nA entries.
 */
size_t orc_match_view_to_query(const uint8_t *A, size_t nA, size_t strideA, const uint8_t *B,
                               size_t nB, size_t strideB, float ratio, int32_t *out_i,
                               int32_t *out_j, int32_t *out_d0) {
    if (nA == 0) return 0;
    int32_t *idx2 = (int32_t *)malloc(sizeof(int32_t) * 2 * nA);
    int32_t *dist2 = (int32_t *)malloc(sizeof(int32_t) * 2 * nA);
    orc_knn2_hamming(A, nA, strideA, B, nB, strideB, idx2, dist2);
    size_t n = 0;
    for (size_t i = 0; i < nA; ++i) {
        if (orc_ratio_pass(dist2[2 * i], dist2[2 * i + 1], ratio)) {
            out_i[n] = (int32_t)i;
            out_j[n] = idx2[2 * i];
            out_d0[n] = dist2[2 * i];
            ++n;
        }
    }
    free(idx2);
    free(dist2);
    return n;
}

/*
 * Post filter of matchAKAZE / trackAKAZE given the raw 2-NN of one pair,
 * MatchUtils.cpp:111-150 (= :200-236).  Quirks kept on purpose:
 *   - a row that passes the float ratio but has d1 == INT_MAX keeps the value-initialised
 *     train index 0 (:111-117);
 *   - the one-to-one scan and the emit loop stop at n-2, so the last row of image I is
 *     never emitted (:125, :146), although it can still knock out an earlier claimant.
 * Returns the number of (i, j) pairs written; out_* hold up to nA entries.
 */
size_t orc_pair_filter(const int32_t *idx2, const int32_t *dist2, size_t nA, float ratio,
                       int32_t *out_i, int32_t *out_j) {
    if (nA == 0) return 0;
    int64_t *m = (int64_t *)calloc(nA, sizeof(int64_t));
    for (size_t i = 0; i < nA; ++i) {
        volatile float q = (0.0f + (float)dist2[2 * i]) / (float)dist2[2 * i + 1];
        if (q < ratio) {
            if (dist2[2 * i + 1] < INT_MAX) m[i] = idx2[2 * i];
        } else {
            m[i] = ORC_NONE;
        }
    }
    for (size_t i = 0; i + 1 < nA; ++i) {
        if (m[i] == ORC_NONE) continue;
        int dup = 0;
        for (size_t j = i + 1; j < nA; ++j) {
            if (m[j] == m[i]) {
                m[j] = ORC_NONE;
                dup = 1;
            }
        }
        if (dup) m[i] = ORC_NONE;
    }
    size_t n = 0;
    for (size_t i = 0; i + 1 < nA; ++i) {
        if (m[i] != ORC_NONE) {
            out_i[n] = (int32_t)i;
            out_j[n] = (int32_t)m[i];
            ++n;
        }
    }
    free(m);
    return n;
}

/*
 * matchAKAZE for one pair (I, J), MatchUtils.cpp:94-150: skip when either image has
 * fewer than two rows (:99-101), else knn2 + orc_pair_filter.
 */
size_t orc_match_pair(const uint8_t *A, size_t nA, size_t strideA, const uint8_t *B, size_t nB,
                      size_t strideB, float ratio, int32_t *out_i, int32_t *out_j) {
    if (nA < 2 || nB < 2) return 0;
    int32_t *idx2 = (int32_t *)malloc(sizeof(int32_t) * 2 * nA);
    int32_t *dist2 = (int32_t *)malloc(sizeof(int32_t) * 2 * nA);
    orc_knn2_hamming(A, nA, strideA, B, nB, strideB, idx2, dist2);
    size_t n = orc_pair_filter(idx2, dist2, nA, ratio, out_i, out_j);
    free(idx2);
    free(dist2);
    return n;
}

/*
 * Track propagation of trackAKAZE, MatchUtils.cpp:239-276.
 *   n_frames        V, the number of views
 *   feat_number[f]  rows of frame f, f = 0..V-2 (:183)
 *   m_off / m_i / m_j   consecutive-frame matches: pair (f, f+1) owns entries
 *                   [m_off[f], m_off[f+1]) of m_i/m_j, f = 0..V-2
 * Appends the propagated matches of pairs (f, frameTo), frameTo >= f+2, in the order the
 * reference generates them; out_f/out_t receive the pair, out_i/out_j the match.
 * Returns the number written (never more than `cap`; the total needed is returned in *need).
 */
size_t orc_track_propagate(size_t n_frames, size_t max_frame_dist, const int32_t *feat_number,
                           const int64_t *m_off, const int32_t *m_i, const int32_t *m_j,
                           int32_t *out_f, int32_t *out_t, int32_t *out_i, int32_t *out_j,
                           size_t cap, size_t *need) {
    size_t n = 0, total = 0;
    if (n_frames < 2) { if (need) *need = 0; return 0; }
    size_t nf = n_frames - 1;
    int32_t **tp = (int32_t **)malloc(sizeof(int32_t *) * nf);
    for (size_t f = 0; f < nf; ++f) {
        size_t r = (size_t)feat_number[f];
        tp[f] = (int32_t *)malloc(sizeof(int32_t) * (r ? r : 1));
        for (size_t i = 0; i < r; ++i) tp[f][i] = -1;
        for (int64_t k = m_off[f]; k < m_off[f + 1]; ++k) tp[f][m_i[k]] = m_j[k];
    }
    for (size_t f = 0; f < nf; ++f) {
        size_t lim = f + max_frame_dist < n_frames ? f + max_frame_dist : n_frames;
        for (size_t to = f + 2; to < lim; ++to) {
            for (size_t i = 0; i < (size_t)feat_number[f]; ++i) {
                int32_t t = tp[f][i];
                if (t == -1) continue;
                /* frame to-1 <= V-2 always holds here because to < V */
                int32_t nx = tp[to - 1][t];
                tp[f][i] = nx;
                if (nx != -1) {
                    if (n < cap) {
                        out_f[n] = (int32_t)f; out_t[n] = (int32_t)to;
                        out_i[n] = (int32_t)i; out_j[n] = nx;
                        ++n;
                    }
                    ++total;
                }
            }
        }
    }
    for (size_t f = 0; f < nf; ++f) free(tp[f]);
    free(tp);
    if (need) *need = total;
    return n;
}

/*
 * 2D-3D assembly, SfMDataUtils.cpp:59-125, on flattened inputs.
 * Matches arrive grouped by view in ascending view id (std::map key order) and, inside
 * a view, in emission order.  For match k: view m_view[k], map feature m_i[k], query
 * feature m_j[k].  landmark_of(view, feat) is given as a sorted table of
 * (view, feat) -> landmark id rows: lm_view/lm_feat/lm_id, n_lm entries sorted by
 * (view, feat).  featDist[(v,q)][j] is "last writer wins" over the view's passing rows
 * (MatchUtils.cpp:351), supplied as fd_view/fd_j/fd_d in emission order.
 * Output: ascending query feature j with the landmark whose featDist is smallest;
 * the first candidate in iteration order wins ties (strict > at :109).
 */
static int64_t lm_lookup(const int32_t *lm_view, const int32_t *lm_feat, const int64_t *lm_id,
                         size_t n_lm, int32_t v, int32_t f) {
    size_t lo = 0, hi = n_lm;
    while (lo < hi) {
        size_t mid = (lo + hi) / 2;
        if (lm_view[mid] < v || (lm_view[mid] == v && lm_feat[mid] < f)) lo = mid + 1;
        else hi = mid;
    }
    if (lo < n_lm && lm_view[lo] == v && lm_feat[lo] == f) return lm_id[lo];
    return -1;
}

size_t orc_match_set(const int32_t *m_view, const int32_t *m_i, const int32_t *m_j, size_t n_m,
                     const int32_t *fd_view, const int32_t *fd_j, const int32_t *fd_d, size_t n_fd,
                     const int32_t *lm_view, const int32_t *lm_feat, const int64_t *lm_id,
                     size_t n_lm, size_t n_query_feat, int32_t *out_j, int64_t *out_lm) {
    int64_t *best_lm = (int64_t *)malloc(sizeof(int64_t) * (n_query_feat ? n_query_feat : 1));
    float *best_d = (float *)malloc(sizeof(float) * (n_query_feat ? n_query_feat : 1));
    for (size_t j = 0; j < n_query_feat; ++j) best_lm[j] = -1;
    for (size_t k = 0; k < n_m; ++k) {
        int64_t lm = lm_lookup(lm_view, lm_feat, lm_id, n_lm, m_view[k], m_i[k]);
        if (lm < 0) continue;
        /* featDist[(v,q)][j]: last entry written for (view, j) */
        int found = 0; int32_t d = 0;
        for (size_t t = 0; t < n_fd; ++t) {
            if (fd_view[t] == m_view[k] && fd_j[t] == m_j[k]) { d = fd_d[t]; found = 1; }
        }
        if (!found) continue;
        size_t j = (size_t)m_j[k];
        if (best_lm[j] < 0 || best_d[j] > (float)d) {
            best_lm[j] = lm;
            best_d[j] = (float)d;
        }
    }
    size_t n = 0;
    for (size_t j = 0; j < n_query_feat; ++j) {
        if (best_lm[j] >= 0) { out_j[n] = (int32_t)j; out_lm[n] = best_lm[j]; ++n; }
    }
    free(best_lm);
    free(best_d);
    return n;
}

/* =====================================================================================
 * Guided matching of the F-matrix geometric filter (bGuided_matching = true; the
 * reconstruction drivers pass -gm by default, ReconstructParam.py:70-71,
 * reconstructGraph.py:156-163).  CPU restatement of OpenMVG 1.1
 *   robust_estimation/guided_matching.hpp  (GuidedMatching with Regions, distanceRatio)
 *   matching_image_collection/F_ACRobust.hpp (Geometry_guided_matching: errorTh =
 *        Square(precision_robust), distRatio = Square(0.6))
 * as called by ImageCollectionGeometricFilter::Robust_model_estimation behind
 * hulo::geometricMatch (MatchUtils.cpp:410-413).  PARITY UNPINNED (OpenMVG not vendored,
 * no reference test).  For every feature i of image I: among the features j of image J
 * whose squared distance to the epipolar line F x_i is < error_th, the nearest and second
 * nearest in Hamming distance; kept iff a second one exists and best < dist_ratio * second
 * (doubles).  The nearest keeps the lowest j on ties (strict < in distanceRatio::update).
 * Known deviation: upstream then removes matches whose position 4-tuple repeats
 * (IndMatchDecorator::getDeduplicated) and emits them in the order of its position-keyed
 * set; orc_guided_dedup keeps the first of each 4-tuple in ascending i and keeps that order.
 * ===================================================================================== */
size_t orc_guided_match(const double *F, const double *xI, const uint8_t *descI, size_t nI, size_t strideI,
                        const double *xJ, const uint8_t *descJ, size_t nJ, size_t strideJ, double error_th,
                        double dist_ratio, int32_t *out_i, int32_t *out_j) {
    int32_t *best = (int32_t *)malloc(sizeof(int32_t) * (nI ? nI : 1));
    const size_t lenI = strideI < 64 ? strideI : 64, lenJ = strideJ < 64 ? strideJ : 64;
    const size_t len = lenI < lenJ ? lenI : lenJ;
#pragma omp parallel for schedule(dynamic, 16)
    for (long long i = 0; i < (long long)nI; ++i) {
        /* explicit fma() in a fixed order: the device evaluates the same expression tree, so the
         * strict comparison with error_th decides identically on both sides */
        const double a = xI[2 * i], b = xI[2 * i + 1];
        const double l0 = fma(F[0], a, fma(F[1], b, F[2])), l1 = fma(F[3], a, fma(F[4], b, F[5])),
                     l2 = fma(F[6], a, fma(F[7], b, F[8]));
        const double den = fma(l0, l0, l1 * l1);
        double bd = 1e300, sbd = 1e300;
        int have2 = 0, have1 = 0;
        int32_t idx = -1;
        for (size_t j = 0; j < nJ; ++j) {
            const double d = fma(l0, xJ[2 * j], fma(l1, xJ[2 * j + 1], l2));
            const double err = d * d / den;
            if (!(err < error_th)) continue;
            int h = row_hamming(descI + (size_t)i * strideI, descJ + j * strideJ, len);
            /* bytes beyond the shorter row count as zero on that side */
            for (size_t k = len; k < lenI; ++k) h += __builtin_popcount(descI[(size_t)i * strideI + k]);
            for (size_t k = len; k < lenJ; ++k) h += __builtin_popcount(descJ[j * strideJ + k]);
            const double dist = (double)h;
            if (dist < bd) { sbd = bd; have2 = have1; bd = dist; idx = (int32_t)j; have1 = 1; }
            else if (dist < sbd) { sbd = dist; have2 = 1; }
        }
        best[i] = (have2 && bd < dist_ratio * sbd) ? idx : -1;
    }
    size_t n = 0;
    for (size_t i = 0; i < nI; ++i)
        if (best[i] >= 0) { out_i[n] = (int32_t)i; out_j[n] = best[i]; ++n; }
    free(best);
    return n;
}

/* Keep the first match (ascending position in the list) of every (xI_i, xJ_j) position
 * 4-tuple compared as floats, order preserved.  Returns the new count (in place). */
size_t orc_guided_dedup(const double *xI, const double *xJ, int32_t *mi, int32_t *mj, size_t n) {
    size_t out = 0;
    for (size_t k = 0; k < n; ++k) {
        const float a = (float)xI[2 * mi[k]], b = (float)xI[2 * mi[k] + 1], c = (float)xJ[2 * mj[k]], d = (float)xJ[2 * mj[k] + 1];
        int dup = 0;
        for (size_t m = 0; m < out && !dup; ++m)
            dup = a == (float)xI[2 * mi[m]] && b == (float)xI[2 * mi[m] + 1] && c == (float)xJ[2 * mj[m]] &&
                  d == (float)xJ[2 * mj[m] + 1];
        if (!dup) { mi[out] = mi[k]; mj[out] = mj[k]; ++out; }
    }
    return out;
}
