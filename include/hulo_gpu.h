/*
 * hulo_gpu.h -- C-ABI of libhulo_gpu.so: the B200 (sm_100a) implementation of the
 * SfMLocalization hot path (exact Hamming 2-NN matching of 64-byte AKAZE/MLDB rows with
 * the ratio test, and scoring of resection-RANSAC hypotheses).
 *
 * Plain C types only; every function returns an int status (HULO_OK == 0) and never
 * throws.  hulo_last_error() returns the message of the last failure on the calling
 * thread.  A hulo_gpu handle owns one device, one stream, its scratch buffers and
 * (optionally) one NCCL communicator; it must be used from one thread at a time, which
 * is how the reference's callers behave (VisionLocalizeServer/src/localizeImage.cc:71-74).
 *
 * There is no CPU fallback: without a CUDA device hulo_gpu_create fails with
 * HULO_ERR_CUDA and nothing else can be called.
 *
 * Each entry point names the reference interface it stands behind (paths relative to
 * the reference repository root).
 */
#ifndef HULO_GPU_H
#define HULO_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HULO_OK 0
#define HULO_ERR_ARG 1      /* bad argument (null pointer, stride, size, capacity) */
#define HULO_ERR_CUDA 2     /* CUDA runtime failure, or no device */
#define HULO_ERR_NCCL 3     /* NCCL failure, or communicator not initialised */
#define HULO_ERR_CAPACITY 4 /* caller buffer too small; *n_out holds the size needed */

#define HULO_ROW_BYTES 64         /* device row width, FileUtils.cpp:77-103 */
#define HULO_DIST_NONE 2147483647 /* INT_MAX: "neighbour not found", MatchUtils.cpp:115 */
#define HULO_IDX_NONE (-1)

typedef struct hulo_gpu hulo_gpu; /* device context */
typedef struct hulo_db hulo_db;   /* device-resident descriptor rows + segment table */

/* ------------------------------------------------------------------ context */

/* Number of CUDA devices visible (0 when there is none); never fails. */
int hulo_device_count(void);

/* Create a context on `device`.  New state with no reference counterpart: the reference
 * keeps no descriptor database in memory (MatchUtils.cpp:328-332 re-reads every view's
 * .desc per query); LocalizeEngine's constructor (LocalizeEngine.cc:84-198) is where a
 * drop-in creates this handle. */
int hulo_gpu_create(int device, hulo_gpu **out);
void hulo_gpu_destroy(hulo_gpu *h);
const char *hulo_last_error(void);
const char *hulo_version(void);

/* Pinned host memory for the buffers of the *_host entry points (optional, faster). */
int hulo_host_alloc(size_t bytes, void **out);
void hulo_host_free(void *p);

/* Device timing of the calls issued since hulo_timer_start on the context's stream
 * (CUDA events); hulo_timer_stop synchronises and returns milliseconds.  This fills the
 * `putMatch` / `PnP` slots of the reference's times[] (LocalizeEngine.cc:643-658). */
int hulo_timer_start(hulo_gpu *h);
int hulo_timer_stop(hulo_gpu *h, float *ms);
int hulo_synchronize(hulo_gpu *h);

/* Number of kernels this library launched on the context since creation. */
uint64_t hulo_launch_count(const hulo_gpu *h);

/* Arithmetic of the flat 2-NN searches (hulo_knn2, hulo_knn2_host, hulo_knn2_sharded).  Both
 * engines give the same exact result, bit for bit, as the cv::flann::Index::knnSearch(k=2)
 * replacement they stand for (MatchUtils.cpp:105-108, 339-340):
 *   HULO_KNN_INT   XOR + popcount on the integer pipes (K1);
 *   HULO_KNN_TC    the 512 bits of a row as +-1 values, distance = (512 - dot) / 2 from a contraction on
 *                  the tensor cores: 4-bit operands with fp32 accumulation (K1t4, tcgen05.mma.kind::mxf4;
 *                  the accumulated integers stay below 2^10, so the result is exact) or, with
 *                  HULO_TC_BITS=8, int8 operands with int32 accumulation (K1t, kind::i8).  Costs one
 *                  expanded copy (256 or 512 bytes per row) of every table it searches;
 *   HULO_KNN_AUTO  (default) K1t when the search is large enough to pay for the expanded copies
 *                  (at least 128 searcher rows, 8192 database rows and 2^28 distances), else K1.
 * The environment variable HULO_KNN_ENGINE=int|tc|auto overrides the default at hulo_gpu_create. */
#define HULO_KNN_INT 0
#define HULO_KNN_TC 1
#define HULO_KNN_AUTO 2
#define HULO_KNN_TC8 3 /* HULO_KNN_TC with int8 operands (K1t): kept for cross-checking the 4-bit form */
int hulo_gpu_set_knn_engine(hulo_gpu *h, int engine);
int hulo_gpu_knn_engine(const hulo_gpu *h);

/* ------------------------------------------------------------ descriptor rows */

/* Upload n rows of `stride` bytes (61..64 used; bytes beyond 64 ignored, rows narrower
 * than 64 zero padded -- hulo::saveAKAZEBin, FileUtils.cpp:77-92) as a device-resident
 * table.  seg_offsets (n_seg+1 ascending row offsets, seg_offsets[0]=0,
 * seg_offsets[n_seg]=n) records which rows belong to which view / image; pass NULL,0
 * for a single segment.  Stands where hulo::readAKAZEBin (FileUtils.cpp:94-103) is
 * called per view per query (MatchUtils.cpp:86-95, 328-332). */
int hulo_db_upload(hulo_gpu *h, const uint8_t *rows, size_t n, size_t stride,
                   const uint64_t *seg_offsets, size_t n_seg, hulo_db **out);
/* Replace the rows of an existing table in place (same or smaller n; one segment). */
int hulo_db_update(hulo_gpu *h, hulo_db *db, const uint8_t *rows, size_t n, size_t stride);
void hulo_db_free(hulo_db *db);
size_t hulo_db_rows(const hulo_db *db);
size_t hulo_db_segments(const hulo_db *db);
/* Copy rows [first, first+n) back to the host as 64-byte rows. */
int hulo_db_download(hulo_gpu *h, const hulo_db *db, size_t first, size_t n, uint8_t *rows64);

/* --------------------------------------------------------------- K1: 2-NN */

/* Exact 2 nearest neighbours, under 512-bit Hamming distance, of every row of A among
 * the rows of B, ordered by (distance, index) ascending.
 *   idx2  nA x 2 int32 row-major: indices into B; HULO_IDX_NONE when B has < 2 (< 1) rows
 *   dist2 nA x 2 int32 row-major: distances; HULO_DIST_NONE for a missing neighbour
 * Replaces cv::flann::Index(...).knnSearch(desc1, idx, dist, 2, ...) at
 * MatchUtils.cpp:105-108, 191-194, 308-310 + 339-340 (same output layout and the same
 * "-1 / INT_MAX" convention), with exact search instead of LSH (BASELINE.json).
 * Outputs are HOST pointers; pass NULL for both to leave the result on the device
 * (used by the throughput benchmark; fetch later with hulo_knn2_fetch). */
int hulo_knn2(hulo_gpu *h, const hulo_db *A, const hulo_db *B, int32_t *idx2, int32_t *dist2);
int hulo_knn2_fetch(hulo_gpu *h, size_t nA, int32_t *idx2, int32_t *dist2);
/* Same with host rows: uploads A and B, searches, downloads.  stride as hulo_db_upload. */
int hulo_knn2_host(hulo_gpu *h, const uint8_t *A, size_t nA, size_t strideA, const uint8_t *B,
                   size_t nB, size_t strideB, int32_t *idx2, int32_t *dist2);

/* ------------------------------------- query localisation (reference direction) */

/* hulo::matchAKAZEToQuery, MatchUtils.cpp:283-367 (decl MatchUtils.h:54-61), for the
 * selected views `views[0..n_views)` (segment numbers of `map`): for every row i of each
 * view, 2-NN among the query image's rows; keep when
 * (0.0f + d0) / d1 < ratio (float32) and d1 < INT_MAX (:347-349).
 * Output, in the reference's iteration order (views in the order given, i ascending):
 *   out_view[k]  position in views[] of the match's view
 *   out_i[k]     feature index inside the view   (IndMatch::i_)
 *   out_j[k]     query feature index             (IndMatch::j_)
 *   out_d0[k]    distance to the nearest query row (featDist, :351)
 * view_counts[v] (n_views entries, may be NULL) receives the matches per view.
 * cap is the capacity of the out_* arrays; *n_out the number of matches.  If cap is too
 * small the call returns HULO_ERR_CAPACITY with *n_out = needed.
 * views == NULL selects every segment of `map` in order. */
int hulo_match_to_query(hulo_gpu *h, const hulo_db *map, const uint32_t *views, size_t n_views,
                        const uint8_t *query, size_t nq, size_t q_stride, float ratio,
                        uint32_t *out_view, uint32_t *out_i, uint32_t *out_j, int32_t *out_d0,
                        size_t cap, size_t *n_out, uint32_t *view_counts);

/* Batched hulo::matchAKAZEToQuery: n_queries query images against the same map in one pass
 * (the server handling concurrent requests, BASELINE.json config 4; the reference has no
 * batched form -- its engine is single-threaded and non-re-entrant, localizeImage.cc:71-74).
 * queries holds the descriptor rows of all images back to back, image q owning rows
 * [q_offsets[q], q_offsets[q+1]).  Output order: query ascending, then as hulo_match_to_query;
 * out_query[k] is the query image of match k; counts (n_queries x n_views, may be NULL) the
 * matches per (query, view).  Capacity protocol as hulo_match_to_query. */
int hulo_match_to_queries(hulo_gpu *h, const hulo_db *map, const uint32_t *views, size_t n_views,
                          const uint8_t *queries, size_t q_stride, const uint64_t *q_offsets, size_t n_queries,
                          float ratio, uint32_t *out_query, uint32_t *out_view, uint32_t *out_i, uint32_t *out_j,
                          int32_t *out_d0, size_t cap, size_t *n_out, uint32_t *counts);

/* ------------------------------------------ image-pair matching (reconstruction) */

#define HULO_PAIR_ONE_TO_ONE 1u /* drop every claimant of a train row claimed twice, :125-143 */
#define HULO_PAIR_DROP_LAST 2u  /* reference quirk: the last row of image I is never emitted, :125,:146 */
#define HULO_PAIR_REFERENCE (HULO_PAIR_ONE_TO_ONE | HULO_PAIR_DROP_LAST)

/* hulo::matchAKAZE, MatchUtils.cpp:73-152 (decl MatchUtils.h:39-42), and the matching half
 * of hulo::trackAKAZE, :164-237: for each pair p = (I, J) = (pairs[2p], pairs[2p+1]) of
 * segments of `db`: skip when either has < 2 rows (:99-101); 2-NN of every row of I among
 * the rows of J; ratio test (:112-121); optional one-to-one filter and last-row quirk.
 *   pair_offsets  n_pairs+1 entries: matches of pair p are [pair_offsets[p], pair_offsets[p+1])
 *   out_i/out_j   IndMatch(i, j), i ascending inside a pair
 * Capacity protocol as hulo_match_to_query. */
int hulo_match_pairs(hulo_gpu *h, const hulo_db *db, const uint32_t *pairs, size_t n_pairs,
                     float ratio, unsigned flags, uint64_t *pair_offsets, uint32_t *out_i,
                     uint32_t *out_j, size_t cap, size_t *n_out);

/* ------------------------------------------------------- K2: resection scoring */

/* Score H pose hypotheses against N 2D-3D correspondences: the body of the model loop of
 * openMVG::robust::ACRANSAC as run by SfM_Localizer::Localize (called at
 * LocalizeEngine.cc:531, localization.cpp:508, adjust_sfm_data.cpp:135-137):
 * squared reprojection residual in K^-1-normalised coordinates for every correspondence,
 * ascending sort, a-contrario NFA scan.
 *   models  H x 12 doubles, row-major 3x4 [R|t] (normalised camera)
 *   x2d     N x 2 doubles, pixel coordinates;  X3d  N x 3 doubles;  K 3x3 row-major
 *   thr_px  >= 0: also count correspondences with residual <= thr_px pixels into n_inl
 * Outputs (each H entries, any may be NULL): nfa (log10 NFA), k_best (inlier count at the
 * NFA minimum), err_k (residual in pixels of the k_best-th correspondence), n_inl. */
int hulo_score_resection(hulo_gpu *h, const double *models, size_t H, const double *x2d,
                         const double *X3d, size_t N, const double *K, double thr_px, float *nfa,
                         int32_t *k_best, float *err_k, int32_t *n_inl);

/* Residuals only (pixels, fp32 on the device), H x N row-major: parity check of K2. */
int hulo_resection_residuals(hulo_gpu *h, const double *models, size_t H, const double *x2d,
                             const double *X3d, size_t N, const double *K, float *res_px);

/* P3P minimal solver on the device for T sample triplets (indices into the
 * correspondences): up to 4 models each (openMVG::euclidean_resection::P3PSolver).
 * models: T x 4 x 12 doubles; n_models: T entries. */
int hulo_p3p(hulo_gpu *h, const uint32_t *triplets, size_t T, const double *x2d, const double *X3d,
             size_t N, const double *K, double *models, int32_t *n_models);

/* openMVG::sfm::SfM_Localizer::Localize (LocalizeEngine.cc:529-532): AC-RANSAC resection
 * with P3P, max_iter iterations (the reference runs OpenMVG's default 4096), batched:
 * triplets are drawn from all correspondences and scored in growing batches (64, 128, ... 512)
 * until a meaningful model (NFA < 0) appears or 90 % of the budget is spent; the reserved 10 %
 * are then drawn from the inliers of the best model so far, like the sequential schedule.
 * Outputs: P = K [R|t] (12 doubles), inliers (capacity N) sorted by residual, *n_inliers,
 * *error_max in pixels.  *found is 1 iff inliers > 2.5 * 3 with NFA < 0. */
int hulo_resect_acransac(hulo_gpu *h, const double *x2d, const double *X3d, size_t N,
                         const double *K, size_t max_iter, uint64_t seed, double *P,
                         int32_t *inliers, size_t *n_inliers, double *error_max, int *found);

/* Host-only self-test (no device needed) of the fp64 rescoring behind hulo_resect_acransac: the
 * production form (radix sort of the residuals' upper key halves with repair of equal runs, NFA
 * minimum found through a float bracket and evaluated exactly only where the bracket allows) against
 * a comparison sort and the full log10 scan, on n_cases seeded sets of n_points correspondences with
 * inliers, outliers, exact duplicates and points on the principal plane.  *n_mismatch = number of
 * sets on which order, k, k-th residual or NFA differ in any bit (expected: 0). */
int hulo_selftest_rescoring(uint64_t seed, size_t n_points, size_t n_cases, size_t *n_mismatch);

/* The same with the SEQUENTIAL schedule of openMVG::robust::ACRANSAC kept to the letter: the
 * global phase ends at the first meaningful model in draw order, the pool narrows again at every
 * improvement during the reserved iterations, and a model replaces the best so far only in draw
 * order.  Iterations are still scored in batches, speculatively; results are committed in order
 * and the rest of a batch is re-drawn when a commit changes the pool.  For one seed this runs the
 * trace of the sequential CPU restatement (same draws, same pool updates); it costs more round
 * trips than hulo_resect_acransac (1.2 ms against 0.28 ms for 700 clean correspondences), which is why
 * the engine uses the batched form.  Inliers and error_max are at the device's scoring precision
 * (fp32 residual keys).  Arguments and outputs as hulo_resect_acransac. */
int hulo_resect_acransac_sequential(hulo_gpu *h, const double *x2d, const double *X3d, size_t N,
                                    const double *K, size_t max_iter, uint64_t seed, double *P,
                                    int32_t *inliers, size_t *n_inliers, double *error_max, int *found);

/* Many independent resections in one call: all views of a reconstruction re-resected against its
 * structure (OpenMVG_BA/src/adjust_sfm_data.cpp:91-146 -- an omp loop around
 * SfM_Localizer::Localize at :135-137), or the queries of a server batch.  Problem p owns the
 * correspondences [offsets[p], offsets[p+1]) of x2d (total x 2) / X3d (total x 3), the intrinsics
 * K + 9 p and the seed seeds[p] (seeds == NULL: seed + 1000003 p).  Each problem runs exactly the
 * schedule of hulo_resect_acransac -- the results are bit-identical to n_problems single calls --
 * but one step of all problems still running is one wave on the device (one P3P launch, one
 * scoring launch per size class, one first-minimum launch, one copy back).  Outputs: P (n x 12),
 * inliers of problem p at inliers + offsets[p] (capacity: its N), n_inliers, error_max, found (n). */
int hulo_resect_acransac_batch(hulo_gpu *h, size_t n_problems, const uint64_t *offsets,
                               const double *x2d, const double *X3d, const double *K, size_t max_iter,
                               uint64_t seed, const uint64_t *seeds, double *P, int32_t *inliers,
                               uint64_t *n_inliers, double *error_max, int32_t *found);

/* Pose extraction from a projection matrix (LocalizeEngine.cc:582-602, localization.cpp:544-547,
 * adjust_sfm_data.cpp:138-142): KRt_From_P, then centre = -R^T t.  Host arithmetic, no launch.
 * P: 12 doubles row-major; outputs K (9, may be NULL), R (9, row-major), center (3). */
int hulo_pose_from_projection(const double *P, double *K, double *R, double *center);

/* --------------------------------------------- K3: F-matrix geometric filter */

/* hulo::geometricMatch, MatchUtils.cpp:372-420 (decl MatchUtils.h:66-72): OpenMVG's
 * GeometricFilter_FMatrix_AC(geomPrec, ransacRound) on every pair of the putative matches
 * (LocalizeEngine.cc:458, localization.cpp:450, computeFeaturesAndMatches.cpp:242), without
 * guided matching: a-contrario RANSAC over the 7-point fundamental-matrix solver, residual =
 * squared distance to the epipolar line in image J, both images preconditioned by their size.
 * All pairs run in one launch (one thread block per pair).
 *   xI, xJ        total x 2 doubles: pixel positions of the matched features in image I / J,
 *                 pairs back to back; pair p owns matches [pair_offsets[p], pair_offsets[p+1])
 *   image_sizes   n_pairs x {wI, hI, wJ, hJ}
 *   precision_px  geomPrec: upper bound of the inlier residual in pixels (INFINITY = none)
 *   max_iter      ransacRound (25 in LocalizeParam.py:35, 200 CLI default, 500 ExtFeatAndMatch)
 *   seed          pair p samples from the stream seeded with seed + 1000003 p, or with
 *                 pair_seeds[p] when pair_seeds is not NULL (results then do not depend on
 *                 how a pair list is split into calls)
 * Outputs per pair: valid[p] = 1 iff a meaningful model (NFA < 0) with more than 2.5 * 7
 * inliers exists (the pair keeps its key in map_geometricMatches); n_inliers[p]; the inlier
 * positions inside the pair, in (residual, index) order like ACRANSAC's vec_inliers, at
 * inliers[pair_offsets[p] ...] (capacity pair_offsets[n_pairs]); F (n_pairs x 9, row-major,
 * pixel coordinates, x_J^T F x_I = 0), error_max (pixels) and nfa (log10) may be NULL.
 * At most 16384 putative matches per pair. */
int hulo_geometric_filter(hulo_gpu *h, const double *xI, const double *xJ, const uint64_t *pair_offsets,
                          size_t n_pairs, const int32_t *image_sizes, double precision_px, size_t max_iter,
                          uint64_t seed, const uint64_t *pair_seeds, int32_t *valid, uint32_t *n_inliers,
                          int32_t *inliers, double *F, double *error_max, double *nfa);

/* Guided matching of the geometric filter (bGuided_matching = true of hulo::geometricMatch,
 * MatchUtils.cpp:372-420; -gm of the CLIs, on by default in the reconstruction drivers,
 * ReconstructParam.py:70-71): OpenMVG's Geometry_guided_matching for the pairs that passed the
 * robust estimation.  For pair p = (I, J) = (pairs[2p], pairs[2p+1]) of segments of `db`, with
 * F[p] (row-major, pixel coordinates, x_J^T F x_I = 0, as hulo_geometric_filter returns it) and
 * error_th[p] = (robust precision in pixels)^2: for every feature i of I the nearest and second
 * nearest descriptor among the features j of J whose squared distance to the epipolar line F x_i
 * is < error_th[p]; the match (i, j) is kept iff a second one exists and best < dist_ratio * second
 * (OpenMVG passes 0.6^2).  xy holds the (undistorted) feature position of every row of `db`,
 * 2 doubles per row.  dedup != 0 drops, per pair, matches whose position 4-tuple repeats as floats
 * (first kept).  Output as hulo_match_pairs: ascending i inside a pair, pair_offsets (n_pairs + 1),
 * capacity protocol of hulo_match_to_query. */
int hulo_guided_match(hulo_gpu *h, const hulo_db *db, const double *xy, const uint32_t *pairs, size_t n_pairs,
                      const double *F, const double *error_th, double dist_ratio, int dedup,
                      uint64_t *pair_offsets, uint32_t *out_i, uint32_t *out_j, size_t cap, size_t *n_out);

/* ------------------------------------------------- query localisation, end to end */

typedef struct hulo_engine hulo_engine;

/* The device-resident state LocalizeEngine builds once per map (LocalizeEngine.cc:84-198):
 * every view's descriptor rows (segments = views, in ascending view id), the
 * (view, feature) -> landmark table of hulo::structureToMapViewFeatTo3D
 * (SfMDataUtils.cpp:33-46) and the landmark positions.
 *   obs_view/obs_feat/obs_landmark  n_obs observations; obs_landmark indexes landmark_X
 *   landmark_X                      n_landmarks x 3 doubles
 *   K                               3x3 row-major pinhole intrinsics of the query camera
 * The engine keeps a reference to `h`, which must outlive it. */
int hulo_engine_create(hulo_gpu *h, const uint8_t *rows, size_t n, size_t stride, const uint64_t *seg_offsets,
                       size_t n_views, const uint32_t *obs_view, const uint32_t *obs_feat,
                       const uint32_t *obs_landmark, size_t n_obs, const double *landmark_X, size_t n_landmarks,
                       const double *K, hulo_engine **out);
void hulo_engine_destroy(hulo_engine *e);

/* Acceptance thresholds of the reference (LocalizeEngine.cc:63-65): a view is kept with at
 * least min_putative putative matches (16, :428-434), resection is tried with more than
 * min_points correspondences (8, :529) and accepted with more than min_inliers inliers (10, :560).
 * max_iter is the AC-RANSAC budget (4096 = OpenMVG's default, which the reference never overrides). */
int hulo_engine_configure(hulo_engine *e, float ratio, int min_putative, int min_points, int min_inliers,
                          size_t max_iter);

/* Schedule of the AC-RANSAC resection behind hulo_engine_localize (SfM_Localizer::Localize,
 * LocalizeEngine.cc:529-532):
 *   HULO_RESECT_BATCHED     (default) growing batches of iterations scored together, the model with the
 *                           smallest NFA of a batch committed; statistically equivalent, fewest round trips;
 *   HULO_RESECT_SEQUENTIAL  the reference's own schedule kept to the letter: every iteration is committed
 *                           in order with the strict-improvement rule, the sampling pool narrows at the
 *                           iteration it would narrow in OpenMVG's loop (hulo_resect_acransac_sequential).
 * The batched localisation (hulo_engine_localize_batch) always uses the batched schedule. */
#define HULO_RESECT_BATCHED 0
#define HULO_RESECT_SEQUENTIAL 1
int hulo_engine_set_resection_schedule(hulo_engine *e, int schedule);

/* Keypoint positions for the geometric filter (the Regions_Provider of LocalizeEngine.cc:458):
 * map_xy holds 2 doubles per descriptor row of the map (same row order), view_wh the image
 * size {w, h} of every view (View::ui_width / ui_height), query_w / query_h that of the
 * query camera. */
int hulo_engine_set_keypoints(hulo_engine *e, const double *map_xy, const int32_t *view_wh, int query_w,
                              int query_h);
/* The size of the next query image (the View added per query at LocalizeEngine.cc:405-409). */
int hulo_engine_set_query_size(hulo_engine *e, int query_w, int query_h);
/* Switch hulo::geometricMatch (LocalizeEngine.cc:458) on or off (off after hulo_engine_create):
 * ransac_round = mRansacRound, precision_px = mRansacPrecision of the LocalizeEngine
 * constructor (LocalizeEngine.cc:84-91).  Switching it off also switches guided matching off. */
int hulo_engine_configure_geometric(hulo_engine *e, int enabled, size_t ransac_round, double precision_px);
/* mGuidedMatching of the LocalizeEngine constructor / -gm of the localisation CLI
 * (bGuided_matching of hulo::geometricMatch, MatchUtils.cpp:407-416): for every (view, query) pair
 * that passed the F-matrix filter, ALL features of the view are re-matched against ALL features of
 * the query behind the epipolar gate of the pair's F (robust precision of the filter, descriptor
 * ratio 0.6^2) and the result replaces the pair's inlier list before the 2D-3D assembly.  Needs
 * the geometric filter enabled (off after hulo_engine_create). */
int hulo_engine_set_guided_matching(hulo_engine *e, int enabled);

/* LocalizeEngine::localize from the putative matching on (LocalizeEngine.cc:423-602) for one
 * query image given its descriptors and (undistorted) keypoint positions:
 *   hulo::matchAKAZEToQuery -> drop views with too few matches -> [hulo::geometricMatch, when
 *   enabled: F-matrix AC-RANSAC per (view, query) pair, hulo_geometric_filter] ->
 *   hulo::matchProviderToMatchSet (2D-3D assembly over the geometric -- else putative --
 *   matches, closest descriptor wins) -> SfM_Localizer::Localize -> KRt_From_P.
 *   views / n_views   selected map views (NULL: all), as the `pairs` argument of matchAKAZEToQuery
 *   pose12            camera centre -R^T t (3) then R row-major (9), as LocalizeEngine.cc:593-602
 *   *localized        1 iff resection succeeded with enough inliers
 *   corr_qfeat / corr_landmark (capacity nq, may be NULL) the 2D-3D pairs, ascending query feature
 *   inliers (capacity nq, may be NULL) indices into the pair list
 *   times_ms[4]       putMatch, assembly, PnP, geoMatch (device + host wall time per stage) */
int hulo_engine_localize(hulo_engine *e, const uint8_t *qdesc, size_t nq, size_t q_stride, const double *qxy,
                         const uint32_t *views, size_t n_views, uint64_t seed, double *pose12, int *localized,
                         uint32_t *corr_qfeat, uint32_t *corr_landmark, size_t *n_corr, int32_t *inliers,
                         size_t *n_inliers, double *times_ms);

/* The same for n_queries images at once (concurrent server requests, BASELINE.json config 4):
 * one batched matching pass (hulo_match_to_queries), then assembly and resection per image.
 * Layout of qdesc / q_offsets as hulo_match_to_queries; qxy holds 2 doubles per descriptor row.
 * Image q uses seed + q.  pose12: n_queries x 12; localized, n_corr, n_inliers: n_queries entries
 * (the last two may be NULL); times_ms[4] accumulates putMatch, assembly, PnP, geoMatch over the batch. */
int hulo_engine_localize_batch(hulo_engine *e, size_t n_queries, const uint8_t *qdesc, size_t q_stride,
                               const uint64_t *q_offsets, const double *qxy, const uint32_t *views, size_t n_views,
                               uint64_t seed, double *pose12, int *localized, uint32_t *n_corr, uint32_t *n_inliers,
                               double *times_ms);

/* ------------------------------------------- K4: 3D-3D transform between two models */

/* ransacAffineTransform (PyVisionLocalizeCommon/src/hulo_sfm/mergeSfM.py:344-388; similarity = 0)
 * and ransacSimilarityTransform (hulo_transform/ransacTransform.py:13-49; similarity = 1), which
 * mergeSfM.ransacTransform (:394-399) runs when two models are merged (:577-579) and when a model
 * is anchored to world coordinates (localizeGlobalCoordinate.py:209, measureAccuracy.py:239):
 * the 3 x 4 matrix M with A ~ M [B; 1].  Every round takes 4 points, fits M to them (exact 4 x 4
 * solve, or rotation + scale + translation), counts the points with || M [B_i; 1] - A_i || < thres;
 * a round replaces the best so far iff it has strictly more inliers and the singular values of its
 * linear part satisfy s_max / s_min < svd_ratio; M is refitted on the inliers of the best round.
 * All rounds are scored in one launch.
 *   A, B      3 x n doubles, row-major (the reference's numpy layout: one row per coordinate)
 *   samples   rounds x 4 point indices (what random.sample returned per round), or NULL to draw
 *             them from `seed`
 *   svd_ratio pass INFINITY for "no condition" (the reference's default sys.float_info.max)
 * Outputs: M (12 doubles, row-major 3 x 4; all zero when nothing was found), inliers (capacity n,
 * ascending like np.where), *n_inliers (0 = the reference's `return [], []`), *best_round (may be
 * NULL) the winning round. */
int hulo_ransac_transform3d(hulo_gpu *h, const double *A, const double *B, size_t n, double thres,
                            const uint32_t *samples, size_t rounds, uint64_t seed, double svd_ratio,
                            int similarity, double *M, int32_t *inliers, size_t *n_inliers,
                            uint32_t *best_round);

/* --------------------------------------- K5: nearest map views in bag-of-features space */

typedef struct hulo_bow hulo_bow;

/* hulo::selectViewByBoF (BoWCommon/src/BoFUtils.cpp:27-68; callers LocalizeEngine.cc:296-332,
 * localization.cpp:386-412): the knn views whose BoF vector is nearest to the query's under L2.
 * The reference rebuilds a FLANN KD-tree (4 trees, 64 checks, approximate) over the candidate views
 * for every query; here the n x d matrix (row v = the vector of view v, what readMatBin returns
 * from <view>.bow, transposed) is uploaded once and the search is exact.
 *   subset / n_subset  candidate views (the viewList argument), NULL = all
 *   knn                must be smaller than the number of candidates (CV_Assert at :30)
 *   idx (knn)          view numbers by (squared distance, view number) ascending; dist may be NULL */
int hulo_bow_create(hulo_gpu *h, const float *bof, size_t n, size_t d, hulo_bow **out);
void hulo_bow_destroy(hulo_bow *b);
int hulo_bow_knn(hulo_bow *b, const float *query, const uint32_t *subset, size_t n_subset, size_t knn,
                 int32_t *idx, float *dist);

/* ------------------------------------------------------------------ multi GPU */

/* One process per GPU.  Rank 0 obtains an id with hulo_comm_unique_id and distributes its
 * 128 bytes out of band; every rank then calls hulo_comm_init. */
int hulo_comm_unique_id(void *id128);
int hulo_comm_init(hulo_gpu *h, const void *id128, int rank, int world);
int hulo_comm_barrier(hulo_gpu *h);
/* max over ranks of a device-time measurement */
int hulo_comm_max_f64(hulo_gpu *h, double *value);

/* All-gather of one host buffer of `bytes` per rank into recv (world x bytes, rank order): staged
 * through the device and ncclAllGather.  A context without a communicator is a world of one. */
int hulo_comm_allgather(hulo_gpu *h, const void *send, size_t bytes, void *recv);
const char *hulo_comm_exchange_kind(const hulo_gpu *h);
int hulo_comm_rank(const hulo_gpu *h);
int hulo_comm_world(const hulo_gpu *h);

/* Reference-direction query over a view-sharded map (the matcher of LocalizeEngine.cc:423 has one
 * independent unit per view): every rank holds the engine and gets the same query; rank r matches a
 * contiguous range of the view list (hulo_partition_views: balanced by descriptor rows), the
 * surviving matches -- a few thousand records of 12 bytes -- are all-gathered, and every rank
 * finishes the query (filter, assembly, resection) on identical inputs: the result on every rank is
 * bit-identical to hulo_engine_localize on one GPU.  Arguments as hulo_engine_localize; collective
 * over the communicator of the engine's context (hulo_comm_init); a world of one is the plain call. */
int hulo_engine_localize_sharded(hulo_engine *e, const uint8_t *qdesc, size_t nq, size_t q_stride,
                                 const double *qxy, const uint32_t *views, size_t n_views, uint64_t seed,
                                 double *pose12, int *localized, uint32_t *corr_qfeat,
                                 uint32_t *corr_landmark, size_t *n_corr, int32_t *inliers,
                                 size_t *n_inliers, double *times_ms);
/* The contiguous ranges of a view list given to the ranks: bounds[r] .. bounds[r + 1] for rank r
 * (world + 1 entries), each holding about 1 / world of the rows.  Host arithmetic. */
int hulo_partition_views(const uint64_t *rows_per_view, size_t n_views, int world, uint64_t *bounds);

/* Row-sharded database: this rank's B holds rows [row_base, row_base + rows(B)) of the
 * global table.  Every rank passes the same A.  Each rank computes its local top-2 with
 * global indices, the candidates (nA x 16 bytes per rank) are exchanged, and every rank
 * merges to the result a single GPU would give (bit-identical).  Outputs as hulo_knn2.
 * The exchange the call uses is what hulo_comm_exchange_kind reports:
 *   "peer-store"      (default) the chunk-merge kernel stores each record straight into every
 *                     rank's exchange buffer (CUDA IPC mappings over NVLink) and raises a flag;
 *                     the merge kernel waits on the flags.  NCCL only carried the set-up.
 *   "nccl-allgather"  one ncclAllGather of the packed records, when the peers cannot be mapped
 *                     or HULO_EXCHANGE=nccl is set;
 *   "none"            a world of one.
 * hulo_gpu_destroy of a context that ran the peer-store exchange is collective (the ranks meet
 * between closing their mappings and freeing the exported buffers). */
int hulo_knn2_sharded(hulo_gpu *h, const hulo_db *A, const hulo_db *B_shard, uint64_t row_base,
                      int32_t *idx2, int32_t *dist2);

/* The same search for a STREAM of searcher tables (a server answering query batch after query
 * batch): hulo_knn2_sharded_submit issues search s and returns at once; hulo_knn2_sharded_collect
 * returns the results of the oldest search not collected yet.  At most two searches may be
 * outstanding (the result arrays are double-buffered), so the pattern is
 *   submit(0); for k = 1..: { hulo_db_update(A, batch k); submit(k); collect(k - 1); }  collect(last);
 * and the device never waits for the host: the exchange of search k - 1 and the copy of its
 * results (on a stream of their own) overlap the K1 of search k.  Results are those of
 * hulo_knn2_sharded, bit for bit.  Works on a world of one as well (no exchange, the result
 * copy still overlaps the next search). */
int hulo_knn2_sharded_submit(hulo_gpu *h, const hulo_db *A, const hulo_db *B_shard, uint64_t row_base);
int hulo_knn2_sharded_collect(hulo_gpu *h, int32_t *idx2, int32_t *dist2, size_t *n_rows);

/* The merge step on its own, for hosts that move the candidates themselves (another
 * transport, several nodes): cand holds `world` lists of nA records {d0, i0, d1, i1} (int32,
 * global indices, HULO_IDX_NONE / HULO_DIST_NONE for a missing neighbour), list after list.
 * Writes the (distance, index)-ordered best two per searcher row. */
int hulo_merge_top2(hulo_gpu *h, const int32_t *cand, size_t nA, int world, int32_t *idx2, int32_t *dist2);

#ifdef __cplusplus
}
#endif
#endif /* HULO_GPU_H */
