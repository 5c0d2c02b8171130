/*
 * oracle_resect.c -- CPU fp64 restatement of the resection half of the hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_match.c for the rule).
 *
 * PARITY UNPINNED.  The arithmetic lives in OpenMVG 1.1 (reference README.md:5,
 * CMakeLists.txt:36-39), a third-party dependency that is neither vendored under
 * /root/reference nor installed in this image, and the reference ships no test or golden
 * vector for it.  The call sites that anchor this restatement are
 *   VisionLocalizeServer/src/LocalizeEngine.cc:503-531,
 *   OpenMVGLocalization_AKAZE/src/localization.cpp:479-509,
 *   OpenMVG_BA/src/adjust_sfm_data.cpp:109-137
 * (all call openMVG::sfm::SfM_Localizer::Localize with error_max = inf and the default
 * max_iteration = 4096).  What is restated is OpenMVG 1.1's published algorithm
 * (SURVEY.md appendix B):
 *   sfm/pipelines/localization/SfM_Localizer.cpp          -> orc_localize
 *   robust_estimation/robust_estimator_ACRansac.hpp        -> orc_acransac, orc_best_nfa,
 *                                                             orc_logcombi
 *   robust_estimation/robust_estimator_ACRansacKernelAdaptator.hpp
 *        (ACKernelAdaptorResection_K)                      -> normalisation, logalpha0,
 *                                                             unormalizeError
 *   multiview/solver_resection_p3p.hpp (Kneip CVPR 2011)   -> orc_p3p
 *   multiview/projection.hpp (Project, KRt_From_P)         -> orc_residuals, orc_krt_from_p
 * The P3P solution set is cross-checked against cv2.solveP3P in tests/.
 *
 * INDEPENDENT ANCHORS (tests/test_oracle_anchors.py; the restatement does not check itself):
 *   - orc_logcombi / orc_best_nfa against exact integer binomials (math.comb) on 1000 residual
 *     lists with ties and threshold cuts;
 *   - the final inlier sets of orc_acransac against cv2.solvePnPRansac(SOLVEPNP_P3P) at the
 *     estimated threshold, and against the planted truth (fixtures: tests/golden/anchors_golden.npz,
 *     generator: tests/golden/make_golden_anchors.py).
 *
 * CONSTANTS TAKEN FROM MEMORY OF THE UPSTREAM SOURCE (SURVEY.md appendix B marks them with a
 * dagger).  Each with the upstream file a maintainer can open to confirm it, and why the value is
 * what it is:
 *   max_iteration = 4096          sfm/pipelines/localization/SfM_Localizer.hpp, member initialiser of
 *                                 Image_Localizer_Match_Data; the reference never overrides it
 *                                 (LocalizeEngine.cc:503-505 sets only pt2D / pt3D).
 *   error_max = infinity          same struct; Localize() passes precision = error_max^2 only when
 *                                 it is finite, so AC-RANSAC runs threshold-free (maxThreshold = inf).
 *   MINIMUM_SAMPLES = 3,          multiview/solver_resection_p3p.hpp, P3PSolver: a P3P problem has 3
 *   MAX_MODELS = 4                points and at most 4 real solutions of its quartic.
 *   logalpha0 = log10(pi)         robust_estimator_ACRansacKernelAdaptator.hpp, ACKernelAdaptorResection_K:
 *                                 the residual is a point-to-point distance in K-normalised image
 *                                 coordinates; the probability that a uniform point falls within
 *                                 distance d of a given point is pi d^2 / (unit area), i.e. alpha0 = pi
 *                                 with the squared error e = d^2 entering as log10(e).
 *   multError = 1.0               same adaptor: the error is already squared (0.5 is used by the
 *                                 kernels that hand over unsquared point-to-line distances).
 *   loge0 = log10(MAX_MODELS *    robust_estimator_ACRansac.hpp: number of tests = models per sample
 *           (N - MINIMUM_SAMPLES))  x number of candidate inlier counts.
 *   NFA_k = loge0 + logalpha *    same file, bestNFA(): the classical a-contrario bound
 *     (k - 3) + log10 C(N,k) +      N_tests * C(N,k) * C(k,3) * alpha^(k-3), in log10; the binomial
 *     log10 C(k,3)                  tables are float (std::vector<float>), reproduced as float here.
 *   + FLT_EPSILON                 same function: log10(e + numeric_limits<float>::epsilon()) so that a
 *                                 zero residual (the three sample points) stays finite.
 *   start index k = 4             same function: loops from MINIMUM_SAMPLES + 1; a model is scored on at
 *                                 least one point beyond its own sample.
 *   reserve = max_iter / 10       ACRANSAC(): nIterReserve = nIter / 10, nIter -= nIterReserve; the
 *                                 reserve is spent on samples drawn from the current inlier set once a
 *                                 meaningful (NFA < 0) model exists.
 *   success iff inliers > 2.5 * 3 SfM_Localizer.cpp, Localize(): "resection_data.vec_inliers.size() >
 *                                 MINIMUM_SAMPLES * OPENMVG_MINIMUM_SAMPLES_COEF" with the coefficient 2.5.
 *   unormalizeError(e) =          adaptor: back to pixels through the normalisation K^-1, whose scale is
 *     sqrt(e) * fx                1 / fx (the reference's cameras have fx ~ fy, K.txt:1-3).
 * What remains UNPINNED after the anchors: bit-level agreement of a whole AC-RANSAC trace with
 * OpenMVG's (impossible in principle: upstream seeds its sampler from the clock), and the
 * constants above, which are confirmed only by the statistical agreement of the outcomes.
 * Known deviations: (1) the sampler is a seeded splitmix64 instead of std::rand (upstream
 * seeds non-deterministically, so traces are not comparable anyway); (2) P3P models with
 * non-finite entries are skipped instead of being sorted with NaN residuals; (3) the narrowed
 * sampling pool is kept in index order (distribution preserving, see orc_acransac).
 */
#include <complex.h>
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---------- small vector helpers ---------- */
static void v_sub(const double *a, const double *b, double *o) { o[0]=a[0]-b[0]; o[1]=a[1]-b[1]; o[2]=a[2]-b[2]; }
static void v_cross(const double *a, const double *b, double *o) {
    double x = a[1]*b[2]-a[2]*b[1], y = a[2]*b[0]-a[0]*b[2], z = a[0]*b[1]-a[1]*b[0];
    o[0]=x; o[1]=y; o[2]=z;
}
static double v_dot(const double *a, const double *b) { return a[0]*b[0]+a[1]*b[1]+a[2]*b[2]; }
static double v_norm(const double *a) { return sqrt(v_dot(a, a)); }
static void v_normalize(double *a) { double n = v_norm(a); a[0]/=n; a[1]/=n; a[2]/=n; }
/* o = M v, M row-major 3x3 */
static void m_mulv(const double *M, const double *v, double *o) {
    double x = M[0]*v[0]+M[1]*v[1]+M[2]*v[2];
    double y = M[3]*v[0]+M[4]*v[1]+M[5]*v[2];
    double z = M[6]*v[0]+M[7]*v[1]+M[8]*v[2];
    o[0]=x; o[1]=y; o[2]=z;
}
/* o = M^T v */
static void m_tmulv(const double *M, const double *v, double *o) {
    double x = M[0]*v[0]+M[3]*v[1]+M[6]*v[2];
    double y = M[1]*v[0]+M[4]*v[1]+M[7]*v[2];
    double z = M[2]*v[0]+M[5]*v[1]+M[8]*v[2];
    o[0]=x; o[1]=y; o[2]=z;
}
static void m_mul(const double *A, const double *B, double *O) {
    double T[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            T[3*i+j] = A[3*i]*B[j] + A[3*i+1]*B[3+j] + A[3*i+2]*B[6+j];
    memcpy(O, T, sizeof T);
}
static void m_transpose(const double *A, double *O) {
    double T[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T[3*i+j] = A[3*j+i];
    memcpy(O, T, sizeof T);
}

/* ---------- quartic, Ferrari closed form (real parts of the four complex roots) ---------- */
static void solve_quartic(const double f[5], double roots[4]) {
    double A = f[0], B = f[1], C = f[2], D = f[3], E = f[4];
    double A2 = A*A, B2 = B*B, A3 = A2*A, B3 = B2*B, A4 = A3*A, B4 = B3*B;
    double alpha = -3.0*B2/(8.0*A2) + C/A;
    double beta = B3/(8.0*A3) - B*C/(2.0*A2) + D/A;
    double gamma = -3.0*B4/(256.0*A4) + B2*C/(16.0*A3) - B*D/(4.0*A2) + E/A;
    double alpha2 = alpha*alpha, alpha3 = alpha2*alpha;
    double complex P = -alpha2/12.0 - gamma;
    double complex Q = -alpha3/108.0 + alpha*gamma/3.0 - beta*beta/8.0;
    double complex R = -Q/2.0 + csqrt(Q*Q/4.0 + P*P*P/27.0);
    double complex U = cpow(R, 1.0/3.0);
    double complex y;
    if (creal(U) == 0.0) y = -5.0*alpha/6.0 - cpow(Q, 1.0/3.0);
    else y = -5.0*alpha/6.0 - P/(3.0*U) + U;
    double complex w = csqrt(alpha + 2.0*y);
    double complex s1 = csqrt(-(3.0*alpha + 2.0*y + 2.0*beta/w));
    double complex s2 = csqrt(-(3.0*alpha + 2.0*y - 2.0*beta/w));
    double sh = -B/(4.0*A);
    roots[0] = creal(sh + 0.5*( w + s1));
    roots[1] = creal(sh + 0.5*( w - s1));
    roots[2] = creal(sh + 0.5*(-w + s2));
    roots[3] = creal(sh + 0.5*(-w - s2));
}

/*
 * Kneip P3P.  f: three unit bearing vectors (f[3*i..]), X: three world points (X[3*i..]).
 * models: up to 4 row-major 3x4 [R|t] with x_cam ~ R X + t.  Returns the number of models
 * with finite entries (upstream always emits four, taking the real part of every root).
 */
int orc_p3p(const double *f_in, const double *X_in, double *models) {
    double P1[3], P2[3], P3[3], f1[3], f2[3], f3[3];
    memcpy(P1, X_in, 24); memcpy(P2, X_in+3, 24); memcpy(P3, X_in+6, 24);
    memcpy(f1, f_in, 24); memcpy(f2, f_in+3, 24); memcpy(f3, f_in+6, 24);

    double t1[3], t2[3], cr[3];
    v_sub(P2, P1, t1); v_sub(P3, P1, t2); v_cross(t1, t2, cr);
    if (v_norm(cr) == 0.0) return 0;               /* collinear world points */

    double T[9], e1[3], e2[3], e3[3], f3t[3];
    for (int pass = 0; pass < 2; ++pass) {
        memcpy(e1, f1, 24);
        v_cross(f1, f2, e3); v_normalize(e3);
        v_cross(e3, e1, e2);
        memcpy(T, e1, 24); memcpy(T+3, e2, 24); memcpy(T+6, e3, 24);
        m_mulv(T, f3, f3t);
        if (pass == 0 && f3t[2] > 0.0) {           /* enforce theta in [0, pi]: swap 1 <-> 2 */
            double tmp[3];
            memcpy(tmp, f1, 24); memcpy(f1, f2, 24); memcpy(f2, tmp, 24);
            memcpy(tmp, P1, 24); memcpy(P1, P2, 24); memcpy(P2, tmp, 24);
            continue;
        }
        break;
    }

    double n1[3], n2[3], n3[3], N[9], d[3];
    v_sub(P2, P1, n1); v_normalize(n1);
    v_sub(P3, P1, d);
    v_cross(n1, d, n3); v_normalize(n3);
    v_cross(n3, n1, n2);
    memcpy(N, n1, 24); memcpy(N+3, n2, 24); memcpy(N+6, n3, 24);

    double P3n[3];
    m_mulv(N, d, P3n);
    v_sub(P2, P1, d);
    double d12 = v_norm(d);
    double phi1 = f3t[0]/f3t[2], phi2 = f3t[1]/f3t[2];
    double p1 = P3n[0], p2 = P3n[1];
    double cosb = v_dot(f1, f2);
    double b = 1.0/(1.0 - cosb*cosb) - 1.0;
    b = cosb < 0.0 ? -sqrt(b) : sqrt(b);

    double phi1_2 = phi1*phi1, phi2_2 = phi2*phi2;
    double p1_2 = p1*p1, p1_3 = p1_2*p1, p1_4 = p1_3*p1;
    double p2_2 = p2*p2, p2_3 = p2_2*p2, p2_4 = p2_3*p2;
    double d12_2 = d12*d12, b_2 = b*b;

    double fac[5];
    fac[0] = -phi2_2*p2_4 - p2_4*phi1_2 - p2_4;
    fac[1] = 2.0*p2_3*d12*b + 2.0*phi2_2*p2_3*d12*b - 2.0*phi2*p2_3*phi1*d12;
    fac[2] = -phi2_2*p2_2*p1_2 - phi2_2*p2_2*d12_2*b_2 - phi2_2*p2_2*d12_2 + phi2_2*p2_4
             + p2_4*phi1_2 + 2.0*p1*p2_2*d12 + 2.0*phi1*phi2*p1*p2_2*d12*b
             - p2_2*p1_2*phi1_2 + 2.0*p1*p2_2*phi2_2*d12 - p2_2*d12_2*b_2 - 2.0*p1_2*p2_2;
    fac[3] = 2.0*p1_2*p2*d12*b + 2.0*phi2*p2_3*phi1*d12 - 2.0*phi2_2*p2_3*d12*b
             - 2.0*p1*p2*d12_2*b;
    fac[4] = -2.0*phi2*p2_2*phi1*p1*d12*b + phi2_2*p2_2*d12_2 + 2.0*p1_3*d12 - p1_2*d12_2
             + phi2_2*p2_2*p1_2 - p1_4 - 2.0*phi2_2*p2_2*p1*d12 + p2_2*phi1_2*p1_2
             + phi2_2*p2_2*d12_2*b_2;

    double roots[4];
    solve_quartic(fac, roots);

    int n_out = 0;
    for (int i = 0; i < 4; ++i) {
        double cot_alpha = (-phi1*p1/phi2 - roots[i]*p2 + d12*b)
                         / (-phi1*roots[i]*p2/phi2 + p1 - d12);
        double cos_theta = roots[i];
        double sin_theta = sqrt(1.0 - roots[i]*roots[i]);
        double sin_alpha = sqrt(1.0/(cot_alpha*cot_alpha + 1.0));
        double cos_alpha = sqrt(1.0 - sin_alpha*sin_alpha);
        if (cot_alpha < 0.0) cos_alpha = -cos_alpha;

        double Cn[3] = { d12*cos_alpha*(sin_alpha*b + cos_alpha),
                         cos_theta*d12*sin_alpha*(sin_alpha*b + cos_alpha),
                         sin_theta*d12*sin_alpha*(sin_alpha*b + cos_alpha) };
        double C[3];
        m_tmulv(N, Cn, C);
        C[0] += P1[0]; C[1] += P1[1]; C[2] += P1[2];

        double Rr[9] = { -cos_alpha, -sin_alpha*cos_theta, -sin_alpha*sin_theta,
                          sin_alpha, -cos_alpha*cos_theta, -cos_alpha*sin_theta,
                          0.0,       -sin_theta,            cos_theta };
        /* camera-to-world rotation  Rcw = N^T Rr^T T ; the model uses Rwc = Rcw^T */
        double Nt[9], Rt[9], Rcw[9], Rwc[9];
        m_transpose(N, Nt); m_transpose(Rr, Rt);
        m_mul(Nt, Rt, Rcw); m_mul(Rcw, T, Rcw);
        m_transpose(Rcw, Rwc);
        double tv[3];
        m_mulv(Rwc, C, tv);
        double M[12] = { Rwc[0], Rwc[1], Rwc[2], -tv[0],
                         Rwc[3], Rwc[4], Rwc[5], -tv[1],
                         Rwc[6], Rwc[7], Rwc[8], -tv[2] };
        int ok = 1;
        for (int k = 0; k < 12; ++k) if (!isfinite(M[k])) ok = 0;
        if (!ok) continue;
        memcpy(models + 12*n_out, M, sizeof M);
        ++n_out;
    }
    return n_out;
}

/* K^-1 for an upper-triangular pinhole K (row-major 3x3, K[8] == 1 not assumed). */
static void k_inverse(const double *K, double *Ki) {
    double fx = K[0], s = K[1], cx = K[2], fy = K[4], cy = K[5], w = K[8];
    Ki[0] = 1.0/fx; Ki[1] = -s/(fx*fy); Ki[2] = (s*cy - cx*fy)/(fx*fy*w);
    Ki[3] = 0.0;    Ki[4] = 1.0/fy;     Ki[5] = -cy/(fy*w);
    Ki[6] = 0.0;    Ki[7] = 0.0;        Ki[8] = 1.0/w;
}

/* x2dn = dehomogenised K^-1 [x;1]   (ACKernelAdaptorResection_K constructor).
 * x2d, x2dn: 2xN column-major (x2d[2*i], x2d[2*i+1]). */
void orc_normalize_points(const double *x2d, size_t N, const double *K, double *x2dn) {
    double Ki[9];
    k_inverse(K, Ki);
    for (size_t i = 0; i < N; ++i) {
        double p[3] = { x2d[2*i], x2d[2*i+1], 1.0 }, q[3];
        m_mulv(Ki, p, q);
        x2dn[2*i] = q[0]/q[2];
        x2dn[2*i+1] = q[1]/q[2];
    }
}

/* Squared reprojection residual of every correspondence under the 3x4 model M (row-major),
 * ResectionSquaredResidualError: || hnormalized(M [X;1]) - x ||^2.  X3d is 3xN column-major. */
void orc_residuals(const double *M, const double *x2dn, const double *X3d, size_t N, double *err) {
    for (size_t i = 0; i < N; ++i) {
        const double *X = X3d + 3*i;
        double u = M[0]*X[0] + M[1]*X[1] + M[2]*X[2]  + M[3];
        double v = M[4]*X[0] + M[5]*X[1] + M[6]*X[2]  + M[7];
        double w = M[8]*X[0] + M[9]*X[1] + M[10]*X[2] + M[11];
        double dx = u/w - x2dn[2*i], dy = v/w - x2dn[2*i+1];
        err[i] = dx*dx + dy*dy;
    }
}

/* log10 of the binomial coefficient C(n,k), accumulated in double, stored as float. */
float orc_logcombi(size_t k, size_t n) {
    if (k >= n || k == 0) return 0.0f;
    if (n - k < k) k = n - k;
    double r = 0.0;
    for (size_t i = 1; i <= k; ++i) r += log10((double)(n - i + 1)) - log10((double)i);
    return (float)r;
}
/* logc_n[k] = log10 C(N,k), logc_k[n] = log10 C(n,3), both N+1 floats. */
void orc_make_logcombi(size_t N, float *logc_n, float *logc_k) {
    for (size_t k = 0; k <= N; ++k) logc_n[k] = orc_logcombi(k, N);
    for (size_t n = 0; n <= N; ++n) logc_k[n] = orc_logcombi(3, n);
}

/*
 * bestNFA over residuals sorted ascending (ties already ordered by index).
 * Returns the minimal NFA and writes its k (number of inliers) to *k_best; first minimum wins.
 */
double orc_best_nfa(const double *sorted_err, size_t N, double logalpha0, double loge0,
                    double max_threshold, const float *logc_n, const float *logc_k,
                    double mult_error, size_t *k_best) {
    const size_t start = 3;
    double best = INFINITY;
    size_t bk = start;
    for (size_t k = start + 1; k <= N && sorted_err[k-1] <= max_threshold; ++k) {
        double logalpha = logalpha0 + mult_error * log10(sorted_err[k-1] + (double)FLT_EPSILON);
        double nfa = loge0 + logalpha * (double)(k - start) + (double)logc_n[k] + (double)logc_k[k];
        if (nfa < best) { best = nfa; bk = k; }
    }
    *k_best = bk;
    return best;
}

typedef struct { double e; size_t i; } err_idx;
static int cmp_err_idx(const void *a, const void *b) {
    const err_idx *x = (const err_idx *)a, *y = (const err_idx *)b;
    if (x->e < y->e) return -1;
    if (x->e > y->e) return 1;
    return (x->i > y->i) - (x->i < y->i);
}

/*
 * Score H hypotheses (row-major 3x4 each) against all N correspondences: the inner body of
 * the ACRANSAC model loop.  Outputs per hypothesis: minimal NFA, its k, the k-th smallest
 * squared residual (normalised units) and, when thr2 >= 0, the count of residuals <= thr2.
 */
void orc_score_hypotheses(const double *models, size_t H, const double *x2dn, const double *X3d,
                          size_t N, double thr2, double *nfa, int32_t *k_best, double *err_k,
                          int32_t *n_inl) {
    float *logc_n = (float *)malloc(sizeof(float) * (N + 1));
    float *logc_k = (float *)malloc(sizeof(float) * (N + 1));
    orc_make_logcombi(N, logc_n, logc_k);
    double loge0 = log10(4.0 * (double)(N - 3));
    double logalpha0 = log10(M_PI);
#pragma omp parallel
    {
        double *err = (double *)malloc(sizeof(double) * (N ? N : 1));
        err_idx *ei = (err_idx *)malloc(sizeof(err_idx) * (N ? N : 1));
#pragma omp for schedule(static)
        for (long long h = 0; h < (long long)H; ++h) {
            orc_residuals(models + 12*h, x2dn, X3d, N, err);
            int32_t cnt = 0;
            for (size_t i = 0; i < N; ++i) { ei[i].e = err[i]; ei[i].i = i; if (thr2 >= 0 && err[i] <= thr2) ++cnt; }
            qsort(ei, N, sizeof(err_idx), cmp_err_idx);
            for (size_t i = 0; i < N; ++i) err[i] = ei[i].e;
            size_t kb;
            nfa[h] = orc_best_nfa(err, N, logalpha0, loge0, INFINITY, logc_n, logc_k, 1.0, &kb);
            k_best[h] = (int32_t)kb;
            err_k[h] = (kb >= 1 && kb <= N) ? err[kb-1] : INFINITY;
            if (n_inl) n_inl[h] = cnt;
        }
        free(err); free(ei);
    }
    free(logc_n); free(logc_k);
}

static int cmp_size_t_fwd(const void *a, const void *b) {
    size_t x = *(const size_t *)a, y = *(const size_t *)b;
    return (x > y) - (x < y);
}

/* ---------- sampler ---------- */
static uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
/* three distinct positions in [0, total) (UniformSample of rand_sampling.hpp) */
void orc_sample3(uint64_t *state, size_t total, size_t out[3]) {
    for (int i = 0; i < 3; ++i) {
        size_t r = (size_t)(splitmix64(state) % (uint64_t)(total - i));
        int j;
        for (j = 0; j < i && r >= out[j]; ++j) ++r;
        for (int k = i; k > j; --k) out[k] = out[k-1];
        out[j] = r;
    }
}

/*
 * ACRANSAC over the P3P kernel (robust_estimator_ACRansac.hpp, ACKernelAdaptorResection_K).
 *   x2d 2xN pixels, X3d 3xN, K 3x3 row-major, max_iter (4096 in the reference), seed.
 * Outputs: P = K [R|t] (row-major 3x4), inliers (indices, sorted by residual), *n_inl,
 * *error_max in pixels, *min_nfa.  Returns 1 when a meaningful model (NFA < 0) was found.
 */
int orc_acransac(const double *x2d, const double *X3d, size_t N, const double *K, size_t max_iter,
                 uint64_t seed, double *P, int32_t *inliers, size_t *n_inl, double *error_max,
                 double *min_nfa_out) {
    *n_inl = 0; *error_max = 0.0; if (min_nfa_out) *min_nfa_out = 0.0;
    if (N <= 3) return 0;
    double *x2dn = (double *)malloc(sizeof(double) * 2 * N);
    orc_normalize_points(x2d, N, K, x2dn);
    float *logc_n = (float *)malloc(sizeof(float) * (N + 1));
    float *logc_k = (float *)malloc(sizeof(float) * (N + 1));
    orc_make_logcombi(N, logc_n, logc_k);
    double loge0 = log10(4.0 * (double)(N - 3));
    double logalpha0 = log10(M_PI);
    double *err = (double *)malloc(sizeof(double) * N);
    err_idx *ei = (err_idx *)malloc(sizeof(err_idx) * N);
    size_t *pool = (size_t *)malloc(sizeof(size_t) * N);
    size_t n_pool = N;
    for (size_t i = 0; i < N; ++i) pool[i] = i;
    size_t *best_inl = (size_t *)malloc(sizeof(size_t) * N);
    size_t n_best = 0;
    double best_model[12] = {0};
    double minNFA = INFINITY, errorMax = INFINITY;
    uint64_t rng = seed;

    size_t nIter = max_iter;
    size_t nIterReserve = nIter / 10;
    nIter -= nIterReserve;

    for (size_t iter = 0; iter < nIter; ++iter) {
        size_t pos[3];
        orc_sample3(&rng, n_pool, pos);
        double f[9], Xs[9];
        for (int s = 0; s < 3; ++s) {
            size_t id = pool[pos[s]];
            double b[3] = { x2dn[2*id], x2dn[2*id+1], 1.0 };
            v_normalize(b);
            memcpy(f + 3*s, b, 24);
            memcpy(Xs + 3*s, X3d + 3*id, 24);
        }
        double models[48];
        int nm = orc_p3p(f, Xs, models);
        int better = 0;
        for (int m = 0; m < nm; ++m) {
            orc_residuals(models + 12*m, x2dn, X3d, N, err);
            for (size_t i = 0; i < N; ++i) { ei[i].e = err[i]; ei[i].i = i; }
            qsort(ei, N, sizeof(err_idx), cmp_err_idx);
            for (size_t i = 0; i < N; ++i) err[i] = ei[i].e;
            size_t kb;
            double nfa = orc_best_nfa(err, N, logalpha0, loge0, INFINITY, logc_n, logc_k, 1.0, &kb);
            if (nfa < minNFA) {
                better = 1;
                minNFA = nfa;
                n_best = kb;
                for (size_t i = 0; i < kb; ++i) best_inl[i] = ei[i].i;
                errorMax = err[kb-1];
                memcpy(best_model, models + 12*m, sizeof best_model);
            }
        }
        if ((better && minNFA < 0) || (iter + 1 == nIter && nIterReserve)) {
            if (n_best == 0) {
                nIter++;
                nIterReserve--;
            } else {
                /* Known deviation (3): the narrowed pool is kept in ascending index order, not in
                 * residual order -- see orc_fmatrix_acransac below for why this changes nothing in
                 * distribution and makes traces comparable between implementations. */
                n_pool = n_best;
                memcpy(pool, best_inl, sizeof(size_t) * n_best);
                qsort(pool, n_pool, sizeof(size_t), cmp_size_t_fwd);
                if (nIterReserve) {
                    nIter = iter + 1 + nIterReserve;
                    nIterReserve = 0;
                }
            }
        }
    }
    int ok = 0;
    if (minNFA >= 0) n_best = 0;
    if (n_best > 0) {
        /* Unnormalize: P = K * model ; error in pixels = sqrt(e) / Kinv(0,0) */
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c)
                P[4*r+c] = K[3*r]*best_model[c] + K[3*r+1]*best_model[4+c] + K[3*r+2]*best_model[8+c];
        *error_max = sqrt(errorMax) * K[0];
        for (size_t i = 0; i < n_best; ++i) inliers[i] = (int32_t)best_inl[i];
        *n_inl = n_best;
        ok = 1;
    }
    if (min_nfa_out) *min_nfa_out = minNFA;
    free(x2dn); free(logc_n); free(logc_k); free(err); free(ei); free(pool); free(best_inl);
    return ok;
}

/*
 * SfM_Localizer::Localize acceptance: success iff #inliers > 2.5 * 3.
 * The engines then require #inliers > 10 (LocalizeEngine.cc:560, localization.cpp:511).
 */
int orc_localize(const double *x2d, const double *X3d, size_t N, const double *K, size_t max_iter,
                 uint64_t seed, double *P, int32_t *inliers, size_t *n_inl, double *error_max) {
    double nfa;
    orc_acransac(x2d, X3d, N, K, max_iter, seed, P, inliers, n_inl, error_max, &nfa);
    return (double)*n_inl > 2.5 * 3.0;
}

/*
 * KRt_From_P for P = K [R|t] with K upper triangular, positive diagonal, K[8] normalised to 1
 * (RQ decomposition of the left 3x3 block by Givens rotations), projection.hpp.
 * Also returns the camera centre c = -R^T t (LocalizeEngine.cc:582-585).
 */
void orc_krt_from_p(const double *P, double *Kout, double *Rout, double *tout, double *centre) {
    double Kk[9] = { P[0], P[1], P[2], P[4], P[5], P[6], P[8], P[9], P[10] };
    double Q[9] = {1,0,0, 0,1,0, 0,0,1};
    /* zero K(2,1) with a rotation about x */
    {
        double c = -Kk[8], s = Kk[7], l = sqrt(c*c + s*s); c /= l; s /= l;
        double Rx[9] = {1,0,0, 0,c,-s, 0,s,c};
        m_mul(Kk, Rx, Kk); m_mul(Q, Rx, Q);
    }
    /* zero K(2,0) with a rotation about y */
    {
        double c = Kk[8], s = Kk[6], l = sqrt(c*c + s*s); c /= l; s /= l;
        double Ry[9] = {c,0,s, 0,1,0, -s,0,c};
        m_mul(Kk, Ry, Kk); m_mul(Q, Ry, Q);
    }
    /* zero K(1,0) with a rotation about z */
    {
        double c = -Kk[4], s = Kk[3], l = sqrt(c*c + s*s); c /= l; s /= l;
        double Rz[9] = {c,-s,0, s,c,0, 0,0,1};
        m_mul(Kk, Rz, Kk); m_mul(Q, Rz, Q);
    }
    double R[9];
    m_transpose(Q, R);
    /* make the diagonal of K positive */
    for (int ax = 0; ax < 3; ++ax) {
        if (Kk[4*ax] < 0) {
            for (int r = 0; r < 3; ++r) Kk[3*r+ax] = -Kk[3*r+ax];
            for (int c = 0; c < 3; ++c) R[3*ax+c] = -R[3*ax+c];
        }
    }
    /* det(R) must be +1 */
    double det = R[0]*(R[4]*R[8]-R[5]*R[7]) - R[1]*(R[3]*R[8]-R[5]*R[6]) + R[2]*(R[3]*R[7]-R[4]*R[6]);
    double p4[3] = { P[3], P[7], P[11] };
    if (det < 0) { for (int k = 0; k < 9; ++k) R[k] = -R[k]; p4[0] = -p4[0]; p4[1] = -p4[1]; p4[2] = -p4[2]; }
    /* t = K^-1 p4 */
    double Ki[9]; k_inverse(Kk, Ki);
    /* k_inverse assumes upper-triangular input, true after the three rotations */
    double t[3]; m_mulv(Ki, p4, t);
    double sc = Kk[8];
    for (int k = 0; k < 9; ++k) Kk[k] /= sc;
    if (Kout) memcpy(Kout, Kk, sizeof Kk);
    if (Rout) memcpy(Rout, R, sizeof R);
    if (tout) memcpy(tout, t, sizeof t);
    if (centre) { double c[3]; m_tmulv(R, t, c); centre[0] = -c[0]; centre[1] = -c[1]; centre[2] = -c[2]; }
}

/* =====================================================================================
 * Fundamental-matrix geometric filter: CPU restatement of hulo::geometricMatch
 * (VisionLocalizeCommon/src/MatchUtils.cpp:372-420), which runs OpenMVG 1.1's
 * GeometricFilter_FMatrix_AC(geomPrec, ransacRound) on every putative pair
 * (callers: LocalizeEngine.cc:458, localization.cpp:450, computeFeaturesAndMatches.cpp:242).
 * PARITY UNPINNED for the same reason as the resection above (OpenMVG is not vendored and
 * the reference has no test).  Upstream files restated:
 *   matching_image_collection/F_ACRobust.hpp                -> orc_fmatrix_acransac
 *   robust_estimation/robust_estimator_ACRansacKernelAdaptator.hpp (ACKernelAdaptor,
 *        point-to-line)                                     -> normalisation, logalpha0,
 *                                                              multError 0.5, unormalizeError
 *   multiview/conditioning.cpp (PreconditionerFromPoints)   -> orc_precondition
 *   multiview/solver_fundamental_kernel.{hpp,cpp} (SevenPointSolver, EpipolarDistanceError)
 *                                                           -> orc_seven_point, orc_epipolar_errors
 *   numeric/poly.h (SolveCubicPolynomial)                   -> solve_cubic
 * The 7-point solution set is cross-checked against cv2.findFundamentalMat(FM_7POINT).
 * Known deviations: seeded splitmix64 sampler (as the resection above); non-finite models
 * skipped; the narrowed sampling pool kept in index order (see orc_fmatrix_acransac).
 * ===================================================================================== */

/* T = [[s,0,-w s/2],[0,s,-h s/2],[0,0,1]], s = 1/sqrt(w h) */
void orc_precondition(int w, int h, double *T) {
    double s = 1.0 / sqrt((double)w * (double)h);
    T[0] = s; T[1] = 0; T[2] = -0.5 * w * s;
    T[3] = 0; T[4] = s; T[5] = -0.5 * h * s;
    T[6] = 0; T[7] = 0; T[8] = 1.0;
}

static double det3r(const double *a, const double *b, const double *c) {
    return a[0]*(b[1]*c[2]-b[2]*c[1]) - a[1]*(b[0]*c[2]-b[2]*c[0]) + a[2]*(b[0]*c[1]-b[1]*c[0]);
}

/* real roots of x^3 + a x^2 + b x + c, ascending (the GSL scheme upstream uses) */
static int solve_cubic(double a, double b, double c, double *x) {
    double q = a*a - 3*b, r = 2*a*a*a - 9*a*b + 27*c;
    double Q = q/9, R = r/54, Q3 = Q*Q*Q, R2 = R*R;
    double CR2 = 729*r*r, CQ3 = 2916*q*q*q;
    if (R == 0 && Q == 0) { x[0] = x[1] = x[2] = -a/3; return 3; }
    if (CR2 == CQ3) {
        double sqrtQ = sqrt(Q);
        if (R > 0) { x[0] = -2*sqrtQ - a/3; x[1] = sqrtQ - a/3; x[2] = sqrtQ - a/3; }
        else       { x[0] = -sqrtQ - a/3;   x[1] = -sqrtQ - a/3; x[2] = 2*sqrtQ - a/3; }
        return 3;
    }
    if (CR2 < CQ3) {
        double sqrtQ = sqrt(Q), sqrtQ3 = sqrtQ*sqrtQ*sqrtQ, theta = acos(R/sqrtQ3), norm = -2*sqrtQ;
        x[0] = norm*cos(theta/3) - a/3;
        x[1] = norm*cos((theta + 2.0*M_PI)/3) - a/3;
        x[2] = norm*cos((theta - 2.0*M_PI)/3) - a/3;
        for (int i = 0; i < 2; ++i) for (int j = 0; j < 2 - i; ++j)
            if (x[j] > x[j+1]) { double t = x[j]; x[j] = x[j+1]; x[j+1] = t; }
        return 3;
    }
    double sgnR = R >= 0 ? 1 : -1;
    double A = -sgnR * pow(fabs(R) + sqrt(R2 - Q3), 1.0/3.0);
    double B = Q / A;
    x[0] = A + B - a/3;
    return 1;
}

/* Two vectors spanning the null space of the 7x9 epipolar system (Gauss-Jordan with full
 * pivoting).  Any basis gives the same solution set {F1 + a F2 : det = 0} up to scale. */
static int nullspace2(double A[7][9], double *f1, double *f2) {
    int colperm[9];
    for (int c = 0; c < 9; ++c) colperm[c] = c;
    for (int k = 0; k < 7; ++k) {
        int pr = k, pc = k; double best = 0;
        for (int r = k; r < 7; ++r) for (int c = k; c < 9; ++c)
            if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); pr = r; pc = c; }
        if (best < 1e-14) return 0;                       /* rank deficient sample */
        if (pr != k) for (int c = 0; c < 9; ++c) { double t = A[k][c]; A[k][c] = A[pr][c]; A[pr][c] = t; }
        if (pc != k) { for (int r = 0; r < 7; ++r) { double t = A[r][k]; A[r][k] = A[r][pc]; A[r][pc] = t; }
                       int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t; }
        double inv = 1.0 / A[k][k];
        for (int c = 0; c < 9; ++c) A[k][c] *= inv;
        for (int r = 0; r < 7; ++r) if (r != k) {
            double m = A[r][k];
            if (m != 0) for (int c = 0; c < 9; ++c) A[r][c] -= m * A[k][c];
        }
    }
    /* free variables: permuted columns 7 and 8 */
    for (int v = 0; v < 2; ++v) {
        double z[9];
        for (int k = 0; k < 7; ++k) z[k] = -A[k][7 + v];
        z[7] = v == 0 ? 1 : 0; z[8] = v == 1 ? 1 : 0;
        double *f = v == 0 ? f1 : f2;
        for (int k = 0; k < 9; ++k) f[colperm[k]] = z[k];
    }
    return 1;
}

/* SevenPointSolver: x1, x2 are 7 x 2 (normalised) points; up to three row-major 3x3 F with
 * x2^T F x1 = 0 and det F = 0, in ascending order of the cubic's root.  Returns the count. */
int orc_seven_point(const double *x1, const double *x2, double *F) {
    double A[7][9];
    for (int i = 0; i < 7; ++i) {
        double a = x1[2*i], b = x1[2*i+1], c = x2[2*i], d = x2[2*i+1];
        A[i][0] = c*a; A[i][1] = c*b; A[i][2] = c;
        A[i][3] = d*a; A[i][4] = d*b; A[i][5] = d;
        A[i][6] = a;   A[i][7] = b;   A[i][8] = 1.0;
    }
    double f1[9], f2[9];
    if (!nullspace2(A, f1, f2)) return 0;
    /* det(F1 + t F2) = P0 + P1 t + P2 t^2 + P3 t^3 by multilinearity in the rows */
    const double *r1 = f1, *r2 = f1 + 3, *r3 = f1 + 6, *s1 = f2, *s2 = f2 + 3, *s3 = f2 + 6;
    double P0 = det3r(r1, r2, r3);
    double P1 = det3r(s1, r2, r3) + det3r(r1, s2, r3) + det3r(r1, r2, s3);
    double P2 = det3r(s1, s2, r3) + det3r(s1, r2, s3) + det3r(r1, s2, s3);
    double P3 = det3r(s1, s2, s3);
    double roots[3];
    int n;
    if (P3 == 0) return 0;
    n = solve_cubic(P2/P3, P1/P3, P0/P3, roots);
    int n_out = 0;
    for (int k = 0; k < n; ++k) {
        int ok = 1;
        for (int e = 0; e < 9; ++e) { F[9*n_out + e] = f1[e] + roots[k]*f2[e]; if (!isfinite(F[9*n_out + e])) ok = 0; }
        if (ok) ++n_out;
    }
    return n_out;
}

/* EpipolarDistanceError (= SimpleError): squared distance of x2 to the epipolar line F x1. */
void orc_epipolar_errors(const double *F, const double *x1, const double *x2, size_t N, double *err) {
    for (size_t i = 0; i < N; ++i) {
        double a = x1[2*i], b = x1[2*i+1];
        double l0 = F[0]*a + F[1]*b + F[2], l1 = F[3]*a + F[4]*b + F[5], l2 = F[6]*a + F[7]*b + F[8];
        double d = l0*x2[2*i] + l1*x2[2*i+1] + l2;
        err[i] = d*d / (l0*l0 + l1*l1);
    }
}

float orc_logcombi_k(size_t k, size_t n) { return orc_logcombi(k, n); }

static int cmp_size_t(const void *a, const void *b) {
    size_t x = *(const size_t *)a, y = *(const size_t *)b;
    return (x > y) - (x < y);
}

/* 7 distinct positions in [0,total), ascending insertion like UniformSample */
static void sample_k(uint64_t *state, size_t total, int k, size_t *out) {
    for (int i = 0; i < k; ++i) {
        size_t r = (size_t)(splitmix64(state) % (uint64_t)(total - i));
        int j;
        for (j = 0; j < i && r >= out[j]; ++j) ++r;
        for (int m = i; m > j; --m) out[m] = out[m-1];
        out[j] = r;
    }
}

/*
 * GeometricFilter_FMatrix_AC::Robust_estimation for one pair.
 *   xI, xJ: N x 2 pixel coordinates of the putative matches; image sizes (wI,hI), (wJ,hJ);
 *   precision_px = geomPrec (infinity allowed), max_iter = ransacRound, seed.
 * Outputs: F (row-major, pixel coordinates: x_J^T F x_I = 0), inliers (indices sorted by
 * residual), *n_inl, *error_max [px], *min_nfa.  Returns 1 iff #inliers > 2.5 * 7.
 */
int orc_fmatrix_acransac(const double *xI, const double *xJ, size_t N, int wI, int hI, int wJ, int hJ,
                         double precision_px, size_t max_iter, uint64_t seed, double *Fout, int32_t *inliers,
                         size_t *n_inl, double *error_max, double *min_nfa_out) {
    *n_inl = 0; *error_max = 0.0; if (min_nfa_out) *min_nfa_out = 0.0;
    const size_t S = 7;
    if (N <= S) return 0;
    double N1[9], N2[9];
    orc_precondition(wI, hI, N1); orc_precondition(wJ, hJ, N2);
    double *x1 = (double *)malloc(sizeof(double)*2*N), *x2 = (double *)malloc(sizeof(double)*2*N);
    for (size_t i = 0; i < N; ++i) {
        x1[2*i] = N1[0]*xI[2*i] + N1[2]; x1[2*i+1] = N1[4]*xI[2*i+1] + N1[5];
        x2[2*i] = N2[0]*xJ[2*i] + N2[2]; x2[2*i+1] = N2[4]*xJ[2*i+1] + N2[5];
    }
    double D = sqrt((double)wJ*wJ + (double)hJ*hJ), Ar = (double)wJ*(double)hJ;
    double logalpha0 = log10(2.0*D/Ar / N2[0]);
    double mult = 0.5;
    double maxThr = isinf(precision_px) ? INFINITY : precision_px*precision_px * N2[0]*N2[0];
    double loge0 = log10(3.0 * (double)(N - S));
    float *logc_n = (float *)malloc(sizeof(float)*(N+1)), *logc_k = (float *)malloc(sizeof(float)*(N+1));
    for (size_t k = 0; k <= N; ++k) { logc_n[k] = orc_logcombi(k, N); logc_k[k] = orc_logcombi(S, k); }
    double *err = (double *)malloc(sizeof(double)*N);
    err_idx *ei = (err_idx *)malloc(sizeof(err_idx)*N);
    size_t *pool = (size_t *)malloc(sizeof(size_t)*N), n_pool = N;
    for (size_t i = 0; i < N; ++i) pool[i] = i;
    size_t *best_inl = (size_t *)malloc(sizeof(size_t)*N), n_best = 0;
    double best_F[9] = {0}, minNFA = INFINITY, errorMax = INFINITY;
    uint64_t rng = seed;
    size_t nIter = max_iter, nIterReserve = nIter/10;
    nIter -= nIterReserve;
    for (size_t iter = 0; iter < nIter; ++iter) {
        size_t pos[7];
        sample_k(&rng, n_pool, 7, pos);
        double s1[14], s2[14];
        for (int s = 0; s < 7; ++s) {
            size_t id = pool[pos[s]];
            s1[2*s] = x1[2*id]; s1[2*s+1] = x1[2*id+1]; s2[2*s] = x2[2*id]; s2[2*s+1] = x2[2*id+1];
        }
        double Fs[27];
        int nm = orc_seven_point(s1, s2, Fs);
        int better = 0;
        for (int m = 0; m < nm; ++m) {
            orc_epipolar_errors(Fs + 9*m, x1, x2, N, err);
            for (size_t i = 0; i < N; ++i) { ei[i].e = isnan(err[i]) ? INFINITY : err[i]; ei[i].i = i; }
            qsort(ei, N, sizeof(err_idx), cmp_err_idx);
            /* bestNFA with startIndex 7 and the precision bound */
            double bn = INFINITY; size_t bk = S;
            for (size_t k = S + 1; k <= N && ei[k-1].e <= maxThr; ++k) {
                double logalpha = logalpha0 + mult * log10(ei[k-1].e + (double)FLT_EPSILON);
                double nfa = loge0 + logalpha*(double)(k - S) + (double)logc_n[k] + (double)logc_k[k];
                if (nfa < bn) { bn = nfa; bk = k; }
            }
            if (bn < minNFA) {
                better = 1; minNFA = bn; n_best = bk;
                for (size_t i = 0; i < bk; ++i) best_inl[i] = ei[i].i;
                errorMax = ei[bk-1].e;
                memcpy(best_F, Fs + 9*m, sizeof best_F);
            }
        }
        if ((better && minNFA < 0) || (iter + 1 == nIter && nIterReserve)) {
            if (n_best == 0) { nIter++; nIterReserve--; }
            else {
                /* Known deviation: the narrowed pool is kept in ascending index order, not in
                 * residual order.  The order of a pool only decides which uniformly drawn
                 * position selects which element, so the sampling distribution is unchanged;
                 * it removes the dependence of the whole trace on the rounding-noise order of
                 * the seven (zero-residual) sample points, which no two floating-point
                 * implementations share.  The inlier list returned stays in residual order. */
                n_pool = n_best;
                memcpy(pool, best_inl, sizeof(size_t)*n_best);
                qsort(pool, n_pool, sizeof(size_t), cmp_size_t);
                if (nIterReserve) { nIter = iter + 1 + nIterReserve; nIterReserve = 0; }
            }
        }
    }
    int ok = 0;
    if (minNFA >= 0) n_best = 0;
    if (n_best > 0) {
        /* Unnormalize: F = N2^T F N1 ; error = sqrt(e) / N2(0,0) */
        double T[9], N2t[9];
        m_transpose(N2, N2t); m_mul(N2t, best_F, T); m_mul(T, N1, Fout);
        *error_max = sqrt(errorMax) / N2[0];
        for (size_t i = 0; i < n_best; ++i) inliers[i] = (int32_t)best_inl[i];
        *n_inl = n_best;
        ok = (double)n_best > 2.5 * 7.0;
    }
    if (min_nfa_out) *min_nfa_out = minNFA;
    free(x1); free(x2); free(logc_n); free(logc_k); free(err); free(ei); free(pool); free(best_inl);
    return ok;
}

/* NFA score of one F hypothesis (normalised coordinates) against all matches: the per-model
 * body of the loop above, for per-hypothesis parity checks of the GPU scoring. */
double orc_fmatrix_score(const double *F, const double *x1n, const double *x2n, size_t N, double logalpha0,
                         double max_thr, size_t *k_best, double *err_k) {
    const size_t S = 7;
    double *err = (double *)malloc(sizeof(double)*(N ? N : 1));
    orc_epipolar_errors(F, x1n, x2n, N, err);
    for (size_t i = 0; i < N; ++i) if (isnan(err[i])) err[i] = INFINITY;
    /* plain ascending sort of the values is enough for the score */
    err_idx *ei = (err_idx *)malloc(sizeof(err_idx)*(N ? N : 1));
    for (size_t i = 0; i < N; ++i) { ei[i].e = err[i]; ei[i].i = i; }
    qsort(ei, N, sizeof(err_idx), cmp_err_idx);
    double loge0 = log10(3.0 * (double)(N > S ? N - S : 1));
    double bn = INFINITY; size_t bk = S;
    for (size_t k = S + 1; k <= N && ei[k-1].e <= max_thr; ++k) {
        double logalpha = logalpha0 + 0.5 * log10(ei[k-1].e + (double)FLT_EPSILON);
        double nfa = loge0 + logalpha*(double)(k - S) + (double)orc_logcombi(k, N) + (double)orc_logcombi(S, k);
        if (nfa < bn) { bn = nfa; bk = k; }
    }
    *k_best = bk;
    *err_k = (bk >= 1 && bk <= N) ? ei[bk-1].e : INFINITY;
    free(err); free(ei);
    return bn;
}
