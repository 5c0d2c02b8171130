// geom.cu -- K3: the fundamental-matrix geometric filter between putative matching and the
// 2D-3D assembly, batched over image pairs, sm_100a.
//
// Stands behind hulo::geometricMatch (VisionLocalizeCommon/src/MatchUtils.cpp:372-420), i.e.
// OpenMVG 1.1's ImageCollectionGeometricFilter::Robust_model_estimation with
// GeometricFilter_FMatrix_AC(geomPrec, ransacRound), called per query at
// VisionLocalizeServer/src/LocalizeEngine.cc:458 and OpenMVGLocalization_AKAZE/src/localization.cpp:450
// and once per dataset at ExtFeatAndMatch/src/computeFeaturesAndMatches.cpp:242.  OpenMVG is
// third-party and not vendored; the algorithm is the published one (F_ACRobust.hpp,
// robust_estimator_ACRansac.hpp, solver_fundamental_kernel.cpp), restated on the CPU by the
// test suite's checker.
//
// One thread block runs the whole AC-RANSAC of one image pair, because the schedule is
// sequential by construction (the sampling pool narrows to the inliers of the best model as
// soon as a meaningful one exists, and the iteration budget changes with it):
//   * the sampler is counter based (splitmix64 at a fixed offset per iteration), so the
//     7-point problems of the next kGeoAhead iterations are solved by that many threads at once
//     (fp64: null space of the 7x9 system by Gauss-Jordan with full pivoting, cubic for det F = 0)
//     and re-solved only when the pool changes;
//   * per model: point-to-epipolar-line residuals of all N matches in fp64, stored as fp32 and
//     packed with the match index into one 64-bit key; only keys within the precision bound
//     are kept (ballot compaction), which makes an outlier-contaminated model cost one pass
//     over the matches and nothing else; survivors are bitonic-sorted in shared memory
//     ((residual, index) order = the order of the sequential std::sort) and scanned for
//     nfa_k = loge0 + (logalpha0 + 0.5 log10(e_k + FLT_EPSILON)) (k-7) + logC(N,k) + logC(k,7)
//     with a block-wide lexicographic (nfa, k) minimum (first minimum wins).
// The narrowed sampling pool is kept in ascending index order (the inlier list that is returned
// stays in residual order): a pool's order only permutes which uniformly drawn position picks
// which element, and this way the trace (samples, pool updates, iteration count) does not hang
// on the rounding-noise order of the seven zero-residual sample points.  For the same seed the
// trace then equals the sequential CPU restatement's; values differ only by device-vs-libm
// rounding of cos/acos/pow/log10 and the fp32 residual keys.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "context.cuh"

namespace hulo {
namespace {

constexpr int kGeoThreads = 128;
constexpr int kGeoAhead = 32;              // iterations whose 7-point problems are solved together
constexpr uint32_t kGeoMaxMatches = 16384; // per pair: 128 KB of sort keys
constexpr uint64_t kGamma = 0x9E3779B97F4A7C15ULL;

struct GeoParams {
    const double2 *xI, *xJ;     // pixel coordinates of the putative matches, all pairs back to back
    const uint64_t *off;        // n_pairs + 1
    const int32_t *sizes;       // n_pairs x {wI, hI, wJ, hJ}
    const double *lfact;        // log10(n!) for n = 0 .. max N
    double precision_px;
    uint32_t max_iter;
    uint64_t seed;
    const uint64_t *pair_seeds; // optional: per-pair sampler seeds
    int32_t *valid;
    uint32_t *n_inl;
    double *F, *err_max, *nfa;
    int32_t *inl, *pool;
};

struct NfaMin { double nfa; int k; };
__device__ __forceinline__ NfaMin nfa_min(NfaMin a, NfaMin b) {
    if (b.nfa < a.nfa || (b.nfa == a.nfa && b.k < a.k)) return b;
    return a;
}

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// seven distinct positions in [0, total), ascending insertion (UniformSample); iteration `it`
// owns draws 7 it .. 7 it + 6 of the splitmix64 stream started at `seed`
__device__ void sample7(uint64_t seed, uint32_t it, uint32_t total, uint32_t out[7]) {
    uint64_t s = seed + (uint64_t)it * 7ull * kGamma;
    for (int i = 0; i < 7; ++i) {
        s += kGamma;
        uint32_t r = (uint32_t)(mix64(s) % (uint64_t)(total - i));
        int j;
        for (j = 0; j < i && r >= out[j]; ++j) ++r;
        for (int m = i; m > j; --m) out[m] = out[m - 1];
        out[j] = r;
    }
}

__device__ __forceinline__ double det3r(const double *a, const double *b, const double *c) {
    return a[0] * (b[1] * c[2] - b[2] * c[1]) - a[1] * (b[0] * c[2] - b[2] * c[0]) + a[2] * (b[0] * c[1] - b[1] * c[0]);
}

// real roots of x^3 + a x^2 + b x + c, ascending
__device__ int solve_cubic(double a, double b, double c, double *x) {
    const double q = a * a - 3 * b, r = 2 * a * a * a - 9 * a * b + 27 * c;
    const double Q = q / 9, R = r / 54, Q3 = Q * Q * Q, R2 = R * R;
    const double CR2 = 729 * r * r, CQ3 = 2916 * q * q * q;
    if (R == 0 && Q == 0) { x[0] = x[1] = x[2] = -a / 3; return 3; }
    if (CR2 == CQ3) {
        const double sqrtQ = sqrt(Q);
        if (R > 0) { x[0] = -2 * sqrtQ - a / 3; x[1] = sqrtQ - a / 3; x[2] = sqrtQ - a / 3; }
        else       { x[0] = -sqrtQ - a / 3;     x[1] = -sqrtQ - a / 3; x[2] = 2 * sqrtQ - a / 3; }
        return 3;
    }
    if (CR2 < CQ3) {
        const double sqrtQ = sqrt(Q), sqrtQ3 = sqrtQ * sqrtQ * sqrtQ, theta = acos(R / sqrtQ3), norm = -2 * sqrtQ;
        x[0] = norm * cos(theta / 3) - a / 3;
        x[1] = norm * cos((theta + 2.0 * M_PI) / 3) - a / 3;
        x[2] = norm * cos((theta - 2.0 * M_PI) / 3) - a / 3;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2 - i; ++j)
                if (x[j] > x[j + 1]) { const double t = x[j]; x[j] = x[j + 1]; x[j + 1] = t; }
        return 3;
    }
    const double sgnR = R >= 0 ? 1 : -1;
    const double A = -sgnR * pow(fabs(R) + sqrt(R2 - Q3), 1.0 / 3.0);
    const double B = Q / A;
    x[0] = A + B - a / 3;
    return 1;
}

// SevenPointSolver: x1, x2 seven normalised points (x, y interleaved); up to three row-major F
// with x2^T F x1 = 0 and det F = 0, ascending in the cubic's root.  Returns the count.
__device__ int seven_point(const double *x1, const double *x2, double *F) {
    double A[7][9];
    int colperm[9];
    for (int i = 0; i < 7; ++i) {
        const double a = x1[2 * i], b = x1[2 * i + 1], c = x2[2 * i], d = x2[2 * i + 1];
        A[i][0] = c * a; A[i][1] = c * b; A[i][2] = c;
        A[i][3] = d * a; A[i][4] = d * b; A[i][5] = d;
        A[i][6] = a;     A[i][7] = b;     A[i][8] = 1.0;
    }
    for (int c = 0; c < 9; ++c) colperm[c] = c;
    for (int k = 0; k < 7; ++k) {
        int pr = k, pc = k;
        double best = 0;
        for (int r = k; r < 7; ++r)
            for (int c = k; c < 9; ++c)
                if (fabs(A[r][c]) > best) { best = fabs(A[r][c]); pr = r; pc = c; }
        if (best < 1e-14) return 0;   // rank deficient sample
        if (pr != k)
            for (int c = 0; c < 9; ++c) { const double t = A[k][c]; A[k][c] = A[pr][c]; A[pr][c] = t; }
        if (pc != k) {
            for (int r = 0; r < 7; ++r) { const double t = A[r][k]; A[r][k] = A[r][pc]; A[r][pc] = t; }
            const int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t;
        }
        const double inv = 1.0 / A[k][k];
        for (int c = 0; c < 9; ++c) A[k][c] *= inv;
        for (int r = 0; r < 7; ++r)
            if (r != k) {
                const double m = A[r][k];
                if (m != 0)
                    for (int c = 0; c < 9; ++c) A[r][c] -= m * A[k][c];
            }
    }
    double f1[9], f2[9];
    for (int v = 0; v < 2; ++v) {
        double *f = v == 0 ? f1 : f2;
        for (int k = 0; k < 7; ++k) f[colperm[k]] = -A[k][7 + v];
        f[colperm[7]] = v == 0 ? 1.0 : 0.0;
        f[colperm[8]] = v == 1 ? 1.0 : 0.0;
    }
    const double *r1 = f1, *r2 = f1 + 3, *r3 = f1 + 6, *s1 = f2, *s2 = f2 + 3, *s3 = f2 + 6;
    const double P0 = det3r(r1, r2, r3);
    const double P1 = det3r(s1, r2, r3) + det3r(r1, s2, r3) + det3r(r1, r2, s3);
    const double P2 = det3r(s1, s2, r3) + det3r(s1, r2, s3) + det3r(r1, s2, s3);
    const double P3 = det3r(s1, s2, s3);
    if (P3 == 0) return 0;
    double roots[3];
    const int n = solve_cubic(P2 / P3, P1 / P3, P0 / P3, roots);
    int n_out = 0;
    for (int k = 0; k < n; ++k) {
        bool ok = true;
        for (int e = 0; e < 9; ++e) {
            const double v = f1[e] + roots[k] * f2[e];
            F[9 * n_out + e] = v;
            ok = ok && isfinite(v);
        }
        if (ok) ++n_out;
    }
    return n_out;
}

// log10 C(n, k) as the sequential definition returns it (0 outside 0 < k < n), from log10 n!
__device__ __forceinline__ float logcombi(const double *__restrict__ lfact, uint32_t k, uint32_t n) {
    if (k >= n || k == 0) return 0.0f;
    return (float)(lfact[n] - lfact[k] - lfact[n - k]);
}

__global__ void __launch_bounds__(kGeoThreads) fmatrix_acransac_kernel(GeoParams g) {
    extern __shared__ unsigned long long s_keys[];
    __shared__ double s_models[kGeoAhead * 27];
    __shared__ int s_nm[kGeoAhead];
    __shared__ double s_bestF[9];
    __shared__ NfaMin s_red[kGeoThreads / 32];
    __shared__ NfaMin s_best;
    __shared__ uint32_t s_cnt;
    __shared__ uint32_t s_mask[kGeoMaxMatches / 32];   // membership of the narrowed pool
    __shared__ uint32_t s_scan[kGeoThreads];

    const uint32_t p = blockIdx.x;
    const int tid = threadIdx.x;
    const uint64_t off = g.off[p];
    const uint32_t N = (uint32_t)(g.off[p + 1] - off);
    if (tid == 0) {
        g.valid[p] = 0;
        g.n_inl[p] = 0;
        g.err_max[p] = 0.0;
        g.nfa[p] = INFINITY;
    }
    if (tid < 9) g.F[(size_t)p * 9 + tid] = 0.0;
    if (N <= 7 || g.max_iter == 0) return;   // ACRANSAC: nothing to do with N <= MINIMUM_SAMPLES

    const double2 *__restrict__ xI = g.xI + off;
    const double2 *__restrict__ xJ = g.xJ + off;
    int32_t *inl = g.inl + off, *pool = g.pool + off;
    // ACKernelAdaptor: both images preconditioned by T = [[s,0,-w s/2],[0,s,-h s/2],[0,0,1]], s = 1/sqrt(w h)
    const double wI = g.sizes[4 * p], hI = g.sizes[4 * p + 1], wJ = g.sizes[4 * p + 2], hJ = g.sizes[4 * p + 3];
    const double s1 = 1.0 / sqrt(wI * hI), c1x = -0.5 * wI * s1, c1y = -0.5 * hI * s1;
    const double s2 = 1.0 / sqrt(wJ * hJ), c2x = -0.5 * wJ * s2, c2y = -0.5 * hJ * s2;
    const double logalpha0 = log10(2.0 * sqrt(wJ * wJ + hJ * hJ) / (wJ * hJ) / s2);
    const double maxThr = isinf(g.precision_px) ? INFINITY : g.precision_px * g.precision_px * s2 * s2;
    const double loge0 = log10(3.0 * (double)(N - 7));
    const uint64_t seed = g.pair_seeds ? g.pair_seeds[p] : g.seed + 1000003ull * p;

    uint32_t nIter = g.max_iter, reserve = nIter / 10;
    nIter -= reserve;
    uint32_t n_pool = N, n_best = 0;
    bool pool_full = true;
    double minNFA = INFINITY;
    float errorMax = INFINITY;
    uint32_t solved_lo = 0, solved_hi = 0;

    for (uint32_t iter = 0; iter < nIter; ++iter) {
        if (iter >= solved_hi) {
            const uint32_t cnt = min((uint32_t)kGeoAhead, nIter - iter);
            __syncthreads();
            if ((uint32_t)tid < cnt) {
                uint32_t pos[7];
                sample7(seed, iter + tid, n_pool, pos);
                double a[14], b[14];
                for (int s = 0; s < 7; ++s) {
                    const uint32_t id = pool_full ? pos[s] : (uint32_t)pool[pos[s]];
                    const double2 u = xI[id], v = xJ[id];
                    a[2 * s] = s1 * u.x + c1x; a[2 * s + 1] = s1 * u.y + c1y;
                    b[2 * s] = s2 * v.x + c2x; b[2 * s + 1] = s2 * v.y + c2y;
                }
                s_nm[tid] = seven_point(a, b, s_models + 27 * tid);
            }
            solved_lo = iter;
            solved_hi = iter + cnt;
            __syncthreads();
        }
        const int nm = s_nm[iter - solved_lo];
        bool better = false;
        for (int m = 0; m < nm; ++m) {
            const double *Fm = s_models + 27 * (iter - solved_lo) + 9 * m;
            const double F0 = Fm[0], F1 = Fm[1], F2 = Fm[2], F3 = Fm[3], F4 = Fm[4], F5 = Fm[5], F6 = Fm[6], F7 = Fm[7],
                         F8 = Fm[8];
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            // residuals (EpipolarDistanceError: squared distance of x2 to the line F x1) + compaction
            const uint32_t n_round = (N + 31u) & ~31u;
            for (uint32_t i = tid; i < n_round; i += kGeoThreads) {
                bool keep = false;
                unsigned long long key = 0;
                if (i < N) {
                    const double2 u = xI[i], v = xJ[i];
                    const double a = s1 * u.x + c1x, b = s1 * u.y + c1y, c = s2 * v.x + c2x, d = s2 * v.y + c2y;
                    const double l0 = fma(F0, a, fma(F1, b, F2)), l1 = fma(F3, a, fma(F4, b, F5)),
                                 l2 = fma(F6, a, fma(F7, b, F8));
                    const double dd = fma(l0, c, fma(l1, d, l2));
                    const float e = (float)(dd * dd / (l0 * l0 + l1 * l1));
                    keep = (e < INFINITY) && ((double)e <= maxThr);   // NaN fails both
                    key = ((unsigned long long)__float_as_uint(e) << 32) | i;
                }
                const unsigned ballot = __ballot_sync(0xffffffffu, keep);
                if (ballot) {
                    const int lane = tid & 31;
                    uint32_t base = 0;
                    if (lane == 0) base = atomicAdd(&s_cnt, (uint32_t)__popc(ballot));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (keep) s_keys[base + __popc(ballot & ((1u << lane) - 1u))] = key;
                }
            }
            __syncthreads();
            const uint32_t M = s_cnt;
            __syncthreads();              // everyone has read the count before it is reset
            if (M < 8) continue;          // no k > 7 within the bound: NFA = +inf, not better
            uint32_t mpad = 32;
            while (mpad < M) mpad <<= 1;
            for (uint32_t i = M + tid; i < mpad; i += kGeoThreads) s_keys[i] = ~0ull;
            __syncthreads();
            for (uint32_t k = 2; k <= mpad; k <<= 1) {
                for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                    for (uint32_t t = tid; t < (mpad >> 1); t += kGeoThreads) {
                        const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                        const uint32_t q = i | j;
                        const unsigned long long x = s_keys[i], y = s_keys[q];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { s_keys[i] = y; s_keys[q] = x; }
                    }
                    __syncthreads();
                }
            }
            NfaMin best{INFINITY, 7};
            for (uint32_t k = 8 + tid; k <= M; k += kGeoThreads) {
                const float e = __uint_as_float((uint32_t)(s_keys[k - 1] >> 32));
                const double logalpha = logalpha0 + 0.5 * log10((double)e + (double)FLT_EPSILON);
                const double nfa = loge0 + logalpha * (double)(k - 7) + (double)logcombi(g.lfact, k, N) +
                                   (double)logcombi(g.lfact, 7, k);
                if (nfa < best.nfa) { best.nfa = nfa; best.k = (int)k; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                NfaMin other;
                other.nfa = __shfl_xor_sync(0xffffffffu, best.nfa, o);
                other.k = __shfl_xor_sync(0xffffffffu, best.k, o);
                best = nfa_min(best, other);
            }
            if ((tid & 31) == 0) s_red[tid >> 5] = best;
            __syncthreads();
            if (tid == 0) {
                NfaMin b = s_red[0];
                for (int w = 1; w < kGeoThreads / 32; ++w) b = nfa_min(b, s_red[w]);
                s_best = b;
            }
            __syncthreads();
            const NfaMin b = s_best;
            if (b.nfa < minNFA) {
                better = true;
                minNFA = b.nfa;
                n_best = (uint32_t)b.k;
                errorMax = __uint_as_float((uint32_t)(s_keys[b.k - 1] >> 32));
                for (uint32_t i = tid; i < n_best; i += kGeoThreads) inl[i] = (int32_t)(uint32_t)s_keys[i];
                if (tid < 9) s_bestF[tid] = Fm[tid];
            }
            __syncthreads();
        }
        if ((better && minNFA < 0) || (iter + 1 == nIter && reserve)) {
            if (n_best == 0) {
                nIter++;
                reserve--;
            } else {
                // the pool is the inlier set in ascending index order (the order of a pool only
                // permutes which uniformly drawn position selects which element): bitmap of the
                // members, block scan of the word counts, ordered expansion
                __syncthreads();
                const uint32_t n_words = (N + 31u) >> 5;
                for (uint32_t wd = tid; wd < n_words; wd += kGeoThreads) s_mask[wd] = 0;
                __syncthreads();
                for (uint32_t i = tid; i < n_best; i += kGeoThreads) {
                    const uint32_t id = (uint32_t)inl[i];
                    atomicOr(&s_mask[id >> 5], 1u << (id & 31));
                }
                __syncthreads();
                const uint32_t per = (n_words + kGeoThreads - 1) / kGeoThreads;   // consecutive words per thread
                const uint32_t w0 = tid * per, w1 = min(n_words, w0 + per);
                uint32_t mine = 0;
                for (uint32_t wd = w0; wd < w1; ++wd) mine += __popc(s_mask[wd]);
                s_scan[tid] = mine;
                __syncthreads();
                for (int o = 1; o < kGeoThreads; o <<= 1) {
                    const uint32_t add = tid >= o ? s_scan[tid - o] : 0;
                    __syncthreads();
                    s_scan[tid] += add;
                    __syncthreads();
                }
                uint32_t pos = s_scan[tid] - mine;
                for (uint32_t wd = w0; wd < w1; ++wd) {
                    uint32_t bits = s_mask[wd];
                    while (bits) {
                        const int b = __ffs(bits) - 1;
                        bits &= bits - 1;
                        pool[pos++] = (int32_t)((wd << 5) + b);
                    }
                }
                n_pool = n_best;
                pool_full = false;
                if (reserve) {
                    nIter = iter + 1 + reserve;
                    reserve = 0;
                }
                solved_hi = iter + 1;    // the problems solved ahead sampled the old pool
                __syncthreads();
            }
        }
    }
    if (!(minNFA < 0)) n_best = 0;
    if (tid == 0) {
        g.nfa[p] = minNFA;
        if (n_best > 0) {
            // Unnormalize: F = N2^T F N1 ; error in pixels = sqrt(e) / N2(0,0)
            const double *B = s_bestF;
            const double N1[9] = {s1, 0, c1x, 0, s1, c1y, 0, 0, 1}, N2[9] = {s2, 0, c2x, 0, s2, c2y, 0, 0, 1};
            double T[9];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) T[3 * i + j] = N2[i] * B[j] + N2[3 + i] * B[3 + j] + N2[6 + i] * B[6 + j];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    g.F[(size_t)p * 9 + 3 * i + j] = T[3 * i] * N1[j] + T[3 * i + 1] * N1[3 + j] + T[3 * i + 2] * N1[6 + j];
            g.err_max[p] = sqrt((double)errorMax) / s2;
            g.n_inl[p] = n_best;
            // GeometricFilter_FMatrix_AC::Robust_estimation: kept iff #inliers > 2.5 * 7
            g.valid[p] = (double)n_best > 2.5 * 7.0 ? 1 : 0;
        }
    }
}

uint32_t geo_pow2(uint32_t v) {
    uint32_t p = 64;
    while (p < v) p <<= 1;
    return p;
}

// log10(n!) for n = 0 .. n_max on the device (grown on demand, kept for the life of the context)
int ensure_lfact(hulo_gpu *h, size_t n_max) {
    if (h->lfact_n > n_max) return HULO_OK;
    const size_t n = std::max<size_t>(n_max + 1, 4096);
    std::vector<double> t(n);
    t[0] = 0.0;
    for (size_t i = 1; i < n; ++i) t[i] = t[i - 1] + log10((double)i);
    HULO_CUDA(h->lfact.reserve(n * sizeof(double)));
    HULO_CUDA(cudaMemcpyAsync(h->lfact.ptr, t.data(), n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    h->lfact_n = n;
    return HULO_OK;
}

}  // namespace
}  // namespace hulo

using namespace hulo;

extern "C" {

int hulo_geometric_filter(hulo_gpu *h, const double *xI, const double *xJ, const uint64_t *pair_offsets,
                          size_t n_pairs, const int32_t *image_sizes, double precision_px, size_t max_iter,
                          uint64_t seed, const uint64_t *pair_seeds, int32_t *valid, uint32_t *n_inliers,
                          int32_t *inliers, double *F, double *error_max, double *nfa) {
    HULO_ARG(h != nullptr, "null context");
    if (n_pairs == 0) return HULO_OK;
    HULO_ARG(pair_offsets != nullptr && image_sizes != nullptr && valid != nullptr && n_inliers != nullptr,
             "null argument");
    HULO_ARG(pair_offsets[0] == 0, "pair_offsets[0] must be 0");
    HULO_ARG(precision_px > 0.0, "precision must be positive (infinity allowed)");
    HULO_ARG(max_iter <= 0xffffffffull / 8, "too many iterations");
    const size_t total = (size_t)pair_offsets[n_pairs];
    HULO_ARG(total == 0 || (xI != nullptr && xJ != nullptr && inliers != nullptr), "null matches");
    size_t n_max = 0;
    for (size_t p = 0; p < n_pairs; ++p) {
        HULO_ARG(pair_offsets[p + 1] >= pair_offsets[p], "pair_offsets must ascend");
        n_max = std::max<size_t>(n_max, (size_t)(pair_offsets[p + 1] - pair_offsets[p]));
        for (int c = 0; c < 4; ++c) HULO_ARG(image_sizes[4 * p + c] > 0, "image sizes must be positive");
    }
    HULO_ARG(n_max <= kGeoMaxMatches, "more than 16384 putative matches in one pair");
    HULO_CUDA(cudaSetDevice(h->device));
    int rc = ensure_lfact(h, n_max);
    if (rc != HULO_OK) return rc;

    // scratch0: xI | xJ | offsets | sizes ; scratch2: per-pair outputs ; scratch3: inliers | pool
    const size_t in_bytes = total * 32 + (n_pairs + 1) * 8 + n_pairs * 8 + n_pairs * 16 + 64;
    HULO_CUDA(h->scratch0.reserve(in_bytes));
    double *d_xI = h->scratch0.as<double>();
    double *d_xJ = d_xI + 2 * total;
    uint64_t *d_off = reinterpret_cast<uint64_t *>(d_xJ + 2 * total);
    uint64_t *d_seeds = d_off + n_pairs + 1;
    int32_t *d_sizes = reinterpret_cast<int32_t *>(d_seeds + n_pairs);
    const size_t out_bytes = n_pairs * (9 + 2) * sizeof(double) + n_pairs * 8 + 64;
    HULO_CUDA(h->scratch2.reserve(out_bytes));
    double *d_F = h->scratch2.as<double>();
    double *d_err = d_F + 9 * n_pairs;
    double *d_nfa = d_err + n_pairs;
    int32_t *d_valid = reinterpret_cast<int32_t *>(d_nfa + n_pairs);
    uint32_t *d_ninl = reinterpret_cast<uint32_t *>(d_valid + n_pairs);
    HULO_CUDA(h->scratch3.reserve(std::max<size_t>(total, 1) * 8));
    int32_t *d_inl = h->scratch3.as<int32_t>();
    int32_t *d_pool = d_inl + total;
    if (total) {
        HULO_CUDA(cudaMemcpyAsync(d_xI, xI, total * 16, cudaMemcpyHostToDevice, h->stream));
        HULO_CUDA(cudaMemcpyAsync(d_xJ, xJ, total * 16, cudaMemcpyHostToDevice, h->stream));
    }
    HULO_CUDA(cudaMemcpyAsync(d_off, pair_offsets, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(cudaMemcpyAsync(d_sizes, image_sizes, n_pairs * 16, cudaMemcpyHostToDevice, h->stream));
    if (pair_seeds) HULO_CUDA(cudaMemcpyAsync(d_seeds, pair_seeds, n_pairs * 8, cudaMemcpyHostToDevice, h->stream));

    GeoParams g;
    g.xI = reinterpret_cast<const double2 *>(d_xI);
    g.xJ = reinterpret_cast<const double2 *>(d_xJ);
    g.off = d_off;
    g.sizes = d_sizes;
    g.lfact = h->lfact.as<double>();
    g.precision_px = precision_px;
    g.max_iter = (uint32_t)max_iter;
    g.seed = seed;
    g.pair_seeds = pair_seeds ? d_seeds : nullptr;
    g.valid = d_valid; g.n_inl = d_ninl; g.F = d_F; g.err_max = d_err; g.nfa = d_nfa;
    g.inl = d_inl; g.pool = d_pool;
    const size_t smem = (size_t)geo_pow2((uint32_t)n_max) * sizeof(unsigned long long);
    if (smem > 32 * 1024 && smem > h->geo_smem_configured) {
        HULO_CUDA(cudaFuncSetAttribute(fmatrix_acransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->geo_smem_configured = smem;
    }
    fmatrix_acransac_kernel<<<(unsigned)n_pairs, kGeoThreads, smem, h->stream>>>(g);
    HULO_CUDA(cudaGetLastError());
    h->launches++;

    HULO_CUDA(cudaMemcpyAsync(valid, d_valid, n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaMemcpyAsync(n_inliers, d_ninl, n_pairs * 4, cudaMemcpyDeviceToHost, h->stream));
    if (F) HULO_CUDA(cudaMemcpyAsync(F, d_F, n_pairs * 72, cudaMemcpyDeviceToHost, h->stream));
    if (error_max) HULO_CUDA(cudaMemcpyAsync(error_max, d_err, n_pairs * 8, cudaMemcpyDeviceToHost, h->stream));
    if (nfa) HULO_CUDA(cudaMemcpyAsync(nfa, d_nfa, n_pairs * 8, cudaMemcpyDeviceToHost, h->stream));
    if (total) HULO_CUDA(cudaMemcpyAsync(inliers, d_inl, total * 4, cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}

}  // extern "C"
