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
// soon as a meaningful one exists, and the iteration budget changes with it); inside the block
// the models of 32 iterations are scored concurrently, one warp each, and committed in order:
//   * the sampler is counter based (splitmix64 at a fixed offset per iteration), so the
//     7-point problems of the next kGeoAhead iterations are solved by that many threads at once
//     (fp64, register resident: null space of the 7x9 system by modified Gram-Schmidt, cubic for
//     det F = 0) and re-solved only when the pool changes;
//   * per model: point-to-epipolar-line residuals of all N matches in fp64, stored as fp32 and
//     packed with the match index into one 64-bit key; only keys within the precision bound
//     are kept (ballot compaction), which makes an outlier-contaminated model cost one pass
//     over the matches and nothing else; survivors are sorted in shared memory (counting sort up
//     to 384 keys, bitonic network above; (residual, index) order = the sequential std::sort) and
//     scanned for
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
#ifndef HULO_GEO_MIN_BLOCKS
#define HULO_GEO_MIN_BLOCKS 4      // 128 registers with a few spills in the solver: four pairs per SM beat two at 254 registers
#endif
constexpr int kGeoAhead = 32;              // iterations whose 7-point problems are solved together
constexpr uint32_t kGeoMaxMatches = 16384; // per pair: 128 KB of sort keys
constexpr uint32_t kGeoRankSortMax = 3 * kGeoThreads;   // counting sort up to here, bitonic network above
constexpr int kGeoWarps = kGeoThreads / 32;
constexpr int kGeoMaxModels = 3 * kGeoAhead;
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

// Two orthonormal vectors spanning the null space of the 7 x 9 epipolar system whose rows are in
// q.  Modified Gram-Schmidt over the rows, then the two largest columns of the projector onto the
// complement.  Every loop is unrolled and every index static (the two column picks go through
// compare-selects), so the 63 + 18 doubles live in registers: the whole solve is a few hundred
// DFMA with short dependency chains instead of a pivoted elimination over thread-local memory,
// which cost ~50 us of latency per batch of samples and dominated the kernel at 25 rounds.
// Any basis of the null space gives the same solution set {F1 + t F2 : det = 0} up to scale.
__device__ __forceinline__ int nullspace2(double (&q)[7][9], double (&f1)[9], double (&f2)[9]) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        double n0 = 0.0;
#pragma unroll
        for (int c = 0; c < 9; ++c) n0 = fma(q[k][c], q[k][c], n0);
#pragma unroll
        for (int m = 0; m < k; ++m) {
            double d = 0.0;
#pragma unroll
            for (int c = 0; c < 9; ++c) d = fma(q[m][c], q[k][c], d);
#pragma unroll
            for (int c = 0; c < 9; ++c) q[k][c] = fma(-d, q[m][c], q[k][c]);
        }
        double n1 = 0.0;
#pragma unroll
        for (int c = 0; c < 9; ++c) n1 = fma(q[k][c], q[k][c], n1);
        if (!(n1 > 1e-24 * n0)) return 0;          // rank deficient sample (also catches NaN)
        const double inv = rsqrt(n1);
#pragma unroll
        for (int c = 0; c < 9; ++c) q[k][c] *= inv;
    }
    double diag[9];                                   // diagonal of P = I - Q^T Q
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        double v = 1.0;
#pragma unroll
        for (int k = 0; k < 7; ++k) v = fma(-q[k][c], q[k][c], v);
        diag[c] = v;
    }
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        double (&f)[9] = pass == 0 ? f1 : f2;
        int a = 0;
        double best = diag[0];
#pragma unroll
        for (int c = 1; c < 9; ++c)
            if (diag[c] > best) { best = diag[c]; a = c; }
        double f1a = 0.0;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            f[c] = c == a ? 1.0 : 0.0;
            if (pass == 1) f1a = c == a ? f1[c] : f1a;
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            double qa = 0.0;
#pragma unroll
            for (int c = 0; c < 9; ++c) qa = c == a ? q[k][c] : qa;
#pragma unroll
            for (int c = 0; c < 9; ++c) f[c] = fma(-qa, q[k][c], f[c]);
        }
        if (pass == 1) {
#pragma unroll
            for (int c = 0; c < 9; ++c) f[c] = fma(-f1a, f1[c], f[c]);
        }
        double n = 0.0;
#pragma unroll
        for (int c = 0; c < 9; ++c) n = fma(f[c], f[c], n);
        if (!(n > 1e-24)) return 0;
        const double inv = rsqrt(n);
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            f[c] *= inv;
            if (pass == 0) diag[c] = fma(-f[c], f[c], diag[c]);
        }
    }
    return 1;
}

// SevenPointSolver: x1, x2 seven normalised points (x, y interleaved); up to three row-major F
// with x2^T F x1 = 0 and det F = 0, ascending in the cubic's root.  Returns the count.
__device__ int seven_point(const double *x1, const double *x2, double *F) {
    double A[7][9];
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        const double a = x1[2 * i], b = x1[2 * i + 1], c = x2[2 * i], d = x2[2 * i + 1];
        A[i][0] = c * a; A[i][1] = c * b; A[i][2] = c;
        A[i][3] = d * a; A[i][4] = d * b; A[i][5] = d;
        A[i][6] = a;     A[i][7] = b;     A[i][8] = 1.0;
    }
    double f1[9], f2[9];
    if (!nullspace2(A, f1, f2)) return 0;
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

// Everything a model is scored against, shared by the two scoring routines.
struct PairCtx {
    const double2 *xI, *xJ;
    uint32_t N;
    double s1, c1x, c1y, s2, c2x, c2y;      // preconditioners of the two images
    double maxThr, logalpha0, loge0;
    const double *lfact;
};

// residual key of match i under model F: fp32 squared distance to the epipolar line in the high
// word, the match index in the low word; ~0 when the residual is NaN/inf or beyond the bound
__device__ __forceinline__ unsigned long long residual_key(const PairCtx &c, const double *F, uint32_t i) {
    const double2 u = c.xI[i], v = c.xJ[i];
    const double a = c.s1 * u.x + c.c1x, b = c.s1 * u.y + c.c1y, x = c.s2 * v.x + c.c2x, y = c.s2 * v.y + c.c2y;
    const double l0 = fma(F[0], a, fma(F[1], b, F[2])), l1 = fma(F[3], a, fma(F[4], b, F[5])),
                 l2 = fma(F[6], a, fma(F[7], b, F[8]));
    const double dd = fma(l0, x, fma(l1, y, l2));
    const float e = (float)(dd * dd / (l0 * l0 + l1 * l1));
    const bool ok = (e < INFINITY) && ((double)e <= c.maxThr);          // NaN fails both
    return ok ? (((unsigned long long)__float_as_uint(e) << 32) | i) : ~0ull;
}

__device__ __forceinline__ double nfa_at(const PairCtx &c, unsigned long long key, uint32_t k) {
    const float e = __uint_as_float((uint32_t)(key >> 32));
    const double logalpha = c.logalpha0 + 0.5 * log10((double)e + (double)FLT_EPSILON);
    return c.loge0 + logalpha * (double)(k - 7) + (double)logcombi(c.lfact, k, c.N) + (double)logcombi(c.lfact, 7, k);
}

struct ModelScore { double nfa; unsigned long long kth_key; int k; };   // k == -1: not scored yet (deferred)

constexpr uint32_t kGeoWarpKeys = 128;      // a warp scores models with up to this many keys within the bound

// One warp scores one model with warp-synchronous code only.  A model with fewer than 8 keys
// within the bound is rejected here (NFA = +inf); up to kGeoWarpKeys keys are sorted by counting
// and scanned; more than that defers the model to the block (k = -1).
__device__ ModelScore score_model_warp(const PairCtx &c, const double *Fm, unsigned long long *keys, int lane) {
    double F[9];
#pragma unroll
    for (int e = 0; e < 9; ++e) F[e] = Fm[e];
    uint32_t M = 0;
    // two independent matches per lane and trip: the fp64 chains of one hide the latency of the other
    for (uint32_t i0 = 0; i0 < c.N; i0 += 64) {
        const uint32_t ia = i0 + lane, ib = i0 + 32 + lane;
        const unsigned long long ka = ia < c.N ? residual_key(c, F, ia) : ~0ull;
        const unsigned long long kb = ib < c.N ? residual_key(c, F, ib) : ~0ull;
        const unsigned ba = __ballot_sync(0xffffffffu, ka != ~0ull), bb = __ballot_sync(0xffffffffu, kb != ~0ull);
        const uint32_t pa = M + __popc(ba & ((1u << lane) - 1u));
        if (ka != ~0ull && pa < kGeoWarpKeys) keys[pa] = ka;
        M += __popc(ba);
        const uint32_t pb = M + __popc(bb & ((1u << lane) - 1u));
        if (kb != ~0ull && pb < kGeoWarpKeys) keys[pb] = kb;
        M += __popc(bb);
    }
    __syncwarp();
    ModelScore out{INFINITY, 0ull, 7};
    if (M < 8) return out;                      // no k > 7 within the bound
    if (M > kGeoWarpKeys) { out.k = -1; return out; }
    unsigned long long mine[4];
    uint32_t rank[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 4; ++e) mine[e] = (uint32_t)(lane + 32 * e) < M ? keys[lane + 32 * e] : ~0ull;
    for (uint32_t j = 0; j < M; ++j) {
        const unsigned long long kj = keys[j];
#pragma unroll
        for (int e = 0; e < 4; ++e) rank[e] += kj < mine[e] ? 1u : 0u;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 4; ++e)
        if ((uint32_t)(lane + 32 * e) < M) keys[rank[e]] = mine[e];
    __syncwarp();
    NfaMin best{INFINITY, 7};
    for (uint32_t k = 8 + lane; k <= M; k += 32) {
        const double nfa = nfa_at(c, keys[k - 1], k);
        if (nfa < best.nfa) { best.nfa = nfa; best.k = (int)k; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        NfaMin other;
        other.nfa = __shfl_xor_sync(0xffffffffu, best.nfa, o);
        other.k = __shfl_xor_sync(0xffffffffu, best.k, o);
        best = nfa_min(best, other);
    }
    out.nfa = best.nfa;
    out.k = best.k;
    out.kth_key = best.nfa < INFINITY ? keys[best.k - 1] : 0ull;
    __syncwarp();                               // every lane has read before the buffer is reused
    return out;
}

// The whole block compacts the keys of model F that are <= limit into keys[] and sorts them
// ((residual, index) order): counting sort up to kGeoRankSortMax keys, bitonic network above.
// Returns the number of keys.  Called by all threads; ends with the keys visible to everyone.
__device__ uint32_t block_sorted_keys(const PairCtx &c, const double *F, unsigned long long limit,
                                      unsigned long long *keys, uint32_t *s_cnt, int tid) {
    const int lane = tid & 31;
    __syncthreads();
    if (tid == 0) *s_cnt = 0;
    __syncthreads();
    const uint32_t n_round = (c.N + 31u) & ~31u;
    for (uint32_t i = tid; i < n_round; i += kGeoThreads) {
        const unsigned long long key = i < c.N ? residual_key(c, F, i) : ~0ull;
        const bool keep = key <= limit && key != ~0ull;
        const unsigned ballot = __ballot_sync(0xffffffffu, keep);
        if (ballot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(s_cnt, (uint32_t)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) keys[base + __popc(ballot & ((1u << lane) - 1u))] = key;
        }
    }
    __syncthreads();
    const uint32_t M = *s_cnt;
    if (M <= kGeoRankSortMax) {
        unsigned long long mine[3];
        uint32_t rank[3] = {0, 0, 0};
#pragma unroll
        for (int e = 0; e < 3; ++e) mine[e] = tid + e * kGeoThreads < (int)M ? keys[tid + e * kGeoThreads] : ~0ull;
        for (uint32_t j = 0; j < M; ++j) {
            const unsigned long long kj = keys[j];
#pragma unroll
            for (int e = 0; e < 3; ++e) rank[e] += kj < mine[e] ? 1u : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 3; ++e)
            if (tid + e * kGeoThreads < (int)M) keys[rank[e]] = mine[e];
        __syncthreads();
    } else {
        uint32_t mpad = 512;
        while (mpad < M) mpad <<= 1;
        for (uint32_t i = M + tid; i < mpad; i += kGeoThreads) keys[i] = ~0ull;
        __syncthreads();
        for (uint32_t k = 2; k <= mpad; k <<= 1) {
            for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t t = tid; t < (mpad >> 1); t += kGeoThreads) {
                    const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const uint32_t q = i | j;
                    const unsigned long long x = keys[i], y = keys[q];
                    const bool up = (i & k) == 0;
                    if ((x > y) == up) { keys[i] = y; keys[q] = x; }
                }
                __syncthreads();
            }
        }
    }
    return M;
}

// The whole block scores one model (the ones a warp deferred: more than kGeoWarpKeys keys).
__device__ ModelScore score_model_block(const PairCtx &c, const double *F, unsigned long long *keys, uint32_t *s_cnt,
                                        NfaMin *s_red, ModelScore *s_out, int tid) {
    const uint32_t M = block_sorted_keys(c, F, ~0ull - 1ull, keys, s_cnt, tid);
    NfaMin best{INFINITY, 7};
    for (uint32_t k = 8 + tid; k <= M; k += kGeoThreads) {
        const double nfa = nfa_at(c, keys[k - 1], k);
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
        for (int w = 1; w < kGeoWarps; ++w) b = nfa_min(b, s_red[w]);
        s_out->nfa = b.nfa;
        s_out->k = b.k;
        s_out->kth_key = b.nfa < INFINITY ? keys[b.k - 1] : 0ull;
    }
    __syncthreads();
    return *s_out;
}

// One block runs the whole AC-RANSAC of one image pair (see the header of this file).  The up to 96
// models of a batch of 32 pre-solved iterations are scored CONCURRENTLY, one warp per model: the
// ~99 % that an outlier contaminates die after one pass over the matches, small inlier sets are
// sorted and scanned inside the warp, and only models with more than 128 keys within the bound are
// left to the block.  The results are then committed in iteration order by the sequential rule;
// deferred models are scored by the whole block when their turn comes (never, if a commit has
// replaced the pool first); models scored against a pool that a commit replaces are discarded and
// their iterations re-solved, exactly where the sequential algorithm would have drawn them from
// the new pool.  While iterating only the SET of inliers of the best model is needed (the pool is
// kept in index order), so it is rebuilt from the model when a commit needs it, and the
// residual-ordered inlier list is produced once at the end.
__global__ void __launch_bounds__(kGeoThreads, HULO_GEO_MIN_BLOCKS) fmatrix_acransac_kernel(GeoParams g) {
    extern __shared__ unsigned long long s_keys[];          // block buffer: pow2 >= N keys (>= 512)
    __shared__ unsigned long long s_wkeys[kGeoWarps][kGeoWarpKeys];
    __shared__ double s_models[kGeoAhead * 27];
    __shared__ int s_nm[kGeoAhead];
    __shared__ int s_first[kGeoAhead + 1];
    __shared__ double s_res_nfa[kGeoMaxModels];
    __shared__ unsigned long long s_res_key[kGeoMaxModels];
    __shared__ int s_res_k[kGeoMaxModels];
    __shared__ double s_bestF[9];
    __shared__ NfaMin s_red[kGeoWarps];
    __shared__ ModelScore s_out;
    __shared__ uint32_t s_cnt;
    __shared__ uint32_t s_mask[kGeoMaxMatches / 32];        // membership of the narrowed pool
    __shared__ uint32_t s_scan[kGeoThreads];

    const uint32_t p = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

    int32_t *inl = g.inl + off, *pool = g.pool + off;
    PairCtx c;
    c.xI = g.xI + off;
    c.xJ = g.xJ + off;
    c.N = N;
    c.lfact = g.lfact;
    {   // ACKernelAdaptor: both images preconditioned by T = [[s,0,-w s/2],[0,s,-h s/2],[0,0,1]], s = 1/sqrt(w h)
        const double wI = g.sizes[4 * p], hI = g.sizes[4 * p + 1], wJ = g.sizes[4 * p + 2], hJ = g.sizes[4 * p + 3];
        c.s1 = 1.0 / sqrt(wI * hI); c.c1x = -0.5 * wI * c.s1; c.c1y = -0.5 * hI * c.s1;
        c.s2 = 1.0 / sqrt(wJ * hJ); c.c2x = -0.5 * wJ * c.s2; c.c2y = -0.5 * hJ * c.s2;
        c.logalpha0 = log10(2.0 * sqrt(wJ * wJ + hJ * hJ) / (wJ * hJ) / c.s2);
        c.maxThr = isinf(g.precision_px) ? INFINITY : g.precision_px * g.precision_px * c.s2 * c.s2;
        c.loge0 = log10(3.0 * (double)(N - 7));
    }
    const uint64_t seed = g.pair_seeds ? g.pair_seeds[p] : g.seed + 1000003ull * p;

    uint32_t nIter = g.max_iter, reserve = nIter / 10;
    nIter -= reserve;
    uint32_t n_pool = N, n_best = 0;
    bool pool_full = true;
    double minNFA = INFINITY;
    unsigned long long best_key = 0ull;      // k-th smallest key of the best model: its inliers are the keys <= it

    uint32_t iter = 0;
    while (iter < nIter) {
        const uint32_t cnt = min((uint32_t)kGeoAhead, nIter - iter);
        // ---- the 7-point problems of the next cnt iterations, one thread each
        __syncthreads();
        if ((uint32_t)tid < cnt) {
            uint32_t pos[7];
            sample7(seed, iter + tid, n_pool, pos);
            double a[14], b[14];
#pragma unroll
            for (int s = 0; s < 7; ++s) {
                const uint32_t id = pool_full ? pos[s] : (uint32_t)pool[pos[s]];
                const double2 u = c.xI[id], v = c.xJ[id];
                a[2 * s] = c.s1 * u.x + c.c1x; a[2 * s + 1] = c.s1 * u.y + c.c1y;
                b[2 * s] = c.s2 * v.x + c.c2x; b[2 * s + 1] = c.s2 * v.y + c.c2y;
            }
            s_nm[tid] = seven_point(a, b, s_models + 27 * tid);
        }
        __syncthreads();
        if (tid == 0) {
            int acc = 0;
            for (uint32_t t = 0; t < cnt; ++t) { s_first[t] = acc; acc += s_nm[t]; }
            s_first[cnt] = acc;
        }
        __syncthreads();
        const int G = s_first[cnt];
        // ---- score the models of the batch, one warp each -- while sampling from all matches, where
        // nearly every model is contaminated.  Once the pool is the inlier set nearly every model is
        // a good one with many keys: those go straight to the block, in order, one at a time.
        if (!pool_full) {
            for (int gi = tid; gi < G; gi += kGeoThreads) s_res_k[gi] = -1;
        } else
        for (int gi = warp; gi < G; gi += kGeoWarps) {
            int t = 0;
            while (s_first[t + 1] <= gi) ++t;           // model gi belongs to iteration iter + t
            const ModelScore r = score_model_warp(c, s_models + 27 * t + 9 * (gi - s_first[t]), s_wkeys[warp], lane);
            if (lane == 0) { s_res_nfa[gi] = r.nfa; s_res_key[gi] = r.kth_key; s_res_k[gi] = r.k; }
        }
        __syncthreads();
        // ---- commit in iteration order (every thread runs the same scalar logic)
        uint32_t next_iter = iter + cnt;
        for (uint32_t t = 0; t < cnt; ++t) {
            const uint32_t it = iter + t;
            bool better = false;
            int best_g = -1;
            for (int gi = s_first[t]; gi < s_first[t + 1]; ++gi) {
                ModelScore r{s_res_nfa[gi], s_res_key[gi], s_res_k[gi]};
                if (r.k < 0) r = score_model_block(c, s_models + 27 * t + 9 * (gi - s_first[t]), s_keys, &s_cnt, s_red, &s_out, tid);
                if (r.nfa < minNFA) {
                    better = true;
                    minNFA = r.nfa;
                    n_best = (uint32_t)r.k;
                    best_key = r.kth_key;
                    best_g = gi;
                }
            }
            if (better) {
                __syncthreads();
                if (tid < 9) s_bestF[tid] = s_models[27 * t + 9 * (best_g - s_first[t]) + tid];
                __syncthreads();
            }
            if ((better && minNFA < 0) || (it + 1 == nIter && reserve)) {
                if (n_best == 0) {
                    nIter++;                             // nothing scored yet: one more global draw
                    reserve--;
                } else {
                    // pool := inliers of the best model, ascending index (bitmap, block scan, expansion)
                    const uint32_t n_words = (N + 31u) >> 5;
                    for (uint32_t wd = tid; wd < n_words; wd += kGeoThreads) s_mask[wd] = 0;
                    __syncthreads();
                    for (uint32_t i = tid; i < N; i += kGeoThreads)
                        if (residual_key(c, s_bestF, i) <= best_key) atomicOr(&s_mask[i >> 5], 1u << (i & 31));
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
                        nIter = it + 1 + reserve;
                        reserve = 0;
                    }
                    next_iter = it + 1;                  // the rest of the batch sampled the old pool: discard
                    __syncthreads();
                    break;
                }
            }
        }
        iter = next_iter;
    }
    if (!(minNFA < 0)) n_best = 0;
    if (n_best > 0) {
        // residual-ordered inlier list of the final model: its keys <= best_key, sorted by the block
        const uint32_t M = block_sorted_keys(c, s_bestF, best_key, s_keys, &s_cnt, tid);    // == n_best
        for (uint32_t i = tid; i < M; i += kGeoThreads) inl[i] = (int32_t)(uint32_t)s_keys[i];
    }
    if (tid == 0) {
        g.nfa[p] = minNFA;
        if (n_best > 0) {
            // Unnormalize: F = N2^T F N1 ; error in pixels = sqrt(e) / N2(0,0)
            const double *B = s_bestF;
            const double N1[9] = {c.s1, 0, c.c1x, 0, c.s1, c.c1y, 0, 0, 1}, N2[9] = {c.s2, 0, c.c2x, 0, c.s2, c.c2y, 0, 0, 1};
            double T[9];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) T[3 * i + j] = N2[i] * B[j] + N2[3 + i] * B[3 + j] + N2[6 + i] * B[6 + j];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    g.F[(size_t)p * 9 + 3 * i + j] = T[3 * i] * N1[j] + T[3 * i + 1] * N1[3 + j] + T[3 * i + 2] * N1[6 + j];
            const float errorMax = __uint_as_float((uint32_t)(best_key >> 32));
            g.err_max[p] = sqrt((double)errorMax) / c.s2;
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
    // A small call (the pairs of one query) is latency: its inputs are packed in pinned memory in the
    // device layout and go up in one copy, its outputs come back in two.  A large one (a
    // reconstruction) copies straight from and to the caller's arrays.
    const bool packed = in_bytes <= (1u << 20);
    if (packed) {
        HULO_CUDA(h->hstage1.reserve(in_bytes));
        uint8_t *hp = h->hstage1.as<uint8_t>();
        if (total) {
            memcpy(hp, xI, total * 16);
            memcpy(hp + total * 16, xJ, total * 16);
        }
        memcpy(hp + total * 32, pair_offsets, (n_pairs + 1) * 8);
        if (pair_seeds) memcpy(hp + total * 32 + (n_pairs + 1) * 8, pair_seeds, n_pairs * 8);
        memcpy(hp + total * 32 + (n_pairs + 1) * 8 + n_pairs * 8, image_sizes, n_pairs * 16);
        HULO_CUDA(cudaMemcpyAsync(d_xI, hp, in_bytes - 64, cudaMemcpyHostToDevice, h->stream));
    } else {
        if (total) {
            HULO_CUDA(cudaMemcpyAsync(d_xI, xI, total * 16, cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemcpyAsync(d_xJ, xJ, total * 16, cudaMemcpyHostToDevice, h->stream));
        }
        HULO_CUDA(cudaMemcpyAsync(d_off, pair_offsets, (n_pairs + 1) * 8, cudaMemcpyHostToDevice, h->stream));
        HULO_CUDA(cudaMemcpyAsync(d_sizes, image_sizes, n_pairs * 16, cudaMemcpyHostToDevice, h->stream));
        if (pair_seeds) HULO_CUDA(cudaMemcpyAsync(d_seeds, pair_seeds, n_pairs * 8, cudaMemcpyHostToDevice, h->stream));
    }

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
    const size_t smem = (size_t)std::max<uint32_t>(geo_pow2((uint32_t)n_max), 512) * sizeof(unsigned long long);
    if (smem > 16 * 1024 && smem > h->geo_smem_configured) {
        HULO_CUDA(cudaFuncSetAttribute(fmatrix_acransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->geo_smem_configured = smem;
    }
    fmatrix_acransac_kernel<<<(unsigned)n_pairs, kGeoThreads, smem, h->stream>>>(g);
    HULO_CUDA(cudaGetLastError());
    h->launches++;

    if (packed) {
        const size_t per_pair = n_pairs * (9 + 2) * sizeof(double) + n_pairs * 8;
        HULO_CUDA(h->hstage0.reserve(per_pair + total * 4 + 64));
        uint8_t *ho = h->hstage0.as<uint8_t>();
        HULO_CUDA(cudaMemcpyAsync(ho, d_F, per_pair, cudaMemcpyDeviceToHost, h->stream));
        if (total) HULO_CUDA(cudaMemcpyAsync(ho + per_pair, d_inl, total * 4, cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaStreamSynchronize(h->stream));
        const uint8_t *o = ho;
        if (F) memcpy(F, o, n_pairs * 72);
        o += n_pairs * 72;
        if (error_max) memcpy(error_max, o, n_pairs * 8);
        o += n_pairs * 8;
        if (nfa) memcpy(nfa, o, n_pairs * 8);
        o += n_pairs * 8;
        memcpy(valid, o, n_pairs * 4);
        o += n_pairs * 4;
        memcpy(n_inliers, o, n_pairs * 4);
        if (total) memcpy(inliers, ho + per_pair, total * 4);
        return HULO_OK;
    }
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
