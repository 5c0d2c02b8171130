// resect.cu -- K2: batched scoring of resection-RANSAC hypotheses, the P3P minimal solver,
// and the batched AC-RANSAC driver, sm_100a.
//
// Stands behind openMVG::sfm::SfM_Localizer::Localize as the reference calls it
// (VisionLocalizeServer/src/LocalizeEngine.cc:503-531, OpenMVGLocalization_AKAZE/src/
// localization.cpp:479-509, OpenMVG_BA/src/adjust_sfm_data.cpp:109-137): AC-RANSAC over the
// P3P kernel, K^-1-normalised squared reprojection residuals, a-contrario NFA, max 4096
// iterations.  OpenMVG 1.1 is third-party and not vendored; SURVEY.md appendix B records the
// published algorithm this follows.
//
// One thread block scores one hypothesis against all N correspondences:
//   projection [R|t] X and the subtraction from the observation in fp64 (12 DFMA per point:
//   keeps the residual within 1e-4 px at f ~ 1860 px, which plain fp32 cannot guarantee),
//   squared residual stored as fp32, ascending bitonic sort in shared memory (the NFA needs the
//   order statistics, so a histogram would not be exact), then the NFA scan
//   nfa_k = loge0 + (logalpha0 + log10(e_k + FLT_EPSILON)) (k-3) + logC(N,k) + logC(k,3)
//   with a block-wide lexicographic (nfa, k) minimum so the first minimum wins like the
//   sequential scan.
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "context.cuh"

namespace hulo {
namespace {

constexpr int kScoreThreads = 256;
constexpr uint32_t kMaxPoints = 32768;

struct Mat3 { double m[9]; };

// ------------------------------------------------------------------ device math
__device__ __forceinline__ void proj_residual(const double *__restrict__ M, const double *__restrict__ X,
                                              const double *__restrict__ x, double &dx, double &dy) {
    const double u = fma(M[0], X[0], fma(M[1], X[1], fma(M[2], X[2], M[3])));
    const double v = fma(M[4], X[0], fma(M[5], X[1], fma(M[6], X[2], M[7])));
    const double w = fma(M[8], X[0], fma(M[9], X[1], fma(M[10], X[2], M[11])));
    dx = u / w - x[0];
    dy = v / w - x[1];
}

struct NfaMin { double nfa; int k; };
__device__ __forceinline__ NfaMin nfa_min(NfaMin a, NfaMin b) {
    if (b.nfa < a.nfa || (b.nfa == a.nfa && b.k < a.k)) return b;
    return a;
}

// One CTA per hypothesis.  smem: npad floats.
__global__ void __launch_bounds__(kScoreThreads) score_kernel(
    const double *__restrict__ models, uint32_t H, const double *__restrict__ x2dn,
    const double *__restrict__ X3d, uint32_t N, uint32_t npad, const float *__restrict__ logc_n,
    const float *__restrict__ logc_k, double loge0, double logalpha0, float thr2, double *__restrict__ out_nfa,
    int32_t *__restrict__ out_k, float *__restrict__ out_errk, int32_t *__restrict__ out_ninl) {
    extern __shared__ float s_e[];
    __shared__ double s_M[12];
    __shared__ NfaMin s_red[kScoreThreads / 32];
    __shared__ int s_cnt[kScoreThreads / 32];
    const uint32_t h = blockIdx.x;
    if (h >= H) return;
    const int tid = threadIdx.x;
    if (tid < 12) s_M[tid] = models[(size_t)h * 12 + tid];
    __syncthreads();
    // a solver slot without a model is marked by a NaN in its first entry
    if (!(s_M[0] == s_M[0]) || N < 4) {
        if (tid == 0) {
            out_nfa[h] = INFINITY;
            if (out_k) out_k[h] = 3;
            if (out_errk) out_errk[h] = INFINITY;
            if (out_ninl) out_ninl[h] = 0;
        }
        return;
    }
    int cnt = 0;
    for (uint32_t i = tid; i < npad; i += kScoreThreads) {
        float e = INFINITY;
        if (i < N) {
            double dx, dy;
            proj_residual(s_M, X3d + 3 * (size_t)i, x2dn + 2 * (size_t)i, dx, dy);
            const float fx = (float)dx, fy = (float)dy;
            e = fmaf(fx, fx, fy * fy);
            if (!(e == e)) e = INFINITY;          // NaN residual (point on the principal plane) sorts last
            if (thr2 >= 0.0f && e <= thr2) ++cnt;
        }
        s_e[i] = e;
    }
    __syncthreads();
    // bitonic sort, ascending
    for (uint32_t k = 2; k <= npad; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t t = tid; t < (npad >> 1); t += kScoreThreads) {
                const uint32_t i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const uint32_t p = i | j;
                const float a = s_e[i], b = s_e[p];
                const bool up = (i & k) == 0;
                if ((a > b) == up) { s_e[i] = b; s_e[p] = a; }
            }
            __syncthreads();
        }
    }
    // NFA scan over k = 4 .. N
    NfaMin best{INFINITY, 3};
    for (uint32_t k = 4 + tid; k <= N; k += kScoreThreads) {
        const float e = s_e[k - 1];
        if (e == INFINITY) continue;
        const double logalpha = logalpha0 + log10((double)e + (double)FLT_EPSILON);
        const double nfa = loge0 + logalpha * (double)(k - 3) + (double)logc_n[k] + (double)logc_k[k];
        if (nfa < best.nfa) { best.nfa = nfa; best.k = (int)k; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        NfaMin other;
        other.nfa = __shfl_xor_sync(0xffffffffu, best.nfa, o);
        other.k = __shfl_xor_sync(0xffffffffu, best.k, o);
        best = nfa_min(best, other);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((tid & 31) == 0) { s_red[tid >> 5] = best; s_cnt[tid >> 5] = cnt; }
    __syncthreads();
    if (tid == 0) {
        NfaMin b = s_red[0];
        int c = s_cnt[0];
        for (int w = 1; w < kScoreThreads / 32; ++w) { b = nfa_min(b, s_red[w]); c += s_cnt[w]; }
        out_nfa[h] = b.nfa;
        if (out_k) out_k[h] = b.k;
        if (out_errk) out_errk[h] = b.k >= 1 ? s_e[b.k - 1] : INFINITY;
        if (out_ninl) out_ninl[h] = c;
    }
}

// The same kernel with the sort in REGISTERS (N <= 4096).  Thread t owns the E = npad / 256
// consecutive elements [E t, E t + E) of the bitonic network: exchanges at distance j < E stay inside
// the thread, E <= j < 32 E are warp shuffles, and only j >= 32 E (at most three sub-stages of the
// last three merges) go through shared memory -- 6 block barriers instead of one per sub-stage (66
// for 2048 elements), and no strided shared-memory traffic (the ncu profile of score_kernel:
// 73 % issue slots, LSU 50 %, 80 M bank conflicts, warps waiting on short scoreboard).
// The block's work for hypothesis h, shared by the single-problem kernel and the batched one (so a
// problem scores bit-identically through either).
template <int E>
__device__ __forceinline__ void score_reg_block(
    const uint32_t h, const double *__restrict__ models, const double *__restrict__ x2dn,
    const double *__restrict__ X3d, uint32_t N, const float *__restrict__ logc_n,
    const float *__restrict__ logc_k, double loge0, double logalpha0, float thr2, double *__restrict__ out_nfa,
    int32_t *__restrict__ out_k, float *__restrict__ out_errk, int32_t *__restrict__ out_ninl) {
    constexpr uint32_t npad = (uint32_t)E * kScoreThreads;
    // one pad word per 32 keeps the blocked accesses (stride E between lanes) free of bank conflicts
    __shared__ float s_e[npad + npad / 32];
    auto at = [](uint32_t i) { return i + (i >> 5); };
    __shared__ double s_M[12];
    __shared__ NfaMin s_red[kScoreThreads / 32];
    __shared__ int s_cnt[kScoreThreads / 32];
    const int tid = threadIdx.x;
    if (tid < 12) s_M[tid] = models[(size_t)h * 12 + tid];
    __syncthreads();
    if (!(s_M[0] == s_M[0]) || N < 4) {
        if (tid == 0) {
            out_nfa[h] = INFINITY;
            if (out_k) out_k[h] = 3;
            if (out_errk) out_errk[h] = INFINITY;
            if (out_ninl) out_ninl[h] = 0;
        }
        return;
    }
    int cnt = 0;
    for (uint32_t i = tid; i < npad; i += kScoreThreads) {
        float e = INFINITY;
        if (i < N) {
            double dx, dy;
            proj_residual(s_M, X3d + 3 * (size_t)i, x2dn + 2 * (size_t)i, dx, dy);
            const float fx = (float)dx, fy = (float)dy;
            e = fmaf(fx, fx, fy * fy);
            if (!(e == e)) e = INFINITY;
            if (thr2 >= 0.0f && e <= thr2) ++cnt;
        }
        s_e[at(i)] = e;
    }
    __syncthreads();
    float v[E];
    const uint32_t base = (uint32_t)E * tid;
#pragma unroll
    for (int a = 0; a < E; ++a) v[a] = s_e[at(base + a)];
    for (uint32_t k = 2; k <= npad; k <<= 1) {
        uint32_t j = k >> 1;
        for (; j >= 32u * E; j >>= 1) {                     // partner in another warp
            __syncthreads();
#pragma unroll
            for (int a = 0; a < E; ++a) s_e[at(base + a)] = v[a];
            __syncthreads();
#pragma unroll
            for (int a = 0; a < E; ++a) {
                const uint32_t i = base + a;
                const float o = s_e[at(i ^ j)];
                const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                v[a] = keep_min ? fminf(v[a], o) : fmaxf(v[a], o);
            }
        }
        for (; j >= (uint32_t)E; j >>= 1) {                 // partner in another lane of the warp
            const int m = (int)(j / E);
#pragma unroll
            for (int a = 0; a < E; ++a) {
                const uint32_t i = base + a;
                const float o = __shfl_xor_sync(0xffffffffu, v[a], m);
                const bool keep_min = ((i & j) == 0) == ((i & k) == 0);
                v[a] = keep_min ? fminf(v[a], o) : fmaxf(v[a], o);
            }
        }
#pragma unroll
        for (int jj = E / 2; jj > 0; jj >>= 1) {            // partner inside the thread
            if ((uint32_t)jj < k) {
#pragma unroll
                for (int a = 0; a < E; ++a) {
                    if ((a & jj) == 0) {
                        const bool up = ((base + a) & k) == 0;
                        const float lo = fminf(v[a], v[a | jj]), hi = fmaxf(v[a], v[a | jj]);
                        v[a] = up ? lo : hi;
                        v[a | jj] = up ? hi : lo;
                    }
                }
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < E; ++a) s_e[at(base + a)] = v[a];
    __syncthreads();
    NfaMin best{INFINITY, 3};
    for (uint32_t k = 4 + tid; k <= N; k += kScoreThreads) {
        const float e = s_e[at(k - 1)];
        if (e == INFINITY) continue;
        const double logalpha = logalpha0 + log10((double)e + (double)FLT_EPSILON);
        const double nfa = loge0 + logalpha * (double)(k - 3) + (double)logc_n[k] + (double)logc_k[k];
        if (nfa < best.nfa) { best.nfa = nfa; best.k = (int)k; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        NfaMin other;
        other.nfa = __shfl_xor_sync(0xffffffffu, best.nfa, o);
        other.k = __shfl_xor_sync(0xffffffffu, best.k, o);
        best = nfa_min(best, other);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((tid & 31) == 0) { s_red[tid >> 5] = best; s_cnt[tid >> 5] = cnt; }
    __syncthreads();
    if (tid == 0) {
        NfaMin b = s_red[0];
        int c = s_cnt[0];
        for (int w = 1; w < kScoreThreads / 32; ++w) { b = nfa_min(b, s_red[w]); c += s_cnt[w]; }
        out_nfa[h] = b.nfa;
        if (out_k) out_k[h] = b.k;
        if (out_errk) out_errk[h] = b.k >= 1 ? s_e[at((uint32_t)b.k - 1)] : INFINITY;
        if (out_ninl) out_ninl[h] = c;
    }
}

template <int E>
__global__ void __launch_bounds__(kScoreThreads) score_kernel_reg(
    const double *__restrict__ models, uint32_t H, const double *__restrict__ x2dn,
    const double *__restrict__ X3d, uint32_t N, const float *__restrict__ logc_n,
    const float *__restrict__ logc_k, double loge0, double logalpha0, float thr2, double *__restrict__ out_nfa,
    int32_t *__restrict__ out_k, float *__restrict__ out_errk, int32_t *__restrict__ out_ninl) {
    if (blockIdx.x >= H) return;
    score_reg_block<E>(blockIdx.x, models, x2dn, X3d, N, logc_n, logc_k, loge0, logalpha0, thr2, out_nfa, out_k,
                       out_errk, out_ninl);
}

// ---- many resection problems in one launch (hulo_resect_acransac_batch)
// One staged problem: where its correspondences and log-binomial tables live in the arena.
struct ResectDesc {
    const double *x2dn, *X3d;
    const float *logc_n, *logc_k;
    double loge0, logalpha0;
    uint32_t N, pad;
};
// One problem's share of a wave: its hypotheses are models[hyp_base .. hyp_base + 4 T) and, inside
// the launch of its size class, the blocks [blk_off, next slot's blk_off).
struct WaveSlot { uint32_t job, hyp_base, blk_off, pad; };

template <int E>
__global__ void __launch_bounds__(kScoreThreads) score_kernel_reg_batch(
    const ResectDesc *__restrict__ desc, const WaveSlot *__restrict__ slots, uint32_t n_slots,
    const double *__restrict__ models, double *__restrict__ out_nfa) {
    // slot of this block: last one with blk_off <= blockIdx.x (uniform over the block, cached)
    uint32_t lo = 0, hi = n_slots;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (slots[mid].blk_off <= blockIdx.x) lo = mid; else hi = mid;
    }
    const WaveSlot sl = slots[lo];
    const ResectDesc d = desc[sl.job];
    score_reg_block<E>(sl.hyp_base + (blockIdx.x - sl.blk_off), models, d.x2dn, d.X3d, d.N, d.logc_n, d.logc_k,
                       d.loge0, d.logalpha0, -1.0f, out_nfa, nullptr, nullptr, nullptr);
}

// First minimum of the H scores (strict <: the earliest hypothesis wins ties, like the sequential
// update rule) and its model, gathered into one small record so a batch costs one D2H copy:
// out = {nfa, (double)index, model[12]}; index = -1 when every score is +inf / NaN.
// kFresh: the inputs were written by other blocks of the same launch (read them past L1).
template <bool kFresh = false>
__device__ __forceinline__ void argmin_block(const double *nfa, uint32_t H, const double *models, double *out) {
    __shared__ double s_v[8];
    __shared__ uint32_t s_i[8];
    double v = INFINITY;
    uint32_t idx = 0xFFFFFFFFu;
    for (uint32_t h = threadIdx.x; h < H; h += 256) {
        const double x = kFresh ? __ldcg(nfa + h) : nfa[h];
        if (x < v) { v = x; idx = h; }          // ascending h per thread: keeps the earliest
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) { v = ov; idx = oi; }
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = v; s_i[threadIdx.x >> 5] = idx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w)
            if (s_v[w] < v || (s_v[w] == v && s_i[w] < idx)) { v = s_v[w]; idx = s_i[w]; }
        out[0] = v;
        out[1] = idx == 0xFFFFFFFFu ? -1.0 : (double)idx;
        for (int k = 0; k < 12; ++k)
            out[2 + k] = idx == 0xFFFFFFFFu ? 0.0 : (kFresh ? __ldcg(models + (size_t)idx * 12 + k) : models[(size_t)idx * 12 + k]);
    }
}
__global__ void __launch_bounds__(256) argmin_kernel(const double *__restrict__ nfa, uint32_t H,
                                                     const double *__restrict__ models, double *__restrict__ out) {
    argmin_block(nfa, H, models, out);
}
// one block per problem of the wave: record a = first minimum over its own hypotheses
__global__ void __launch_bounds__(256) argmin_batch_kernel(const double *__restrict__ nfa,
                                                           const uint2 *__restrict__ ranges,
                                                           const double *__restrict__ models,
                                                           double *__restrict__ out) {
    const uint2 r = ranges[blockIdx.x];
    argmin_block(nfa + r.x, r.y, models + (size_t)r.x * 12, out + (size_t)blockIdx.x * 14);
}

// residuals in pixels (sqrt(e) * fx), H x N, for the parity test of the projection arithmetic
__global__ void residual_kernel(const double *__restrict__ models, uint32_t H, const double *__restrict__ x2dn,
                                const double *__restrict__ X3d, uint32_t N, float fx, float *__restrict__ res) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t h = blockIdx.y;
    if (i >= N || h >= H) return;
    double dx, dy;
    proj_residual(models + (size_t)h * 12, X3d + 3 * (size_t)i, x2dn + 2 * (size_t)i, dx, dy);
    const float fxx = (float)dx, fyy = (float)dy;
    res[(size_t)h * N + i] = sqrtf(fmaf(fxx, fxx, fyy * fyy)) * fx;
}

// ------------------------------------------------------------------ P3P (Kneip, CVPR 2011)
struct Cplx { double re, im; };
__device__ __forceinline__ Cplx c_add(Cplx a, Cplx b) { return {a.re + b.re, a.im + b.im}; }
__device__ __forceinline__ Cplx c_sub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx c_mul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cplx c_scale(Cplx a, double s) { return {a.re * s, a.im * s}; }
__device__ __forceinline__ Cplx c_div(Cplx a, Cplx b) {
    const double d = b.re * b.re + b.im * b.im;
    return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
__device__ __forceinline__ Cplx c_sqrt(Cplx z) {   // principal branch
    const double r = hypot(z.re, z.im);
    if (r == 0.0) return {0.0, 0.0};
    double a = sqrt(0.5 * (r + fabs(z.re)));
    double b = 0.5 * z.im / a;
    if (z.re >= 0.0) return {a, b};
    return {fabs(b), copysign(a, z.im)};
}
__device__ __forceinline__ Cplx c_cbrt(Cplx z) {   // principal value of z^(1/3)
    const double r = hypot(z.re, z.im);
    if (r == 0.0) return {0.0, 0.0};
    const double th = atan2(z.im, z.re) / 3.0;
    const double m = cbrt(r);
    double s, c;
    sincos(th, &s, &c);
    return {m * c, m * s};
}

__device__ void solve_quartic(const double f[5], double roots[4]) {
    const double A = f[0], B = f[1], C = f[2], D = f[3], E = f[4];
    const double A2 = A * A, B2 = B * B, A3 = A2 * A, B3 = B2 * B, A4 = A3 * A, B4 = B3 * B;
    const double alpha = -3.0 * B2 / (8.0 * A2) + C / A;
    const double beta = B3 / (8.0 * A3) - B * C / (2.0 * A2) + D / A;
    const double gamma = -3.0 * B4 / (256.0 * A4) + B2 * C / (16.0 * A3) - B * D / (4.0 * A2) + E / A;
    const double alpha2 = alpha * alpha, alpha3 = alpha2 * alpha;
    const Cplx P{-alpha2 / 12.0 - gamma, 0.0};
    const Cplx Q{-alpha3 / 108.0 + alpha * gamma / 3.0 - beta * beta / 8.0, 0.0};
    const Cplx disc = c_add(c_scale(c_mul(Q, Q), 0.25), c_scale(c_mul(c_mul(P, P), P), 1.0 / 27.0));
    const Cplx R = c_add(c_scale(Q, -0.5), c_sqrt(disc));
    const Cplx U = c_cbrt(R);
    Cplx y;
    if (U.re == 0.0) y = c_sub(Cplx{-5.0 * alpha / 6.0, 0.0}, c_cbrt(Q));
    else y = c_add(c_sub(Cplx{-5.0 * alpha / 6.0, 0.0}, c_div(P, c_scale(U, 3.0))), U);
    const Cplx w = c_sqrt(c_add(Cplx{alpha, 0.0}, c_scale(y, 2.0)));
    const Cplx bw = c_div(Cplx{2.0 * beta, 0.0}, w);
    const Cplx base = c_add(Cplx{3.0 * alpha, 0.0}, c_scale(y, 2.0));
    const Cplx s1 = c_sqrt(c_scale(c_add(base, bw), -1.0));
    const Cplx s2 = c_sqrt(c_scale(c_sub(base, bw), -1.0));
    const double sh = -B / (4.0 * A);
    roots[0] = sh + 0.5 * (w.re + s1.re);
    roots[1] = sh + 0.5 * (w.re - s1.re);
    roots[2] = sh + 0.5 * (-w.re + s2.re);
    roots[3] = sh + 0.5 * (-w.re - s2.re);
}

__device__ __forceinline__ void cross3(const double *a, const double *b, double *o) {
    const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
    o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void norm3(double *a) {
    const double n = sqrt(dot3(a, a));
    a[0] /= n; a[1] /= n; a[2] /= n;
}

// The solver for one sample triplet (i0, i1, i2): the models of the quartic's roots [root_lo,
// root_hi) are written to out[0], out[12], ... (finite ones only, packed) and counted.  One
// compiled body (never inlined) serves the per-triplet kernel (all four roots) and the fused wave
// kernel (one root per CTA), so a model has the same bits through either.
__device__ __noinline__ int p3p_models(uint32_t i0, uint32_t i1, uint32_t i2, const double *__restrict__ x2dn,
                                       const double *__restrict__ X3d, int root_lo, int root_hi,
                                       double *__restrict__ out) {
    double P1[3], P2[3], P3[3], f1[3], f2[3], f3[3];
    {
        for (int c = 0; c < 3; ++c) { P1[c] = X3d[3 * (size_t)i0 + c]; P2[c] = X3d[3 * (size_t)i1 + c]; P3[c] = X3d[3 * (size_t)i2 + c]; }
        f1[0] = x2dn[2 * (size_t)i0]; f1[1] = x2dn[2 * (size_t)i0 + 1]; f1[2] = 1.0;
        f2[0] = x2dn[2 * (size_t)i1]; f2[1] = x2dn[2 * (size_t)i1 + 1]; f2[2] = 1.0;
        f3[0] = x2dn[2 * (size_t)i2]; f3[1] = x2dn[2 * (size_t)i2 + 1]; f3[2] = 1.0;
        norm3(f1); norm3(f2); norm3(f3);
    }
    int n_out = 0;
    double d1[3], d2[3], cr[3];
    for (int c = 0; c < 3; ++c) { d1[c] = P2[c] - P1[c]; d2[c] = P3[c] - P1[c]; }
    cross3(d1, d2, cr);
    if (dot3(cr, cr) == 0.0) return 0;   // collinear world points

    double e1[3], e2[3], e3[3], f3t[3];
    for (int pass = 0; pass < 2; ++pass) {
        for (int c = 0; c < 3; ++c) e1[c] = f1[c];
        cross3(f1, f2, e3); norm3(e3);
        cross3(e3, e1, e2);
        f3t[0] = dot3(e1, f3); f3t[1] = dot3(e2, f3); f3t[2] = dot3(e3, f3);
        if (pass == 0 && f3t[2] > 0.0) {
            for (int c = 0; c < 3; ++c) {
                double tmp = f1[c]; f1[c] = f2[c]; f2[c] = tmp;
                tmp = P1[c]; P1[c] = P2[c]; P2[c] = tmp;
            }
            continue;
        }
        break;
    }
    double n1[3], n2[3], n3[3], dd[3];
    for (int c = 0; c < 3; ++c) { n1[c] = P2[c] - P1[c]; dd[c] = P3[c] - P1[c]; }
    const double d12 = sqrt(dot3(n1, n1));
    norm3(n1);
    cross3(n1, dd, n3); norm3(n3);
    cross3(n3, n1, n2);
    const double p1 = dot3(n1, dd), p2 = dot3(n2, dd);
    const double phi1 = f3t[0] / f3t[2], phi2 = f3t[1] / f3t[2];
    const double cosb = dot3(f1, f2);
    double b = 1.0 / (1.0 - cosb * cosb) - 1.0;
    b = cosb < 0.0 ? -sqrt(b) : sqrt(b);

    const double phi1_2 = phi1 * phi1, phi2_2 = phi2 * phi2;
    const double p1_2 = p1 * p1, p1_3 = p1_2 * p1, p1_4 = p1_3 * p1;
    const double p2_2 = p2 * p2, p2_3 = p2_2 * p2, p2_4 = p2_3 * p2;
    const double d12_2 = d12 * d12, b_2 = b * b;
    double fac[5];
    fac[0] = -phi2_2 * p2_4 - p2_4 * phi1_2 - p2_4;
    fac[1] = 2.0 * p2_3 * d12 * b + 2.0 * phi2_2 * p2_3 * d12 * b - 2.0 * phi2 * p2_3 * phi1 * d12;
    fac[2] = -phi2_2 * p2_2 * p1_2 - phi2_2 * p2_2 * d12_2 * b_2 - phi2_2 * p2_2 * d12_2 + phi2_2 * p2_4 +
             p2_4 * phi1_2 + 2.0 * p1 * p2_2 * d12 + 2.0 * phi1 * phi2 * p1 * p2_2 * d12 * b -
             p2_2 * p1_2 * phi1_2 + 2.0 * p1 * p2_2 * phi2_2 * d12 - p2_2 * d12_2 * b_2 - 2.0 * p1_2 * p2_2;
    fac[3] = 2.0 * p1_2 * p2 * d12 * b + 2.0 * phi2 * p2_3 * phi1 * d12 - 2.0 * phi2_2 * p2_3 * d12 * b -
             2.0 * p1 * p2 * d12_2 * b;
    fac[4] = -2.0 * phi2 * p2_2 * phi1 * p1 * d12 * b + phi2_2 * p2_2 * d12_2 + 2.0 * p1_3 * d12 - p1_2 * d12_2 +
             phi2_2 * p2_2 * p1_2 - p1_4 - 2.0 * phi2_2 * p2_2 * p1 * d12 + p2_2 * phi1_2 * p1_2 +
             phi2_2 * p2_2 * d12_2 * b_2;
    double roots[4];
    solve_quartic(fac, roots);

    for (int i = root_lo; i < root_hi; ++i) {
        const double cot_alpha = (-phi1 * p1 / phi2 - roots[i] * p2 + d12 * b) / (-phi1 * roots[i] * p2 / phi2 + p1 - d12);
        const double cos_theta = roots[i];
        const double sin_theta = sqrt(1.0 - roots[i] * roots[i]);
        const double sin_alpha = sqrt(1.0 / (cot_alpha * cot_alpha + 1.0));
        double cos_alpha = sqrt(1.0 - sin_alpha * sin_alpha);
        if (cot_alpha < 0.0) cos_alpha = -cos_alpha;
        const double sc = d12 * (sin_alpha * b + cos_alpha);
        const double Cn[3] = {cos_alpha * sc, cos_theta * sin_alpha * sc, sin_theta * sin_alpha * sc};
        double C[3];
        for (int c = 0; c < 3; ++c) C[c] = P1[c] + n1[c] * Cn[0] + n2[c] * Cn[1] + n3[c] * Cn[2];
        // Rr (rows), camera-to-world rotation Rcw = N^T Rr^T T, model rotation = Rcw^T = T^T Rr N
        const double Rr[9] = {-cos_alpha, -sin_alpha * cos_theta, -sin_alpha * sin_theta,
                              sin_alpha,  -cos_alpha * cos_theta, -cos_alpha * sin_theta,
                              0.0,        -sin_theta,             cos_theta};
        const double *Nrow[3] = {n1, n2, n3};
        const double *Trow[3] = {e1, e2, e3};
        double RrN[9];   // Rr * N
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c)
                RrN[3 * r + c] = Rr[3 * r] * Nrow[0][c] + Rr[3 * r + 1] * Nrow[1][c] + Rr[3 * r + 2] * Nrow[2][c];
        double M[12];
        bool ok = true;
        for (int r = 0; r < 3; ++r) {
            for (int c = 0; c < 3; ++c)   // (T^T)[r][k] = Trow[k][r]
                M[4 * r + c] = Trow[0][r] * RrN[c] + Trow[1][r] * RrN[3 + c] + Trow[2][r] * RrN[6 + c];
            M[4 * r + 3] = -(M[4 * r] * C[0] + M[4 * r + 1] * C[1] + M[4 * r + 2] * C[2]);
        }
        for (int k = 0; k < 12; ++k) ok = ok && isfinite(M[k]);
        if (!ok) continue;
        for (int k = 0; k < 12; ++k) out[12 * n_out + k] = M[k];
        ++n_out;
    }
    return n_out;
}

// One thread per sample triplet.  models: T x 4 x 12; an absent model has NaN in entry 0.
__global__ void p3p_kernel(const uint32_t *__restrict__ triplets, uint32_t T, const double *__restrict__ x2dn,
                           const double *__restrict__ X3d, double *__restrict__ models,
                           int32_t *__restrict__ n_models) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    double *out = models + (size_t)t * 48;
    for (int m = 0; m < 4; ++m) out[12 * m] = NAN;
    n_models[t] = p3p_models(triplets[3 * t], triplets[3 * t + 1], triplets[3 * t + 2], x2dn, X3d, 0, 4, out);
}

// ---- one wave of a single problem in TWO launches: (draw + solve), (score + first minimum)
// The host sampler is a counter-based generator (its n-th output is a function of seed + n), so the
// device draws sample t itself; four threads share a sample and each solves one root of its
// quartic (slot 4 t + i, NaN-marked when the root gives no model: the order of the per-triplet
// kernel's packed models, so the first minimum is the same hypothesis).  The scoring kernel's
// last block to finish reduces the wave to the record {nfa, index, model[12]} the host reads.
__device__ __forceinline__ uint64_t splitmix64_at(uint64_t s0, uint64_t n) {
    uint64_t z = s0 + n * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(128) p3p_draw_kernel(uint64_t rng0, uint32_t T, uint32_t total,
                                                       const uint32_t *__restrict__ pool,
                                                       const double *__restrict__ x2dn,
                                                       const double *__restrict__ X3d, double *__restrict__ models) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= 4 * T) return;
    const uint32_t t = h >> 2;
    // three distinct positions in [0, total), ascending insertion (sample3 on the host)
    uint32_t pos[3];
    for (int i = 0; i < 3; ++i) {
        uint32_t r = (uint32_t)(splitmix64_at(rng0, 3ull * t + i + 1) % (uint64_t)(total - i));
        int j;
        for (j = 0; j < i && r >= pos[j]; ++j) ++r;
        for (int k = i; k > j; --k) pos[k] = pos[k - 1];
        pos[j] = r;
    }
    if (pool) { pos[0] = pool[pos[0]]; pos[1] = pool[pos[1]]; pos[2] = pool[pos[2]]; }
    double M[12];
    M[0] = NAN;
    for (int k = 1; k < 12; ++k) M[k] = 0.0;
    const int root = (int)(h & 3);
    p3p_models(pos[0], pos[1], pos[2], x2dn, X3d, root, root + 1, M);
    for (int k = 0; k < 12; ++k) models[(size_t)h * 12 + k] = M[k];
}

template <int E>
__global__ void __launch_bounds__(kScoreThreads) score_wave_kernel(
    const double *__restrict__ models, const double *__restrict__ x2dn, const double *__restrict__ X3d, uint32_t N,
    const float *__restrict__ logc_n, const float *__restrict__ logc_k, double loge0, double logalpha0,
    double *__restrict__ out_nfa, unsigned int *__restrict__ counter, double *__restrict__ rec, uint64_t seq) {
    score_reg_block<E>(blockIdx.x, models, x2dn, X3d, N, logc_n, logc_k, loge0, logalpha0, -1.0f, out_nfa, nullptr,
                       nullptr, nullptr);
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) *counter = 0;                 // ready for the next wave
    // rec is pinned host memory: the record goes straight to the host, then the wave's sequence
    // number behind a system-wide fence -- the host polls that word instead of waiting on a copy
    argmin_block<true>(out_nfa, gridDim.x, models, rec);
    if (threadIdx.x == 0) {
        __threadfence_system();
        *reinterpret_cast<volatile uint64_t *>(rec + 14) = seq;
    }
}

// ------------------------------------------------------------------ host helpers
void k_inverse(const double *K, double *Ki) {
    const double fx = K[0], s = K[1], cx = K[2], fy = K[4], cy = K[5], w = K[8];
    Ki[0] = 1.0 / fx; Ki[1] = -s / (fx * fy); Ki[2] = (s * cy - cx * fy) / (fx * fy * w);
    Ki[3] = 0.0; Ki[4] = 1.0 / fy; Ki[5] = -cy / (fy * w);
    Ki[6] = 0.0; Ki[7] = 0.0; Ki[8] = 1.0 / w;
}

// x2dn = dehomogenised K^-1 [x; 1]  (ACKernelAdaptorResection_K)
void normalize_points(const double *x2d, size_t N, const double *K, std::vector<double> &out) {
    double Ki[9];
    k_inverse(K, Ki);
    out.resize(2 * N);
    for (size_t i = 0; i < N; ++i) {
        const double x = x2d[2 * i], y = x2d[2 * i + 1];
        const double a = Ki[0] * x + Ki[1] * y + Ki[2];
        const double b = Ki[3] * x + Ki[4] * y + Ki[5];
        const double c = Ki[6] * x + Ki[7] * y + Ki[8];
        out[2 * i] = a / c;
        out[2 * i + 1] = b / c;
    }
}

// log10 C(N,k) for k = 0..N and log10 C(n,3) for n = 0..N, double sums stored as float with the
// same summation order as the sequential definition (prefix sums reproduce it exactly).
// log10 of the integers, grown on demand and kept per thread: the log-binomial tables below are
// rebuilt for every problem size, and the calls to log10 were most of their cost
static const double *log10_int(size_t n_max) {
    static thread_local std::vector<double> t(2, 0.0);      // t[0] unused, t[1] = 0
    if (t.size() <= n_max) {
        const size_t old = t.size();
        t.resize(n_max + 1 + n_max / 2);
        for (size_t i = old; i < t.size(); ++i) t[i] = log10((double)i);
    }
    return t.data();
}

void make_logcombi(size_t N, std::vector<float> &logc_n, std::vector<float> &logc_k) {
    logc_n.assign(N + 1, 0.0f);
    logc_k.assign(N + 1, 0.0f);
    const double *lg = log10_int(N + 1);
    std::vector<double> prefix(N / 2 + 2, 0.0);   // prefix[j] = sum_{i=1..j} log10(N-i+1) - log10(i)
    for (size_t j = 1; j < prefix.size() && j <= N; ++j)
        prefix[j] = prefix[j - 1] + (lg[N - j + 1] - lg[j]);
    for (size_t k = 0; k <= N; ++k) {
        if (k >= N || k == 0) { logc_n[k] = 0.0f; continue; }
        const size_t kk = (N - k < k) ? N - k : k;
        logc_n[k] = (float)prefix[kk];
    }
    for (size_t n = 0; n <= N; ++n) {
        size_t k = 3;
        if (k >= n) { logc_k[n] = 0.0f; continue; }
        if (n - k < k) k = n - k;
        double r = 0.0;
        for (size_t i = 1; i <= k; ++i) r += lg[n - i + 1] - lg[i];
        logc_k[n] = (float)r;
    }
}

uint32_t next_pow2(uint32_t v) {
    uint32_t p = 64;
    while (p < v) p <<= 1;
    return p;
}

// Device-side staging of one resection problem.
struct Problem {
    size_t N = 0;
    double *d_x2dn = nullptr, *d_X3d = nullptr;
    float *d_logc_n = nullptr, *d_logc_k = nullptr;
    double loge0 = 0, logalpha0 = 0;
    std::vector<float> lcn, lck;       // the same tables on the host (final fp64 rescoring)
};

// scratch0: x2dn | X3d | logc_n | logc_k   (problem);  scratch1: models; scratch2: outputs; scratch3: triplets
int stage_problem(hulo_gpu *h, const std::vector<double> &x2dn, const double *X3d, size_t N, Problem &pb) {
    std::vector<float> &lcn = pb.lcn, &lck = pb.lck;
    if (lcn.size() != N + 1 || lck.size() != N + 1) make_logcombi(N, lcn, lck);   // else: handed in by the caller
    const size_t bytes = N * 5 * sizeof(double) + 2 * (N + 1) * sizeof(float) + 64;
    HULO_CUDA(h->scratch0.reserve(bytes));
    pb.N = N;
    pb.d_x2dn = h->scratch0.as<double>();
    pb.d_X3d = pb.d_x2dn + 2 * N;
    pb.d_logc_n = reinterpret_cast<float *>(pb.d_X3d + 3 * N);
    pb.d_logc_k = pb.d_logc_n + (N + 1);
    pb.loge0 = log10(4.0 * (double)(N > 3 ? N - 3 : 1));
    pb.logalpha0 = log10(M_PI);
    // one copy: the four arrays are packed in pinned memory in the device layout.  Every caller
    // synchronises the stream before the staging buffer can be written again.
    HULO_CUDA(h->hstage1.reserve(bytes));
    uint8_t *hp = h->hstage1.as<uint8_t>();
    if (N > 0) {
        memcpy(hp, x2dn.data(), 2 * N * sizeof(double));
        memcpy(hp + 2 * N * sizeof(double), X3d, 3 * N * sizeof(double));
    }
    memcpy(hp + 5 * N * sizeof(double), lcn.data(), (N + 1) * sizeof(float));
    memcpy(hp + 5 * N * sizeof(double) + (N + 1) * sizeof(float), lck.data(), (N + 1) * sizeof(float));
    HULO_CUDA(cudaMemcpyAsync(pb.d_x2dn, hp, 5 * N * sizeof(double) + 2 * (N + 1) * sizeof(float), cudaMemcpyHostToDevice,
                              h->stream));
    return HULO_OK;
}

struct ScoreOut { double *nfa; int32_t *k; float *errk; int32_t *ninl; };

int launch_score(hulo_gpu *h, const Problem &pb, const double *d_models, size_t H, float thr2, ScoreOut &o) {
    const size_t bytes = H * (sizeof(double) + 2 * sizeof(int32_t) + sizeof(float)) + 64 + 16 * sizeof(double);
    HULO_CUDA(h->scratch2.reserve(bytes));
    o.nfa = h->scratch2.as<double>();
    o.k = reinterpret_cast<int32_t *>(o.nfa + H);
    o.errk = reinterpret_cast<float *>(o.k + H);
    o.ninl = reinterpret_cast<int32_t *>(o.errk + H);
    if (H == 0) return HULO_OK;
    if (pb.N <= 4096 && !getenv("HULO_K2_SMEM_SORT")) {
        // register-resident sort: E elements per thread, npad = 256 E
        uint32_t np = kScoreThreads;
        while (np < pb.N) np <<= 1;
#define HULO_K2_REG(EE)                                                                                            \
    score_kernel_reg<EE><<<(unsigned)H, kScoreThreads, 0, h->stream>>>(d_models, (uint32_t)H, pb.d_x2dn, pb.d_X3d,     \
                                                                     (uint32_t)pb.N, pb.d_logc_n, pb.d_logc_k, pb.loge0, \
                                                                     pb.logalpha0, thr2, o.nfa, o.k, o.errk, o.ninl)
        switch (np / kScoreThreads) {
            case 1: HULO_K2_REG(1); break;
            case 2: HULO_K2_REG(2); break;
            case 4: HULO_K2_REG(4); break;
            case 8: HULO_K2_REG(8); break;
            default: HULO_K2_REG(16); break;
        }
#undef HULO_K2_REG
        HULO_CUDA(cudaGetLastError());
        h->launches++;
        return HULO_OK;
    }
    const uint32_t npad = next_pow2((uint32_t)pb.N);
    const size_t smem = npad * sizeof(float);
    if (smem > 48 * 1024 && smem > h->score_smem_configured) {   // per device, so per context
        HULO_CUDA(cudaFuncSetAttribute(score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        h->score_smem_configured = smem;
    }
    score_kernel<<<(unsigned)H, kScoreThreads, smem, h->stream>>>(d_models, (uint32_t)H, pb.d_x2dn, pb.d_X3d,
                                                                (uint32_t)pb.N, npad, pb.d_logc_n, pb.d_logc_k,
                                                                pb.loge0, pb.logalpha0, thr2, o.nfa, o.k, o.errk,
                                                                o.ninl);
    HULO_CUDA(cudaGetLastError());
    h->launches++;
    return HULO_OK;
}

uint64_t splitmix64(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// three distinct positions in [0, total), ascending insertion (UniformSample, rand_sampling.hpp)
void sample3(uint64_t &state, size_t total, size_t out[3]) {
    for (int i = 0; i < 3; ++i) {
        size_t r = (size_t)(splitmix64(state) % (uint64_t)(total - i));
        int j;
        for (j = 0; j < i && r >= out[j]; ++j) ++r;
        for (int k = i; k > j; --k) out[k] = out[k - 1];
        out[j] = r;
    }
}


// ---- one AC-RANSAC resection as a resumable schedule
// The host side of hulo_resect_acransac: which triplets to draw next, and what a scored batch means.
// The device work of a step (P3P over the drawn triplets, K2 scoring, first minimum) is done by the
// caller -- alone for one problem, or for a whole wave of problems at once
// (hulo_resect_acransac_batch) -- and comes back as the record {nfa, index, model[12]}.
//
// Schedule of the sequential algorithm: 10 % of the iterations are reserved; once a meaningful model
// (NFA < 0) exists the sampler draws from its inliers and only the reserved iterations remain.
// Batched: global batches until a meaningful model appears (at most max_iter - reserve draws), then
// one batch of `reserve` draws from the inliers.  The first batches are small and grow (64, 128, 256,
// 512, 512, ...): the sequential algorithm leaves the global phase at its first meaningful model,
// which on clean correspondence sets is one of the very first draws, so a large first batch would
// only score models it discards.
struct ResectJob {
    enum Phase { kGlobal, kFocused, kReserveGlobal, kDone };
    struct EI { double e; size_t i; };

    size_t N = 0;
    const double *X3d = nullptr;
    const double *K = nullptr;
    std::vector<double> x2dn;
    std::vector<float> lcn, lck;
    double loge0 = 0, logalpha0 = 0;

    Phase phase = kDone;
    size_t reserve = 0, global_budget = 0, batch = 0, drawn = 0, last_T = 0;
    uint64_t rng = 0;
    std::vector<size_t> pool;
    double best_nfa = INFINITY, final_nfa = INFINITY;
    double best_model[12] = {0}, prev_model[12] = {0};
    double prev_nfa = INFINITY;
    size_t best_k = 0;
    double best_err = 0.0;
    std::vector<EI> ei;

    void init(const double *x2d, const double *X3d_, size_t N_, const double *K_, size_t max_iter, uint64_t seed) {
        N = N_; X3d = X3d_; K = K_;
        normalize_points(x2d, N, K, x2dn);
        make_logcombi(N, lcn, lck);
        loge0 = log10(4.0 * (double)(N > 3 ? N - 3 : 1));
        logalpha0 = log10(M_PI);
        reserve = max_iter / 10;
        global_budget = max_iter - reserve;
        batch = std::min<size_t>(global_budget, 64);
        rng = seed;
        pool.resize(N);
        for (size_t i = 0; i < N; ++i) pool[i] = i;
        ei.resize(N);
        phase = kGlobal;
    }
    // triplets of the next step (0: finished)
    size_t next_T() const {
        switch (phase) {
            case kGlobal: return std::min(batch, global_budget - drawn);
            case kFocused: case kReserveGlobal: return reserve;
            default: return 0;
        }
    }
    // draws T triplets; indices are offset by `base` (the problem's first row in a shared arena)
    void draw(size_t T, uint32_t *tri, uint32_t base) {
        for (size_t t = 0; t < T; ++t) {
            size_t pos[3];
            sample3(rng, pool.size(), pos);
            for (int s = 0; s < 3; ++s) tri[3 * t + s] = base + (uint32_t)pool[pos[s]];
        }
        last_T = T;
    }
    // the device drew T triplets itself: advance the generator past them (three outputs each)
    void skip_draws(size_t T) {
        rng += (uint64_t)(3 * T) * 0x9E3779B97F4A7C15ULL;
        last_T = T;
    }
    // (residual, index) ascending.  The residuals are non-negative doubles (or +inf), whose bit
    // patterns order like the values: a stable radix sort of the upper 32 bits (three passes of
    // 11 / 11 / 10 bits from the index order), then the rare runs that agree in those bits are put
    // in full (residual, index) order -- exactly the order of the comparison sort it replaces.
    std::vector<EI> ei_tmp;
    std::vector<uint32_t> hist;
    void sort_by_residual() {
        auto less = [](const EI &a, const EI &b) { return a.e < b.e || (a.e == b.e && a.i < b.i); };
        if (N < 64) {
            std::sort(ei.begin(), ei.end(), less);
            return;
        }
        hist.assign(2048 + 2048 + 1024, 0);
        uint32_t *h0 = hist.data(), *h1 = h0 + 2048, *h2 = h1 + 2048;
        auto hi32 = [](const EI &x) { uint64_t k; memcpy(&k, &x.e, 8); return (uint32_t)(k >> 32); };
        for (size_t i = 0; i < N; ++i) {
            if (ei[i].e == 0.0) ei[i].e = 0.0;            // -0.0 would sort as the largest key
            const uint32_t k = hi32(ei[i]);
            ++h0[k & 2047]; ++h1[(k >> 11) & 2047]; ++h2[k >> 22];
        }
        ei_tmp.resize(N);
        EI *src = ei.data(), *dst = ei_tmp.data();
        const int shift[3] = {0, 11, 22}, bins[3] = {2048, 2048, 1024};
        uint32_t *hh[3] = {h0, h1, h2};
        for (int b = 0; b < 3; ++b) {
            uint32_t *h = hh[b];
            bool trivial = false;
            uint32_t run = 0;
            for (int d = 0; d < bins[b]; ++d) {
                const uint32_t c = h[d];
                if (c == N) { trivial = true; break; }
                h[d] = run; run += c;
            }
            if (trivial) continue;
            const uint32_t mask = (uint32_t)bins[b] - 1;
            for (size_t i = 0; i < N; ++i) dst[h[(hi32(src[i]) >> shift[b]) & mask]++] = src[i];
            std::swap(src, dst);
        }
        if (src != ei.data()) memcpy(ei.data(), src, N * sizeof(EI));
        for (size_t i = 0; i + 1 < N;) {
            size_t j = i + 1;
            const uint32_t k = hi32(ei[i]);
            while (j < N && hi32(ei[j]) == k) ++j;
            if (j - i > 1) std::sort(ei.begin() + i, ei.begin() + j, less);
            i = j;
        }
    }

    // Final scoring of a model in fp64 on the host: residuals, (residual, index) order, NFA.
    // The NFA minimum is found in two passes: a float log2 brackets every term
    // (|log10 x - log2f(x) log10(2)| < 3e-6 for x in [1e-7, 1e30]: 1 ulp of log2f plus the
    // float rounding of x; kSlack leaves a factor 3), and only the terms whose bracket reaches
    // below the smallest upper bound are evaluated with the double log10 -- the same terms, in the
    // same order and with the same strict comparison the full scan would have kept.
    std::vector<double> nfa_lo;
    double finalize(const double *M) {
        for (size_t i = 0; i < N; ++i) {
            const double *X = X3d + 3 * i;
            const double u = M[0] * X[0] + M[1] * X[1] + M[2] * X[2] + M[3];
            const double v = M[4] * X[0] + M[5] * X[1] + M[6] * X[2] + M[7];
            const double w = M[8] * X[0] + M[9] * X[1] + M[10] * X[2] + M[11];
            const double dx = u / w - x2dn[2 * i], dy = v / w - x2dn[2 * i + 1];
            double e = dx * dx + dy * dy;
            if (!(e == e)) e = INFINITY;
            ei[i] = EI{e, i};
        }
        sort_by_residual();
        constexpr double kSlack = 1e-5, kLog10of2 = 0.30102999566398120;
        nfa_lo.resize(N + 1);
        double ub_min = INFINITY;
        size_t k_end = 3;
        for (size_t k = 4; k <= N; ++k) {
            if (!(ei[k - 1].e < INFINITY)) break;
            k_end = k;
            const double x = ei[k - 1].e + (double)FLT_EPSILON;
            if (!(x < 1e30)) { nfa_lo[k] = -INFINITY; continue; }        // outside the bracket's range: always exact
            const double la = logalpha0 + (double)log2f((float)x) * kLog10of2;
            const double nf = loge0 + la * (double)(k - 3) + (double)lcn[k] + (double)lck[k];
            const double slack = kSlack * (double)(k - 3);
            nfa_lo[k] = nf - slack;
            ub_min = std::min(ub_min, nf + slack);
        }
        double bn = INFINITY;
        size_t bk = 3;
        for (size_t k = 4; k <= k_end; ++k) {
            if (!(nfa_lo[k] <= ub_min)) continue;
            const double logalpha = logalpha0 + log10(ei[k - 1].e + (double)FLT_EPSILON);
            const double nfa = loge0 + logalpha * (double)(k - 3) + (double)lcn[k] + (double)lck[k];
            if (nfa < bn) { bn = nfa; bk = k; }
        }
        best_k = bk;
        best_err = bk >= 1 ? ei[bk - 1].e : 0.0;
        return bn;
    }
    // the scored step comes back: rec = {nfa, index or -1, model[12]} of its first minimum
    void absorb(const double *rec) {
        // sequential update rule: strict <, so the earliest hypothesis wins ties
        if (rec[1] >= 0.0 && rec[0] < best_nfa) {
            best_nfa = rec[0];
            memcpy(best_model, rec + 2, 12 * sizeof(double));
        }
        if (phase == kGlobal) {
            drawn += last_T;
            if (!(best_nfa < 0.0) && drawn < global_budget) {
                batch = std::min<size_t>(batch * 2, 512);
                return;
            }
            final_nfa = INFINITY;
            if (best_nfa < INFINITY) final_nfa = finalize(best_model);
            if (reserve > 0 && final_nfa < 0.0 && best_k >= 3) {
                // focused sampling among the inliers of the best model so far
                pool.resize(best_k);
                for (size_t i = 0; i < best_k; ++i) pool[i] = ei[i].i;
                memcpy(prev_model, best_model, sizeof prev_model);
                prev_nfa = best_nfa;
                phase = kFocused;
            } else if (reserve > 0 && !(final_nfa < 0.0)) {
                phase = kReserveGlobal;   // no meaningful model yet: the reserved iterations keep sampling globally
            } else {
                phase = kDone;
            }
        } else if (phase == kFocused) {
            if (best_nfa < prev_nfa) {
                const double keep_nfa = final_nfa;
                const size_t keep_k = best_k;
                const double keep_err = best_err;
                std::vector<EI> keep_ei(ei.begin(), ei.begin() + keep_k);
                const double cand = finalize(best_model);
                if (cand < keep_nfa) {
                    final_nfa = cand;
                } else {   // fp32 ranking disagreed with the fp64 rescoring: keep the earlier model
                    memcpy(best_model, prev_model, sizeof prev_model);
                    best_k = keep_k; best_err = keep_err;
                    std::copy(keep_ei.begin(), keep_ei.end(), ei.begin());
                }
            }
            phase = kDone;
        } else if (phase == kReserveGlobal) {
            if (best_nfa < INFINITY) final_nfa = finalize(best_model);
            phase = kDone;
        }
    }
    // outputs of SfM_Localizer::Localize
    void emit(double *P, int32_t *inliers, size_t *n_inliers, double *error_max, int *found) const {
        *n_inliers = 0; *error_max = 0.0; *found = 0;
        if (!(final_nfa < 0.0)) return;   // minNFA >= 0: inliers cleared, not found
        // Unnormalize: P = K * model ; error in pixels = sqrt(e) / Kinv(0,0)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c)
                P[4 * r + c] = K[3 * r] * best_model[c] + K[3 * r + 1] * best_model[4 + c] + K[3 * r + 2] * best_model[8 + c];
        *error_max = sqrt(best_err) * K[0];
        for (size_t i = 0; i < best_k; ++i) inliers[i] = (int32_t)ei[i].i;
        *n_inliers = best_k;
        // SfM_Localizer::Localize: resection succeeded iff #inliers > 2.5 * MINIMUM_SAMPLES
        *found = (double)best_k > 2.5 * 3.0 ? 1 : 0;
    }
};

template <class F>
void parallel_for(size_t n, size_t grain, F f) {
    const size_t hw = std::max<size_t>(1, std::min<size_t>(std::thread::hardware_concurrency(), 32));
    const size_t nt = std::min(hw, (n + grain - 1) / std::max<size_t>(grain, 1));
    if (nt <= 1) { for (size_t i = 0; i < n; ++i) f(i); return; }
    std::atomic<size_t> next{0};
    std::vector<std::thread> th;
    auto body = [&]() {
        for (;;) {
            const size_t b = next.fetch_add(grain);
            if (b >= n) return;
            for (size_t i = b; i < std::min(n, b + grain); ++i) f(i);
        }
    };
    for (size_t t = 1; t < nt; ++t) th.emplace_back(body);
    body();
    for (auto &t : th) t.join();
}

}  // namespace
}  // namespace hulo

using namespace hulo;

extern "C" {

int hulo_score_resection(hulo_gpu *h, const double *models, size_t H, const double *x2d, const double *X3d,
                         size_t N, const double *K, double thr_px, float *nfa, int32_t *k_best, float *err_k,
                         int32_t *n_inl) {
    HULO_ARG(h != nullptr && K != nullptr, "null argument");
    HULO_ARG(H == 0 || models != nullptr, "models is null");
    HULO_ARG(N == 0 || (x2d != nullptr && X3d != nullptr), "null correspondences");
    HULO_ARG(N <= kMaxPoints, "more than 32768 correspondences");
    HULO_CUDA(cudaSetDevice(h->device));
    std::vector<double> x2dn;
    normalize_points(x2d, N, K, x2dn);
    Problem pb;
    int rc = stage_problem(h, x2dn, X3d, N, pb);
    if (rc != HULO_OK) return rc;
    HULO_CUDA(h->scratch1.reserve(std::max<size_t>(H, 1) * 12 * sizeof(double)));
    if (H) HULO_CUDA(cudaMemcpyAsync(h->scratch1.ptr, models, H * 12 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    const double fx = K[0];
    const float thr2 = thr_px >= 0.0 ? (float)((thr_px / fx) * (thr_px / fx)) : -1.0f;
    ScoreOut o;
    rc = launch_score(h, pb, h->scratch1.as<double>(), H, thr2, o);
    if (rc != HULO_OK) return rc;
    if (H) {
        std::vector<double> h_nfa(H);
        std::vector<float> h_err(H);
        HULO_CUDA(cudaMemcpyAsync(h_nfa.data(), o.nfa, H * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        if (k_best) HULO_CUDA(cudaMemcpyAsync(k_best, o.k, H * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaMemcpyAsync(h_err.data(), o.errk, H * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        if (n_inl) HULO_CUDA(cudaMemcpyAsync(n_inl, o.ninl, H * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaStreamSynchronize(h->stream));
        for (size_t i = 0; i < H; ++i) {
            if (nfa) nfa[i] = (float)h_nfa[i];
            // unormalizeError: sqrt(e) / Kinv(0,0) = sqrt(e) * fx
            if (err_k) err_k[i] = (float)(sqrt((double)h_err[i]) * fx);
        }
    }
    return HULO_OK;
}

int hulo_resection_residuals(hulo_gpu *h, const double *models, size_t H, const double *x2d, const double *X3d,
                             size_t N, const double *K, float *res_px) {
    HULO_ARG(h != nullptr && K != nullptr, "null argument");
    HULO_ARG(H == 0 || models != nullptr, "models is null");
    HULO_ARG(N == 0 || (x2d != nullptr && X3d != nullptr), "null correspondences");
    if (H == 0 || N == 0) return HULO_OK;
    HULO_ARG(res_px != nullptr, "null output");
    HULO_ARG(H <= 65535, "more than 65535 hypotheses in one residual dump");
    HULO_CUDA(cudaSetDevice(h->device));
    std::vector<double> x2dn;
    normalize_points(x2d, N, K, x2dn);
    Problem pb;
    int rc = stage_problem(h, x2dn, X3d, N, pb);
    if (rc != HULO_OK) return rc;
    HULO_CUDA(h->scratch1.reserve(H * 12 * sizeof(double)));
    HULO_CUDA(cudaMemcpyAsync(h->scratch1.ptr, models, H * 12 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(h->scratch2.reserve(H * N * sizeof(float)));
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)H);
    residual_kernel<<<grid, 256, 0, h->stream>>>(h->scratch1.as<double>(), (uint32_t)H, pb.d_x2dn, pb.d_X3d,
                                                 (uint32_t)N, (float)K[0], h->scratch2.as<float>());
    HULO_CUDA(cudaGetLastError());
    h->launches++;
    HULO_CUDA(cudaMemcpyAsync(res_px, h->scratch2.ptr, H * N * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}

int hulo_p3p(hulo_gpu *h, const uint32_t *triplets, size_t T, const double *x2d, const double *X3d, size_t N,
             const double *K, double *models, int32_t *n_models) {
    HULO_ARG(h != nullptr && K != nullptr, "null argument");
    HULO_ARG(T == 0 || (triplets != nullptr && models != nullptr && n_models != nullptr), "null argument");
    HULO_ARG(N == 0 || (x2d != nullptr && X3d != nullptr), "null correspondences");
    for (size_t t = 0; t < 3 * T; ++t) HULO_ARG(triplets[t] < N, "triplet index out of range");
    if (T == 0) return HULO_OK;
    HULO_CUDA(cudaSetDevice(h->device));
    std::vector<double> x2dn;
    normalize_points(x2d, N, K, x2dn);
    Problem pb;
    int rc = stage_problem(h, x2dn, X3d, N, pb);
    if (rc != HULO_OK) return rc;
    HULO_CUDA(h->scratch3.reserve(T * 3 * sizeof(uint32_t)));
    HULO_CUDA(h->scratch1.reserve(T * 48 * sizeof(double) + T * sizeof(int32_t)));
    HULO_CUDA(cudaMemcpyAsync(h->scratch3.ptr, triplets, T * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    double *d_models = h->scratch1.as<double>();
    int32_t *d_nm = reinterpret_cast<int32_t *>(d_models + T * 48);
    p3p_kernel<<<(unsigned)((T + 127) / 128), 128, 0, h->stream>>>(h->scratch3.as<uint32_t>(), (uint32_t)T, pb.d_x2dn,
                                                                 pb.d_X3d, d_models, d_nm);
    HULO_CUDA(cudaGetLastError());
    h->launches++;
    HULO_CUDA(cudaMemcpyAsync(models, d_models, T * 48 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaMemcpyAsync(n_models, d_nm, T * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}

int hulo_resect_acransac(hulo_gpu *h, const double *x2d, const double *X3d, size_t N, const double *K,
                         size_t max_iter, uint64_t seed, double *P, int32_t *inliers, size_t *n_inliers,
                         double *error_max, int *found) {
    HULO_ARG(h != nullptr && K != nullptr && P != nullptr && n_inliers != nullptr && error_max != nullptr &&
                 found != nullptr, "null argument");
    HULO_ARG(N == 0 || (x2d != nullptr && X3d != nullptr && inliers != nullptr), "null correspondences");
    HULO_ARG(N <= kMaxPoints, "more than 32768 correspondences");
    *n_inliers = 0; *error_max = 0.0; *found = 0;
    // ACRANSAC: nothing to do with N <= MINIMUM_SAMPLES
    if (N <= 3 || max_iter == 0) return HULO_OK;
    HULO_CUDA(cudaSetDevice(h->device));
    static const bool trace = getenv("HULO_RESECT_TRACE") != nullptr;      // host time per phase on stderr
    auto now = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    double t_mark = trace ? now() : 0.0;
    auto lap = [&](const char *what) {
        if (!trace) return;
        const double t = now();
        fprintf(stderr, "[resect] %-10s %7.1f us\n", what, t - t_mark);
        t_mark = t;
    };
    ResectJob job;
    job.init(x2d, X3d, N, K, max_iter, seed);
    lap("init");
    Problem pb;
    pb.lcn = job.lcn;
    pb.lck = job.lck;
    int rc = stage_problem(h, job.x2dn, X3d, N, pb);
    if (rc != HULO_OK) return rc;
    lap("stage");

    // Two launches per wave (draw + P3P, scoring + first minimum) when the sort fits in registers:
    // nothing goes up but the focused phase's pool; the host-drawn three-kernel form below otherwise.
    static const bool three_kernels = getenv("HULO_RESECT_UNFUSED") != nullptr || getenv("HULO_K2_SMEM_SORT") != nullptr;
    if (N <= 4096 && !three_kernels) {
        if (!h->wave_counter.ptr) {
            HULO_CUDA(h->wave_counter.reserve(sizeof(unsigned int)));
            HULO_CUDA(cudaMemsetAsync(h->wave_counter.ptr, 0, sizeof(unsigned int), h->stream));
        }
        if (!h->wave_rec.ptr) {
            HULO_CUDA(h->wave_rec.reserve(16 * sizeof(double)));
            memset(h->wave_rec.ptr, 0, 16 * sizeof(double));
        }
        double *rec = h->wave_rec.as<double>();
        volatile uint64_t *rec_seq = reinterpret_cast<volatile uint64_t *>(rec + 14);
        for (size_t T = job.next_T(); T > 0; T = job.next_T()) {
            const size_t H = 4 * T;
            HULO_CUDA(h->scratch1.reserve(H * 12 * sizeof(double)));
            HULO_CUDA(h->scratch2.reserve(H * sizeof(double)));
            const uint32_t *d_pool = nullptr;
            if (job.phase == ResectJob::kFocused) {
                // focused phase: the sampler draws among the inliers of the best model (the problem's
                // staging copy has completed: the previous wave was waited for)
                HULO_CUDA(h->scratch3.reserve(job.pool.size() * sizeof(uint32_t)));
                HULO_CUDA(h->hstage1.reserve(job.pool.size() * sizeof(uint32_t)));
                uint32_t *hp = h->hstage1.as<uint32_t>();
                for (size_t i = 0; i < job.pool.size(); ++i) hp[i] = (uint32_t)job.pool[i];
                HULO_CUDA(cudaMemcpyAsync(h->scratch3.ptr, hp, job.pool.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                          h->stream));
                d_pool = h->scratch3.as<uint32_t>();
            }
            double *d_models = h->scratch1.as<double>();
            double *d_nfa = h->scratch2.as<double>();
            const uint64_t seq = ++h->wave_seq;
            uint32_t np = kScoreThreads;
            while (np < N) np <<= 1;
            p3p_draw_kernel<<<(unsigned)((H + 127) / 128), 128, 0, h->stream>>>(job.rng, (uint32_t)T, (uint32_t)job.pool.size(),
                                                                                d_pool, pb.d_x2dn, pb.d_X3d, d_models);
            HULO_CUDA(cudaGetLastError());
            h->launches++;
#define HULO_WAVE(EE)                                                                                              \
    score_wave_kernel<EE><<<(unsigned)H, kScoreThreads, 0, h->stream>>>(                                               \
        d_models, pb.d_x2dn, pb.d_X3d, (uint32_t)N, pb.d_logc_n, pb.d_logc_k, pb.loge0, pb.logalpha0, d_nfa,         \
        h->wave_counter.as<unsigned int>(), rec, seq)
            switch (np / kScoreThreads) {
                case 1: HULO_WAVE(1); break;
                case 2: HULO_WAVE(2); break;
                case 4: HULO_WAVE(4); break;
                case 8: HULO_WAVE(8); break;
                default: HULO_WAVE(16); break;
            }
#undef HULO_WAVE
            HULO_CUDA(cudaGetLastError());
            h->launches++;
            job.skip_draws(T);
            lap("enqueue");
            // poll the record's sequence word; the stream is queried now and then so that a failed
            // launch is reported instead of waited for
            for (uint32_t spins = 1; *rec_seq != seq; ++spins) {
                if ((spins & 0xFFFu) == 0) {
                    const cudaError_t q = cudaStreamQuery(h->stream);
                    if (q != cudaErrorNotReady) {
                        HULO_CUDA(q);
                        if (*rec_seq != seq) {
                            // an earlier wave died half-way and left its blocks-done count behind: start clean next time
                            cudaMemsetAsync(h->wave_counter.ptr, 0, sizeof(unsigned int), h->stream);
                            set_error("hulo_resect_acransac: the wave finished without its record");
                            return HULO_ERR_CUDA;
                        }
                    }
                }
#if defined(__x86_64__)
                __builtin_ia32_pause();
#endif
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            lap("wait");
            job.absorb(rec);
            lap("absorb");
        }
        job.emit(P, inliers, n_inliers, error_max, found);
        lap("emit");
        return HULO_OK;
    }

    std::vector<uint32_t> tri;
    for (size_t T = job.next_T(); T > 0; T = job.next_T()) {
        tri.resize(3 * T);
        job.draw(T, tri.data(), 0);
        lap("draw");
        HULO_CUDA(h->scratch3.reserve(T * 3 * sizeof(uint32_t)));
        HULO_CUDA(h->scratch1.reserve(T * 48 * sizeof(double) + T * sizeof(int32_t)));
        HULO_CUDA(cudaMemcpyAsync(h->scratch3.ptr, tri.data(), T * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        double *d_models = h->scratch1.as<double>();
        int32_t *d_nm = reinterpret_cast<int32_t *>(d_models + T * 48);
        p3p_kernel<<<(unsigned)((T + 127) / 128), 128, 0, h->stream>>>(h->scratch3.as<uint32_t>(), (uint32_t)T,
                                                                     pb.d_x2dn, pb.d_X3d, d_models, d_nm);
        HULO_CUDA(cudaGetLastError());
        h->launches++;
        ScoreOut o;
        rc = launch_score(h, pb, d_models, 4 * T, -1.0f, o);
        if (rc != HULO_OK) return rc;
        // first minimum of the batch and its model in one record, one copy, one synchronisation
        double rec[14];
        double *d_rec = reinterpret_cast<double *>(o.ninl + 4 * T + 2);
        d_rec = reinterpret_cast<double *>((reinterpret_cast<uintptr_t>(d_rec) + 7) & ~(uintptr_t)7);
        argmin_kernel<<<1, 256, 0, h->stream>>>(o.nfa, (uint32_t)(4 * T), d_models, d_rec);
        HULO_CUDA(cudaGetLastError());
        h->launches++;
        lap("enqueue");
        HULO_CUDA(cudaMemcpyAsync(rec, d_rec, sizeof rec, cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaStreamSynchronize(h->stream));
        lap("wait");
        job.absorb(rec);
        lap("absorb");
    }
    job.emit(P, inliers, n_inliers, error_max, found);
    lap("emit");
    return HULO_OK;
}

// Many independent resections at once: the views of a reconstruction re-resected against its own
// structure (OpenMVG_BA/src/adjust_sfm_data.cpp:91-146, an omp loop of SfM_Localizer::Localize), or
// the queries of a server batch.  Every problem follows exactly the schedule of
// hulo_resect_acransac with its own seed -- same draws, same kernels' arithmetic, same decisions,
// so the results are bit-identical to n_problems single calls -- but a step of all problems still
// running is ONE wave on the device: one P3P launch over all drawn triplets, one scoring launch
// per size class (the register-resident sort is compiled per points-per-thread), one first-minimum
// launch, one copy back.  The host part of a step (fp64 rescoring when a problem changes phase)
// runs on the host threads.
int hulo_resect_acransac_batch(hulo_gpu *h, size_t n_problems, const uint64_t *offsets, const double *x2d,
                               const double *X3d, const double *K, size_t max_iter, uint64_t seed,
                               const uint64_t *seeds, double *P, int32_t *inliers, uint64_t *n_inliers,
                               double *error_max, int32_t *found) {
    HULO_ARG(h != nullptr, "null handle");
    if (n_problems == 0) return HULO_OK;
    HULO_ARG(offsets != nullptr && K != nullptr && P != nullptr && n_inliers != nullptr && error_max != nullptr &&
                 found != nullptr, "null argument");
    const size_t total = (size_t)offsets[n_problems];
    HULO_ARG(total == 0 || (x2d != nullptr && X3d != nullptr && inliers != nullptr), "null correspondences");
    HULO_ARG(total < 0xFFFFFFFFull, "too many correspondences in one call");
    for (size_t p = 0; p < n_problems; ++p) {
        HULO_ARG(offsets[p] <= offsets[p + 1], "offsets must ascend");
        HULO_ARG(offsets[p + 1] - offsets[p] <= kMaxPoints, "more than 32768 correspondences in one problem");
        n_inliers[p] = 0; error_max[p] = 0.0; found[p] = 0;
    }
    HULO_CUDA(cudaSetDevice(h->device));
    auto job_seed = [&](size_t p) { return seeds ? seeds[p] : seed + 1000003ull * (uint64_t)p; };

    // problems the batched kernels take (register-resident sort: N <= 4096); the rest go one by one
    std::vector<ResectJob> jobs(n_problems);
    std::vector<uint32_t> batched, single;
    const bool smem_sort = getenv("HULO_K2_SMEM_SORT") != nullptr;
    for (size_t p = 0; p < n_problems; ++p) {
        const size_t N = (size_t)(offsets[p + 1] - offsets[p]);
        if (N <= 3 || max_iter == 0) continue;
        if (N > 4096 || smem_sort) single.push_back((uint32_t)p); else batched.push_back((uint32_t)p);
    }
    parallel_for(batched.size(), 8, [&](size_t b) {
        const size_t p = batched[b];
        jobs[p].init(x2d + 2 * offsets[p], X3d + 3 * offsets[p], (size_t)(offsets[p + 1] - offsets[p]), K + 9 * p,
                     max_iter, job_seed(p));
    });

    if (!batched.empty()) {
        // ---- arena: normalised observations, points, log-binomial tables, descriptors
        std::vector<double> h_x2dn(2 * total);
        std::vector<uint64_t> tab_off(n_problems + 1, 0);
        for (size_t p = 0; p < n_problems; ++p) tab_off[p + 1] = tab_off[p] + (jobs[p].phase != ResectJob::kDone ? jobs[p].N + 1 : 0);
        std::vector<float> h_lcn(tab_off[n_problems]), h_lck(tab_off[n_problems]);
        for (uint32_t p : batched) {
            memcpy(h_x2dn.data() + 2 * offsets[p], jobs[p].x2dn.data(), jobs[p].x2dn.size() * sizeof(double));
            memcpy(h_lcn.data() + tab_off[p], jobs[p].lcn.data(), (jobs[p].N + 1) * sizeof(float));
            memcpy(h_lck.data() + tab_off[p], jobs[p].lck.data(), (jobs[p].N + 1) * sizeof(float));
        }
        const size_t b_x = 2 * total * sizeof(double), b_X = 3 * total * sizeof(double);
        const size_t b_t = ((tab_off[n_problems] * sizeof(float) + 15) / 16) * 16;
        HULO_CUDA(h->scratch0.reserve(b_x + b_X + 2 * b_t + n_problems * sizeof(ResectDesc) + 64));
        double *d_x2dn = h->scratch0.as<double>();
        double *d_X3d = d_x2dn + 2 * total;
        float *d_lcn = reinterpret_cast<float *>(d_X3d + 3 * total);
        float *d_lck = reinterpret_cast<float *>(reinterpret_cast<char *>(d_lcn) + b_t);
        ResectDesc *d_desc = reinterpret_cast<ResectDesc *>(reinterpret_cast<char *>(d_lck) + b_t);
        std::vector<ResectDesc> h_desc(n_problems);
        for (size_t p = 0; p < n_problems; ++p) {
            ResectDesc &d = h_desc[p];
            d.x2dn = d_x2dn + 2 * offsets[p];
            d.X3d = d_X3d + 3 * offsets[p];
            d.logc_n = d_lcn + tab_off[p];
            d.logc_k = d_lck + tab_off[p];
            d.loge0 = jobs[p].loge0;
            d.logalpha0 = jobs[p].logalpha0;
            d.N = (uint32_t)jobs[p].N;
            d.pad = 0;
        }
        HULO_CUDA(cudaMemcpyAsync(d_x2dn, h_x2dn.data(), b_x, cudaMemcpyHostToDevice, h->stream));
        HULO_CUDA(cudaMemcpyAsync(d_X3d, X3d, b_X, cudaMemcpyHostToDevice, h->stream));
        if (!h_lcn.empty()) {
            HULO_CUDA(cudaMemcpyAsync(d_lcn, h_lcn.data(), h_lcn.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemcpyAsync(d_lck, h_lck.data(), h_lck.size() * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        }
        HULO_CUDA(cudaMemcpyAsync(d_desc, h_desc.data(), n_problems * sizeof(ResectDesc), cudaMemcpyHostToDevice, h->stream));

        // ---- waves
        std::vector<uint32_t> active(batched), tri;
        std::vector<uint64_t> trip_off;
        std::vector<WaveSlot> slots;
        std::vector<uint2> ranges;
        std::vector<double> recs;
        while (!active.empty()) {
            const size_t na = active.size();
            trip_off.assign(na + 1, 0);
            for (size_t a = 0; a < na; ++a) trip_off[a + 1] = trip_off[a] + jobs[active[a]].next_T();
            const size_t Ttot = (size_t)trip_off[na];
            HULO_ARG(4 * Ttot < 0xFFFFFFFFull, "too many hypotheses in one wave");
            tri.resize(3 * Ttot);
            parallel_for(na, 16, [&](size_t a) {
                ResectJob &j = jobs[active[a]];
                j.draw(j.next_T(), tri.data() + 3 * trip_off[a], (uint32_t)offsets[active[a]]);
            });
            // slots grouped by size class (points per thread of the register-resident sort)
            ranges.resize(na);
            slots.clear();
            uint32_t cls_begin[6] = {0}, cls_blocks[5] = {0};
            for (int c = 0; c < 5; ++c) {
                cls_begin[c] = (uint32_t)slots.size();
                uint32_t blk = 0;
                for (size_t a = 0; a < na; ++a) {
                    const ResectJob &j = jobs[active[a]];
                    uint32_t np = kScoreThreads;
                    while (np < j.N) np <<= 1;
                    const uint32_t e = np / kScoreThreads;
                    const int cls = e == 1 ? 0 : e == 2 ? 1 : e == 4 ? 2 : e == 8 ? 3 : 4;
                    if (cls != c) continue;
                    const uint32_t nh = (uint32_t)(4 * (trip_off[a + 1] - trip_off[a]));
                    slots.push_back(WaveSlot{active[a], (uint32_t)(4 * trip_off[a]), blk, 0});
                    blk += nh;
                }
                cls_blocks[c] = blk;
            }
            cls_begin[5] = (uint32_t)slots.size();
            for (size_t a = 0; a < na; ++a)
                ranges[a] = make_uint2((uint32_t)(4 * trip_off[a]), (uint32_t)(4 * (trip_off[a + 1] - trip_off[a])));

            const size_t b_tri = ((3 * Ttot * sizeof(uint32_t) + 15) / 16) * 16;
            const size_t b_slots = slots.size() * sizeof(WaveSlot);
            HULO_CUDA(h->scratch3.reserve(b_tri + b_slots + na * sizeof(uint2) + 64));
            uint32_t *d_tri = h->scratch3.as<uint32_t>();
            WaveSlot *d_slots = reinterpret_cast<WaveSlot *>(reinterpret_cast<char *>(d_tri) + b_tri);
            uint2 *d_ranges = reinterpret_cast<uint2 *>(reinterpret_cast<char *>(d_slots) + b_slots);
            HULO_CUDA(h->scratch1.reserve(Ttot * 48 * sizeof(double) + Ttot * sizeof(int32_t)));
            double *d_models = h->scratch1.as<double>();
            int32_t *d_nm = reinterpret_cast<int32_t *>(d_models + Ttot * 48);
            HULO_CUDA(h->scratch2.reserve((4 * Ttot + 14 * na) * sizeof(double)));
            double *d_nfa = h->scratch2.as<double>();
            double *d_rec = d_nfa + 4 * Ttot;
            HULO_CUDA(cudaMemcpyAsync(d_tri, tri.data(), 3 * Ttot * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemcpyAsync(d_slots, slots.data(), b_slots, cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemcpyAsync(d_ranges, ranges.data(), na * sizeof(uint2), cudaMemcpyHostToDevice, h->stream));
            p3p_kernel<<<(unsigned)((Ttot + 127) / 128), 128, 0, h->stream>>>(d_tri, (uint32_t)Ttot, d_x2dn, d_X3d,
                                                                            d_models, d_nm);
            HULO_CUDA(cudaGetLastError());
            h->launches++;
#define HULO_K2_BATCH(C, EE)                                                                                        \
    if (cls_blocks[C] > 0) {                                                                                        \
        score_kernel_reg_batch<EE><<<cls_blocks[C], kScoreThreads, 0, h->stream>>>(                                 \
            d_desc, d_slots + cls_begin[C], cls_begin[C + 1] - cls_begin[C], d_models, d_nfa);                      \
        HULO_CUDA(cudaGetLastError());                                                                              \
        h->launches++;                                                                                              \
    }
            HULO_K2_BATCH(0, 1) HULO_K2_BATCH(1, 2) HULO_K2_BATCH(2, 4) HULO_K2_BATCH(3, 8) HULO_K2_BATCH(4, 16)
#undef HULO_K2_BATCH
            argmin_batch_kernel<<<(unsigned)na, 256, 0, h->stream>>>(d_nfa, d_ranges, d_models, d_rec);
            HULO_CUDA(cudaGetLastError());
            h->launches++;
            recs.resize(14 * na);
            HULO_CUDA(cudaMemcpyAsync(recs.data(), d_rec, 14 * na * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            HULO_CUDA(cudaStreamSynchronize(h->stream));
            parallel_for(na, 4, [&](size_t a) { jobs[active[a]].absorb(recs.data() + 14 * a); });
            size_t keep = 0;
            for (size_t a = 0; a < na; ++a)
                if (jobs[active[a]].next_T() > 0) active[keep++] = active[a];
            active.resize(keep);
        }
        for (uint32_t p : batched) {
            size_t ni = 0;
            int f = 0;
            jobs[p].emit(P + 12 * p, inliers + offsets[p], &ni, error_max + p, &f);
            n_inliers[p] = ni;
            found[p] = f;
        }
    }
    for (uint32_t p : single) {
        size_t ni = 0;
        int f = 0;
        int rc = hulo_resect_acransac(h, x2d + 2 * offsets[p], X3d + 3 * offsets[p], (size_t)(offsets[p + 1] - offsets[p]),
                                      K + 9 * p, max_iter, job_seed(p), P + 12 * p, inliers + offsets[p], &ni,
                                      error_max + p, &f);
        if (rc != HULO_OK) return rc;
        n_inliers[p] = ni;
        found[p] = f;
    }
    return HULO_OK;
}

// The residual key of correspondence i under model M exactly as score_kernel forms it: projection and
// difference in fp64 with the same fma tree, squared norm in fp32; NaN sorts last.
static inline float device_residual(const double *M, const double *X, const double *xn) {
    const double u = std::fma(M[0], X[0], std::fma(M[1], X[1], std::fma(M[2], X[2], M[3])));
    const double v = std::fma(M[4], X[0], std::fma(M[5], X[1], std::fma(M[6], X[2], M[7])));
    const double w = std::fma(M[8], X[0], std::fma(M[9], X[1], std::fma(M[10], X[2], M[11])));
    const float fx = (float)(u / w - xn[0]), fy = (float)(v / w - xn[1]);
    const float e = std::fmaf(fx, fx, fy * fy);
    return e == e ? e : INFINITY;
}

// Host-only self-test of the fp64 rescoring (no device involved): ResectJob::finalize -- radix sort
// of the upper key halves + repair, bracketed NFA scan -- against a comparison sort and the full
// scan with log10 on every term, over seeded correspondence sets with clustered inliers, uniform
// outliers, exact duplicates (ties in the residual) and points on the principal plane (infinite or
// NaN residuals).  The two must agree bit for bit: order, k, the k-th residual and the NFA.
int hulo_selftest_rescoring(uint64_t seed, size_t n_points, size_t n_cases, size_t *n_mismatch) {
    HULO_ARG(n_mismatch != nullptr, "null argument");
    HULO_ARG(n_points <= kMaxPoints, "more than 32768 correspondences");
    *n_mismatch = 0;
    const size_t N = n_points;
    if (N <= 3) return HULO_OK;
    const double K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    const double M[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    std::vector<double> x2d(2 * N), X3d(3 * N);
    uint64_t rng = seed * 0x9E3779B97F4A7C15ULL + 12345;
    auto uni = [&]() { return (double)(splitmix64(rng) >> 11) * (1.0 / 9007199254740992.0); };
    for (size_t c = 0; c < n_cases; ++c) {
        const double inlier_frac = 0.1 + 0.8 * uni(), sigma = std::pow(10.0, -5.0 + 3.0 * uni());
        for (size_t i = 0; i < N; ++i) {
            const double r = uni();
            if (i > 0 && r < 0.05) {                           // exact duplicate of an earlier point
                const size_t j = (size_t)(uni() * (double)i);
                for (int k = 0; k < 3; ++k) X3d[3 * i + k] = X3d[3 * j + k];
                x2d[2 * i] = x2d[2 * j]; x2d[2 * i + 1] = x2d[2 * j + 1];
                continue;
            }
            const double X = 2.0 * uni() - 1.0, Y = 2.0 * uni() - 1.0;
            double Z = 1.0 + uni();
            if (r > 0.98) Z = 0.0;                             // on the principal plane
            const bool inl = uni() < inlier_frac;
            const double dx = inl ? sigma * (uni() - 0.5) : 2.0 * uni() - 1.0;
            const double dy = inl ? sigma * (uni() - 0.5) : 2.0 * uni() - 1.0;
            X3d[3 * i] = X * Z; X3d[3 * i + 1] = Y * Z; X3d[3 * i + 2] = Z;
            x2d[2 * i] = X + dx; x2d[2 * i + 1] = Y + dy;
            if (r > 0.99) { X3d[3 * i] = 0.0; X3d[3 * i + 1] = 0.0; }     // 0 / 0
        }
        ResectJob job;
        job.init(x2d.data(), X3d.data(), N, K, 4096, seed + c);
        const double got = job.finalize(M);
        // the plain form
        std::vector<ResectJob::EI> ref(N);
        for (size_t i = 0; i < N; ++i) {
            const double *X = X3d.data() + 3 * i;
            const double u = M[0] * X[0] + M[1] * X[1] + M[2] * X[2] + M[3];
            const double v = M[4] * X[0] + M[5] * X[1] + M[6] * X[2] + M[7];
            const double w = M[8] * X[0] + M[9] * X[1] + M[10] * X[2] + M[11];
            const double dx = u / w - job.x2dn[2 * i], dy = v / w - job.x2dn[2 * i + 1];
            double e = dx * dx + dy * dy;
            if (!(e == e)) e = INFINITY;
            ref[i] = ResectJob::EI{e, i};
        }
        std::sort(ref.begin(), ref.end(),
                  [](const ResectJob::EI &a, const ResectJob::EI &b) { return a.e < b.e || (a.e == b.e && a.i < b.i); });
        double bn = INFINITY;
        size_t bk = 3;
        for (size_t k = 4; k <= N; ++k) {
            if (!(ref[k - 1].e < INFINITY)) break;
            const double logalpha = job.logalpha0 + log10(ref[k - 1].e + (double)FLT_EPSILON);
            const double nfa = job.loge0 + logalpha * (double)(k - 3) + (double)job.lcn[k] + (double)job.lck[k];
            if (nfa < bn) { bn = nfa; bk = k; }
        }
        bool same = memcmp(&bn, &got, sizeof bn) == 0 && bk == job.best_k;
        const double want_err = bk >= 1 ? ref[bk - 1].e : 0.0;
        same = same && memcmp(&want_err, &job.best_err, sizeof want_err) == 0;
        for (size_t i = 0; same && i < N; ++i)
            same = ref[i].i == job.ei[i].i && memcmp(&ref[i].e, &job.ei[i].e, sizeof(double)) == 0;
        if (!same) ++*n_mismatch;
    }
    return HULO_OK;
}

int hulo_resect_acransac_sequential(hulo_gpu *h, const double *x2d, const double *X3d, size_t N, const double *K,
                                    size_t max_iter, uint64_t seed, double *P, int32_t *inliers, size_t *n_inliers,
                                    double *error_max, int *found) {
    HULO_ARG(h != nullptr && K != nullptr && P != nullptr && n_inliers != nullptr && error_max != nullptr &&
                 found != nullptr, "null argument");
    HULO_ARG(N == 0 || (x2d != nullptr && X3d != nullptr && inliers != nullptr), "null correspondences");
    HULO_ARG(N <= kMaxPoints, "more than 32768 correspondences");
    *n_inliers = 0; *error_max = 0.0; *found = 0;
    if (N <= 3 || max_iter == 0) return HULO_OK;
    HULO_CUDA(cudaSetDevice(h->device));
    std::vector<double> x2dn;
    normalize_points(x2d, N, K, x2dn);
    Problem pb;
    int rc = stage_problem(h, x2dn, X3d, N, pb);
    if (rc != HULO_OK) return rc;

    size_t nIter = max_iter, reserve = nIter / 10;
    nIter -= reserve;
    std::vector<size_t> pool(N);
    for (size_t i = 0; i < N; ++i) pool[i] = i;
    double minNFA = INFINITY, best_model[12] = {0};
    size_t n_best = 0;
    std::vector<uint32_t> tri;
    std::vector<double> h_nfa, h_models;
    std::vector<int32_t> h_k;
    struct EI { float e; uint32_t i; };
    std::vector<EI> ei(N);
    // the inliers of a model at device precision: the n_best smallest (residual, index) keys
    auto sort_keys = [&](const double *M) {
        for (size_t i = 0; i < N; ++i) ei[i] = EI{device_residual(M, X3d + 3 * i, x2dn.data() + 2 * i), (uint32_t)i};
        std::sort(ei.begin(), ei.end(), [](const EI &a, const EI &b) { return a.e < b.e || (a.e == b.e && a.i < b.i); });
    };

    size_t iter = 0, chunk = 64;
    while (iter < nIter) {
        // the next iterations are drawn and scored together, speculatively: their results are
        // committed in order below, and thrown away from the first commit that changes the pool
        const size_t B = std::min(chunk, nIter - iter);
        tri.resize(3 * B);
        for (size_t t = 0; t < B; ++t) {
            uint64_t st = seed + (uint64_t)(iter + t) * 3ull * 0x9E3779B97F4A7C15ULL;   // 3 draws per iteration
            size_t pos[3];
            sample3(st, pool.size(), pos);
            for (int s3 = 0; s3 < 3; ++s3) tri[3 * t + s3] = (uint32_t)pool[pos[s3]];
        }
        HULO_CUDA(h->scratch3.reserve(B * 3 * sizeof(uint32_t)));
        HULO_CUDA(h->scratch1.reserve(B * 48 * sizeof(double) + B * sizeof(int32_t)));
        HULO_CUDA(cudaMemcpyAsync(h->scratch3.ptr, tri.data(), B * 3 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        double *d_models = h->scratch1.as<double>();
        int32_t *d_nm = reinterpret_cast<int32_t *>(d_models + B * 48);
        p3p_kernel<<<(unsigned)((B + 127) / 128), 128, 0, h->stream>>>(h->scratch3.as<uint32_t>(), (uint32_t)B, pb.d_x2dn,
                                                                     pb.d_X3d, d_models, d_nm);
        HULO_CUDA(cudaGetLastError());
        h->launches++;
        ScoreOut o;
        rc = launch_score(h, pb, d_models, 4 * B, -1.0f, o);
        if (rc != HULO_OK) return rc;
        h_nfa.resize(4 * B); h_k.resize(4 * B); h_models.resize(48 * B);
        HULO_CUDA(cudaMemcpyAsync(h_nfa.data(), o.nfa, 4 * B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaMemcpyAsync(h_k.data(), o.k, 4 * B * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaMemcpyAsync(h_models.data(), d_models, 48 * B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaStreamSynchronize(h->stream));
        size_t next_iter = iter + B;
        for (size_t t = 0; t < B; ++t) {
            const size_t it = iter + t;
            bool better = false;
            // the solver compacts its models to the front; absent slots are NaN-marked and score +inf
            for (int m = 0; m < 4; ++m)
                if (h_nfa[4 * t + m] < minNFA) {
                    better = true;
                    minNFA = h_nfa[4 * t + m];
                    n_best = (size_t)h_k[4 * t + m];
                    memcpy(best_model, &h_models[48 * t + 12 * m], sizeof best_model);
                }
            if ((better && minNFA < 0) || (it + 1 == nIter && reserve)) {
                if (n_best == 0) {
                    nIter++;
                    reserve--;
                } else {
                    sort_keys(best_model);
                    pool.resize(n_best);
                    for (size_t i = 0; i < n_best; ++i) pool[i] = ei[i].i;
                    std::sort(pool.begin(), pool.end());        // the pool is kept in index order
                    if (reserve) { nIter = it + 1 + reserve; reserve = 0; }
                    next_iter = it + 1;                          // the rest sampled the old pool
                    break;
                }
            }
        }
        if (next_iter == iter + B) chunk = std::min<size_t>(chunk * 2, 512);
        iter = next_iter;
    }
    if (!(minNFA < 0.0) || n_best == 0) return HULO_OK;
    sort_keys(best_model);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
            P[4 * r + c] = K[3 * r] * best_model[c] + K[3 * r + 1] * best_model[4 + c] + K[3 * r + 2] * best_model[8 + c];
    *error_max = sqrt((double)ei[n_best - 1].e) * K[0];
    for (size_t i = 0; i < n_best; ++i) inliers[i] = (int32_t)ei[i].i;
    *n_inliers = n_best;
    *found = (double)n_best > 2.5 * 3.0 ? 1 : 0;
    return HULO_OK;
}

}  // extern "C"
