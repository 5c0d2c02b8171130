// merge3d.cu -- K4: RANSAC of a 3D-3D transform between two models, sm_100a.
//
// Stands behind ransacAffineTransform (PyVisionLocalizeCommon/src/hulo_sfm/mergeSfM.py:344-388) and
// ransacSimilarityTransform (PyVisionLocalizeCommon/src/hulo_transform/ransacTransform.py:13-49),
// which mergeSfM.ransacTransform (:394-399) runs with ransacRound = 100 x #matches (:577) when two
// models are merged and when a model is anchored to world coordinates
// (localizeGlobalCoordinate.py:209, measureAccuracy.py:239): find the 3 x 4 matrix M with
// A ~ M [B; 1] from 4-point samples, score every sample by the number of points within `thres`,
// keep the best one whose linear part is well conditioned (singular value ratio < svdRatio), and
// refit on its inliers.  The reference is a Python loop with one lstsq / SVD and one 3 x n product
// per round; here every round is a hypothesis scored in parallel.
//   hypotheses_kernel  one thread per round: the 4 x 4 system [B_sel; 1]^T X = A_sel^T by Gaussian
//                      elimination with partial pivoting (affine), or Kabsch + scale from a one-sided
//                      Jacobi SVD of the 3 x 3 covariance (similarity); fp64
//   count_kernel       one warp per hypothesis: || M [B_i; 1] - A_i || < thres over all points, fp64
// The sequential selection rule (strictly more inliers than the best so far AND conditioned) and the
// refit run on the host over the counts; samples can be passed in (tests replay the reference's own
// random.sample sequence) or are drawn from a counter-based generator.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "context.cuh"

namespace hulo {
namespace {

constexpr int kMergeThreads = 128;

__device__ __forceinline__ uint64_t mix64m(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// One-sided Jacobi SVD of a 3 x 3 matrix H = U diag(s) V^T (columns of U, V in u, v).
__host__ __device__ inline void svd3(const double H[9], double U[9], double S[3], double V[9]) {
    double a[9];
    for (int k = 0; k < 9; ++k) { a[k] = H[k]; V[k] = (k % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < 3; ++r) {
                    alpha += a[3 * r + p] * a[3 * r + p];
                    beta += a[3 * r + q] * a[3 * r + q];
                    gamma += a[3 * r + p] * a[3 * r + q];
                }
                off = fmax(off, fabs(gamma) / sqrt(fmax(alpha * beta, 1e-300)));
                if (gamma == 0.0) continue;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < 3; ++r) {
                    const double x = a[3 * r + p], y = a[3 * r + q];
                    a[3 * r + p] = c * x - s * y;
                    a[3 * r + q] = s * x + c * y;
                    const double vx = V[3 * r + p], vy = V[3 * r + q];
                    V[3 * r + p] = c * vx - s * vy;
                    V[3 * r + q] = s * vx + c * vy;
                }
            }
        if (off < 1e-15) break;
    }
    for (int q = 0; q < 3; ++q) {
        double n = 0;
        for (int r = 0; r < 3; ++r) n += a[3 * r + q] * a[3 * r + q];
        n = sqrt(n);
        S[q] = n;
        for (int r = 0; r < 3; ++r) U[3 * r + q] = n > 0 ? a[3 * r + q] / n : 0.0;
    }
}

// superimposition_matrix(v0 = B points, v1 = A points, scale=True) from the centred moments:
// H = sum (a - ma)(b - mb)^T, sa / sb the centred sums of squares.  R = U V^T of the SVD of H with
// the direction of the smallest singular value flipped when det < 0 (Kabsch), scale sqrt(sa / sb).
__host__ __device__ inline bool similarity_from_moments(const double H[9], double sa, double sb, const double ma[3],
                                                        const double mb[3], double *M) {
    if (!(sb > 0.0)) return false;
    double U[9], S[3], V[9];
    svd3(H, U, S, V);
    int smallest = 0;
    for (int q = 1; q < 3; ++q)
        if (S[q] < S[smallest]) smallest = q;
    double R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[3 * i + j] = U[3 * i] * V[3 * j] + U[3 * i + 1] * V[3 * j + 1] + U[3 * i + 2] * V[3 * j + 2];
    const double det = R[0] * (R[4] * R[8] - R[5] * R[7]) - R[1] * (R[3] * R[8] - R[5] * R[6]) + R[2] * (R[3] * R[7] - R[4] * R[6]);
    if (det < 0.0)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[3 * i + j] -= 2.0 * U[3 * i + smallest] * V[3 * j + smallest];
    const double sc = sqrt(sa / sb);
    for (int i = 0; i < 3; ++i) {
        double t = ma[i];
        for (int j = 0; j < 3; ++j) { M[4 * i + j] = sc * R[3 * i + j]; t -= sc * R[3 * i + j] * mb[j]; }
        M[4 * i + 3] = t;
    }
    return true;
}

// models: rounds x 12 (row-major 3 x 4); an unusable sample has NaN in entry 0.
__global__ void hypotheses_kernel(const double *__restrict__ A, const double *__restrict__ B, uint32_t n,
                                  const uint32_t *__restrict__ samples, uint32_t rounds, uint64_t seed, int similarity,
                                  double *__restrict__ models, uint32_t *__restrict__ samples_out) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rounds) return;
    uint32_t sel[4];
    if (samples) {
        for (int k = 0; k < 4; ++k) sel[k] = samples[4 * r + k];
    } else {
        // four distinct indices (random.sample(listInd, 4)): counter-based draws, rejection of repeats
        uint64_t s = seed + (uint64_t)r * 0x9E3779B97F4A7C15ULL * 8ull;
        int got = 0;
        while (got < 4) {
            s += 0x9E3779B97F4A7C15ULL;
            const uint32_t c = (uint32_t)(mix64m(s) % n);
            bool dup = false;
            for (int k = 0; k < got; ++k) dup = dup || sel[k] == c;
            if (!dup) sel[got++] = c;
        }
    }
    if (samples_out)
        for (int k = 0; k < 4; ++k) samples_out[4 * r + k] = sel[k];
    double *M = models + 12 * (size_t)r;
    double a[4][3], b[4][3];
    for (int k = 0; k < 4; ++k)
        for (int c = 0; c < 3; ++c) { a[k][c] = A[(size_t)c * n + sel[k]]; b[k][c] = B[(size_t)c * n + sel[k]]; }
    if (!similarity) {
        // [b_k 1] X = a_k for the four points: augmented 4 x 7 system, partial pivoting
        double T[4][7];
        for (int k = 0; k < 4; ++k) {
            T[k][0] = b[k][0]; T[k][1] = b[k][1]; T[k][2] = b[k][2]; T[k][3] = 1.0;
            T[k][4] = a[k][0]; T[k][5] = a[k][1]; T[k][6] = a[k][2];
        }
        bool ok = true;
        for (int k = 0; k < 4 && ok; ++k) {
            int piv = k;
            for (int rr = k + 1; rr < 4; ++rr)
                if (fabs(T[rr][k]) > fabs(T[piv][k])) piv = rr;
            if (!(fabs(T[piv][k]) > 1e-300)) { ok = false; break; }
            if (piv != k)
                for (int c = 0; c < 7; ++c) { const double t = T[k][c]; T[k][c] = T[piv][c]; T[piv][c] = t; }
            const double inv = 1.0 / T[k][k];
            for (int rr = 0; rr < 4; ++rr) {
                if (rr == k) continue;
                const double f = T[rr][k] * inv;
                for (int c = k; c < 7; ++c) T[rr][c] -= f * T[k][c];
            }
        }
        if (!ok) { M[0] = NAN; return; }
        // X (4 x 3): X[k][c] = T[k][4 + c] / T[k][k];  M = X^T
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < 4; ++k) M[4 * c + k] = T[k][4 + c] / T[k][k];
    } else {
        // superimposition_matrix(B_sel, A_sel, scale=True): Kabsch rotation + scale + translation
        double mb[3] = {0, 0, 0}, ma[3] = {0, 0, 0};
        for (int k = 0; k < 4; ++k)
            for (int c = 0; c < 3; ++c) { mb[c] += b[k][c] * 0.25; ma[c] += a[k][c] * 0.25; }
        double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, sa = 0, sb = 0;
        for (int k = 0; k < 4; ++k) {
            double da[3], db[3];
            for (int c = 0; c < 3; ++c) { da[c] = a[k][c] - ma[c]; db[c] = b[k][c] - mb[c]; sa += da[c] * da[c]; sb += db[c] * db[c]; }
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) H[3 * i + j] += da[i] * db[j];      // v1 v0^T
        }
        if (!similarity_from_moments(H, sa, sb, ma, mb, M)) { M[0] = NAN; return; }
    }
    bool fin = true;
    for (int k = 0; k < 12; ++k) fin = fin && isfinite(M[k]);
    if (!fin) M[0] = NAN;
}

// One warp per hypothesis: number of points with || M [B_i; 1] - A_i || < thres.
__global__ void __launch_bounds__(kMergeThreads) count_kernel(const double *__restrict__ A, const double *__restrict__ B,
                                                             uint32_t n, const double *__restrict__ models,
                                                             uint32_t rounds, double thres, uint32_t *__restrict__ counts) {
    const uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (h >= rounds) return;
    const double *M = models + 12 * (size_t)h;
    double m[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) m[k] = M[k];
    uint32_t cnt = 0;
    if (m[0] == m[0]) {
        for (uint32_t i = lane; i < n; i += 32) {
            const double x = B[i], y = B[n + i], z = B[2 * (size_t)n + i];
            const double dx = (m[0] * x + m[1] * y + m[2] * z + m[3]) - A[i];
            const double dy = (m[4] * x + m[5] * y + m[6] * z + m[7]) - A[n + i];
            const double dz = (m[8] * x + m[9] * y + m[10] * z + m[11]) - A[2 * (size_t)n + i];
            cnt += sqrt(dx * dx + dy * dy + dz * dz) < thres ? 1u : 0u;
        }
    }
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) counts[h] = cnt;
}

// singular values of the 3 x 3 linear part (Jacobi on M^T M), descending
void singular_values3(const double *M, double s[3]) {
    double a[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            a[3 * i + j] = 0;
            for (int k = 0; k < 3; ++k) a[3 * i + j] += M[4 * k + i] * M[4 * k + j];
        }
    for (int sweep = 0; sweep < 50; ++sweep) {
        const double off = fabs(a[1]) + fabs(a[2]) + fabs(a[5]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (a[3 * p + q] == 0.0) continue;
                const double th = (a[3 * q + q] - a[3 * p + p]) / (2.0 * a[3 * p + q]);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double x = a[3 * k + p], y = a[3 * k + q];
                    a[3 * k + p] = c * x - sn * y;
                    a[3 * k + q] = sn * x + c * y;
                }
                for (int k = 0; k < 3; ++k) {
                    const double x = a[3 * p + k], y = a[3 * q + k];
                    a[3 * p + k] = c * x - sn * y;
                    a[3 * q + k] = sn * x + c * y;
                }
            }
    }
    s[0] = sqrt(std::max(a[0], 0.0)); s[1] = sqrt(std::max(a[4], 0.0)); s[2] = sqrt(std::max(a[8], 0.0));
    std::sort(s, s + 3, [](double x, double y) { return x > y; });
}

// least squares M (3 x 4) with A_inl ~ M [B_inl; 1]: normal equations in long double
bool refit_affine(const double *A, const double *B, size_t n, const std::vector<int32_t> &inl, double *M) {
    long double G[4][4] = {{0}}, rhs[4][3] = {{0}};
    for (int32_t i : inl) {
        const long double b[4] = {B[i], B[n + i], B[2 * n + i], 1.0L};
        const long double a[3] = {A[i], A[n + i], A[2 * n + i]};
        for (int r = 0; r < 4; ++r) {
            for (int c = 0; c < 4; ++c) G[r][c] += b[r] * b[c];
            for (int c = 0; c < 3; ++c) rhs[r][c] += b[r] * a[c];
        }
    }
    for (int k = 0; k < 4; ++k) {
        int piv = k;
        for (int r = k + 1; r < 4; ++r)
            if (fabsl(G[r][k]) > fabsl(G[piv][k])) piv = r;
        if (!(fabsl(G[piv][k]) > 0)) return false;
        if (piv != k) {
            for (int c = 0; c < 4; ++c) std::swap(G[k][c], G[piv][c]);
            for (int c = 0; c < 3; ++c) std::swap(rhs[k][c], rhs[piv][c]);
        }
        for (int r = 0; r < 4; ++r) {
            if (r == k) continue;
            const long double f = G[r][k] / G[k][k];
            for (int c = k; c < 4; ++c) G[r][c] -= f * G[k][c];
            for (int c = 0; c < 3; ++c) rhs[r][c] -= f * rhs[k][c];
        }
    }
    for (int c = 0; c < 3; ++c)
        for (int k = 0; k < 4; ++k) M[4 * c + k] = (double)(rhs[k][c] / G[k][k]);
    return true;
}

}  // namespace
}  // namespace hulo

using namespace hulo;

extern "C" {

int hulo_ransac_transform3d(hulo_gpu *h, const double *A, const double *B, size_t n, double thres,
                            const uint32_t *samples, size_t rounds, uint64_t seed, double svd_ratio, int similarity,
                            double *M, int32_t *inliers, size_t *n_inliers, uint32_t *best_round) {
    HULO_ARG(h != nullptr && M != nullptr && n_inliers != nullptr, "null argument");
    *n_inliers = 0;
    if (best_round) *best_round = 0xFFFFFFFFu;
    for (int k = 0; k < 12; ++k) M[k] = 0.0;
    HULO_ARG(n == 0 || (A != nullptr && B != nullptr && inliers != nullptr), "null points");
    HULO_ARG(n < (size_t)0x7fffffff && rounds < (size_t)0x7fffffff, "too many points or rounds");
    // random.sample(listInd, 4) raises for fewer than four points; nothing to estimate either way
    if (n < 4 || rounds == 0) return HULO_OK;
    if (samples)
        for (size_t k = 0; k < 4 * rounds; ++k) HULO_ARG(samples[k] < n, "sample index out of range");
    HULO_CUDA(cudaSetDevice(h->device));
    // scratch0: A | B ; scratch1: models ; scratch2: counts ; scratch3: samples
    HULO_CUDA(h->scratch0.reserve(6 * n * sizeof(double)));
    HULO_CUDA(h->scratch1.reserve(rounds * 12 * sizeof(double)));
    HULO_CUDA(h->scratch2.reserve(rounds * sizeof(uint32_t)));
    HULO_CUDA(h->scratch3.reserve(rounds * 4 * sizeof(uint32_t)));
    double *d_A = h->scratch0.as<double>(), *d_B = d_A + 3 * n;
    HULO_CUDA(cudaMemcpyAsync(d_A, A, 3 * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(cudaMemcpyAsync(d_B, B, 3 * n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    uint32_t *d_samples = h->scratch3.as<uint32_t>();
    if (samples) HULO_CUDA(cudaMemcpyAsync(d_samples, samples, rounds * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    hypotheses_kernel<<<(unsigned)((rounds + 127) / 128), 128, 0, h->stream>>>(
        d_A, d_B, (uint32_t)n, samples ? d_samples : nullptr, (uint32_t)rounds, seed, similarity, h->scratch1.as<double>(),
        samples ? nullptr : d_samples);
    HULO_CUDA(cudaGetLastError());
    count_kernel<<<(unsigned)((rounds * 32 + kMergeThreads - 1) / kMergeThreads), kMergeThreads, 0, h->stream>>>(
        d_A, d_B, (uint32_t)n, h->scratch1.as<double>(), (uint32_t)rounds, thres, h->scratch2.as<uint32_t>());
    HULO_CUDA(cudaGetLastError());
    h->launches += 2;
    std::vector<uint32_t> counts(rounds);
    HULO_CUDA(cudaMemcpyAsync(counts.data(), h->scratch2.ptr, rounds * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));

    // the sequential rule: a round replaces the best so far iff it has strictly more inliers and
    // its linear part is conditioned (s_max / s_min < svdRatio); models are fetched only for the
    // rounds that beat the running count
    uint32_t best = 0;
    long best_r = -1;
    double Mb[12] = {0}, Mr[12];
    // the models come over in one copy the first time one is needed (an ill-conditioned leader would
    // otherwise cost one blocking 96-byte copy per round that beats the running count)
    std::vector<double> all_models;
    for (size_t r = 0; r < rounds; ++r) {
        if (counts[r] <= best) continue;
        if (all_models.empty()) {
            all_models.resize(12 * rounds);
            HULO_CUDA(cudaMemcpyAsync(all_models.data(), h->scratch1.ptr, 12 * rounds * sizeof(double), cudaMemcpyDeviceToHost,
                                      h->stream));
            HULO_CUDA(cudaStreamSynchronize(h->stream));
        }
        memcpy(Mr, &all_models[12 * r], sizeof Mr);
        if (!(Mr[0] == Mr[0])) continue;
        double s[3];
        singular_values3(Mr, s);
        if (!(s[0] / s[2] < svd_ratio)) continue;
        best = counts[r];
        best_r = (long)r;
        memcpy(Mb, Mr, sizeof Mb);
    }
    if (best_r < 0) return HULO_OK;
    std::vector<int32_t> inl;
    for (size_t i = 0; i < n; ++i) {
        const double x = B[i], y = B[n + i], z = B[2 * n + i];
        const double dx = (Mb[0] * x + Mb[1] * y + Mb[2] * z + Mb[3]) - A[i];
        const double dy = (Mb[4] * x + Mb[5] * y + Mb[6] * z + Mb[7]) - A[n + i];
        const double dz = (Mb[8] * x + Mb[9] * y + Mb[10] * z + Mb[11]) - A[2 * n + i];
        if (sqrt(dx * dx + dy * dy + dz * dz) < thres) inl.push_back((int32_t)i);
    }
    if (inl.size() < 4) return HULO_OK;                 // `if len(inliers) < 4: return [], []`
    if (best_round) *best_round = (uint32_t)best_r;
    if (!similarity) {
        if (!refit_affine(A, B, n, inl, M)) return HULO_OK;
    } else {
        // superimposition_matrix over the inliers: the same closed form from their centred moments
        double mb[3] = {0, 0, 0}, ma[3] = {0, 0, 0};
        for (int32_t i : inl)
            for (int c = 0; c < 3; ++c) { mb[c] += B[c * n + i]; ma[c] += A[c * n + i]; }
        for (int c = 0; c < 3; ++c) { mb[c] /= (double)inl.size(); ma[c] /= (double)inl.size(); }
        double H[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, sa = 0, sb = 0;
        for (int32_t i : inl) {
            double da[3], db[3];
            for (int c = 0; c < 3; ++c) { da[c] = A[c * n + i] - ma[c]; db[c] = B[c * n + i] - mb[c]; sa += da[c] * da[c]; sb += db[c] * db[c]; }
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) H[3 * a + b] += da[a] * db[b];
        }
        if (!similarity_from_moments(H, sa, sb, ma, mb, M)) return HULO_OK;
    }
    memcpy(inliers, inl.data(), inl.size() * sizeof(int32_t));
    *n_inliers = inl.size();
    return HULO_OK;
}

}  // extern "C"
