// bow.cu -- K5: k nearest map views of a query image in bag-of-features space,
// sm_100a.
//
// Stands behind hulo::selectViewByBoF (BoWCommon/src/BoFUtils.cpp:27-68), the view pre-selection of
// LocalizeEngine::localize (VisionLocalizeServer/src/LocalizeEngine.cc:296-332) and of the CLI
// (localization.cpp:386-412): the reference stacks every view's BoF vector (read from <view>.bow,
// FileUtils.cpp:60-75), builds a FLANN KD-tree index (4 trees, 64 checks -- approximate, OpenCV 3.0,
// not vendored) PER QUERY and takes the knn nearest rows under L2.  Here the matrix is uploaded once
// and stays on the device, and the search is exact: one block per view accumulates the squared L2
// distance to the query in fp32 (coalesced float4 reads: the kernel is a pure HBM stream of
// n x d x 4 bytes), the host picks the knn smallest by (distance, index).
#include <algorithm>
#include <new>
#include <vector>

#include "context.cuh"

struct hulo_bow {
    hulo_gpu *owner = nullptr;
    float *rows = nullptr;       // n x d_pad, zero padded to a multiple of 4
    float *query = nullptr;      // d_pad
    float *dist = nullptr;       // n
    size_t n = 0, d = 0, d_pad = 0;
};

namespace hulo {
namespace {

constexpr int kBowThreads = 256;

__global__ void __launch_bounds__(kBowThreads) bow_dist_kernel(const float4 *__restrict__ rows,
                                                               const float4 *__restrict__ query, uint32_t d4,
                                                               float *__restrict__ dist) {
    __shared__ float s_part[kBowThreads / 32];
    const float4 *row = rows + (size_t)blockIdx.x * d4;
    float acc = 0.0f;
    for (uint32_t k = threadIdx.x; k < d4; k += kBowThreads) {
        const float4 a = __ldg(row + k), q = __ldg(query + k);
        const float x = a.x - q.x, y = a.y - q.y, z = a.z - q.z, w = a.w - q.w;
        acc = fmaf(x, x, fmaf(y, y, fmaf(z, z, fmaf(w, w, acc))));
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < kBowThreads / 32; ++w) t += s_part[w];
        dist[blockIdx.x] = t;
    }
}

}  // namespace
}  // namespace hulo

using namespace hulo;

extern "C" {

void hulo_bow_destroy(hulo_bow *b);

int hulo_bow_create(hulo_gpu *h, const float *bof, size_t n, size_t d, hulo_bow **out) {
    HULO_ARG(h != nullptr && out != nullptr, "null argument");
    *out = nullptr;
    HULO_ARG(n == 0 || bof != nullptr, "null matrix");
    HULO_ARG(d >= 1 && n < (size_t)0x7fffffff && d < (size_t)0x7fffffff, "bad shape");
    HULO_CUDA(cudaSetDevice(h->device));
    hulo_bow *b = new (std::nothrow) hulo_bow();
    HULO_ARG(b != nullptr, "out of host memory");
    b->owner = h; b->n = n; b->d = d; b->d_pad = (d + 3) & ~(size_t)3;
    cudaError_t e = cudaMalloc(&b->rows, std::max<size_t>(n, 1) * b->d_pad * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&b->query, b->d_pad * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(&b->dist, std::max<size_t>(n, 1) * sizeof(float));
    if (e != cudaSuccess) {
        set_error("hulo_bow_create: cudaMalloc -> %s", cudaGetErrorString(e));
        if (b->rows) cudaFree(b->rows);
        if (b->query) cudaFree(b->query);
        if (b->dist) cudaFree(b->dist);
        delete b;
        return HULO_ERR_CUDA;
    }
    if (n) {
        e = cudaMemsetAsync(b->rows, 0, n * b->d_pad * sizeof(float), h->stream);
        if (e == cudaSuccess)
            e = cudaMemcpy2DAsync(b->rows, b->d_pad * sizeof(float), bof, d * sizeof(float), d * sizeof(float), n,
                                  cudaMemcpyHostToDevice, h->stream);
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(b->query, 0, b->d_pad * sizeof(float), h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e != cudaSuccess) {
        set_error("hulo_bow_create: upload -> %s", cudaGetErrorString(e));
        hulo_bow_destroy(b);
        return HULO_ERR_CUDA;
    }
    *out = b;
    return HULO_OK;
}

void hulo_bow_destroy(hulo_bow *b) {
    if (!b) return;
    if (b->owner) cudaSetDevice(b->owner->device);
    if (b->rows) cudaFree(b->rows);
    if (b->query) cudaFree(b->query);
    if (b->dist) cudaFree(b->dist);
    delete b;
}

int hulo_bow_knn(hulo_bow *b, const float *query, const uint32_t *subset, size_t n_subset, size_t knn,
                 int32_t *idx, float *dist) {
    HULO_ARG(b != nullptr && query != nullptr, "null argument");
    HULO_ARG(knn == 0 || idx != nullptr, "null output");
    hulo_gpu *h = b->owner;
    const size_t n_cand = subset ? n_subset : b->n;
    // CV_Assert(knn < viewList.size()), BoFUtils.cpp:30
    HULO_ARG(knn < n_cand || (knn == 0 && n_cand == 0), "knn must be smaller than the number of candidate views");
    for (size_t k = 0; subset && k < n_subset; ++k) HULO_ARG(subset[k] < b->n, "view index out of range");
    if (knn == 0) return HULO_OK;
    HULO_CUDA(cudaSetDevice(h->device));
    HULO_CUDA(cudaMemcpyAsync(b->query, query, b->d * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    bow_dist_kernel<<<(unsigned)b->n, kBowThreads, 0, h->stream>>>(reinterpret_cast<const float4 *>(b->rows),
                                                                  reinterpret_cast<const float4 *>(b->query),
                                                                  (uint32_t)(b->d_pad / 4), b->dist);
    HULO_CUDA(cudaGetLastError());
    h->launches++;
    std::vector<float> all(b->n);
    HULO_CUDA(cudaMemcpyAsync(all.data(), b->dist, b->n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    std::vector<std::pair<float, int32_t>> cand(n_cand);
    for (size_t k = 0; k < n_cand; ++k) {
        const int32_t v = subset ? (int32_t)subset[k] : (int32_t)k;
        cand[k] = std::make_pair(all[(size_t)v], v);
    }
    std::partial_sort(cand.begin(), cand.begin() + knn, cand.end());      // (distance, index) ascending
    for (size_t k = 0; k < knn; ++k) {
        idx[k] = cand[k].second;
        if (dist) dist[k] = cand[k].first;
    }
    return HULO_OK;
}

}  // extern "C"
