// match_post.cu -- see match_post.cuh.
#include "match_post.cuh"

#include <climits>

#include "../../include/hulo_gpu.h"
#include "knn2.cuh"

namespace hulo {
namespace {

constexpr uint64_t kNone64 = ~0ull;

__device__ __forceinline__ void top2_insert(uint64_t k, uint64_t &m0, uint64_t &m1) {
    const uint64_t hi = k > m0 ? k : m0;
    m0 = k < m0 ? k : m0;
    m1 = hi < m1 ? hi : m1;
}

// (0.0f + d0) / d1 < ratio in IEEE float32, then d1 < INT_MAX  (MatchUtils.cpp:347-349).
// 0/0 -> NaN -> false.
__device__ __forceinline__ bool ratio_pass(int32_t d0, int32_t d1, float ratio) {
    const float q = __fdiv_rn(__fadd_rn(0.0f, __int2float_rn(d0)), __int2float_rn(d1));
    return (q < ratio) && (d1 < INT_MAX);
}

// largest s with off[s] <= row  (off ascending, off[0] == 0, n + 1 entries; duplicates allowed)
__device__ __forceinline__ uint32_t seg_of_row(const uint64_t *__restrict__ off, uint32_t n, uint64_t row) {
    uint32_t lo = 0, hi = n;   // invariant: off[lo] <= row < off[hi]
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (off[mid] <= row) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void post_query_kernel(const uint2 *__restrict__ partial, uint32_t n_rows, uint32_t n_chunks,
                                  uint64_t slot_stride, uint32_t rows_per_chunk, float ratio,
                                  int32_t *__restrict__ val, int32_t *__restrict__ dist) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    uint64_t m0 = kNone64, m1 = kNone64;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint2 k = partial[(uint64_t)c * slot_stride + row];
        const uint32_t base = c * rows_per_chunk;
        if (k.x != kKeyNone)
            top2_insert(((uint64_t)(k.x >> kKeyIdxBits) << 32) | (uint64_t)(base + (k.x & kKeyIdxMask)), m0, m1);
        if (k.y != kKeyNone)
            top2_insert(((uint64_t)(k.y >> kKeyIdxBits) << 32) | (uint64_t)(base + (k.y & kKeyIdxMask)), m0, m1);
    }
    const int32_t d0 = m0 == kNone64 ? INT_MAX : (int32_t)(m0 >> 32);
    const int32_t d1 = m1 == kNone64 ? INT_MAX : (int32_t)(m1 >> 32);
    const bool ok = ratio_pass(d0, d1, ratio);
    val[row] = ok ? (int32_t)(uint32_t)m0 : -1;
    dist[row] = d0;
}

__global__ void post_pair_claim_kernel(const uint2 *__restrict__ partial, uint32_t n_rows,
                                       const uint64_t *__restrict__ row_off,
                                       const uint64_t *__restrict__ hist_off, uint32_t n_pairs, float ratio,
                                       int32_t *__restrict__ val, uint32_t *__restrict__ hist) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint2 k = partial[row];
    const int32_t d0 = k.x == kKeyNone ? INT_MAX : (int32_t)(k.x >> kKeyIdxBits);
    const int32_t d1 = k.y == kKeyNone ? INT_MAX : (int32_t)(k.y >> kKeyIdxBits);
    int32_t m = -1;
    if (ratio_pass(d0, d1, ratio)) m = (int32_t)(k.x & kKeyIdxMask);
    val[row] = m;
    if (m >= 0) {
        const uint32_t p = seg_of_row(row_off, n_pairs, row);
        atomicAdd(&hist[hist_off[p] + (uint32_t)m], 1u);
    }
}

__global__ void post_pair_filter_kernel(uint32_t n_rows, const uint64_t *__restrict__ row_off,
                                        const uint64_t *__restrict__ hist_off, uint32_t n_pairs,
                                        unsigned flags, int32_t *__restrict__ val,
                                        const uint32_t *__restrict__ hist) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const int32_t m = val[row];
    if (m < 0) return;
    const uint32_t p = seg_of_row(row_off, n_pairs, row);
    bool keep = true;
    if ((flags & HULO_PAIR_ONE_TO_ONE) && hist[hist_off[p] + (uint32_t)m] != 1u) keep = false;
    if ((flags & HULO_PAIR_DROP_LAST) && (uint64_t)row + 1 == row_off[p + 1]) keep = false;
    if (!keep) val[row] = -1;
}

// ---- compaction: count per block, scan the block counts, write
constexpr int kCompactThreads = 256;
constexpr int kRowsPerThread = kCompactBlockRows / kCompactThreads;   // 8 consecutive rows

__global__ void compact_count_kernel(const int32_t *__restrict__ val, uint32_t n_rows,
                                     uint32_t *__restrict__ block_counts) {
    __shared__ uint32_t s_warp[kCompactThreads / 32];
    const uint32_t base = blockIdx.x * kCompactBlockRows + threadIdx.x * kRowsPerThread;
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kRowsPerThread; ++k) {
        const uint32_t r = base + k;
        if (r < n_rows && val[r] >= 0) ++c;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kCompactThreads / 32; ++w) t += s_warp[w];
        block_counts[blockIdx.x] = t;
    }
}

// single block: exclusive scan of block_counts[0..n_blocks) in place; total -> block_counts[n_blocks], *total
__global__ void compact_scan_kernel(uint32_t *block_counts, uint32_t n_blocks, uint64_t *total) {
    __shared__ uint32_t s_part[1024];
    const uint32_t per = (n_blocks + blockDim.x - 1) / blockDim.x;
    const uint32_t lo = min(threadIdx.x * per, n_blocks), hi = min(lo + per, n_blocks);
    uint32_t sum = 0;
    for (uint32_t i = lo; i < hi; ++i) sum += block_counts[i];
    s_part[threadIdx.x] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (uint32_t t = 0; t < blockDim.x; ++t) {
            const uint32_t v = s_part[t];
            s_part[t] = run;
            run += v;
        }
        block_counts[n_blocks] = run;
        *total = run;
    }
    __syncthreads();
    uint32_t run = s_part[threadIdx.x];
    for (uint32_t i = lo; i < hi; ++i) {
        const uint32_t v = block_counts[i];
        block_counts[i] = run;
        run += v;
    }
}

// SELF_SCAN: block_offsets holds the raw per-block counts and the block sums its predecessors
// itself (small launches: saves the scan kernel); otherwise block_offsets is the exclusive scan.
// seg_out_off[s] = number of survivors in compact rows [0, seg_off[s]): written by the thread whose
// rows hold the boundary (its exclusive prefix + the survivors of its own rows before it);
// boundaries at or beyond n_rows get the total from the last block.
template <bool SELF_SCAN>
__global__ void compact_write_kernel(const int32_t *__restrict__ val, const int32_t *__restrict__ dist,
                                     uint32_t n_rows, const uint64_t *__restrict__ seg_off, uint32_t n_seg,
                                     const uint32_t *__restrict__ block_offsets, uint32_t *__restrict__ out_seg,
                                     uint32_t *__restrict__ out_i, uint32_t *__restrict__ out_j,
                                     int32_t *__restrict__ out_d, uint64_t *__restrict__ seg_out_off,
                                     uint64_t *__restrict__ total) {
    __shared__ uint32_t s_warp[kCompactThreads / 32];
    __shared__ uint32_t s_pre[kCompactThreads / 32];
    const uint32_t base = blockIdx.x * kCompactBlockRows + threadIdx.x * kRowsPerThread;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t block_base;
    if (SELF_SCAN) {
        uint32_t pre = 0;
        for (uint32_t b = threadIdx.x; b < blockIdx.x; b += kCompactThreads) pre += block_offsets[b];
        for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
        if (lane == 0) s_pre[warp] = pre;
        __syncthreads();
        block_base = 0;
        for (int w = 0; w < kCompactThreads / 32; ++w) block_base += s_pre[w];
    } else {
        block_base = block_offsets[blockIdx.x];
    }
    int32_t v[kRowsPerThread];
    uint32_t c = 0;
#pragma unroll
    for (int k = 0; k < kRowsPerThread; ++k) {
        const uint32_t r = base + k;
        v[k] = r < n_rows ? val[r] : -1;
        if (v[k] >= 0) ++c;
    }
    // exclusive scan of c over the block
    uint32_t inc = c;
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t warp_base = 0, block_count = 0;
    for (int w = 0; w < kCompactThreads / 32; ++w) {
        if (w < warp) warp_base += s_warp[w];
        block_count += s_warp[w];
    }
    const uint32_t pos0 = block_base + warp_base + inc - c;
    uint32_t pos = pos0;
#pragma unroll
    for (int k = 0; k < kRowsPerThread; ++k) {
        if (v[k] >= 0) {
            const uint32_t r = base + k;
            const uint32_t s = seg_of_row(seg_off, n_seg, r);
            if (out_seg) out_seg[pos] = s;
            out_i[pos] = (uint32_t)(r - seg_off[s]);
            out_j[pos] = (uint32_t)v[k];
            if (out_d) out_d[pos] = dist[r];
            ++pos;
        }
    }
    if (seg_out_off) {
        if (base < n_rows) {
            // first boundary at or after this thread's first row
            uint32_t lo = 0, hi = n_seg + 1;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (seg_off[mid] < base) lo = mid + 1; else hi = mid;
            }
            const uint64_t end = min((uint64_t)base + kRowsPerThread, (uint64_t)n_rows);
            for (uint32_t s = lo; s <= n_seg && seg_off[s] < end; ++s) {
                const uint32_t upto = (uint32_t)(seg_off[s] - base);
                uint32_t before = 0;
#pragma unroll
                for (int k = 0; k < kRowsPerThread; ++k)
                    if ((uint32_t)k < upto && v[k] >= 0) ++before;
                seg_out_off[s] = (uint64_t)pos0 + before;
            }
        }
        if (blockIdx.x == gridDim.x - 1) {
            uint32_t lo = 0, hi = n_seg + 1;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (seg_off[mid] < n_rows) lo = mid + 1; else hi = mid;
            }
            for (uint32_t s = lo + threadIdx.x; s <= n_seg; s += kCompactThreads)
                seg_out_off[s] = (uint64_t)block_base + block_count;
        }
    }
    if (SELF_SCAN && total && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *total = (uint64_t)block_base + block_count;
}

}  // namespace

cudaError_t post_query_launch(const uint2 *partial, uint32_t n_rows, uint32_t n_chunks, uint64_t slot_stride,
                              uint32_t rows_per_chunk, float ratio, int32_t *val, int32_t *dist,
                              cudaStream_t stream) {
    if (n_rows == 0) return cudaSuccess;
    post_query_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(partial, n_rows, n_chunks, slot_stride,
                                                              rows_per_chunk, ratio, val, dist);
    return cudaGetLastError();
}

cudaError_t post_pair_claim_launch(const uint2 *partial, uint32_t n_rows, const uint64_t *row_off,
                                   const uint64_t *hist_off, uint32_t n_pairs, float ratio, int32_t *val,
                                   uint32_t *hist, cudaStream_t stream) {
    if (n_rows == 0) return cudaSuccess;
    post_pair_claim_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(partial, n_rows, row_off, hist_off,
                                                                   n_pairs, ratio, val, hist);
    return cudaGetLastError();
}

cudaError_t post_pair_filter_launch(uint32_t n_rows, const uint64_t *row_off, const uint64_t *hist_off,
                                    uint32_t n_pairs, unsigned flags, int32_t *val, const uint32_t *hist,
                                    cudaStream_t stream) {
    if (n_rows == 0) return cudaSuccess;
    post_pair_filter_kernel<<<(n_rows + 255) / 256, 256, 0, stream>>>(n_rows, row_off, hist_off, n_pairs,
                                                                    flags, val, hist);
    return cudaGetLastError();
}

cudaError_t compact_launch(const int32_t *val, const int32_t *dist, uint32_t n_rows, const uint64_t *seg_off,
                           uint32_t n_seg, uint32_t *block_counts, uint32_t *out_seg, uint32_t *out_i,
                           uint32_t *out_j, int32_t *out_d, uint64_t *seg_out_off, uint64_t *total,
                           cudaStream_t stream, int *n_launches) {
    const uint32_t n_blocks = (n_rows + kCompactBlockRows - 1) / kCompactBlockRows;
    int launches = 0;
    if (n_blocks == 0) {
        // nothing to compact: total and every segment offset are zero
        cudaError_t e = cudaMemsetAsync(total, 0, sizeof(uint64_t), stream);
        if (e != cudaSuccess) return e;
        if (seg_out_off) e = cudaMemsetAsync(seg_out_off, 0, ((size_t)n_seg + 1) * sizeof(uint64_t), stream);
        if (n_launches) *n_launches = 0;
        return e;
    }
    compact_count_kernel<<<n_blocks, kCompactThreads, 0, stream>>>(val, n_rows, block_counts);
    ++launches;
    if (n_blocks <= 1024) {
        // few blocks: each one sums the counts of its predecessors itself
        compact_write_kernel<true><<<n_blocks, kCompactThreads, 0, stream>>>(val, dist, n_rows, seg_off, n_seg, block_counts,
                                                                            out_seg, out_i, out_j, out_d, seg_out_off, total);
        ++launches;
    } else {
        compact_scan_kernel<<<1, 1024, 0, stream>>>(block_counts, n_blocks, total);
        compact_write_kernel<false><<<n_blocks, kCompactThreads, 0, stream>>>(val, dist, n_rows, seg_off, n_seg, block_counts,
                                                                             out_seg, out_i, out_j, out_d, seg_out_off, total);
        launches += 2;
    }
    if (n_launches) *n_launches = launches;
    return cudaGetLastError();
}

}  // namespace hulo
