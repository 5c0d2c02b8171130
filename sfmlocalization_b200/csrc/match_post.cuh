// match_post.cuh -- device post-processing between K1 and the reference's result containers:
// ratio test (MatchUtils.cpp:113-116, 347-349), one-to-one filter (:125-143), last-row quirk
// (:125, :146) and order-preserving compaction of the survivors.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hulo {

// Query mode.  For every compact searcher row r (rows of the selected views, concatenated):
// fold its n_chunks partial keys, apply the ratio test and store val[r] = j (query feature)
// or -1, dist[r] = d0.
cudaError_t post_query_launch(const uint2 *partial, uint32_t n_rows, uint32_t n_chunks, uint64_t slot_stride,
                              uint32_t rows_per_chunk, float ratio, int32_t *val, int32_t *dist,
                              cudaStream_t stream);

// Pair mode, step 1: per compact row (rows of image I of every pair, concatenated) decode the
// single-chunk keys, ratio test -> val[r] = j or -1, and count claimants per train row in
// hist[hist_off(pair) + j].  pair_of_row is found by binary search in row_off (n_pairs+1).
cudaError_t post_pair_claim_launch(const uint2 *partial, uint32_t n_rows, const uint64_t *row_off,
                                   const uint64_t *hist_off, uint32_t n_pairs, float ratio, int32_t *val,
                                   uint32_t *hist, cudaStream_t stream);
// Pair mode, step 2: apply one-to-one (hist == 1) and / or the last-row quirk in place.
cudaError_t post_pair_filter_launch(uint32_t n_rows, const uint64_t *row_off, const uint64_t *hist_off,
                                    uint32_t n_pairs, unsigned flags, int32_t *val, const uint32_t *hist,
                                    cudaStream_t stream);

// Order-preserving compaction of rows with val[r] >= 0.
//   seg_off (n_seg+1): compact-row offset of every segment (view or pair)
// Outputs: out_seg[k], out_i[k] (row inside its segment), out_j[k] = val, out_d[k] = dist
// (out_seg / out_d / dist may be null), seg_out_off[n_seg+1] = output offset of each segment,
// *total (device) = number of survivors.  block_counts needs ceil(n_rows / 2048) + 1 entries.
// *n_launches (optional) = kernels launched, for hulo_launch_count: 2 up to 1024 blocks, 3 above.
cudaError_t compact_launch(const int32_t *val, const int32_t *dist, uint32_t n_rows, const uint64_t *seg_off,
                           uint32_t n_seg, uint32_t *block_counts, uint32_t *out_seg, uint32_t *out_i,
                           uint32_t *out_j, int32_t *out_d, uint64_t *seg_out_off, uint64_t *total,
                           cudaStream_t stream, int *n_launches = nullptr);
constexpr uint32_t kCompactBlockRows = 2048;

}  // namespace hulo
