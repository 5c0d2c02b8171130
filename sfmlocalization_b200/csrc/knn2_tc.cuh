// knn2_tc.cuh -- K1t: exact Hamming 2-NN as an int8 contraction on the 5th-generation tensor
// cores (tcgen05.mma.kind::i8, accumulators in TMEM), sm_100a.
//
// Same contract as K1 (knn2.cuh): replaces the cv::flann::Index::knnSearch(k=2) calls of the
// reference (VisionLocalizeCommon/src/MatchUtils.cpp:105-108, 191-194, 339-340) with exact search and
// leaves per-(chunk, searcher row) packed keys (distance << 22 | row-in-chunk) in the partial
// buffer that knn2_merge_* folds.  Only the distance arithmetic differs:
//
//   every bit b of a 512-bit row becomes the int8 value 2b - 1, so for two rows a, b
//       dot(a, b) = #equal bits - #different bits = 512 - 2 * hamming(a, b)
//   exactly, in int32 accumulation.  Padding bits (486..511) are zero on both sides and count as
//   equal, so the identity also holds for 61-byte AKAZE rows.
//
// Rows are expanded once into a TILE IMAGE that is byte for byte what the tensor core reads from
// shared memory (K-major, no swizzle: 8 rows x 16 bytes core matrices), so the loads are plain
// contiguous bulk copies (cp.async.bulk, SASS UBLKCP) with no tensor map:
//
//   offset(row r, byte k of 512) = (r / 128) * 65536        128-row tile
//                                + (k / 128) * 16384        K chunk of 128 bytes (one ring stage)
//                                + ((r % 128) / 8) * 1024   8-row group      (descriptor SBO)
//                                + ((k % 128) / 16) * 128   core matrix in K (descriptor LBO)
//                                + (r % 8) * 16 + k % 16
//
// The same image serves as the A operand (searcher tile, M = 128, resident in shared memory for a
// whole work item) and as the B operand (two consecutive tiles = N = 256 database rows per MMA).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "knn2.cuh"

namespace hulo {

constexpr uint32_t kTcTileRows = 128;                 // rows per image tile
constexpr uint32_t kTcRowBytes = 512;                 // expanded row: one int8 per bit
constexpr size_t kTcTileBytes = (size_t)kTcTileRows * kTcRowBytes;   // 65536
constexpr uint32_t kTcTileN = 256;                    // database rows per accumulator tile

// Bytes of the tile image of an n-row table (an even number of tiles: B is read two tiles at a time).
inline size_t knn2_tc_image_bytes(size_t n) {
    const size_t tiles = ((n + kTcTileN - 1) / kTcTileN) * 2;
    return (tiles < 2 ? 2 : tiles) * kTcTileBytes;
}

// FOLDED 64-byte rows (the device layout of K1) -> tile image.  Rows past n inside the last
// tiles are written as zeros.
cudaError_t knn2_tc_expand_launch(const uint4 *folded_rows, size_t n, uint8_t *image, cudaStream_t stream);

// The same for a table cut into segments (views, images, query images), every segment starting on
// an even tile so that it can serve as searcher (128-row tiles) and as database (256-row tiles):
// image tile t holds rows tile_src[t] .. tile_src[t] + tile_rows[t] - 1 of the table, zeros after
// (tile_rows[t] = 0: a padding tile).  Both tables are device pointers of n_tiles entries.
cudaError_t knn2_tc_expand_tiles_launch(const uint4 *folded_rows, const uint32_t *tile_src, const uint32_t *tile_rows,
                                        size_t n_tiles, uint8_t *image, cudaStream_t stream);

// One unit of work of the item mode (32 bytes): searcher tile `a_tile` of image A against
// `b_rows` database rows starting at tile `b_tile0` (even) of image B.
struct alignas(16) TcItem {
    uint32_t a_tile;
    uint32_t a_rows;     // rows of the tile whose keys are written (<= 128)
    uint32_t b_tile0;
    uint32_t b_rows;     // <= kMaxChunkRows
    uint64_t out_slot0;  // partial[out_slot0 + r] receives the keys of searcher row r of the tile
    uint64_t pad;
};

struct TcParams {
    const uint8_t *imgA;      // searcher tile image
    const uint8_t *imgB;      // database tile image
    const TcItem *items = nullptr;   // item mode when set (device pointer), else the flat fields below
    uint32_t n_items = 0;
    uint32_t nA, nB;
    uint32_t n_mtiles;        // ceil(nA / 128)
    uint32_t n_chunks;
    uint32_t rows_per_chunk;  // multiple of 256, <= kMaxChunkRows
    uint64_t slot_stride;     // partial slot of (chunk, row) = chunk * slot_stride + row
    uint2 *partial;           // packed keys (best, second), same format as K1
    // bring-up hooks (tools/k1t_probe.cu): when dbg_dots is set, the raw accumulators of the first
    // tile of work item 0 are written there (128 x 256 int32)
    int32_t *dbg_dots = nullptr;
    uint32_t lbo = 128, sbo = 1024;   // shared-memory descriptor strides of the tile image
    int cluster = 1;                  // CTAs per cluster sharing every database stage (1, 2 or 4); 1 measured fastest
};

// Persistent launch, one CTA per SM in clusters of p.cluster CTAs; work item w -> group of p.cluster
// consecutive searcher tiles (w % n_groups), chunk (w / n_groups), items dealt round-robin over the
// clusters, so the CTAs running at any moment share a handful of database chunks through L2 and the
// CTAs of a cluster share every stage through one multicast copy.
cudaError_t knn2_tc_launch(const TcParams &p, int grid, cudaStream_t stream);

// Chunk size for a K1t run (multiple of 256 rows).
void knn2_tc_plan(size_t nA, size_t nB, int n_ctas, uint32_t *n_mtiles, uint32_t *n_chunks, uint32_t *rows_per_chunk);


// ---- K1t4 (knn2_tc4.cu): the flat search with 4-bit operands (kind::mxf4, block scales 1.0, fp32
// accumulation -- exact for these integers).  Its own image format (256 bytes per row); TcParams as
// for the flat mode of K1t with rows_per_chunk a multiple of 224.
size_t knn2_tc4_image_bytes(size_t n);
cudaError_t knn2_tc4_expand_launch(const uint4 *folded_rows, size_t n, uint8_t *image, cudaStream_t stream);
void knn2_tc4_plan(size_t nA, size_t nB, int n_ctas, uint32_t *n_mtiles, uint32_t *n_chunks, uint32_t *rows_per_chunk);
cudaError_t knn2_tc4_launch(const TcParams &p, int grid, cudaStream_t stream);
// Item mode of K1t4: TcItem::a_tile / b_tile0 are 8-row GROUP indices into the images (any group can
// start a searcher tile or a database range, so the segments of a table only need 8-row alignment).
// The segmented image: group g holds table rows group_src[g] .. + group_rows[g] - 1 (device tables).
size_t knn2_tc4_groups_image_bytes(size_t n_groups);
cudaError_t knn2_tc4_expand_groups_launch(const uint4 *folded_rows, const uint32_t *group_src, const uint8_t *group_rows,
                                          size_t n_groups, uint8_t *image, cudaStream_t stream);

}  // namespace hulo
