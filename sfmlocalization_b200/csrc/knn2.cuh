// knn2.cuh -- K1: exact Hamming 2-NN over 64-byte rows, sm_100a.
//
// Replaces the cv::flann::Index::knnSearch(k=2) calls of the reference
// (VisionLocalizeCommon/src/MatchUtils.cpp:105-108, 191-194, 339-340) with exact search.
//
// Work decomposition: a work item is (tile of searcher rows A) x (chunk of database rows B).
// Each thread keeps QPT searcher rows in registers (16 x u32 each); database rows are
// streamed global -> shared by 1-D TMA bulk copies (cp.async.bulk + mbarrier, SASS UBLKCP)
// through a small ring and read back as warp-wide broadcast LDS.128.  Per (searcher, row):
// 16 XOR (LOP3), a carry-save adder tree in LOP3 that folds the 16 words to fewer words of
// weight 1/2/4/8, POPC of those, and a running best-2 kept as packed keys
// (distance << 22 | row-in-chunk) so "lowest index wins a tie" is one unsigned min.
// Results per item go to a partial buffer; knn2_merge folds the chunks of a row with the
// same (distance, global index) order, which is associative, so the answer is bit-identical
// to a sequential scan.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hulo {

constexpr int kKeyIdxBits = 22;                       // rows per chunk <= 4 Mi
constexpr uint32_t kKeyIdxMask = (1u << kKeyIdxBits) - 1u;
constexpr uint32_t kKeyNone = 0xFFFFFFFFu;            // > any real key ((512 << 22) | idx)
constexpr uint32_t kMaxChunkRows = 1u << kKeyIdxBits;

// One unit of work for the K1 kernel (32 bytes).
struct alignas(16) KnnItem {
    uint32_t a_row0;    // first searcher row (row index into A)
    uint32_t a_rows;    // searcher rows in this tile (<= THREADS * QPT)
    uint32_t b_row0;    // first database row of the chunk
    uint32_t b_rows;    // rows in the chunk (<= kMaxChunkRows)
    uint64_t out_slot0; // partial[out_slot0 + r] receives the keys of searcher row a_row0 + r
    uint64_t pad;
};

struct KnnParams {
    const uint4 *A;        // searcher rows, 4 x uint4 each
    const uint4 *B;        // database rows
    const KnnItem *items;  // nullptr: flat mode, items derived from the fields below
    uint32_t n_items;
    // flat mode: item w -> tile (w % n_tiles), chunk (w / n_tiles)
    uint32_t nA, nB, n_tiles, rows_per_chunk;
    uint64_t slot_stride;  // partial slot of (chunk, row) = chunk * slot_stride + row
    uint32_t key_unit;     // 1 << kKeyIdxBits (a kernel argument so IMAD multipliers stay in registers)
    uint2 *partial;        // packed keys (best, second)
    unsigned int *counter; // dynamic item counter, zeroed before launch
};

struct KnnConfig {
    int threads;
    int qpt;
    int csa;   // carry-save depth: number of CSAs applied before POPC (0, 5, 7, 8, 9, 11); 89 / 889: 8 and 9 mixed per searcher row
    int opt;   // bit 0: adds on the FMA pipe (IMAD); bit 1: two rows per best-2 update (VIMNMX3)
};

// Launch K1 (variant chosen by cfg) on `stream`; grid_ctas persistent CTAs.
cudaError_t knn2_launch(const KnnParams &p, const KnnConfig &cfg, int grid_ctas, cudaStream_t stream);
// Rows of A one item covers for a config.
inline uint32_t knn2_tile_rows(const KnnConfig &cfg) { return (uint32_t)(cfg.threads * cfg.qpt); }
// Registers / occupancy report used by tests and the bench (cudaFuncGetAttributes).
cudaError_t knn2_kernel_info(const KnnConfig &cfg, int *regs, int *max_ctas_per_sm, size_t *smem);

// Device layout of a descriptor row: FOLDED -- words 3i+2 (i = 0..4) hold w[3i] ^ w[3i+1] ^ w[3i+2]
// and word 15 holds w9 ^ ... ^ w15 (see hamming_key in knn2.cu).  Every table and staging buffer
// K1 reads must be folded after its upload; a download copies and unfolds on the host.
cudaError_t knn2_fold_rows_launch(uint4 *rows, size_t n, cudaStream_t stream);
void knn2_unfold_rows_host(uint8_t *rows64, size_t n);

// Merge the per-chunk keys of every searcher row into final (idx, dist) pairs.
//   slot(row, c) = c * slot_stride + row ; global index = row_base + c * rows_per_chunk + (key & mask)
// out_idx2/out_dist2: nA x 2 int32.  out_packed (optional): nA x int4 {d0, i0, d1, i1} with
// 32-bit global indices, the record exchanged by the row-sharded all-gather.
cudaError_t knn2_merge_launch(const uint2 *partial, uint32_t nA, uint32_t n_chunks, uint64_t slot_stride,
                              uint32_t rows_per_chunk, uint32_t row_base, int32_t *out_idx2,
                              int32_t *out_dist2, int4 *out_packed, cudaStream_t stream);

// Merge `world` all-gathered candidate lists (world x nA int4 {d0,i0,d1,i1}) per row.
cudaError_t knn2_merge_ranks_launch(const int4 *gathered, uint32_t nA, int world, int32_t *out_idx2,
                                    int32_t *out_dist2, cudaStream_t stream);

// ---- fused exchange over peer memory (row-sharded database, one process per GPU) ----
// Every rank owns an exchange buffer that all peers have mapped (CUDA IPC over NVLink):
//   records: [2 parities][world][cap] int4 {d0,i0,d1,i1};  flags: [2][world] uint32 sequence numbers.
constexpr int kMaxPeers = 16;
struct PeerExchange {
    int4 *records[kMaxPeers];       // records[g]: rank g's buffer as mapped in this process
    uint32_t *flags[kMaxPeers];     // flags[g]:   rank g's flag words
    int rank, world;
    uint32_t cap;                   // rows per (parity, rank) slot
    uint32_t seq;                   // sequence number of this exchange (>= 1)
    unsigned int *done_counter;     // local: blocks of the store kernel that have finished
    unsigned int *status;           // local: set to 1 when a wait timed out
};
// Step 1 (the K1 chunk-merge epilogue): fold the chunk keys of every searcher row and STORE the
// packed top-2 record straight into slot [seq&1][rank] of every rank's buffer (peer stores over
// NVLink), then publish flags[g][seq&1][rank] = seq on every rank.
cudaError_t knn2_merge_store_peers_launch(const uint2 *partial, uint32_t nA, uint32_t n_chunks,
                                          uint64_t slot_stride, uint32_t rows_per_chunk, uint32_t row_base,
                                          const PeerExchange &px, cudaStream_t stream);
// Step 2: wait until every rank's record list for `seq` has landed in the local buffer, then
// merge them per row on (distance, global index).
cudaError_t knn2_merge_from_peers_launch(const PeerExchange &px, uint32_t nA, int32_t *out_idx2,
                                         int32_t *out_dist2, cudaStream_t stream);
// The view-sharded query on the same buffers (16 bytes of a slot = 4 words): store this rank's block
// {n_matches, counts[max_nv], (i, j, d0)[<= slots]} into every rank's buffer, then wait for all blocks.
cudaError_t query_matches_store_peers_launch(const uint64_t *d_total, const uint64_t *d_seg_out, uint32_t nv_local,
                                             uint32_t max_nv, uint32_t slots, const uint32_t *o_i, const uint32_t *o_j,
                                             const int32_t *o_d, const PeerExchange &px, cudaStream_t stream);
cudaError_t peers_wait_launch(const PeerExchange &px, cudaStream_t stream);

}  // namespace hulo
