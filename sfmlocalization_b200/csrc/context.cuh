// context.cuh -- internal state behind the opaque handles of include/hulo_gpu.h.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <string>
#include <vector>

#include "../../include/hulo_gpu.h"
#include "knn2.cuh"
#include "knn2_tc.cuh"

namespace hulo {

void set_error(const char *fmt, ...);

// what a flat K1 launch left in the partial-key buffer
struct FlatRun {
    uint32_t n_chunks, rows_per_chunk;
    uint64_t slot_stride;
};

#define HULO_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t e__ = (expr);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            hulo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__)); \
            return HULO_ERR_CUDA;                                                             \
        }                                                                                     \
    } while (0)

#define HULO_ARG(cond, msg)                                  \
    do {                                                     \
        if (!(cond)) {                                       \
            hulo::set_error("%s: %s", __func__, msg);        \
            return HULO_ERR_ARG;                             \
        }                                                    \
    } while (0)

// A device buffer that only ever grows (scratch reused across calls).
struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMalloc(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

// Same for pinned host staging.
struct HostBuf {
    void *ptr = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&ptr, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (ptr) cudaFreeHost(ptr);
        ptr = nullptr;
        cap = 0;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(ptr); }
};

}  // namespace hulo

struct hulo_db {
    hulo_gpu *owner = nullptr;
    uint4 *rows = nullptr;             // n x 64 bytes
    size_t n = 0;
    size_t cap_rows = 0;
    std::vector<uint64_t> seg;         // n_seg + 1 offsets
};

struct hulo_gpu {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;
    uint64_t launches = 0;
    hulo::KnnConfig knn_cfg{512, 4, 89, 1};  // best of the sweeps on C3 (profiles/r1_k1_variant_sweep.txt, r1_k1_sweep_folded.txt)
    bool knn_cfg_forced = false;

    // K1 arithmetic of the flat searches: HULO_KNN_INT = integer pipes (knn2.cu), HULO_KNN_TC = int8
    // contraction on the tensor cores (knn2_tc.cu), HULO_KNN_AUTO = K1t for large searches
    int knn_engine = HULO_KNN_AUTO;
    int tc_bits = 4;           // operand width of the tensor-core engine: 4 (K1t4, kind::mxf4) or 8 (K1t, kind::i8)
    // tile images of K1t: one per registered table (built on first use, dropped when the table
    // changes), and two scratch images for staged rows
    // kind: kTcFlat8 = int8 image of a flat table, kTcSeg8 = the segmented int8 form (every segment
    // of the table starts on an even tile, tile0[s] = its first tile) used by the item-mode searches,
    // kTcFlat4 / kTcSeg4 = the 4-bit images of K1t4 (segments start on an 8-row group, tile0 = first group).
    enum { kTcFlat8 = 0, kTcSeg8 = 1, kTcFlat4 = 2, kTcSeg4 = 3 };
    struct TcImage { const void *rows; size_t n; bool valid; int kind; std::vector<uint32_t> tile0; hulo::DevBuf img; };
    std::vector<TcImage> tc_images;
    hulo::DevBuf tc_scratchA, tc_scratchB, tc_tiles;
    // item list of the last hulo_match_to_query on the device: a server asks the same (map, view list,
    // query size) again and again, and the list (32 bytes per searcher tile) need not go up each time
    struct TcItemCache { const void *rows = nullptr; size_t n_views = 0, nq = 0, n_items = 0; bool all_views = false; std::vector<uint32_t> views; int bits = 0; bool valid = false; } tc_qitems_key;
    hulo::DevBuf tc_qitems;

    hulo::DevBuf partial;      // K1 per-item keys
    hulo::DevBuf counter;      // K1 dynamic item counter
    hulo::DevBuf items;        // K1 item list (pair / view modes)
    hulo::DevBuf knn_idx, knn_dist;   // final nA x 2 results
    hulo::DevBuf packed;       // nA x int4 (sharded exchange)
    hulo::DevBuf gathered;     // world x nA x int4
    hulo::DevBuf comm_scratch; // set-up collectives (IPC handles, barriers): never a buffer a search has results in
    hulo::DevBuf stageA, stageB;      // uploaded rows of the *_host entry points
    hulo::DevBuf scratch0, scratch1, scratch2, scratch3;   // post-processing / K2
    hulo::HostBuf hstage0, hstage1;
    size_t last_nA = 0;
    size_t score_smem_configured = 0;   // dynamic smem opt-in already set for K2 on this device
    hulo::DevBuf wave_counter; // K2: blocks-done counter of the fused resection wave (rests at 0)
    hulo::HostBuf wave_rec;    // K2: the wave's record {nfa, index, model[12], -, sequence}, written by the kernel
    uint64_t wave_seq = 0;
    hulo::DevBuf lfact;        // K3: log10(n!) table
    size_t lfact_n = 0;
    size_t geo_smem_configured = 0;

    // NCCL (loaded lazily)
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;
    // fused exchange over peer memory (CUDA IPC): set up on the first sharded search
    hulo::PeerExchange px{};
    void *px_own = nullptr;            // this rank's exchange buffer (records + flags)
    void *px_peer_base[hulo::kMaxPeers] = {nullptr};
    bool px_ready = false, px_disabled = false;
    uint32_t px_seq = 0;
    // The exchange runs on its own stream so that the next search's K1 overlaps it: ev_k1[p] = K1 of
    // the step with parity p has left its keys, ev_x[p] = that step's exchange is complete.  Keys are
    // double-buffered (partial / partial_alt) like the record slots.  Every other entry point that
    // touches the result buffers joins the exchange stream first (hulo::join_exchange).
    cudaStream_t xstream = nullptr;
    cudaEvent_t ev_k1[2] = {nullptr, nullptr}, ev_x[2] = {nullptr, nullptr};
    hulo::DevBuf partial_alt;
    bool x_pending = false;
    int x_last = 0;
    // hulo_knn2_sharded_submit / _collect: two result sets, one event each (results complete),
    // a stream for the result copies so that they never queue behind the next search's K1
    hulo::DevBuf knn_idx_alt, knn_dist_alt;
    struct PipeSlot { const void *idx = nullptr, *dist = nullptr; size_t nA = 0; bool busy = false; } pipe_slot[2];
    cudaEvent_t ev_res[2] = {nullptr, nullptr};
    cudaStream_t cstream = nullptr;
    uint64_t pipe_submitted = 0, pipe_collected = 0;
    bool in_submit = false;            // the search being issued is a submit (it may run while others are outstanding)
};

namespace hulo {
// What hulo_match_to_query leaves on the device before its results are fetched: the view-sharded
// query hands these to the peer exchange instead of copying them out.
struct QueryMatchesDev {
    const uint64_t *d_total = nullptr;     // total survivors; nullptr: nothing was searched (no rows / no views)
    const uint64_t *d_seg_out = nullptr;   // n_views + 1 output offsets (view v: [v], [v + 1])
    const uint32_t *o_view = nullptr, *o_i = nullptr, *o_j = nullptr;
    const int32_t *o_d = nullptr;
    size_t n_views = 0;
    uint64_t n_rows = 0;
};
int match_to_query_device(hulo_gpu *h, const hulo_db *map, const uint32_t *views, size_t n_views, const uint8_t *query,
                          size_t nq, size_t q_stride, float ratio, QueryMatchesDev *out);
// The blocks {n_matches, counts[max_nv], (i, j, d0)[slots]} of all ranks, in rank order, into
// all_words (world x (1 + max_nv + 3 slots) words, host).  Returns kNoPeerExchange when the
// peers are not mapped (the caller falls back to the host-staged all-gather).
int gather_query_matches(hulo_gpu *h, const QueryMatchesDev &dm, size_t nv_local, size_t max_nv, size_t slots,
                         uint32_t *all_words);
constexpr int kNoPeerExchange = 1000;
// Make the main stream wait for every exchange still in flight on the exchange stream.
inline cudaError_t join_exchange(hulo_gpu *h) {
    if (!h->x_pending) return cudaSuccess;
    h->x_pending = false;
    return cudaStreamWaitEvent(h->stream, h->ev_x[h->x_last], 0);
}
}  // namespace hulo
