// comm.cu -- the one exchange step of the path: a row-sharded database returns its local
// top-2 per searcher row and the candidates (16 bytes per row per rank) are merged.  Default data
// plane: stores into peer-mapped buffers (CUDA IPC over NVLink) fused into the chunk-merge kernel,
// with a flag handshake (knn2.cu); NCCL carries the set-up traffic (IPC handles, barriers) and a
// single ncclAllGather is the fallback data plane.  NCCL is bound lazily with dlopen and its few
// types are declared here, so neither building nor single-GPU use needs NCCL installed.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <dlfcn.h>

// The slice of nccl.h this file uses (NCCL 2.x ABI): declared locally so that the library builds on
// a machine without the NCCL headers.
extern "C" {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt8 = 0, ncclChar = 0, ncclFloat64 = 8, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;
}

#include "context.cuh"

namespace hulo {

int run_flat_packed(hulo_gpu *h, const uint4 *A, size_t nA, const uint4 *B, size_t nB, uint32_t row_base);
int run_flat_k1(hulo_gpu *h, const uint4 *A, size_t nA, const uint4 *B, size_t nB, uint32_t row_base, FlatRun *run);

namespace {

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.lib) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        set_error("NCCL: cannot dlopen libnccl.so.2 (%s)", dlerror());
        return false;
    }
    NcclApi a;
    a.lib = lib;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(lib, "ncclAllGather");
    a.AllReduce = (decltype(a.AllReduce))dlsym(lib, "ncclAllReduce");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather || !a.AllReduce || !a.GetErrorString) {
        set_error("NCCL: libnccl.so.2 lacks a required symbol");
        dlclose(lib);
        return false;
    }
    g_nccl = a;
    return true;
}

#define HULO_NCCL(expr)                                                                              \
    do {                                                                                             \
        ncclResult_t r__ = (expr);                                                                   \
        if (r__ != ncclSuccess) {                                                                    \
            hulo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
            return HULO_ERR_NCCL;                                                                    \
        }                                                                                            \
    } while (0)

// ---- exchange buffers over CUDA IPC.  Layout of one rank's buffer:
//   [2][world][cap] int4 records | [2][world] uint32 flags | done counter | status
size_t px_bytes(int world, uint32_t cap) { return (size_t)2 * world * cap * sizeof(int4) + 4096; }

// Collective over the communicator when peers are mapped: every rank closes its mappings of the
// peers' buffers, all ranks meet, and only then is the exported buffer freed (freeing memory that an
// importer still has open is undefined behaviour).
void px_release(hulo_gpu *h) {
    if (h->xstream) cudaStreamSynchronize(h->xstream);
    h->x_pending = false;
    bool had_peers = false;
    for (int g = 0; g < h->world && g < kMaxPeers; ++g) {
        if (g != h->rank && h->px_peer_base[g]) { cudaIpcCloseMemHandle(h->px_peer_base[g]); had_peers = true; }
        h->px_peer_base[g] = nullptr;
    }
    if (had_peers && h->nccl_comm) hulo_comm_barrier(h);
    if (h->px_own) cudaFree(h->px_own);
    h->px_own = nullptr;
    h->px_ready = false;
}

// (Re)build the exchange buffers for `rows` searcher rows: collective over the communicator.
int px_setup(hulo_gpu *h, size_t rows) {
    if (h->world > kMaxPeers) { h->px_disabled = true; return HULO_OK; }
    px_release(h);
    const uint32_t cap = (uint32_t)((rows + 255) & ~(size_t)255);
    const size_t bytes = px_bytes(h->world, cap);
    // a local failure must not leave the peers alone inside the collectives below: carry it to the
    // agreement step instead of returning
    bool ok = true;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (cudaMalloc(&h->px_own, bytes) != cudaSuccess) { cudaGetLastError(); h->px_own = nullptr; ok = false; }
    if (ok && cudaMemsetAsync(h->px_own, 0, bytes, h->stream) != cudaSuccess) { cudaGetLastError(); ok = false; }
    if (ok && cudaIpcGetMemHandle(&mine, h->px_own) != cudaSuccess) { cudaGetLastError(); ok = false; }
    // all-gather the 64-byte handles with the communicator that already exists
    HULO_CUDA(h->comm_scratch.reserve((size_t)(h->world + 1) * sizeof(cudaIpcMemHandle_t)));
    uint8_t *d_all = h->comm_scratch.as<uint8_t>();
    uint8_t *d_mine = d_all + (size_t)h->world * sizeof(cudaIpcMemHandle_t);
    HULO_CUDA(cudaMemcpyAsync(d_mine, &mine, sizeof mine, cudaMemcpyHostToDevice, h->stream));
    HULO_NCCL(g_nccl.AllGather(d_mine, d_all, sizeof mine, ncclChar, (ncclComm_t)h->nccl_comm, h->stream));
    std::vector<cudaIpcMemHandle_t> all((size_t)h->world);
    HULO_CUDA(cudaMemcpyAsync(all.data(), d_all, all.size() * sizeof mine, cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    // did every rank get as far as exporting a buffer?
    {
        double neg = ok ? -1.0 : 0.0;
        int rc = hulo_comm_max_f64(h, &neg);
        if (rc != HULO_OK) return rc;
        if (-neg < 0.5) {
            if (h->px_own) cudaFree(h->px_own);
            h->px_own = nullptr;
            h->px_disabled = true;
            return HULO_OK;
        }
    }
    for (int g = 0; g < h->world; ++g) {
        if (g == h->rank) { h->px_peer_base[g] = h->px_own; continue; }
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, all[(size_t)g], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        h->px_peer_base[g] = p;
    }
    // every rank must agree on whether peer mapping works (min over ranks)
    double flag = ok ? 1.0 : 0.0, neg = -flag;
    int rc = hulo_comm_max_f64(h, &neg);
    if (rc != HULO_OK) return rc;
    if (-neg < 0.5) {
        px_release(h);
        h->px_disabled = true;                  // fall back to the NCCL all-gather exchange
        return HULO_OK;
    }
    PeerExchange &px = h->px;
    px.rank = h->rank; px.world = h->world; px.cap = cap; px.seq = 0;
    for (int g = 0; g < h->world; ++g) {
        uint8_t *base = static_cast<uint8_t *>(h->px_peer_base[g]);
        px.records[g] = reinterpret_cast<int4 *>(base);
        px.flags[g] = reinterpret_cast<uint32_t *>(base + (size_t)2 * h->world * cap * sizeof(int4));
    }
    uint8_t *tail = static_cast<uint8_t *>(h->px_own) + (size_t)2 * h->world * cap * sizeof(int4);
    px.done_counter = reinterpret_cast<unsigned int *>(tail + 2048);
    px.status = reinterpret_cast<unsigned int *>(tail + 2052);
    h->px_seq = 0;
    h->px_ready = true;
    // nobody may start storing into a buffer that a peer is still zeroing
    return hulo_comm_barrier(h);
}

}  // namespace

int gather_query_matches(hulo_gpu *h, const QueryMatchesDev &dm, size_t nv_local, size_t max_nv, size_t slots,
                         uint32_t *all_words) {
    const char *ex = getenv("HULO_EXCHANGE");
    if (h->world <= 1 || h->px_disabled || (ex && strcmp(ex, "nccl") == 0)) return kNoPeerExchange;
    HULO_CUDA(cudaSetDevice(h->device));
    const size_t words = 1 + max_nv + 3 * slots;
    const size_t cap16 = (words * sizeof(uint32_t) + 15) / 16;          // 16-byte records per slot
    if (!h->px_ready || h->px.cap < cap16) {
        int rc = px_setup(h, std::max<size_t>(cap16, 4096));             // collective: words is the same on every rank
        if (rc != HULO_OK) return rc;
        if (!h->px_ready) return kNoPeerExchange;
    }
    HULO_CUDA(join_exchange(h));
    h->px.seq = ++h->px_seq;
    HULO_CUDA(query_matches_store_peers_launch(dm.d_total, dm.d_seg_out, (uint32_t)nv_local, (uint32_t)max_nv,
                                               (uint32_t)slots, dm.o_i, dm.o_j, dm.o_d, h->px, h->stream));
    HULO_CUDA(peers_wait_launch(h->px, h->stream));
    h->launches += 2;
    // this rank's buffer now holds every rank's block for this parity, one per slot
    const size_t pitch = (size_t)h->px.cap * sizeof(int4);
    HULO_CUDA(h->hstage0.reserve(words * sizeof(uint32_t) * (size_t)h->world + sizeof(unsigned int)));
    uint8_t *hp = h->hstage0.as<uint8_t>();
    const uint8_t *src = reinterpret_cast<const uint8_t *>(h->px.records[h->rank]) + (size_t)(h->px.seq & 1u) * h->world * pitch;
    HULO_CUDA(cudaMemcpy2DAsync(hp, words * sizeof(uint32_t), src, pitch, words * sizeof(uint32_t), (size_t)h->world,
                                cudaMemcpyDeviceToHost, h->stream));
    unsigned int *h_status = reinterpret_cast<unsigned int *>(hp + words * sizeof(uint32_t) * (size_t)h->world);
    HULO_CUDA(cudaMemcpyAsync(h_status, h->px.status, sizeof(unsigned int), cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    if (*h_status) {
        HULO_CUDA(cudaMemset(h->px.status, 0, sizeof(unsigned int)));
        set_error("gather_query_matches: a peer did not deliver its matches within 10 s");
        return HULO_ERR_NCCL;
    }
    memcpy(all_words, hp, words * sizeof(uint32_t) * (size_t)h->world);
    return HULO_OK;
}

}  // namespace hulo

using namespace hulo;

extern "C" {

void hulo_comm_destroy_internal(hulo_gpu *h) {
    if (h) px_release(h);
    if (h && h->nccl_comm && g_nccl.lib) {
        g_nccl.CommDestroy((ncclComm_t)h->nccl_comm);
        h->nccl_comm = nullptr;
    }
}

int hulo_comm_unique_id(void *id128) {
    HULO_ARG(id128 != nullptr, "null id buffer");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    if (!load_nccl()) return HULO_ERR_NCCL;
    ncclUniqueId id;
    HULO_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id128, &id, sizeof id);
    return HULO_OK;
}

int hulo_comm_init(hulo_gpu *h, const void *id128, int rank, int world) {
    HULO_ARG(h != nullptr && id128 != nullptr, "null argument");
    HULO_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
    if (!load_nccl()) return HULO_ERR_NCCL;
    HULO_CUDA(cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t comm;
    HULO_NCCL(g_nccl.CommInitRank(&comm, world, id, rank));
    h->nccl_comm = comm;
    h->rank = rank;
    h->world = world;
    return HULO_OK;
}

int hulo_comm_barrier(hulo_gpu *h) {
    double v = 0.0;
    return hulo_comm_max_f64(h, &v);
}

int hulo_comm_max_f64(hulo_gpu *h, double *value) {
    HULO_ARG(h != nullptr && value != nullptr, "null argument");
    if (h->world == 1 && !h->nccl_comm) return hulo_synchronize(h);
    if (!h->nccl_comm) { set_error("hulo_comm_max_f64: communicator not initialised"); return HULO_ERR_NCCL; }
    HULO_CUDA(cudaSetDevice(h->device));
    HULO_CUDA(join_exchange(h));
    HULO_CUDA(h->comm_scratch.reserve(64));
    double *d = h->comm_scratch.as<double>();
    HULO_CUDA(cudaMemcpyAsync(d, value, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    HULO_NCCL(g_nccl.AllReduce(d, d, 1, ncclDouble, ncclMax, (ncclComm_t)h->nccl_comm, h->stream));
    HULO_CUDA(cudaMemcpyAsync(value, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}

int hulo_comm_allgather(hulo_gpu *h, const void *send, size_t bytes, void *recv) {
    HULO_ARG(h != nullptr && (bytes == 0 || (send != nullptr && recv != nullptr)), "null argument");
    if (bytes == 0) return HULO_OK;
    if (h->world == 1 && !h->nccl_comm) { memcpy(recv, send, bytes); return HULO_OK; }
    if (!h->nccl_comm) { set_error("hulo_comm_allgather: communicator not initialised"); return HULO_ERR_NCCL; }
    HULO_CUDA(cudaSetDevice(h->device));
    const size_t slot = (bytes + 15) & ~(size_t)15;
    const size_t world = (size_t)h->world;
    HULO_CUDA(h->comm_scratch.reserve(slot * (world + 1)));
    uint8_t *d_all = h->comm_scratch.as<uint8_t>();
    uint8_t *d_mine = d_all + slot * world;
    // both directions through pinned memory: the caller's buffers are usually pageable, and a
    // pageable copy of this size costs more than the collective
    HULO_CUDA(h->hstage0.reserve(slot * (world + 1)));
    uint8_t *p_all = h->hstage0.as<uint8_t>();
    uint8_t *p_mine = p_all + slot * world;
    memcpy(p_mine, send, bytes);
    HULO_CUDA(cudaMemcpyAsync(d_mine, p_mine, bytes, cudaMemcpyHostToDevice, h->stream));
    HULO_NCCL(g_nccl.AllGather(d_mine, d_all, slot, ncclChar, (ncclComm_t)h->nccl_comm, h->stream));
    HULO_CUDA(cudaMemcpyAsync(p_all, d_all, slot * world, cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    for (size_t r = 0; r < world; ++r) memcpy(static_cast<uint8_t *>(recv) + r * bytes, p_all + r * slot, bytes);
    return HULO_OK;
}

const char *hulo_comm_exchange_kind(const hulo_gpu *h) {
    if (!h || h->world <= 1) return "none";
    const char *ex = getenv("HULO_EXCHANGE");
    if (h->px_disabled || (ex && strcmp(ex, "nccl") == 0)) return "nccl-allgather";
    return "peer-store";
}

int hulo_comm_rank(const hulo_gpu *h) { return h ? h->rank : 0; }
int hulo_comm_world(const hulo_gpu *h) { return h ? std::max(h->world, 1) : 1; }

int hulo_knn2_sharded(hulo_gpu *h, const hulo_db *A, const hulo_db *B_shard, uint64_t row_base, int32_t *idx2,
                      int32_t *dist2) {
    HULO_ARG(h != nullptr && A != nullptr && B_shard != nullptr, "null argument");
    HULO_ARG(row_base + B_shard->n < (uint64_t)INT32_MAX, "global row index exceeds int32");
    HULO_CUDA(cudaSetDevice(h->device));
    const size_t nA = A->n;
    if (h->world > 1 && !h->nccl_comm) { set_error("hulo_knn2_sharded: communicator not initialised"); return HULO_ERR_NCCL; }
    // exchange flavour: stores into peer-mapped buffers fused into the merge epilogue (default),
    // or one ncclAllGather (HULO_EXCHANGE=nccl, or when the peers cannot be mapped)
    const char *ex = getenv("HULO_EXCHANGE");
    bool use_peers = h->world > 1 && !h->px_disabled && !(ex && strcmp(ex, "nccl") == 0);
    if (use_peers && (!h->px_ready || h->px.cap < nA)) {
        int rc = px_setup(h, std::max<size_t>(nA, 4096));
        if (rc != HULO_OK) return rc;
        use_peers = h->px_ready;
    }
    if (use_peers) {
        if (!h->xstream) {
            HULO_CUDA(cudaStreamCreateWithFlags(&h->xstream, cudaStreamNonBlocking));
            for (int p = 0; p < 2; ++p) {
                HULO_CUDA(cudaEventCreateWithFlags(&h->ev_k1[p], cudaEventDisableTiming));
                HULO_CUDA(cudaEventCreateWithFlags(&h->ev_x[p], cudaEventDisableTiming));
            }
        }
        // K1 on the main stream, the exchange on its own: the next call's K1 starts while this
        // call's candidates are still crossing NVLink.  This parity's key buffer and record slots
        // were last used two calls ago; wait for that exchange before K1 overwrites the keys.
        const int p = (int)((h->px_seq + 1) & 1);
        HULO_CUDA(cudaStreamWaitEvent(h->stream, h->ev_x[p], 0));
        std::swap(h->partial, h->partial_alt);
        FlatRun run;
        int rc = run_flat_k1(h, A->rows, nA, B_shard->rows, B_shard->n, (uint32_t)row_base, &run);
        if (rc != HULO_OK) return rc;
        HULO_CUDA(cudaEventRecord(h->ev_k1[p], h->stream));
        HULO_CUDA(cudaStreamWaitEvent(h->xstream, h->ev_k1[p], 0));
        h->px.seq = ++h->px_seq;
        HULO_CUDA(knn2_merge_store_peers_launch(h->partial.as<uint2>(), (uint32_t)nA, run.n_chunks, run.slot_stride,
                                                run.rows_per_chunk, (uint32_t)row_base, h->px, h->xstream));
        HULO_CUDA(knn2_merge_from_peers_launch(h->px, (uint32_t)nA, h->knn_idx.as<int32_t>(),
                                               h->knn_dist.as<int32_t>(), h->xstream));
        HULO_CUDA(cudaEventRecord(h->ev_x[p], h->xstream));
        h->x_pending = true;
        h->x_last = p;
        h->launches += 2;
        if (idx2 || dist2) {
            int rc2 = hulo_knn2_fetch(h, nA, idx2, dist2);
            if (rc2 != HULO_OK) return rc2;
            unsigned int st = 0;
            HULO_CUDA(cudaMemcpy(&st, h->px.status, sizeof st, cudaMemcpyDeviceToHost));
            if (st) {
                HULO_CUDA(cudaMemset(h->px.status, 0, sizeof st));      // reported once, not on every later fetch
                set_error("hulo_knn2_sharded: a peer did not deliver its candidates within 10 s");
                return HULO_ERR_NCCL;
            }
        }
        return HULO_OK;
    }
    HULO_CUDA(join_exchange(h));
    int rc = run_flat_packed(h, A->rows, nA, B_shard->rows, B_shard->n, (uint32_t)row_base);
    if (rc != HULO_OK) return rc;
    if (h->world > 1 && nA > 0) {
        HULO_CUDA(h->gathered.reserve((size_t)h->world * nA * sizeof(int4)));
        // one all-gather of nA x 16 bytes per rank
        HULO_NCCL(g_nccl.AllGather(h->packed.ptr, h->gathered.ptr, nA * sizeof(int4), ncclChar,
                                   (ncclComm_t)h->nccl_comm, h->stream));
        HULO_CUDA(knn2_merge_ranks_launch(h->gathered.as<int4>(), (uint32_t)nA, h->world,
                                          h->knn_idx.as<int32_t>(), h->knn_dist.as<int32_t>(), h->stream));
        h->launches++;
    }
    if (idx2 || dist2) return hulo_knn2_fetch(h, nA, idx2, dist2);
    return HULO_OK;
}

int hulo_knn2_sharded_submit(hulo_gpu *h, const hulo_db *A, const hulo_db *B_shard, uint64_t row_base) {
    HULO_ARG(h != nullptr && A != nullptr && B_shard != nullptr, "null argument");
    HULO_ARG(h->pipe_submitted - h->pipe_collected < 2, "two searches are outstanding: collect one first");
    HULO_CUDA(cudaSetDevice(h->device));
    if (!h->cstream) {
        HULO_CUDA(cudaStreamCreateWithFlags(&h->cstream, cudaStreamNonBlocking));
        for (int q = 0; q < 2; ++q) HULO_CUDA(cudaEventCreateWithFlags(&h->ev_res[q], cudaEventDisableTiming));
    }
    const int q = (int)(h->pipe_submitted & 1);
    // this search's results go to the set that was collected longest ago
    std::swap(h->knn_idx, h->knn_idx_alt);
    std::swap(h->knn_dist, h->knn_dist_alt);
    h->in_submit = true;
    int rc = hulo_knn2_sharded(h, A, B_shard, row_base, nullptr, nullptr);
    h->in_submit = false;
    if (rc != HULO_OK) return rc;
    // the results are complete where the search ends: on the exchange stream when it ran there
    HULO_CUDA(cudaEventRecord(h->ev_res[q], h->x_pending ? h->xstream : h->stream));
    h->pipe_slot[q].idx = h->knn_idx.ptr;
    h->pipe_slot[q].dist = h->knn_dist.ptr;
    h->pipe_slot[q].nA = A->n;
    h->pipe_slot[q].busy = true;
    ++h->pipe_submitted;
    return HULO_OK;
}

int hulo_knn2_sharded_collect(hulo_gpu *h, int32_t *idx2, int32_t *dist2, size_t *n_rows) {
    HULO_ARG(h != nullptr, "null context");
    HULO_ARG(h->pipe_collected < h->pipe_submitted, "nothing was submitted");
    HULO_CUDA(cudaSetDevice(h->device));
    const int q = (int)(h->pipe_collected & 1);
    hulo_gpu::PipeSlot &sl = h->pipe_slot[q];
    if (n_rows) *n_rows = sl.nA;
    HULO_CUDA(cudaStreamWaitEvent(h->cstream, h->ev_res[q], 0));
    if (sl.nA > 0) {
        if (idx2) HULO_CUDA(cudaMemcpyAsync(idx2, sl.idx, sl.nA * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->cstream));
        if (dist2) HULO_CUDA(cudaMemcpyAsync(dist2, sl.dist, sl.nA * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->cstream));
    }
    HULO_CUDA(cudaStreamSynchronize(h->cstream));
    sl.busy = false;
    ++h->pipe_collected;
    if (h->world > 1 && h->px_ready) {
        unsigned int st = 0;
        HULO_CUDA(cudaMemcpyAsync(&st, h->px.status, sizeof st, cudaMemcpyDeviceToHost, h->cstream));
        HULO_CUDA(cudaStreamSynchronize(h->cstream));
        if (st) {
            HULO_CUDA(cudaMemsetAsync(h->px.status, 0, sizeof st, h->cstream));
            set_error("hulo_knn2_sharded_collect: a peer did not deliver its candidates within 10 s");
            return HULO_ERR_NCCL;
        }
    }
    return HULO_OK;
}

int hulo_merge_top2(hulo_gpu *h, const int32_t *cand, size_t nA, int world, int32_t *idx2, int32_t *dist2) {
    HULO_ARG(h != nullptr && world >= 1, "bad argument");
    HULO_ARG(nA == 0 || (cand != nullptr && idx2 != nullptr && dist2 != nullptr), "null argument");
    HULO_ARG(h->pipe_submitted == h->pipe_collected,
             "searches issued with hulo_knn2_sharded_submit are outstanding: collect them first");
    HULO_CUDA(cudaSetDevice(h->device));
    HULO_CUDA(join_exchange(h));
    HULO_CUDA(h->knn_idx.reserve(std::max<size_t>(nA, 1) * 2 * sizeof(int32_t)));
    HULO_CUDA(h->knn_dist.reserve(std::max<size_t>(nA, 1) * 2 * sizeof(int32_t)));
    h->last_nA = nA;
    if (nA == 0) return HULO_OK;
    HULO_CUDA(h->gathered.reserve((size_t)world * nA * sizeof(int4)));
    HULO_CUDA(cudaMemcpyAsync(h->gathered.ptr, cand, (size_t)world * nA * sizeof(int4), cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(knn2_merge_ranks_launch(h->gathered.as<int4>(), (uint32_t)nA, world, h->knn_idx.as<int32_t>(),
                                      h->knn_dist.as<int32_t>(), h->stream));
    h->launches++;
    return hulo_knn2_fetch(h, nA, idx2, dist2);
}

}  // extern "C"
