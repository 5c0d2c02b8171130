// comm.cu -- the one exchange step of the path: a row-sharded database returns its local
// top-2 per searcher row and the candidates (16 bytes per row per rank) are merged after a
// single ncclAllGather over NVLink.  NCCL is bound lazily with dlopen so that single-GPU use
// and the CPU-side symbol check do not need it.
#include <algorithm>
#include <dlfcn.h>
#include <nccl.h>

#include "context.cuh"

namespace hulo {

int run_flat_packed(hulo_gpu *h, const uint4 *A, size_t nA, const uint4 *B, size_t nB, uint32_t row_base);

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

}  // namespace
}  // namespace hulo

using namespace hulo;

extern "C" {

void hulo_comm_destroy_internal(hulo_gpu *h) {
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
    HULO_CUDA(h->scratch2.reserve(64));
    double *d = h->scratch2.as<double>();
    HULO_CUDA(cudaMemcpyAsync(d, value, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    HULO_NCCL(g_nccl.AllReduce(d, d, 1, ncclDouble, ncclMax, (ncclComm_t)h->nccl_comm, h->stream));
    HULO_CUDA(cudaMemcpyAsync(value, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}

int hulo_knn2_sharded(hulo_gpu *h, const hulo_db *A, const hulo_db *B_shard, uint64_t row_base, int32_t *idx2,
                      int32_t *dist2) {
    HULO_ARG(h != nullptr && A != nullptr && B_shard != nullptr, "null argument");
    HULO_ARG(row_base + B_shard->n < (uint64_t)INT32_MAX, "global row index exceeds int32");
    HULO_CUDA(cudaSetDevice(h->device));
    const size_t nA = A->n;
    int rc = run_flat_packed(h, A->rows, nA, B_shard->rows, B_shard->n, (uint32_t)row_base);
    if (rc != HULO_OK) return rc;
    if (h->world > 1) {
        if (!h->nccl_comm) { set_error("hulo_knn2_sharded: communicator not initialised"); return HULO_ERR_NCCL; }
        if (nA > 0) {
            HULO_CUDA(h->gathered.reserve((size_t)h->world * nA * sizeof(int4)));
            // one all-gather of nA x 16 bytes per rank
            HULO_NCCL(g_nccl.AllGather(h->packed.ptr, h->gathered.ptr, nA * sizeof(int4), ncclChar,
                                       (ncclComm_t)h->nccl_comm, h->stream));
            HULO_CUDA(knn2_merge_ranks_launch(h->gathered.as<int4>(), (uint32_t)nA, h->world,
                                              h->knn_idx.as<int32_t>(), h->knn_dist.as<int32_t>(), h->stream));
            h->launches++;
        }
    }
    if (idx2 || dist2) return hulo_knn2_fetch(h, nA, idx2, dist2);
    return HULO_OK;
}

int hulo_merge_top2(hulo_gpu *h, const int32_t *cand, size_t nA, int world, int32_t *idx2, int32_t *dist2) {
    HULO_ARG(h != nullptr && world >= 1, "bad argument");
    HULO_ARG(nA == 0 || (cand != nullptr && idx2 != nullptr && dist2 != nullptr), "null argument");
    HULO_CUDA(cudaSetDevice(h->device));
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
