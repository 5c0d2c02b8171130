// api.cu -- C-ABI of libhulo_gpu.so (include/hulo_gpu.h): context, descriptor tables,
// K1 entry points (flat 2-NN, query localisation, image-pair matching).
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <new>

#include "context.cuh"
#include "match_post.cuh"

namespace hulo {

static thread_local std::string g_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
}

// ---------------------------------------------------------------- K1 planning
struct FlatPlan {
    KnnConfig cfg;
    uint32_t n_tiles = 0, n_chunks = 0, rows_per_chunk = 0;
    uint64_t slot_stride = 0;
    int grid = 0;
};

static int env_int(const char *name, int dflt) {
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Searcher-tile shape: the widest tile (fewest B re-reads) that the searcher count still fills.
// "threads,qpt,csa,opt" from the environment (tuning sweeps), else the default
static KnnConfig env_config(const char *name, KnnConfig dflt) {
    const char *s = getenv(name);
    KnnConfig c = dflt;
    if (s && sscanf(s, "%d,%d,%d,%d", &c.threads, &c.qpt, &c.csa, &c.opt) == 4 &&
        knn2_kernel_info(c, nullptr, nullptr, nullptr) == cudaSuccess)
        return c;
    cudaGetLastError();
    return dflt;
}
static KnnConfig small_config() { static const KnnConfig c = env_config("HULO_KNN_SMALL", KnnConfig{128, 4, 8, 1}); return c; }
static KnnConfig mid_config() { static const KnnConfig c = env_config("HULO_KNN_MID", KnnConfig{256, 4, 89, 1}); return c; }

static KnnConfig choose_config(const hulo_gpu *h, size_t nA) {
    if (h->knn_cfg_forced) return h->knn_cfg;
    KnnConfig c = h->knn_cfg;
    if (nA <= 512) c = small_config();
    else if (nA <= 1024) c = mid_config();
    return c;
}

// Split nB database rows into chunks so that tiles x chunks fills the persistent grid evenly.
// Items are handed out dynamically, so the makespan is (items per CTA, rounded up) x (cost of one
// item); the cost of an item is its chunk rows plus a fixed per-item overhead (loading the
// searcher tile into registers, writing its keys), expressed in equivalent rows.
static void plan_chunks(uint32_t n_tiles, size_t nB, int n_ctas, uint32_t *n_chunks, uint32_t *rows_per_chunk) {
    if (nB == 0) { *n_chunks = 0; *rows_per_chunk = 1; return; }
    if (n_tiles == 0) n_tiles = 1;
    const uint64_t overhead_rows = 24;
    const uint32_t min_rows = 128;
    const uint32_t c_min = (uint32_t)((nB + kMaxChunkRows - 1) / kMaxChunkRows);
    uint32_t c_max = (uint32_t)std::max<size_t>(c_min, nB / min_rows);
    // no point in more than ~16 items per CTA
    c_max = std::min<uint32_t>(c_max, std::max<uint32_t>(c_min, (uint32_t)((16ull * n_ctas + n_tiles - 1) / n_tiles)));
    uint32_t best_c = c_min;
    uint64_t best_cost = ~0ull;
    for (uint32_t c = c_min; c <= c_max; ++c) {
        const uint64_t rpc = (nB + c - 1) / c;
        const uint64_t chunks = (nB + rpc - 1) / rpc;
        const uint64_t items = (uint64_t)n_tiles * chunks;
        const uint64_t per_cta = (items + n_ctas - 1) / n_ctas;
        const uint64_t cost = per_cta * (rpc + overhead_rows);
        if (cost < best_cost) { best_cost = cost; best_c = c; }
    }
    const uint32_t rpc = (uint32_t)((nB + best_c - 1) / best_c);
    *rows_per_chunk = rpc;
    *n_chunks = (uint32_t)((nB + rpc - 1) / rpc);
}

static FlatPlan plan_flat(const hulo_gpu *h, size_t nA, size_t nB) {
    FlatPlan p;
    p.cfg = choose_config(h, nA);
    const uint32_t tile = knn2_tile_rows(p.cfg);
    p.n_tiles = (uint32_t)((nA + tile - 1) / tile);
    int ctas_per_sm = 1;
    knn2_kernel_info(p.cfg, nullptr, &ctas_per_sm, nullptr);
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    p.grid = h->sm_count * ctas_per_sm;
    plan_chunks(p.n_tiles, nB, p.grid, &p.n_chunks, &p.rows_per_chunk);
    p.slot_stride = (nA + 31) & ~(size_t)31;
    const uint64_t items = (uint64_t)p.n_tiles * p.n_chunks;
    if ((uint64_t)p.grid > items) p.grid = (int)std::max<uint64_t>(items, 1);
    return p;
}

// ------------------------------------------------------------- K1t tile images
void tc_image_register(hulo_gpu *h, const void *rows) {
    if (h->tc_qitems_key.rows == rows) h->tc_qitems_key.valid = false;
    for (auto &e : h->tc_images) if (e.rows == rows) e.valid = false;
    for (auto &e : h->tc_images) if (e.rows == rows && e.kind == hulo_gpu::kTcFlat4) return;
    h->tc_images.push_back(hulo_gpu::TcImage{rows, 0, false, hulo_gpu::kTcFlat4, {}, DevBuf{}});
}
void tc_image_invalidate(hulo_gpu *h, const void *rows) {
    for (auto &e : h->tc_images) if (e.rows == rows) e.valid = false;
    if (h->tc_qitems_key.rows == rows) h->tc_qitems_key.valid = false;
}
void tc_image_drop(hulo_gpu *h, const void *rows) {
    if (h->tc_qitems_key.rows == rows) h->tc_qitems_key.valid = false;
    for (size_t k = h->tc_images.size(); k-- > 0;)
        if (h->tc_images[k].rows == rows) {
            h->tc_images[k].img.release();
            h->tc_images.erase(h->tc_images.begin() + (long)k);
        }
}
// Tile image of `n` folded rows at `rows` (kind: kTcFlat8 or kTcFlat4): the cached one when the rows
// are a registered table, else expanded into `scratch`.
static int tc_image_for(hulo_gpu *h, const uint4 *rows, size_t n, int kind, DevBuf &scratch, const uint8_t **out) {
    const size_t bytes = kind == hulo_gpu::kTcFlat4 ? knn2_tc4_image_bytes(n) : knn2_tc_image_bytes(n);
    auto expand = [&](uint8_t *img) {
        return kind == hulo_gpu::kTcFlat4 ? knn2_tc4_expand_launch(rows, n, img, h->stream)
                                          : knn2_tc_expand_launch(rows, n, img, h->stream);
    };
    bool registered = false;
    for (auto &e : h->tc_images) registered = registered || e.rows == rows;
    if (registered) {
        hulo_gpu::TcImage *e = nullptr;
        for (auto &x : h->tc_images) if (x.rows == rows && x.kind == kind) e = &x;
        if (!e) {
            h->tc_images.push_back(hulo_gpu::TcImage{rows, 0, false, kind, {}, DevBuf{}});
            e = &h->tc_images.back();
        }
        if (!e->valid || e->n != n) {
            HULO_CUDA(e->img.reserve(bytes));
            HULO_CUDA(expand(e->img.as<uint8_t>()));
            h->launches++;
            e->n = n;
            e->valid = true;
        }
        *out = e->img.as<uint8_t>();
        return HULO_OK;
    }
    HULO_CUDA(scratch.reserve(bytes));
    HULO_CUDA(expand(scratch.as<uint8_t>()));
    h->launches++;
    *out = scratch.as<uint8_t>();
    return HULO_OK;
}

// Segmented tile image: every segment [seg[s], seg[s + 1]) starts on an even tile, so that it can be
// read as searcher (128-row tiles) and as database (256-row tiles).  Fills tile0 (n_seg + 1 entries).
static int tc_build_seg_image(hulo_gpu *h, const uint4 *rows, const uint64_t *seg, size_t n_seg, DevBuf &img,
                              std::vector<uint32_t> &tile0) {
    tile0.assign(n_seg + 1, 0);
    std::vector<uint32_t> tab;                 // tile_src | tile_rows
    for (size_t s = 0; s < n_seg; ++s) tile0[s + 1] = tile0[s] + 2u * (uint32_t)((seg[s + 1] - seg[s] + kTcTileN - 1) / kTcTileN);
    const size_t n_tiles = std::max<size_t>(tile0[n_seg], 2);
    tab.assign(2 * n_tiles, 0);
    for (size_t s = 0; s < n_seg; ++s) {
        const uint64_t n = seg[s + 1] - seg[s];
        for (uint32_t t = tile0[s]; t < tile0[s + 1]; ++t) {
            const uint64_t first = (uint64_t)(t - tile0[s]) * kTcTileRows;
            tab[t] = (uint32_t)(seg[s] + std::min<uint64_t>(first, n));
            tab[n_tiles + t] = first < n ? (uint32_t)std::min<uint64_t>(kTcTileRows, n - first) : 0u;
        }
    }
    HULO_CUDA(img.reserve(n_tiles * kTcTileBytes));
    HULO_CUDA(h->tc_tiles.reserve(tab.size() * sizeof(uint32_t)));
    HULO_CUDA(cudaMemcpyAsync(h->tc_tiles.ptr, tab.data(), tab.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(knn2_tc_expand_tiles_launch(rows, h->tc_tiles.as<uint32_t>(), h->tc_tiles.as<uint32_t>() + n_tiles, n_tiles,
                                          img.as<uint8_t>(), h->stream));
    h->launches++;
    // the table is host memory of this call: the copy must have left it before it is reused
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}

// The same for K1t4: every segment starts on an 8-row group; grp0[s] = its first group.
static int tc_build_seg_image4(hulo_gpu *h, const uint4 *rows, const uint64_t *seg, size_t n_seg, DevBuf &img,
                               std::vector<uint32_t> &grp0) {
    grp0.assign(n_seg + 1, 0);
    for (size_t s = 0; s < n_seg; ++s) grp0[s + 1] = grp0[s] + (uint32_t)((seg[s + 1] - seg[s] + 7) / 8);
    const size_t n_groups = grp0[n_seg];
    std::vector<uint32_t> src(std::max<size_t>(n_groups, 1), 0);
    std::vector<uint8_t> cnt(std::max<size_t>(n_groups, 1), 0);
    for (size_t s = 0; s < n_seg; ++s) {
        const uint64_t n = seg[s + 1] - seg[s];
        for (uint32_t g = grp0[s]; g < grp0[s + 1]; ++g) {
            const uint64_t first = (uint64_t)(g - grp0[s]) * 8;
            src[g] = (uint32_t)(seg[s] + first);
            cnt[g] = (uint8_t)std::min<uint64_t>(8, n - first);
        }
    }
    HULO_CUDA(img.reserve(knn2_tc4_groups_image_bytes(n_groups)));
    const size_t cnt_off = (src.size() * sizeof(uint32_t) + 15) & ~(size_t)15;
    HULO_CUDA(h->tc_tiles.reserve(cnt_off + cnt.size()));
    HULO_CUDA(cudaMemcpyAsync(h->tc_tiles.ptr, src.data(), src.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(cudaMemcpyAsync(h->tc_tiles.as<uint8_t>() + cnt_off, cnt.data(), cnt.size(), cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(knn2_tc4_expand_groups_launch(rows, h->tc_tiles.as<uint32_t>(), h->tc_tiles.as<uint8_t>() + cnt_off, n_groups,
                                            img.as<uint8_t>(), h->stream));
    h->launches++;
    HULO_CUDA(cudaStreamSynchronize(h->stream));      // the tables are host memory of this call
    return HULO_OK;
}

// Operand width of the tensor-core engine on this context: 4-bit (K1t4) unless HULO_KNN_TC8 / HULO_TC_BITS=8.
static bool tc_items_four(const hulo_gpu *h) { return h->tc_bits != 8; }
// A searcher tile / database range of a segmented image in the units of the engine: image tiles of
// 128 rows (K1t) or groups of 8 rows (K1t4); `first[s]` = tile0 or grp0 of the segment.
static inline uint32_t tc_unit_of_row(const hulo_gpu *h, const uint32_t *first, size_t s, uint64_t row_in_seg) {
    return first[s] + (uint32_t)(row_in_seg / (tc_items_four(h) ? 8u : kTcTileRows));
}

// The cached segmented image of a resident table.
static int tc_seg_image_for_db(hulo_gpu *h, const hulo_db *db, const uint8_t **img, const uint32_t **tile0) {
    hulo_gpu::TcImage *e = nullptr;
    const int kind = tc_items_four(h) ? hulo_gpu::kTcSeg4 : hulo_gpu::kTcSeg8;
    for (auto &x : h->tc_images) if (x.rows == db->rows && x.kind == kind) e = &x;
    if (!e) {
        h->tc_images.push_back(hulo_gpu::TcImage{db->rows, 0, false, kind, {}, DevBuf{}});
        e = &h->tc_images.back();
    }
    if (!e->valid || e->n != db->n || e->tile0.size() != db->seg.size()) {
        int rc = tc_items_four(h) ? tc_build_seg_image4(h, db->rows, db->seg.data(), db->seg.size() - 1, e->img, e->tile0)
                                 : tc_build_seg_image(h, db->rows, db->seg.data(), db->seg.size() - 1, e->img, e->tile0);
        if (rc != HULO_OK) return rc;
        e->n = db->n;
        e->valid = true;
    }
    *img = e->img.as<uint8_t>();
    *tile0 = e->tile0.data();
    return HULO_OK;
}

// Engine choice for a search of `dist` distance evaluations.
static bool tc_wanted(const hulo_gpu *h, uint64_t dist, uint64_t min_dist) {
    return h->knn_engine == HULO_KNN_TC || (h->knn_engine == HULO_KNN_AUTO && dist >= min_dist);
}

// K1t over an item list that is already on the device; the caller has reserved h->partial.
static int run_items_tc_dev(hulo_gpu *h, const uint8_t *imgA, const uint8_t *imgB, const TcItem *d_items, size_t n_items);

// K1t over an item list; the caller has reserved h->partial.
static int run_items_tc(hulo_gpu *h, const uint8_t *imgA, const uint8_t *imgB, const std::vector<TcItem> &items) {
    if (items.empty()) return HULO_OK;
    HULO_CUDA(h->items.reserve(items.size() * sizeof(TcItem)));
    HULO_CUDA(cudaMemcpyAsync(h->items.ptr, items.data(), items.size() * sizeof(TcItem), cudaMemcpyHostToDevice, h->stream));
    return run_items_tc_dev(h, imgA, imgB, h->items.as<TcItem>(), items.size());
}

static int run_items_tc_dev(hulo_gpu *h, const uint8_t *imgA, const uint8_t *imgB, const TcItem *d_items, size_t n_items) {
    if (n_items == 0) return HULO_OK;
    TcParams tp{};
    tp.imgA = imgA; tp.imgB = imgB;
    tp.items = d_items;
    tp.n_items = (uint32_t)n_items;
    tp.partial = h->partial.as<uint2>();
    tp.cluster = 1;
    if (tc_items_four(h)) HULO_CUDA(knn2_tc4_launch(tp, h->sm_count, h->stream));
    else HULO_CUDA(knn2_tc_launch(tp, h->sm_count, h->stream));
    h->launches++;
    return HULO_OK;
}

// K1t for a flat searcher table against a flat database: same partial-key format as K1.  The 4-bit
// form (K1t4, kind::mxf4) by default; HULO_TC_BITS=8 selects the int8 form (K1t, kind::i8).
static int run_flat_k1_tc(hulo_gpu *h, const uint4 *A, size_t nA, const uint4 *B, size_t nB, FlatRun *run) {
    const bool four = tc_items_four(h);
    const int kind = four ? hulo_gpu::kTcFlat4 : hulo_gpu::kTcFlat8;
    uint32_t n_mtiles = 0, n_chunks = 0, rpc = 0;
    if (four) knn2_tc4_plan(nA, nB, h->sm_count, &n_mtiles, &n_chunks, &rpc);
    else knn2_tc_plan(nA, nB, h->sm_count, &n_mtiles, &n_chunks, &rpc);
    run->n_chunks = n_chunks; run->rows_per_chunk = rpc; run->slot_stride = (nA + 31) & ~(size_t)31;
    if (n_chunks == 0) return HULO_OK;
    const uint8_t *imgA = nullptr, *imgB = nullptr;
    int rc = tc_image_for(h, A, nA, kind, h->tc_scratchA, &imgA);
    if (rc != HULO_OK) return rc;
    rc = tc_image_for(h, B, nB, kind, h->tc_scratchB, &imgB);
    if (rc != HULO_OK) return rc;
    HULO_CUDA(h->partial.reserve((size_t)n_chunks * run->slot_stride * sizeof(uint2)));
    TcParams tp{};
    tp.imgA = imgA; tp.imgB = imgB;
    tp.nA = (uint32_t)nA; tp.nB = (uint32_t)nB;
    tp.n_mtiles = n_mtiles; tp.n_chunks = n_chunks; tp.rows_per_chunk = rpc;
    tp.slot_stride = run->slot_stride;
    tp.partial = h->partial.as<uint2>();
    if (four) {
        HULO_CUDA(knn2_tc4_launch(tp, h->sm_count, h->stream));
    } else {
        static const int cluster = env_int("HULO_TC_CLUSTER", 1);      // tuning sweeps
        tp.cluster = cluster;
        HULO_CUDA(knn2_tc_launch(tp, h->sm_count, h->stream));
    }
    h->launches++;
    return HULO_OK;
}

// K1 for a flat searcher table against a flat database: per-chunk keys left in h->partial.
int run_flat_k1(hulo_gpu *h, const uint4 *A, size_t nA, const uint4 *B, size_t nB, uint32_t row_base,
                FlatRun *run) {
    HULO_ARG(nA < (size_t)INT_MAX && nB + (size_t)row_base < (size_t)INT_MAX, "table too large for int32 indices");
    // a plain search would overwrite the result set of the latest submitted one
    HULO_ARG(h->in_submit || h->pipe_submitted == h->pipe_collected,
             "searches issued with hulo_knn2_sharded_submit are outstanding: collect them first");
    HULO_CUDA(h->knn_idx.reserve(std::max<size_t>(nA, 1) * 2 * sizeof(int32_t)));
    HULO_CUDA(h->knn_dist.reserve(std::max<size_t>(nA, 1) * 2 * sizeof(int32_t)));
    h->last_nA = nA;
    run->n_chunks = 0; run->rows_per_chunk = 1; run->slot_stride = 0;
    if (nA == 0) return HULO_OK;
    const bool big = nA >= 128 && nB >= 8192 && (uint64_t)nA * nB >= (1ull << 28);
    if (h->knn_engine == HULO_KNN_TC || (h->knn_engine == HULO_KNN_AUTO && big)) return run_flat_k1_tc(h, A, nA, B, nB, run);
    FlatPlan pl = plan_flat(h, nA, nB);
    run->n_chunks = pl.n_chunks; run->rows_per_chunk = pl.rows_per_chunk; run->slot_stride = pl.slot_stride;
    if (pl.n_chunks > 0) {
        HULO_CUDA(h->partial.reserve((size_t)pl.n_chunks * pl.slot_stride * sizeof(uint2)));
        HULO_CUDA(h->counter.reserve(sizeof(unsigned int)));
        HULO_CUDA(cudaMemsetAsync(h->counter.ptr, 0, sizeof(unsigned int), h->stream));
        KnnParams kp{};
        kp.A = A; kp.B = B; kp.items = nullptr;
        kp.n_items = pl.n_tiles * pl.n_chunks;
        kp.nA = (uint32_t)nA; kp.nB = (uint32_t)nB;
        kp.n_tiles = pl.n_tiles; kp.rows_per_chunk = pl.rows_per_chunk;
        kp.slot_stride = pl.slot_stride;
        kp.key_unit = 1u << kKeyIdxBits;
        kp.partial = h->partial.as<uint2>();
        kp.counter = h->counter.as<unsigned int>();
        HULO_CUDA(knn2_launch(kp, pl.cfg, pl.grid, h->stream));
        h->launches++;
    }
    return HULO_OK;
}

// K1 + chunk merge, results left in h->knn_idx / h->knn_dist (and h->packed when want_packed).
static int run_flat(hulo_gpu *h, const uint4 *A, size_t nA, const uint4 *B, size_t nB, uint32_t row_base,
                    bool want_packed) {
    FlatRun run;
    int rc = run_flat_k1(h, A, nA, B, nB, row_base, &run);
    if (rc != HULO_OK) return rc;
    if (want_packed) HULO_CUDA(h->packed.reserve(std::max<size_t>(nA, 1) * sizeof(int4)));
    if (nA == 0) return HULO_OK;
    HULO_CUDA(knn2_merge_launch(h->partial.as<uint2>(), (uint32_t)nA, run.n_chunks, run.slot_stride,
                                run.rows_per_chunk, row_base, h->knn_idx.as<int32_t>(),
                                h->knn_dist.as<int32_t>(), want_packed ? h->packed.as<int4>() : nullptr,
                                h->stream));
    h->launches++;
    return HULO_OK;
}

// Expand rows of `stride` bytes into zero-padded 64-byte rows inside a pinned staging buffer.
static const uint8_t *stage_rows(HostBuf &hb, const uint8_t *rows, size_t n, size_t stride, cudaError_t *err) {
    *err = cudaSuccess;
    if (stride == HULO_ROW_BYTES) return rows;
    *err = hb.reserve(std::max<size_t>(n, 1) * HULO_ROW_BYTES);
    if (*err != cudaSuccess) return nullptr;
    uint8_t *dst = hb.as<uint8_t>();
    const size_t w = std::min<size_t>(stride, HULO_ROW_BYTES);
    for (size_t i = 0; i < n; ++i) {
        memcpy(dst + i * HULO_ROW_BYTES, rows + i * stride, w);
        if (w < HULO_ROW_BYTES) memset(dst + i * HULO_ROW_BYTES + w, 0, HULO_ROW_BYTES - w);
    }
    return dst;
}

// The same into pinned memory for every stride, with the fold of the device layout (knn2.cuh)
// applied on the way: a query image is a few thousand rows, the host folds them while it copies and
// the upload needs neither a pageable transfer nor the fold kernel.
static const uint8_t *stage_rows_folded(HostBuf &hb, const uint8_t *rows, size_t n, size_t stride, cudaError_t *err) {
    *err = hb.reserve(std::max<size_t>(n, 1) * HULO_ROW_BYTES);
    if (*err != cudaSuccess) return nullptr;
    uint8_t *dst = hb.as<uint8_t>();
    const size_t w = std::min<size_t>(stride, HULO_ROW_BYTES);
    for (size_t i = 0; i < n; ++i) {
        uint32_t r[16];
        memcpy(r, rows + i * stride, w);
        if (w < HULO_ROW_BYTES) memset(reinterpret_cast<uint8_t *>(r) + w, 0, HULO_ROW_BYTES - w);
        for (int k = 0; k < 5; ++k) r[3 * k + 2] ^= r[3 * k] ^ r[3 * k + 1];
        r[15] ^= r[11] ^ r[14];                                   // folded w11, w14: w15 = w9 ^ ... ^ w15
        memcpy(dst + i * HULO_ROW_BYTES, r, HULO_ROW_BYTES);
    }
    return dst;
}

int run_flat_packed(hulo_gpu *h, const uint4 *A, size_t nA, const uint4 *B, size_t nB, uint32_t row_base) {
    return run_flat(h, A, nA, B, nB, row_base, true);
}

}  // namespace hulo

using namespace hulo;

extern "C" {

int hulo_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *hulo_last_error(void) { return g_error.c_str(); }
const char *hulo_version(void) { return "sfmlocalization_b200 0.1 (sm_100a)"; }

void hulo_gpu_destroy(hulo_gpu *h);

int hulo_gpu_create(int device, hulo_gpu **out) {
    HULO_ARG(out != nullptr, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        set_error("hulo_gpu_create: no CUDA device (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return HULO_ERR_CUDA;
    }
    HULO_ARG(device >= 0 && device < n, "device index out of range");
    HULO_CUDA(cudaSetDevice(device));
    hulo_gpu *h = new (std::nothrow) hulo_gpu();
    HULO_ARG(h != nullptr, "out of host memory");
    h->device = device;
    cudaDeviceProp prop;
    // from here on a failure releases what exists already (hulo_gpu_destroy copes with a half-built context)
#define HULO_CREATE_CUDA(expr)                                                                                    \
    do {                                                                                                          \
        cudaError_t e__ = (expr);                                                                                 \
        if (e__ != cudaSuccess) {                                                                                 \
            hulo_gpu_destroy(h);                                                                                  \
            hulo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(e__));               \
            return HULO_ERR_CUDA;                                                                                 \
        }                                                                                                         \
    } while (0)
    HULO_CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("hulo_gpu_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                  prop.major, prop.minor);
        delete h;
        return HULO_ERR_CUDA;
    }
    h->sm_count = prop.multiProcessorCount;
    HULO_CREATE_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    HULO_CREATE_CUDA(cudaEventCreate(&h->ev_start));
    HULO_CREATE_CUDA(cudaEventCreate(&h->ev_stop));
#undef HULO_CREATE_CUDA
    // tuning overrides (benchmark sweeps): HULO_KNN_THREADS / HULO_KNN_QPT / HULO_KNN_CSA
    const int t = env_int("HULO_KNN_THREADS", 0), q = env_int("HULO_KNN_QPT", 0), c = env_int("HULO_KNN_CSA", -1);
    const int o = env_int("HULO_KNN_OPT", -1);
    if (t > 0 || q > 0 || c >= 0 || o >= 0) {
        if (t > 0) h->knn_cfg.threads = t;
        if (q > 0) h->knn_cfg.qpt = q;
        if (c >= 0) h->knn_cfg.csa = c;
        if (o >= 0) h->knn_cfg.opt = o;
        h->knn_cfg_forced = true;
        if (knn2_kernel_info(h->knn_cfg, nullptr, nullptr, nullptr) != cudaSuccess) {
            cudaGetLastError();
            set_error("hulo_gpu_create: no K1 variant threads=%d qpt=%d csa=%d opt=%d", h->knn_cfg.threads,
                      h->knn_cfg.qpt, h->knn_cfg.csa, h->knn_cfg.opt);
            hulo_gpu_destroy(h);
            return HULO_ERR_ARG;
        }
    }
    {
        h->tc_bits = env_int("HULO_TC_BITS", 4) == 8 ? 8 : 4;
        const char *eng = getenv("HULO_KNN_ENGINE");
        if (eng && (!strcmp(eng, "tc") || !strcmp(eng, "1"))) h->knn_engine = HULO_KNN_TC;
        else if (eng && (!strcmp(eng, "int") || !strcmp(eng, "0"))) h->knn_engine = HULO_KNN_INT;
        else if (eng && *eng && strcmp(eng, "auto") && strcmp(eng, "2")) {
            set_error("hulo_gpu_create: HULO_KNN_ENGINE=%s (expected int, tc or auto)", eng);
            hulo_gpu_destroy(h);
            return HULO_ERR_ARG;
        }
    }
    *out = h;
    return HULO_OK;
}

int hulo_gpu_set_knn_engine(hulo_gpu *h, int engine) {
    HULO_ARG(h != nullptr, "null context");
    HULO_ARG(engine == HULO_KNN_INT || engine == HULO_KNN_TC || engine == HULO_KNN_AUTO || engine == HULO_KNN_TC8,
             "unknown engine");
    if (engine == HULO_KNN_TC8) { h->knn_engine = HULO_KNN_TC; h->tc_bits = 8; }
    else { h->knn_engine = engine; if (engine == HULO_KNN_TC) h->tc_bits = 4; }
    return HULO_OK;
}
int hulo_gpu_knn_engine(const hulo_gpu *h) {
    if (!h) return -1;
    return h->knn_engine == HULO_KNN_TC && h->tc_bits == 8 ? HULO_KNN_TC8 : h->knn_engine;
}

void hulo_comm_destroy_internal(hulo_gpu *h);

void hulo_gpu_destroy(hulo_gpu *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    hulo_comm_destroy_internal(h);
    if (h->xstream) cudaStreamSynchronize(h->xstream);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int p = 0; p < 2; ++p) {
        if (h->ev_k1[p]) cudaEventDestroy(h->ev_k1[p]);
        if (h->ev_x[p]) cudaEventDestroy(h->ev_x[p]);
    }
    if (h->xstream) cudaStreamDestroy(h->xstream);
    if (h->cstream) { cudaStreamSynchronize(h->cstream); cudaStreamDestroy(h->cstream); }
    for (int q = 0; q < 2; ++q)
        if (h->ev_res[q]) cudaEventDestroy(h->ev_res[q]);
    DevBuf *bufs[] = {&h->partial, &h->counter, &h->items, &h->knn_idx, &h->knn_dist, &h->packed, &h->gathered,
                      &h->stageA, &h->stageB, &h->scratch0, &h->scratch1, &h->scratch2, &h->scratch3, &h->lfact, &h->wave_counter, &h->comm_scratch, &h->partial_alt, &h->knn_idx_alt, &h->knn_dist_alt};
    for (DevBuf *b : bufs) b->release();
    for (auto &e : h->tc_images) e.img.release();
    h->tc_scratchA.release();
    h->tc_scratchB.release();
    h->tc_tiles.release();
    h->tc_qitems.release();
    h->wave_rec.release();
    h->hstage0.release();
    h->hstage1.release();
    if (h->ev_start) cudaEventDestroy(h->ev_start);
    if (h->ev_stop) cudaEventDestroy(h->ev_stop);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int hulo_host_alloc(size_t bytes, void **out) {
    HULO_ARG(out != nullptr, "out is null");
    HULO_CUDA(cudaMallocHost(out, std::max<size_t>(bytes, 1)));
    return HULO_OK;
}
void hulo_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int hulo_timer_start(hulo_gpu *h) {
    HULO_ARG(h != nullptr, "null context");
    HULO_CUDA(cudaEventRecord(h->ev_start, h->stream));
    return HULO_OK;
}
int hulo_timer_stop(hulo_gpu *h, float *ms) {
    HULO_ARG(h != nullptr && ms != nullptr, "null argument");
    HULO_CUDA(join_exchange(h));
    HULO_CUDA(cudaEventRecord(h->ev_stop, h->stream));
    HULO_CUDA(cudaEventSynchronize(h->ev_stop));
    HULO_CUDA(cudaEventElapsedTime(ms, h->ev_start, h->ev_stop));
    return HULO_OK;
}
int hulo_synchronize(hulo_gpu *h) {
    HULO_ARG(h != nullptr, "null context");
    HULO_CUDA(join_exchange(h));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}
uint64_t hulo_launch_count(const hulo_gpu *h) { return h ? h->launches : 0; }

// ------------------------------------------------------------ descriptor rows
int hulo_db_upload(hulo_gpu *h, const uint8_t *rows, size_t n, size_t stride, const uint64_t *seg_offsets,
                   size_t n_seg, hulo_db **out) {
    HULO_ARG(h != nullptr && out != nullptr, "null argument");
    *out = nullptr;
    HULO_ARG(n == 0 || rows != nullptr, "rows is null");
    HULO_ARG(stride >= 1, "stride must be >= 1");
    HULO_ARG(n < (size_t)INT_MAX, "too many rows");
    if (seg_offsets) {
        HULO_ARG(n_seg >= 1, "n_seg must be >= 1 when seg_offsets is given");
        HULO_ARG(seg_offsets[0] == 0 && seg_offsets[n_seg] == n, "seg_offsets must start at 0 and end at n");
        for (size_t s = 0; s < n_seg; ++s) HULO_ARG(seg_offsets[s] <= seg_offsets[s + 1], "seg_offsets not ascending");
    }
    HULO_CUDA(cudaSetDevice(h->device));
    hulo_db *db = new (std::nothrow) hulo_db();
    HULO_ARG(db != nullptr, "out of host memory");
    db->owner = h;
    db->n = n;
    db->cap_rows = std::max<size_t>(n, 1);
    cudaError_t e = cudaMalloc(&db->rows, db->cap_rows * HULO_ROW_BYTES);
    if (e != cudaSuccess) {
        set_error("hulo_db_upload: cudaMalloc(%zu) -> %s", db->cap_rows * (size_t)HULO_ROW_BYTES, cudaGetErrorString(e));
        delete db;
        return HULO_ERR_CUDA;
    }
    if (seg_offsets) db->seg.assign(seg_offsets, seg_offsets + n_seg + 1);
    else db->seg = {0, (uint64_t)n};
    tc_image_register(h, db->rows);
    int rc = hulo_db_update(h, db, rows, n, stride);
    if (rc != HULO_OK) { hulo_db_free(db); return rc; }
    *out = db;
    return HULO_OK;
}

int hulo_db_update(hulo_gpu *h, hulo_db *db, const uint8_t *rows, size_t n, size_t stride) {
    HULO_ARG(h != nullptr && db != nullptr, "null argument");
    HULO_ARG(n <= db->cap_rows, "more rows than the table was created with");
    HULO_ARG(n == 0 || rows != nullptr, "rows is null");
    HULO_ARG(stride >= 1, "stride must be >= 1");
    if (n != db->n) { db->n = n; db->seg = {0, (uint64_t)n}; }
    tc_image_invalidate(h, db->rows);
    if (n == 0) return HULO_OK;
    if (stride == HULO_ROW_BYTES) {
        HULO_CUDA(cudaMemcpyAsync(db->rows, rows, n * HULO_ROW_BYTES, cudaMemcpyHostToDevice, h->stream));
    } else if (stride > HULO_ROW_BYTES) {
        HULO_CUDA(cudaMemcpy2DAsync(db->rows, HULO_ROW_BYTES, rows, stride, HULO_ROW_BYTES, n,
                                    cudaMemcpyHostToDevice, h->stream));
    } else {
        HULO_CUDA(cudaMemsetAsync(db->rows, 0, n * HULO_ROW_BYTES, h->stream));
        HULO_CUDA(cudaMemcpy2DAsync(db->rows, HULO_ROW_BYTES, rows, stride, stride, n, cudaMemcpyHostToDevice,
                                    h->stream));
    }
    HULO_CUDA(knn2_fold_rows_launch(db->rows, n, h->stream));     // device layout of K1 (knn2.cuh)
    h->launches++;
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}

void hulo_db_free(hulo_db *db) {
    if (!db) return;
    if (db->owner) cudaSetDevice(db->owner->device);
    if (db->owner) tc_image_drop(db->owner, db->rows);
    if (db->rows) cudaFree(db->rows);
    delete db;
}
size_t hulo_db_rows(const hulo_db *db) { return db ? db->n : 0; }
size_t hulo_db_segments(const hulo_db *db) { return db ? db->seg.size() - 1 : 0; }

int hulo_db_download(hulo_gpu *h, const hulo_db *db, size_t first, size_t n, uint8_t *rows64) {
    HULO_ARG(h != nullptr && db != nullptr, "null argument");
    HULO_ARG(first + n <= db->n, "row range out of bounds");
    if (n == 0) return HULO_OK;
    HULO_ARG(rows64 != nullptr, "rows64 is null");
    HULO_CUDA(cudaMemcpyAsync(rows64, db->rows + first * 4, n * HULO_ROW_BYTES, cudaMemcpyDeviceToHost, h->stream));
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    knn2_unfold_rows_host(rows64, n);                             // undo the device layout
    return HULO_OK;
}

// ------------------------------------------------------------------- K1: 2-NN
int hulo_knn2_fetch(hulo_gpu *h, size_t nA, int32_t *idx2, int32_t *dist2) {
    HULO_ARG(h != nullptr, "null context");
    HULO_ARG(nA <= h->last_nA, "more rows requested than the last search produced");
    HULO_CUDA(join_exchange(h));
    if (nA > 0) {
        if (idx2) HULO_CUDA(cudaMemcpyAsync(idx2, h->knn_idx.ptr, nA * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        if (dist2) HULO_CUDA(cudaMemcpyAsync(dist2, h->knn_dist.ptr, nA * 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    return HULO_OK;
}

int hulo_knn2(hulo_gpu *h, const hulo_db *A, const hulo_db *B, int32_t *idx2, int32_t *dist2) {
    HULO_ARG(h != nullptr && A != nullptr && B != nullptr, "null argument");
    HULO_CUDA(cudaSetDevice(h->device));
    HULO_CUDA(join_exchange(h));
    int rc = run_flat(h, A->rows, A->n, B->rows, B->n, 0, false);
    if (rc != HULO_OK) return rc;
    if (idx2 || dist2) return hulo_knn2_fetch(h, A->n, idx2, dist2);
    return HULO_OK;
}

int hulo_knn2_host(hulo_gpu *h, const uint8_t *A, size_t nA, size_t strideA, const uint8_t *B, size_t nB,
                   size_t strideB, int32_t *idx2, int32_t *dist2) {
    HULO_ARG(h != nullptr, "null context");
    HULO_ARG((nA == 0 || A != nullptr) && (nB == 0 || B != nullptr), "null rows");
    HULO_ARG(strideA >= 1 && strideB >= 1, "stride must be >= 1");
    HULO_ARG(nA == 0 || (idx2 != nullptr && dist2 != nullptr), "null output");
    HULO_CUDA(cudaSetDevice(h->device));
    HULO_CUDA(join_exchange(h));
    HULO_CUDA(h->stageA.reserve(std::max<size_t>(nA, 1) * HULO_ROW_BYTES));
    HULO_CUDA(h->stageB.reserve(std::max<size_t>(nB, 1) * HULO_ROW_BYTES));
    cudaError_t e;
    const uint8_t *a64 = stage_rows(h->hstage0, A, nA, strideA, &e);
    HULO_CUDA(e);
    const uint8_t *b64 = stage_rows(h->hstage1, B, nB, strideB, &e);
    HULO_CUDA(e);
    if (nA) HULO_CUDA(cudaMemcpyAsync(h->stageA.ptr, a64, nA * HULO_ROW_BYTES, cudaMemcpyHostToDevice, h->stream));
    if (nB) HULO_CUDA(cudaMemcpyAsync(h->stageB.ptr, b64, nB * HULO_ROW_BYTES, cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(knn2_fold_rows_launch(h->stageA.as<uint4>(), nA, h->stream));
    HULO_CUDA(knn2_fold_rows_launch(h->stageB.as<uint4>(), nB, h->stream));
    h->launches += (nA ? 1 : 0) + (nB ? 1 : 0);
    int rc = run_flat(h, h->stageA.as<uint4>(), nA, h->stageB.as<uint4>(), nB, 0, false);
    if (rc != HULO_OK) return rc;
    return hulo_knn2_fetch(h, nA, idx2, dist2);
}

// --------------------------------------------- query localisation, reference direction
}  // extern "C"

// hulo_match_to_query up to and including the compaction: the survivors stay on the device
// (out->d_total == nullptr when nothing was searched).  The stream is NOT synchronised.
int hulo::match_to_query_device(hulo_gpu *h, const hulo_db *map, const uint32_t *views, size_t n_views,
                                const uint8_t *query, size_t nq, size_t q_stride, float ratio, QueryMatchesDev *out) {
    HULO_ARG(h != nullptr && map != nullptr && out != nullptr, "null argument");
    HULO_ARG(nq == 0 || query != nullptr, "query is null");
    HULO_ARG(q_stride >= 1, "stride must be >= 1");
    HULO_ARG(nq <= kMaxChunkRows * 64ull, "query too large");
    *out = QueryMatchesDev{};
    HULO_CUDA(cudaSetDevice(h->device));
    HULO_CUDA(join_exchange(h));
    const size_t n_seg = map->seg.size() - 1;
    if (views == nullptr) n_views = n_seg;
    for (size_t v = 0; views && v < n_views; ++v) HULO_ARG(views[v] < n_seg, "view index out of range");
    out->n_views = n_views;
    // MatchUtils.cpp:299-301: nothing to do without query rows
    if (nq < 1) return HULO_OK;

    // upload the query rows.  They stay in the staging buffer, folded, until the next matching
    // call: the engine's guided matching reads them there (also on a rank of a view-sharded query
    // that was given no view, hence before the early returns).
    HULO_CUDA(h->stageB.reserve(nq * HULO_ROW_BYTES));
    cudaError_t e;
    if (nq <= 65536) {
        const uint8_t *q64 = stage_rows_folded(h->hstage1, query, nq, q_stride, &e);
        HULO_CUDA(e);
        HULO_CUDA(cudaMemcpyAsync(h->stageB.ptr, q64, nq * HULO_ROW_BYTES, cudaMemcpyHostToDevice, h->stream));
    } else {
        const uint8_t *q64 = stage_rows(h->hstage1, query, nq, q_stride, &e);
        HULO_CUDA(e);
        HULO_CUDA(cudaMemcpyAsync(h->stageB.ptr, q64, nq * HULO_ROW_BYTES, cudaMemcpyHostToDevice, h->stream));
        HULO_CUDA(knn2_fold_rows_launch(h->stageB.as<uint4>(), nq, h->stream));
        h->launches++;
    }
    if (n_views == 0) return hulo_synchronize(h);

    // compact row space: the rows of the selected views, concatenated in the order given
    std::vector<uint64_t> sel_off(n_views + 1, 0);
    for (size_t v = 0; v < n_views; ++v) {
        const size_t s = views ? views[v] : v;
        sel_off[v + 1] = sel_off[v] + (map->seg[s + 1] - map->seg[s]);
    }
    const uint64_t n_rows = sel_off[n_views];
    HULO_ARG(n_rows < (uint64_t)INT_MAX, "too many rows selected");
    if (n_rows == 0) return hulo_synchronize(h);

    const uint64_t slot_stride = (n_rows + 31) & ~31ull;
    uint32_t n_chunks = 0, rows_per_chunk = 1;
    const bool use_tc = nq <= kMaxChunkRows && tc_wanted(h, n_rows * (uint64_t)nq, 1ull << 26);
    // scratch0: val | dist ; scratch1: seg offsets (device) ; scratch2: block counts, seg_out_off, total
    // scratch3: compacted outputs view | i | j | d0
    const size_t n_blocks = (n_rows + kCompactBlockRows - 1) / kCompactBlockRows;
    HULO_CUDA(h->scratch0.reserve(n_rows * 2 * sizeof(int32_t)));
    HULO_CUDA(h->scratch1.reserve((n_views + 1) * sizeof(uint64_t)));
    HULO_CUDA(h->scratch2.reserve((n_blocks + 2) * sizeof(uint32_t) + (n_views + 2) * sizeof(uint64_t) + 64));
    HULO_CUDA(h->scratch3.reserve(n_rows * 4 * sizeof(uint32_t)));
    HULO_CUDA(cudaMemcpyAsync(h->scratch1.ptr, sel_off.data(), (n_views + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
    if (use_tc) {
        // K1t: searcher tiles of 128 rows per view out of the map's segmented image, the query image
        // as one database range (a single chunk)
        const uint8_t *imgA = nullptr, *imgB = nullptr;
        const uint32_t *tile0 = nullptr;
        int rc = tc_seg_image_for_db(h, map, &imgA, &tile0);
        if (rc != HULO_OK) return rc;
        rc = tc_image_for(h, h->stageB.as<uint4>(), nq, tc_items_four(h) ? hulo_gpu::kTcFlat4 : hulo_gpu::kTcFlat8,
                          h->tc_scratchB, &imgB);
        if (rc != HULO_OK) return rc;
        n_chunks = 1; rows_per_chunk = kMaxChunkRows;
        // the item list stays on the device between calls with the same map, view list and query size
        hulo_gpu::TcItemCache &ck = h->tc_qitems_key;
        const int bits = tc_items_four(h) ? 4 : 8;
        const bool same_views = ck.all_views == (views == nullptr) && ck.n_views == n_views &&
                                (views == nullptr || std::equal(views, views + n_views, ck.views.begin()));
        if (!(ck.valid && ck.rows == map->rows && same_views && ck.nq == nq && ck.bits == bits)) {
            std::vector<TcItem> items;
            for (size_t v = 0; v < n_views; ++v) {
                const size_t s = views ? views[v] : v;
                const uint64_t rows = map->seg[s + 1] - map->seg[s];
                for (uint64_t t0 = 0; t0 < rows; t0 += kTcTileRows) {
                    TcItem it{};
                    it.a_tile = tc_unit_of_row(h, tile0, s, t0);
                    it.a_rows = (uint32_t)std::min<uint64_t>(kTcTileRows, rows - t0);
                    it.b_tile0 = 0;
                    it.b_rows = (uint32_t)nq;
                    it.out_slot0 = sel_off[v] + t0;
                    items.push_back(it);
                }
            }
            HULO_CUDA(h->tc_qitems.reserve(std::max<size_t>(items.size(), 1) * sizeof(TcItem)));
            HULO_CUDA(cudaMemcpyAsync(h->tc_qitems.ptr, items.data(), items.size() * sizeof(TcItem), cudaMemcpyHostToDevice,
                                      h->stream));
            HULO_CUDA(cudaStreamSynchronize(h->stream));          // `items` is about to go out of scope
            ck.rows = map->rows; ck.n_views = n_views; ck.nq = nq; ck.bits = bits;
            ck.all_views = views == nullptr;
            if (views) ck.views.assign(views, views + n_views); else ck.views.clear();
            ck.n_items = items.size();
            ck.valid = true;
        }
        HULO_CUDA(h->partial.reserve((size_t)slot_stride * sizeof(uint2)));
        rc = run_items_tc_dev(h, imgA, imgB, h->tc_qitems.as<TcItem>(), ck.n_items);
        if (rc != HULO_OK) return rc;
    } else {
    // items: maximal runs of views that are contiguous in the table, tiled, x chunks of the query
    const KnnConfig cfg = choose_config(h, (size_t)n_rows);
    const uint32_t tile = knn2_tile_rows(cfg);
    struct Run { uint64_t a0, rows, slot0; };
    std::vector<Run> runs;
    for (size_t v = 0; v < n_views; ++v) {
        const size_t s = views ? views[v] : v;
        const uint64_t a0 = map->seg[s], rows = map->seg[s + 1] - map->seg[s];
        if (rows == 0) continue;
        if (!runs.empty() && runs.back().a0 + runs.back().rows == a0) runs.back().rows += rows;
        else runs.push_back(Run{a0, rows, sel_off[v]});
    }
    uint64_t n_tiles = 0;
    for (const Run &r : runs) n_tiles += (r.rows + tile - 1) / tile;
    int ctas_per_sm = 1;
    knn2_kernel_info(cfg, nullptr, &ctas_per_sm, nullptr);
    int grid = h->sm_count * std::max(1, ctas_per_sm);
    plan_chunks((uint32_t)n_tiles, nq, grid, &n_chunks, &rows_per_chunk);
    std::vector<KnnItem> items;
    items.reserve((size_t)n_tiles * n_chunks);
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint32_t b0 = c * rows_per_chunk;
        const uint32_t brows = (uint32_t)std::min<uint64_t>(rows_per_chunk, nq - b0);
        for (const Run &r : runs)
            for (uint64_t t0 = 0; t0 < r.rows; t0 += tile) {
                KnnItem it{};
                it.a_row0 = (uint32_t)(r.a0 + t0);
                it.a_rows = (uint32_t)std::min<uint64_t>(tile, r.rows - t0);
                it.b_row0 = b0;
                it.b_rows = brows;
                it.out_slot0 = (uint64_t)c * slot_stride + r.slot0 + t0;
                items.push_back(it);
            }
    }
    grid = (int)std::min<size_t>((size_t)grid, items.size());

    HULO_CUDA(h->items.reserve(items.size() * sizeof(KnnItem)));
    HULO_CUDA(cudaMemcpyAsync(h->items.ptr, items.data(), items.size() * sizeof(KnnItem), cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(h->partial.reserve((size_t)n_chunks * slot_stride * sizeof(uint2)));
    HULO_CUDA(h->counter.reserve(sizeof(unsigned int)));
    HULO_CUDA(cudaMemsetAsync(h->counter.ptr, 0, sizeof(unsigned int), h->stream));

    KnnParams kp{};
    kp.A = map->rows; kp.B = h->stageB.as<uint4>();
    kp.items = h->items.as<KnnItem>();
    kp.n_items = (uint32_t)items.size();
    kp.key_unit = 1u << kKeyIdxBits;
    kp.partial = h->partial.as<uint2>();
    kp.counter = h->counter.as<unsigned int>();
    HULO_CUDA(knn2_launch(kp, cfg, grid, h->stream));
    h->launches++;
    }

    int32_t *val = h->scratch0.as<int32_t>();
    int32_t *dist = val + n_rows;
    HULO_CUDA(post_query_launch(h->partial.as<uint2>(), (uint32_t)n_rows, n_chunks, slot_stride, rows_per_chunk,
                                ratio, val, dist, h->stream));
    h->launches++;
    uint8_t *s2 = h->scratch2.as<uint8_t>();
    uint64_t *d_total = reinterpret_cast<uint64_t *>(s2);
    uint64_t *d_seg_out = d_total + 1;
    uint32_t *d_blocks = reinterpret_cast<uint32_t *>(d_seg_out + (n_views + 1));
    uint32_t *o_view = h->scratch3.as<uint32_t>();
    uint32_t *o_i = o_view + n_rows, *o_j = o_i + n_rows;
    int32_t *o_d = reinterpret_cast<int32_t *>(o_j + n_rows);
    { int nl = 0; HULO_CUDA(compact_launch(val, dist, (uint32_t)n_rows, h->scratch1.as<uint64_t>(), (uint32_t)n_views, d_blocks,
                             o_view, o_i, o_j, o_d, d_seg_out, d_total, h->stream, &nl)); h->launches += nl; }
    out->d_total = d_total; out->d_seg_out = d_seg_out;
    out->o_view = o_view; out->o_i = o_i; out->o_j = o_j; out->o_d = o_d;
    out->n_rows = n_rows;
    return HULO_OK;
}

extern "C" {

int hulo_match_to_query(hulo_gpu *h, const hulo_db *map, const uint32_t *views, size_t n_views,
                        const uint8_t *query, size_t nq, size_t q_stride, float ratio, uint32_t *out_view,
                        uint32_t *out_i, uint32_t *out_j, int32_t *out_d0, size_t cap, size_t *n_out,
                        uint32_t *view_counts) {
    HULO_ARG(n_out != nullptr, "null argument");
    *n_out = 0;
    QueryMatchesDev dm;
    int rc = match_to_query_device(h, map, views, n_views, query, nq, q_stride, ratio, &dm);
    if (rc != HULO_OK) return rc;
    n_views = dm.n_views;
    if (view_counts) memset(view_counts, 0, n_views * sizeof(uint32_t));
    if (dm.d_total == nullptr) return nq < 1 ? HULO_OK : hulo_synchronize(h);
    const uint64_t n_rows = dm.n_rows;
    const uint64_t *d_total = dm.d_total;
    const uint32_t *o_view = dm.o_view, *o_i = dm.o_i, *o_j = dm.o_j;
    const int32_t *o_d = dm.o_d;

    // One round trip in the common case: the totals and per-view offsets travel together with the
    // first `spec` survivors of every output array into pinned memory; only a query with more
    // survivors than that fetches the rest in a second round.
    const size_t spec = std::min<size_t>(std::min<size_t>((size_t)n_rows, cap), 16384);
    const size_t head_bytes = ((n_views + 2) * sizeof(uint64_t) + 15) & ~(size_t)15;
    HULO_CUDA(h->hstage0.reserve(head_bytes + 4 * spec * sizeof(uint32_t)));
    uint64_t *h_tot = h->hstage0.as<uint64_t>();
    uint32_t *h_spec = reinterpret_cast<uint32_t *>(h->hstage0.as<uint8_t>() + head_bytes);
    HULO_CUDA(cudaMemcpyAsync(h_tot, d_total, (n_views + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    if (spec > 0) {
        if (out_view) HULO_CUDA(cudaMemcpyAsync(h_spec, o_view, spec * 4, cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaMemcpyAsync(h_spec + spec, o_i, spec * 4, cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaMemcpyAsync(h_spec + 2 * spec, o_j, spec * 4, cudaMemcpyDeviceToHost, h->stream));
        if (out_d0) HULO_CUDA(cudaMemcpyAsync(h_spec + 3 * spec, o_d, spec * 4, cudaMemcpyDeviceToHost, h->stream));
    }
    HULO_CUDA(cudaStreamSynchronize(h->stream));
    const uint64_t total = h_tot[0];
    *n_out = (size_t)total;
    if (view_counts)
        for (size_t v = 0; v < n_views; ++v) view_counts[v] = (uint32_t)(h_tot[2 + v] - h_tot[1 + v]);
    if (total > cap) {
        set_error("hulo_match_to_query: %llu matches, capacity %zu", (unsigned long long)total, cap);
        return HULO_ERR_CAPACITY;
    }
    if (total > 0) {
        HULO_ARG(out_i != nullptr && out_j != nullptr, "null output");
        const size_t got = std::min<size_t>((size_t)total, spec);
        if (out_view) memcpy(out_view, h_spec, got * 4);
        memcpy(out_i, h_spec + spec, got * 4);
        memcpy(out_j, h_spec + 2 * spec, got * 4);
        if (out_d0) memcpy(out_d0, h_spec + 3 * spec, got * 4);
        if (total > got) {
            const size_t rest = (size_t)total - got;
            if (out_view) HULO_CUDA(cudaMemcpyAsync(out_view + got, o_view + got, rest * 4, cudaMemcpyDeviceToHost, h->stream));
            HULO_CUDA(cudaMemcpyAsync(out_i + got, o_i + got, rest * 4, cudaMemcpyDeviceToHost, h->stream));
            HULO_CUDA(cudaMemcpyAsync(out_j + got, o_j + got, rest * 4, cudaMemcpyDeviceToHost, h->stream));
            if (out_d0) HULO_CUDA(cudaMemcpyAsync(out_d0 + got, o_d + got, rest * 4, cudaMemcpyDeviceToHost, h->stream));
            HULO_CUDA(cudaStreamSynchronize(h->stream));
        }
    }
    return HULO_OK;
}

// Batched form: several query images against the same map in one pass (BASELINE.json config 4).
int hulo_match_to_queries(hulo_gpu *h, const hulo_db *map, const uint32_t *views, size_t n_views,
                          const uint8_t *queries, size_t q_stride, const uint64_t *q_offsets, size_t n_queries,
                          float ratio, uint32_t *out_query, uint32_t *out_view, uint32_t *out_i, uint32_t *out_j,
                          int32_t *out_d0, size_t cap, size_t *n_out, uint32_t *counts) {
    HULO_ARG(h != nullptr && map != nullptr && n_out != nullptr, "null argument");
    HULO_ARG(n_queries == 0 || q_offsets != nullptr, "q_offsets is null");
    HULO_ARG(q_stride >= 1, "stride must be >= 1");
    *n_out = 0;
    HULO_CUDA(cudaSetDevice(h->device));
    HULO_CUDA(join_exchange(h));
    const size_t n_seg = map->seg.size() - 1;
    if (views == nullptr) n_views = n_seg;
    for (size_t v = 0; views && v < n_views; ++v) HULO_ARG(views[v] < n_seg, "view index out of range");
    if (counts) memset(counts, 0, n_queries * n_views * sizeof(uint32_t));
    if (n_queries == 0 || n_views == 0) return HULO_OK;
    const uint64_t total_q_rows = q_offsets[n_queries];
    for (size_t q = 0; q < n_queries; ++q) {
        HULO_ARG(q_offsets[q] <= q_offsets[q + 1], "q_offsets not ascending");
        HULO_ARG(q_offsets[q + 1] - q_offsets[q] <= kMaxChunkRows, "query image with more than 4 Mi descriptors");
    }
    HULO_ARG(total_q_rows == 0 || queries != nullptr, "queries is null");
    HULO_ARG(total_q_rows < (uint64_t)INT_MAX, "too many query rows");

    std::vector<uint64_t> sel_off(n_views + 1, 0);
    for (size_t v = 0; v < n_views; ++v) {
        const size_t s = views ? views[v] : v;
        sel_off[v + 1] = sel_off[v] + (map->seg[s + 1] - map->seg[s]);
    }
    const uint64_t n_rows = sel_off[n_views];
    if (n_rows == 0 || total_q_rows == 0) return HULO_OK;

    // all query rows go up once
    HULO_CUDA(h->stageB.reserve(total_q_rows * HULO_ROW_BYTES));
    cudaError_t e;
    const uint8_t *q64 = stage_rows(h->hstage1, queries, total_q_rows, q_stride, &e);
    HULO_CUDA(e);
    HULO_CUDA(cudaMemcpyAsync(h->stageB.ptr, q64, total_q_rows * HULO_ROW_BYTES, cudaMemcpyHostToDevice, h->stream));
    HULO_CUDA(knn2_fold_rows_launch(h->stageB.as<uint4>(), total_q_rows, h->stream));
    h->launches++;

    // K1t: the map's segmented image as searcher tiles, one segmented image of all query images as
    // database ranges; per (view tile, query) one item, the items of a view tile next to each other
    const bool use_tc = tc_wanted(h, n_rows * total_q_rows, 1ull << 26);
    const uint8_t *imgA = nullptr, *imgB = nullptr;
    const uint32_t *tile0_map = nullptr;
    std::vector<uint32_t> tile0_q;
    std::vector<TcItem> tc_items;
    if (use_tc) {
        int rc = tc_seg_image_for_db(h, map, &imgA, &tile0_map);
        if (rc != HULO_OK) return rc;
        rc = tc_items_four(h) ? tc_build_seg_image4(h, h->stageB.as<uint4>(), q_offsets, n_queries, h->tc_scratchB, tile0_q)
                             : tc_build_seg_image(h, h->stageB.as<uint4>(), q_offsets, n_queries, h->tc_scratchB, tile0_q);
        if (rc != HULO_OK) return rc;
        imgB = h->tc_scratchB.as<uint8_t>();
    }

    const KnnConfig cfg = choose_config(h, (size_t)n_rows);
    const uint32_t tile = knn2_tile_rows(cfg);
    struct Run { uint64_t a0, rows, slot0; };
    std::vector<Run> runs;
    for (size_t v = 0; v < n_views; ++v) {
        const size_t s = views ? views[v] : v;
        const uint64_t a0 = map->seg[s], rows = map->seg[s + 1] - map->seg[s];
        if (rows == 0) continue;
        if (!runs.empty() && runs.back().a0 + runs.back().rows == a0) runs.back().rows += rows;
        else runs.push_back(Run{a0, rows, sel_off[v]});
    }
    int ctas_per_sm = 1;
    knn2_kernel_info(cfg, nullptr, &ctas_per_sm, nullptr);
    const int full_grid = h->sm_count * std::max(1, ctas_per_sm);

    // sub-batches of queries bounded by compact rows (queries x selected map rows)
    const uint64_t budget = (uint64_t)std::max(1, env_int("HULO_QUERY_BATCH_ROWS", 48 << 20));
    HULO_ARG(n_rows <= budget, "selected map rows exceed HULO_QUERY_BATCH_ROWS");
    const size_t per_batch = (size_t)std::max<uint64_t>(1, budget / n_rows);
    std::vector<KnnItem> items;
    std::vector<uint64_t> seg_off, h_seg_out;
    std::vector<size_t> members;
    size_t total_out = 0;
    bool overflow = false;
    for (size_t q0 = 0; q0 < n_queries; q0 += per_batch) {
        const size_t q1 = std::min(n_queries, q0 + per_batch);
        members.clear();
        for (size_t q = q0; q < q1; ++q)
            if (q_offsets[q + 1] > q_offsets[q]) members.push_back(q);     // MatchUtils.cpp:299-301
        if (members.empty()) continue;
        const size_t nb = members.size();
        const uint64_t rows_total = (uint64_t)nb * n_rows;
        items.clear();
        tc_items.clear();
        seg_off.assign(nb * n_views + 1, 0);
        if (use_tc) {
            for (size_t v = 0; v < n_views; ++v) {
                const size_t s = views ? views[v] : v;
                const uint64_t rows = map->seg[s + 1] - map->seg[s];
                for (uint64_t t0 = 0; t0 < rows; t0 += kTcTileRows)
                    for (size_t ql = 0; ql < nb; ++ql) {
                        const size_t q = members[ql];
                        TcItem it{};
                        it.a_tile = tc_unit_of_row(h, tile0_map, s, t0);
                        it.a_rows = (uint32_t)std::min<uint64_t>(kTcTileRows, rows - t0);
                        it.b_tile0 = tile0_q[q];
                        it.b_rows = (uint32_t)(q_offsets[q + 1] - q_offsets[q]);
                        it.out_slot0 = (uint64_t)ql * n_rows + sel_off[v] + t0;
                        tc_items.push_back(it);
                    }
            }
        }
        for (size_t ql = 0; ql < nb; ++ql) {
            const size_t q = members[ql];
            if (!use_tc)
            for (const Run &r : runs)
                for (uint64_t t0 = 0; t0 < r.rows; t0 += tile) {
                    KnnItem it{};
                    it.a_row0 = (uint32_t)(r.a0 + t0);
                    it.a_rows = (uint32_t)std::min<uint64_t>(tile, r.rows - t0);
                    it.b_row0 = (uint32_t)q_offsets[q];
                    it.b_rows = (uint32_t)(q_offsets[q + 1] - q_offsets[q]);
                    it.out_slot0 = (uint64_t)ql * n_rows + r.slot0 + t0;
                    items.push_back(it);
                }
            for (size_t v = 0; v < n_views; ++v) seg_off[ql * n_views + v] = (uint64_t)ql * n_rows + sel_off[v];
        }
        seg_off[nb * n_views] = rows_total;
        const size_t n_segs = nb * n_views;
        const size_t n_blocks = (rows_total + kCompactBlockRows - 1) / kCompactBlockRows;
        HULO_CUDA(h->partial.reserve(rows_total * sizeof(uint2)));
        if (!use_tc) {
        HULO_CUDA(h->items.reserve(items.size() * sizeof(KnnItem)));
        HULO_CUDA(cudaMemcpyAsync(h->items.ptr, items.data(), items.size() * sizeof(KnnItem), cudaMemcpyHostToDevice, h->stream));
        HULO_CUDA(h->counter.reserve(sizeof(unsigned int)));
        HULO_CUDA(cudaMemsetAsync(h->counter.ptr, 0, sizeof(unsigned int), h->stream));
        }
        HULO_CUDA(h->scratch0.reserve(rows_total * 2 * sizeof(int32_t)));
        HULO_CUDA(h->scratch1.reserve((n_segs + 1) * sizeof(uint64_t)));
        HULO_CUDA(h->scratch2.reserve((n_blocks + 2) * sizeof(uint32_t) + (n_segs + 2) * sizeof(uint64_t) + 64));
        HULO_CUDA(h->scratch3.reserve(rows_total * 3 * sizeof(uint32_t)));
        HULO_CUDA(cudaMemcpyAsync(h->scratch1.ptr, seg_off.data(), (n_segs + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
        if (use_tc) {
            int rc = run_items_tc(h, imgA, imgB, tc_items);
            if (rc != HULO_OK) return rc;
        } else {
        KnnParams kp{};
        kp.A = map->rows; kp.B = h->stageB.as<uint4>();
        kp.items = h->items.as<KnnItem>();
        kp.n_items = (uint32_t)items.size();
        kp.key_unit = 1u << kKeyIdxBits;
        kp.partial = h->partial.as<uint2>();
        kp.counter = h->counter.as<unsigned int>();
        HULO_CUDA(knn2_launch(kp, cfg, (int)std::min<size_t>((size_t)full_grid, items.size()), h->stream));
        h->launches++;
        }
        int32_t *val = h->scratch0.as<int32_t>();
        int32_t *dist = val + rows_total;
        HULO_CUDA(post_query_launch(h->partial.as<uint2>(), (uint32_t)rows_total, 1, rows_total, kMaxChunkRows, ratio, val,
                                    dist, h->stream));
        h->launches++;
        uint8_t *s2 = h->scratch2.as<uint8_t>();
        uint64_t *d_total = reinterpret_cast<uint64_t *>(s2);
        uint64_t *d_seg_out = d_total + 1;
        uint32_t *d_blocks = reinterpret_cast<uint32_t *>(d_seg_out + (n_segs + 1));
        uint32_t *o_i = h->scratch3.as<uint32_t>(), *o_j = o_i + rows_total;
        int32_t *o_d = reinterpret_cast<int32_t *>(o_j + rows_total);
        { int nl = 0; HULO_CUDA(compact_launch(val, dist, (uint32_t)rows_total, h->scratch1.as<uint64_t>(), (uint32_t)n_segs, d_blocks,
                                 nullptr, o_i, o_j, o_d, d_seg_out, d_total, h->stream, &nl)); h->launches += nl; }
        h_seg_out.resize(n_segs + 2);
        HULO_CUDA(cudaMemcpyAsync(h_seg_out.data(), d_total, (n_segs + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
        HULO_CUDA(cudaStreamSynchronize(h->stream));
        const uint64_t total = h_seg_out[0];
        if (!overflow && total_out + total <= cap) {
            if (total > 0) {
                HULO_ARG(out_i != nullptr && out_j != nullptr, "null output");
                HULO_CUDA(cudaMemcpyAsync(out_i + total_out, o_i, total * 4, cudaMemcpyDeviceToHost, h->stream));
                HULO_CUDA(cudaMemcpyAsync(out_j + total_out, o_j, total * 4, cudaMemcpyDeviceToHost, h->stream));
                if (out_d0) HULO_CUDA(cudaMemcpyAsync(out_d0 + total_out, o_d, total * 4, cudaMemcpyDeviceToHost, h->stream));
                HULO_CUDA(cudaStreamSynchronize(h->stream));
            }
            size_t k = total_out;
            for (size_t ql = 0; ql < nb; ++ql)
                for (size_t v = 0; v < n_views; ++v) {
                    const uint32_t c = (uint32_t)(h_seg_out[2 + ql * n_views + v] - h_seg_out[1 + ql * n_views + v]);
                    if (counts) counts[members[ql] * n_views + v] = c;
                    for (uint32_t m = 0; m < c; ++m, ++k) {
                        if (out_query) out_query[k] = (uint32_t)members[ql];
                        if (out_view) out_view[k] = (uint32_t)v;
                    }
                }
        } else {
            overflow = true;
        }
        total_out += total;
    }
    *n_out = total_out;
    if (overflow) {
        set_error("hulo_match_to_queries: %zu matches, capacity %zu", total_out, cap);
        return HULO_ERR_CAPACITY;
    }
    return HULO_OK;
}

// ------------------------------------------------ image-pair matching (reconstruction)
int hulo_match_pairs(hulo_gpu *h, const hulo_db *db, const uint32_t *pairs, size_t n_pairs, float ratio,
                     unsigned flags, uint64_t *pair_offsets, uint32_t *out_i, uint32_t *out_j, size_t cap,
                     size_t *n_out) {
    HULO_ARG(h != nullptr && db != nullptr && n_out != nullptr, "null argument");
    HULO_ARG(n_pairs == 0 || pairs != nullptr, "pairs is null");
    *n_out = 0;
    HULO_CUDA(cudaSetDevice(h->device));
    HULO_CUDA(join_exchange(h));
    const size_t n_seg = db->seg.size() - 1;
    for (size_t p = 0; p < n_pairs; ++p)
        HULO_ARG(pairs[2 * p] < n_seg && pairs[2 * p + 1] < n_seg, "pair refers to a segment that does not exist");
    if (pair_offsets) pair_offsets[0] = 0;

    const KnnConfig cfg = h->knn_cfg_forced ? h->knn_cfg : mid_config();   // 1024-row tiles fit 5000-row images with 2 % waste
    const uint32_t tile = knn2_tile_rows(cfg);
    int ctas_per_sm = 1;
    knn2_kernel_info(cfg, nullptr, &ctas_per_sm, nullptr);
    const int full_grid = h->sm_count * std::max(1, ctas_per_sm);

    // K1t: both sides of a pair out of the table's segmented image
    uint64_t all_dist = 0;
    for (size_t p = 0; p < n_pairs; ++p)
        all_dist += (db->seg[pairs[2 * p] + 1] - db->seg[pairs[2 * p]]) * (db->seg[pairs[2 * p + 1] + 1] - db->seg[pairs[2 * p + 1]]);
    const bool use_tc = tc_wanted(h, all_dist, 1ull << 26);
    const uint8_t *img = nullptr;
    const uint32_t *tile0 = nullptr;
    std::vector<TcItem> tc_items;
    if (use_tc) {
        int rc = tc_seg_image_for_db(h, db, &img, &tile0);
        if (rc != HULO_OK) return rc;
    }

    // batches bounded by compact searcher rows so the scratch stays modest
    const uint64_t max_batch_rows = (uint64_t)std::max(1, env_int("HULO_PAIR_BATCH_ROWS", 16 << 20));
    size_t total_out = 0;
    bool overflow = false;
    size_t p0 = 0;
    std::vector<uint64_t> row_off, hist_off, h_seg_out;
    std::vector<KnnItem> items;
    while (p0 < n_pairs) {
        row_off.assign(1, 0);
        hist_off.assign(1, 0);
        items.clear();
        tc_items.clear();
        size_t p1 = p0;
        while (p1 < n_pairs) {
            const uint32_t I = pairs[2 * p1], J = pairs[2 * p1 + 1];
            const uint64_t nI = db->seg[I + 1] - db->seg[I], nJ = db->seg[J + 1] - db->seg[J];
            // MatchUtils.cpp:99-101: pairs with fewer than two rows on either side are skipped
            const bool skip = nI < 2 || nJ < 2;
            HULO_ARG(skip || nJ <= kMaxChunkRows, "image with more than 4 Mi descriptors");
            const uint64_t add = skip ? 0 : nI;
            if (p1 > p0 && row_off.back() + add > max_batch_rows) break;
            if (!skip && use_tc) {
                for (uint64_t t0 = 0; t0 < nI; t0 += kTcTileRows) {
                    TcItem it{};
                    it.a_tile = tc_unit_of_row(h, tile0, I, t0);
                    it.a_rows = (uint32_t)std::min<uint64_t>(kTcTileRows, nI - t0);
                    it.b_tile0 = tile0[J];
                    it.b_rows = (uint32_t)nJ;
                    it.out_slot0 = row_off.back() + t0;
                    tc_items.push_back(it);
                }
            } else if (!skip) {
                for (uint64_t t0 = 0; t0 < nI; t0 += tile) {
                    KnnItem it{};
                    it.a_row0 = (uint32_t)(db->seg[I] + t0);
                    it.a_rows = (uint32_t)std::min<uint64_t>(tile, nI - t0);
                    it.b_row0 = (uint32_t)db->seg[J];
                    it.b_rows = (uint32_t)nJ;
                    it.out_slot0 = row_off.back() + t0;
                    items.push_back(it);
                }
            }
            row_off.push_back(row_off.back() + add);
            hist_off.push_back(hist_off.back() + (skip ? 0 : nJ));
            ++p1;
        }
        const size_t bp = p1 - p0;
        const uint64_t n_rows = row_off.back();
        HULO_ARG(n_rows < (uint64_t)INT_MAX, "batch too large");
        if (n_rows > 0) {
            const size_t n_blocks = (n_rows + kCompactBlockRows - 1) / kCompactBlockRows;
            HULO_CUDA(h->partial.reserve(n_rows * sizeof(uint2)));
            if (!use_tc) {
                HULO_CUDA(h->items.reserve(items.size() * sizeof(KnnItem)));
                HULO_CUDA(cudaMemcpyAsync(h->items.ptr, items.data(), items.size() * sizeof(KnnItem), cudaMemcpyHostToDevice, h->stream));
                HULO_CUDA(h->counter.reserve(sizeof(unsigned int)));
                HULO_CUDA(cudaMemsetAsync(h->counter.ptr, 0, sizeof(unsigned int), h->stream));
            }
            // scratch0: val ; scratch1: row_off | hist_off (device) ; scratch2: total, seg_out_off, block counts
            // scratch3: out_i | out_j ; gathered (reused): claim histogram
            HULO_CUDA(h->scratch0.reserve(n_rows * sizeof(int32_t)));
            HULO_CUDA(h->scratch1.reserve(2 * (bp + 1) * sizeof(uint64_t)));
            HULO_CUDA(h->scratch2.reserve((n_blocks + 2) * sizeof(uint32_t) + (bp + 2) * sizeof(uint64_t) + 64));
            HULO_CUDA(h->scratch3.reserve(n_rows * 2 * sizeof(uint32_t)));
            HULO_CUDA(h->gathered.reserve(std::max<uint64_t>(hist_off.back(), 1) * sizeof(uint32_t)));
            uint64_t *d_row_off = h->scratch1.as<uint64_t>();
            uint64_t *d_hist_off = d_row_off + (bp + 1);
            HULO_CUDA(cudaMemcpyAsync(d_row_off, row_off.data(), (bp + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemcpyAsync(d_hist_off, hist_off.data(), (bp + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, h->stream));
            HULO_CUDA(cudaMemsetAsync(h->gathered.ptr, 0, std::max<uint64_t>(hist_off.back(), 1) * sizeof(uint32_t), h->stream));

            if (use_tc) {
                int rc = run_items_tc(h, img, img, tc_items);
                if (rc != HULO_OK) return rc;
            } else {
                KnnParams kp{};
                kp.A = db->rows; kp.B = db->rows;
                kp.items = h->items.as<KnnItem>();
                kp.n_items = (uint32_t)items.size();
                kp.key_unit = 1u << kKeyIdxBits;
                kp.partial = h->partial.as<uint2>();
                kp.counter = h->counter.as<unsigned int>();
                HULO_CUDA(knn2_launch(kp, cfg, (int)std::min<size_t>((size_t)full_grid, items.size()), h->stream));
                h->launches++;
            }

            int32_t *val = h->scratch0.as<int32_t>();
            uint32_t *hist = h->gathered.as<uint32_t>();
            HULO_CUDA(post_pair_claim_launch(h->partial.as<uint2>(), (uint32_t)n_rows, d_row_off, d_hist_off,
                                             (uint32_t)bp, ratio, val, hist, h->stream));
            h->launches++;
            if (flags & (HULO_PAIR_ONE_TO_ONE | HULO_PAIR_DROP_LAST)) {
                HULO_CUDA(post_pair_filter_launch((uint32_t)n_rows, d_row_off, d_hist_off, (uint32_t)bp, flags, val,
                                                  hist, h->stream));
                h->launches++;
            }
            uint8_t *s2 = h->scratch2.as<uint8_t>();
            uint64_t *d_total = reinterpret_cast<uint64_t *>(s2);
            uint64_t *d_seg_out = d_total + 1;
            uint32_t *d_blocks = reinterpret_cast<uint32_t *>(d_seg_out + (bp + 1));
            uint32_t *o_i = h->scratch3.as<uint32_t>(), *o_j = o_i + n_rows;
            { int nl = 0; HULO_CUDA(compact_launch(val, nullptr, (uint32_t)n_rows, d_row_off, (uint32_t)bp, d_blocks, nullptr, o_i,
                                     o_j, nullptr, d_seg_out, d_total, h->stream, &nl)); h->launches += nl; }

            h_seg_out.resize(bp + 2);
            HULO_CUDA(cudaMemcpyAsync(h_seg_out.data(), d_total, (bp + 2) * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
            HULO_CUDA(cudaStreamSynchronize(h->stream));
            const uint64_t total = h_seg_out[0];
            if (pair_offsets)
                for (size_t p = 0; p < bp; ++p) pair_offsets[p0 + p + 1] = total_out + h_seg_out[2 + p];
            if (!overflow && total_out + total <= cap) {
                if (total > 0) {
                    HULO_ARG(out_i != nullptr && out_j != nullptr, "null output");
                    HULO_CUDA(cudaMemcpyAsync(out_i + total_out, o_i, total * 4, cudaMemcpyDeviceToHost, h->stream));
                    HULO_CUDA(cudaMemcpyAsync(out_j + total_out, o_j, total * 4, cudaMemcpyDeviceToHost, h->stream));
                    HULO_CUDA(cudaStreamSynchronize(h->stream));
                }
            } else {
                overflow = true;
            }
            total_out += total;
        } else if (pair_offsets) {
            for (size_t p = 0; p < bp; ++p) pair_offsets[p0 + p + 1] = total_out;
        }
        p0 = p1;
    }
    *n_out = total_out;
    if (overflow) {
        set_error("hulo_match_pairs: %zu matches, capacity %zu", total_out, cap);
        return HULO_ERR_CAPACITY;
    }
    return HULO_OK;
}

}  // extern "C"
