// knn2_tc.cu -- K1t kernels (see knn2_tc.cuh for the design).
#include "knn2_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>

namespace hulo {

namespace {

using namespace tcptx;

constexpr int kStages = 4;                       // ring of B stages (256 rows x 128 K-bytes = 32 KB each)
constexpr uint32_t kStageBytes = 32768;
constexpr uint32_t kKChunks = 4;                 // 512 K-bytes = 4 stages per accumulator tile
constexpr uint32_t kABytes = 65536;              // resident searcher tile
constexpr uint32_t kThreads = 320;               // warp 0: copies, warp 1: MMA issue, warps 2..5 / 6..9: epilogue of even / odd tiles
constexpr uint32_t kTmemCols = 512;              // two 128 x 256 int32 accumulators
constexpr size_t kSmemBytes = 1024 + kABytes + (size_t)kStages * kStageBytes + 256 + 2 * 128 * sizeof(uint2);

constexpr int32_t kThrNone = -1024;              // "second best" threshold while fewer than two candidates
constexpr int32_t kDotMasked = -2048;            // accumulator value given to columns past the chunk end

// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M = 128, N = 256, K = 32
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor): D = s32, A = B = signed 8 bit,
// both K-major, N = 256, M = 128
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((kTcTileN >> 3) << 17) | ((128u >> 4) << 24);

// 32 lanes x 64 consecutive 32-bit columns of TMEM -> 64 registers per thread (SASS: LDTM)
#define HULO_LDTM64(v, taddr)                                                                                       \
    asm volatile(                                                                                                   \
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "                                                                   \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                  \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "                          \
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "                          \
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"                   \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),          \
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),    \
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),  \
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),  \
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),  \
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),  \
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),  \
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])   \
        : "r"(taddr)                                                                                                \
        : "memory")

// tcgen05.wait::ld with the 64 destination registers as read-write operands, so the compiler cannot
// move a use of them above the wait
#define HULO_WAIT_LD64(v) \
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]), "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]), "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31]), "+r"(v[32]), "+r"(v[33]), "+r"(v[34]), "+r"(v[35]), "+r"(v[36]), "+r"(v[37]), "+r"(v[38]), "+r"(v[39]), "+r"(v[40]), "+r"(v[41]), "+r"(v[42]), "+r"(v[43]), "+r"(v[44]), "+r"(v[45]), "+r"(v[46]), "+r"(v[47]), "+r"(v[48]), "+r"(v[49]), "+r"(v[50]), "+r"(v[51]), "+r"(v[52]), "+r"(v[53]), "+r"(v[54]), "+r"(v[55]), "+r"(v[56]), "+r"(v[57]), "+r"(v[58]), "+r"(v[59]), "+r"(v[60]), "+r"(v[61]), "+r"(v[62]), "+r"(v[63]) : : "memory")

__device__ __forceinline__ int32_t max3(int32_t a, int32_t b, int32_t c) { return max(max(a, b), c); }

// One block of 64 accumulator columns of this thread's searcher row.  v[e] = 512 - 2 * distance to
// database row (row0 + e) of the chunk.  Fast path per 16 columns: the largest dot against the dot
// of the current second best (thr); columns arrive in ascending row order, so a column that only
// TIES the second best loses on the index and a strict compare is exact.  Slow path: the packed
// keys of the 16 columns through the same min/max update as K1.
__device__ __forceinline__ void scan_block(const int32_t (&v)[64], uint32_t row0, uint32_t &best0, uint32_t &best1,
                                           int32_t &thr) {
    // the four group maxima first: 32 independent min/max instructions the scheduler can overlap
    // (the compares below depend on each other through thr)
    int32_t gm[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int32_t *w = &v[16 * g];
        const int32_t m0 = max3(w[0], w[1], w[2]), m1 = max3(w[3], w[4], w[5]), m2 = max3(w[6], w[7], w[8]);
        const int32_t m3 = max3(w[9], w[10], w[11]), m4 = max3(w[12], w[13], w[14]);
        gm[g] = max(max3(m0, m1, m2), max3(m3, m4, w[15]));
    }
    if (max(max(gm[0], gm[1]), max(gm[2], gm[3])) <= thr) return;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const int32_t *w = &v[16 * g];
        if (gm[g] > thr) {
            const int32_t t0 = thr;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                if (w[e] > t0) {
                    const uint32_t key = ((uint32_t)(512 - w[e]) << (kKeyIdxBits - 1)) + (row0 + 16 * g + e);
                    const uint32_t hi = max(best0, key);
                    best0 = min(best0, key);
                    best1 = min(best1, hi);
                }
            }
            thr = best1 == kKeyNone ? kThrNone : 512 - 2 * (int32_t)(best1 >> kKeyIdxBits);
        }
    }
}

// One unit of work as every warp role sees it.
struct TcWork {
    uint32_t a_tile;     // image tile of the searcher rows
    uint32_t a_rows;     // rows of it whose keys are written (0: none)
    uint32_t b_tile0;    // first image tile (128 rows) of the database range, even
    uint32_t b_rows;     // database rows scanned
    uint64_t out_slot0;  // partial[out_slot0 + r] receives the keys of searcher row r of the tile
};
// Flat mode: item w -> group of CL consecutive searcher tiles (w % n_mgroups), chunk (w / n_mgroups).
// Item mode: the list entry.
template <int CL>
__device__ __forceinline__ TcWork tc_work(const TcParams &p, uint32_t w, uint32_t n_mgroups, uint32_t crank) {
    TcWork k;
    if (p.items != nullptr) {
        const TcItem it = p.items[w];
        k.a_tile = it.a_tile; k.a_rows = it.a_rows; k.b_tile0 = it.b_tile0; k.b_rows = it.b_rows;
        k.out_slot0 = it.out_slot0;
    } else {
        const uint32_t mt = (w % n_mgroups) * CL + crank, c = w / n_mgroups;
        const uint32_t b_row0 = c * p.rows_per_chunk;
        // a cluster past the last searcher tile repeats it (its results are not written)
        k.a_tile = min(mt, p.n_mtiles - 1u);
        k.a_rows = mt < p.n_mtiles ? min(kTcTileRows, p.nA - mt * kTcTileRows) : 0u;
        k.b_tile0 = b_row0 / kTcTileRows;
        k.b_rows = min(p.rows_per_chunk, p.nB - b_row0);
        k.out_slot0 = (uint64_t)c * p.slot_stride + (uint64_t)mt * kTcTileRows;
    }
    return k;
}

// CL = CTAs per cluster.  The CTAs of a cluster work on CL consecutive searcher tiles against the
// SAME database chunk in lockstep: every stage of B is fetched once per cluster, each CTA issuing
// 1/CL of it as a multicast bulk copy into all CL shared memories, which divides the L2 -> SM
// traffic (the first limit of the single-CTA form) by CL.  A stage is refilled when the MMAs of all
// CL CTAs have released it (multicast tcgen05.commit onto every CTA's "empty" barrier).
template <int CL>
__global__ void __launch_bounds__(kThreads, 1) knn2_tc_kernel(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base;
    const uint32_t sB = smem_base + kABytes;
    const uint32_t bars = sB + kStages * kStageBytes;
    const uint32_t bar_a_full = bars, bar_a_empty = bars + 8;
    const uint32_t bar_b_full = bars + 16, bar_b_empty = bar_b_full + 8 * kStages;
    const uint32_t bar_acc_full = bar_b_empty + 8 * kStages, bar_acc_empty = bar_acc_full + 16;
    const uint32_t tmem_slot = bar_acc_empty + 16;
    const uint32_t xchg = bars + 256;                    // [2 item parities][128 rows] uint2: keys of the odd-tile group
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));
    uint2 *xchg_gen = reinterpret_cast<uint2 *>(smem_gen + (xchg - smem_base));

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_empty, 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_b_full + 8 * s, 1);
            mbar_init(bar_b_empty + 8 * s, CL);           // one commit per CTA of the cluster
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + 8 * b, 1);
            mbar_init(bar_acc_empty + 8 * b, 4);          // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();             // peers' barriers exist before anything is sent to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
    const uint32_t cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
    const uint32_t n_mgroups = (p.n_mtiles + CL - 1) / CL;
    constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
    // flat mode: items dealt round-robin over the clusters (the CTAs running at any moment share a few
    // database chunks through L2); item mode: a contiguous block of the list per CTA, so that items
    // with the same searcher tile follow each other and the tile stays in shared memory
    uint32_t w_begin, w_end, w_step;
    if (p.items != nullptr) {
        w_begin = (uint32_t)((uint64_t)blockIdx.x * p.n_items / gridDim.x);
        w_end = (uint32_t)((uint64_t)(blockIdx.x + 1) * p.n_items / gridDim.x);
        w_step = 1;
    } else {
        w_begin = cluster_id; w_end = n_mgroups * p.n_chunks; w_step = n_clusters;
    }

    if (warp == 0) {
        // ===== producer: searcher tile once per item, database stages through the ring =====
        if (lane == 0) {
            uint32_t stage = 0, ph = 0, a_loaded = 0xFFFFFFFFu, a_loads = 0;
            for (uint32_t w = w_begin; w < w_end; w += w_step) {
                const TcWork k = tc_work<CL>(p, w, n_mgroups, crank);
                const uint32_t n_tiles = (k.b_rows + kTcTileN - 1) / kTcTileN;
                if (k.a_tile != a_loaded) {
                    // a new searcher tile: the MMAs that read the previous one must be done (the MMA
                    // thread commits "A empty" when IT reaches this item; both sides count tile loads)
                    if (a_loads > 0) mbar_wait(bar_a_empty, (a_loads - 1u) & 1u);
                    mbar_expect_tx(bar_a_full, kABytes);
                    bulk_load(sA, p.imgA + (size_t)k.a_tile * kTcTileBytes, kABytes, bar_a_full);
                    a_loaded = k.a_tile;
                    ++a_loads;
                }
                const uint8_t *src = p.imgB + (size_t)k.b_tile0 * kTcTileBytes;
                for (uint32_t t = 0; t < n_tiles; ++t) {
                    for (uint32_t kc = 0; kc < kKChunks; ++kc) {
                        mbar_wait(bar_b_empty + 8 * stage, ph ^ 1u);
                        mbar_expect_tx(bar_b_full + 8 * stage, kStageBytes);
                        const uint8_t *s0 = src + (size_t)(2 * t) * kTcTileBytes + kc * 16384u;
                        if constexpr (CL == 1) {
                            bulk_load(sB + stage * kStageBytes, s0, 16384u, bar_b_full + 8 * stage);
                            bulk_load(sB + stage * kStageBytes + 16384u, s0 + kTcTileBytes, 16384u, bar_b_full + 8 * stage);
                        } else {
                            // this CTA's share of the stage (32 KB / CL), to every CTA of the cluster
                            constexpr uint32_t kShare = kStageBytes / CL;
                            const uint32_t off = crank * kShare;              // offset inside the 32 KB stage
                            const uint8_t *sp = s0 + (size_t)(off >> 14) * kTcTileBytes + (off & 16383u);
                            bulk_load_multicast(sB + stage * kStageBytes + off, sp, kShare, bar_b_full + 8 * stage, kMask);
                        }
                        if (++stage == kStages) { stage = 0; ph ^= 1u; }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===== MMA issuer: one thread =====
        if (lane == 0) {
            uint32_t stage = 0, ph = 0, acc_it = 0, a_loaded = 0xFFFFFFFFu, a_loads = 0;
            for (uint32_t w = w_begin; w < w_end; w += w_step) {
                const TcWork k = tc_work<CL>(p, w, n_mgroups, crank);
                const uint32_t n_tiles = (k.b_rows + kTcTileN - 1) / kTcTileN;
                if (k.a_tile != a_loaded) {
                    if (a_loads > 0) tc_commit(bar_a_empty);           // arrives when every MMA on the old tile is done
                    mbar_wait(bar_a_full, a_loads & 1u);
                    a_loaded = k.a_tile;
                    ++a_loads;
                }
                for (uint32_t t = 0; t < n_tiles; ++t, ++acc_it) {
                    const uint32_t buf = acc_it & 1u;
                    mbar_wait(bar_acc_empty + 8 * buf, ((acc_it >> 1) & 1u) ^ 1u);   // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * kTcTileN;
                    for (uint32_t kc = 0; kc < kKChunks; ++kc) {
                        mbar_wait(bar_b_full + 8 * stage, ph);
                        tc_fence_after();
#pragma unroll
                        for (uint32_t j = 0; j < 4; ++j) {
                            const uint64_t da = smem_desc(sA + kc * 16384u + j * 256u, p.lbo, p.sbo);
                            const uint64_t db = smem_desc(sB + stage * kStageBytes + j * 256u, p.lbo, p.sbo);
                            tc_mma_i8(tmem_d, da, db, kIdesc, (kc | j) != 0u);
                        }
                        // stage free once these MMAs have read it (in every CTA of the cluster)
                        if constexpr (CL == 1) tc_commit(bar_b_empty + 8 * stage);
                        else tc_commit_multicast(bar_b_empty + 8 * stage, kMask);
                        if (++stage == kStages) { stage = 0; ph ^= 1u; }
                    }
                    tc_commit(bar_acc_full + 8 * buf);
                }
            }
        }
        __syncwarp();
    } else {
        // ===== epilogue: warp q of a warpgroup reads TMEM lanes 32q .. 32q+31.  Two groups of four
        // warps: group 0 (warps 2..5) drains accumulator 0, group 1 (warps 6..9) accumulator 1, i.e.
        // they take alternate tiles, so each has two MMA tile times per tile.  Each thread keeps the
        // best two of ITS tiles; the groups are merged per item through shared memory. =====
        const uint32_t quarter = warp & 3u;
        const uint32_t grp = (warp - 2u) >> 2;                   // accumulator buffer this warp drains
        const uint32_t row = quarter * 32u + lane;               // searcher row within the tile
        uint32_t acc_base = 0, uses = 0, it = 0;                 // tiles before this item; uses of this group's buffer
        for (uint32_t w = w_begin; w < w_end; w += w_step, ++it) {
            const TcWork k = tc_work<CL>(p, w, n_mgroups, crank);
            const uint32_t b_rows = k.b_rows;
            const uint32_t n_tiles = (b_rows + kTcTileN - 1) / kTcTileN;
            uint32_t best0 = kKeyNone, best1 = kKeyNone;
            int32_t thr = kThrNone;
            const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + grp * kTcTileN;
            for (uint32_t t = (acc_base + grp) & 1u; t < n_tiles; t += 2, ++uses) {
                mbar_wait(bar_acc_full + 8 * grp, uses & 1u);
                tc_fence_after();
                const uint32_t n_valid = min(kTcTileN, b_rows - t * kTcTileN);
                int32_t va[64], vb[64];
                HULO_LDTM64(va, taddr);
                HULO_WAIT_LD64(va);
                const bool dump = p.dbg_dots != nullptr && w == 0 && t == 0 && crank == 0 && p.items == nullptr;
                if (dump) _Pragma("unroll") for (int e = 0; e < 64; ++e) p.dbg_dots[row * 256 + e] = va[e];
                HULO_LDTM64(vb, taddr + 64u);
                if (n_valid < 64u) {
#pragma unroll
                    for (int e = 0; e < 64; ++e) if ((uint32_t)e >= n_valid) va[e] = kDotMasked;
                }
                scan_block(va, t * kTcTileN, best0, best1, thr);
                HULO_WAIT_LD64(vb);
                if (dump) _Pragma("unroll") for (int e = 0; e < 64; ++e) p.dbg_dots[row * 256 + 64 + e] = vb[e];
                HULO_LDTM64(va, taddr + 128u);
                if (n_valid < 128u) {
#pragma unroll
                    for (int e = 0; e < 64; ++e) if ((uint32_t)(64 + e) >= n_valid) vb[e] = kDotMasked;
                }
                scan_block(vb, t * kTcTileN + 64u, best0, best1, thr);
                HULO_WAIT_LD64(va);
                if (dump) _Pragma("unroll") for (int e = 0; e < 64; ++e) p.dbg_dots[row * 256 + 128 + e] = va[e];
                HULO_LDTM64(vb, taddr + 192u);
                if (n_valid < 192u) {
#pragma unroll
                    for (int e = 0; e < 64; ++e) if ((uint32_t)(128 + e) >= n_valid) va[e] = kDotMasked;
                }
                scan_block(va, t * kTcTileN + 128u, best0, best1, thr);
                HULO_WAIT_LD64(vb);
                if (dump) _Pragma("unroll") for (int e = 0; e < 64; ++e) p.dbg_dots[row * 256 + 192 + e] = vb[e];
                // every column of this accumulator is in registers: hand it back before the last scan
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_acc_empty + 8 * grp);
                if (n_valid < 256u) {
#pragma unroll
                    for (int e = 0; e < 64; ++e) if ((uint32_t)(192 + e) >= n_valid) vb[e] = kDotMasked;
                }
                scan_block(vb, t * kTcTileN + 192u, best0, best1, thr);
            }
            acc_base += n_tiles;
            // merge the two groups: keys are unique, so min/max on them is the (distance, index) order
            uint2 *slot = xchg_gen + (it & 1u) * 128u + row;
            if (grp == 1u) *slot = make_uint2(best0, best1);
            asm volatile("bar.sync 1, 256;" ::: "memory");        // the eight epilogue warps
            if (grp == 0u) {
                const uint2 o = *slot;
                const uint32_t lo = min(best0, o.x), mid = max(best0, o.x);
                const uint32_t second = min(mid, min(best1, o.y));
                if (row < k.a_rows) p.partial[k.out_slot0 + row] = make_uint2(lo, second);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (CL > 1) cluster_sync_all();             // no peer may still send into this CTA's shared memory
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------ tile image
// The 16 bytes of the image holding bits [16 j, 16 j + 16) of a folded row (j = kbyte / 2).
__device__ __forceinline__ uint4 expand_piece(const uint32_t *f, uint32_t kbyte) {
    const uint32_t wi = kbyte >> 2;
    uint32_t wv = __ldg(f + wi);
    // undo the K1 fold (knn2.cuh): words 2, 5, 8, 11, 14 hold the XOR of their triple, word 15 of w9..w15
    if (wi == 15u) wv ^= __ldg(f + 11) ^ __ldg(f + 14);
    else if (wi % 3u == 2u) wv ^= __ldg(f + wi - 1) ^ __ldg(f + wi - 2);
    const uint32_t bits = (wv >> ((kbyte & 2u) * 8u)) & 0xFFFFu;
    uint32_t o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t x = (bits >> (4 * q)) & 15u;
        const uint32_t ones = (x * 0x00204081u) & 0x01010101u;            // bit b -> byte b
        o[q] = ones | ((ones ^ 0x01010101u) * 0xFFu);                     // 1 -> +1, 0 -> -1
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// One thread per 16-byte piece of the image (16 consecutive K positions of one row).
// tile_src == nullptr: image row r is table row r (rows past n are zeros); else image tile t holds
// table rows tile_src[t] .. + tile_rows[t] - 1.
__global__ void knn2_tc_expand_kernel(const uint32_t *__restrict__ folded, size_t n, const uint32_t *__restrict__ tile_src,
                                      const uint32_t *__restrict__ tile_rows, uint4 *__restrict__ image, size_t n_pieces) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pieces) return;
    const size_t tile = idx >> 12;                       // 4096 pieces per tile
    const uint32_t rem = (uint32_t)(idx & 4095u);
    const uint32_t kc = rem >> 10, g = (rem >> 6) & 15u, cm = (rem >> 3) & 7u, i = rem & 7u;
    const uint32_t rt = g * 8u + i;                      // row within the tile
    size_t r;
    bool valid;
    if (tile_src != nullptr) {
        valid = rt < __ldg(tile_rows + tile);
        r = (size_t)__ldg(tile_src + tile) + rt;
    } else {
        r = tile * kTcTileRows + rt;
        valid = r < n;
    }
    image[idx] = valid ? expand_piece(folded + r * 16, kc * 16u + cm * 2u) : make_uint4(0u, 0u, 0u, 0u);
}

}  // namespace

cudaError_t knn2_tc_expand_launch(const uint4 *folded_rows, size_t n, uint8_t *image, cudaStream_t stream) {
    const size_t n_pieces = knn2_tc_image_bytes(n) / 16;
    const int threads = 256;
    knn2_tc_expand_kernel<<<(unsigned)((n_pieces + threads - 1) / threads), threads, 0, stream>>>(
        reinterpret_cast<const uint32_t *>(folded_rows), n, nullptr, nullptr, reinterpret_cast<uint4 *>(image), n_pieces);
    return cudaGetLastError();
}

cudaError_t knn2_tc_expand_tiles_launch(const uint4 *folded_rows, const uint32_t *tile_src, const uint32_t *tile_rows,
                                        size_t n_tiles, uint8_t *image, cudaStream_t stream) {
    if (n_tiles == 0) return cudaSuccess;
    const size_t n_pieces = n_tiles * (kTcTileBytes / 16);
    const int threads = 256;
    knn2_tc_expand_kernel<<<(unsigned)((n_pieces + threads - 1) / threads), threads, 0, stream>>>(
        reinterpret_cast<const uint32_t *>(folded_rows), 0, tile_src, tile_rows, reinterpret_cast<uint4 *>(image), n_pieces);
    return cudaGetLastError();
}

void knn2_tc_plan(size_t nA, size_t nB, int n_ctas, uint32_t *n_mtiles, uint32_t *n_chunks, uint32_t *rows_per_chunk) {
    const uint32_t mt = (uint32_t)((nA + kTcTileRows - 1) / kTcTileRows);
    *n_mtiles = mt;
    if (nB == 0 || mt == 0) { *n_chunks = 0; *rows_per_chunk = kTcTileN; return; }
    // Items (searcher tile x chunk) are dealt round-robin, so the makespan is (items per CTA, rounded
    // up) x (cost of an item).  Every item pays a fixed cost on top of its rows: the searcher tile
    // load, the pipeline refill and, above all, the best-2 thresholds restarting from nothing (the
    // first few thousand rows of a chunk take the slow path of the epilogue often).
    const uint64_t overhead_rows = 4096;
    const uint64_t tiles_b = (nB + kTcTileN - 1) / kTcTileN;
    const uint64_t c_min = (nB + kMaxChunkRows - 1) / kMaxChunkRows;
    uint64_t c_max = std::min<uint64_t>(tiles_b, std::max<uint64_t>(c_min, (32ull * n_ctas + mt - 1) / mt));
    uint64_t best_c = c_min, best_cost = ~0ull;
    for (uint64_t c = c_min; c <= c_max; ++c) {
        uint64_t rpc = ((nB + c - 1) / c + kTcTileN - 1) / kTcTileN * kTcTileN;
        if (rpc > kMaxChunkRows) continue;
        const uint64_t chunks = (nB + rpc - 1) / rpc;
        const uint64_t per_cta = ((uint64_t)mt * chunks + n_ctas - 1) / n_ctas;
        const uint64_t cost = per_cta * (rpc + overhead_rows);
        if (cost < best_cost) { best_cost = cost; best_c = c; }
    }
    uint64_t rpc = ((nB + best_c - 1) / best_c + kTcTileN - 1) / kTcTileN * kTcTileN;
    rpc = std::min<uint64_t>(rpc, kMaxChunkRows);
    *rows_per_chunk = (uint32_t)rpc;
    *n_chunks = (uint32_t)((nB + rpc - 1) / rpc);
}

template <int CL>
static cudaError_t launch_cl(const TcParams &p, int grid, cudaStream_t stream) {
    static thread_local int configured_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_device != dev) {
        cudaError_t e = cudaFuncSetAttribute(knn2_tc_kernel<CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
        configured_device = dev;
    }
    const uint32_t n_mgroups = (p.n_mtiles + CL - 1) / CL;
    const uint64_t n_items = p.items != nullptr ? p.n_items : (uint64_t)n_mgroups * p.n_chunks;
    if (n_items == 0) return cudaSuccess;
    uint64_t clusters = (uint64_t)grid / CL;
    if (clusters > n_items) clusters = n_items;
    if (clusters == 0) clusters = 1;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(clusters * CL), 1, 1);
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, knn2_tc_kernel<CL>, p);
}

cudaError_t knn2_tc_launch(const TcParams &p, int grid, cudaStream_t stream) {
    if (p.items != nullptr) return launch_cl<1>(p, grid, stream);      // item lists are not clustered
    switch (p.cluster) {
        case 1: return launch_cl<1>(p, grid, stream);
        case 2: return launch_cl<2>(p, grid, stream);
        case 4: return launch_cl<4>(p, grid, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace hulo
