// knn2_tc4.cu -- K1t4: the flat Hamming 2-NN of K1t (knn2_tc.cu) with 4-bit operands:
// tcgen05.mma.kind::mxf4.block_scale, every bit b of a row as the e2m1 value 2b - 1 (+1.0 = 0x2,
// -1.0 = 0xA), all block scales 1.0 (ue8m0 0x7F), fp32 accumulation.  dot = 512 - 2 * hamming is an
// integer of magnitude <= 512, far below 2^24, so the fp32 accumulator holds it exactly and the
// (distance, index) keys are those of K1 / K1t bit for bit.  Half the operand bytes and half the
// tensor-pipe cycles per distance of the int8 form.
//
// Image layout (its own: rows are 256 bytes here): groups of 8 rows, 2 KB each --
//   offset(row r, byte k of 256) = (r / 8) * 2048 + (k / 16) * 128 + (r % 8) * 16 + k % 16
// so any range of whole groups is contiguous: the searcher tile (128 rows, 32 KB) and a database
// tile (224 rows, 56 KB, the whole K of it: one ring stage) are one bulk copy each.
// TMEM: accumulators at columns 0 and 256 (224 used of each), scale factors at column 480.
#include "knn2_tc.cuh"
#include "tc_ptx.cuh"

#include <algorithm>
#include <cstdlib>

namespace hulo {

namespace {

using namespace tcptx;

constexpr uint32_t kN4 = 224;                     // database rows per accumulator tile, flat mode
constexpr uint32_t kN4Items = 192;                // item mode: narrower tiles leave room for a second searcher tile
constexpr uint32_t kRowBytes4 = 256;
constexpr uint32_t kGroupBytes4 = 8 * kRowBytes4; // 2048
constexpr uint32_t kABytes4 = 128 * kRowBytes4;   // 32768
constexpr uint32_t kStageBytes4 = kN4 * kRowBytes4;   // 57344 (the widest stage: sizes the slack of the images)
constexpr int kStages4 = 3;
constexpr uint32_t kTmemCols4 = 512;
constexpr uint32_t kSfCol = 480;
template <uint32_t N, bool A2>
constexpr size_t smem_bytes4() {
    return 1024 + (A2 ? 2 : 1) * kABytes4 + (size_t)kStages4 * N * kRowBytes4 + 256 + 2 * 2 * 128 * sizeof(uint2);
}

constexpr float kThrNoneF = -1024.0f;
constexpr float kDotPastEnd = -514.0f;          // 512 - 2 * 513
constexpr uint32_t kLazyMinRows = 16384;   // shorter ranges: the epilogue's threshold test does not pay (scan32)
constexpr uint32_t kKeyPastEnd = 513u << kKeyIdxBits;

// D[tmem] (+)= A[smem] * B[smem]^T, e2m1 x e2m1 with one ue8m0 scale per 32 elements -> fp32; M = 128, K = 64
__device__ __forceinline__ void tc_mma_mxf4(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// cute::UMMA::InstrDescriptorBlockScaled: A = B = e2m1 (MXF4Format 1), K-major, N, ue8m0 scales, M = 128, K = 64
template <uint32_t N>
constexpr uint32_t idesc4() { return (1u << 7) | (1u << 10) | ((N >> 3) << 17) | (1u << 23) | ((128u >> 4) << 24); }

#define HULO_LDTM32(v, taddr)                                                                                       \
    asm volatile(                                                                                                   \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                   \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                  \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                   \
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),          \
          "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),    \
          "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),  \
          "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])   \
        : "r"(taddr)                                                                                                \
        : "memory")
#define HULO_LDTM16(v, taddr)                                                                                       \
    asm volatile(                                                                                                   \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                   \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                            \
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),          \
          "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])     \
        : "r"(taddr)                                                                                                \
        : "memory")
#define HULO_WAIT_LD32(v)                                                                                            \
    asm volatile("tcgen05.wait::ld.sync.aligned;"                                                                    \
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]),    \
                   "+f"(v[8]), "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), \
                   "+f"(v[16]), "+f"(v[17]), "+f"(v[18]), "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), \
                   "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]), "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31]) \
                 :                                                                                                    \
                 : "memory")

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// two sorted key pairs -> the smallest two of the four
__device__ __forceinline__ void merge2(uint32_t &p0, uint32_t &p1, uint32_t q0, uint32_t q1) {
    const uint32_t m = max(p0, q0);
    p0 = min(p0, q0);
    p1 = min(min(m, p1), q1);
}
// the same on two 16-bit keys per register (SASS VIMNMX.U16x2 / VIMNMX3.U16x2)
__device__ __forceinline__ void merge2x2(uint32_t &p0, uint32_t &p1, uint32_t q0, uint32_t q1) {
    const uint32_t m = __vmaxu2(p0, q0);
    p0 = __vminu2(p0, q0);
    p1 = __vimin3_u16x2(m, p1, q1);
}

// 32 accumulator columns of this thread's searcher row.  v[e] = 512 - 2 * distance to database row
// (row0 + e) of the range, an exact integer in fp32.  Fast path: the maximum of the 32 against the dot
// of the row's current second best (columns arrive in ascending row order, so a column that only ties
// the second best loses on the index and a strict compare is exact).  Slow path, branch-free: a 16-bit
// key (distance << 5 | e) per column, two to a register, and a tournament for the smallest two of
// each 16-bit lane with packed min/max -- log depth, 1.9 ALU instructions per column, which is what
// short database ranges (thresholds restart with every item) spend most of their epilogue in.
// Columns past the end of a range are given the dot kDotPastEnd by the caller ("distance 513"): they
// lose against every real row and are turned into "none" when the item's keys are written.
// `lazy` = false skips the fast path's test: against a short range (a query image's few thousand
// rows) some lane of the warp improves its pair in nearly every block, and the test only costs.
__device__ __forceinline__ void scan32(const float (&v)[32], uint32_t row0, uint32_t &best0, uint32_t &best1, float &thr,
                                       bool lazy) {
    if (lazy) {
        float m[11];
#pragma unroll
        for (int i = 0; i < 10; ++i) m[i] = fmax3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        m[10] = fmaxf(v[30], v[31]);
        const float bm = fmax3(fmax3(m[0], m[1], m[2]), fmax3(m[3], m[4], m[5]),
                               fmax3(fmax3(m[6], m[7], m[8]), m[9], m[10]));
        if (bm <= thr) return;
    }
    // key16(e) = distance * 32 + e = 8192 + e - 16 v, read off the low mantissa bits of
    // fma(v, -16, 2^23 + 8192 + e) (exact: the value is an integer in [2^23, 2^24)); columns e and
    // e + 16 share a register: (bits(e + 16) << 16) + bits(e) leaves key(e) in the low half and
    // key(e + 16) + 0x4B00 (the exponent bits of the low word, a constant) in the high half
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint32_t ta = (uint32_t)__float_as_int(fmaf(v[i], -16.0f, 8388608.0f + 8192.0f + (float)i));
        const uint32_t tb = (uint32_t)__float_as_int(fmaf(v[i + 16], -16.0f, 8388608.0f + 8192.0f + (float)(i + 16)));
        pk[i] = tb * 65536u + ta;
    }
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { lo[i] = __vminu2(pk[i], pk[i + 8]); hi[i] = __vmaxu2(pk[i], pk[i + 8]); }
#pragma unroll
    for (int i = 0; i < 4; ++i) merge2x2(lo[i], hi[i], lo[i + 4], hi[i + 4]);
    merge2x2(lo[0], hi[0], lo[2], hi[2]);
    merge2x2(lo[1], hi[1], lo[3], hi[3]);
    merge2x2(lo[0], hi[0], lo[1], hi[1]);
    // the best two of columns 0..15 (low halves) and of 16..31 (high halves) as full keys
    auto full = [&](uint32_t k16) { return ((k16 >> 5) << kKeyIdxBits) + row0 + (k16 & 31u); };
    uint32_t a0 = full(lo[0] & 0xFFFFu), a1 = full(hi[0] & 0xFFFFu);
    const uint32_t b0 = full((lo[0] >> 16) - 0x4B00u), b1 = full((hi[0] >> 16) - 0x4B00u);
    merge2(a0, a1, b0, b1);
    merge2(best0, best1, a0, a1);
    thr = best1 == kKeyNone ? kThrNoneF : (float)(512 - 2 * (int32_t)(best1 >> kKeyIdxBits));
}

// The same over 64 columns at once (two register blocks of one group): 16-bit keys
// (distance << 6 | e), one tournament and one widening per 64 columns instead of two.
__device__ __forceinline__ void scan64(const float (&va)[32], const float (&vb)[32], uint32_t row0, uint32_t &best0,
                                       uint32_t &best1, float &thr, bool lazy) {
    if (lazy) {
        float m[11];
#pragma unroll
        for (int i = 0; i < 10; ++i) m[i] = fmax3(va[3 * i], va[3 * i + 1], va[3 * i + 2]);
        m[10] = fmaxf(va[30], va[31]);
        float bm = fmax3(fmax3(m[0], m[1], m[2]), fmax3(m[3], m[4], m[5]), fmax3(fmax3(m[6], m[7], m[8]), m[9], m[10]));
#pragma unroll
        for (int i = 0; i < 10; ++i) m[i] = fmax3(vb[3 * i], vb[3 * i + 1], vb[3 * i + 2]);
        m[10] = fmaxf(vb[30], vb[31]);
        bm = fmaxf(bm, fmax3(fmax3(m[0], m[1], m[2]), fmax3(m[3], m[4], m[5]), fmax3(fmax3(m[6], m[7], m[8]), m[9], m[10])));
        if (bm <= thr) return;
    }
    // key16(e) = distance * 64 + e = 16384 + e - 32 v; columns e (block a) and e + 32 (block b) share a register
    uint32_t pk[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const uint32_t ta = (uint32_t)__float_as_int(fmaf(va[i], -32.0f, 8388608.0f + 16384.0f + (float)i));
        const uint32_t tb = (uint32_t)__float_as_int(fmaf(vb[i], -32.0f, 8388608.0f + 16384.0f + (float)(i + 32)));
        pk[i] = tb * 65536u + ta;
    }
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { lo[i] = __vminu2(pk[i], pk[i + 16]); hi[i] = __vmaxu2(pk[i], pk[i + 16]); }
#pragma unroll
    for (int i = 0; i < 8; ++i) merge2x2(lo[i], hi[i], lo[i + 8], hi[i + 8]);
#pragma unroll
    for (int i = 0; i < 4; ++i) merge2x2(lo[i], hi[i], lo[i + 4], hi[i + 4]);
    merge2x2(lo[0], hi[0], lo[2], hi[2]);
    merge2x2(lo[1], hi[1], lo[3], hi[3]);
    merge2x2(lo[0], hi[0], lo[1], hi[1]);
    auto full = [&](uint32_t k16) { return ((k16 >> 6) << kKeyIdxBits) + row0 + (k16 & 63u); };
    uint32_t a0 = full(lo[0] & 0xFFFFu), a1 = full(hi[0] & 0xFFFFu);
    const uint32_t b0 = full((lo[0] >> 16) - 0x4B00u), b1 = full((hi[0] >> 16) - 0x4B00u);
    merge2(a0, a1, b0, b1);
    merge2(best0, best1, a0, a1);
    thr = best1 == kKeyNone ? kThrNoneF : (float)(512 - 2 * (int32_t)(best1 >> kKeyIdxBits));
}

struct TcWork4 {
    uint32_t a_group;    // first 8-row group of the searcher tile in image A
    uint32_t a_rows;     // rows of the tile whose keys are written
    uint32_t b_group;    // first group of the database range in image B
    uint32_t b_rows;
    uint64_t out_slot0;
};
// Flat mode: item w -> searcher tile (w % n_mtiles), chunk (w / n_mtiles); item mode: the list entry
// (TcItem::a_tile / b_tile0 hold GROUP indices for this kernel).
__device__ __forceinline__ TcWork4 tc_work4(const TcParams &p, uint32_t w) {
    TcWork4 k;
    if (p.items != nullptr) {
        const TcItem it = p.items[w];
        k.a_group = it.a_tile; k.a_rows = it.a_rows; k.b_group = it.b_tile0; k.b_rows = it.b_rows;
        k.out_slot0 = it.out_slot0;
    } else {
        const uint32_t mt = w % p.n_mtiles, c = w / p.n_mtiles;
        const uint32_t b_row0 = c * p.rows_per_chunk;
        k.a_group = mt * 16u;
        k.a_rows = min(kTcTileRows, p.nA - mt * kTcTileRows);
        k.b_group = b_row0 / 8u;
        k.b_rows = min(p.rows_per_chunk, p.nB - b_row0);
        k.out_slot0 = (uint64_t)c * p.slot_stride + (uint64_t)mt * kTcTileRows;
    }
    return k;
}

// N = database rows per accumulator tile (a multiple of 64).  A2 = two searcher-tile buffers: the
// producer fetches the searcher tile of the next item while the MMAs of the current one run (item
// mode, where every few tiles bring a new searcher tile; with one buffer each change drained the ring).
// G = epilogue groups of four warps: every group drains N / G columns of every accumulator.  Two for
// the flat searches (the lazy epilogue keeps up with the tensor pipe); three for item lists, whose
// short ranges run the full update on nearly every block and are bound by ALU issue -- four more
// warps per SM to issue from.
template <uint32_t N, bool A2, uint32_t G>
__global__ void __launch_bounds__(64 + 128 * G, 1) knn2_tc4_kernel(const TcParams p) {
    constexpr uint32_t kStageB = N * kRowBytes4;
    constexpr uint32_t kNA = A2 ? 2u : 1u;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = smem_base;
    const uint32_t sB = smem_base + kNA * kABytes4;
    const uint32_t bars = sB + kStages4 * kStageB;
    const uint32_t bar_a_full = bars, bar_a_empty = bars + 16;          // [2] each
    const uint32_t bar_b_full = bars + 32, bar_b_empty = bar_b_full + 8 * kStages4;
    const uint32_t bar_acc_full = bar_b_empty + 8 * kStages4, bar_acc_empty = bar_acc_full + 16;
    const uint32_t tmem_slot = bar_acc_empty + 16;
    const uint32_t xchg = bars + 256;
    uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    volatile uint32_t *tmem_slot_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));
    uint2 *xchg_gen = reinterpret_cast<uint2 *>(smem_gen + (xchg - smem_base));

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_a_full + 8 * a, 1);
            mbar_init(bar_a_empty + 8 * a, 1);
        }
        for (int s = 0; s < kStages4; ++s) {
            mbar_init(bar_b_full + 8 * s, 1);
            mbar_init(bar_b_empty + 8 * s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_acc_full + 8 * b, 1);
            mbar_init(bar_acc_empty + 8 * b, 4 * G);       // every epilogue warp drains a part of every tile
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols4)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    // scale factors: every byte of columns 480..511 on all 128 lanes = 0x7F (ue8m0 1.0), whatever
    // layout the instruction reads them in
    if (warp >= 2 && warp < 6) {
        const uint32_t taddr = tmem_base + (((warp & 3u) * 32u) << 16) + kSfCol;
        const uint32_t one = 0x7F7F7F7Fu;
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
            "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
            "r"(one)
            : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // flat mode: items round-robin over the CTAs; item mode: a contiguous block of the list per CTA
    uint32_t w_begin, w_end, w_step;
    if (p.items != nullptr) {
        w_begin = (uint32_t)((uint64_t)blockIdx.x * p.n_items / gridDim.x);
        w_end = (uint32_t)((uint64_t)(blockIdx.x + 1) * p.n_items / gridDim.x);
        w_step = 1;
    } else {
        w_begin = blockIdx.x; w_end = p.n_mtiles * p.n_chunks; w_step = gridDim.x;
    }

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, ph = 0, a_loaded = 0xFFFFFFFFu, a_loads = 0;
            for (uint32_t w = w_begin; w < w_end; w += w_step) {
                const TcWork4 k = tc_work4(p, w);
                const uint32_t n_tiles = (k.b_rows + N - 1) / N;
                if (k.a_group != a_loaded) {
                    // load number a_loads goes to buffer a_loads % kNA; the buffer's previous tenant
                    // (load a_loads - kNA) is released by the MMA thread when it moves past it
                    const uint32_t slot = a_loads % kNA, use = a_loads / kNA;
                    if (use > 0) mbar_wait(bar_a_empty + 8 * slot, (use - 1u) & 1u);
                    mbar_expect_tx(bar_a_full + 8 * slot, kABytes4);
                    bulk_load(sA + slot * kABytes4, p.imgA + (size_t)k.a_group * kGroupBytes4, kABytes4, bar_a_full + 8 * slot);
                    a_loaded = k.a_group;
                    ++a_loads;
                }
                const uint8_t *src = p.imgB + (size_t)k.b_group * kGroupBytes4;
                for (uint32_t t = 0; t < n_tiles; ++t) {
                    mbar_wait(bar_b_empty + 8 * stage, ph ^ 1u);
                    mbar_expect_tx(bar_b_full + 8 * stage, kStageB);
                    bulk_load(sB + stage * kStageB, src + (size_t)t * kStageB, kStageB, bar_b_full + 8 * stage);
                    if (++stage == kStages4) { stage = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t stage = 0, ph = 0, acc_it = 0, a_loaded = 0xFFFFFFFFu, a_loads = 0, a_slot = 0;
            const uint32_t tsf = tmem_base + kSfCol;
            for (uint32_t w = w_begin; w < w_end; w += w_step) {
                const TcWork4 k = tc_work4(p, w);
                const uint32_t n_tiles = (k.b_rows + N - 1) / N;
                if (k.a_group != a_loaded) {
                    // every MMA issued so far read an earlier tile: the buffer of the previous one is free
                    // once they are done
                    if (a_loads > 0) tc_commit(bar_a_empty + 8 * ((a_loads - 1u) % kNA));
                    a_slot = a_loads % kNA;
                    mbar_wait(bar_a_full + 8 * a_slot, (a_loads / kNA) & 1u);
                    a_loaded = k.a_group;
                    ++a_loads;
                }
                for (uint32_t t = 0; t < n_tiles; ++t, ++acc_it) {
                    const uint32_t buf = acc_it & 1u;
                    mbar_wait(bar_acc_empty + 8 * buf, ((acc_it >> 1) & 1u) ^ 1u);
                    mbar_wait(bar_b_full + 8 * stage, ph);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + buf * 256u;
#pragma unroll
                    for (uint32_t j = 0; j < 8; ++j) {
                        const uint64_t da = smem_desc(sA + a_slot * kABytes4 + j * 256u, 128u, kGroupBytes4);
                        const uint64_t db = smem_desc(sB + stage * kStageB + j * 256u, 128u, kGroupBytes4);
                        tc_mma_mxf4(tmem_d, da, db, idesc4<N>(), tsf, tsf, j != 0u);
                    }
                    tc_commit(bar_b_empty + 8 * stage);
                    tc_commit(bar_acc_full + 8 * buf);
                    if (++stage == kStages4) { stage = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        const uint32_t quarter = warp & 3u;
        const uint32_t grp = (warp - 2u) >> 2;
        const uint32_t row = quarter * 32u + lane;
        // Both groups work on EVERY tile, each on half of its columns (group g: [112 g, 112 g + 112)), so an
        // accumulator is drained in half the time and handed back to the MMA thread sooner: with one
        // group per accumulator the tensor pipe sat at 63 % (buffer cycle = MMA time + a whole drain).
        uint32_t acc_base = 0, it = 0;
        constexpr uint32_t kHalf = N / G;                        // 112 = 32 + 32 + 32 + 16, 96 = 32 + 32 + 32, or 64 = 32 + 32
        constexpr bool kTail16 = (kHalf % 32u) != 0u;
        constexpr bool kThree = kHalf / 32u == 3u;
        static_assert(kHalf * G == N && (kThree || (kHalf == 64u && !kTail16)),
                      "the epilogue is written for three full blocks (+16 columns) or exactly two per group");
        for (uint32_t w = w_begin; w < w_end; w += w_step, ++it) {
            const TcWork4 k = tc_work4(p, w);
            const uint32_t b_rows = k.b_rows;
            const uint32_t n_tiles = (b_rows + N - 1) / N;
            uint32_t best0 = kKeyNone, best1 = kKeyNone;
            float thr = kThrNoneF;
            const bool lazy = b_rows > kLazyMinRows;
            for (uint32_t t = 0; t < n_tiles; ++t) {
                const uint32_t acc_it = acc_base + t, buf = acc_it & 1u;
                mbar_wait(bar_acc_full + 8 * buf, (acc_it >> 1) & 1u);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + buf * 256u + grp * kHalf;
                const uint32_t n_valid = min(N, b_rows - t * N);           // valid columns of the tile
                const uint32_t c0 = grp * kHalf;                           // this group's first column
                const bool dump = p.dbg_dots != nullptr && w == 0 && t == 0 && p.items == nullptr;
                float va[32], vb[32];
                auto process = [&](float (&v)[32], uint32_t blk) {
                    const uint32_t col = c0 + 32u * blk;
                    if (dump) _Pragma("unroll") for (int e = 0; e < 32; ++e)
                        if (blk < 3u || e < 16) p.dbg_dots[row * 256 + col + e] = (int32_t)v[e];
                    if (n_valid < col + 32u) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = col + (uint32_t)e < n_valid ? v[e] : kDotPastEnd;
                    }
                    scan32(v, t * N + col, best0, best1, thr, lazy);
                };
                auto release = [&]() {
                    // every column of this half is in registers: hand the accumulator back before the last scan
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
                };
                if constexpr (!kThree) {
                    // two blocks per group: both in registers, one 64-column update
                    HULO_LDTM32(va, taddr);
                    HULO_LDTM32(vb, taddr + 32u);
                    HULO_WAIT_LD32(va);
                    HULO_WAIT_LD32(vb);
                    release();
                    if (n_valid < c0 + 64u) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) {
                            va[e] = c0 + (uint32_t)e < n_valid ? va[e] : kDotPastEnd;
                            vb[e] = c0 + 32u + (uint32_t)e < n_valid ? vb[e] : kDotPastEnd;
                        }
                    }
                    scan64(va, vb, t * N + c0, best0, best1, thr, lazy);
                    continue;
                }
                HULO_LDTM32(va, taddr);
                HULO_WAIT_LD32(va);
                HULO_LDTM32(vb, taddr + 32u);
                process(va, 0u);
                HULO_WAIT_LD32(vb);
                HULO_LDTM32(va, taddr + 64u);
                process(vb, 1u);
                HULO_WAIT_LD32(va);
                if constexpr (kTail16) {
                    HULO_LDTM16(vb, taddr + 96u);                          // the last 16 columns of the half
#pragma unroll
                    for (int e = 16; e < 32; ++e) vb[e] = kDotPastEnd;
                    process(va, 2u);
                    HULO_WAIT_LD32(vb);
                    release();
                    process(vb, 3u);
                } else {
                    release();
                    process(va, 2u);
                }
            }
            acc_base += n_tiles;
            if (best0 >= kKeyPastEnd) best0 = kKeyNone;           // columns past the end of the range
            if (best1 >= kKeyPastEnd) best1 = kKeyNone;
            uint2 *slot = xchg_gen + (it & 1u) * 256u + row;     // [parity][group - 1][row]
            if (grp != 0u) slot[(grp - 1u) * 128u] = make_uint2(best0, best1);
            asm volatile("bar.sync 1, %0;" ::"n"(128 * G) : "memory");
            if (grp == 0u) {
#pragma unroll
                for (uint32_t o_g = 0; o_g + 1u < G; ++o_g) {
                    const uint2 o = slot[o_g * 128u];
                    merge2(best0, best1, o.x, o.y);
                }
                if (row < k.a_rows) p.partial[k.out_slot0 + row] = make_uint2(best0, best1);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols4) : "memory");
    }
}

// One thread per 16-byte piece of the image = 32 consecutive bits of one row.
// group_src == nullptr: image row r is table row r; else image group g holds table rows
// group_src[g] .. + group_rows[g] - 1 (segmented tables: every segment starts on a group).
__global__ void knn2_tc4_expand_kernel(const uint32_t *__restrict__ folded, size_t n, const uint32_t *__restrict__ group_src,
                                       const uint8_t *__restrict__ group_rows, size_t n_groups_tab,
                                       uint4 *__restrict__ image, size_t n_pieces) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_pieces) return;
    const size_t group = idx >> 7;                       // 128 pieces per group of 8 rows
    const uint32_t rem = (uint32_t)(idx & 127u);
    const uint32_t core = rem >> 3, i = rem & 7u;        // word `core` of the row
    size_t r;
    bool valid;
    if (group_src != nullptr) {
        valid = group < n_groups_tab && i < __ldg(group_rows + group);
        r = valid ? (size_t)__ldg(group_src + group) + i : 0;
    } else {
        r = group * 8 + i;
        valid = r < n;
    }
    uint4 out = make_uint4(0u, 0u, 0u, 0u);              // padding rows: 0.0 everywhere
    if (valid) {
        const uint32_t *f = folded + r * 16;
        uint32_t wv = __ldg(f + core);
        if (core == 15u) wv ^= __ldg(f + 11) ^ __ldg(f + 14);
        else if (core % 3u == 2u) wv ^= __ldg(f + core - 1) ^ __ldg(f + core - 2);
        uint32_t o[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t x = (wv >> (8 * q)) & 255u;
            // bit b of x -> nibble b: 0x2 | (bit << 3)
            uint32_t sp = (x | (x << 12)) & 0x000F000Fu;          // 4 bits per 16-bit half
            sp = (sp | (sp << 6)) & 0x03030303u;                  // 2 bits per byte
            sp = (sp | (sp << 3)) & 0x11111111u;                  // 1 bit per nibble
            o[q] = 0x22222222u | (sp << 3);
        }
        out = make_uint4(o[0], o[1], o[2], o[3]);
    }
    image[idx] = out;
}

}  // namespace

size_t knn2_tc4_image_bytes(size_t n) {
    // whole groups, plus one database tile of slack: the last tile of a chunk is always read in full
    return ((n + 7) / 8) * (size_t)kGroupBytes4 + kStageBytes4 + kABytes4;
}

cudaError_t knn2_tc4_expand_launch(const uint4 *folded_rows, size_t n, uint8_t *image, cudaStream_t stream) {
    const size_t n_pieces = knn2_tc4_image_bytes(n) / 16;
    const int threads = 256;
    knn2_tc4_expand_kernel<<<(unsigned)((n_pieces + threads - 1) / threads), threads, 0, stream>>>(
        reinterpret_cast<const uint32_t *>(folded_rows), n, nullptr, nullptr, 0, reinterpret_cast<uint4 *>(image), n_pieces);
    return cudaGetLastError();
}

size_t knn2_tc4_groups_image_bytes(size_t n_groups) { return n_groups * (size_t)kGroupBytes4 + kStageBytes4 + kABytes4; }

cudaError_t knn2_tc4_expand_groups_launch(const uint4 *folded_rows, const uint32_t *group_src, const uint8_t *group_rows,
                                          size_t n_groups, uint8_t *image, cudaStream_t stream) {
    const size_t n_pieces = knn2_tc4_groups_image_bytes(n_groups) / 16;
    const int threads = 256;
    knn2_tc4_expand_kernel<<<(unsigned)((n_pieces + threads - 1) / threads), threads, 0, stream>>>(
        reinterpret_cast<const uint32_t *>(folded_rows), 0, group_src, group_rows, n_groups, reinterpret_cast<uint4 *>(image),
        n_pieces);
    return cudaGetLastError();
}

void knn2_tc4_plan(size_t nA, size_t nB, int n_ctas, uint32_t *n_mtiles, uint32_t *n_chunks, uint32_t *rows_per_chunk) {
    const uint32_t mt = (uint32_t)((nA + kTcTileRows - 1) / kTcTileRows);
    *n_mtiles = mt;
    if (nB == 0 || mt == 0) { *n_chunks = 0; *rows_per_chunk = kN4; return; }
    const uint64_t overhead_rows = 4096;
    const uint64_t tiles_b = (nB + kN4 - 1) / kN4;
    const uint64_t c_min = (nB + kMaxChunkRows - 1) / kMaxChunkRows;
    const uint64_t c_max = std::min<uint64_t>(tiles_b, std::max<uint64_t>(c_min, (32ull * n_ctas + mt - 1) / mt));
    uint64_t best_c = c_min, best_cost = ~0ull;
    for (uint64_t c = c_min; c <= c_max; ++c) {
        const uint64_t rpc = ((nB + c - 1) / c + kN4 - 1) / kN4 * kN4;
        if (rpc > kMaxChunkRows) continue;
        const uint64_t chunks = (nB + rpc - 1) / rpc;
        const uint64_t per_cta = ((uint64_t)mt * chunks + n_ctas - 1) / n_ctas;
        const uint64_t cost = per_cta * (rpc + overhead_rows);
        if (cost < best_cost) { best_cost = cost; best_c = c; }
    }
    uint64_t rpc = ((nB + best_c - 1) / best_c + kN4 - 1) / kN4 * kN4;
    rpc = std::min<uint64_t>(rpc, kMaxChunkRows / kN4 * kN4);
    *rows_per_chunk = (uint32_t)rpc;
    *n_chunks = (uint32_t)((nB + rpc - 1) / rpc);
}

template <uint32_t N, bool A2, uint32_t G>
static cudaError_t launch4(const TcParams &p, int grid, cudaStream_t stream) {
    static thread_local int configured_device = -1;
    constexpr size_t smem = smem_bytes4<N, A2>();
    int dev = 0;
    cudaGetDevice(&dev);
    if (configured_device != dev) {
        cudaError_t e = cudaFuncSetAttribute(knn2_tc4_kernel<N, A2, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured_device = dev;
    }
    const uint64_t n_items = p.items != nullptr ? p.n_items : (uint64_t)p.n_mtiles * p.n_chunks;
    if (n_items == 0) return cudaSuccess;
    if ((uint64_t)grid > n_items) grid = (int)n_items;
    knn2_tc4_kernel<N, A2, G><<<grid, 64 + 128 * G, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t knn2_tc4_launch(const TcParams &p, int grid, cudaStream_t stream) {
    // item lists: 192-row tiles and two searcher-tile buffers; flat searches: 224-row tiles, one buffer
    if (p.items != nullptr) {
        static const bool two = getenv("HULO_TC4_ITEM_GROUPS") != nullptr && atoi(getenv("HULO_TC4_ITEM_GROUPS")) == 2;
        return two ? launch4<kN4Items, true, 2>(p, grid, stream) : launch4<kN4Items, true, 3>(p, grid, stream);
    }
    return launch4<kN4, false, 2>(p, grid, stream);
}

}  // namespace hulo
