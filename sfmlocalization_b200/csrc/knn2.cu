// knn2.cu -- K1 kernels (see knn2.cuh for the design).
#include "knn2.cuh"

#include <climits>

namespace hulo {

namespace {

constexpr int kTileRows = 64;  // database rows per smem stage (4 KB)
constexpr int kStages = 3;

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ uint32_t xor3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ uint32_t maj3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// carry of a + b + c when the third operand is given as the SUM bit s = a ^ b ^ c:
// maj(a, b, a ^ b ^ s), one LOP3 (truth table 0xD4)
__device__ __forceinline__ uint32_t carry_from_sum(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xD4;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}
// carry-save adder: three words of weight w -> one of weight w (lo) and one of weight 2w (hi)
#define HULO_CSA(hi, lo, a, b, c) \
    do {                          \
        lo = xor3(a, b, c);       \
        hi = maj3(a, b, c);       \
    } while (0)

// acc + m * popc(x): the multiply-add runs on the FMA pipe (IMAD), which K1 leaves idle, instead
// of an IADD3 on the ALU pipe that the XORs and adders saturate.
__device__ __forceinline__ uint32_t popc_mad(uint32_t x, uint32_t m, uint32_t acc) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"((uint32_t)__popc(x)), "r"(m), "r"(acc));
    return r;
}

// Rows are stored FOLDED (knn2_fold_rows): words 3i+2 (i = 0..4) hold w[3i] ^ w[3i+1] ^ w[3i+2],
// and word 15 holds w9 ^ w10 ^ ... ^ w15.  XOR is linear, so q'[3i+2] ^ w'[3i+2] is directly the sum
// output of the first-level carry-save adder over the triple (x[3i], x[3i+1], x[3i+2]) of the
// unfolded difference, and its carry follows from the two plain differences and that sum in one
// LOP3: 4 instead of 5 LOP3 per triple.  Likewise q'[15] ^ w'[15] is the sum output of the
// second-level adder over (s3, s4, x[15]).  6 fewer ALU instructions per distance, bit-identical
// distances.
//
// key = base + (distance << kKeyIdxBits) for a register-resident searcher row q and a database
// row w; distance = 512-bit Hamming.  CSA = number of carry-save adders applied before the POPCs
// (0: plain 16 POPC).  `unit` is 1 << kKeyIdxBits held in a register so ptxas keeps the IMADs.
template <int CSA, bool IMAD>
__device__ __forceinline__ uint32_t hamming_key(const uint32_t (&q)[16], const uint32_t (&w)[16], uint32_t base,
                                                uint32_t unit) {
    uint32_t x[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) x[k] = q[k] ^ w[k];
    uint32_t n1 = 0, n2 = 0, n4 = 0, n8 = 0;      // counts of weight 1 / 2 / 4 / 8 (plain-add path)
    uint32_t key = base;                           // IMAD path accumulates straight into the key
#define HULO_ACC(word, weight, cnt)                                  \
    do {                                                             \
        if constexpr (IMAD) key = popc_mad(word, unit * (weight), key); \
        else cnt += __popc(word);                                    \
    } while (0)
    if constexpr (CSA == 0) {
        // plain 16 POPC baseline: undo the fold of the difference first
        x[15] = xor3(x[15], x[11], x[14]);
#pragma unroll
        for (int i = 0; i < 5; ++i) x[3 * i + 2] = xor3(x[3 * i], x[3 * i + 1], x[3 * i + 2]);
#pragma unroll
        for (int k = 0; k < 16; ++k) HULO_ACC(x[k], 1u, n1);
    } else {
        // level 1: 15 words -> 5 sums (weight 1) + 5 carries (weight 2); x[15] left over.
        // The folded layout delivers the sums; only the carries are computed.
        const uint32_t s0 = x[2], s1 = x[5], s2 = x[8], s3 = x[11], s4 = x[14];
        const uint32_t c0 = carry_from_sum(x[0], x[1], s0);
        const uint32_t c1 = carry_from_sum(x[3], x[4], s1);
        const uint32_t c2 = carry_from_sum(x[6], x[7], s2);
        const uint32_t c3 = carry_from_sum(x[9], x[10], s3);
        const uint32_t c4 = carry_from_sum(x[12], x[13], s4);
        if constexpr (CSA == 5) {
            HULO_ACC(s0, 1u, n1); HULO_ACC(s1, 1u, n1); HULO_ACC(s2, 1u, n1); HULO_ACC(s3, 1u, n1);
            HULO_ACC(s4, 1u, n1); HULO_ACC(xor3(x[15], s3, s4), 1u, n1);
            HULO_ACC(c0, 2u, n2); HULO_ACC(c1, 2u, n2); HULO_ACC(c2, 2u, n2); HULO_ACC(c3, 2u, n2);
            HULO_ACC(c4, 2u, n2);
        } else {
            // level 2: the six weight-1 words -> 2 sums + 2 carries
            uint32_t s5, c5;
            HULO_CSA(c5, s5, s0, s1, s2);
            const uint32_t s6 = x[15];                             // folded: s3 ^ s4 ^ (plain x15)
            const uint32_t c6 = carry_from_sum(s3, s4, s6);
            HULO_ACC(s5, 1u, n1); HULO_ACC(s6, 1u, n1);
            if constexpr (CSA == 7) {
                HULO_ACC(c0, 2u, n2); HULO_ACC(c1, 2u, n2); HULO_ACC(c2, 2u, n2); HULO_ACC(c3, 2u, n2);
                HULO_ACC(c4, 2u, n2); HULO_ACC(c5, 2u, n2); HULO_ACC(c6, 2u, n2);
            } else {
                // level 3: weight-2 words folded three at a time
                uint32_t t0, f0;
                HULO_CSA(f0, t0, c0, c1, c2);
                if constexpr (CSA == 8) {
                    HULO_ACC(t0, 2u, n2); HULO_ACC(c3, 2u, n2); HULO_ACC(c4, 2u, n2); HULO_ACC(c5, 2u, n2);
                    HULO_ACC(c6, 2u, n2); HULO_ACC(f0, 4u, n4);
                } else {
                    uint32_t t1, f1;
                    HULO_CSA(f1, t1, c3, c4, c5);
                    if constexpr (CSA == 9) {
                        HULO_ACC(t0, 2u, n2); HULO_ACC(t1, 2u, n2); HULO_ACC(c6, 2u, n2);
                        HULO_ACC(f0, 4u, n4); HULO_ACC(f1, 4u, n4);
                    } else {
                        static_assert(CSA == 11, "unsupported CSA depth");
                        uint32_t t2, f2, f3, e0;
                        HULO_CSA(f2, t2, t0, t1, c6);
                        HULO_CSA(e0, f3, f0, f1, f2);
                        HULO_ACC(t2, 2u, n2); HULO_ACC(f3, 4u, n4); HULO_ACC(e0, 8u, n8);
                    }
                }
            }
        }
    }
#undef HULO_ACC
    if constexpr (IMAD) return key;
    else return base + ((n1 + 2u * n2 + 4u * n4 + 8u * n8) << kKeyIdxBits);
}

// One database row against the QPT register-resident searcher rows of a thread, unrolled at
// compile time.  CSA == 89 is a mix: the last searcher row of the thread uses nine carry-save
// adders, the others eight -- per distance that is 27.5 LOP3 + 7.75 POPC, which puts the ALU
// pipe (2 cycles per warp instruction) and the XU pipe (8 cycles per POPC) at the same load.
template <int QI, int QPT, int CSA, bool IMAD>
struct RowVsQueries {
    static __device__ __forceinline__ void run(const uint32_t (&q)[QPT][16], const uint32_t (&w)[16], uint32_t base,
                                               uint32_t unit, uint32_t (&best0)[QPT], uint32_t (&best1)[QPT]) {
        constexpr int kCsa = CSA == 89 ? (QI == QPT - 1 ? 9 : 8) : CSA == 889 ? (2 * QI >= QPT ? 9 : 8) : CSA;
        const uint32_t key = hamming_key<kCsa, IMAD>(q[QI], w, base, unit);
        const uint32_t hi = max(best0[QI], key);
        best0[QI] = min(best0[QI], key);
        best1[QI] = min(best1[QI], hi);
        if constexpr (QI + 1 < QPT) RowVsQueries<QI + 1, QPT, CSA, IMAD>::run(q, w, base, unit, best0, best1);
    }
};

// The same with a LAZY best-2 update: the keys of the row are computed for every searcher row of
// the thread, but the three min/max per key only run when some lane of the warp holds a key below
// its current second best -- one compare per key and one vote per row otherwise.  After a few
// thousand rows of a chunk almost no row qualifies, so on long chunks this takes two ALU
// instructions per distance out of the loop; on short chunks (a few hundred rows) it loses.
template <int QI, int QPT, int CSA, bool IMAD>
struct RowKeys {
    static __device__ __forceinline__ bool run(const uint32_t (&q)[QPT][16], const uint32_t (&w)[16], uint32_t base,
                                               uint32_t unit, const uint32_t (&best1)[QPT], uint32_t (&keys)[QPT]) {
        constexpr int kCsa = CSA == 89 ? (QI == QPT - 1 ? 9 : 8) : CSA == 889 ? (2 * QI >= QPT ? 9 : 8) : CSA;
        keys[QI] = hamming_key<kCsa, IMAD>(q[QI], w, base, unit);
        bool hit = keys[QI] < best1[QI];
        if constexpr (QI + 1 < QPT) hit = RowKeys<QI + 1, QPT, CSA, IMAD>::run(q, w, base, unit, best1, keys) || hit;
        return hit;
    }
};

// OPT bit 0: adds on the FMA pipe (IMAD); bit 1: rows taken two at a time so the best-2 update
// uses 3-input min/max (VIMNMX3): 5 instead of 6 instructions per two keys; bit 2: lazy best-2
// update (RowKeys).
template <int THREADS, int QPT, int CSA, int OPT>
__global__ void __launch_bounds__(THREADS, 1) knn2_kernel(const KnnParams p) {
    constexpr bool kImad = (OPT & 1) != 0;
    constexpr bool kPairRows = (OPT & 2) != 0;
    constexpr bool kLazy = (OPT & 4) != 0;
    constexpr int kRowUnroll = (OPT & 8) ? 1 : (OPT & 16) ? 4 : 2;     // rows per trip of the inner loop
    constexpr int kWarps = THREADS / 32;
    __shared__ __align__(128) uint4 s_tiles[kStages][kTileRows * 4];
    __shared__ __align__(8) uint64_t s_full[kStages];
    __shared__ __align__(8) uint64_t s_empty[kStages];
    __shared__ uint32_t s_item;

    const int tid = threadIdx.x;
    const int lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], kWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    uint32_t n_cons = 0;   // tiles consumed so far by this thread (all threads agree)
    uint32_t n_prod = 0;   // tiles issued so far (thread 0)

    for (;;) {
        if (tid == 0) s_item = atomicAdd(p.counter, 1u);
        __syncthreads();
        const uint32_t item = s_item;
        __syncthreads();
        if (item >= p.n_items) break;

        uint32_t a_row0, a_rows, b_row0, b_rows;
        uint64_t out_slot0;
        if (p.items != nullptr) {
            const KnnItem it = p.items[item];
            a_row0 = it.a_row0; a_rows = it.a_rows; b_row0 = it.b_row0; b_rows = it.b_rows;
            out_slot0 = it.out_slot0;
        } else {
            const uint32_t tile = item % p.n_tiles, chunk = item / p.n_tiles;
            a_row0 = tile * (uint32_t)(THREADS * QPT);
            a_rows = min((uint32_t)(THREADS * QPT), p.nA - a_row0);
            b_row0 = chunk * p.rows_per_chunk;
            b_rows = min(p.rows_per_chunk, p.nB - b_row0);
            out_slot0 = (uint64_t)chunk * p.slot_stride + a_row0;
        }

        // searcher rows -> registers
        uint32_t q[QPT][16];
        uint32_t best0[QPT], best1[QPT];
#pragma unroll
        for (int qi = 0; qi < QPT; ++qi) {
            const uint32_t r = (uint32_t)(qi * THREADS + tid);
            best0[qi] = kKeyNone;
            best1[qi] = kKeyNone;
            if (r < a_rows) {
                const uint4 *src = p.A + (size_t)(a_row0 + r) * 4;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint4 t = __ldg(src + v);
                    q[qi][4 * v + 0] = t.x; q[qi][4 * v + 1] = t.y;
                    q[qi][4 * v + 2] = t.z; q[qi][4 * v + 3] = t.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < 16; ++k) q[qi][k] = 0u;
            }
        }

        const uint32_t n_tiles = (b_rows + kTileRows - 1) / kTileRows;
        const uint4 *bsrc = p.B + (size_t)b_row0 * 4;

        // producer prologue: fill the ring
        if (tid == 0) {
            const uint32_t pre = n_tiles < (uint32_t)kStages ? n_tiles : (uint32_t)kStages;
            for (uint32_t t = 0; t < pre; ++t) {
                const uint32_t stage = n_prod % kStages;
                if (n_prod >= (uint32_t)kStages) mbar_wait(&s_empty[stage], ((n_prod / kStages) - 1) & 1);
                const uint32_t rows = min((uint32_t)kTileRows, b_rows - t * kTileRows);
                mbar_expect_tx(&s_full[stage], rows * 64u);
                tma_load_1d(&s_tiles[stage][0], bsrc + (size_t)t * kTileRows * 4, rows * 64u, &s_full[stage]);
                ++n_prod;
            }
        }

        for (uint32_t t = 0; t < n_tiles; ++t) {
            const uint32_t stage = n_cons % kStages;
            mbar_wait(&s_full[stage], (n_cons / kStages) & 1);
            const uint32_t rows = min((uint32_t)kTileRows, b_rows - t * kTileRows);
            const uint4 *tile = &s_tiles[stage][0];
            const uint32_t key_base = t * kTileRows;
            const uint32_t unit = p.key_unit;      // 1 << kKeyIdxBits, opaque to the compiler
            auto load_row = [&](uint32_t r, uint32_t (&w)[16]) {
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint4 tv = tile[r * 4 + v];   // same address in every lane: broadcast
                    w[4 * v + 0] = tv.x; w[4 * v + 1] = tv.y; w[4 * v + 2] = tv.z; w[4 * v + 3] = tv.w;
                }
            };
            uint32_t r = 0;
            if constexpr (kPairRows) {
                for (; r + 2 <= rows; r += 2) {
                    uint32_t w0[16], w1[16];
                    load_row(r, w0);
                    load_row(r + 1, w1);
#pragma unroll
                    for (int qi = 0; qi < QPT; ++qi) {
                        constexpr int kCsaPair = (CSA == 89 || CSA == 889) ? 8 : CSA;
                        const uint32_t k0 = hamming_key<kCsaPair, kImad>(q[qi], w0, key_base + r, unit);
                        const uint32_t k1 = hamming_key<kCsaPair, kImad>(q[qi], w1, key_base + r + 1, unit);
                        const uint32_t lo = min(k0, k1), hi = max(k0, k1);
                        const uint32_t m = max(best0[qi], lo);
                        best0[qi] = min(best0[qi], lo);
                        best1[qi] = min(min(best1[qi], hi), m);
                    }
                }
            }
#pragma unroll kRowUnroll
            for (; r < rows; ++r) {
                uint32_t w[16];
                load_row(r, w);
                if constexpr (kLazy) {
                    uint32_t keys[QPT];
                    const bool hit = RowKeys<0, QPT, CSA, kImad>::run(q, w, key_base + r, unit, best1, keys);
                    if (__any_sync(0xffffffffu, hit)) {
#pragma unroll
                        for (int qi = 0; qi < QPT; ++qi) {
                            const uint32_t hi = max(best0[qi], keys[qi]);
                            best0[qi] = min(best0[qi], keys[qi]);
                            best1[qi] = min(best1[qi], hi);
                        }
                    }
                } else {
                    RowVsQueries<0, QPT, CSA, kImad>::run(q, w, key_base + r, unit, best0, best1);
                }
            }
            ++n_cons;
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[stage]);
            // producer: refill this stage with tile t + kStages
            if (tid == 0 && t + kStages < n_tiles) {
                const uint32_t tn = t + kStages;
                const uint32_t pstage = n_prod % kStages;   // == stage
                mbar_wait(&s_empty[pstage], ((n_prod / kStages) - 1) & 1);
                const uint32_t prow = min((uint32_t)kTileRows, b_rows - tn * kTileRows);
                mbar_expect_tx(&s_full[pstage], prow * 64u);
                tma_load_1d(&s_tiles[pstage][0], bsrc + (size_t)tn * kTileRows * 4, prow * 64u, &s_full[pstage]);
                ++n_prod;
            }
        }

#pragma unroll
        for (int qi = 0; qi < QPT; ++qi) {
            const uint32_t r = (uint32_t)(qi * THREADS + tid);
            if (r < a_rows) p.partial[out_slot0 + r] = make_uint2(best0[qi], best1[qi]);
        }
    }
}

// ------------------------------------------------------------------- merge
__device__ __forceinline__ void top2_insert(uint64_t k, uint64_t &m0, uint64_t &m1) {
    const uint64_t hi = k > m0 ? k : m0;
    m0 = k < m0 ? k : m0;
    m1 = hi < m1 ? hi : m1;
}

constexpr uint64_t kNone64 = ~0ull;

__device__ __forceinline__ void write_result(uint64_t m0, uint64_t m1, uint32_t row, int32_t *out_idx2,
                                             int32_t *out_dist2, int4 *out_packed) {
    const int32_t d0 = m0 == kNone64 ? INT_MAX : (int32_t)(m0 >> 32);
    const int32_t i0 = m0 == kNone64 ? -1 : (int32_t)(uint32_t)m0;
    const int32_t d1 = m1 == kNone64 ? INT_MAX : (int32_t)(m1 >> 32);
    const int32_t i1 = m1 == kNone64 ? -1 : (int32_t)(uint32_t)m1;
    if (out_idx2) reinterpret_cast<int2 *>(out_idx2)[row] = make_int2(i0, i1);
    if (out_dist2) reinterpret_cast<int2 *>(out_dist2)[row] = make_int2(d0, d1);
    if (out_packed) out_packed[row] = make_int4(d0, i0, d1, i1);
}

__global__ void knn2_merge_kernel(const uint2 *__restrict__ partial, uint32_t nA, uint32_t n_chunks,
                                  uint64_t slot_stride, uint32_t rows_per_chunk, uint32_t row_base,
                                  int32_t *out_idx2, int32_t *out_dist2, int4 *out_packed) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nA) return;
    uint64_t m0 = kNone64, m1 = kNone64;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint2 k = partial[(uint64_t)c * slot_stride + row];
        const uint32_t base = row_base + c * rows_per_chunk;
        if (k.x != kKeyNone)
            top2_insert(((uint64_t)(k.x >> kKeyIdxBits) << 32) | (uint64_t)(base + (k.x & kKeyIdxMask)), m0, m1);
        if (k.y != kKeyNone)
            top2_insert(((uint64_t)(k.y >> kKeyIdxBits) << 32) | (uint64_t)(base + (k.y & kKeyIdxMask)), m0, m1);
    }
    write_result(m0, m1, row, out_idx2, out_dist2, out_packed);
}

__global__ void knn2_merge_ranks_kernel(const int4 *__restrict__ gathered, uint32_t nA, int world,
                                        int32_t *out_idx2, int32_t *out_dist2) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nA) return;
    uint64_t m0 = kNone64, m1 = kNone64;
    for (int g = 0; g < world; ++g) {
        const int4 c = gathered[(size_t)g * nA + row];
        if (c.y >= 0) top2_insert(((uint64_t)(uint32_t)c.x << 32) | (uint32_t)c.y, m0, m1);
        if (c.w >= 0) top2_insert(((uint64_t)(uint32_t)c.z << 32) | (uint32_t)c.w, m0, m1);
    }
    write_result(m0, m1, row, out_idx2, out_dist2, nullptr);
}

// ---- fused exchange over peer memory
__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int4 ld_relaxed_sys_v4(const int4 *p) {
    int4 v;
    asm volatile("ld.relaxed.sys.global.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

__global__ void knn2_merge_store_peers_kernel(const uint2 *__restrict__ partial, uint32_t nA, uint32_t n_chunks,
                                              uint64_t slot_stride, uint32_t rows_per_chunk, uint32_t row_base,
                                              const PeerExchange px) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row < nA) {
        uint64_t m0 = kNone64, m1 = kNone64;
        for (uint32_t c = 0; c < n_chunks; ++c) {
            const uint2 k = partial[(uint64_t)c * slot_stride + row];
            const uint32_t base = row_base + c * rows_per_chunk;
            if (k.x != kKeyNone)
                top2_insert(((uint64_t)(k.x >> kKeyIdxBits) << 32) | (uint64_t)(base + (k.x & kKeyIdxMask)), m0, m1);
            if (k.y != kKeyNone)
                top2_insert(((uint64_t)(k.y >> kKeyIdxBits) << 32) | (uint64_t)(base + (k.y & kKeyIdxMask)), m0, m1);
        }
        const int4 rec = make_int4(m0 == kNone64 ? INT_MAX : (int32_t)(m0 >> 32), m0 == kNone64 ? -1 : (int32_t)(uint32_t)m0,
                                   m1 == kNone64 ? INT_MAX : (int32_t)(m1 >> 32), m1 == kNone64 ? -1 : (int32_t)(uint32_t)m1);
        const size_t slot = ((size_t)(px.seq & 1u) * px.world + px.rank) * px.cap + row;
        for (int g = 0; g < px.world; ++g) px.records[g][slot] = rec;      // 16-byte stores, peers over NVLink
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(px.done_counter, 1u) + 1u;
        if (done == gridDim.x) {               // last block: every record of this rank is visible
            *px.done_counter = 0u;
            __threadfence_system();
            for (int g = 0; g < px.world; ++g)
                st_release_sys(px.flags[g] + (size_t)(px.seq & 1u) * px.world + px.rank, px.seq);
        }
    }
}

// The view-sharded query's exchange on the same buffers: this rank's surviving matches go from
// the compaction's output arrays straight into slot [seq&1][rank] of every rank's buffer as the
// block {n_matches, counts[max_nv], (i, j, d0)[min(n_matches, slots)]} of 32-bit words.
__global__ void query_matches_store_peers_kernel(const uint64_t *__restrict__ d_total,
                                                 const uint64_t *__restrict__ d_seg_out, uint32_t nv_local,
                                                 uint32_t max_nv, uint32_t slots, const uint32_t *__restrict__ o_i,
                                                 const uint32_t *__restrict__ o_j, const int32_t *__restrict__ o_d,
                                                 const PeerExchange px) {
    const uint64_t total = d_total ? *d_total : 0ull;
    const uint32_t n_fit = (uint32_t)(total < (uint64_t)slots ? total : (uint64_t)slots);
    const uint32_t n_words = 1u + max_nv + 3u * n_fit;
    const size_t slot = ((size_t)(px.seq & 1u) * px.world + px.rank) * px.cap;
    for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += gridDim.x * blockDim.x) {
        uint32_t v;
        if (w == 0) {
            v = (uint32_t)total;
        } else if (w <= max_nv) {
            const uint32_t k = w - 1u;
            v = (d_seg_out && k < nv_local) ? (uint32_t)(d_seg_out[k + 1] - d_seg_out[k]) : 0u;
        } else {
            const uint32_t r = w - 1u - max_nv, k = r / 3u, c = r - 3u * k;
            v = c == 0u ? o_i[k] : c == 1u ? o_j[k] : (uint32_t)o_d[k];
        }
        for (int g = 0; g < px.world; ++g) reinterpret_cast<uint32_t *>(px.records[g] + slot)[w] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int done = atomicAdd(px.done_counter, 1u) + 1u;
        if (done == gridDim.x) {
            *px.done_counter = 0u;
            __threadfence_system();
            for (int g = 0; g < px.world; ++g)
                st_release_sys(px.flags[g] + (size_t)(px.seq & 1u) * px.world + px.rank, px.seq);
        }
    }
}

// Wait (bounded) until every rank's block for `seq` has landed in the local buffer.
__global__ void peers_wait_kernel(const PeerExchange px) {
    const int g = threadIdx.x;
    if (g >= px.world) return;
    const uint32_t *fl = px.flags[px.rank] + (size_t)(px.seq & 1u) * px.world;
    const long long t0 = clock64();
    while (ld_acquire_sys(fl + g) != px.seq) {
        if (clock64() - t0 > 20000000000ll) { *px.status = 1u; break; }      // ~10 s
        __nanosleep(200);
    }
}

__global__ void knn2_merge_from_peers_kernel(const PeerExchange px, uint32_t nA, int32_t *out_idx2,
                                             int32_t *out_dist2) {
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        // all ranks run this step concurrently (one process per GPU); bounded spin so that a
        // dead peer surfaces as an error instead of a hang
        const uint32_t *fl = px.flags[px.rank] + (size_t)(px.seq & 1u) * px.world;
        int ok = 1;
        const long long t0 = clock64();
        for (int g = 0; g < px.world; ++g) {
            while (ld_acquire_sys(fl + g) != px.seq) {
                if (clock64() - t0 > 20000000000ll) { ok = 0; break; }      // ~10 s
                __nanosleep(200);
            }
            if (!ok) break;
        }
        if (!ok) *px.status = 1u;
        s_ok = ok;
    }
    __syncthreads();
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nA) return;
    uint64_t m0 = kNone64, m1 = kNone64;
    if (s_ok) {
        const int4 *mine = px.records[px.rank] + (size_t)(px.seq & 1u) * px.world * px.cap;
        for (int g = 0; g < px.world; ++g) {
            const int4 c = ld_relaxed_sys_v4(mine + (size_t)g * px.cap + row);
            if (c.y >= 0) top2_insert(((uint64_t)(uint32_t)c.x << 32) | (uint32_t)c.y, m0, m1);
            if (c.w >= 0) top2_insert(((uint64_t)(uint32_t)c.z << 32) | (uint32_t)c.w, m0, m1);
        }
    }
    write_result(m0, m1, row, out_idx2, out_dist2, nullptr);
}

// In-place fold of 64-byte rows: w[3i+2] ^= w[3i] ^ w[3i+1] for i = 0..4, then w15 ^= w11 ^ w14.
__global__ void knn2_fold_rows_kernel(uint4 *__restrict__ rows, size_t n) {
    const size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    uint4 *p = rows + r * 4;
    uint4 a = p[0], b = p[1], c = p[2], d = p[3];
    a.z ^= a.x ^ a.y;      // w2  ^= w0 ^ w1
    b.y ^= a.w ^ b.x;      // w5  ^= w3 ^ w4
    c.x ^= b.z ^ b.w;      // w8  ^= w6 ^ w7
    c.w ^= c.y ^ c.z;      // w11 ^= w9 ^ w10
    d.z ^= d.x ^ d.y;      // w14 ^= w12 ^ w13
    d.w ^= c.w ^ d.z;      // w15 ^= w11' ^ w14'  (= w9 ^ ... ^ w15)
    p[0] = a; p[1] = b; p[2] = c; p[3] = d;
}

template <int THREADS, int QPT, int CSA, int OPT>
cudaError_t launch_variant(const KnnParams &p, int grid, cudaStream_t stream) {
    knn2_kernel<THREADS, QPT, CSA, OPT><<<grid, THREADS, 0, stream>>>(p);
    return cudaGetLastError();
}
template <int THREADS, int QPT, int CSA, int OPT>
cudaError_t info_variant(int *regs, int *ctas, size_t *smem) {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, knn2_kernel<THREADS, QPT, CSA, OPT>);
    if (e != cudaSuccess) return e;
    if (regs) *regs = a.numRegs;
    if (smem) *smem = a.sharedSizeBytes;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, knn2_kernel<THREADS, QPT, CSA, OPT>, THREADS, 0);
    if (ctas) *ctas = n;
    return e;
}

}  // namespace

#define HULO_KNN_VARIANTS(X) \
    X(256, 8, 0, 0) X(256, 8, 5, 0) X(256, 8, 7, 0) X(256, 8, 9, 0) X(256, 8, 11, 0) \
    X(256, 8, 7, 1) X(256, 8, 7, 2) X(256, 8, 7, 3) X(256, 8, 8, 0) X(256, 8, 8, 1) X(256, 8, 8, 3) \
    X(512, 4, 7, 0) X(512, 4, 7, 1) X(512, 4, 7, 3) X(512, 4, 8, 1) X(512, 4, 8, 3) X(512, 4, 9, 0) \
    X(256, 4, 7, 0) X(256, 4, 7, 3) X(256, 4, 8, 3) X(128, 8, 7, 0) X(128, 4, 7, 0) X(128, 4, 7, 3) \
    X(768, 2, 8, 1) X(768, 2, 7, 1) X(768, 2, 8, 0) X(640, 3, 8, 1) X(640, 3, 7, 1) X(640, 3, 8, 0) X(384, 5, 8, 1) \
    X(512, 4, 89, 1) X(512, 4, 9, 1) X(256, 8, 89, 1) X(256, 8, 9, 1) X(256, 4, 89, 1) X(256, 4, 8, 1) X(128, 4, 8, 1) \
    X(512, 4, 889, 1) X(256, 4, 889, 1) X(256, 4, 9, 1) X(128, 4, 89, 1) X(128, 4, 889, 1) \
    X(512, 4, 9, 5) X(512, 4, 889, 5)

cudaError_t knn2_launch(const KnnParams &p, const KnnConfig &cfg, int grid_ctas, cudaStream_t stream) {
#define X(T, Q, C, O) \
    if (cfg.threads == T && cfg.qpt == Q && cfg.csa == C && cfg.opt == O) \
        return launch_variant<T, Q, C, O>(p, grid_ctas, stream);
    HULO_KNN_VARIANTS(X)
#undef X
    return cudaErrorInvalidValue;
}

static cudaError_t knn2_kernel_info_query(const KnnConfig &cfg, int *regs, int *max_ctas_per_sm, size_t *smem) {
#define X(T, Q, C, O) \
    if (cfg.threads == T && cfg.qpt == Q && cfg.csa == C && cfg.opt == O) \
        return info_variant<T, Q, C, O>(regs, max_ctas_per_sm, smem);
    HULO_KNN_VARIANTS(X)
#undef X
    return cudaErrorInvalidValue;
}

// The attributes of a variant do not change: asked once per (thread, device, variant) -- the
// per-query entry points call this on every request.
cudaError_t knn2_kernel_info(const KnnConfig &cfg, int *regs, int *max_ctas_per_sm, size_t *smem) {
    struct Entry { KnnConfig cfg; int device, regs, ctas; size_t smem; };
    static thread_local Entry cache[8];
    static thread_local int n_cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    for (int k = 0; k < n_cached; ++k) {
        const Entry &e = cache[k];
        if (e.device == dev && e.cfg.threads == cfg.threads && e.cfg.qpt == cfg.qpt && e.cfg.csa == cfg.csa &&
            e.cfg.opt == cfg.opt) {
            if (regs) *regs = e.regs;
            if (max_ctas_per_sm) *max_ctas_per_sm = e.ctas;
            if (smem) *smem = e.smem;
            return cudaSuccess;
        }
    }
    Entry e{cfg, dev, 0, 0, 0};
    const cudaError_t rc = knn2_kernel_info_query(cfg, &e.regs, &e.ctas, &e.smem);
    if (rc != cudaSuccess) return rc;
    if (n_cached < 8) cache[n_cached++] = e;
    if (regs) *regs = e.regs;
    if (max_ctas_per_sm) *max_ctas_per_sm = e.ctas;
    if (smem) *smem = e.smem;
    return cudaSuccess;
}

cudaError_t knn2_fold_rows_launch(uint4 *rows, size_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    const int threads = 256;
    knn2_fold_rows_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, stream>>>(rows, n);
    return cudaGetLastError();
}

void knn2_unfold_rows_host(uint8_t *rows64, size_t n) {
    for (size_t r = 0; r < n; ++r) {
        uint32_t *w = reinterpret_cast<uint32_t *>(rows64 + r * 64);
        w[15] ^= w[11] ^ w[14];                                   // while w11, w14 are still folded
        for (int i = 0; i < 5; ++i) w[3 * i + 2] ^= w[3 * i] ^ w[3 * i + 1];
    }
}

cudaError_t knn2_merge_launch(const uint2 *partial, uint32_t nA, uint32_t n_chunks, uint64_t slot_stride,
                              uint32_t rows_per_chunk, uint32_t row_base, int32_t *out_idx2,
                              int32_t *out_dist2, int4 *out_packed, cudaStream_t stream) {
    if (nA == 0) return cudaSuccess;
    const int threads = 256;
    knn2_merge_kernel<<<(nA + threads - 1) / threads, threads, 0, stream>>>(
        partial, nA, n_chunks, slot_stride, rows_per_chunk, row_base, out_idx2, out_dist2, out_packed);
    return cudaGetLastError();
}

cudaError_t knn2_merge_store_peers_launch(const uint2 *partial, uint32_t nA, uint32_t n_chunks,
                                          uint64_t slot_stride, uint32_t rows_per_chunk, uint32_t row_base,
                                          const PeerExchange &px, cudaStream_t stream) {
    const int threads = 256;
    const uint32_t blocks = nA == 0 ? 1 : (nA + threads - 1) / threads;
    knn2_merge_store_peers_kernel<<<blocks, threads, 0, stream>>>(partial, nA, n_chunks, slot_stride,
                                                                 rows_per_chunk, row_base, px);
    return cudaGetLastError();
}

cudaError_t knn2_merge_from_peers_launch(const PeerExchange &px, uint32_t nA, int32_t *out_idx2,
                                         int32_t *out_dist2, cudaStream_t stream) {
    const int threads = 256;
    const uint32_t blocks = nA == 0 ? 1 : (nA + threads - 1) / threads;
    knn2_merge_from_peers_kernel<<<blocks, threads, 0, stream>>>(px, nA, out_idx2, out_dist2);
    return cudaGetLastError();
}

cudaError_t query_matches_store_peers_launch(const uint64_t *d_total, const uint64_t *d_seg_out, uint32_t nv_local,
                                             uint32_t max_nv, uint32_t slots, const uint32_t *o_i, const uint32_t *o_j,
                                             const int32_t *o_d, const PeerExchange &px, cudaStream_t stream) {
    query_matches_store_peers_kernel<<<16, 256, 0, stream>>>(d_total, d_seg_out, nv_local, max_nv, slots, o_i, o_j, o_d, px);
    return cudaGetLastError();
}

cudaError_t peers_wait_launch(const PeerExchange &px, cudaStream_t stream) {
    peers_wait_kernel<<<1, 32, 0, stream>>>(px);
    return cudaGetLastError();
}

cudaError_t knn2_merge_ranks_launch(const int4 *gathered, uint32_t nA, int world, int32_t *out_idx2,
                                    int32_t *out_dist2, cudaStream_t stream) {
    if (nA == 0) return cudaSuccess;
    const int threads = 256;
    knn2_merge_ranks_kernel<<<(nA + threads - 1) / threads, threads, 0, stream>>>(gathered, nA, world,
                                                                                  out_idx2, out_dist2);
    return cudaGetLastError();
}

}  // namespace hulo
