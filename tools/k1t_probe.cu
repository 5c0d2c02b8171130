// k1t_probe.cu -- bring-up and timing harness for K1t (knn2_tc.cu), linked against the library's
// object files.  1: raw accumulators of one 128 x 256 tile against the CPU for the descriptor
// stride candidates; 2: top-2 keys on a ragged shape against a CPU scan; 3: throughput.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "../sfmlocalization_b200/csrc/knn2_tc.cuh"

using namespace hulo;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static void fold_rows(std::vector<uint32_t> &w, size_t n) {
    for (size_t r = 0; r < n; ++r) {
        uint32_t *p = &w[r * 16];
        for (int i = 0; i < 5; ++i) p[3 * i + 2] ^= p[3 * i] ^ p[3 * i + 1];
        p[15] ^= p[11] ^ p[14];
    }
}
static int hamming(const uint32_t *a, const uint32_t *b) {
    int d = 0;
    for (int k = 0; k < 16; ++k) d += __builtin_popcount(a[k] ^ b[k]);
    return d;
}
static std::vector<uint32_t> random_rows(size_t n, uint64_t seed) {
    std::mt19937_64 g(seed);
    std::vector<uint32_t> w(n * 16);
    for (auto &x : w) x = (uint32_t)g();
    return w;
}

struct Dev {
    uint4 *rows = nullptr; uint8_t *img = nullptr; size_t n = 0;
    void make(const std::vector<uint32_t> &plain, size_t n_) {
        n = n_;
        std::vector<uint32_t> f = plain;
        fold_rows(f, n);
        CK(cudaMalloc(&rows, std::max<size_t>(n, 1) * 64));
        CK(cudaMemcpy(rows, f.data(), n * 64, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&img, knn2_tc_image_bytes(n)));
        CK(knn2_tc_expand_launch(rows, n, img, 0));
        CK(cudaDeviceSynchronize());
    }
    void free() { cudaFree(rows); cudaFree(img); }
};

int main(int argc, char **argv) {
    const size_t big_nB = argc > 1 ? strtoull(argv[1], nullptr, 10) : 2000000;
    const int reps = argc > 2 ? atoi(argv[2]) : 12;            // timed launches of section 3
    const int only_cluster = argc > 3 ? atoi(argv[3]) : 0;     // 0: sweep 1, 2, 4
    const bool timing_only = argc > 4 && atoi(argv[4]) != 0;   // skip the correctness sections
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    printf("SMs %d\n", sms);

    // ---- 1: one tile, raw dots
    if (!timing_only) {
        const size_t nA = 128, nB = 256;
        auto A = random_rows(nA, 1), B = random_rows(nB, 2);
        Dev dA, dB; dA.make(A, nA); dB.make(B, nB);
        int32_t *dots; uint2 *partial;
        CK(cudaMalloc(&dots, 128 * 256 * 4));
        CK(cudaMalloc(&partial, 128 * sizeof(uint2)));
        const uint32_t cand[2][2] = {{128, 1024}, {1024, 128}};
        for (int v = 0; v < 2; ++v) {
            CK(cudaMemset(dots, 0x7f, 128 * 256 * 4));
            TcParams p{};
            p.imgA = dA.img; p.imgB = dB.img; p.nA = nA; p.nB = nB; p.n_mtiles = 1; p.n_chunks = 1;
            p.rows_per_chunk = 256; p.slot_stride = 128; p.partial = partial; p.dbg_dots = dots;
            p.lbo = cand[v][0]; p.sbo = cand[v][1]; p.cluster = 1;
            CK(knn2_tc_launch(p, sms, 0));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant lbo=%u sbo=%u: kernel failed: %s\n", p.lbo, p.sbo, cudaGetErrorString(e)); return 3; }
            std::vector<int32_t> h(128 * 256);
            CK(cudaMemcpy(h.data(), dots, h.size() * 4, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t i = 0; i < nA; ++i)
                for (size_t j = 0; j < nB; ++j)
                    if (h[i * 256 + j] != 512 - 2 * hamming(&A[i * 16], &B[j * 16])) ++bad;
            printf("tile test lbo=%u sbo=%u: %zu / %zu mismatches;  got[0][0..3] = %d %d %d %d  want %d %d %d %d; got[1][0]=%d want %d; got[8][0]=%d want %d\n",
                   p.lbo, p.sbo, bad, nA * nB, h[0], h[1], h[2], h[3], 512 - 2 * hamming(&A[0], &B[0]),
                   512 - 2 * hamming(&A[0], &B[16]), 512 - 2 * hamming(&A[0], &B[32]), 512 - 2 * hamming(&A[0], &B[48]),
                   h[256], 512 - 2 * hamming(&A[16], &B[0]), h[8 * 256], 512 - 2 * hamming(&A[8 * 16], &B[0]));
        }
        dA.free(); dB.free(); cudaFree(dots); cudaFree(partial);
    }

    // ---- 2: ragged shape, keys against a CPU scan
    if (!timing_only) {
        const size_t nA = 300, nB = 70000;
        auto A = random_rows(nA, 3), B = random_rows(nB, 4);
        // make ties and duplicates likely: copy some rows
        for (size_t j = 100; j < 140; ++j) memcpy(&B[j * 16], &B[7 * 16], 64);
        for (size_t i = 0; i < 20; ++i) memcpy(&A[i * 16], &B[(i * 37) * 16], 64);
        Dev dA, dB; dA.make(A, nA); dB.make(B, nB);
        uint32_t mt, nc, rpc;
        knn2_tc_plan(nA, nB, sms, &mt, &nc, &rpc);
        const uint64_t stride = (nA + 31) & ~31ull;
        uint2 *partial;
        CK(cudaMalloc(&partial, (size_t)nc * stride * sizeof(uint2)));
        CK(cudaMemset(partial, 0xEE, (size_t)nc * stride * sizeof(uint2)));
      for (int cl = 1; cl <= 4; cl *= 2) {
        CK(cudaMemset(partial, 0xEE, (size_t)nc * stride * sizeof(uint2)));
        TcParams p{};
        p.imgA = dA.img; p.imgB = dB.img; p.nA = nA; p.nB = nB; p.n_mtiles = mt; p.n_chunks = nc;
        p.rows_per_chunk = rpc; p.slot_stride = stride; p.partial = partial; p.cluster = cl;
        CK(knn2_tc_launch(p, sms, 0));
        CK(cudaDeviceSynchronize());
        std::vector<uint2> h((size_t)nc * stride);
        CK(cudaMemcpy(h.data(), partial, h.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
        size_t bad = 0;
        for (size_t i = 0; i < nA; ++i) {
            uint64_t m0 = ~0ull, m1 = ~0ull;
            for (size_t j = 0; j < nB; ++j) {
                const uint64_t k = ((uint64_t)hamming(&A[i * 16], &B[j * 16]) << 32) | j;
                if (k < m0) { m1 = m0; m0 = k; } else if (k < m1) m1 = k;
            }
            uint64_t g0 = ~0ull, g1 = ~0ull;
            for (uint32_t c = 0; c < nc; ++c) {
                const uint2 k = h[(size_t)c * stride + i];
                const uint32_t ks[2] = {k.x, k.y};
                for (uint32_t kk : ks) {
                    if (kk == kKeyNone) continue;
                    const uint64_t key = ((uint64_t)(kk >> kKeyIdxBits) << 32) | (uint64_t)(c * rpc + (kk & kKeyIdxMask));
                    if (key < g0) { g1 = g0; g0 = key; } else if (key < g1) g1 = key;
                }
            }
            if (g0 != m0 || g1 != m1) {
                if (bad < 5) printf("  row %zu: got (%llu,%llu) (%llu,%llu) want (%llu,%llu) (%llu,%llu)\n", i,
                                    (unsigned long long)(g0 >> 32), (unsigned long long)(g0 & 0xffffffff),
                                    (unsigned long long)(g1 >> 32), (unsigned long long)(g1 & 0xffffffff),
                                    (unsigned long long)(m0 >> 32), (unsigned long long)(m0 & 0xffffffff),
                                    (unsigned long long)(m1 >> 32), (unsigned long long)(m1 & 0xffffffff));
                ++bad;
            }
        }
        printf("ragged test %zu x %zu cluster %d (mtiles %u chunks %u rpc %u): %zu rows wrong\n", nA, nB, cl, mt, nc, rpc, bad);
      }
        dA.free(); dB.free(); cudaFree(partial);
    }

    // ---- 3: throughput
    if (!(argc > 5 && atoi(argv[5]) == 4)) {
        const size_t nA = 4096, nB = big_nB;
        auto A = random_rows(nA, 5);
        Dev dA; dA.make(A, nA);
        // database rows generated on the device side would need another kernel: upload random rows in slabs
        uint4 *rows; uint8_t *img;
        CK(cudaMalloc(&rows, nB * 64));
        {
            const size_t slab = 1 << 20;
            std::vector<uint32_t> buf = random_rows(slab, 6);
            for (size_t off = 0; off < nB; off += slab) {
                const size_t m = std::min(slab, nB - off);
                buf[0] = (uint32_t)off;    // slabs differ a little
                CK(cudaMemcpy(rows + off * 4, buf.data(), m * 64, cudaMemcpyHostToDevice));
            }
        }
        CK(cudaMalloc(&img, knn2_tc_image_bytes(nB)));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        CK(knn2_tc_expand_launch(rows, nB, img, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("expand %zu rows: %.3f ms (%.1f GB/s written)\n", nB, ms, knn2_tc_image_bytes(nB) / ms * 1e-6);
        uint32_t mt, nc, rpc;
        knn2_tc_plan(nA, nB, sms, &mt, &nc, &rpc);
        const uint64_t stride = nA;
        uint2 *partial;
        CK(cudaMalloc(&partial, (size_t)nc * stride * sizeof(uint2)));
        TcParams p{};
        p.imgA = dA.img; p.imgB = img; p.nA = nA; p.nB = nB; p.n_mtiles = mt; p.n_chunks = nc;
        p.rows_per_chunk = rpc; p.slot_stride = stride; p.partial = partial;
        for (int rep = 0; rep < reps; ++rep) {
            p.cluster = only_cluster ? only_cluster : 1 << ((rep * 3) / reps);
            CK(cudaEventRecord(e0));
            CK(knn2_tc_launch(p, sms, 0));
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("K1t %zu x %zu cluster %d (mtiles %u chunks %u rpc %u): %.3f ms  %.1f Gdist/s\n", nA, nB, p.cluster, mt, nc, rpc, ms,
                   (double)nA * nB / ms * 1e-6);
        }
        cudaFree(rows); cudaFree(img); cudaFree(partial); dA.free();
    }

    // ---- 4: the 4-bit engine (K1t4): one tile of raw dots, a ragged shape, timing
    if (argc > 5 && atoi(argv[5]) == 4) {
        {
            const size_t nA = 128, nB = 224;
            auto A = random_rows(nA, 11), B = random_rows(nB, 12);
            std::vector<uint32_t> fa = A, fb = B;
            fold_rows(fa, nA); fold_rows(fb, nB);
            uint4 *ra, *rb; uint8_t *ia, *ib; int32_t *dots; uint2 *partial;
            CK(cudaMalloc(&ra, nA * 64)); CK(cudaMalloc(&rb, nB * 64));
            CK(cudaMemcpy(ra, fa.data(), nA * 64, cudaMemcpyHostToDevice)); CK(cudaMemcpy(rb, fb.data(), nB * 64, cudaMemcpyHostToDevice));
            CK(cudaMalloc(&ia, knn2_tc4_image_bytes(nA))); CK(cudaMalloc(&ib, knn2_tc4_image_bytes(nB)));
            CK(knn2_tc4_expand_launch(ra, nA, ia, 0)); CK(knn2_tc4_expand_launch(rb, nB, ib, 0));
            CK(cudaMalloc(&dots, 128 * 256 * 4)); CK(cudaMemset(dots, 0x7f, 128 * 256 * 4));
            CK(cudaMalloc(&partial, 128 * sizeof(uint2)));
            TcParams p{};
            p.imgA = ia; p.imgB = ib; p.nA = nA; p.nB = nB; p.n_mtiles = 1; p.n_chunks = 1; p.rows_per_chunk = 224;
            p.slot_stride = 128; p.partial = partial; p.dbg_dots = dots;
            CK(knn2_tc4_launch(p, sms, 0));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("tc4 tile: kernel failed: %s\n", cudaGetErrorString(e)); return 3; }
            std::vector<int32_t> h(128 * 256);
            CK(cudaMemcpy(h.data(), dots, h.size() * 4, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t i = 0; i < nA; ++i)
                for (size_t j = 0; j < nB; ++j)
                    if (h[i * 256 + j] != 512 - 2 * hamming(&A[i * 16], &B[j * 16])) ++bad;
            printf("tc4 tile test: %zu / %zu mismatches; got[0][0..3] = %d %d %d %d want %d %d %d %d; got[1][0]=%d want %d\n", bad,
                   nA * nB, h[0], h[1], h[2], h[3], 512 - 2 * hamming(&A[0], &B[0]), 512 - 2 * hamming(&A[0], &B[16]),
                   512 - 2 * hamming(&A[0], &B[32]), 512 - 2 * hamming(&A[0], &B[48]), h[256], 512 - 2 * hamming(&A[16], &B[0]));
        }
        {
            const size_t nA = 300, nB = 70000;
            auto A = random_rows(nA, 3), B = random_rows(nB, 4);
            for (size_t j = 100; j < 140; ++j) memcpy(&B[j * 16], &B[7 * 16], 64);
            for (size_t i = 0; i < 20; ++i) memcpy(&A[i * 16], &B[(i * 37) * 16], 64);
            std::vector<uint32_t> fa = A, fb = B;
            fold_rows(fa, nA); fold_rows(fb, nB);
            uint4 *ra, *rb; uint8_t *ia, *ib;
            CK(cudaMalloc(&ra, nA * 64)); CK(cudaMalloc(&rb, nB * 64));
            CK(cudaMemcpy(ra, fa.data(), nA * 64, cudaMemcpyHostToDevice)); CK(cudaMemcpy(rb, fb.data(), nB * 64, cudaMemcpyHostToDevice));
            CK(cudaMalloc(&ia, knn2_tc4_image_bytes(nA))); CK(cudaMalloc(&ib, knn2_tc4_image_bytes(nB)));
            CK(knn2_tc4_expand_launch(ra, nA, ia, 0)); CK(knn2_tc4_expand_launch(rb, nB, ib, 0));
            uint32_t mt, nc, rpc;
            knn2_tc4_plan(nA, nB, sms, &mt, &nc, &rpc);
            const uint64_t stride = (nA + 31) & ~31ull;
            uint2 *partial;
            CK(cudaMalloc(&partial, (size_t)nc * stride * sizeof(uint2)));
            CK(cudaMemset(partial, 0xEE, (size_t)nc * stride * sizeof(uint2)));
            TcParams p{};
            p.imgA = ia; p.imgB = ib; p.nA = nA; p.nB = nB; p.n_mtiles = mt; p.n_chunks = nc; p.rows_per_chunk = rpc;
            p.slot_stride = stride; p.partial = partial;
            CK(knn2_tc4_launch(p, sms, 0));
            CK(cudaDeviceSynchronize());
            std::vector<uint2> h((size_t)nc * stride);
            CK(cudaMemcpy(h.data(), partial, h.size() * sizeof(uint2), cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t i = 0; i < nA; ++i) {
                uint64_t m0 = ~0ull, m1 = ~0ull;
                for (size_t j = 0; j < nB; ++j) {
                    const uint64_t k = ((uint64_t)hamming(&A[i * 16], &B[j * 16]) << 32) | j;
                    if (k < m0) { m1 = m0; m0 = k; } else if (k < m1) m1 = k;
                }
                uint64_t g0 = ~0ull, g1 = ~0ull;
                for (uint32_t c = 0; c < nc; ++c) {
                    const uint2 k = h[(size_t)c * stride + i];
                    const uint32_t ks[2] = {k.x, k.y};
                    for (uint32_t kk : ks) {
                        if (kk == kKeyNone) continue;
                        const uint64_t key = ((uint64_t)(kk >> kKeyIdxBits) << 32) | (uint64_t)(c * rpc + (kk & kKeyIdxMask));
                        if (key < g0) { g1 = g0; g0 = key; } else if (key < g1) g1 = key;
                    }
                }
                if (g0 != m0 || g1 != m1) ++bad;
            }
            printf("tc4 ragged test %zu x %zu (mtiles %u chunks %u rpc %u): %zu rows wrong\n", nA, nB, mt, nc, rpc, bad);
        }
        {
            const size_t nA = 4096, nB = big_nB;
            auto A = random_rows(nA, 5);
            std::vector<uint32_t> fa = A;
            fold_rows(fa, nA);
            uint4 *ra, *rows; uint8_t *ia, *img;
            CK(cudaMalloc(&ra, nA * 64)); CK(cudaMemcpy(ra, fa.data(), nA * 64, cudaMemcpyHostToDevice));
            CK(cudaMalloc(&ia, knn2_tc4_image_bytes(nA))); CK(knn2_tc4_expand_launch(ra, nA, ia, 0));
            CK(cudaMalloc(&rows, nB * 64));
            {
                const size_t slab = 1 << 20;
                std::vector<uint32_t> buf = random_rows(slab, 6);
                for (size_t off = 0; off < nB; off += slab) {
                    const size_t m = std::min(slab, nB - off);
                    buf[0] = (uint32_t)off;
                    CK(cudaMemcpy(rows + off * 4, buf.data(), m * 64, cudaMemcpyHostToDevice));
                }
            }
            CK(cudaMalloc(&img, knn2_tc4_image_bytes(nB)));
            CK(knn2_tc4_expand_launch(rows, nB, img, 0));
            uint32_t mt, nc, rpc;
            knn2_tc4_plan(nA, nB, sms, &mt, &nc, &rpc);
            uint2 *partial;
            CK(cudaMalloc(&partial, (size_t)nc * nA * sizeof(uint2)));
            TcParams p{};
            p.imgA = ia; p.imgB = img; p.nA = nA; p.nB = nB; p.n_mtiles = mt; p.n_chunks = nc; p.rows_per_chunk = rpc;
            p.slot_stride = nA; p.partial = partial;
            cudaEvent_t e0, e1;
            CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            for (int rep = 0; rep < reps; ++rep) {
                float ms = 0;
                CK(cudaEventRecord(e0));
                CK(knn2_tc4_launch(p, sms, 0));
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                printf("K1t4 %zu x %zu (mtiles %u chunks %u rpc %u): %.3f ms  %.1f Gdist/s\n", nA, nB, mt, nc, rpc, ms,
                       (double)nA * nB / ms * 1e-6);
            }
        }
    }
    printf("done\n");
    return 0;
}
