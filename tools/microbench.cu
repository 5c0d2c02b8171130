// microbench.cu -- instruction-pipe throughput probes for the K1 roofline (B200, sm_100a).
// Prints one JSON object: lanes per clock per SM for POPC, LOP3, IADD3, IMAD, VIMNMX and for the
// mixes K1 issues.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHECK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int MODE>
__global__ void probe(uint32_t *out, uint32_t seed, long long *cycles) {
    uint32_t v[ILP], w[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) { v[k] = seed + threadIdx.x * 7919u + k * 104729u; w[k] = v[k] ^ 0x9E3779B9u; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < ILP; ++k) {
            if (MODE == 0) {            // POPC only (dependent chain per k)
                asm volatile("popc.b32 %0, %0;" : "+r"(v[k]));
            } else if (MODE == 1) {     // LOP3 only
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[k]) : "r"(w[k]), "r"(seed));
            } else if (MODE == 2) {     // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(v[k]) : "r"(w[k]));
            } else if (MODE == 3) {     // IMAD
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(v[k]) : "r"(w[k]), "r"(seed));
            } else if (MODE == 4) {     // VIMNMX
                asm volatile("min.u32 %0, %0, %1;" : "+r"(v[k]) : "r"(w[k]));
                asm volatile("add.u32 %0, %0, 1;" : "+r"(w[k]));
            } else if (MODE == 5) {     // 1 POPC : 4 LOP3  (K1-like mix, csa=7..9)
                asm volatile("popc.b32 %0, %1;" : "=r"(v[k]) : "r"(w[k]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
            } else if (MODE == 6) {     // 1 POPC : 1 LOP3 (naive popcount mix)
                asm volatile("popc.b32 %0, %1;" : "=r"(v[k]) : "r"(w[k]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
            } else if (MODE == 7) {     // 1 POPC : 4 LOP3 : 1 IMAD
                asm volatile("popc.b32 %0, %1;" : "=r"(v[k]) : "r"(w[k]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(w[k]) : "r"(v[k]), "r"(seed));
            } else if (MODE == 8) {     // LOP3 + IMAD 1:1 (ALU and FMA pipes together)
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[k]) : "r"(w[k]), "r"(seed));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(w[k]) : "r"(seed), "r"(seed));
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) acc ^= v[k] ^ w[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
int run(const char *name, int ops_per_iter, int sms, uint32_t *out, long long *cyc, bool last) {
    const int threads = 1024, blocks = sms * 2;   // 64 warps per SM
    probe<MODE><<<blocks, threads>>>(out, 12345u, cyc);
    CHECK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1);
    CHECK(cudaEventSynchronize(e1));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    long long h[1024];
    CHECK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += (double)h[i]; avg /= blocks;
    // lane-ops per SM = 2 blocks * 1024 threads * ITERS * ILP * ops_per_iter ; cycles = avg per-block span
    double lane_ops = 2.0 * 1024 * ITERS * ILP * ops_per_iter;
    double total = lane_ops * sms;
    printf("  \"%s\": {\"lane_ops_per_clk_per_sm\": %.2f, \"gops_wall\": %.1f, \"ms\": %.4f}%s\n", name,
           lane_ops / avg, total / (ms * 1e6), ms, last ? "" : ",");
    return 0;
}

int main() {
    cudaDeviceProp p; CHECK(cudaGetDeviceProperties(&p, 0));
    uint32_t *out; long long *cyc;
    CHECK(cudaMalloc(&out, sizeof(uint32_t) * p.multiProcessorCount * 2 * 1024));
    CHECK(cudaMalloc(&cyc, sizeof(long long) * 1024));
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, p.multiProcessorCount, p.clockRate);
    int s = p.multiProcessorCount;
    run<0>("popc", 1, s, out, cyc, false);
    run<1>("lop3", 1, s, out, cyc, false);
    run<2>("iadd", 1, s, out, cyc, false);
    run<3>("imad", 1, s, out, cyc, false);
    run<4>("vimnmx_plus_add", 2, s, out, cyc, false);
    run<5>("mix_1popc_4lop3", 5, s, out, cyc, false);
    run<6>("mix_1popc_1lop3", 2, s, out, cyc, false);
    run<7>("mix_1popc_4lop3_1imad", 6, s, out, cyc, false);
    run<8>("mix_1lop3_1imad", 2, s, out, cyc, true);
    printf("}\n");
    return 0;
}
