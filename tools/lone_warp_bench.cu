// lone_warp_bench.cu -- how fast does ONE warp per SM sub-partition run the dense kernel's k-loop?
// In the real kernel the two warps of a sub-partition cover each other's boundaries (step seeds, stage 2, chunk
// hand-over); while one is away the other runs alone, and ncu's sampling says the pair is then well below the DMMA issue
// rate.  This probe isolates the loop (24 accumulator tiles, A fragment from a shared-memory table, one LDS.128 of B per
// DMMA pair) with 8 warps (two per sub-partition) and with 4 (one per sub-partition), in two forms:
//   V0  the loop as the kernel writes it: ptxas picks the LDS -> DMMA distance (two pairs = 64 clk);
//   V1  the same loads and DMMAs as volatile inline PTX in an explicit software pipeline, B loaded D pairs ahead.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lone_warp_bench lone_warp_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma_v(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void lds128_v(double2& v, uint32_t addr) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
}
__device__ __forceinline__ void lds64_v(double& v, uint32_t addr) {
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
}

constexpr int KS = 8;          // k-steps per "chunk" held in shared memory
constexpr int NP = 12;         // n-tile pairs per k-step

template <int D>               // D = 0: compiler-scheduled (V0); D > 0: explicit pipeline, B loaded D pairs ahead (V1)
__global__ void __launch_bounds__(256, 1) k(double* out, const double* in, int iters) {
    extern __shared__ double2 sB[];
    for (int i = threadIdx.x; i < NP * 32 * KS; i += blockDim.x) sB[i] = make_double2(in[i & 63], in[(i + 7) & 63]);
    double* sA = reinterpret_cast<double*>(sB + NP * 32 * KS) + (threadIdx.x >> 5) * 32 * 32;
    for (int i = threadIdx.x & 31; i < 32 * 32; i += 32) sA[i] = in[i & 63];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc[24][2];
#pragma unroll
    for (int t = 0; t < 24; t++) { acc[t][0] = 0; acc[t][1] = 0; }
    double p = in[lane];
    if constexpr (D == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll 1
            for (int s0 = 0; s0 < KS; s0 += 4) {
#pragma unroll
                for (int su = 0; su < 4; su++) {
                    const int s = s0 + su;
                    p = sA[((it & 3) * 8 + s) * 32 + lane];
#pragma unroll
                    for (int tp = 0; tp < NP; tp++) {
                        double2 b = sB[(s * NP + tp) * 32 + lane];
                        dmma(acc[2 * tp][0], acc[2 * tp][1], p, b.x);
                        dmma(acc[2 * tp + 1][0], acc[2 * tp + 1][1], p, b.y);
                    }
                }
            }
        }
    } else {
        constexpr int NBUF = D + 1;
        const uint32_t bB = (uint32_t)__cvta_generic_to_shared(sB) + lane * 16;
        const uint32_t bA = (uint32_t)__cvta_generic_to_shared(sA) + lane * 8;
        for (int it = 0; it < iters; it++) {
#pragma unroll 1
            for (int s0 = 0; s0 < KS; s0 += 4) {
                const uint32_t base = bB + (uint32_t)s0 * NP * 512;
                const uint32_t abase = bA + (uint32_t)(((it & 3) * 8 + s0) * 256);
                double2 buf[NBUF];
                double av[4];
#pragma unroll
                for (int su = 0; su < 4; su++) lds64_v(av[su], abase + su * 256);
#pragma unroll
                for (int j = 0; j < D; j++) lds128_v(buf[j], base + j * 512);
#pragma unroll
                for (int i = 0; i < 4 * NP; i++) {
                    if (i + D < 4 * NP) lds128_v(buf[(i + D) % NBUF], base + (i + D) * 512);
                    const int tp = i % NP;
                    dmma_v(acc[2 * tp][0], acc[2 * tp][1], av[i / NP], buf[i % NBUF].x);
                    dmma_v(acc[2 * tp + 1][0], acc[2 * tp + 1][1], av[i / NP], buf[i % NBUF].y);
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < 24; t++) s += acc[t][0] + acc[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + p;
}

template <int D> void run(const char* name, int warps, double* out, const double* in) {
    int iters = 800, smem = NP * 32 * KS * 16 + 8 * 32 * 32 * 8;
    cudaFuncSetAttribute(k<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<D><<<148, 32 * warps, smem>>>(out, in, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<D><<<148, 32 * warps, smem>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 148.0 * warps * iters * KS * 24 * 512;
    printf("{\"name\":\"%s\",\"warps_per_sm\":%d,\"lds_distance_pairs\":%d,\"ms\":%.4f,\"dmma_tflops\":%.3f,\"err\":\"%s\"}\n", name, warps, D, best,
           flops / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 256 * 8);
    double h[128]; for (int i = 0; i < 128; i++) h[i] = 1e-3 * (i % 7) - 2e-3; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    for (int warps : {8, 4}) {
        run<0>("compiler_scheduled", warps, out, in);
        run<2>("explicit_pipeline", warps, out, in);
        run<3>("explicit_pipeline", warps, out, in);
        run<4>("explicit_pipeline", warps, out, in);
        run<6>("explicit_pipeline", warps, out, in);
        run<8>("explicit_pipeline", warps, out, in);
    }
    return 0;
}
