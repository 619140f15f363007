// fp64_peak.cu -- microbenchmark: what is the fp64 ceiling of this GPU, and do the
// DMMA (tensor) and DFMA (vector) pipes overlap?  Prints one JSON object.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// NACC independent accumulator pairs per warp; ITERS iterations.
template <int NACC, int NFMA>
__global__ void __launch_bounds__(1024) mix_kernel(double* out, const double* in, int iters) {
    double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    double c[NACC > 0 ? NACC : 1][2];
    double f[NFMA > 0 ? NFMA : 1];
#pragma unroll
    for (int i = 0; i < (NACC > 0 ? NACC : 1); i++) { c[i][0] = 0; c[i][1] = 0; }
#pragma unroll
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); i++) f[i] = in[64 + i];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NACC > 0 ? NACC : 1); i++) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); i++) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC, int NFMA>
void run(const char* name, int warps_per_sm, int nsm, double* out, const double* in, bool last = false) {
    int iters = 20000;
    dim3 grid(nsm), block(32 * warps_per_sm);
    mix_kernel<NACC, NFMA><<<grid, block>>>(out, in, 100);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        mix_kernel<NACC, NFMA><<<grid, block>>>(out, in, iters);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warps = (double)nsm * warps_per_sm;
    double fl_mma = warps * iters * (double)NACC * 512.0;
    double fl_fma = warps * iters * (double)NFMA * 64.0;
    printf("  {\"name\":\"%s\",\"warps_per_sm\":%d,\"nacc\":%d,\"nfma\":%d,\"ms\":%.4f,\"dmma_tflops\":%.3f,\"dfma_tflops\":%.3f,\"total_tflops\":%.3f}%s\n",
           name, warps_per_sm, NACC, NFMA, best, fl_mma / best * 1e-9, fl_fma / best * 1e-9, (fl_mma + fl_fma) / best * 1e-9, last ? "" : ",");
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    double *in, *out; CK(cudaMalloc(&in, 4096)); CK(cudaMalloc(&out, nsm * 1024 * 8));
    double h[128]; for (int i = 0; i < 128; i++) h[i] = 1e-3 * (i % 7) - 2e-3; CK(cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice));
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"gpu\":\"%s\",\"sms\":%d,\"clock_khz\":%d,\"runs\":[\n", p.name, nsm, clk);
    run<8, 0>("dmma_only", 4, nsm, out, in);
    run<8, 0>("dmma_only", 8, nsm, out, in);
    run<8, 0>("dmma_only", 16, nsm, out, in);
    run<24, 0>("dmma_only", 8, nsm, out, in);
    run<0, 8>("dfma_only", 4, nsm, out, in);
    run<0, 8>("dfma_only", 8, nsm, out, in);
    run<0, 8>("dfma_only", 16, nsm, out, in);
    run<0, 16>("dfma_only", 16, nsm, out, in);
    run<8, 8>("mix_8mma_8fma", 8, nsm, out, in);      // 8 DMMA (4096 fl) + 8 DFMA (512 fl)
    run<8, 32>("mix_8mma_32fma", 8, nsm, out, in);    // 4096 + 2048
    run<8, 64>("mix_8mma_64fma", 8, nsm, out, in);    // 4096 + 4096 : equal flops
    run<8, 64>("mix_8mma_64fma", 16, nsm, out, in);
    run<4, 64>("mix_4mma_64fma", 8, nsm, out, in, true);
    printf("]}\n");
    return 0;
}
