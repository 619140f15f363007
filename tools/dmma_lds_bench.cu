// dmma_lds_bench.cu -- isolates the in-loop efficiency of the spectral kernel's inner loop:
// 24 accumulator tiles per warp, one LDS.128 per two DMMAs, optional twiddle DFMAs, 8 warps/SM.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// 0: DMMA only (B in regs), 1: + LDS.128 per pair, 2: + LDS + 4 DFMA twiddle per k-step,
// 3: twiddles for 4 k-steps computed in one burst (16 DFMA) every 4 k-steps, 4: Reinsch 2-op recurrence per k-step,
// 5: Reinsch burst for 4 k-steps, 6: A fragment from a shared-memory table (one LDS.64 per k-step, no DFMA in the loop),
// 7: own element only, partner's by shuffle (2 DFMA + SHFL per k-step)
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, const double* in, int iters) {
    extern __shared__ double2 sB[];
    for (int i = threadIdx.x; i < 12 * 32 * 8; i += blockDim.x) sB[i] = make_double2(in[i & 63], in[(i + 7) & 63]);
    double* sA = reinterpret_cast<double*>(sB + 12 * 32 * 8) + (threadIdx.x >> 5) * 32 * 32;     // per-warp twiddle table
    for (int i = threadIdx.x & 31; i < 32 * 32; i += 32) sA[i] = in[i & 63];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double acc[24][2];
#pragma unroll
    for (int t = 0; t < 24; t++) { acc[t][0] = 0; acc[t][1] = 0; }
    double p = in[lane], q = in[lane + 32], dc = -1e-3, ds = 2e-3;
    double2 breg = make_double2(in[lane], in[lane + 1]);
    double pv[4] = {p, p, p, p};
    double dl = 1e-4;
    for (int it = 0; it < iters; it++) {
        if (MODE == 3 || MODE == 5) {
#pragma unroll 1
            for (int s4 = 0; s4 < 8; s4 += 4) {
#pragma unroll
                for (int u = 0; u < 4; u++) {       // burst: twiddles of the next four k-steps
                    if (MODE == 3) { double np = fma(p, dc, fma(q, ds, p)), nq = fma(q, dc, fma(-p, ds, q)); p = np; q = nq; }
                    else { dl = fma(dc, p, dl); p = p + dl; }
                    pv[u] = p;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
#pragma unroll
                    for (int tp = 0; tp < 12; tp++) {
                        double2 b = sB[((s4 + u) * 12 + tp) * 32 + lane];
                        dmma(acc[2 * tp][0], acc[2 * tp][1], pv[u], b.x);
                        dmma(acc[2 * tp + 1][0], acc[2 * tp + 1][1], pv[u], b.y);
                    }
                }
            }
        } else {
#pragma unroll 2
        for (int s = 0; s < 8; s++) {
#pragma unroll
            for (int tp = 0; tp < 12; tp++) {
                double2 b = MODE >= 1 ? sB[(s * 12 + tp) * 32 + lane] : breg;
                if (MODE == 6 && tp == 0) p = sA[((it & 3) * 8 + s) * 32 + lane];
                dmma(acc[2 * tp][0], acc[2 * tp][1], p, b.x);
                dmma(acc[2 * tp + 1][0], acc[2 * tp + 1][1], p, b.y);
            }
            if (MODE == 2) {
                double np = fma(p, dc, fma(q, ds, p)), nq = fma(q, dc, fma(-p, ds, q));
                p = np; q = nq;
            }
            if (MODE == 4) { dl = fma(dc, p, dl); p = p + dl; }
            if (MODE == 7) { const double qq = __shfl_xor_sync(0xffffffffu, p, 1); p = fma(p, dc, fma(qq, ds, p)); }
        }
        }
    }
    double s = 0;
#pragma unroll
    for (int t = 0; t < 24; t++) s += acc[t][0] + acc[t][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + p + q + dl + pv[0] + pv[3];
}
template <int MODE> void run(const char* name, double* out, const double* in) {
    int iters = 400, smem = 12 * 32 * 8 * 16 + 8 * 32 * 32 * 8;
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<MODE><<<148, 256, smem>>>(out, in, 10);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); k<MODE><<<148, 256, smem>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double flops = 148.0 * 8 * iters * 8 * 24 * 512;
    printf("{\"name\":\"%s\",\"ms\":%.4f,\"dmma_tflops\":%.3f,\"err\":\"%s\"}\n", name, best, flops / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    double *in, *out; cudaMalloc(&in, 4096); cudaMalloc(&out, 148 * 256 * 8);
    double h[128]; for (int i = 0; i < 128; i++) h[i] = 1e-3 * (i % 7) - 2e-3; cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    run<0>("dmma_only_24acc", out, in); run<1>("dmma_lds128", out, in); run<2>("dmma_lds128_twiddle", out, in); run<3>("burst4_rotation", out, in); run<4>("reinsch_per_kstep", out, in); run<5>("burst4_reinsch", out, in);
    run<6>("a_table_lds64", out, in); run<7>("own_element_shfl", out, in);
    return 0;
}
