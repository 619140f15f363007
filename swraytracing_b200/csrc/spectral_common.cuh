// spectral_common.cuh -- device helpers shared by the dense-contraction kernels (spectral_kernels.cu: evaluation and
// leapfrog; spectral_rk4_kernels.cu: the fused step_packet / step_packet_xka steppers): mbarrier / bulk-copy PTX, the
// fp64 DMMA, the reduced angle of interpolate.m:21, complex helpers and the pinned half-drift.
#pragma once
#include "swrt_internal.h"

namespace swrt {

#ifndef SWRT_KUNROLL
#define SWRT_KUNROLL 4
#endif
constexpr int kKUnroll = SWRT_KUNROLL;   // k-steps per unrolled loop body

// ------------------------------------------------------------------------------------------------
// PTX helpers (sm_100a)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// non-blocking probe of an mbarrier phase (test_wait never suspends the thread)
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// reduced angle of interpolate.m:21: xl = mod(x/dx, nx); returns t = 2*xl/nx so that theta = pi*t
// When nx is a power of two (inv_nx > 0) q/nx, floor(.)*nx and the subtraction are all exact, so
// the result equals fmod's to the last bit without fmod's loop.
__device__ __forceinline__ double reduced_turns(double x, double dx, double nxd, double inv_nx) {
    const double q = x / dx;
    if (inv_nx > 0.0) {
        const double r = q - floor(q * inv_nx) * nxd;
        return r * (2.0 * inv_nx);
    }
    double r = fmod(q, nxd);
    if (r < 0.0) r += nxd;
    return 2.0 * r / nxd;
}

struct Cplx { double re, im; };
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) {
    Cplx r;
    r.re = fma(a.re, b.re, -a.im * b.im);
    r.im = fma(a.re, b.im, a.im * b.re);
    return r;
}
// dt/2 * gH*k/omega(k) with every rounding pinned (explicit fma / mul), so that the value is the same
// doubles wherever it is evaluated (tile start, after a kick, fused or single-step launches)
__device__ __forceinline__ void half_drift(double k, double l, double f2, double gH, double hg, double& hx, double& hy) {
    const double K2 = fma(k, k, __dmul_rn(l, l));
    const double s = __dmul_rn(hg, rsqrt(fma(gH, K2, f2)));
    hx = __dmul_rn(s, k);
    hy = __dmul_rn(s, l);
}
// e^N by binary exponentiation, N a compile-time constant
template <int N>
__device__ __forceinline__ Cplx cpow(Cplx e) {
    if constexpr (N == 1) return e;
    else if constexpr (N % 2 == 0) { Cplx h = cpow<N / 2>(e); return cmul(h, h); }
    else return cmul(cpow<N - 1>(e), e);
}


}  // namespace swrt
