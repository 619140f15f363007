// nufft_kernels.cu -- NUFFT mode of libswrt: the SAME exact Fourier-series semantics as SPECTRAL mode
// (F(x,y) = sum over the half-plane spectrum), evaluated as a type-2 non-uniform FFT instead of a
// dense sum:
//   setup (per flow frame, untimed):  u-hat, v-hat  ->  divide by the kernel's Fourier transform (deconvolution)
//       ->  zero-pad to an oversampled nf = 2 nx grid  ->  cuFFT inverse  ->  fine grids u_f, v_f (interleaved);
//   per evaluation:  F(x,y) = sum_{a,b=0..15} phi(z_a(x)) phi(z_b(y)) g[i0+a, j0+b],  an 18 x 18 gather with the
//       "exponential of semicircle" kernel phi(z) = exp(beta (sqrt(1 - z^2) - 1)), beta = 2.30 w, w = 18
//       (Barnett, Magland & af Klinteberg, SIAM J. Sci. Comput. 41, 2019); the four velocity gradients come from the
//       ANALYTIC derivative phi'(z) applied to the same u_f, v_f nodes, i.e. they are the spectral derivatives
//       i kx u-hat, ... of SpectralScheme.m:20-23 / grid_U.m:6-9.
// Accuracy (tests/test_gpu_parity.py): <= 1e-12 of max|plane| against the exact trig-sum oracle on all six planes
// (measured 1e-14 for u,v and 2e-13 for the gradients).  Work per evaluation is 324 nodes x 16 bytes, independent of
// nx, against 6 nx^2 flops for the dense contraction: the crossover is below nx = 64.
// Replaces: SpectralScheme.U / grad_U (SpectralScheme.m:45-68), interpolate_U.m:5-23, ode_symplectic.m:13-37.
#include "swrt_internal.h"

namespace swrt {

namespace {

constexpr int W = kNufftW;          // kernel width in fine-grid points
constexpr double HALF_W = W / 2.0;
#ifndef SWRT_NUFFT_UNROLL
#define SWRT_NUFFT_UNROLL 1
#endif
constexpr int kRowUnroll = SWRT_NUFFT_UNROLL;   // rows of the stencil per unrolled loop body (1, 4, 18: within 5 %; L1TEX-bound)
constexpr int kBlock = 64;          // one packet per thread; small blocks balance small ensembles over 148 SMs (164 regs -> 6 blocks/SM)

__device__ __forceinline__ double reduced_coord(double x, double dx, double nxd) {
    // xl = mod(x/dx, nx), the reduced coordinate of interpolate.m:21 (same as the other two modes)
    double r = fmod(x / dx, nxd);
    if (r < 0.0) r += nxd;
    return r;
}

// phi and dphi/dz of the ES kernel at z in [-1, 1]
__device__ __forceinline__ void es_kernel(double z, double beta, double& p0, double& p1) {
    const double s2 = fma(-z, z, 1.0);
    if (s2 <= 0.0) { p0 = 0.0; p1 = 0.0; return; }      // |z| = 1: phi = e^-beta ~ 1e-16 of the peak, dropped
    const double s = sqrt(s2);
    const double e = exp(beta * (s - 1.0));
    p0 = e;
    p1 = -e * beta * z / s;
}

// first fine-grid index of the stencil and t0 = i0 - xf, so that z_a = (t0 + a) / (w/2)
__device__ __forceinline__ void stencil_origin(double xl, int nf, int& base, double& t0) {
    const double xf = xl * kNufftSigma;                  // sigma = 2: exact
    const double c = ceil(xf - HALF_W);
    t0 = c - xf;
    int i0 = (int)c % nf;
    if (i0 < 0) i0 += nf;
    base = i0;
}

// u,v,ux,uy,vx,vy at one point.  grid: double2 (u,v) at [(iy*nf + ix)], x fastest (cuFFT's output order).
__device__ __forceinline__ void nufft_eval6(const double2* __restrict__ grid, int nf, double beta, double dscale,
                                            double xl, double yl, double* F) {
    int ib, jb; double tx, ty;
    stencil_origin(xl, nf, ib, tx);
    stencil_origin(yl, nf, jb, ty);
    double wx0[W], wx1[W];
#pragma unroll
    for (int a = 0; a < W; a++) es_kernel((tx + (double)a) * (1.0 / HALF_W), beta, wx0[a], wx1[a]);
    double U = 0, V = 0, Ux = 0, Uy = 0, Vx = 0, Vy = 0;
#pragma unroll (kRowUnroll)
    for (int b = 0; b < W; b++) {
        double wy0, wy1;
        es_kernel((ty + (double)b) * (1.0 / HALF_W), beta, wy0, wy1);
        int iy = jb + b; if (iy >= nf) iy -= nf;
        const double2* row = grid + (size_t)iy * nf;
        double su0 = 0, su1 = 0, sv0 = 0, sv1 = 0;
#pragma unroll
        for (int a = 0; a < W; a++) {
            int ix = ib + a; if (ix >= nf) ix -= nf;
            const double2 g = __ldg(row + ix);
            su0 = fma(wx0[a], g.x, su0); su1 = fma(wx1[a], g.x, su1);
            sv0 = fma(wx0[a], g.y, sv0); sv1 = fma(wx1[a], g.y, sv1);
        }
        U = fma(wy0, su0, U);  Ux = fma(wy0, su1, Ux); Uy = fma(wy1, su0, Uy);
        V = fma(wy0, sv0, V);  Vx = fma(wy0, sv1, Vx); Vy = fma(wy1, sv0, Vy);
    }
    F[0] = U; F[1] = V; F[2] = Ux * dscale; F[3] = Uy * dscale; F[4] = Vx * dscale; F[5] = Vy * dscale;
}

__global__ void __launch_bounds__(kBlock, 6) nufft_eval_kernel(const NufftArgs a) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n) return;
    double F[6];
    nufft_eval6(a.grid, a.nf, a.beta, a.dscale, reduced_coord(a.xin[p], a.dx, a.nxd), reduced_coord(a.yin[p], a.dx, a.nxd), F);
#pragma unroll
    for (int c = 0; c < 6; c++)
        if (a.out[c]) a.out[c][p] = F[c];
}

// ode_symplectic.m:13-21,33-37 with the six planes from the NUFFT evaluation at x1
__global__ void __launch_bounds__(kBlock, 6) nufft_leapfrog_kernel(const NufftArgs a) {
    long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n) return;
    double x = a.x[p], y = a.y[p], k = a.k[p], l = a.l[p];
    const double h = 0.5 * a.dt;
    for (int st = 0; st < a.nsteps; st++) {
        double om = sqrt(a.f2 + a.gH * (k * k + l * l));
        x = x + h * (a.gH * k / om);
        y = y + h * (a.gH * l / om);
        double F[6];
        nufft_eval6(a.grid, a.nf, a.beta, a.dscale, reduced_coord(x, a.dx, a.nxd), reduced_coord(y, a.dx, a.nxd), F);
        x = x + a.dt * F[0];
        y = y + a.dt * F[1];
        const double k0 = k, l0 = l;
        k = k0 - a.dt * (F[2] * k0 + F[4] * l0);
        l = l0 - a.dt * (F[3] * k0 + F[5] * l0);
        om = sqrt(a.f2 + a.gH * (k * k + l * l));
        x = x + h * (a.gH * k / om);
        y = y + h * (a.gH * l / om);
    }
    a.x[p] = x; a.y[p] = y; a.k[p] = k; a.l[p] = l;
}

// half-plane coefficients (g2k layout, the ky = 0 symmetrisation of fulspec.m:16 applied) -> deconvolved, zero-padded
// full spectrum of the nf x nf fine grid in FFT order: full[ky'*nf + kx'] (x fastest)
__global__ void nufft_spread_kernel(const double2* __restrict__ half, int nx, int nf, const double* __restrict__ invphi,
                                    double2* __restrict__ full) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nf * nf) return;
    const int nkx = nx - 1, kmax = nx / 2 - 1;
    const int ixp = (int)(idx % nf), iyp = (int)(idx / nf);
    const int kx = ixp <= nf / 2 ? ixp : ixp - nf;
    const int ky = iyp <= nf / 2 ? iyp : iyp - nf;
    double2 v = make_double2(0.0, 0.0);
    if (kx >= -kmax && kx <= kmax && ky >= -kmax && ky <= kmax) {
        if (ky > 0) v = half[(size_t)ky * nkx + kx + kmax];
        else if (ky < 0) { v = half[(size_t)(-ky) * nkx + (-kx) + kmax]; v.y = -v.y; }
        else if (kx >= 0) { v = half[kx + kmax]; if (kx == 0) v.y = 0.0; }
        else { v = half[-kx + kmax]; v.y = -v.y; }
        const double s = invphi[kx < 0 ? -kx : kx] * invphi[ky < 0 ? -ky : ky];
        v.x *= s; v.y *= s;
    }
    full[idx] = v;
}
// real part of the inverse transform -> component c of the interleaved (u,v) fine grid
__global__ void nufft_store_kernel(const double2* __restrict__ full, size_t n, int c, double* __restrict__ grid2) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) grid2[2 * i + c] = full[i].x;
}

}  // namespace

void launch_nufft_spread(const double2* half, int nx, int nf, const double* invphi_dev, double2* full, cudaStream_t st) {
    size_t n = (size_t)nf * nf;
    nufft_spread_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(half, nx, nf, invphi_dev, full);
}
void launch_nufft_store(const double2* full, int nf, int c, double* grid2, cudaStream_t st) {
    size_t n = (size_t)nf * nf;
    nufft_store_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(full, n, c, grid2);
}
cudaError_t launch_nufft_eval(const NufftArgs& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    nufft_eval_kernel<<<(unsigned)((a.n + kBlock - 1) / kBlock), kBlock, 0, st>>>(a);
    return cudaGetLastError();
}
cudaError_t launch_nufft_leapfrog(const NufftArgs& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    nufft_leapfrog_kernel<<<(unsigned)((a.n + kBlock - 1) / kBlock), kBlock, 0, st>>>(a);
    return cudaGetLastError();
}

}  // namespace swrt
