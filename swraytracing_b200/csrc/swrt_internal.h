// swrt_internal.h -- shared declarations between the kernels and the C-ABI layer of libswrt.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstddef>

namespace swrt {

constexpr int kMaxPlanes = 7;          // u,v,ux,uy,vx,vy,H
constexpr int kNumSrcPlanes = 10;      // + the three psi-hat moment planes N0,N1,N2 (ids 7,8,9)
#ifndef SWRT_CONSUMER_WARPS
#define SWRT_CONSUMER_WARPS 8
#endif
constexpr int kConsumerWarps = SWRT_CONSUMER_WARPS;      // MMA warps per CTA in the spectral kernel
#ifndef SWRT_CTAS_PER_SM
#define SWRT_CTAS_PER_SM 1
#endif
constexpr int kCtasPerSm = SWRT_CTAS_PER_SM;             // independent CTAs per SM (each with its own smem ring)
constexpr int kSpecThreads = kConsumerWarps * 32;         // warp 0 lane 0 also issues the bulk copies

// ---------------------------------------------------------------------------------------------
// Packed coefficient stack ("B operand") geometry.  See DESIGN.md section 3.
//   A pass covers KYP = 4*G consecutive ky for all NPL planes: NT = NPL*G n-tiles of 8 real
//   columns (4 complex (re,im) column pairs).  One k-step covers two kx>=0 wavenumbers
//   (rows Er(kx),Ei(kx),Er(kx+1),Ei(kx+1)).  Storage order (doubles):
//       [pass][kstep][tile pair tp][lane 0..31][2]
//   so that lane reads ONE 16-byte word per pair of n-tiles and a whole chunk of KC k-steps is
//   one contiguous block that a single cp.async.bulk moves into shared memory.
// ---------------------------------------------------------------------------------------------
struct PackGeom {
    int nx, nkx, nky, kmax;
    int npl, G, NT;
    int npass;            // ceil(nky / (4G))
    int ksteps;           // k-steps per pass, padded to a multiple of kc
    int kc;               // k-steps per chunk (pipeline stage)
    int chunks_per_eval;  // npass * ksteps / kc
    int nstages;
    int lag;              // a stage is refilled `lag` chunks after warp 0 released it
    int desync_ns;        // start-up skew of warps 4..7 (0 = none)
    int atab;             // x twiddles of a step: 0 = rotated in registers inside the k-loop, 1 = tabulated in shared memory
                          //    (A fragments by LDS.64), 2 = tabulated in a per-CTA global scratch (L2), prefetched by LDG
    size_t chunk_doubles; // kc * NT * 32
    size_t total_doubles; // npass * ksteps * NT * 32
    int plane_ids[kMaxPlanes];   // which of the 7 source planes each stack plane is
};

PackGeom make_geom(int nx, int npl, const int* plane_ids, int mtiles, int twiddle_pref = 0);

// one source slot: complex planes on the device, [plane][ky][kx+kmax] as double2 (kx fastest)
void launch_pack(const PackGeom& g, const double2* const* planes_dev /*[kNumSrcPlanes] device ptrs (host array)*/,
                 double* stack_dev, cudaStream_t st);
void launch_psi_moments(const double2* psik, double2* n0, double2* n1, double2* n2, int nkx, int nky, cudaStream_t st);
void launch_psi_to_planes(const double2* psik, double2* const* planes /*host array of 6 dev ptrs*/,
                          int nkx, int nky, double kappa, double u_mean, cudaStream_t st);
void launch_axpby(double* out, const double* a, const double* b, double wa, double wb, size_t n,
                  cudaStream_t st);
// m blends in one launch: out[j*n + i] = (1 - al_j)*a[i] + al_j*b[i], al_j = alpha0 + (j0 + j)*dalpha, j < m
void launch_axpby_multi(double* out, const double* a, const double* b, double alpha0, double dalpha, int j0, int m, size_t n,
                        cudaStream_t st);

struct SpecArgs {
    const double* stack;
    PackGeom g;
    long long n;            // packets
    const double* xin; const double* yin;    // EVAL: positions
    double* x; double* y; double* k; double* l;   // LEAPFROG: state (in/out)
    double* out[kMaxPlanes];                  // EVAL: outputs per stack plane (may be null)
    double dx, nxd;         // grid spacing and nx as double
    double inv_nx;          // 1/nx when nx is a power of two (exact), else 0
    double f2, gH, dt;
    int nsteps;
    bool psi;               // stack = psi-hat moment planes; assemble the six planes in stage 2
    double kappa;           // psi mode: 2*pi/L
    double u_mean0, u_mean1;   // psi mode: mean shear of the two flow slots; evaluation j adds (1-al_j)*u_mean0 + al_j*u_mean1
    double alpha0, dalpha;  //   with al_j = alpha0 + (j0 + j)*dalpha (EVAL: j = 0)
    int j0;
    int nstack;             // LEAPFROG: `stack` holds nstack (= nsteps) consecutive stacks, one per fused step (a
                            //   time-dependent flow pre-blended at al_j for every step of the launch); 0/1 = one stack
    double* twid;           // atab == 2: per-CTA twiddle scratch, gridDim.x * 8 warps * ksteps * 32 doubles
    unsigned long long* trace;   // developer timeline buffer (only read when built with -DSWRT_TRACE)
};

enum SpecMode { SPEC_EVAL = 0, SPEC_LEAPFROG = 1 };

// fused step_packet / step_packet_xka in SPECTRAL mode (spectral_rk4_kernels.cu): two packed stacks share one chunk ring
struct SpecRk4Args {
    const double* stackA;   // stage evaluations: (u,v) or (u,v,H)
    const double* stackB;   // the big evaluation: six planes / the psi-hat moment planes (old position) or seven planes (new position)
    PackGeom gA, gB;
    int nstages, lag, atab, tab_ksteps;      // joint ring geometry (filled by spectral_rk4_geometry)
    uint32_t stage_bytes;
    long long n;
    double *x, *y, *k, *l, *a;
    double dx, nxd, inv_nx;
    double f, C0, dt;
    int nsteps;
    int nstack;             // > 1: one pre-blended pair of stacks per step (time-dependent flow), stored back to back
    bool psiB;              // stackB = psi-hat moment planes
    double* twid;           // per-CTA global twiddle scratch (atab == 2): gridDim.x * 8 warps * tab_ksteps * 32 doubles
    double kappa, u_mean0, u_mean1, alpha0, dalpha;
    int j0;
};
bool spectral_rk4_geometry(SpecRk4Args& a, size_t* smem_bytes);
cudaError_t launch_spectral_rk4(const SpecRk4Args& a, bool xka, int num_sms, cudaStream_t st);

// returns cudaError; grid is sized to the SM count (persistent CTAs)
cudaError_t launch_spectral(const SpecArgs& a, int mode, int mtiles, int num_sms, cudaStream_t st);
size_t spectral_smem_bytes(const PackGeom& g);

// ---------------------------------------------------------------------------------------------
// Lagrange (reference-semantics) kernels.  Grid planes are node-interleaved: F[ix][iy][NPL].
// ---------------------------------------------------------------------------------------------
struct LagArgs {
    const double* grid;     // [nx][nx][npl] doubles
    const double* grid2;    // second frame for the exact two-frame blend (eval / leapfrog kernels), else null
    double alpha;           // blend weight of grid2 (step j of a fused run: alpha + (j0 + j)*dalpha)
    double dalpha; int j0;
    size_t gstride;         // fused runs on pre-blended grids: step j reads grid + j*gstride (doubles); 0 = one grid
    int nx, npl;            // npl = 6 or 7 (7th = H)
    long long n;
    const double* xin; const double* yin;
    double* x; double* y; double* k; double* l; double* a;
    double* out[kMaxPlanes];
    double dx, bump;
    double f, gH, C0, dt;
    int nsteps;
};
void launch_interleave_grid(const double* const* planes_dev, int npl, int nx, double* grid, cudaStream_t st);
cudaError_t launch_lagrange_eval(const LagArgs& a, cudaStream_t st);
cudaError_t launch_lagrange_leapfrog(const LagArgs& a, cudaStream_t st);
cudaError_t launch_lagrange_rk4(const LagArgs& a, bool xka, cudaStream_t st);
cudaError_t launch_interpolate_single(const double* F, int nx, int ny, const double* x, const double* y,
                                      long long n, double dx, double dy, double bump, double* out, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// Elementwise / diagnostics kernels
// ---------------------------------------------------------------------------------------------
struct Rk4Args {   // spectral-mode RK4 glue (point-wise composition), all device SoA arrays
    long long n;
    double *x, *y, *k, *l, *a;
    double *xs, *ys;                 // stage positions
    double *ax, *ay;                 // accumulated (x1+2x2+2x3+x4)
    const double *u, *v, *H;         // evaluated at stage position
    const double *ux, *uy, *vx, *vy; // final evaluation
    double dt, f, C0;
    int stage; bool xka;
};
void launch_rk4_stage(const Rk4Args& a, cudaStream_t st);
void launch_rk4_final(const Rk4Args& a, cudaStream_t st);
void launch_rhs(long long n, const double* k, const double* l, const double* const* e6, double f, double gH, double cgfac,
                double* dxdt, double* dydt, double* dkdt, double* dldt, cudaStream_t st);
void launch_omega(long long n, const double* k, const double* l, const double* u, const double* v,
                  double f, double gH, double* omega, double* Omega_abs, cudaStream_t st);
void launch_hist(long long n, const double* w, const double* edges_dev, int nedges,
                 unsigned long long* counts_dev, cudaStream_t st);
void launch_ideal_hist(long long npts, const double* u, const double* v, const double* kvx, const double* kvy, int nang,
                       double omega0, const double* edges_dev, int nedges, unsigned long long* counts_dev, cudaStream_t st);
void launch_diag(long long n, const double* x, const double* y, const double* k, const double* l, const double* a,
                 const double* omega, const double* Omega_abs, double* out8_dev, cudaStream_t st);
void launch_fill(double* p, double v, long long n, cudaStream_t st);
double gather_probe(size_t table_bytes, int iters, int reps, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// NUFFT mode (type-2 non-uniform FFT evaluation of the same Fourier series; nufft_kernels.cu)
// ---------------------------------------------------------------------------------------------
constexpr int kNufftW = 18;             // ES-kernel width in fine-grid points (16 gives 2e-12 on the gradients at 512^2, 18 gives 1.4e-13)
constexpr double kNufftSigma = 2.0;     // oversampling: nf = 2 nx
constexpr double kNufftBeta = 2.30 * kNufftW;
struct NufftArgs {
    const double2* grid;    // (u,v) fine grid, [iy*nf + ix]
    const double* hgrid;    // optional (u,v,H,0) fine grid, 4 doubles per node, same indexing (flows that carry H)
    size_t gstride, hstride;   // fused runs on pre-blended fine grids: step j reads grid + j*gstride (double2) /
                               // hgrid + j*hstride (doubles); 0 = one grid for every step
    int nf;
    long long n;
    const double* xin; const double* yin;
    double* x; double* y; double* k; double* l; double* a;
    double* out[7];
    double dx, nxd, beta, dscale;       // dscale = -sigma / (dx * w/2): d/dx of the kernel argument
    double f2, gH, dt;
    double f, C0;                       // RK4 steppers (step_packet / step_packet_xka)
    int nsteps;
};
void launch_nufft_spread(const double2* half, int nx, int nf, const double* invphi_dev, double2* full, cudaStream_t st);
void launch_nufft_store(const double2* full, int nf, int c, int stride, double* grid2, cudaStream_t st);
cudaError_t launch_nufft_eval(const NufftArgs& a, cudaStream_t st);
cudaError_t launch_nufft_leapfrog(const NufftArgs& a, cudaStream_t st);
cudaError_t launch_nufft_rk4(const NufftArgs& a, bool xka, cudaStream_t st);

struct Bs23Args {          // ode23 work arrays, component order x,y,k,l
    long long n;
    double* y[4];          // the packet state itself
    double* yt[4];         // stage state / ynew
    double* f[4][4];       // f1..f4
};
void launch_bs23_stage(const Bs23Args& a, double hb1, double hb2, double hb3, cudaStream_t st);
void launch_bs23_norm(const Bs23Args& a, int mode, double thr, unsigned long long* out, cudaStream_t st);
void launch_bs23_accept(const Bs23Args& a, cudaStream_t st);
void launch_bs23_interp(const Bs23Args& a, const double w[4], double* const out[4], cudaStream_t st);

}  // namespace swrt
