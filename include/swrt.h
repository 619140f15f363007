/* swrt.h -- C ABI of libswrt.so: B200-native engine for SWRaytracing's packet hot path.
 *
 * The reference (ndefilippis/SWRaytracing) is MATLAB and has no FFI; its callers are MATLAB
 * scripts that invoke a handful of functions by name.  Each entry point below replaces one of
 * those function-level interfaces (cited as file:line relative to the reference tree); the MEX
 * gateway (matlab/swrt_mex.c) and the Python ctypes mirror (swraytracing_b200/engine.py) bind
 * exactly these symbols.  See INTEGRATION.md for the reference-side stubs.
 *
 * Conventions
 *   - all arrays are caller-owned HOST pointers unless the name says _dev; fp64; MATLAB
 *     column-major; complex data as separate real/imag arrays (Octave's MEX API is
 *     non-interleaved);
 *   - every function returns 0 on success, <0 on error; swrt_last_error() gives the message;
 *     nothing throws or longjmps; no caller pointer is kept after return;
 *   - one host thread drives a handle; a handle owns swrt_params.ngpu CUDA devices (default 1) and the
 *     device-resident SoA packet state (x,y,k,l,a) plus the flow-coefficient stacks.  With ngpu > 1 the
 *     packets are sharded in contiguous index ranges over the devices device .. device+ngpu-1, the flow
 *     is replicated on each, every call below keeps its meaning for the WHOLE ensemble, and the only
 *     exchanges are NCCL all-reduces (single-process ncclCommInitAll communicator) of the u64 histogram
 *     counts, the diagnostic scalars and the ode23 error norm.  The caller is one MATLAB / Octave /
 *     Python process (qgsw_raytrace.m:121-150 runs the whole loop in one interpreter);
 *   - there is NO CPU fallback: without a CUDA device every call fails with SWRT_ERR_CUDA.
 */
#ifndef SWRT_H
#define SWRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWRT_VERSION 200

/* error codes */
#define SWRT_OK            0
#define SWRT_ERR_ARG      -1   /* bad argument (null pointer, size mismatch, bad enum) */
#define SWRT_ERR_STATE    -2   /* call order (no flow set, no packets set, wrong mode) */
#define SWRT_ERR_CUDA     -3   /* CUDA runtime / no device */
#define SWRT_ERR_ALLOC    -4   /* out of device or host memory */
#define SWRT_ERR_NCCL     -5

/* field-evaluation mode */
#define SWRT_MODE_SPECTRAL  0  /* exact Fourier-series sum (dense DMMA contraction)            */
#define SWRT_MODE_LAGRANGE6 1  /* the reference's 6x6 Lagrange stencil, interpolate.m:12-49     */
#define SWRT_MODE_NUFFT     2  /* the SAME exact Fourier series as SPECTRAL (<= 1e-12 of max|plane|), evaluated as a
                                  type-2 non-uniform FFT: oversampled cuFFT grid of u,v per frame (setup) + an 18x18
                                  kernel gather per evaluation, gradients from the kernel's analytic derivative (= the
                                  spectral derivatives of SpectralScheme.m:20-23 / grid_U.m:6-9).  Cost independent of
                                  nx.  An H plane (step_packet_xka) rides on a second fine grid.  The ux,uy,vx,vy planes
                                  handed to swrt_set_flow_planes_spectral / swrt_set_flow_grid are NOT used in this mode:
                                  the gradients are always the spectral derivatives of the u and v supplied.      */

/* integrator */
#define SWRT_SCHEME_LEAPFROG   0  /* ode_symplectic.m:13-21,33-37                               */
#define SWRT_SCHEME_RK4_PACKET 1  /* ray_trace_sw/step_packet.m:37-78                           */
#define SWRT_SCHEME_RK4_XKA    2  /* ray_trace_sw/step_packet_xka.m:38-91 (+ cg_sw.m:15-31)     */

/* swrt_params.flags */
#define SWRT_FLAG_RHS_GH    1  /* odefun group velocity gH*k/omega (SW_zero_background_raytracing.m:182-184,
                                  initialize_raytracing :134-145) instead of Cg*k/omega (qgsw_raytrace.m:262) */

/* histogram kind */
#define SWRT_HIST_INTRINSIC 0  /* omega = sqrt(f^2 + gH K^2), analysis/load_data.m:33           */
#define SWRT_HIST_ABSOLUTE  1  /* Omega = omega + U.k, symplectic_full_fourier.m:41,55          */

typedef struct swrt_handle swrt_handle;

typedef struct swrt_params {
    int32_t nx;        /* grid size (even); spectral arrays are (nx-1) x (nx/2), g2k.m:5-9      */
    int32_t mode;      /* SWRT_MODE_*                                                           */
    int32_t device;    /* CUDA device ordinal                                                   */
    int32_t flags;     /* SWRT_FLAG_* bits (0 = the qgsw_raytrace.m conventions)                  */
    double  L;         /* domain side; dx = L/nx (qgsw_raytrace.m:13-14)                        */
    double  f;         /* Coriolis parameter                                                    */
    double  gH;        /* Cg^2 = C0^2 (ode_symplectic.m:10-11)                                  */
    double  bump;      /* Lagrange "bump": 1e-13 (ray_trace_sw/interpolate.m:13) or 1e-10 (qg_flow_ray_trace/interpolate.m:13 -- the
                          copy interpolate_U.m and the QG drivers' odefun run beside -- and interpolate_par.m:13) */
    int32_t ngpu;      /* devices device .. device+ngpu-1 share the packets (0 or 1 = one device)  */
    int32_t reserved;  /* must be 0                                                             */
} swrt_params;

/* ---- lifetime ----------------------------------------------------------------------------- */
int  swrt_version(void);
int  swrt_device_count(void);
int  swrt_create(const swrt_params* p, swrt_handle** h);
int  swrt_destroy(swrt_handle* h);
/* message of the last failure on this handle (h may be NULL: last failure of swrt_create) */
const char* swrt_last_error(const swrt_handle* h);

/* ---- background flow (slots 0 and 1 = the two time frames of interpolate_U.m:5-17) --------- */
/* psi-hat in g2k layout; builds u,v,ux,uy,vx,vy = SpectralScheme.m:18-25 / grid_U.m:3-9 with
 * wavenumbers kappa*kx, kappa*ky (kappa = 2*pi/L) and adds u_mean to u (grid_U.m:11).          */
int swrt_set_flow_spectral(swrt_handle* h, int slot, const double* psik_re, const double* psik_im,
                           int nkx, int nky, double u_mean);
/* caller-supplied coefficient planes in the order u,v,ux,uy,vx,vy[,H]; nplanes = 6 or 7.
 * Plane 7 holds the coefficients of H = 1 + eta_g (raytrace_sw.m:44-45), mean included.        */
int swrt_set_flow_planes_spectral(swrt_handle* h, int slot, const double* const* planes_re,
                                  const double* const* planes_im, int nplanes, int nkx, int nky);
/* gridded planes nx x nx (column-major, x fastest) = grid_U.m:11-17 output / SpectralScheme
 * fields; H may be NULL.  In SPECTRAL mode the grids are transformed with g2k on the device.   */
int swrt_set_flow_grid(swrt_handle* h, int slot, const double* u, const double* v,
                       const double* ux, const double* uy, const double* vx, const double* vy,
                       const double* H, int nx);

/* ---- packets ------------------------------------------------------------------------------ */
int swrt_set_packets(swrt_handle* h, int64_t n, const double* x, const double* y,
                     const double* k, const double* l, const double* a /* may be NULL -> 1 */);
int swrt_get_packets(swrt_handle* h, double* x, double* y, double* k, double* l,
                     double* a /* may be NULL */);
int64_t swrt_num_packets(const swrt_handle* h);
/* devices behind the handle, and shard i's device ordinal and packet range [lo, lo+n) (any pointer may be NULL) */
int swrt_num_devices(const swrt_handle* h);
int swrt_shard_info(const swrt_handle* h, int i, int* device, int64_t* lo, int64_t* n);
/* device-resident SoA buffers (for callers that already live on the GPU, e.g. torch) */
int swrt_packets_alloc_dev(swrt_handle* h, int64_t n);
int swrt_packets_dev(swrt_handle* h, double** x, double** y, double** k, double** l, double** a);

/* ---- evaluation --------------------------------------------------------------------------- */
/* U,V,Ux,Uy,Vx,Vy at the current packet positions, flow = (1-alpha)*slot0 + alpha*slot1
 * (interpolate_U.m:19-23; SpectralScheme.U / grad_U, SpectralScheme.m:45-68).  Any output
 * pointer may be NULL.  If slot 1 is unset alpha must be 0.                                    */
int swrt_eval(swrt_handle* h, double alpha, double* U, double* V, double* Ux, double* Uy,
              double* Vx, double* Vy);
/* the same at caller-given positions (n host doubles each); does not touch the packet state.  */
int swrt_eval_at(swrt_handle* h, double alpha, int64_t n, const double* x, const double* y,
                 double* U, double* V, double* Ux, double* Uy, double* Vx, double* Vy, double* H);
/* odefun of qgsw_raytrace.m:259-265: dx/dt = U + Cg k/omega, dk/dt = -(grad U)^T k             */
int swrt_rhs(swrt_handle* h, double alpha, double* dxdt, double* dydt, double* dkdt, double* dldt);
/* standalone FI = interpolate(x,y,F,dx,dy) (interpolate.m:1-50): one nx x ny grid, n points.   */
int swrt_interpolate(int device, const double* x, const double* y, int64_t n, const double* F,
                     int nx, int ny, double dx, double dy, double bump, double* FI);

/* ---- stepping ----------------------------------------------------------------------------- */
/* nsteps fused steps of the chosen scheme; step j evaluates the flow at alpha0 + j*dalpha.    */
int swrt_step(swrt_handle* h, int scheme, double dt, int nsteps, double alpha0, double dalpha);
/* the same launches without the host wait (swrt_step returns when the kernels are done; this returns when they are
 * queued): host work of the previous diagnostic interval overlaps the kernel.  swrt_synchronize completes it.   */
int swrt_step_async(swrt_handle* h, int scheme, double dt, int nsteps, double alpha0, double dalpha);
/* The reference's integrators take HOST arrays and return HOST arrays (ode_symplectic.m:1-4: x0,k0 in, x,k out;
 * step_packet.m:1 / step_packet_xka.m:1: P in, Pout out).  swrt_step_host is that call shape in one entry point:
 * swrt_set_packets(n, x,y,k,l,a) + swrt_step(scheme, dt, nsteps, alpha0, dalpha) + swrt_get_packets(xo,yo,ko,lo,ao),
 * with identical results, but pipelined over packet chunks: the buffers may be ordinary pageable memory, they are
 * staged through an internal pinned ring by helper threads, and chunk i computes while chunk i+1 uploads and chunk
 * i-1 downloads.  Output pointers may alias the inputs; a / ao may be NULL (a = 1).  The device keeps the final
 * packets, as after swrt_set_packets + swrt_step.                                                            */
int swrt_step_host(swrt_handle* h, int scheme, double dt, int nsteps, double alpha0, double dalpha, int64_t n,
                   const double* x, const double* y, const double* k, const double* l, const double* a,
                   double* xo, double* yo, double* ko, double* lo, double* ao);

/* ---- ode23 (Bogacki-Shampine 3(2), MATLAB builtin called by qgsw_raytrace.m:149 and
 * qg2layersw_raytrace.m:195 on y = [x;y;k;l]) -- device building blocks.  The HOST owns the step-size
 * controller (swraytracing_b200/reference_api.py: ode23) so that a multi-GPU host can MAX-all-reduce
 * the error norm, which in the reference couples all packets.  alpha = t/tmax per stage.
 *   begin:   f1 = odefun(alpha, y); rh = max_i |f1_i| / max(|y_i|, threshold)
 *   attempt: stages 2,3, ynew, f4; err = max_i |(f*E)_i| / max(|y_i|, |ynew_i|, threshold)
 *   accept:  y <- ynew, f1 <- f4 (first-same-as-last)                                           */
int swrt_bs23_begin(swrt_handle* h, double alpha, double threshold, double* rh_norm);
int swrt_bs23_attempt(swrt_handle* h, double hstep, const double alpha[3], double threshold, double* err_norm);
int swrt_bs23_accept(swrt_handle* h);
/* dense output of the LAST attempted step (MATLAB ntrp23, used by ode23 when tspan lists output times:
 * SW_zero_background_raytracing.m:73-78): y(t + s*hstep) = y + hstep * [f1 f2 f3 f4] * BI * [s; s^2; s^3],
 * BI = [1 -4/3 5/9; 0 1 -2/3; 0 4/3 -8/9; 0 -1 1].  Call between attempt and accept; s in (0, 1].     */
int swrt_bs23_interp(swrt_handle* h, double hstep, double s, double* x, double* y, double* k, double* l);

/* ---- diagnostics -------------------------------------------------------------------------- */
/* histcounts(omega, edges) (analysis/load_data.m:39-47): counts[nedges-1], bin i = [e_i,e_i+1),
 * last bin closed; accumulate != 0 adds to counts instead of overwriting.                      */
int swrt_hist_omega(swrt_handle* h, int kind, double alpha, const double* edges, int nedges,
                    uint64_t* counts, int accumulate);
/* the same, but the counts stay on the device (u64[nedges-1], owned by the handle, valid until the
 * next histogram call) so that a multi-GPU host can all-reduce them in place (NCCL) without a
 * round trip through host memory; the handle's stream is synchronised on return              */
int swrt_hist_omega_dev(swrt_handle* h, int kind, double alpha, const double* edges, int nedges,
                        uint64_t** counts_dev);
/* pipelined form for multi-GPU diagnostics: _launch queues the histogram kernel behind the work already on the
 * handle's stream and returns at once; _wait blocks until THAT kernel has finished (not later work), after which
 * counts_dev may be snapshotted and all-reduced while the next interval's packet kernel runs                  */
int swrt_hist_omega_launch(swrt_handle* h, int kind, double alpha, const double* edges, int nedges,
                           uint64_t** counts_dev);
int swrt_hist_omega_wait(swrt_handle* h);
/* theoretical omega pdf of ideal_omega_distribution.m:3-11: omega_abs = omega0 + U(x_i).kvec_j over npts
 * points (the caller's grid XX(:),YY(:)) x nangles wavevectors (kvx,kvy = k_0*[cos(t) sin(t)]), counted
 * into histcounts-style bins; counts[nedges-1].                                                   */
int swrt_ideal_omega_hist(swrt_handle* h, double alpha, int64_t npts, const double* x, const double* y,
                          const double* kvx, const double* kvy, int nangles, double omega0,
                          const double* edges, int nedges, uint64_t* counts);
/* out[0]=sum omega, out[1]=sum Omega (omega+U.k), out[2]=max omega, out[3]=min omega,
 * out[4]=number of non-finite packets, out[5]=sum a, out[6]=n, out[7]=sum omega*a              */
int swrt_diag(swrt_handle* h, double alpha, double out[8]);
/* per-packet intrinsic and absolute frequency (symplectic_full_fourier.m:41,54-56)             */
int swrt_omega(swrt_handle* h, double alpha, double* omega, double* Omega_abs);

/* ---- spectral <-> grid kit on the device (setup, not hot path) ----------------------------- */
/* fk = g2k(fg) (g2k.m:5-9) and fg = k2g(fk) (k2g.m:5-6, fulspec.m:10-19), column-major.         */
int swrt_g2k(int device, const double* fg, int nx, double* fk_re, double* fk_im);
int swrt_k2g(int device, const double* fk_re, const double* fk_im, int nx, double* fg);

/* ---- on-device one-layer QG frame producer (setup for time-evolving runs; not the hot path) --
 * qgsw_raytrace.m:111-137 (AB3 loop), :222-230 (filter), :216-220 (forcing), :270-286 (update()).
 * swrt_set_flow_from_qg fills a flow slot with psi = -q/(K_d2+K2) (grid_U.m:2) without leaving
 * the device.                                                                                   */
typedef struct swrt_qg swrt_qg;
int swrt_qg_create(int device, int nx, double L, double K_d2, double beta, double r_drag,
                   double force_strength, double f, double Cg, double dt,
                   const double* qk_re, const double* qk_im, swrt_qg** out);
int swrt_qg_step(swrt_qg* q, int nsteps);
int swrt_qg_get(swrt_qg* q, double* qk_re, double* qk_im);
int swrt_qg_get_grid(swrt_qg* q, double* qgrid);   /* q = k2g(qk), nx x nx column-major (the pv frame of :165-170) */
int swrt_qg_destroy(swrt_qg* q);
int swrt_set_flow_from_qg(swrt_handle* h, int slot, swrt_qg* q, double u_mean);

/* ---- on-device TWO-layer QG frame producer (qg2layersw_raytrace.m:120-181, update() :309-323) ------------
 * q1,q2: layer PV spectra (g2k layout).  The inversion matrix B (:137-143), the linear operator factor_L
 * (:146-150: shear advection + hyperdiffusion nu*K^(2 alpha) + drag r + beta) and expm(factor_L*dt) live on
 * the device; the CFL logic that picks dt (:156-165) is host control flow built on swrt_qg2_max_speed.     */
typedef struct swrt_qg2 swrt_qg2;
int swrt_qg2_create(int device, int nx, double L, double K_d2, double beta, double shear_strength, double r, double nu,
                    double alpha, const double* q1_re, const double* q1_im, const double* q2_re, const double* q2_im,
                    swrt_qg2** out);
int swrt_qg2_max_speed(swrt_qg2* q, double* U0);   /* sqrt(max(u.^2+v.^2)) of grid_U over both layers (:155-157) */
int swrt_qg2_step(swrt_qg2* q, double dt);         /* one AB1/AB2/AB3 + integrating-factor step (:166-181)       */
int swrt_qg2_get(swrt_qg2* q, int layer, double* qk_re, double* qk_im);
int swrt_qg2_destroy(swrt_qg2* q);
/* flow slot <- grid_U(qk(:,:,1), ..., shear_strength): top layer, one-layer inversion, mean shear (:187-188)  */
int swrt_set_flow_from_qg2(swrt_handle* h, int slot, swrt_qg2* q);

/* ---- instrumentation ---------------------------------------------------------------------- */
/* number of kernel launches issued by this handle since creation / since the last reset        */
int64_t swrt_launch_count(swrt_handle* h, int reset);
/* milliseconds spent in the dominant kernel of the most recent swrt_step / swrt_eval call,
 * measured with CUDA events on the handle's stream; nlaunch receives how many launches         */
double  swrt_last_kernel_ms(swrt_handle* h, int* nlaunch);
/* executed real fp64 flops (spectral) or gathered bytes (Lagrange) per packet per evaluation   */
double  swrt_work_per_eval(const swrt_handle* h, int nplanes);
int     swrt_synchronize(swrt_handle* h);
/* run on a caller-owned CUDA stream (a cudaStream_t passed as void*; NULL = the handle's own) so
 * that the caller's events (e.g. torch.cuda.Event on torch's current stream) bracket the work   */
int     swrt_set_stream(swrt_handle* h, void* cuda_stream);
/* device-side stopwatch on the handle's stream: start records an event, stop records another,
 * waits for it and returns the elapsed milliseconds (<0 on error)                               */
int     swrt_timer_start(swrt_handle* h);
double  swrt_timer_stop(swrt_handle* h);
/* tuning knobs: 1 or 2 m-tiles per warp in the spectral kernel (0 = automatic); flags bits 3-4 = where the dense
 * kernel keeps the x twiddles of a step (0 automatic, 1 rotate in registers, 2 global / L2 table, 3 shared-memory
 * table with shrunk chunks else global); flags bit 0 = do
 * not use the psi-hat moment contraction (always contract the six planes); flags bit 1 (LAGRANGE6) =
 * blend two flow frames on the grid before the gather (half the gathers) instead of interpolating both
 * frames and blending the results as interpolate_U.m:19-23 does (the default, bit-faithful); flags
 * bit 2 (SPECTRAL, NUFFT) = run step_packet / step_packet_xka as separate evaluation + stage launches (the
 * composed route: 4-5 evaluation kernels + 5 point-wise kernels per step, state and planes through
 * HBM) instead of the fused kernel (one launch per run of steps, packet in registers)            */
int     swrt_set_tuning(swrt_handle* h, int mtiles, int flags);
/* planes the spectral kernel contracts for a six-plane evaluation: 3 when every flow slot was
 * given as psi-hat (moments N0,N1,N2; 6 nx^2 flops), else 6 (12 nx^2 flops); 0 in LAGRANGE6    */
int     swrt_contracted_planes(const swrt_handle* h);
/* diagnostic, host-only (no device needed): the launch geometry the dense kernel would use for an
 * nx^2 grid contracting `nplanes` planes with `mtiles` m-tiles per warp.  out[0..9] = n-tiles per
 * pass, ky passes, k-steps per pass (padded), k-steps per chunk, ring stages, chunk bytes, x twiddles
 * (0 = rotated in registers, 1 = per-step table in shared memory, 2 = per-step table in a per-CTA
 * global / L2 scratch), shared-memory twiddle-table bytes, dynamic shared memory bytes per CTA,
 * packed-stack bytes per flow slot.  Returns SWRT_OK or SWRT_ERR_ARG.                             */
int     swrt_spectral_geometry(int nx, int nplanes, int mtiles, int64_t out[10]);
/* roofline probe for the gather-bound modes (LAGRANGE6, NUFFT): the rate (GB/s of useful bytes) at which the device serves
 * scattered 64-byte segments of an L2-resident table of `table_bytes` (power of two, e.g. 16 MiB = the 512^2 NUFFT fine
 * grid) to quads of lanes, eight loads in flight per lane; best of `reps`.  Returns SWRT_OK or an error code.      */
int     swrt_gather_probe(int device, int64_t table_bytes, int reps, double* gbytes_per_s);

#ifdef __cplusplus
}
#endif
#endif /* SWRT_H */
