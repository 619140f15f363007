// spectral_kernels.cu -- SPECTRAL mode of libswrt: evaluate the flow planes at every packet by the
// exact Fourier-series sum, as a dense real-embedded complex contraction on the fp64 tensor pipe
// (DMMA m8n8k4), fused with the leapfrog stepper of ode_symplectic.m:13-21,33-37.
//
//   F_c(p) = sum_ky w_ky Re[ e^{i ky ty_p} * G_c(p,ky) ],   G_c(p,ky) = sum_kx C_c(kx,ky) e^{i kx tx_p}
//
// Stage 1 (G) is the GEMM:  M = packets, K = kx, N = (plane, ky).  The +kx / -kx halves are folded
// at pack time (S = C(+kx)+C(-kx), D = C(+kx)-C(-kx)), so K runs over kx >= 0 only:
//       Gr += Er*Sr - Ei*Di ,   Gi += Er*Si + Ei*Dr          (E = e^{i kx tx})
// which is the real GEMM  [Er Ei] (P x 2K)  x  [[Sr Si],[-Di Dr]] (2K x 2N).  A (the twiddles) is
// generated in registers by rotation recurrence; B (the packed stack) streams L2 -> shared memory
// with cp.async.bulk + mbarrier from one producer warp; stage 2 (the ky sum) is applied to the
// accumulator fragments in registers, followed by a quad shuffle reduction.
//
// Reference behaviour replaced: SpectralScheme.U / grad_U (SpectralScheme.m:45-68) + interpolate
// (interpolate.m:12-49) evaluated spectrally, and ode_symplectic's stage loop.
#include "swrt_internal.h"
#include "spectral_common.cuh"
#include <cstdio>
#include <cstdlib>

namespace swrt {

// ------------------------------------------------------------------------------------------------
// pack / setup kernels
// ------------------------------------------------------------------------------------------------
struct PlanePtrs { const double2* p[kNumSrcPlanes]; };

__device__ __forceinline__ double2 load_coef(const double2* pl, int kx, int ky, int kmax, int nkx) {
    // half-plane value with the ky=0 conjugate symmetrisation of fulspec.m:16 applied
    if (ky == 0) {
        if (kx < 0) { double2 v = pl[(-kx + kmax)]; v.y = -v.y; return v; }
        double2 v = pl[kx + kmax];
        if (kx == 0) v.y = 0.0;
        return v;
    }
    return pl[(size_t)ky * nkx + (kx + kmax)];
}

__global__ void pack_kernel(PackGeom g, PlanePtrs src, double* __restrict__ stack) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= g.total_doubles) return;
    int which = idx & 1;
    int lane = (idx >> 1) & 31;
    size_t rest = idx >> 6;
    int half_nt = g.NT / 2;
    int tp = rest % half_nt; rest /= half_nt;
    int s = rest % g.ksteps;
    int pass = rest / g.ksteps;
    int t = 2 * tp + which;
    int c = t / g.G, gg = t % g.G;
    int krow = lane & 3, ncol = lane >> 2;
    int kx = 2 * s + (krow >> 1), part_k = krow & 1;
    int j = ncol >> 1, part_n = ncol & 1;
    int ky = pass * 4 * g.G + 4 * gg + j;
    double val = 0.0;
    if (kx <= g.kmax && ky <= g.kmax) {
        const double2* pl = src.p[g.plane_ids[c]];
        double2 cp = load_coef(pl, kx, ky, g.kmax, g.nkx);
        double2 cm = make_double2(0.0, 0.0);
        double2 S, D;
        if (kx == 0) { S = cp; D = make_double2(0.0, 0.0); }
        else {
            cm = load_coef(pl, -kx, ky, g.kmax, g.nkx);
            S = make_double2(cp.x + cm.x, cp.y + cm.y);
            D = make_double2(cp.x - cm.x, cp.y - cm.y);
        }
        double w = (ky == 0) ? 1.0 : 2.0;
        if (part_k == 0) val = part_n == 0 ? S.x : S.y;
        else             val = part_n == 0 ? -D.y : D.x;
        val *= w;
    }
    stack[idx] = val;
}

void launch_pack(const PackGeom& g, const double2* const* planes_dev, double* stack_dev, cudaStream_t st) {
    PlanePtrs pp;
    for (int i = 0; i < kNumSrcPlanes; i++) pp.p[i] = planes_dev[i];
    size_t n = g.total_doubles;
    int bs = 256;
    pack_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, st>>>(g, pp, stack_dev);
}

struct PlaneOut { double2* p[6]; };
// SpectralScheme.m:18-25 / grid_U.m:3-9 on the device: six planes from psi-hat.
__global__ void psi_to_planes_kernel(const double2* __restrict__ psik, PlaneOut out, int nkx, int nky,
                                     double kappa, double u_mean) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nkx * nky) return;
    int kmax = (nkx - 1) / 2;
    double kx = kappa * (double)((idx % nkx) - kmax);
    double ky = kappa * (double)(idx / nkx);
    double2 psi = psik[idx];
    // u = -i ky psi ; v = i kx psi
    double2 u = make_double2(ky * psi.y, -ky * psi.x);
    double2 v = make_double2(-kx * psi.y, kx * psi.x);
    // ux = i kx u, uy = i ky u, vx = i kx v, vy = i ky v
    double2 ux = make_double2(-kx * u.y, kx * u.x);
    double2 uy = make_double2(-ky * u.y, ky * u.x);
    double2 vx = make_double2(-kx * v.y, kx * v.x);
    double2 vy = make_double2(-ky * v.y, ky * v.x);
    if (idx == kmax) u.x += u_mean;   // (kx,ky) = (0,0): mean shear, grid_U.m:11
    out.p[0][idx] = u; out.p[1][idx] = v; out.p[2][idx] = ux; out.p[3][idx] = uy; out.p[4][idx] = vx; out.p[5][idx] = vy;
}

// psi-hat moment planes (spectra of REAL fields, so the ky = 0 symmetrisation of fulspec.m:16 applies):
//   N0 = psi, N1 = i kx psi, N2 = -kx^2 psi   with INTEGER kx
__global__ void psi_moments_kernel(const double2* __restrict__ psik, double2* n0, double2* n1, double2* n2, int nkx, int nky) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= nkx * nky) return;
    const int kmax = (nkx - 1) / 2;
    const double kx = (double)((idx % nkx) - kmax);
    const double2 psi = psik[idx];
    n0[idx] = psi;
    n1[idx] = make_double2(-kx * psi.y, kx * psi.x);
    n2[idx] = make_double2(-kx * kx * psi.x, -kx * kx * psi.y);
}
void launch_psi_moments(const double2* psik, double2* n0, double2* n1, double2* n2, int nkx, int nky, cudaStream_t st) {
    int n = nkx * nky;
    psi_moments_kernel<<<(n + 255) / 256, 256, 0, st>>>(psik, n0, n1, n2, nkx, nky);
}

void launch_psi_to_planes(const double2* psik, double2* const* planes, int nkx, int nky, double kappa,
                          double u_mean, cudaStream_t st) {
    PlaneOut po;
    for (int i = 0; i < 6; i++) po.p[i] = planes[i];
    int n = nkx * nky;
    psi_to_planes_kernel<<<(n + 255) / 256, 256, 0, st>>>(psik, po, nkx, nky, kappa, u_mean);
}

// out = wa*a + wb*b : the on-device frame blend of interpolate_U.m:19-23 applied to the
// coefficient stack (linear, so identical to blending the evaluated fields).
// (the expression is pinned -- one rounded product, one fma -- so that the single blend and the multi-blend kernel below
// produce the same doubles: a fused run of m time-dependent steps equals m single-step launches bit for bit)
__global__ void axpby_kernel(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b,
                             double wa, double wb, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] = fma(wa, a[i], __dmul_rn(wb, b[i]));
}
// out[j][i] = (1 - al_j)*a[i] + al_j*b[i] for the m steps of a fused run, al_j = alpha0 + (j0 + j)*dalpha (the host's
// expression for step j0 + j of swrt_step): every operand is read once for all m blends
__global__ void axpby_multi_kernel(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b,
                                   double alpha0, double dalpha, int j0, int m, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const double av = a[i], bv = b[i];
        for (int j = 0; j < m; j++) {
            const double al = __dadd_rn(alpha0, __dmul_rn((double)(j0 + j), dalpha));      // un-fused: the host's alpha0 + j*dalpha
            out[(size_t)j * n + i] = fma(__dsub_rn(1.0, al), av, __dmul_rn(al, bv));
        }
    }
}
void launch_axpby_multi(double* out, const double* a, const double* b, double alpha0, double dalpha, int j0, int m, size_t n,
                        cudaStream_t st) {
    int bs = 256;
    size_t nb = (n + bs - 1) / bs;
    if (nb > 148 * 16) nb = 148 * 16;
    axpby_multi_kernel<<<(unsigned)nb, bs, 0, st>>>(out, a, b, alpha0, dalpha, j0, m, n);
}
void launch_axpby(double* out, const double* a, const double* b, double wa, double wb, size_t n, cudaStream_t st) {
    int bs = 256;
    size_t nb = (n + bs - 1) / bs;
    if (nb > 148 * 16) nb = 148 * 16;
    axpby_kernel<<<(unsigned)nb, bs, 0, st>>>(out, a, b, wa, wb, n);
}

PackGeom make_geom(int nx, int npl, const int* plane_ids, int mtiles, int twiddle_pref) {
    PackGeom g{};
    g.nx = nx; g.nkx = nx - 1; g.nky = nx / 2; g.kmax = nx / 2 - 1;
    g.npl = npl;
    // accumulator budget: MT * NT * 2 doubles per thread = 96..112 registers-pairs -> NT*MT ~ 24
    int nt_target = 24 / mtiles;
    int G = nt_target / npl; if (G < 1) G = 1;
    if ((npl * G) & 1) G += 1;            // NT must be even (n-tiles are fetched in pairs)
    // supported instantiations (see launch_spectral): keep in sync
    g.G = G; g.NT = npl * G;
    int kyp = 4 * G;
    g.npass = (g.nky + kyp - 1) / kyp;
    // ~48 KB chunks (kc k-steps of NT*256 bytes), kc a multiple of the k-loop unroll
    int kc0 = 192 / g.NT;
    kc0 = kc0 >= 16 ? 16 : (kc0 >= 8 ? 8 : 4);
    const int ks = (g.kmax + 1 + 1) / 2;
    auto layout = [&](int kc) {
        if (kc % kKUnroll) kc = ((kc + kKUnroll - 1) / kKUnroll) * kKUnroll;
        g.kc = kc;
        g.ksteps = ((ks + kc - 1) / kc) * kc;
        g.chunks_per_eval = g.npass * (g.ksteps / kc);
        g.chunk_doubles = (size_t)kc * g.NT * 32;
        g.total_doubles = (size_t)g.npass * g.ksteps * g.NT * 32;
    };
    // Twiddle table: one double per lane per k-step per warp (the A fragments of a whole step), written and read back
    // by the same lane.  With it the k-loop holds no fp64 instruction but the DMMAs (tools/dmma_lds_bench.cu: 36.3
    // against 34.9 TFLOP/s for the loop with the 4-DFMA rotation per k-step).  It needs ksteps * 2 KB of shared memory
    // next to a ring of at least three chunks: 48 KB chunks up to nx = 128, 24 KB chunks at nx = 256 (measured there:
    // 33.0 TFLOP/s rotating in registers with 48 KB chunks, 32.0 with 24 KB chunks, 33.7 with 24 KB chunks and the table);
    // larger grids rotate in registers.
    constexpr size_t kSmemBudget = 216 * 1024;
    auto atab_fits = [&]() { return (size_t)g.ksteps * 32 * 8 * kConsumerWarps + 3 * g.chunk_doubles * 8 <= kSmemBudget; };
    // twiddle_pref: 0 = automatic (shared-memory table when it fits beside a ring of full-size chunks, else the global / L2
    // table), 1 = rotate in registers, 2 = global table wherever a table is possible, 3 = shared-memory table even with
    // shrunk chunks (the round-1 layout at 256^2), else the global one.  Measured (profiles/README.md, TFLOP/s, rotation /
    // shared / global): 128^2 31.1 / 32.2 / 31.6; 256^2 32.9 / 33.9 (24 KB chunks) / 34.2; 512^2 33.4 / does not fit / 34.8.
    bool want_atab = mtiles == 1 && kCtasPerSm == 1 && twiddle_pref != 1 && twiddle_pref != 2;
    const bool want_global = mtiles == 1 && kCtasPerSm == 1 && twiddle_pref != 1;
    const bool shrink_for_table = twiddle_pref == 3;
    int kc_forced = 0;
#ifdef SWRT_DEV_TUNING      // developer experiments only: the shipped library reads no environment variables
    if (const char* e = getenv("SWRT_ATAB")) want_atab = want_atab && atoi(e) != 0;
    if (const char* e = getenv("SWRT_KC")) { kc_forced = atoi(e); if (kc_forced < kKUnroll) kc_forced = kKUnroll; }
#endif
    if (kc_forced) {
        layout(kc_forced);
        g.atab = want_atab && atab_fits();
    } else {
        layout(kc0);
        g.atab = want_atab && atab_fits();
        if (want_atab && !g.atab && kc0 > 4 && shrink_for_table) {
            layout(4);
            g.atab = atab_fits();
            if (!g.atab) layout(kc0);
        }
    }
    if (!g.atab && want_global) g.atab = 2;       // table in the per-CTA global scratch: full-size chunks, full-size ring
    const size_t chunk_bytes = g.chunk_doubles * 8;
    const size_t ring_budget = g.atab == 1 ? (kSmemBudget - (size_t)g.ksteps * 32 * 8 * kConsumerWarps) : (size_t)(200 * 1024 / kCtasPerSm);
    g.nstages = (int)(ring_budget / chunk_bytes);
    if (g.nstages > 8) g.nstages = 8;
    if (g.nstages < 3) g.nstages = 3;
    // Warps 4..7 (the second warp of every SM sub-partition) start ~1.5 chunks after warps 0..3, so
    // that one group's non-MMA phases (chunk hand-over, per-pass stage 2, per-step twiddle seeds)
    // run under the other group's DMMAs instead of leaving the fp64 pipe idle.  A stage is
    // refilled `lag` chunks after its issuing warp left it, which must exceed the skew.
    g.lag = g.nstages - 1;
    if (g.lag > 4) g.lag = 4;
    g.desync_ns = (int)((g.lag >= 3 ? 1.5 : 0.6) * g.kc * g.NT * 16 / 1.9);
#ifdef SWRT_DEV_TUNING
    if (const char* e = getenv("SWRT_LAG")) g.lag = atoi(e);
    if (const char* e = getenv("SWRT_DESYNC_NS")) g.desync_ns = atoi(e);
    if (const char* e = getenv("SWRT_NSTAGES")) { g.nstages = atoi(e); if (g.nstages < 2) g.nstages = 2; }
#endif
    if (g.lag >= g.nstages) g.lag = g.nstages - 1;
    if (g.lag < 1) g.lag = 1;
    for (int i = 0; i < kMaxPlanes; i++) g.plane_ids[i] = i < npl ? plane_ids[i] : 0;
    return g;
}

size_t spectral_smem_bytes(const PackGeom& g) {
    // ring | full/empty barriers (padded to 128 bytes) | twiddle table
    return (size_t)g.nstages * g.chunk_doubles * 8 + 128 + (g.atab == 1 ? (size_t)g.ksteps * 32 * 8 * kConsumerWarps : 0);
}

// ------------------------------------------------------------------------------------------------
// the contraction kernel
// ------------------------------------------------------------------------------------------------
// PSI = true: the stack holds the three psi-hat moment planes N0 = psi, N1 = i kx psi, N2 = -kx^2 psi
// (integer kx); the six velocity/gradient planes of SpectralScheme.m:18-25 are assembled in stage 2:
//   u = kap ky Im[T G0], v = kap Re[T G1], ux = kap^2 ky Im[T G1], uy = kap^2 ky^2 Re[T G0],
//   vx = kap^2 Re[T G2], vy = -ux   (T = e^{i ky ty}),  which halves the DMMA work (6 nx^2 flops).
// ATAB: 0 = x twiddles rotated in registers inside the k-loop; 1 = tabulated once per step in shared memory; 2 = tabulated
// once per step in a per-CTA global scratch (L2-resident; grids whose table does not fit in shared memory), read back
// with plain loads one unrolled body ahead -- in both table forms the k-loop holds no fp64 instruction but the DMMAs.
template <int NPL, int G, int MT, int MODE, bool PSI, int ATAB>
__global__ void __launch_bounds__(kSpecThreads, kCtasPerSm) spectral_kernel(const SpecArgs a) {
    static_assert(!ATAB || MT == 1, "the twiddle table is laid out for one m-tile per warp");
    constexpr int NT = NPL * G;
    constexpr int NF = PSI ? 6 : NPL;          // planes produced per packet
    static_assert(!PSI || NPL == 3, "psi mode contracts exactly three moment planes");
    constexpr int HALF_NT = NT / 2;
    constexpr int TILE_P = kConsumerWarps * 8 * MT;
    static_assert(NT % 2 == 0, "n-tiles are fetched in pairs");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const PackGeom& g = a.g;
    const int nstages = g.nstages;
    const uint32_t chunk_bytes = (uint32_t)(g.chunk_doubles * 8);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nstages * chunk_bytes);
    uint64_t* empty_bar = full_bar + nstages;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // this lane's column of the warp's twiddle table: sA[s * 32] = A element of k-step s (written and read by this lane only)
    double* sA;
    if constexpr (ATAB == 2) sA = a.twid + ((size_t)blockIdx.x * kConsumerWarps + warp) * g.ksteps * 32 + lane;
    else sA = reinterpret_cast<double*>(smem_raw + (size_t)nstages * chunk_bytes + 128) + (size_t)warp * g.ksteps * 32 + lane;
    const long long ntiles = (a.n + TILE_P - 1) / TILE_P;
    const int nevals = (MODE == SPEC_LEAPFROG) ? a.nsteps : 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; s++) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], kConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    long long my_tiles = 0;
    if ((long long)blockIdx.x < ntiles) my_tiles = (ntiles - 1 - blockIdx.x) / gridDim.x + 1;
    const long long total_chunks = my_tiles * nevals * (long long)g.chunks_per_eval;
    // The packed stack streams through a ring of nstages smem buffers with cp.async.bulk; chunk j
    // lives in stage j % nstages.  Producer duty rotates over the warps (lane 0 of warp j % 8 issues
    // chunk j) so that no single warp paces the CTA.  Chunk j is issued when its warp STARTS chunk
    // j - D (D = nstages - lag): the stage it refills was left `lag` chunks ago by that warp, and lag
    // exceeds the deliberate skew between the two warp groups, so the empty-barrier wait inside is
    // (almost) never a stall.  No deadlock: the issue of chunk j happens before its warp waits on the
    // full barrier of chunk j - D, and only needs chunks < j - D, which were issued earlier.
    const int D = nstages - g.lag;
    // This warp issues the chunks pj = warp, warp + 8, warp + 16, ...; the ring stage, the ring iteration and the source
    // chunk of pj are carried incrementally (64-bit divisions by run-time values in the issue path cost the issuing
    // warp ~700 cycles per turn, i.e. once per step at C2).
    // a fused run on a time-dependent flow reads one pre-blended stack per step: the nstack stacks are contiguous, so the
    // source chunk index simply wraps after chunks_per_eval * nstack (every tile walks its evaluations 0..nevals-1 in order)
    const int src_period = g.chunks_per_eval * ((MODE == SPEC_LEAPFROG && a.nstack > 1) ? a.nstack : 1);
    long long pj = warp;
    int p_st = warp % nstages, p_src = warp % src_period;
    long long p_it = warp / nstages;
    auto producer_issue = [&]() {          // lane 0 only; issues chunk pj
        if (p_it > 0) mbar_wait(&empty_bar[p_st], (uint32_t)((p_it - 1) & 1));
        mbar_expect_tx(&full_bar[p_st], chunk_bytes);
        bulk_g2s(smem_raw + (size_t)p_st * chunk_bytes, a.stack + (size_t)p_src * g.chunk_doubles, chunk_bytes, &full_bar[p_st]);
    };
    auto producer_advance = [&]() {        // all lanes (uniform): pj += kConsumerWarps
        pj += kConsumerWarps;
        p_st += kConsumerWarps;
        while (p_st >= nstages) { p_st -= nstages; p_it++; }
        p_src += kConsumerWarps;
        while (p_src >= src_period) p_src -= src_period;
    };
    static_assert(kConsumerWarps >= 8, "the prologue assumes at most one chunk per warp (D < kConsumerWarps)");
    if (pj < D && pj < total_chunks) {
        if (lane == 0) producer_issue();
        producer_advance();
    }

    if (warp >= kConsumerWarps / 2 && g.desync_ns > 0 && kCtasPerSm == 1) {
        // __nanosleep returns far too early for this purpose (measured ~250 cycles for 2000 ns): spin on
        // the SM clock instead (desync_ns is converted with the nominal 1.9 GHz it was computed for)
        const long long until = clock64() + (long long)(g.desync_ns * 1.9);
        while (clock64() < until) {}
    }

    // ===== consumer warps =====
    const int quad_row = lane >> 2;       // packet row inside an m-tile
    const int jq = lane & 3;              // complex column inside an n-tile (C fragment) / k row (A fragment)
    const int kx0 = jq >> 1;              // first kx handled by this lane's A element (0 or 1)
    const bool odd = lane & 1;            // A element is Ei (odd) or Er (even)
    const int chunks_per_pass = g.ksteps / g.kc;
    const int kyp_passes = g.npass;

    int stage = 0; uint32_t phase = 0;
    long long ci = 0;        // chunks consumed so far by this warp
    bool ready = false;      // next chunk's full barrier already observed complete
#ifdef SWRT_TRACE
    int trace_n = 0;
#define SWRT_TRACE_EV(id)                                                                      \
    do {                                                                                       \
        if (a.trace && blockIdx.x == 3 && lane == 0 && trace_n < 4 * 160)                      \
            a.trace[(size_t)warp * 4 * 160 + trace_n++] = (unsigned long long)clock64();       \
    } while (0)
#else
#define SWRT_TRACE_EV(id) do {} while (0)
#endif

    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        long long prow[MT];
        double px[MT], py[MT], pk[MT], pl[MT];
#pragma unroll
        for (int mt = 0; mt < MT; mt++) {
            long long r = tile * TILE_P + (long long)warp * 8 * MT + mt * 8 + quad_row;
            prow[mt] = r;
            long long rc = r < a.n ? r : a.n - 1;
            if (MODE == SPEC_LEAPFROG) { px[mt] = a.x[rc]; py[mt] = a.y[rc]; pk[mt] = a.k[rc]; pl[mt] = a.l[rc]; }
            else { px[mt] = a.xin[rc]; py[mt] = a.yin[rc]; pk[mt] = 0; pl[mt] = 0; }
        }

        // half-drift displacement dt/2 * gH*k/omega(k): k only changes in the kick, so the value computed
        // after a kick serves the second half drift of that step AND the first half drift of the next
        double hcx[MT], hcy[MT];
        if (MODE == SPEC_LEAPFROG) {
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                half_drift(pk[mt], pl[mt], a.f2, a.gH, __dmul_rn(__dmul_rn(0.5, a.dt), a.gH), hcx[mt], hcy[mt]);
            }
        }

        for (int ev = 0; ev < nevals; ev++) {
            if (MODE == SPEC_LEAPFROG) {
                // phi1(dt/2): x += dt/2 * gH*k/omega(k)   (ode_symplectic.m:13-16,34)
#pragma unroll
                for (int mt = 0; mt < MT; mt++) { px[mt] = __dadd_rn(px[mt], hcx[mt]); py[mt] = __dadd_rn(py[mt], hcy[mt]); }
            }
            // ---- twiddle seeds -------------------------------------------------------------
            double tp_[MT], tq_[MT];            // x twiddle in (p,q) form: p = this lane's A element
            double ep0[MT], eq0[MT];            // value at the start of every pass
            double xdc[MT], xds[MT];            // rotation by e^{i 4 tx} minus one (two k-steps), sign-adjusted
            double ep1[MT], eq1[MT];            // second chain: value at k-step 1 of every pass
            double tpB[MT], tqB[MT];
            Cplx ytw[MT][G];                    // e^{i ky ty} for this lane's ky of each group
            Cplx yrot[MT];                      // e^{i 4G ty}
            double F[MT][NF];
            double kyd[MT];                     // PSI: this lane's ky of group 0 in the current pass (as double)
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                double s1, c1;
                sincospi(reduced_turns(px[mt], a.dx, a.nxd, a.inv_nx), &s1, &c1);
                double er = kx0 ? c1 : 1.0, ei = kx0 ? s1 : 0.0;
                ep0[mt] = odd ? ei : er;
                eq0[mt] = odd ? er : ei;
                xdc[mt] = -2.0 * s1 * s1;
                double ds = 2.0 * s1 * c1;
                xds[mt] = odd ? ds : -ds;
                // two interleaved recurrences (even / odd k-steps), each advancing by e^{i 4 tx}: half the dependent
                // chain length per k-step; ep1 = ep0 rotated once by e^{i 2 tx}
                ep1[mt] = fma(ep0[mt], xdc[mt], fma(eq0[mt], xds[mt], ep0[mt]));
                eq1[mt] = fma(eq0[mt], xdc[mt], fma(-ep0[mt], xds[mt], eq0[mt]));
                {
                    const double s2 = ds, c2 = fma(-2.0 * s1, s1, 1.0);     // sin 2tx, cos 2tx
                    xdc[mt] = -2.0 * s2 * s2;
                    const double ds2 = 2.0 * s2 * c2;
                    xds[mt] = odd ? ds2 : -ds2;
                }
                double sy, cy;
                sincospi(reduced_turns(py[mt], a.dx, a.nxd, a.inv_nx), &sy, &cy);
                Cplx e1{cy, sy};
                Cplx e2 = cmul(e1, e1);
                Cplx e3 = cmul(e2, e1);
                Cplx e4 = cmul(e2, e2);
                Cplx cur = jq == 0 ? Cplx{1.0, 0.0} : (jq == 1 ? e1 : (jq == 2 ? e2 : e3));
#pragma unroll
                for (int gg = 0; gg < G; gg++) {
                    ytw[mt][gg] = cur;
                    cur = cmul(cur, e4);
                }
                yrot[mt] = cpow<G>(e4);
#pragma unroll
                for (int c = 0; c < NF; c++) F[mt][c] = 0.0;
                kyd[mt] = (double)jq;
            }

            if constexpr (ATAB) {
                // the same two recurrences the in-loop rotation runs, tabulated once per step for all passes
                double ap = ep0[0], aq = eq0[0], bp = ep1[0], bq = eq1[0];
                for (int s = 0; s < g.ksteps; s += 2) {
                    sA[s * 32] = ap; sA[(s + 1) * 32] = bp;
                    const double nap = fma(ap, xdc[0], fma(aq, xds[0], ap)), naq = fma(aq, xdc[0], fma(-ap, xds[0], aq));
                    const double nbp = fma(bp, xdc[0], fma(bq, xds[0], bp)), nbq = fma(bq, xdc[0], fma(-bp, xds[0], bq));
                    ap = nap; aq = naq; bp = nbp; bq = nbq;
                }
            }
            double apf[kKUnroll];                  // ATAB 2: the A elements of the next unrolled body, in flight from L2
            if constexpr (ATAB == 2) {
#pragma unroll
                for (int su = 0; su < kKUnroll; su++) apf[su] = sA[su * 32];
            }
            // ---- passes over ky blocks -----------------------------------------------------
            for (int pass = 0; pass < kyp_passes; pass++) {
                double acc[MT][NT][2];
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                    for (int t = 0; t < NT; t++) { acc[mt][t][0] = 0.0; acc[mt][t][1] = 0.0; }
                    if constexpr (!ATAB) {
                        tp_[mt] = ep0[mt]; tq_[mt] = eq0[mt];
                        tpB[mt] = ep1[mt]; tqB[mt] = eq1[mt];
                    }
                }
                for (int ch = 0; ch < chunks_per_pass; ch++) {
                    SWRT_TRACE_EV(0);
                    {   // producer duty for chunk ci + D (one warp in eight, lane 0 only)
                        if (ci + D == pj && pj < total_chunks) {
                            if (lane == 0) producer_issue();
                            producer_advance();
                        }
                    }
                    SWRT_TRACE_EV(1);
                    if (!ready) mbar_wait(&full_bar[stage], phase);
                    SWRT_TRACE_EV(2);
                    const double2* sB = reinterpret_cast<const double2*>(smem_raw + (size_t)stage * chunk_bytes) + lane;
                    const double* sAc = sA + (size_t)ch * g.kc * 32;
                    int nstage = stage + 1; uint32_t nphase = phase;
                    if (nstage == nstages) { nstage = 0; nphase ^= 1; }
                    for (int s0 = 0; s0 < g.kc; s0 += kKUnroll) {
                        // probe the NEXT chunk's full barrier while this chunk's last k-steps run, so the
                        // hand-over at the chunk boundary does not wait on the barrier round trip
                        if (s0 + kKUnroll >= g.kc) ready = mbar_test(&full_bar[nstage], nphase);
                        double acur[kKUnroll];
                        if constexpr (ATAB == 2) {
                            int nb = ch * g.kc + s0 + kKUnroll;          // next body of this pass; the table serves every pass
                            if (nb >= g.ksteps) nb = 0;
#pragma unroll
                            for (int su = 0; su < kKUnroll; su++) { acur[su] = apf[su]; apf[su] = sA[(nb + su) * 32]; }
                        }
#pragma unroll
                        for (int su = 0; su < kKUnroll; su++) {
                            const int s = s0 + su;
                            if constexpr (ATAB == 1) tp_[0] = sAc[s * 32];
                            if constexpr (ATAB == 2) tp_[0] = acur[su];
#pragma unroll
                            for (int tp = 0; tp < HALF_NT; tp++) {
                                double2 b = sB[(s * HALF_NT + tp) * 32];
#pragma unroll
                                for (int mt = 0; mt < MT; mt++) {
                                    dmma884(acc[mt][2 * tp][0], acc[mt][2 * tp][1], tp_[mt], b.x);
                                    dmma884(acc[mt][2 * tp + 1][0], acc[mt][2 * tp + 1][1], tp_[mt], b.y);
                                }
                            }
#pragma unroll
                            for (int mt = 0; mt < (ATAB ? 0 : MT); mt++) {
                                // this chain's next value (two k-steps ahead): E <- E + E*(e^{i 4 tx} - 1); then swap chains
                                double np = fma(tp_[mt], xdc[mt], fma(tq_[mt], xds[mt], tp_[mt]));
                                double nq = fma(tq_[mt], xdc[mt], fma(-tp_[mt], xds[mt], tq_[mt]));
                                tp_[mt] = tpB[mt]; tq_[mt] = tqB[mt];
                                tpB[mt] = np; tqB[mt] = nq;
                            }
                        }
                    }
                    SWRT_TRACE_EV(3);
                    // every lane's LDS of this stage has been consumed by the DMMAs above (mma.sync is
                    // warp-convergent), so lane 0 may hand the stage back
                    if (lane == 0) mbar_arrive(&empty_bar[stage]);
                    stage = nstage; phase = nphase;
                    ++ci;
                }
                // ---- stage 2: F_c += Gr*cos(ky ty) - Gi*sin(ky ty) for this lane's ky ----------
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                    for (int gg = 0; gg < G; gg++) {
                        const double cy = ytw[mt][gg].re, sy = ytw[mt][gg].im;
                        if constexpr (PSI) {
                            const double ky = kyd[mt] + (double)(4 * gg);
                            const double g0r = acc[mt][0 * G + gg][0], g0i = acc[mt][0 * G + gg][1];
                            const double g1r = acc[mt][1 * G + gg][0], g1i = acc[mt][1 * G + gg][1];
                            const double g2r = acc[mt][2 * G + gg][0], g2i = acc[mt][2 * G + gg][1];
                            const double a0 = fma(g0r, cy, -g0i * sy), b0 = fma(g0i, cy, g0r * sy);
                            const double a1 = fma(g1r, cy, -g1i * sy), b1 = fma(g1i, cy, g1r * sy);
                            const double a2 = fma(g2r, cy, -g2i * sy);
                            F[mt][0] = fma(ky, b0, F[mt][0]);          // u  / kap
                            F[mt][1] += a1;                            // v  / kap
                            F[mt][2] = fma(ky, b1, F[mt][2]);          // ux / kap^2
                            F[mt][3] = fma(ky * ky, a0, F[mt][3]);     // uy / kap^2
                            F[mt][4] += a2;                            // vx / kap^2
                        } else {
#pragma unroll
                            for (int c = 0; c < NPL; c++) {
                                F[mt][c] = fma(acc[mt][c * G + gg][0], cy, F[mt][c]);
                                F[mt][c] = fma(-acc[mt][c * G + gg][1], sy, F[mt][c]);
                            }
                        }
                        ytw[mt][gg] = cmul(ytw[mt][gg], yrot[mt]);
                    }
                    kyd[mt] += (double)(4 * G);
                }
            }
            // ---- quad reduction: the 4 lanes of a quad hold 4 interleaved ky subsets ----------
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
#pragma unroll
                for (int c = 0; c < (PSI ? 5 : NPL); c++) {
                    double v = F[mt][c];
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    F[mt][c] = v;
                }
                if constexpr (PSI) {
                    const double kap = a.kappa, kap2 = a.kappa * a.kappa;
                    // mean shear of the blended frame (grid_U.m:11; interpolate_U.m:19-23), every rounding pinned
                    const double al = __dadd_rn(a.alpha0, __dmul_rn((double)(a.j0 + ev), a.dalpha));
                    const double um = __dadd_rn(__dmul_rn(__dsub_rn(1.0, al), a.u_mean0), __dmul_rn(al, a.u_mean1));
                    F[mt][0] = fma(kap, F[mt][0], um);
                    F[mt][1] = kap * F[mt][1];
                    F[mt][2] = kap2 * F[mt][2];
                    F[mt][3] = kap2 * F[mt][3];
                    F[mt][4] = kap2 * F[mt][4];
                    F[mt][5] = -F[mt][2];
                }
            }
            if (MODE == SPEC_LEAPFROG) {
                if constexpr (NF >= 6) {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        // phi2(dt): x += dt*U(x1); k -= dt*(grad U)^T k with the OLD k (ode_symplectic.m:18-21)
                        const double u = F[mt][0], v = F[mt][1], ux = F[mt][2], uy = F[mt][3], vx = F[mt][4], vy = F[mt][5];
                        px[mt] = px[mt] + a.dt * u;
                        py[mt] = py[mt] + a.dt * v;
                        const double k0 = pk[mt], l0 = pl[mt];
                        pk[mt] = k0 - a.dt * (ux * k0 + vx * l0);
                        pl[mt] = l0 - a.dt * (uy * k0 + vy * l0);
                        // phi1(dt/2) with the new k
                        // (explicit un-fused mul/add so that a fused run of n steps and n single-step
                        // launches produce bit-identical states)
                        half_drift(pk[mt], pl[mt], a.f2, a.gH, __dmul_rn(__dmul_rn(0.5, a.dt), a.gH), hcx[mt], hcy[mt]);
                        px[mt] = __dadd_rn(px[mt], hcx[mt]);
                        py[mt] = __dadd_rn(py[mt], hcy[mt]);
                    }
                }
            } else {
                if (jq == 0) {
#pragma unroll
                    for (int mt = 0; mt < MT; mt++) {
                        if (prow[mt] < a.n) {
#pragma unroll
                            for (int c = 0; c < NF; c++)
                                if (a.out[c]) a.out[c][prow[mt]] = F[mt][c];
                        }
                    }
                }
            }
        }
        if (MODE == SPEC_LEAPFROG) {
            if (jq == 0) {
#pragma unroll
                for (int mt = 0; mt < MT; mt++) {
                    if (prow[mt] < a.n) {
                        a.x[prow[mt]] = px[mt]; a.y[prow[mt]] = py[mt];
                        a.k[prow[mt]] = pk[mt]; a.l[prow[mt]] = pl[mt];
                    }
                }
            }
        }
    }
}

template <int NPL, int G, int MT, int MODE, bool PSI, int ATAB>
static cudaError_t launch_inst2(const SpecArgs& a, int num_sms, cudaStream_t st) {
    constexpr int TILE_P = kConsumerWarps * 8 * MT;
    size_t smem = spectral_smem_bytes(a.g);
    auto kern = spectral_kernel<NPL, G, MT, MODE, PSI, ATAB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    long long ntiles = (a.n + TILE_P - 1) / TILE_P;
    int grid = (int)(ntiles < (long long)num_sms * kCtasPerSm ? ntiles : (long long)num_sms * kCtasPerSm);
    if (grid < 1) grid = 1;
    kern<<<grid, kSpecThreads, smem, st>>>(a);
    return cudaGetLastError();
}

template <int NPL, int G, int MT, int MODE, bool PSI>
static cudaError_t launch_inst(const SpecArgs& a, int num_sms, cudaStream_t st) {
    if constexpr (MT == 1) {
        if (a.g.atab == 1) return launch_inst2<NPL, G, MT, MODE, PSI, 1>(a, num_sms, st);
        if (a.g.atab == 2 && a.twid) return launch_inst2<NPL, G, MT, MODE, PSI, 2>(a, num_sms, st);
    }
    return launch_inst2<NPL, G, MT, MODE, PSI, 0>(a, num_sms, st);
}

template <int NPL, int G, int MT>
static cudaError_t dispatch_mode(const SpecArgs& a, int mode, int num_sms, cudaStream_t st) {
    if (a.psi) {
        if constexpr (NPL == 3) {
            if (mode == SPEC_LEAPFROG) return launch_inst<NPL, G, MT, SPEC_LEAPFROG, true>(a, num_sms, st);
            return launch_inst<NPL, G, MT, SPEC_EVAL, true>(a, num_sms, st);
        } else return cudaErrorInvalidValue;
    }
    if (mode == SPEC_LEAPFROG) {
        if constexpr (NPL == 6) return launch_inst<NPL, G, MT, SPEC_LEAPFROG, false>(a, num_sms, st);
        else return cudaErrorInvalidValue;
    }
    return launch_inst<NPL, G, MT, SPEC_EVAL, false>(a, num_sms, st);
}

cudaError_t launch_spectral(const SpecArgs& a, int mode, int mtiles, int num_sms, cudaStream_t st) {
    const int npl = a.g.npl, G = a.g.G;
#define SWRT_INST(NPL_, G_, MT_) \
    if (npl == NPL_ && G == G_ && mtiles == MT_) return dispatch_mode<NPL_, G_, MT_>(a, mode, num_sms, st);
    SWRT_INST(6, 4, 1) SWRT_INST(6, 2, 2)
    SWRT_INST(7, 4, 1) SWRT_INST(7, 2, 2)
    SWRT_INST(2, 12, 1) SWRT_INST(2, 6, 2)
    SWRT_INST(3, 8, 1) SWRT_INST(3, 4, 2)
    SWRT_INST(4, 6, 1) SWRT_INST(4, 3, 2)
    SWRT_INST(1, 24, 1) SWRT_INST(1, 12, 2)
#undef SWRT_INST
    return cudaErrorInvalidValue;
}

}  // namespace swrt
