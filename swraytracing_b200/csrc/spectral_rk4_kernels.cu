// spectral_rk4_kernels.cu -- SPECTRAL mode: step_packet (ray_trace_sw/step_packet.m:37-78) and step_packet_xka
// (step_packet_xka.m:38-91 with cg_sw.m:15-31 evaluated at the point) as ONE kernel launch per run of steps.
//
// Per step and packet the reference needs five field evaluations:
//   step_packet      stage 0 at the old position: u, v AND the four gradients (:41-42, :58-61) -> the "big" stack B
//                    (six planes, or the three psi-hat moment planes when the flow was given as psi-hat);
//                    stages 1..3 at the RK4 stage positions: u, v only (:44-51)                    -> the stage stack A (u,v)
//   step_packet_xka  stages 0..3: u, v, H (:42-52; gH = C0^2 H enters the group velocity)          -> A = (u,v,H)
//                    then the seven planes at the NEW position (:59-65)                           -> B = (u,v,ux,uy,vx,vy,H)
// Every evaluation is the same dense DMMA contraction as the leapfrog kernel (spectral_kernels.cu: +-kx folded stack
// streamed L2 -> shared memory by cp.async.bulk through a full/empty mbarrier ring, twiddles by recurrence or from the
// per-step shared-memory table, ky sum on the accumulator fragments, quad shuffle reduction); here the two stacks share ONE
// ring (stage size = the larger chunk) and the producer walks the per-step sequence A,A,A,A,B (xka) or B,A,A,A.  The
// packet (x, y, k, l, a), the RK4 accumulators and the evaluated planes never leave registers between stages; the state
// is read once and written once per launch.  The stage arithmetic is the expression-for-expression twin of
// rk4_stage_kernel / rk4_final_kernel (misc_kernels.cu), which the composed route (swrt_set_tuning flag 4) launches.
#include "spectral_common.cuh"

namespace swrt {

namespace {

// this thread's view of the shared chunk ring (consumer cursor + the producer duty of its warp)
struct Ring {
    unsigned char* smem;
    uint64_t* full_bar;
    uint64_t* empty_bar;
    double* sA;                 // this lane's column of the warp's twiddle table
    int nstages, D;
    uint32_t stage_bytes;
    // consumer
    int stage; uint32_t phase; long long ci; bool ready;
    // producer: this warp issues the chunks pj = warp, warp + 8, ...; (p_step, p_r) = position inside the source cycle
    long long pj, p_it, total_chunks;
    int p_st, p_step, p_r;
    const double* stackA; const double* stackB;
    size_t totalA, totalB, chunkA, chunkB;     // doubles
    uint32_t bytesA, bytesB;
    int cpeA, cpeB, cyc, nstack;
    bool b_first;               // step_packet: the big evaluation comes first; step_packet_xka: last
};

__device__ __forceinline__ void ring_issue(const Ring& r) {          // lane 0 only: issue chunk pj
    if (r.p_it > 0) mbar_wait(&r.empty_bar[r.p_st], (uint32_t)((r.p_it - 1) & 1));
    int q = r.p_r;
    bool is_b;
    if (r.b_first) {
        if (q < r.cpeB) is_b = true;
        else { is_b = false; q -= r.cpeB; while (q >= r.cpeA) q -= r.cpeA; }
    } else {
        if (q < 4 * r.cpeA) { is_b = false; while (q >= r.cpeA) q -= r.cpeA; }
        else { is_b = true; q -= 4 * r.cpeA; }
    }
    const double* src = is_b ? r.stackB + (size_t)r.p_step * r.totalB + (size_t)q * r.chunkB
                             : r.stackA + (size_t)r.p_step * r.totalA + (size_t)q * r.chunkA;
    const uint32_t bytes = is_b ? r.bytesB : r.bytesA;
    mbar_expect_tx(&r.full_bar[r.p_st], bytes);
    bulk_g2s(r.smem + (size_t)r.p_st * r.stage_bytes, src, bytes, &r.full_bar[r.p_st]);
}
__device__ __forceinline__ void ring_advance(Ring& r) {              // all lanes (uniform): pj += kConsumerWarps
    r.pj += kConsumerWarps;
    r.p_st += kConsumerWarps;
    while (r.p_st >= r.nstages) { r.p_st -= r.nstages; r.p_it++; }
    r.p_r += kConsumerWarps;
    while (r.p_r >= r.cyc) { r.p_r -= r.cyc; if (++r.p_step >= r.nstack) r.p_step = 0; }
}

// One evaluation of NPL stack planes at (px, py) for this lane's packet row: F[0..NF) on every lane of the quad.
// PSI: the stack holds the psi-hat moment planes and the six velocity / gradient planes are assembled (see
// spectral_kernels.cu); um = mean shear added to u.
template <int NPL, int G, bool PSI, int ATAB>
__device__ __forceinline__ void contract(Ring& r, const PackGeom& g, double px, double py, double dx, double nxd, double inv_nx,
                                         int lane, double kappa, double um, double* F) {
    constexpr int NT = NPL * G;
    constexpr int HALF_NT = NT / 2;
    constexpr int NF = PSI ? 6 : NPL;
    static_assert(NT % 2 == 0, "n-tiles are fetched in pairs");
    const int jq = lane & 3, kx0 = jq >> 1;
    const bool odd = lane & 1;
    // ---- twiddle seeds (the same recurrences as the leapfrog kernel) ----
    double s1, c1;
    sincospi(reduced_turns(px, dx, nxd, inv_nx), &s1, &c1);
    const double er = kx0 ? c1 : 1.0, ei = kx0 ? s1 : 0.0;
    const double ep0 = odd ? ei : er, eq0 = odd ? er : ei;
    double xdc = -2.0 * s1 * s1;
    const double ds = 2.0 * s1 * c1;
    double xds = odd ? ds : -ds;
    const double ep1 = fma(ep0, xdc, fma(eq0, xds, ep0));
    const double eq1 = fma(eq0, xdc, fma(-ep0, xds, eq0));
    {
        const double s2 = ds, c2 = fma(-2.0 * s1, s1, 1.0);
        xdc = -2.0 * s2 * s2;
        const double ds2 = 2.0 * s2 * c2;
        xds = odd ? ds2 : -ds2;
    }
    double sy, cy;
    sincospi(reduced_turns(py, dx, nxd, inv_nx), &sy, &cy);
    const Cplx e1{cy, sy};
    const Cplx e2 = cmul(e1, e1), e3 = cmul(e2, e1), e4 = cmul(e2, e2);
    Cplx ytw[G];
    {
        Cplx cur = jq == 0 ? Cplx{1.0, 0.0} : (jq == 1 ? e1 : (jq == 2 ? e2 : e3));
#pragma unroll
        for (int gg = 0; gg < G; gg++) { ytw[gg] = cur; cur = cmul(cur, e4); }
    }
    const Cplx yrot = cpow<G>(e4);
#pragma unroll
    for (int c = 0; c < NF; c++) F[c] = 0.0;
    double kyd = (double)jq;
    if constexpr (ATAB) {
        double ap = ep0, aq = eq0, bp = ep1, bq = eq1;
        for (int s = 0; s < g.ksteps; s += 2) {
            r.sA[s * 32] = ap; r.sA[(s + 1) * 32] = bp;
            const double nap = fma(ap, xdc, fma(aq, xds, ap)), naq = fma(aq, xdc, fma(-ap, xds, aq));
            const double nbp = fma(bp, xdc, fma(bq, xds, bp)), nbq = fma(bq, xdc, fma(-bp, xds, bq));
            ap = nap; aq = naq; bp = nbp; bq = nbq;
        }
    }
    double apf[kKUnroll];                          // ATAB 2: the A elements of the next unrolled body, in flight from L2
    if constexpr (ATAB == 2) {
#pragma unroll
        for (int su = 0; su < kKUnroll; su++) apf[su] = r.sA[su * 32];
    }
    const int chunks_per_pass = g.ksteps / g.kc;
    for (int pass = 0; pass < g.npass; pass++) {
        double acc[NT][2];
#pragma unroll
        for (int t = 0; t < NT; t++) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
        double tp = ep0, tq = eq0, tpB = ep1, tqB = eq1;
        for (int ch = 0; ch < chunks_per_pass; ch++) {
            if (r.ci + r.D == r.pj && r.pj < r.total_chunks) {      // producer duty for chunk ci + D (one warp in eight)
                if (lane == 0) ring_issue(r);
                ring_advance(r);
            }
            if (!r.ready) mbar_wait(&r.full_bar[r.stage], r.phase);
            const double2* sB = reinterpret_cast<const double2*>(r.smem + (size_t)r.stage * r.stage_bytes) + lane;
            const double* sAc = r.sA + (size_t)ch * g.kc * 32;
            int nstage = r.stage + 1; uint32_t nphase = r.phase;
            if (nstage == r.nstages) { nstage = 0; nphase ^= 1; }
            for (int s0 = 0; s0 < g.kc; s0 += kKUnroll) {
                if (s0 + kKUnroll >= g.kc) r.ready = mbar_test(&r.full_bar[nstage], nphase);
                double acur[kKUnroll];
                if constexpr (ATAB == 2) {
                    int nb = ch * g.kc + s0 + kKUnroll;
                    if (nb >= g.ksteps) nb = 0;
#pragma unroll
                    for (int su = 0; su < kKUnroll; su++) { acur[su] = apf[su]; apf[su] = r.sA[(nb + su) * 32]; }
                }
#pragma unroll
                for (int su = 0; su < kKUnroll; su++) {
                    const int s = s0 + su;
                    if constexpr (ATAB == 1) tp = sAc[s * 32];
                    if constexpr (ATAB == 2) tp = acur[su];
#pragma unroll
                    for (int t2 = 0; t2 < HALF_NT; t2++) {
                        const double2 b = sB[(s * HALF_NT + t2) * 32];
                        dmma884(acc[2 * t2][0], acc[2 * t2][1], tp, b.x);
                        dmma884(acc[2 * t2 + 1][0], acc[2 * t2 + 1][1], tp, b.y);
                    }
                    if constexpr (!ATAB) {
                        const double np_ = fma(tp, xdc, fma(tq, xds, tp));
                        const double nq_ = fma(tq, xdc, fma(-tp, xds, tq));
                        tp = tpB; tq = tqB; tpB = np_; tqB = nq_;
                    }
                }
            }
            if (lane == 0) mbar_arrive(&r.empty_bar[r.stage]);
            r.stage = nstage; r.phase = nphase;
            ++r.ci;
        }
        // ---- stage 2: the ky sum on the accumulator fragments ----
#pragma unroll
        for (int gg = 0; gg < G; gg++) {
            const double cyv = ytw[gg].re, syv = ytw[gg].im;
            if constexpr (PSI) {
                const double ky = kyd + (double)(4 * gg);
                const double g0r = acc[0 * G + gg][0], g0i = acc[0 * G + gg][1];
                const double g1r = acc[1 * G + gg][0], g1i = acc[1 * G + gg][1];
                const double g2r = acc[2 * G + gg][0], g2i = acc[2 * G + gg][1];
                const double a0 = fma(g0r, cyv, -g0i * syv), b0 = fma(g0i, cyv, g0r * syv);
                const double a1 = fma(g1r, cyv, -g1i * syv), b1 = fma(g1i, cyv, g1r * syv);
                const double a2 = fma(g2r, cyv, -g2i * syv);
                F[0] = fma(ky, b0, F[0]);
                F[1] += a1;
                F[2] = fma(ky, b1, F[2]);
                F[3] = fma(ky * ky, a0, F[3]);
                F[4] += a2;
            } else {
#pragma unroll
                for (int c = 0; c < NPL; c++) {
                    F[c] = fma(acc[c * G + gg][0], cyv, F[c]);
                    F[c] = fma(-acc[c * G + gg][1], syv, F[c]);
                }
            }
            ytw[gg] = cmul(ytw[gg], yrot);
        }
        kyd += (double)(4 * G);
    }
#pragma unroll
    for (int c = 0; c < (PSI ? 5 : NPL); c++) {
        double v = F[c];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        F[c] = v;
    }
    if constexpr (PSI) {
        const double kap2 = kappa * kappa;
        F[0] = fma(kappa, F[0], um);
        F[1] = kappa * F[1];
        F[2] = kap2 * F[2];
        F[3] = kap2 * F[3];
        F[4] = kap2 * F[4];
        F[5] = -F[2];
    }
}

// XKA: step_packet_xka (A = u,v,H; B = seven planes at the new position); else step_packet (B first, at the old position;
// PSIB: B = the three psi-hat moment planes).
template <bool XKA, bool PSIB, int ATAB>
__global__ void __launch_bounds__(kSpecThreads, 1) spectral_rk4_kernel(const SpecRk4Args a) {
    constexpr int NPL_A = XKA ? 3 : 2, G_A = XKA ? 8 : 12;
    constexpr int NPL_B = XKA ? 7 : (PSIB ? 3 : 6), G_B = XKA ? 4 : (PSIB ? 8 : 4);
    constexpr int TILE_P = kConsumerWarps * 8;
    static_assert(!(XKA && PSIB), "step_packet_xka needs the H plane: no psi-moment form");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    Ring r;
    r.smem = smem_raw;
    r.nstages = a.nstages;
    r.stage_bytes = a.stage_bytes;
    r.full_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)a.nstages * a.stage_bytes);
    r.empty_bar = r.full_bar + a.nstages;
    if constexpr (ATAB == 2) r.sA = a.twid + ((size_t)blockIdx.x * kConsumerWarps + warp) * a.tab_ksteps * 32 + lane;
    else r.sA = reinterpret_cast<double*>(smem_raw + (size_t)a.nstages * a.stage_bytes + 128) + (size_t)warp * a.tab_ksteps * 32 + lane;
    if (threadIdx.x == 0) {
        for (int s = 0; s < a.nstages; s++) { mbar_init(&r.full_bar[s], 1); mbar_init(&r.empty_bar[s], kConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const long long ntiles = (a.n + TILE_P - 1) / TILE_P;
    long long my_tiles = 0;
    if ((long long)blockIdx.x < ntiles) my_tiles = (ntiles - 1 - blockIdx.x) / gridDim.x + 1;
    r.stackA = a.stackA; r.stackB = a.stackB;
    r.totalA = a.gA.total_doubles; r.totalB = a.gB.total_doubles;
    r.chunkA = a.gA.chunk_doubles; r.chunkB = a.gB.chunk_doubles;
    r.bytesA = (uint32_t)(a.gA.chunk_doubles * 8); r.bytesB = (uint32_t)(a.gB.chunk_doubles * 8);
    r.cpeA = a.gA.chunks_per_eval; r.cpeB = a.gB.chunks_per_eval;
    r.cyc = (XKA ? 4 : 3) * r.cpeA + r.cpeB;
    r.nstack = a.nstack > 1 ? a.nstack : 1;
    r.b_first = !XKA;
    r.total_chunks = my_tiles * (long long)a.nsteps * r.cyc;
    r.D = a.nstages - a.lag;
    r.pj = warp;
    r.p_st = warp % a.nstages; r.p_it = warp / a.nstages;
    r.p_step = 0; r.p_r = warp;
    while (r.p_r >= r.cyc) { r.p_r -= r.cyc; if (++r.p_step >= r.nstack) r.p_step = 0; }
    r.stage = 0; r.phase = 0; r.ci = 0; r.ready = false;
    if (r.pj < r.D && r.pj < r.total_chunks) {
        if (lane == 0) ring_issue(r);
        ring_advance(r);
    }

    const int quad_row = lane >> 2, jq = lane & 3;
    const double dt = a.dt, C02 = a.C0 * a.C0, f = a.f;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const long long prow = tile * TILE_P + (long long)warp * 8 + quad_row;
        const long long rc = prow < a.n ? prow : a.n - 1;
        double x = a.x[rc], y = a.y[rc], k = a.k[rc], l = a.l[rc];
        double am = XKA ? a.a[rc] : 0.0;
        for (int st = 0; st < a.nsteps; st++) {
            // mean shear of the (blended) frame for the psi-moment form; the A stack carries it in its u plane
            const double al = __dadd_rn(a.alpha0, __dmul_rn((double)(a.j0 + st), a.dalpha));
            const double um = __dadd_rn(__dmul_rn(__dsub_rn(1.0, al), a.u_mean0), __dmul_rn(al, a.u_mean1));
            const double K2 = k * k + l * l;
            double gux = 0, guy = 0, gvx = 0, gvy = 0;
            double xs = x, ys = y, ax = 0, ay = 0;
#pragma unroll 1
            for (int stage = 0; stage < 4; stage++) {
                double u, v, Hh = 1.0;
                if (!XKA && stage == 0) {
                    double F[6];
                    contract<NPL_B, G_B, PSIB, ATAB>(r, a.gB, xs, ys, a.dx, a.nxd, a.inv_nx, lane, a.kappa, um, F);
                    u = F[0]; v = F[1]; gux = F[2]; guy = F[3]; gvx = F[4]; gvy = F[5];      // gradients at the OLD position (:58-61)
                } else {
                    double F[NPL_A];
                    contract<NPL_A, G_A, false, ATAB>(r, a.gA, xs, ys, a.dx, a.nxd, a.inv_nx, lane, 0.0, 0.0, F);
                    u = F[0]; v = F[1];
                    if constexpr (XKA) Hh = F[2];
                }
                const double gH = XKA ? C02 * Hh : C02;
                const double om = sqrt(f * f + gH * K2);
                const double dxs = dt * (u + gH * k / om);
                const double dys = dt * (v + gH * l / om);
                switch (stage) {
                    case 0: ax = dxs; ay = dys; xs = x + dxs / 2; ys = y + dys / 2; break;
                    case 1: ax += 2 * dxs; ay += 2 * dys; xs = x + dxs / 2; ys = y + dys / 2; break;
                    case 2: ax += 2 * dxs; ay += 2 * dys; xs = x + dxs; ys = y + dys; break;
                    default: {
                        const double sx = ax + dxs, sy = ay + dys;
                        xs = x + sx / 6; ys = y + sy / 6;       // = Pout.x, Pout.y
                    }
                }
            }
            double oxi = 0.0, oyi = 0.0, dci = 0.0;
            if constexpr (XKA) {
                double F[7];
                contract<NPL_B, G_B, false, ATAB>(r, a.gB, xs, ys, a.dx, a.nxd, a.inv_nx, lane, 0.0, 0.0, F);
                gux = F[2]; guy = F[3]; gvx = F[4]; gvy = F[5];
                const double gH = C02 * F[6];
                const double om = sqrt(f * f + gH * K2);
                const double cx = gH * k / om, cy = gH * l / om;
                oxi = f * K2 * F[1] / (2 * om);
                oyi = -f * K2 * F[0] / (2 * om);
                dci = (k * f * F[1] - l * f * F[0] - cx * cx - cy * cy) / om;
            }
            const double k1 = dt * (-gux * k - gvx * l - oxi);
            const double l1 = dt * (-guy * k - gvy * l - oyi);
            const double k2 = dt * (-gux * (k + k1 / 2) - gvx * (l + l1 / 2) - oxi);
            const double l2 = dt * (-guy * (k + k1 / 2) - gvy * (l + l1 / 2) - oyi);
            const double k3 = dt * (-gux * (k + k2 / 2) - gvx * (l + l2 / 2) - oxi);
            const double l3 = dt * (-guy * (k + k2 / 2) - gvy * (l + l2 / 2) - oyi);
            const double k4 = dt * (-gux * (k + k3) - gvx * (l + l3) - oxi);
            const double l4 = dt * (-guy * (k + k3) - gvy * (l + l3) - oyi);
            k = k + (k1 + 2 * k2 + 2 * k3 + k4) / 6;
            l = l + (l1 + 2 * l2 + 2 * l3 + l4) / 6;
            if constexpr (XKA) {
                const double a1 = dt * (-am * dci);
                const double a2 = dt * (-(am + a1 / 2) * dci);
                const double a3 = dt * (-(am + a2 / 2) * dci);
                const double a4 = dt * (-(am + a3) * dci);
                am = am + (a1 + 2 * a2 + 2 * a3 + a4) / 6;
            }
            x = xs; y = ys;
        }
        if (jq == 0 && prow < a.n) {
            a.x[prow] = x; a.y[prow] = y; a.k[prow] = k; a.l[prow] = l;
            if (XKA) a.a[prow] = am;
        }
    }
}

template <bool XKA, bool PSIB, int ATAB>
cudaError_t launch_rk4_inst(const SpecRk4Args& a, size_t smem, int num_sms, cudaStream_t st) {
    auto kern = spectral_rk4_kernel<XKA, PSIB, ATAB>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    constexpr int TILE_P = kConsumerWarps * 8;
    const long long ntiles = (a.n + TILE_P - 1) / TILE_P;
    int grid = (int)(ntiles < (long long)num_sms ? ntiles : (long long)num_sms);
    if (grid < 1) grid = 1;
    kern<<<grid, kSpecThreads, smem, st>>>(a);
    return cudaGetLastError();
}

}  // namespace

// joint ring geometry of the two stacks: stage = the larger chunk; the twiddle table (sized for the longer pass) is used
// when it fits beside a ring of at least three stages
bool spectral_rk4_geometry(SpecRk4Args& a, size_t* smem_bytes) {
    constexpr size_t kSmemBudget = 216 * 1024;
    const size_t cb = (a.gA.chunk_doubles > a.gB.chunk_doubles ? a.gA.chunk_doubles : a.gB.chunk_doubles) * 8;
    a.stage_bytes = (uint32_t)cb;
    a.tab_ksteps = a.gA.ksteps > a.gB.ksteps ? a.gA.ksteps : a.gB.ksteps;
    const size_t table = (size_t)a.tab_ksteps * 32 * 8 * kConsumerWarps;
    a.atab = (a.gA.atab == 1 && a.gB.atab == 1 && table + 3 * cb <= kSmemBudget) ? 1 : ((a.gA.atab && a.gB.atab) ? 2 : 0);
    const size_t ring_budget = a.atab == 1 ? kSmemBudget - table : (size_t)200 * 1024;
    int ns = (int)(ring_budget / cb);
    if (ns > 8) ns = 8;
    if (ns < 3) ns = 3;
    a.nstages = ns;
    a.lag = ns - 1 > 4 ? 4 : ns - 1;
    if (a.lag < 1) a.lag = 1;
    *smem_bytes = (size_t)ns * cb + 128 + (a.atab == 1 ? table : 0);
    return *smem_bytes <= 227 * 1024;
}

cudaError_t launch_spectral_rk4(const SpecRk4Args& a_in, bool xka, int num_sms, cudaStream_t st) {
    if (a_in.n <= 0 || a_in.nsteps <= 0) return cudaSuccess;
    SpecRk4Args a = a_in;
    size_t smem = 0;
    if (!spectral_rk4_geometry(a, &smem)) return cudaErrorInvalidValue;
    // the instantiations assume the n-tile grouping make_geom chose for these plane counts
    const bool ok = xka ? (a.gA.npl == 3 && a.gA.G == 8 && a.gB.npl == 7 && a.gB.G == 4 && !a.psiB)
                        : (a.gA.npl == 2 && a.gA.G == 12 && (a.psiB ? (a.gB.npl == 3 && a.gB.G == 8) : (a.gB.npl == 6 && a.gB.G == 4)));
    if (!ok) return cudaErrorInvalidValue;
    if (a.atab == 2 && !a.twid) a.atab = 0;
#define SWRT_RK4_DISPATCH(XKA_, PSI_)                                                                  \
    switch (a.atab) {                                                                                 \
        case 1: return launch_rk4_inst<XKA_, PSI_, 1>(a, smem, num_sms, st);                          \
        case 2: return launch_rk4_inst<XKA_, PSI_, 2>(a, smem, num_sms, st);                          \
        default: return launch_rk4_inst<XKA_, PSI_, 0>(a, smem, num_sms, st);                         \
    }
    if (xka) { SWRT_RK4_DISPATCH(true, false) }
    if (a.psiB) { SWRT_RK4_DISPATCH(false, true) }
    SWRT_RK4_DISPATCH(false, false)
#undef SWRT_RK4_DISPATCH
}

}  // namespace swrt
