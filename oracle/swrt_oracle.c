/* swrt_oracle.c -- plain-C restatement of the SWRaytracing packet hot path.
 *
 * TEST INFRASTRUCTURE ONLY: this is the CPU checker/baseline.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it; the product (libswrt.so) never
 * links or calls it.  Parity status: the packet arithmetic restated here (interpolate, leapfrog,
 * step_packet*) is PINNED to the outputs of the unmodified reference .m files, executed from
 * /root/reference by oracle/minimat (the MATLAB-subset interpreter of this repository; no MATLAB /
 * Octave in the image): tests/golden/octave_out/*.bin, held to 1e-12 / 1e-9 by
 * tests/test_octave_goldens.py::test_c_port_against_reference_outputs, and this file is bit-identical
 * to oracle/swrt_oracle.py.  histcounts is a MATLAB builtin (nothing of the reference's to execute):
 * pinned by known-answer tests (tests/test_oracle_kat.py).  The spectral kit / initial-condition
 * chain of the numpy oracle is pinned against the reference's stored run logs and pv_time stream
 * (tests/test_reference_goldens.py).
 *
 * Every function cites the reference file:line it follows (relative to /root/reference).
 * Arrays are MATLAB column-major: F[ix + nx*iy]; spectral planes fk[(kx+kmax) + nkx*ky].
 * Loops over packets are OpenMP-parallel (packets are independent); the per-packet arithmetic keeps
 * the reference's operation order.
 *
 * Build: gcc -O3 -march=x86-64-v3 -ffp-contract=off -fopenmp -fPIC -shared -o oracle/build/liboracle.so oracle/swrt_oracle.c -lm
 * (no -ffast-math: operation order is part of the contract)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define IORD 2
#define NW 6

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* MATLAB mod(a,m), m>0 */
static inline double mmod(double a, double m) {
    double r = fmod(a, m);
    if (r < 0.0) r += m;
    return r;
}

/* interpolate.m:33-41 -- factor-by-factor multiply then divide, j ascending */
static inline void lag_weights(double a, double bump, double* w) {
    for (int i = -IORD; i <= IORD + 1; i++) {
        double wi = 1.0;
        for (int j = -IORD; j <= IORD + 1; j++)
            if (i != j) wi = wi * (a - (double)j + bump) / (double)(j - i);
        w[i + IORD] = wi;
    }
}

typedef struct { int ig[NW], jg[NW]; double wx[NW], wy[NW]; } stencil_t;

/* interpolate.m:21-31,45-46 (both indices wrap with nx) */
static inline void make_stencil(double x, double y, double dx, double dy, int nx, int ny, double bump, stencil_t* s) {
    double xl = mmod(x / dx, (double)nx);
    double yl = mmod(y / dy, (double)ny);
    double i0 = 1.0 + floor(xl), j0 = 1.0 + floor(yl);
    double ax = 1.0 + xl - i0, ay = 1.0 + yl - j0;
    lag_weights(ax, bump, s->wx);
    lag_weights(ay, bump, s->wy);
    long ii = (long)i0, jj = (long)j0;
    for (int i = -IORD; i <= IORD + 1; i++) {
        long a = (ii + i - 1) % nx; if (a < 0) a += nx;
        long b = (jj + i - 1) % nx; if (b < 0) b += nx;
        s->ig[i + IORD] = (int)a; s->jg[i + IORD] = (int)b;
    }
}

/* interpolate.m:43-49: i outer, j inner, term (wx_i*wy_j)*F(ig,jg) */
static inline double stencil_sum(const stencil_t* s, const double* F, int nx) {
    double acc = 0.0;
    for (int i = 0; i < NW; i++)
        for (int j = 0; j < NW; j++) acc = acc + s->wx[i] * s->wy[j] * F[(size_t)s->jg[j] * nx + s->ig[i]];
    return acc;
}

/* FI = interpolate(x,y,F,dx,dy): ray_trace_sw/interpolate.m:1-50 */
void orc_interpolate(const double* x, const double* y, int64_t n, const double* F, int nx, int ny, double dx,
                     double dy, double bump, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; m++) {
        stencil_t s;
        make_stencil(x[m], y[m], dx, dy, nx, ny, bump, &s);
        out[m] = stencil_sum(&s, F, nx);
    }
}

/* six planes at once; grids[c] column-major.  interpolate_U.m:5-17 for one frame / SpectralScheme.m:45-68 */
void orc_interpolate6(const double* x, const double* y, int64_t n, const double* const* grids, int npl, int nx,
                      double dx, double bump, double* const* out) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; m++) {
        stencil_t s;
        make_stencil(x[m], y[m], dx, dx, nx, nx, bump, &s);
        for (int c = 0; c < npl; c++) out[c][m] = stencil_sum(&s, grids[c], nx);
    }
}

/* ode_symplectic.m:13-21,33-37 with SpectralScheme.U / grad_U (SpectralScheme.m:45-68):
 * nsteps leapfrog steps, Lagrange-interpolated gridded planes u,v,u_x,u_y,v_x,v_y. */
void orc_leapfrog_lagrange(double* x, double* y, double* k, double* l, int64_t n, const double* const* grids, int nx,
                           double dx, double bump, double f, double gH, double dt, int nsteps) {
    const double h = dt / 2, f2 = f * f;
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; m++) {
        double px = x[m], py = y[m], pk = k[m], pl = l[m];
        for (int st = 0; st < nsteps; st++) {
            double om = sqrt(f2 + gH * (pk * pk + pl * pl));
            px = px + h * (gH * pk / om);
            py = py + h * (gH * pl / om);
            stencil_t s;
            make_stencil(px, py, dx, dx, nx, nx, bump, &s);
            double F[6];
            for (int c = 0; c < 6; c++) F[c] = stencil_sum(&s, grids[c], nx);
            px = px + dt * F[0];
            py = py + dt * F[1];
            double k0 = pk, l0 = pl;
            pk = k0 - dt * (F[2] * k0 + F[4] * l0);
            pl = l0 - dt * (F[3] * k0 + F[5] * l0);
            om = sqrt(f2 + gH * (pk * pk + pl * pl));
            px = px + h * (gH * pk / om);
            py = py + h * (gH * pl / om);
        }
        x[m] = px; y[m] = py; k[m] = pk; l[m] = pl;
    }
}

/* The same leapfrog on a TIME-DEPENDENT flow: two stored frames, every plane interpolated in BOTH and blended
 * (1-alpha)*F1 + alpha*F2 exactly as qg_flow_ray_trace/interpolate_U.m:5-23 does; sub-step j evaluates at
 * alpha_j = alpha0 + j*dalpha (the time-centred choice of SURVEY 7.0; the production drivers pass t/tmax, qgsw_raytrace.m:261). */
void orc_leapfrog_lagrange2(double* x, double* y, double* k, double* l, int64_t n, const double* const* g1,
                            const double* const* g2, int nx, double dx, double bump, double f, double gH, double dt,
                            int nsteps, double alpha0, double dalpha) {
    const double h = dt / 2, f2 = f * f;
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; m++) {
        double px = x[m], py = y[m], pk = k[m], pl = l[m];
        for (int st = 0; st < nsteps; st++) {
            const double al = alpha0 + (double)st * dalpha;
            double om = sqrt(f2 + gH * (pk * pk + pl * pl));
            px = px + h * (gH * pk / om);
            py = py + h * (gH * pl / om);
            stencil_t s;
            make_stencil(px, py, dx, dx, nx, nx, bump, &s);
            double F[6];
            for (int c = 0; c < 6; c++) F[c] = (1.0 - al) * stencil_sum(&s, g1[c], nx) + al * stencil_sum(&s, g2[c], nx);
            px = px + dt * F[0];
            py = py + dt * F[1];
            double k0 = pk, l0 = pl;
            pk = k0 - dt * (F[2] * k0 + F[4] * l0);
            pl = l0 - dt * (F[3] * k0 + F[5] * l0);
            om = sqrt(f2 + gH * (pk * pk + pl * pl));
            px = px + h * (gH * pk / om);
            py = py + h * (gH * pl / om);
        }
        x[m] = px; y[m] = py; k[m] = pk; l[m] = pl;
    }
}

/* ---- exact trig-sum evaluation (SPECTRAL mode oracle; pattern scratch/fourier_interpolate_test.m:125-136)
 *   F = sum_kx Fs(kx,0) e^{i kx tx} + 2 Re sum_{ky>=1} sum_kx F(kx,ky) e^{i(kx tx + ky ty)},
 *   tx = 2 pi mod(x/dx, nx)/nx.  The ky=0 column is conjugate-symmetrised (fulspec.m:16).
 * Template-by-macro on the accumulation type: long double = checker, double = timed CPU port. */
#define DEFINE_SPECTRAL_POINT(NAME, T, SIN, COS)                                                              \
    static void NAME(double x, double y, const double* const* pre, const double* const* pim, int npl, int nx, \
                     double dx, T* cx, T* sx, double* out) {                                                  \
        const int kmax = nx / 2 - 1, nkx = nx - 1, nky = nx / 2;                                              \
        const T two_pi = 2 * (T)3.14159265358979323846264338327950288L;                                       \
        T tx = two_pi * (T)mmod(x / dx, (double)nx) / (T)nx;                                                  \
        T ty = two_pi * (T)mmod(y / dx, (double)nx) / (T)nx;                                                  \
        for (int i = 0; i < nkx; i++) { T a = (T)(i - kmax) * tx; cx[i] = COS(a); sx[i] = SIN(a); }           \
        for (int c = 0; c < npl; c++) {                                                                       \
            const double* re = pre[c]; const double* im = pim[c];                                             \
            T acc = 0;                                                                                        \
            for (int ky = 0; ky < nky; ky++) {                                                                \
                T gr = 0, gi = 0;                                                                             \
                if (ky == 0) {                                                                                \
                    /* symmetrised column: F(0,0).re + 2 sum_{kx>0} Re(F(kx,0) e^{i kx tx}) */               \
                    gr = (T)re[kmax];                                                                         \
                    for (int i = kmax + 1; i < nkx; i++) gr += 2 * ((T)re[i] * cx[i] - (T)im[i] * sx[i]);    \
                    acc += gr;                                                                                \
                } else {                                                                                      \
                    const double* r = re + (size_t)ky * nkx; const double* q = im + (size_t)ky * nkx;         \
                    for (int i = 0; i < nkx; i++) {                                                           \
                        gr += (T)r[i] * cx[i] - (T)q[i] * sx[i];                                              \
                        gi += (T)r[i] * sx[i] + (T)q[i] * cx[i];                                              \
                    }                                                                                         \
                    T a = (T)ky * ty;                                                                         \
                    acc += 2 * (gr * COS(a) - gi * SIN(a));                                                   \
                }                                                                                             \
            }                                                                                                 \
            out[c] = (double)acc;                                                                             \
        }                                                                                                     \
    }

DEFINE_SPECTRAL_POINT(spectral_point_ld, long double, sinl, cosl)
DEFINE_SPECTRAL_POINT(spectral_point_d, double, sin, cos)

/* planes: npl pointers each to col-major (nkx x nky) re / im.  precise != 0 -> long double sums. */
void orc_spectral_eval(const double* x, const double* y, int64_t n, const double* const* pre, const double* const* pim,
                       int npl, int nx, double dx, int precise, double* const* out) {
    const int nkx = nx - 1;
#pragma omp parallel
    {
        long double* cl = (long double*)malloc(sizeof(long double) * 2 * nkx);
        double* cd = (double*)malloc(sizeof(double) * 2 * nkx);
        double F[8];
#pragma omp for schedule(static)
        for (int64_t m = 0; m < n; m++) {
            if (precise) spectral_point_ld(x[m], y[m], pre, pim, npl, nx, dx, cl, cl + nkx, F);
            else spectral_point_d(x[m], y[m], pre, pim, npl, nx, dx, cd, cd + nkx, F);
            for (int c = 0; c < npl; c++) out[c][m] = F[c];
        }
        free(cl); free(cd);
    }
}

/* leapfrog with exact trig-sum planes (ode_symplectic.m:13-21,33-37 + spectral evaluation) */
void orc_leapfrog_spectral(double* x, double* y, double* k, double* l, int64_t n, const double* const* pre,
                           const double* const* pim, int nx, double dx, int precise, double f, double gH, double dt,
                           int nsteps) {
    const double h = dt / 2, f2 = f * f;
    const int nkx = nx - 1;
#pragma omp parallel
    {
        long double* cl = (long double*)malloc(sizeof(long double) * 2 * nkx);
        double* cd = (double*)malloc(sizeof(double) * 2 * nkx);
#pragma omp for schedule(static)
        for (int64_t m = 0; m < n; m++) {
            double px = x[m], py = y[m], pk = k[m], pl = l[m];
            for (int st = 0; st < nsteps; st++) {
                double om = sqrt(f2 + gH * (pk * pk + pl * pl));
                px = px + h * (gH * pk / om);
                py = py + h * (gH * pl / om);
                double F[8];
                if (precise) spectral_point_ld(px, py, pre, pim, 6, nx, dx, cl, cl + nkx, F);
                else spectral_point_d(px, py, pre, pim, 6, nx, dx, cd, cd + nkx, F);
                px = px + dt * F[0];
                py = py + dt * F[1];
                double k0 = pk, l0 = pl;
                pk = k0 - dt * (F[2] * k0 + F[4] * l0);
                pl = l0 - dt * (F[3] * k0 + F[5] * l0);
                om = sqrt(f2 + gH * (pk * pk + pl * pl));
                px = px + h * (gH * pk / om);
                py = py + h * (gH * pl / om);
            }
            x[m] = px; y[m] = py; k[m] = pk; l[m] = pl;
        }
        free(cl); free(cd);
    }
}

/* ---- RK4 packet steps, Lagrange semantics: step_packet.m:37-78, step_packet_xka.m:38-91, cg_sw.m:15-31.
 * The reference adds the per-packet group velocity to every grid node and interpolates the sum; only
 * the 36 stencil nodes contribute, so the sum is formed at those nodes (same values, same order). */
static inline void vel_stage(const stencil_t* s, const double* const* g, int nx, int xka, double k, double l, double K2,
                             double C02, double f2, double Cx, double Cy, double* vx, double* vy) {
    double ax = 0.0, ay = 0.0;
    for (int i = 0; i < NW; i++)
        for (int j = 0; j < NW; j++) {
            size_t id = (size_t)s->jg[j] * nx + s->ig[i];
            double w = s->wx[i] * s->wy[j];
            if (xka) {
                double gH = C02 * g[6][id];
                double om = sqrt(f2 + gH * K2);
                ax = ax + w * (g[0][id] + gH * k / om);
                ay = ay + w * (g[1][id] + gH * l / om);
            } else {
                ax = ax + w * (g[0][id] + Cx);
                ay = ay + w * (g[1][id] + Cy);
            }
        }
    *vx = ax; *vy = ay;
}

void orc_rk4_lagrange(double* x, double* y, double* k, double* l, double* a, int64_t n, const double* const* g, int nx,
                      double dx, double bump, double f, double C0, double dt, int nsteps, int xka) {
    const double f2 = f * f, C02 = C0 * C0;
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; m++) {
        double px = x[m], py = y[m], pk = k[m], pl = l[m], pa = a ? a[m] : 1.0;
        for (int st = 0; st < nsteps; st++) {
            double K2 = pk * pk + pl * pl, Cx = 0, Cy = 0;
            if (!xka) { double om = sqrt(f2 + C02 * K2); Cx = C02 * pk / om; Cy = C02 * pl / om; }
            stencil_t s;
            double vx, vy, gr[4] = {0, 0, 0, 0};
            make_stencil(px, py, dx, dx, nx, nx, bump, &s);
            if (!xka) for (int c = 0; c < 4; c++) gr[c] = stencil_sum(&s, g[2 + c], nx);   /* old position */
            vel_stage(&s, g, nx, xka, pk, pl, K2, C02, f2, Cx, Cy, &vx, &vy);
            double x1 = dt * vx, y1 = dt * vy;
            make_stencil(px + x1 / 2, py + y1 / 2, dx, dx, nx, nx, bump, &s);
            vel_stage(&s, g, nx, xka, pk, pl, K2, C02, f2, Cx, Cy, &vx, &vy);
            double x2 = dt * vx, y2 = dt * vy;
            make_stencil(px + x2 / 2, py + y2 / 2, dx, dx, nx, nx, bump, &s);
            vel_stage(&s, g, nx, xka, pk, pl, K2, C02, f2, Cx, Cy, &vx, &vy);
            double x3 = dt * vx, y3 = dt * vy;
            make_stencil(px + x3, py + y3, dx, dx, nx, nx, bump, &s);
            vel_stage(&s, g, nx, xka, pk, pl, K2, C02, f2, Cx, Cy, &vx, &vy);
            double x4 = dt * vx, y4 = dt * vy;
            double xn = px + (x1 + 2 * x2 + 2 * x3 + x4) / 6;
            double yn = py + (y1 + 2 * y2 + 2 * y3 + y4) / 6;
            double oxi = 0, oyi = 0, dci = 0;
            if (xka) {
                make_stencil(xn, yn, dx, dx, nx, nx, bump, &s);
                for (int c = 0; c < 4; c++) gr[c] = stencil_sum(&s, g[2 + c], nx);
                for (int i = 0; i < NW; i++)
                    for (int j = 0; j < NW; j++) {
                        size_t id = (size_t)s.jg[j] * nx + s.ig[i];
                        double w = s.wx[i] * s.wy[j];
                        double u = g[0][id], v = g[1][id];
                        double gH = C02 * g[6][id];
                        double om = sqrt(f2 + gH * K2);
                        double cx = gH * pk / om, cy = gH * pl / om;
                        oxi = oxi + w * (f * K2 * v / (2 * om));
                        oyi = oyi + w * (-f * K2 * u / (2 * om));
                        dci = dci + w * ((pk * f * v - pl * f * u - cx * cx - cy * cy) / om);
                    }
            }
            double k1 = dt * (-gr[0] * pk - gr[2] * pl - oxi), l1 = dt * (-gr[1] * pk - gr[3] * pl - oyi);
            double k2 = dt * (-gr[0] * (pk + k1 / 2) - gr[2] * (pl + l1 / 2) - oxi);
            double l2 = dt * (-gr[1] * (pk + k1 / 2) - gr[3] * (pl + l1 / 2) - oyi);
            double k3 = dt * (-gr[0] * (pk + k2 / 2) - gr[2] * (pl + l2 / 2) - oxi);
            double l3 = dt * (-gr[1] * (pk + k2 / 2) - gr[3] * (pl + l2 / 2) - oyi);
            double k4 = dt * (-gr[0] * (pk + k3) - gr[2] * (pl + l3) - oxi);
            double l4 = dt * (-gr[1] * (pk + k3) - gr[3] * (pl + l3) - oyi);
            pk = pk + (k1 + 2 * k2 + 2 * k3 + k4) / 6;
            pl = pl + (l1 + 2 * l2 + 2 * l3 + l4) / 6;
            if (xka) {
                double a1 = dt * (-pa * dci), a2 = dt * (-(pa + a1 / 2) * dci);
                double a3 = dt * (-(pa + a2 / 2) * dci), a4 = dt * (-(pa + a3) * dci);
                pa = pa + (a1 + 2 * a2 + 2 * a3 + a4) / 6;
            }
            px = xn; py = yn;
        }
        x[m] = px; y[m] = py; k[m] = pk; l[m] = pl;
        if (a) a[m] = pa;
    }
}

/* odefun, qgsw_raytrace.m:259-265, given the six evaluated planes e[c][m] */
void orc_rhs(const double* k, const double* l, int64_t n, const double* const* e, double f, double Cg, double* dxdt,
             double* dydt, double* dkdt, double* dldt) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < n; m++) {
        double w = sqrt(f * f + Cg * Cg * (k[m] * k[m] + l[m] * l[m]));
        dxdt[m] = e[0][m] + Cg * k[m] / w;
        dydt[m] = e[1][m] + Cg * l[m] / w;
        dkdt[m] = -(e[2][m] * k[m] + e[4][m] * l[m]);
        dldt[m] = -(e[3][m] * k[m] + e[5][m] * l[m]);
    }
}

/* histcounts(w, edges), analysis/load_data.m:47: [e_i, e_i+1), last bin closed, NaN/out-of-range dropped */
void orc_histcounts(const double* w, int64_t n, const double* edges, int nedges, uint64_t* counts) {
    const int nb = nedges - 1;
    memset(counts, 0, sizeof(uint64_t) * nb);
    for (int64_t m = 0; m < n; m++) {
        double v = w[m];
        if (!(v >= edges[0] && v <= edges[nb])) continue;
        int a = 0, b = nb;
        while (b - a > 1) { int mid = (a + b) >> 1; if (edges[mid] <= v) a = mid; else b = mid; }
        counts[a]++;
    }
}
