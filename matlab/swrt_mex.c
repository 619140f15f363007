/* swrt_mex.c -- MEX gateway over libswrt.so (include/swrt.h) for MATLAB / GNU Octave.
 *
 *   out = swrt_mex('command', handle, args...)
 *
 * The gateway holds no logic: it unpacks mxArrays (column-major fp64, separate real/imag), calls one
 * C-ABI function and packs the result.  Handles travel as uint64 scalars.  Errors are raised with
 * mexErrMsgIdAndTxt("swrt:...") only after the C call has returned (libswrt never throws).
 * Build:   mkoctfile --mex swrt_mex.c -I../include -L../swraytracing_b200 -lswrt
 *          mex swrt_mex.c -I../include -L../swraytracing_b200 -lswrt          (MATLAB)
 * Neither tool exists in this image; the file is compile-checked against matlab/stub/mex.h. */
#include <string.h>
#include "mex.h"
#include "swrt.h"

#define MAXH 64
static swrt_handle* g_handles[MAXH];
static int g_locked = 0;

static void destroy_all(void) {
    for (int i = 0; i < MAXH; i++) if (g_handles[i]) { swrt_destroy(g_handles[i]); g_handles[i] = NULL; }
}
static void fail(swrt_handle* h, const char* what) {
    mexErrMsgIdAndTxt("swrt:call", "%s: %s", what, swrt_last_error(h));
}
static swrt_handle* H(const mxArray* a) {
    uint64_t id = mxIsDouble(a) ? (uint64_t)mxGetScalar(a) : *(uint64_t*)mxGetData(a);
    if (id < 1 || id > MAXH || !g_handles[id - 1]) mexErrMsgIdAndTxt("swrt:handle", "invalid handle");
    return g_handles[id - 1];
}
static double* vec(mxArray** out, size_t n) { *out = mxCreateDoubleMatrix(n, 1, mxREAL); return mxGetPr(*out); }
static double* opt(const mxArray* prhs[], int nrhs, int i) { return (i < nrhs && !mxIsEmpty(prhs[i])) ? mxGetPr(prhs[i]) : NULL; }

void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    char cmd[48];
    if (nrhs < 1 || !mxIsChar(prhs[0]) || mxGetString(prhs[0], cmd, sizeof cmd)) mexErrMsgIdAndTxt("swrt:usage", "swrt_mex('command', ...)");
    if (!g_locked) { mexLock(); mexAtExit(destroy_all); g_locked = 1; }
    (void)nlhs;

    if (!strcmp(cmd, "create")) {           /* h = swrt_mex('create', nx, L, f, gH, mode, device, bump, flags) */
        swrt_params p; memset(&p, 0, sizeof p);
        p.nx = (int)mxGetScalar(prhs[1]); p.L = mxGetScalar(prhs[2]); p.f = mxGetScalar(prhs[3]); p.gH = mxGetScalar(prhs[4]);
        p.mode = nrhs > 5 ? (int)mxGetScalar(prhs[5]) : SWRT_MODE_SPECTRAL;
        p.device = nrhs > 6 ? (int)mxGetScalar(prhs[6]) : 0;
        p.bump = nrhs > 7 ? mxGetScalar(prhs[7]) : 1e-13;
        p.flags = nrhs > 8 ? (int)mxGetScalar(prhs[8]) : 0;
        int slot = 0; while (slot < MAXH && g_handles[slot]) slot++;
        if (slot == MAXH) mexErrMsgIdAndTxt("swrt:handle", "too many handles");
        if (swrt_create(&p, &g_handles[slot])) fail(NULL, "swrt_create");
        plhs[0] = mxCreateNumericMatrix(1, 1, mxUINT64_CLASS, mxREAL);
        *(uint64_t*)mxGetData(plhs[0]) = (uint64_t)slot + 1;
        return;
    }
    if (!strcmp(cmd, "interpolate")) {      /* FI = swrt_mex('interpolate', x, y, F, dx, dy [, bump]) */
        size_t n = mxGetNumberOfElements(prhs[1]);
        plhs[0] = mxCreateDoubleMatrix(mxGetM(prhs[1]), mxGetN(prhs[1]), mxREAL);
        if (swrt_interpolate(0, mxGetPr(prhs[1]), mxGetPr(prhs[2]), (int64_t)n, mxGetPr(prhs[3]), (int)mxGetM(prhs[3]), (int)mxGetN(prhs[3]),
                             mxGetScalar(prhs[4]), mxGetScalar(prhs[5]), nrhs > 6 ? mxGetScalar(prhs[6]) : 1e-13, mxGetPr(plhs[0])))
            fail(NULL, "swrt_interpolate");
        return;
    }
    if (!strcmp(cmd, "k2g")) {              /* fg = swrt_mex('k2g', fk) */
        int nx = (int)mxGetM(prhs[1]) + 1; size_t nh = mxGetNumberOfElements(prhs[1]);
        plhs[0] = mxCreateDoubleMatrix(nx, nx, mxREAL);
        mxArray* z = mxCreateDoubleMatrix(nh, 1, mxREAL);      /* zero imaginary part for real input */
        if (swrt_k2g(0, mxGetPr(prhs[1]), mxIsComplex(prhs[1]) ? mxGetPi(prhs[1]) : mxGetPr(z), nx, mxGetPr(plhs[0]))) fail(NULL, "swrt_k2g");
        return;
    }
    if (!strcmp(cmd, "g2k")) {              /* fk = swrt_mex('g2k', fg) */
        int nx = (int)mxGetM(prhs[1]);
        plhs[0] = mxCreateDoubleMatrix(nx - 1, nx / 2, mxCOMPLEX);
        if (swrt_g2k(0, mxGetPr(prhs[1]), nx, mxGetPr(plhs[0]), mxGetPi(plhs[0]))) fail(NULL, "swrt_g2k");
        return;
    }
    if (nrhs < 2) mexErrMsgIdAndTxt("swrt:usage", "missing handle");
    swrt_handle* h = H(prhs[1]);
    size_t n = (size_t)swrt_num_packets(h);

    if (!strcmp(cmd, "destroy")) {
        uint64_t id = mxIsDouble(prhs[1]) ? (uint64_t)mxGetScalar(prhs[1]) : *(uint64_t*)mxGetData(prhs[1]);
        swrt_destroy(h); g_handles[id - 1] = NULL;
    } else if (!strcmp(cmd, "set_flow_spectral")) {   /* (h, slot, psik [, u_mean]) */
        if (swrt_set_flow_spectral(h, (int)mxGetScalar(prhs[2]), mxGetPr(prhs[3]), mxGetPi(prhs[3]), (int)mxGetM(prhs[3]), (int)mxGetN(prhs[3]),
                                   nrhs > 4 ? mxGetScalar(prhs[4]) : 0.0)) fail(h, cmd);
    } else if (!strcmp(cmd, "set_flow_grid")) {       /* (h, slot, u, v, ux, uy, vx, vy [, H]) */
        if (swrt_set_flow_grid(h, (int)mxGetScalar(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(prhs[5]), mxGetPr(prhs[6]), mxGetPr(prhs[7]),
                               mxGetPr(prhs[8]), opt(prhs, nrhs, 9), (int)mxGetM(prhs[3]))) fail(h, cmd);
    } else if (!strcmp(cmd, "set_packets")) {         /* (h, x, y, k, l [, a]) */
        if (swrt_set_packets(h, (int64_t)mxGetNumberOfElements(prhs[2]), mxGetPr(prhs[2]), mxGetPr(prhs[3]), mxGetPr(prhs[4]), mxGetPr(prhs[5]),
                             opt(prhs, nrhs, 6))) fail(h, cmd);
    } else if (!strcmp(cmd, "get_packets")) {         /* [x, y, k, l, a] = ... */
        double* o[5]; for (int i = 0; i < 5; i++) o[i] = vec(&plhs[i], n);
        if (swrt_get_packets(h, o[0], o[1], o[2], o[3], o[4])) fail(h, cmd);
    } else if (!strcmp(cmd, "eval")) {                /* [U, V, Ux, Uy, Vx, Vy] = (h, alpha) */
        double* o[6]; for (int i = 0; i < 6; i++) o[i] = vec(&plhs[i], n);
        if (swrt_eval(h, mxGetScalar(prhs[2]), o[0], o[1], o[2], o[3], o[4], o[5])) fail(h, cmd);
    } else if (!strcmp(cmd, "eval_at")) {             /* [U, V, Ux, Uy, Vx, Vy] = (h, alpha, x, y) */
        size_t m = mxGetNumberOfElements(prhs[3]); double* o[6]; for (int i = 0; i < 6; i++) o[i] = vec(&plhs[i], m);
        if (swrt_eval_at(h, mxGetScalar(prhs[2]), (int64_t)m, mxGetPr(prhs[3]), mxGetPr(prhs[4]), o[0], o[1], o[2], o[3], o[4], o[5], NULL)) fail(h, cmd);
    } else if (!strcmp(cmd, "rhs")) {                 /* [dxdt, dydt, dkdt, dldt] = (h, alpha) */
        double* o[4]; for (int i = 0; i < 4; i++) o[i] = vec(&plhs[i], n);
        if (swrt_rhs(h, mxGetScalar(prhs[2]), o[0], o[1], o[2], o[3])) fail(h, cmd);
    } else if (!strcmp(cmd, "step")) {                /* (h, scheme, dt, nsteps [, alpha0, dalpha]) */
        if (swrt_step(h, (int)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), (int)mxGetScalar(prhs[4]), nrhs > 5 ? mxGetScalar(prhs[5]) : 0.0,
                      nrhs > 6 ? mxGetScalar(prhs[6]) : 0.0)) fail(h, cmd);
    } else if (!strcmp(cmd, "bs23_begin")) {          /* rh = (h, alpha, threshold)          -- ode23 building blocks */
        plhs[0] = mxCreateDoubleMatrix(1, 1, mxREAL);
        if (swrt_bs23_begin(h, mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), mxGetPr(plhs[0]))) fail(h, cmd);
    } else if (!strcmp(cmd, "bs23_attempt")) {        /* err = (h, hstep, [a2 a3 a4], threshold) */
        plhs[0] = mxCreateDoubleMatrix(1, 1, mxREAL);
        if (mxGetNumberOfElements(prhs[3]) != 3) mexErrMsgIdAndTxt("swrt:usage", "bs23_attempt: alpha must have 3 entries");
        if (swrt_bs23_attempt(h, mxGetScalar(prhs[2]), mxGetPr(prhs[3]), mxGetScalar(prhs[4]), mxGetPr(plhs[0]))) fail(h, cmd);
    } else if (!strcmp(cmd, "bs23_accept")) {         /* (h) */
        if (swrt_bs23_accept(h)) fail(h, cmd);
    } else if (!strcmp(cmd, "bs23_interp")) {         /* [x, y, k, l] = (h, hstep, s)        -- ntrp23 */
        double* o[4]; for (int i = 0; i < 4; i++) o[i] = vec(&plhs[i], n);
        if (swrt_bs23_interp(h, mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), o[0], o[1], o[2], o[3])) fail(h, cmd);
    } else if (!strcmp(cmd, "hist_omega")) {          /* counts = (h, kind, alpha, edges) */
        int ne = (int)mxGetNumberOfElements(prhs[4]);
        plhs[0] = mxCreateNumericMatrix(1, ne - 1, mxUINT64_CLASS, mxREAL);
        if (swrt_hist_omega(h, (int)mxGetScalar(prhs[2]), mxGetScalar(prhs[3]), mxGetPr(prhs[4]), ne, (uint64_t*)mxGetData(plhs[0]), 0)) fail(h, cmd);
    } else if (!strcmp(cmd, "diag")) {                /* d = (h, alpha) */
        if (swrt_diag(h, mxGetScalar(prhs[2]), vec(&plhs[0], 8))) fail(h, cmd);
    } else if (!strcmp(cmd, "omega")) {               /* [omega, Omega] = (h, alpha) */
        double* a = vec(&plhs[0], n); double* b = vec(&plhs[1], n);
        if (swrt_omega(h, mxGetScalar(prhs[2]), a, b)) fail(h, cmd);
    } else {
        mexErrMsgIdAndTxt("swrt:usage", "unknown command '%s'", cmd);
    }
}
