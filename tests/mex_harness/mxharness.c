/* mxharness.c -- a small implementation of the mx / mex functions matlab/swrt_mex.c calls, so that the gateway's
 * mexFunction can be EXECUTED without MATLAB or GNU Octave (neither exists in the development image).
 *
 * TEST INFRASTRUCTURE: built by tests/mex_harness/build.py into libmexharness{,_ic}.so together with the unmodified
 * matlab/swrt_mex.c and linked against libswrt.so; tests/test_mex_gateway.py drives it through ctypes.  Two builds:
 * separate real / imaginary storage (Octave, MATLAB -R2017b) and interleaved complex (MATLAB -R2018a,
 * -DMX_HAS_INTERLEAVED_COMPLEX=1), because the gateway has a code path for each.
 *
 * Semantics reproduced: column-major mxArrays; mxCalloc memory and arrays created inside mexFunction are released by the
 * harness when the call ends (as MATLAB's memory manager does); mexErrMsgIdAndTxt does not return -- it longjmps out of
 * mexFunction to the harness' call wrapper, which reports the identifier and message. */
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mex.h"

struct mxArray_tag {
    mxClassID cls;
    size_t m, n;
    int is_complex;
    double* re;              /* separate layout: real part; interleaved layout: real data or interleaved (re,im) pairs */
    double* im;              /* separate layout only */
    void* raw;               /* non-double classes (uint64, char) */
};

#define MAXTMP 4096
static void* g_tmp[MAXTMP];
static int g_ntmp = 0;
static jmp_buf g_jmp;
static int g_in_call = 0;
static char g_err_id[128], g_err_msg[1024];
static void (*g_atexit)(void) = NULL;
static int g_lock_count = 0;

static void* track(void* p) { if (g_in_call && p && g_ntmp < MAXTMP) g_tmp[g_ntmp++] = p; return p; }

static mxArray* new_array(mxClassID cls, size_t m, size_t n, int cplx) {
    mxArray* a = (mxArray*)calloc(1, sizeof(mxArray));
    a->cls = cls; a->m = m; a->n = n; a->is_complex = cplx;
    size_t cnt = (m * n > 0) ? m * n : 1;
    if (cls == mxDOUBLE_CLASS) {
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
        a->re = (double*)calloc(cnt * (cplx ? 2 : 1), sizeof(double));
#else
        a->re = (double*)calloc(cnt, sizeof(double));
        if (cplx) a->im = (double*)calloc(cnt, sizeof(double));
#endif
    } else {
        a->raw = calloc(cnt, 8);
    }
    return a;
}

/* ---- the MEX API subset ---- */
double* mxGetPr(const mxArray* a) { return a->cls == mxDOUBLE_CLASS ? a->re : NULL; }
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
double* mxGetDoubles(const mxArray* a) { return (a->cls == mxDOUBLE_CLASS && !a->is_complex) ? a->re : NULL; }
mxComplexDouble* mxGetComplexDoubles(const mxArray* a) { return (a->cls == mxDOUBLE_CLASS && a->is_complex) ? (mxComplexDouble*)a->re : NULL; }
#else
double* mxGetPi(const mxArray* a) { return a->im; }
#endif
void* mxGetData(const mxArray* a) { return a->cls == mxDOUBLE_CLASS ? (void*)a->re : a->raw; }
double mxGetScalar(const mxArray* a) {
    if (a->m * a->n == 0) return 0.0;
    if (a->cls == mxDOUBLE_CLASS) return a->re[0];
    if (a->cls == mxUINT64_CLASS) return (double)*(uint64_t*)a->raw;
    return 0.0;
}
size_t mxGetM(const mxArray* a) { return a->m; }
size_t mxGetN(const mxArray* a) { return a->n; }
size_t mxGetNumberOfElements(const mxArray* a) { return a->m * a->n; }
int mxIsDouble(const mxArray* a) { return a->cls == mxDOUBLE_CLASS; }
int mxIsComplex(const mxArray* a) { return a->is_complex; }
int mxIsEmpty(const mxArray* a) { return a->m * a->n == 0; }
int mxIsChar(const mxArray* a) { return a->cls == mxCHAR_CLASS; }
int mxGetString(const mxArray* a, char* buf, mwSize buflen) {
    if (a->cls != mxCHAR_CLASS) return 1;
    size_t len = a->m * a->n;
    if (len + 1 > buflen) return 1;
    memcpy(buf, a->raw, len); buf[len] = 0;
    return 0;
}
void* mxCalloc(size_t n, size_t sz) { return track(calloc(n ? n : 1, sz ? sz : 1)); }
void mxFree(void* p) {
    for (int i = 0; i < g_ntmp; i++) if (g_tmp[i] == p) { g_tmp[i] = NULL; break; }
    free(p);
}
mxArray* mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) { return new_array(mxDOUBLE_CLASS, m, n, c == mxCOMPLEX); }
mxArray* mxCreateDoubleScalar(double v) { mxArray* a = new_array(mxDOUBLE_CLASS, 1, 1, 0); a->re[0] = v; return a; }
mxArray* mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c) { return new_array(cls, m, n, c == mxCOMPLEX); }
void mexErrMsgIdAndTxt(const char* id, const char* fmt, ...) {
    va_list ap;
    snprintf(g_err_id, sizeof g_err_id, "%s", id);
    va_start(ap, fmt); vsnprintf(g_err_msg, sizeof g_err_msg, fmt, ap); va_end(ap);
    if (g_in_call) longjmp(g_jmp, 1);
    fprintf(stderr, "mexErrMsgIdAndTxt outside a call: %s: %s\n", g_err_id, g_err_msg);
    abort();
}
void mexLock(void) { g_lock_count++; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }

/* ---- driver API (called from Python through ctypes) ---- */
mxArray* hx_double(size_t m, size_t n, int cplx) { return new_array(mxDOUBLE_CLASS, m, n, cplx); }
mxArray* hx_uint64(uint64_t v) { mxArray* a = new_array(mxUINT64_CLASS, 1, 1, 0); *(uint64_t*)a->raw = v; return a; }
mxArray* hx_string(const char* s) {
    size_t len = strlen(s);
    mxArray* a = new_array(mxCHAR_CLASS, 1, len, 0);
    memcpy(a->raw, s, len);
    return a;
}
/* copy separate real / imaginary host buffers into / out of an array, whatever the storage layout */
void hx_set(mxArray* a, const double* re, const double* im) {
    size_t cnt = a->m * a->n;
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
    if (a->is_complex) { for (size_t i = 0; i < cnt; i++) { a->re[2 * i] = re[i]; a->re[2 * i + 1] = im ? im[i] : 0.0; } return; }
#else
    if (a->is_complex && im) memcpy(a->im, im, cnt * sizeof(double));
#endif
    memcpy(a->re, re, cnt * sizeof(double));
}
void hx_get(const mxArray* a, double* re, double* im) {
    size_t cnt = a->m * a->n;
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
    if (a->is_complex) { for (size_t i = 0; i < cnt; i++) { re[i] = a->re[2 * i]; if (im) im[i] = a->re[2 * i + 1]; } return; }
#else
    if (a->is_complex && im) memcpy(im, a->im, cnt * sizeof(double));
#endif
    memcpy(re, a->re, cnt * sizeof(double));
}
void hx_get_u64(const mxArray* a, uint64_t* out) { memcpy(out, a->raw, a->m * a->n * 8); }
void hx_dims(const mxArray* a, size_t* m, size_t* n, int* cls, int* cplx) { *m = a->m; *n = a->n; *cls = (int)a->cls; *cplx = a->is_complex; }
void hx_free(mxArray* a) { if (!a) return; free(a->re); free(a->im); free(a->raw); free(a); }
int hx_interleaved(void) {
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
    return 1;
#else
    return 0;
#endif
}
/* run mexFunction; 0 = returned normally, 1 = mexErrMsgIdAndTxt was raised (hx_error_id / hx_error_msg tell which) */
int hx_call(int nlhs, mxArray** plhs, int nrhs, const mxArray** prhs) {
    g_err_id[0] = g_err_msg[0] = 0;
    g_ntmp = 0;
    g_in_call = 1;
    int rc = 0;
    if (setjmp(g_jmp) == 0) mexFunction(nlhs, plhs, nrhs, prhs);
    else rc = 1;
    g_in_call = 0;
    for (int i = 0; i < g_ntmp; i++) free(g_tmp[i]);      /* what MATLAB's memory manager does at the end of a MEX call */
    g_ntmp = 0;
    return rc;
}
const char* hx_error_id(void) { return g_err_id; }
const char* hx_error_msg(void) { return g_err_msg; }
int hx_locked(void) { return g_lock_count; }
void hx_run_atexit(void) { if (g_atexit) g_atexit(); }
