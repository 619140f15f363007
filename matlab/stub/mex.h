/* Minimal declaration of the MEX API subset matlab/swrt_mex.c uses, for an image that has neither MATLAB nor GNU Octave.
 * Declarations follow the documented C MEX interface: separate real/imaginary storage by default (Octave's mkoctfile
 * --mex, MATLAB -R2017b), the interleaved-complex accessors when MX_HAS_INTERLEAVED_COMPLEX is defined to 1 (MATLAB
 * -R2018a).  tests/mex_harness/mxharness.c implements these functions so that the gateway can be EXECUTED in tests;
 * a real build uses the tool's own mex.h.  Never shipped. */
#ifndef SWRT_STUB_MEX_H
#define SWRT_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxCHAR_CLASS = 4, mxDOUBLE_CLASS = 6, mxUINT64_CLASS = 13 } mxClassID;
typedef struct { double real, imag; } mxComplexDouble;
#ifdef __cplusplus
extern "C" {
#endif
#if defined(MX_HAS_INTERLEAVED_COMPLEX) && MX_HAS_INTERLEAVED_COMPLEX
double* mxGetDoubles(const mxArray*);
mxComplexDouble* mxGetComplexDoubles(const mxArray*);
#else
double* mxGetPi(const mxArray*);
#endif
double* mxGetPr(const mxArray*);
void* mxGetData(const mxArray*);
double mxGetScalar(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
int mxIsDouble(const mxArray*);
int mxIsComplex(const mxArray*);
int mxIsEmpty(const mxArray*);
int mxIsChar(const mxArray*);
int mxGetString(const mxArray*, char*, mwSize);
void* mxCalloc(size_t, size_t);
void mxFree(void*);
mxArray* mxCreateDoubleMatrix(mwSize, mwSize, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateNumericMatrix(mwSize, mwSize, mxClassID, mxComplexity);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
void mexLock(void);
int mexAtExit(void (*)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);
#ifdef __cplusplus
}
#endif
#endif
