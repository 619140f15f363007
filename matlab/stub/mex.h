/* Minimal stand-in for the MEX API, ONLY to compile-check matlab/swrt_mex.c in an image that has
 * neither MATLAB nor GNU Octave.  Declarations follow the documented C MEX interface (separate
 * real/imaginary storage, as Octave's mkoctfile --mex provides).  Never shipped, never linked. */
#ifndef SWRT_STUB_MEX_H
#define SWRT_STUB_MEX_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxUINT64_CLASS = 13 } mxClassID;
#ifdef __cplusplus
extern "C" {
#endif
double* mxGetPr(const mxArray*);
double* mxGetPi(const mxArray*);
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
