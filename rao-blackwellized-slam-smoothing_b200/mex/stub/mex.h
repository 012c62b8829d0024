/* Minimal stand-in for MATLAB's mex.h, used ONLY to compile-check rbslam_mex.cpp in an
 * image without MATLAB (tests/test_capi_host.py).  Declarations follow the documented
 * MEX C API (R2018a+ interleaved-complex API names); nothing here is linked. */
#ifndef RBSLAM_STUB_MEX_H
#define RBSLAM_STUB_MEX_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6, mxINT32_CLASS = 12 } mxClassID;
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
double *mxGetDoubles(const mxArray *a);
void *mxGetData(const mxArray *a);
size_t mxGetM(const mxArray *a);
size_t mxGetN(const mxArray *a);
size_t mxGetNumberOfElements(const mxArray *a);
mwSize mxGetNumberOfDimensions(const mxArray *a);
const mwSize *mxGetDimensions(const mxArray *a);
int mxIsEmpty(const mxArray *a);
int mxIsDouble(const mxArray *a);
int mxIsStruct(const mxArray *a);
int mxIsChar(const mxArray *a);
double mxGetScalar(const mxArray *a);
char *mxArrayToString(const mxArray *a);
void mxFree(void *p);
mxArray *mxGetField(const mxArray *s, size_t index, const char *name);
mxArray *mxCreateDoubleMatrix(size_t m, size_t n, mxComplexity c);
mxArray *mxCreateNumericArray(mwSize ndim, const mwSize *dims, mxClassID cls, mxComplexity c);
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
int mexPrintf(const char *fmt, ...);
void mexLock(void);
void mexUnlock(void);
int mexCallMATLAB(int nlhs, mxArray *plhs[], int nrhs, mxArray *prhs[], const char *name);
void mxDestroyArray(mxArray *a);
int mxIsClass(const mxArray *a, const char *classname);
int mexAtExit(void (*fn)(void));
#ifdef __cplusplus
}
#endif
#endif
