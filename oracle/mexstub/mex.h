/*
 * Stub mex.h -- TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * A minimal stand-in for MathWorks' mex.h so that MEX gateway sources can be
 * compiled with plain gcc/g++ and driven through ctypes.  It declares exactly
 * the mx/mex symbols used by the reference's five C files
 * (/root/reference/private/{SparseMatrixMinusCluster,SparseMatrixInnerProduct,
 * SparseMatrixColumnNormSq,hadamard,hadamard_pthreads}.c) plus the handful our
 * own shims in mex/ need.  MATLAB is not installed in this image; nothing in
 * the product path includes this file.
 *
 * Error semantics: mexErrMsgTxt / mexErrMsgIdAndTxt never return in MATLAB.
 * Here they record the message and longjmp back into mexstub_call().
 */
#ifndef SKM_ORACLE_STUB_MEX_H
#define SKM_ORACLE_STUB_MEX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef size_t mwSize;    /* -largeArrayDims: 64-bit sizes (setup_kmeans.m:19) */
typedef size_t mwIndex;
typedef ptrdiff_t mwSignedIndex;

typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum {
    mxUNKNOWN_CLASS = 0, mxCHAR_CLASS = 4, mxDOUBLE_CLASS = 6, mxSINGLE_CLASS = 7,
    mxINT32_CLASS = 12, mxUINT64_CLASS = 15
} mxClassID;

typedef struct mxArray_tag {
    mwSize   m, n;
    void    *pr;          /* data (double unless classid says otherwise) */
    mwIndex *ir;          /* sparse only */
    mwIndex *jc;          /* sparse only */
    int      is_sparse;
    int      is_complex;
    int      classid;     /* mxClassID */
    int      owns;        /* 1 = allocated by the stub (mxCreate*) */
    mwSize   nzmax;
} mxArray;

mwSize   mxGetM(const mxArray *a);
mwSize   mxGetN(const mxArray *a);
double  *mxGetPr(const mxArray *a);
void    *mxGetData(const mxArray *a);
mwIndex *mxGetIr(const mxArray *a);
mwIndex *mxGetJc(const mxArray *a);
int      mxIsSparse(const mxArray *a);
int      mxIsComplex(const mxArray *a);
int      mxIsDouble(const mxArray *a);
int      mxIsSingle(const mxArray *a);
int      mxIsEmpty(const mxArray *a);
mwSize   mxGetNumberOfElements(const mxArray *a);
double   mxGetScalar(const mxArray *a);
int      mxIsChar(const mxArray *a);
char    *mxArrayToString(const mxArray *a);   /* stub char arrays hold one byte per character */
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c);
mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c);
mxArray *mxCreateDoubleScalar(double v);
void     mxDestroyArray(mxArray *a);
void    *mxMalloc(size_t n);
void    *mxCalloc(size_t n, size_t sz);
void     mxFree(void *p);
int      mexPrintf(const char *fmt, ...);
void     mexErrMsgTxt(const char *msg);
void     mexErrMsgIdAndTxt(const char *id, const char *fmt, ...);
void     mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...);
int      mexAtExit(void (*fn)(void));

/* every MEX source defines this */
void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);

/* ---- stub driver API (called through ctypes) ---- */
/* returns 0 on success, 1 if the gateway raised a MEX error */
int         mexstub_call(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]);
const char *mexstub_last_error(void);
const char *mexstub_last_error_id(void);
void        mexstub_run_atexit(void);

#ifdef __cplusplus
}
#endif
#endif
