/*
 * Stub MEX runtime -- TEST INFRASTRUCTURE ONLY (oracle/).  See mex.h here.
 * Linked into every library under oracle/_ref and into the test build of the mex/ shims.
 */
#include "mex.h"
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static __thread jmp_buf g_jmp;
static __thread int g_active = 0;
static __thread char g_err[1024];
static __thread char g_errid[256];
static void (*g_atexit)(void) = NULL;

mwSize mxGetM(const mxArray *a) { return a->m; }
mwSize mxGetN(const mxArray *a) { return a->n; }
double *mxGetPr(const mxArray *a) { return (double *)a->pr; }
void *mxGetData(const mxArray *a) { return a->pr; }
mwIndex *mxGetIr(const mxArray *a) { return a->ir; }
mwIndex *mxGetJc(const mxArray *a) { return a->jc; }
int mxIsSparse(const mxArray *a) { return a->is_sparse; }
int mxIsComplex(const mxArray *a) { return a->is_complex; }
int mxIsDouble(const mxArray *a) { return a->classid == mxDOUBLE_CLASS; }
int mxIsSingle(const mxArray *a) { return a->classid == mxSINGLE_CLASS; }
mwSize mxGetNumberOfElements(const mxArray *a) { return a->m * a->n; }
int mxIsEmpty(const mxArray *a) { return a->m == 0 || a->n == 0; }
double mxGetScalar(const mxArray *a) { return ((double *)a->pr)[0]; }
int mxIsChar(const mxArray *a) { return a->classid == mxCHAR_CLASS; }
char *mxArrayToString(const mxArray *a) {
    size_t n = a->m * a->n;
    char *s = (char *)malloc(n + 1);
    if (a->classid != mxCHAR_CLASS) { free(s); return NULL; }
    memcpy(s, a->pr, n);
    s[n] = 0;
    return s;
}

static size_t class_size(mxClassID c) {
    switch (c) {
        case mxCHAR_CLASS: return 1;
        case mxDOUBLE_CLASS: return 8;
        case mxSINGLE_CLASS: return 4;
        case mxINT32_CLASS: return 4;
        case mxUINT64_CLASS: return 8;
        default: return 8;
    }
}

mxArray *mxCreateNumericMatrix(mwSize m, mwSize n, mxClassID cls, mxComplexity c) {
    mxArray *a = (mxArray *)calloc(1, sizeof(mxArray));
    size_t cnt = m * n;
    a->m = m; a->n = n;
    a->pr = calloc(cnt ? cnt : 1, class_size(cls));
    a->classid = cls;
    a->is_complex = (c == mxCOMPLEX);
    a->owns = 1;
    return a;
}
mxArray *mxCreateDoubleMatrix(mwSize m, mwSize n, mxComplexity c) {
    return mxCreateNumericMatrix(m, n, mxDOUBLE_CLASS, c);
}
mxArray *mxCreateDoubleScalar(double v) {
    mxArray *a = mxCreateDoubleMatrix(1, 1, mxREAL);
    ((double *)a->pr)[0] = v;
    return a;
}
void mxDestroyArray(mxArray *a) {
    if (!a) return;
    if (a->owns) { free(a->pr); free(a->ir); free(a->jc); }
    free(a);
}
void *mxMalloc(size_t n) { return malloc(n ? n : 1); }
void *mxCalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz ? sz : 1); }
void mxFree(void *p) { free(p); }

int mexPrintf(const char *fmt, ...) { (void)fmt; return 0; } /* usage text is noise here */

static void raise_err(void) {
    if (g_active) longjmp(g_jmp, 1);
    fprintf(stderr, "mexstub: error outside mexstub_call: %s\n", g_err);
    abort();
}
void mexErrMsgTxt(const char *msg) {
    snprintf(g_err, sizeof g_err, "%s", msg ? msg : "");
    g_errid[0] = 0;
    raise_err();
}
void mexErrMsgIdAndTxt(const char *id, const char *fmt, ...) {
    va_list ap; va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt ? fmt : "", ap);
    va_end(ap);
    snprintf(g_errid, sizeof g_errid, "%s", id ? id : "");
    raise_err();
}
void mexWarnMsgIdAndTxt(const char *id, const char *fmt, ...) { (void)id; (void)fmt; }
int mexAtExit(void (*fn)(void)) { g_atexit = fn; return 0; }
void mexstub_run_atexit(void) { if (g_atexit) { g_atexit(); g_atexit = NULL; } }

int mexstub_call(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[]) {
    g_err[0] = 0; g_errid[0] = 0;
    g_active = 1;
    if (setjmp(g_jmp)) { g_active = 0; return 1; }
    mexFunction(nlhs, plhs, nrhs, prhs);
    g_active = 0;
    return 0;
}
const char *mexstub_last_error(void) { return g_err; }
const char *mexstub_last_error_id(void) { return g_errid; }
