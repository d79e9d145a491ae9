/*
 * Handle-based gateway so X stays resident on the GPU across Lloyd iterations (level 2 of
 * include/skm_b200.h).  Usage from MATLAB (see INTEGRATION.md for the patched Lloyd loop):
 *
 *   h  = skm_lloyd_mex('upload', X, K)                 X sparse p x n (points = columns)
 *   h  = skm_lloyd_mex('upload', X, K, ngpus)          same, columns sharded over ngpus GPUs of this box
 *                                                      (0 = every visible GPU), driven from this one process:
 *                                                      skm_multi_* -- peer-memory reduction fused into K3
 *   [assignments, distances, centers, dff, sumsq, counts] = ...
 *        skm_lloyd_mex('iterate', h, centers, gammaDist, gammaUpdate, MLcorrection)
 *                                                      gammaDist = [] -> no division (findClusterAssignments.m:79-80)
 *   [assignments, distances] = skm_lloyd_mex('assign', h, centers, gamma)
 *   skm_lloyd_mex('free', h)
 *
 * 'iterate' is one pass of kmeans_sparsified.m:420-471 for non-empty clusters: K1 assign,
 * K2 accumulate, K3 finalize.  Empty clusters are reported through `counts`; the MATLAB side
 * applies EmptyAction exactly as the reference does (:432-445).
 */
#include <string.h>
#include "skm_mex_common.h"

#define MAX_H 64
static struct { skm_dataset *ds; skm_lloyd *L; int64_t K; skm_multi_dataset *mds; skm_multi_lloyd *ML; } g_h[MAX_H];
static skm_multi *g_multi = NULL;          /* one device group per MATLAB process, created by the first multi upload */
static int g_multi_ndev = 0;

static void free_handle(int i)
{
    if (g_h[i].L) skm_lloyd_destroy(g_h[i].L);
    if (g_h[i].ds) skm_dataset_destroy(g_h[i].ds);
    if (g_h[i].ML) skm_multi_lloyd_destroy(g_h[i].ML);
    if (g_h[i].mds) skm_multi_dataset_destroy(g_h[i].mds);
    g_h[i].L = NULL; g_h[i].ds = NULL; g_h[i].ML = NULL; g_h[i].mds = NULL;
}

static void free_all(void)
{
    for (int i = 0; i < MAX_H; ++i) free_handle(i);
    if (g_multi) { skm_multi_destroy(g_multi); g_multi = NULL; }
    skm_mex_atexit();
}

static int handle_of(const mxArray *a)
{
    int h = (int)mxGetScalar(a);
    if (h < 0 || h >= MAX_H || !(g_h[h].ds || g_h[h].mds)) mexErrMsgIdAndTxt("skm_b200:badHandle", "invalid handle");
    return h;
}

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    char cmd[32];
    if (nrhs < 1 || !mxIsChar(prhs[0]) || mxGetNumberOfElements(prhs[0]) >= sizeof cmd)
        mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "First argument must be a command string");
    {
        char *c = mxArrayToString(prhs[0]);
        strncpy(cmd, c ? c : "", sizeof cmd - 1);
        cmd[sizeof cmd - 1] = 0;
        mxFree(c);
    }
    mexAtExit(free_all);

    if (!strcmp(cmd, "upload") && nrhs == 4) {
        /* several GPUs behind the same handle */
        if (!mxIsSparse(prhs[1]) || mxIsComplex(prhs[1]) || !mxIsDouble(prhs[1]))
            mexErrMsgTxt("Requires X to be a real double sparse matrix");
        int h = 0;
        while (h < MAX_H && (g_h[h].ds || g_h[h].mds)) ++h;
        if (h == MAX_H) mexErrMsgIdAndTxt("skm_b200:tooManyHandles", "too many live datasets");
        const int want = (int)mxGetScalar(prhs[3]);
        if (g_multi && want != g_multi_ndev && want != 0) {
            for (int i = 0; i < MAX_H; ++i)
                if (g_h[i].mds) mexErrMsgIdAndTxt("skm_b200:deviceGroup", "free the datasets of the current device group before changing ngpus");
            skm_multi_destroy(g_multi); g_multi = NULL;
        }
        if (!g_multi) { skm_mex_check(skm_multi_create(want, NULL, &g_multi), NULL); g_multi_ndev = want; }
        const int64_t p = (int64_t)mxGetM(prhs[1]), n = (int64_t)mxGetN(prhs[1]);
        g_h[h].K = (int64_t)mxGetScalar(prhs[2]);
        skm_mex_check(skm_multi_dataset_create_csc(g_multi, p, n, mxGetJc(prhs[1]), SKM_I64, mxGetIr(prhs[1]), SKM_I64,
                                                   mxGetPr(prhs[1]), SKM_F64, SKM_F32, &g_h[h].mds), NULL);
        int rc = skm_multi_lloyd_create(g_h[h].mds, g_h[h].K, &g_h[h].ML);
        if (rc != SKM_OK) { skm_multi_dataset_destroy(g_h[h].mds); g_h[h].mds = NULL; skm_mex_check(rc, NULL); }
        plhs[0] = mxCreateDoubleScalar((double)h);
        return;
    }
    if ((!strcmp(cmd, "assign") || !strcmp(cmd, "iterate")) && nrhs >= 2 && g_h[handle_of(prhs[1])].mds) {
        const int iterate = cmd[0] == 'i';
        if (nrhs != (iterate ? 6 : 4)) mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "wrong number of inputs");
        const int h = handle_of(prhs[1]);
        skm_dataset_info info;
        skm_multi_dataset_get_info(g_h[h].mds, &info);
        if ((int64_t)mxGetM(prhs[2]) != info.p || (int64_t)mxGetN(prhs[2]) != g_h[h].K || mxIsSparse(prhs[2]))
            mexErrMsgIdAndTxt("findClusterAssignments:size", "Array of centers not of correct size");
        const int has_gamma = !mxIsEmpty(prhs[3]);
        const double gamma = has_gamma ? mxGetScalar(prhs[3]) : 0.0;
        skm_multi_lloyd *ML = g_h[h].ML;
        skm_iter_stats st;
        memset(&st, 0, sizeof st);
        skm_mex_check(skm_multi_lloyd_set_centers(ML, mxGetPr(prhs[2])), NULL);
        /* 'assign' = an iteration whose centre update is discarded: the centres are set again on the next call */
        skm_mex_check(skm_multi_lloyd_step(ML, has_gamma, gamma, iterate ? mxGetScalar(prhs[4]) : 1.0,
                                           iterate ? (mxGetScalar(prhs[5]) != 0.0) : 1, 0, &st), NULL);
        mxArray *ai = mxCreateNumericMatrix(1, (mwSize)info.n, mxINT32_CLASS, mxREAL);
        mxArray *d = mxCreateDoubleMatrix(1, (mwSize)info.n, mxREAL);
        int rc = skm_multi_lloyd_get_assignments(ML, (int32_t *)mxGetData(ai), mxGetPr(d));
        if (rc != SKM_OK) { mxDestroyArray(ai); skm_mex_check(rc, d); }
        plhs[0] = mxCreateDoubleMatrix(1, (mwSize)info.n, mxREAL);
        {
            const int32_t *s = (const int32_t *)mxGetData(ai);
            double *o = mxGetPr(plhs[0]);
            for (int64_t j = 0; j < info.n; ++j) o[j] = (double)s[j];
        }
        mxDestroyArray(ai);
        if (nlhs > 1) plhs[1] = d; else mxDestroyArray(d);
        if (iterate) {
            if (nlhs > 2) { plhs[2] = mxCreateDoubleMatrix((mwSize)info.p, (mwSize)g_h[h].K, mxREAL);
                            skm_mex_check(skm_multi_lloyd_get_centers(ML, mxGetPr(plhs[2])), NULL); }
            if (nlhs > 3) plhs[3] = mxCreateDoubleScalar(st.dff);
            if (nlhs > 4) plhs[4] = mxCreateDoubleScalar(st.sumsq);
            if (nlhs > 5) {
                plhs[5] = mxCreateDoubleMatrix(1, (mwSize)g_h[h].K, mxREAL);
                int64_t cnt[4096];
                if (g_h[h].K <= 4096 && skm_multi_lloyd_get_counts(ML, cnt) == SKM_OK)
                    for (int64_t k = 0; k < g_h[h].K; ++k) mxGetPr(plhs[5])[k] = (double)cnt[k];
            }
        }
        return;
    }
    if (!strcmp(cmd, "free") && nrhs == 2 && g_h[handle_of(prhs[1])].mds) {
        free_handle(handle_of(prhs[1]));
        return;
    }
    skm_ctx *ctx = skm_mex_ctx();

    if (!strcmp(cmd, "upload")) {
        if (nrhs != 3) mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "upload needs (X, K [, ngpus])");
        if (!mxIsSparse(prhs[1]) || mxIsComplex(prhs[1]) || !mxIsDouble(prhs[1]))
            mexErrMsgTxt("Requires X to be a real double sparse matrix");
        int h = 0;
        while (h < MAX_H && (g_h[h].ds || g_h[h].mds)) ++h;
        if (h == MAX_H) mexErrMsgIdAndTxt("skm_b200:tooManyHandles", "too many live datasets");
        const int64_t p = (int64_t)mxGetM(prhs[1]), n = (int64_t)mxGetN(prhs[1]);
        g_h[h].K = (int64_t)mxGetScalar(prhs[2]);
        /* K as a hint: a large X is uploaded in chunks and the entry order for K centres is built under the upload */
        skm_mex_check(skm_dataset_create_csc_hint(ctx, p, n, mxGetJc(prhs[1]), SKM_I64, mxGetIr(prhs[1]), SKM_I64,
                                                  mxGetPr(prhs[1]), SKM_F64, SKM_F32, 0, g_h[h].K, &g_h[h].ds), NULL);
        int rc = skm_lloyd_create(g_h[h].ds, g_h[h].K, &g_h[h].L);
        if (rc != SKM_OK) { skm_dataset_destroy(g_h[h].ds); g_h[h].ds = NULL; skm_mex_check(rc, NULL); }
        plhs[0] = mxCreateDoubleScalar((double)h);
        return;
    }
    if (!strcmp(cmd, "free")) {
        if (nrhs != 2) mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "free needs (h)");
        int h = handle_of(prhs[1]);
        skm_lloyd_destroy(g_h[h].L); skm_dataset_destroy(g_h[h].ds);
        g_h[h].L = NULL; g_h[h].ds = NULL;
        return;
    }
    if (!strcmp(cmd, "assign") || !strcmp(cmd, "iterate")) {
        const int iterate = cmd[0] == 'i';
        if (nrhs != (iterate ? 6 : 4)) mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "wrong number of inputs");
        int h = handle_of(prhs[1]);
        skm_dataset_info info;
        skm_dataset_get_info(g_h[h].ds, &info);
        if ((int64_t)mxGetM(prhs[2]) != info.p || (int64_t)mxGetN(prhs[2]) != g_h[h].K || mxIsSparse(prhs[2]))
            mexErrMsgIdAndTxt("findClusterAssignments:size", "Array of centers not of correct size");
        const int has_gamma = !mxIsEmpty(prhs[3]);
        const double gamma = has_gamma ? mxGetScalar(prhs[3]) : 0.0;
        skm_lloyd *L = g_h[h].L;
        skm_mex_check(skm_lloyd_set_centers(L, mxGetPr(prhs[2])), NULL);
        skm_mex_check(skm_lloyd_assign(L, has_gamma, gamma), NULL);
        skm_iter_stats st;
        memset(&st, 0, sizeof st);
        if (iterate) {
            skm_mex_check(skm_lloyd_accumulate(L), NULL);
            skm_mex_check(skm_lloyd_finalize(L, mxGetScalar(prhs[4]), mxGetScalar(prhs[5]) != 0.0, &st), NULL);
        }
        /* assignments come back as doubles holding 1..K, like MATLAB's min */
        mxArray *ai = mxCreateNumericMatrix(1, (mwSize)info.n, mxINT32_CLASS, mxREAL);
        mxArray *d = mxCreateDoubleMatrix(1, (mwSize)info.n, mxREAL);
        int rc = skm_lloyd_get_assignments(L, (int32_t *)mxGetData(ai), mxGetPr(d));
        if (rc != SKM_OK) { mxDestroyArray(ai); skm_mex_check(rc, d); }
        plhs[0] = mxCreateDoubleMatrix(1, (mwSize)info.n, mxREAL);
        {
            const int32_t *s = (const int32_t *)mxGetData(ai);
            double *o = mxGetPr(plhs[0]);
            for (int64_t j = 0; j < info.n; ++j) o[j] = (double)s[j];
        }
        mxDestroyArray(ai);
        if (nlhs > 1) plhs[1] = d; else mxDestroyArray(d);
        if (iterate) {
            if (nlhs > 2) { plhs[2] = mxCreateDoubleMatrix((mwSize)info.p, (mwSize)g_h[h].K, mxREAL);
                            skm_mex_check(skm_lloyd_get_centers(L, mxGetPr(plhs[2])), NULL); }
            if (nlhs > 3) plhs[3] = mxCreateDoubleScalar(st.dff);
            if (nlhs > 4) plhs[4] = mxCreateDoubleScalar(st.sumsq);
            if (nlhs > 5) {
                plhs[5] = mxCreateDoubleMatrix(1, (mwSize)g_h[h].K, mxREAL);
                int64_t cnt[4096];
                if (g_h[h].K <= 4096 && skm_lloyd_get_counts(L, cnt) == SKM_OK)
                    for (int64_t k = 0; k < g_h[h].K; ++k) mxGetPr(plhs[5])[k] = (double)cnt[k];
            }
        }
        return;
    }
    mexErrMsgIdAndTxt("skm_b200:badCommand", "unknown command '%s'", cmd);
}
