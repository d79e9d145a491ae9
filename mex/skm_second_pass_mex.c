/*
 * [centers2, assignments2, distances2] = skm_second_pass_mex(XFull, bestCenters, bestAssignments)
 *
 * GPU drop-in for the two-pass block of kmeans_sparsified.m:542-560 (in core) and for
 * private/recalculateAssignmentLargeFile.m:85-113 (call it per chunk, or once on a matrix that
 * fits in host memory):
 *     centers2(:,k)  = mean( XFull(:, bestAssignments == k), 2 )             (:545-551)
 *     [assignments2, distances2] = findClusterAssignments( XFull, bestCenters )   dense branch (:558)
 * XFull is the dense p x n double matrix (points are columns) AFTER X*(1+2*eps) (:292,:310);
 * bestCenters p x K double; bestAssignments 1 x n double holding 1..K (0 = unassigned).
 * Calls skm_second_pass.  With nlhs <= 1 only the means are computed.
 */
#include "skm_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs != 3)
        mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "Three input arguments required.");
    if (nlhs > 3)
        mexErrMsgIdAndTxt("MATLAB:mexFile:maxlhs", "Too many output arguments.");
    if (mxIsSparse(prhs[0]) || mxIsComplex(prhs[0]) || !mxIsDouble(prhs[0]))
        mexErrMsgTxt("Requires first input to be a real full double matrix");
    if (mxIsSparse(prhs[1]) || mxIsComplex(prhs[1]) || !mxIsDouble(prhs[1]))
        mexErrMsgTxt("Requires second input to be a real full double matrix");
    const mwSize p = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    const mwSize K = mxGetN(prhs[1]);
    if (mxGetM(prhs[1]) != p)
        mexErrMsgTxt("Array of centers not of correct size");
    if (mxGetM(prhs[2]) * mxGetN(prhs[2]) != n || !mxIsDouble(prhs[2]))
        mexErrMsgTxt("Requires one assignment per column of X");
    if (K < 1 || p < 1)
        mexErrMsgTxt("Requires at least one center and one row");

    int32_t *a_in = (int32_t *)mxMalloc(sizeof(int32_t) * (n ? n : 1));
    const double *ad = mxGetPr(prhs[2]);
    for (mwSize j = 0; j < n; ++j) a_in[j] = (int32_t)ad[j];
    int32_t *a_out = nlhs > 1 ? (int32_t *)mxMalloc(sizeof(int32_t) * (n ? n : 1)) : NULL;

    plhs[0] = mxCreateDoubleMatrix(p, K, mxREAL);
    if (nlhs > 1) plhs[1] = mxCreateDoubleMatrix(1, n, mxREAL);
    if (nlhs > 2) plhs[2] = mxCreateDoubleMatrix(1, n, mxREAL);
    const int rc = skm_second_pass(skm_mex_ctx(), (int64_t)p, (int64_t)n, mxGetPr(prhs[0]), SKM_F64, 0, 1.0,
                                   mxGetPr(prhs[1]), (int64_t)K, a_in, mxGetPr(plhs[0]), NULL, a_out,
                                   nlhs > 2 ? mxGetPr(plhs[2]) : NULL, 0, NULL);
    if (rc == SKM_OK && nlhs > 1) {
        double *o = mxGetPr(plhs[1]);
        for (mwSize j = 0; j < n; ++j) o[j] = (double)a_out[j];
    }
    mxFree(a_in);
    if (a_out) mxFree(a_out);
    if (rc != SKM_OK) {
        if (nlhs > 2) mxDestroyArray(plhs[2]);
        if (nlhs > 1) mxDestroyArray(plhs[1]);
        skm_mex_check(rc, plhs[0]);
    }
}
