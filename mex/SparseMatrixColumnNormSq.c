/*
 * normX2 = SparseMatrixColumnNormSq(X)      X sparse p x n  ->  1 x n
 *
 * GPU drop-in for private/SparseMatrixColumnNormSq.c (gateway :35-38, loop :71-77).
 * Calls skm_sparse_matrix_column_normsq.
 */
#include "skm_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs != 1)
        mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "One input argument required.");
    if (nlhs > 1)
        mexErrMsgIdAndTxt("MATLAB:mexFile:maxlhs", "Too many output arguments.");
    if (!mxIsSparse(prhs[0]))
        mexErrMsgTxt("Requires first input to be a sparse matrix");
    if (mxIsComplex(prhs[0]) || !mxIsDouble(prhs[0]))
        mexErrMsgTxt("Requires a real double sparse matrix");
    const mwSize p = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    plhs[0] = mxCreateDoubleMatrix(1, n, mxREAL);
    if (n == 0) return;
    skm_mex_check(skm_sparse_matrix_column_normsq(skm_mex_ctx(), (int64_t)p, (int64_t)n,
                                                  (const uint64_t *)mxGetJc(prhs[0]), mxGetPr(prhs[0]),
                                                  mxGetPr(plhs[0])), plhs[0]);
}
