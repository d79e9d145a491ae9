/*
 * dist = SparseMatrixMinusCluster(X, c)        X sparse p x n, c dense p x K  ->  K x n
 * dist = SparseMatrixMinusCluster(X, c, beta)  K == 1 only
 *
 * GPU drop-in for private/SparseMatrixMinusCluster.c of the reference (gateway :44-47,
 * argument checks :57-75,:104-107, K x n output :113).  Calls skm_sparse_matrix_minus_cluster.
 */
#include "skm_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs != 2 && nrhs != 3)
        mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "Two input arguments required.");
    if (nlhs > 1)
        mexErrMsgIdAndTxt("MATLAB:mexFile:maxlhs", "Too many output arguments.");
    if (!mxIsSparse(prhs[0]))
        mexErrMsgTxt("Requires first input to be a sparse matrix");
    if (mxIsComplex(prhs[0]) || !mxIsDouble(prhs[0]) || mxIsSparse(prhs[1]) || !mxIsDouble(prhs[1]))
        mexErrMsgTxt("Requires a real double sparse matrix and a full double center array");
    const mwSize p = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    if (mxGetM(prhs[1]) != p)
        mexErrMsgTxt("Center vector must be or pxk, but this vector did not have p rows");
    const mwSize K = mxGetN(prhs[1]);
    const int has_beta = (nrhs == 3);
    const double beta = has_beta ? mxGetScalar(prhs[2]) : 0.0;
    if (has_beta && K != 1)
        mexErrMsgTxt("Have not yet implemented case for using 'beta' with p x k (k!=`) centers");
    plhs[0] = mxCreateDoubleMatrix(K, n, mxREAL);
    if (n == 0 || K == 0) return;
    skm_mex_check(skm_sparse_matrix_minus_cluster(skm_mex_ctx(), (int64_t)p, (int64_t)n, (int64_t)K,
                                                  (const uint64_t *)mxGetJc(prhs[0]),
                                                  (const uint64_t *)mxGetIr(prhs[0]), mxGetPr(prhs[0]),
                                                  mxGetPr(prhs[1]), has_beta, beta, mxGetPr(plhs[0])),
                  plhs[0]);
}
