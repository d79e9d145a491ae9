/*
 * [innerProd, normX2] = SparseMatrixInnerProduct(X, c)     X sparse p x n, c dense with p entries
 *
 * GPU drop-in for private/SparseMatrixInnerProduct.c (gateway :40-43, checks :48-77, 1 x n
 * outputs :80-83).  The reference validates the centre length against n (:71-77), which is a
 * typo for p (it indexes c by row); this shim checks p.  Calls skm_sparse_matrix_inner_product.
 */
#include "skm_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs != 2)
        mexErrMsgIdAndTxt("MATLAB:mexFile:invalidNumInputs", "Two input arguments required.");
    if (nlhs > 2)
        mexErrMsgIdAndTxt("MATLAB:mexFile:maxlhs", "Too many output arguments, needs 1 or 2 outputs.");
    if (!mxIsSparse(prhs[0]))
        mexErrMsgTxt("Requires first input to be a sparse matrix");
    if (mxIsComplex(prhs[0]) || !mxIsDouble(prhs[0]) || mxIsSparse(prhs[1]) || !mxIsDouble(prhs[1]))
        mexErrMsgTxt("Requires a real double sparse matrix and a full double center vector");
    const mwSize p = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    if (mxGetNumberOfElements(prhs[1]) != p)
        mexErrMsgTxt("Center vector must be 1xp or px1");
    plhs[0] = mxCreateDoubleMatrix(1, n, mxREAL);
    mxArray *nrm = mxCreateDoubleMatrix(1, n, mxREAL);
    if (n > 0) {
        int rc = skm_sparse_matrix_inner_product(skm_mex_ctx(), (int64_t)p, (int64_t)n,
                                                 (const uint64_t *)mxGetJc(prhs[0]), (const uint64_t *)mxGetIr(prhs[0]),
                                                 mxGetPr(prhs[0]), mxGetPr(prhs[1]), mxGetPr(plhs[0]), mxGetPr(nrm));
        if (rc != SKM_OK) { mxDestroyArray(nrm); skm_mex_check(rc, plhs[0]); }
    }
    if (nlhs > 1) plhs[1] = nrm;
    else mxDestroyArray(nrm);
}
