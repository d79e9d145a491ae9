/*
 * w = hadamard(x)    x full real double m x n, m a power of two >= 2
 *
 * GPU drop-in for private/hadamard.c (gateway :115-152, size check :97-111) and, compiled as
 * hadamard_pthreads, for private/hadamard_pthreads.c (gateway :227-264): unnormalised
 * natural-order Walsh-Hadamard transform of every column.  Calls skm_hadamard.
 */
#include "skm_mex_common.h"

void mexFunction(int nlhs, mxArray *plhs[], int nrhs, const mxArray *prhs[])
{
    if (nrhs != 1)
        mexErrMsgTxt("One and only one input required; must be a column vector or matrix, with # rows a power of 2.");
    if (nlhs > 1)
        mexErrMsgTxt("Too many output arguments.");
    const mwSize m = mxGetM(prhs[0]), n = mxGetN(prhs[0]);
    if (m <= 1)
        mexErrMsgTxt("Vector length must be greater than 1.");
    if ((m & (m - 1)) != 0)
        mexErrMsgTxt("Vector length must be power of 2.");
    if (mxIsComplex(prhs[0]))
        mexErrMsgTxt("Input must be real.");
    else if (mxIsSparse(prhs[0]))
        mexErrMsgTxt("Input must be a full matrix, not sparse.");
    else if (!mxIsDouble(prhs[0]))
        mexErrMsgTxt("Input must be of type double.");
    plhs[0] = mxCreateDoubleMatrix(m, n, mxREAL);
    if (n == 0) return;
    skm_mex_check(skm_hadamard(skm_mex_ctx(), (int64_t)m, (int64_t)n, mxGetPr(prhs[0]), mxGetPr(plhs[0])), plhs[0]);
}
