/*
 * skm_mex_common.h -- shared helpers of the MEX shims over libskm_b200.so.
 *
 * Each shim in this directory is a drop-in for one MEX file of the reference
 * (same name, same usage, same error identifiers/messages) that only unpacks its
 * mxArray arguments and calls ONE function of include/skm_b200.h.  They compile
 * against MathWorks' mex.h (`mex -largeArrayDims -I../include X.c -L.. -lskm_b200`)
 * and, in this repository's tests, against the stub in oracle/mexstub/mex.h.
 *
 * The library never long-jumps; on a non-zero status the shim releases what it
 * holds and raises the MATLAB error itself.
 */
#ifndef SKM_MEX_COMMON_H
#define SKM_MEX_COMMON_H

#include <stdint.h>
#include "mex.h"
#include "skm_b200.h"

static skm_ctx *g_skm_ctx = NULL;

static void skm_mex_atexit(void)
{
    if (g_skm_ctx) { skm_ctx_destroy(g_skm_ctx); g_skm_ctx = NULL; }
}

/* one context per MATLAB process, created on first use on device 0 */
static skm_ctx *skm_mex_ctx(void)
{
    if (!g_skm_ctx) {
        if (skm_ctx_create(0, NULL, &g_skm_ctx) != SKM_OK)
            mexErrMsgIdAndTxt("skm_b200:noDevice", "%s", skm_last_error(NULL));
        mexAtExit(skm_mex_atexit);
    }
    return g_skm_ctx;
}

static void skm_mex_check(int rc, mxArray *owned)
{
    if (rc == SKM_OK) return;
    if (owned) mxDestroyArray(owned);
    mexErrMsgIdAndTxt(rc == SKM_ERR_INVALID ? "skm_b200:invalidInput" : "skm_b200:runtime", "%s",
                      skm_last_error(g_skm_ctx));
}

#endif
