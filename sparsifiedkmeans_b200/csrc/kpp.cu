// kpp.cu -- K5: the distance part of k-means++ (private/Arthur_initialization.m:39-53).
// The reference recomputes the distance of every point to ALL chosen centres each round and
// takes the minimum; min is exact, so folding the distance to the newest centre into a
// running minimum gives bit-identical values with one centre-pass per round.  Distances are
// evaluated in fp64 in the reference's order (SparseMatrixMinusCluster.c:133-141, K = 1).
#include "common.cuh"
#include <stdlib.h>
#include <vector>

namespace {

template <typename VT>
__global__ void k_kpp_update(int64_t n, const int64_t *__restrict__ colptr,
                             const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                             const double *__restrict__ c, int first, int masked, double *__restrict__ mind)
{
    // thread per column, grid-stride.  Measured 2.3-2.4 TB/s algorithmic whatever the number of resident warps
    // (8 ... 64 per SM, tools/debug/kpp_probe.py): the walk is a dependent chain per thread (index load -> centre
    // gather -> three fp64 operations in the reference's order), not a capacity problem.  A staged variant
    // (warp-cooperative coalesced pass writing the rounded squares to shared memory, then per-lane ordered sums)
    // was bit-identical but 3.6x slower: 20 KB of squares per warp leaves 8 warps per SM and too few bytes in
    // flight; it would need bulk asynchronous copies of the slice to pay off.
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        double s = 0.0;
        for (int64_t t = colptr[j]; t < colptr[j + 1]; ++t) {
            const double cv = c[rowidx[t]];
            // sparse-centre form (findClusterAssignments.m:70-74, Arthur_initialization.m:26 without gamma): the
            // sum runs over supp(x_j) /\ supp(c) only -- X(ind,:) keeps the ascending row order
            if (masked && cv == 0.0) continue;
            const double d = __dsub_rn((double)val[t], cv);
            s = __dadd_rn(s, __dmul_rn(d, d));
        }
        const double d = __dsqrt_rn(s);
        if (first) mind[j] = d;
        else {
            const double o = mind[j];
            // MATLAB min ignores NaN unless both are NaN
            mind[j] = (d != d) ? o : ((o != o) ? d : (d < o ? d : o));
        }
    }
}

// deterministic block sums of mind^2 over fixed blocks of 1024 columns
__global__ void k_block_sumsq(int64_t n, const double *__restrict__ mind, double *__restrict__ bsum)
{
    __shared__ double s[1024];
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    double v = 0.0;
    if (i < n) { v = mind[i]; v = v * v; }
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 512; o; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) bsum[blockIdx.x] = s[0];
}

}  // namespace

int skm_launch_kpp_update(skm_ctx *ctx, const skm_dataset *ds, const double *c_scaled, int first,
                          double *mind, int masked)
{
    if (ds->n == 0) return SKM_OK;
    int64_t blocks = (ds->n + 255) / 256;
    {
        static const char *e = getenv("SKM_KPP_CTAS");           // tuning knob: resident 256-thread CTAs per SM
        const int per_sm = e ? atoi(e) : 0;
        if (per_sm > 0 && blocks > (int64_t)ctx->sm_count * per_sm) blocks = (int64_t)ctx->sm_count * per_sm;
    }
    if (ds->store_dtype == SKM_F32)
        k_kpp_update<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->n, ds->colptr, ds->rowidx,
                                                                      (const float *)ds->val, c_scaled, first, masked, mind);
    else
        k_kpp_update<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->n, ds->colptr, ds->rowidx,
                                                                       (const double *)ds->val, c_scaled, first, masked, mind);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// cum[b] = sum of mind^2 over blocks 0..b (host-side sequential sum of the block sums)
int skm_launch_scan_sq(skm_ctx *ctx, int64_t n, const double *mind, double *bsum_dev)
{
    if (n == 0) return SKM_OK;
    int64_t nb = (n + 1023) / 1024;
    k_block_sumsq<<<(unsigned)nb, 1024, 0, ctx->stream>>>(n, mind, bsum_dev);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}
