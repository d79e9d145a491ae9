// update.cu -- K2 (per-cluster row sums S, support counts N, member counts, sum of squared
// distances) and K3 (centre finalisation, centre change) of the Lloyd iteration.
//   K2 replaces kmeans_sparsified.m:430-453  (sum(X(:,ind),2), sum(NormalizationMatrix(:,ind),2))
//   K3 replaces kmeans_sparsified.m:448,450,470-471
// The partials buffer is [S (p*K) | N (p*K) | counts (K) | sumsq (1)] in doubles, column-major
// per cluster (S[k*p + r]); it is what the multi-GPU all-reduce sums.
#include "common.cuh"

namespace {

// v1: one warp per column, lanes over the column's stored entries, fp64 atomics to L2.
template <typename VT>
__global__ void k_accumulate_csc(int64_t p, int64_t n, int64_t K, const int64_t *__restrict__ colptr,
                                 const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                                 const int32_t *__restrict__ assign, const float *__restrict__ dist32,
                                 const double *__restrict__ dist64, double *__restrict__ partials)
{
    double *S = partials, *N = partials + p * K, *counts = partials + 2 * p * K;
    double *sumsq = counts + K;
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double local_sq = 0.0;
    for (int64_t j = warp; j < n; j += nwarps) {
        const int64_t k = assign[j];
        if (k < 0 || k >= K) continue;
        const int64_t t0 = colptr[j], t1 = colptr[j + 1];
        for (int64_t t = t0 + lane; t < t1; t += 32) {
            const int64_t r = rowidx[t];
            atomicAdd(&S[k * p + r], (double)val[t]);
            atomicAdd(&N[k * p + r], 1.0);
        }
        if (lane == 0) {
            atomicAdd(&counts[k], 1.0);
            const double d = dist64 ? dist64[j] : (double)dist32[j];
            local_sq += d * d;
        }
    }
    // block reduction of the squared distances (lane 0 of each warp holds a partial)
    __shared__ double red[32];
    if (lane == 0) red[threadIdx.x >> 5] = local_sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        if (s != 0.0 || s != s) atomicAdd(sumsq, s);
    }
}

// stats[0] += sum (old-new)^2 ; stats[1] = 1 if any NaN in new centres
__global__ void k_finalize(int64_t p, int64_t K, const double *__restrict__ partials, double gamma,
                           int ml, double *__restrict__ centers, double *__restrict__ centers_old,
                           double *__restrict__ stats)
{
    const double *S = partials, *N = partials + p * K, *counts = partials + 2 * p * K;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d2 = 0.0;
    int nan = 0;
    if (idx < p * K) {
        const int64_t k = idx / p;
        const double old = centers[idx];
        double nw = old;
        const double cnt = counts[k];
        if (cnt > 0.0) {
            if (ml) nw = __ddiv_rn(__dmul_rn(gamma, S[idx]), __dadd_rn(N[idx], 1e-16));
            else    nw = __ddiv_rn(S[idx], cnt);
        }
        centers_old[idx] = old;
        centers[idx] = nw;
        const double d = old - nw;
        d2 = d * d;
        nan = (nw != nw);
    }
    __shared__ double red[32];
    __shared__ int rnan;
    if (threadIdx.x == 0) rnan = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
    if (nan) rnan = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(&stats[0], s);
        if (rnan) stats[1] = 1.0;
    }
}

// centre change without an update (after the host patched columns for EmptyAction)
__global__ void k_diff(int64_t total, const double *__restrict__ centers,
                       const double *__restrict__ centers_old, double *__restrict__ stats)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d2 = 0.0;
    int nan = 0;
    if (idx < total) {
        const double d = centers_old[idx] - centers[idx];
        d2 = d * d;
        nan = (centers[idx] != centers[idx]);
    }
    __shared__ double red[32];
    __shared__ int rnan;
    if (threadIdx.x == 0) rnan = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
    if (nan) rnan = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(&stats[0], s);
        if (rnan) stats[1] = 1.0;
    }
}

// first index attaining the maximum (MATLAB max): per-block candidates, then one block
template <typename T>
__global__ void k_argmax_blocks(int64_t n, const T *__restrict__ d, double *__restrict__ bval,
                                int64_t *__restrict__ bidx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double bv = 0.0;
    int64_t bi = -1;
    for (; i < n; i += stride) {
        const double v = (double)d[i];
        if (v != v) continue;                                  // max skips NaN
        if (bi < 0 || v > bv) { bv = v; bi = i; }
    }
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x >> 1; o; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double ov = sv[threadIdx.x + o];
            const int64_t oi = si[threadIdx.x + o];
            const double mv = sv[threadIdx.x];
            const int64_t mi = si[threadIdx.x];
            if (oi >= 0 && (mi < 0 || ov > mv || (ov == mv && oi < mi))) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { bval[blockIdx.x] = sv[0]; bidx[blockIdx.x] = si[0]; }
}

__global__ void k_argmax_final(int nb, const double *__restrict__ bval, const int64_t *__restrict__ bidx,
                               double *__restrict__ out_val, int64_t *__restrict__ out_idx)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double bv = 0.0;
    int64_t bi = -1;
    for (int b = 0; b < nb; ++b) {
        const double ov = bval[b];
        const int64_t oi = bidx[b];
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    *out_val = bi >= 0 ? bv : __longlong_as_double(0x7ff8000000000000LL);
    *out_idx = bi >= 0 ? bi : 0;
}

}  // namespace

int skm_launch_accumulate(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const int32_t *assign,
                          const float *dist32, const double *dist64, double *partials)
{
    const int64_t p = ds->p, n = ds->n;
    SKM_CUDA(cudaMemsetAsync(partials, 0, sizeof(double) * (size_t)(2 * p * K + K + 1), ctx->stream));
    if (n == 0) return SKM_OK;
    int64_t blocks = (n * 32 + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (ds->store_dtype == SKM_F32)
        k_accumulate_csc<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            p, n, K, ds->colptr, ds->rowidx, (const float *)ds->val, assign, dist32, dist64, partials);
    else
        k_accumulate_csc<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            p, n, K, ds->colptr, ds->rowidx, (const double *)ds->val, assign, dist32, dist64, partials);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_finalize(skm_ctx *ctx, int64_t p, int64_t K, const double *partials, double gamma,
                        int ml_correction, double *centers, double *centers_old, double *stats)
{
    SKM_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 8, ctx->stream));
    int64_t total = p * K;
    if (total == 0) return SKM_OK;
    int64_t blocks = (total + 255) / 256;
    if (partials)
        k_finalize<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, K, partials, gamma, ml_correction, centers,
                                                             centers_old, stats);
    else
        k_diff<<<(unsigned)blocks, 256, 0, ctx->stream>>>(total, centers, centers_old, stats);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_argmax(skm_ctx *ctx, int64_t n, const float *dist32, const double *dist64,
                      double *out_val, int64_t *out_idx)
{
    const int nb = 256;
    DevBuf bv, bi;
    SKM_TRY(bv.alloc(sizeof(double) * nb));
    SKM_TRY(bi.alloc(sizeof(int64_t) * nb));
    if (dist64)
        k_argmax_blocks<double><<<nb, 256, 0, ctx->stream>>>(n, dist64, bv.as<double>(), bi.as<int64_t>());
    else
        k_argmax_blocks<float><<<nb, 256, 0, ctx->stream>>>(n, dist32, bv.as<double>(), bi.as<int64_t>());
    SKM_CHECK_LAUNCH(ctx);
    k_argmax_final<<<1, 32, 0, ctx->stream>>>(nb, bv.as<double>(), bi.as<int64_t>(), out_val, out_idx);
    SKM_CHECK_LAUNCH(ctx);
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));   // temporaries are freed on return
    return SKM_OK;
}
