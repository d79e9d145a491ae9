// exact.cu -- IEEE-double kernels that reproduce the reference's arithmetic bit for bit:
// every sum is accumulated sequentially in stored order with separately rounded multiply
// and add (the stock `mex -O` x86-64 build has no FMA), sqrt is applied before comparing,
// and the minimum follows MATLAB's first-occurrence / NaN-skipping rule.
//   masked distance  : private/SparseMatrixMinusCluster.c:117-183
//   min              : private/findClusterAssignments.m:168-171
//   inner product    : private/SparseMatrixInnerProduct.c:87-100
//   column norms     : private/SparseMatrixColumnNormSq.c:71-77
// These back the level-1 MEX replacements, SKM_F64 datasets, the sparse-centres branch and
// the re-evaluation of columns the fast fp32 kernel could not certify.
#include "common.cuh"

namespace {

template <typename VT>
__device__ __forceinline__ double masked_sum(const ExactArgs &a, int64_t j, int64_t k)
{
    const VT *val = (const VT *)a.val;
    const int64_t t0 = a.colptr[j], t1 = a.colptr[j + 1];
    const int64_t K = a.K;
    double s = 0.0;
    if (a.mask == nullptr) {
        for (int64_t t = t0; t < t1; ++t) {
            const int64_t r = a.rowidx[t];
            const double d = __dsub_rn((double)val[t], a.ct[r * K + k]);
            s = __dadd_rn(s, __dmul_rn(d, d));
        }
    } else {
        const double xd = a.xdiv[k];
        for (int64_t t = t0; t < t1; ++t) {
            const int64_t r = a.rowidx[t];
            if (!a.mask[r * K + k]) continue;
            const double xv = __ddiv_rn((double)val[t], xd);     // X(ind,:)/gamma_center
            const double d = __dsub_rn(xv, a.ct[r * K + k]);
            s = __dadd_rn(s, __dmul_rn(d, d));
        }
    }
    return s;
}

// one thread per (centre, column); k fastest so a warp shares its columns' entries
template <typename VT>
__global__ void k_exact_dist(ExactArgs a, int64_t j0, int64_t count, double *__restrict__ dist)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    int64_t j = j0 + idx / a.K, k = idx % a.K;
    dist[idx] = __dsqrt_rn(masked_sum<VT>(a, j, k));
}

// beta variant, one centre: sum of x*x + (-2beta)*x*c + c*c, left to right
template <typename VT>
__global__ void k_exact_dist_beta(ExactArgs a, double b, double *__restrict__ dist)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n) return;
    const VT *val = (const VT *)a.val;
    double s = 0.0;
    for (int64_t t = a.colptr[j]; t < a.colptr[j + 1]; ++t) {
        const double xv = (double)val[t], cv = a.ct[(int64_t)a.rowidx[t]];
        double term = __dadd_rn(__dmul_rn(xv, xv), __dmul_rn(__dmul_rn(b, xv), cv));
        term = __dadd_rn(term, __dmul_rn(cv, cv));
        s = __dadd_rn(s, term);
    }
    dist[j] = __dsqrt_rn(s);
}

// one warp per column (or per entry of `subset`): lanes stride over the centres.
// lb (optional): a lower bound on the distance to every centre but the winner (the second-smallest
// distance, rounded down) -- the state of the bounded assignment (bounded.cu).
template <typename VT>
__global__ void k_exact_assign(ExactArgs a, int32_t *__restrict__ assign, double *__restrict__ dist64,
                               float *__restrict__ dist32, const int32_t *__restrict__ subset,
                               const int *__restrict__ subset_count, int64_t total, float *__restrict__ lb)
{
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (subset_count) total = min(total, (int64_t)*subset_count);
    const double DINF = __longlong_as_double(0x7ff0000000000000LL);
    for (; warp < total; warp += nwarps) {
        const int64_t j = subset ? (int64_t)subset[warp] : warp;
        double bv = 0.0, sv = DINF;
        int bk = -1;
        for (int64_t k = lane; k < a.K; k += 32) {
            const double v = __dsqrt_rn(masked_sum<VT>(a, j, k));
            if (v != v) continue;                       // MATLAB min skips NaN
            if (bk < 0) { bv = v; bk = (int)k; }
            else if (v < bv) { sv = bv; bv = v; bk = (int)k; }
            else if (v < sv) sv = v;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const double osv = __shfl_xor_sync(0xffffffffu, sv, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ok >= 0) {
                if (bk < 0) { bv = ov; bk = ok; sv = osv; }
                else {
                    const bool other = ov < bv || (ov == bv && ok < bk);
                    const double loser = other ? bv : ov;
                    sv = fmin(fmin(sv, osv), loser);
                    if (other) { bv = ov; bk = ok; }
                }
            }
        }
        if (lane == 0) {
            if (bk < 0) { bv = __longlong_as_double(0x7ff8000000000000LL); bk = 0; sv = 0.0; }   // all NaN
            assign[j] = bk;
            if (dist64) dist64[j] = bv;
            if (dist32) dist32[j] = (float)bv;
            if (lb) lb[j] = __double2float_rd(sv * (1.0 - 1e-9));
        }
    }
}

__global__ void k_inner_product(int64_t n, const int64_t *__restrict__ colptr,
                                const int32_t *__restrict__ rowidx, const double *__restrict__ val,
                                const double *__restrict__ c, double *__restrict__ inner,
                                double *__restrict__ normsq)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double ip = 0.0, nn = 0.0;
    for (int64_t t = colptr[j]; t < colptr[j + 1]; ++t) {
        const double xv = val[t];
        if (c) ip = __dadd_rn(ip, __dmul_rn(xv, c[rowidx[t]]));
        nn = __dadd_rn(nn, __dmul_rn(xv, xv));
    }
    if (inner) inner[j] = ip;
    if (normsq) normsq[j] = nn;
}

// centres (p x K column-major) -> ct (row-major [(p+1)][K], divided by gamma, last row zero),
// optional support mask of the RAW centres (find(centers(:,k))).
__global__ void k_prep_centers(int64_t p, int64_t K, const double *__restrict__ centers, int has_gamma,
                               double gamma, double *__restrict__ ct, uint8_t *__restrict__ mask)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t total = (p + 1) * K;
    if (idx >= total) return;
    int64_t r = idx / K, k = idx % K;
    if (r == p) { ct[idx] = 0.0; if (mask) mask[idx] = 0; return; }
    const double c = centers[k * p + r];
    ct[idx] = has_gamma ? __ddiv_rn(c, gamma) : c;
    if (mask) mask[idx] = (c != 0.0) ? 1 : 0;     // NaN counts as nonzero, as in MATLAB find()
}

// xdiv[k] = nnz(centers(:,k)) / p   (findClusterAssignments.m:67), or 1 without gamma
__global__ void k_center_nnz(int64_t p, int64_t K, const double *__restrict__ centers, int has_gamma,
                             double *__restrict__ xdiv)
{
    __shared__ int cnt[32];
    int64_t k = blockIdx.x;
    int c = 0;
    for (int64_t r = threadIdx.x; r < p; r += blockDim.x) c += (centers[k * p + r] != 0.0);
#pragma unroll
    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += cnt[w];
        xdiv[k] = has_gamma ? __ddiv_rn((double)tot, (double)p) : 1.0;
    }
}

}  // namespace

int skm_launch_exact_dist(skm_ctx *ctx, const ExactArgs &a, int64_t j0, int64_t j1, double *dist)
{
    int64_t count = (j1 - j0) * a.K;
    if (count <= 0) return SKM_OK;
    int64_t blocks = (count + 255) / 256;
    if (blocks > 0x7fffffffLL) { skm_set_error("exact_dist: too many elements in one launch"); return SKM_ERR_INVALID; }
    if (a.val_type == SKM_F32)
        k_exact_dist<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, j0, count, dist);
    else
        k_exact_dist<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, j0, count, dist);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_exact_dist_beta(skm_ctx *ctx, const ExactArgs &a, double beta, double *dist)
{
    if (a.n <= 0) return SKM_OK;
    int64_t blocks = (a.n + 255) / 256;
    double b = beta * -2.;                      // SparseMatrixMinusCluster.c:121
    if (a.val_type == SKM_F32)
        k_exact_dist_beta<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, b, dist);
    else
        k_exact_dist_beta<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, b, dist);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_exact_assign(skm_ctx *ctx, const ExactArgs &a, int32_t *assign, double *dist64,
                            float *dist32, const int32_t *subset, const int *subset_count_dev,
                            int64_t subset_max, float *lb)
{
    int64_t total = subset ? subset_max : a.n;
    if (total <= 0) return SKM_OK;
    int64_t blocks = (total * 32 + 255) / 256;
    int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    if (a.val_type == SKM_F32)
        k_exact_assign<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, assign, dist64, dist32, subset,
                                                                        subset_count_dev, total, lb);
    else
        k_exact_assign<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a, assign, dist64, dist32, subset,
                                                                         subset_count_dev, total, lb);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_inner_product(skm_ctx *ctx, int64_t n, const int64_t *colptr, const int32_t *rowidx,
                             const double *val, const double *c, double *inner, double *normsq)
{
    if (n <= 0) return SKM_OK;
    int64_t blocks = (n + 255) / 256;
    k_inner_product<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, colptr, rowidx, val, c, inner, normsq);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_prep_centers(skm_ctx *ctx, int64_t p, int64_t K, const double *centers, int has_gamma,
                            double gamma, double *ct, uint8_t *mask, double *xdiv)
{
    int64_t total = (p + 1) * K;
    int64_t blocks = (total + 255) / 256;
    k_prep_centers<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, K, centers, has_gamma, gamma, ct, mask);
    SKM_CHECK_LAUNCH(ctx);
    if (xdiv) {
        k_center_nnz<<<(unsigned)K, 256, 0, ctx->stream>>>(p, K, centers, has_gamma, xdiv);
        SKM_CHECK_LAUNCH(ctx);
    }
    return SKM_OK;
}
