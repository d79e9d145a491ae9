// kpp.cu -- K5: the distance part of k-means++ (private/Arthur_initialization.m:39-53).
// The reference recomputes the distance of every point to ALL chosen centres each round and
// takes the minimum; min is exact, so folding the distance to the newest centre into a
// running minimum gives bit-identical values with one centre-pass per round.  Distances are
// evaluated in fp64 in the reference's order (SparseMatrixMinusCluster.c:133-141, K = 1).
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>

namespace {

template <typename VT>
__global__ void k_kpp_update(int64_t n, const int64_t *__restrict__ colptr,
                             const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                             const double *__restrict__ c, int first, int masked, double *__restrict__ mind,
                             const int32_t *__restrict__ subset, const int *__restrict__ subset_count)
{
    // thread per column, grid-stride.  Measured 2.3-2.4 TB/s algorithmic whatever the number of resident warps
    // (8 ... 64 per SM, tools/debug/kpp_probe.py): the walk is a dependent chain per thread (index load -> centre
    // gather -> three fp64 operations in the reference's order), not a capacity problem.  A staged variant
    // (warp-cooperative coalesced pass writing the rounded squares to shared memory, then per-lane ordered sums)
    // was bit-identical but 3.6x slower: 20 KB of squares per warp leaves 8 warps per SM and too few bytes in
    // flight; it would need bulk asynchronous copies of the slice to pay off.
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t total = subset ? (int64_t)*subset_count : n;           // subset: the columns the fp32 filter could not skip
    for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < total; f += stride) {
        const int64_t j = subset ? (int64_t)subset[f] : f;
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

// ---- fp32 filter in front of the exact pass (rounds after the first) ----
// Adding a centre changes the running minimum of few columns (about n/k in round k); for the others it is enough to
// PROVE that the new centre is farther than the current minimum.  One pass over the streamed SELL image (8 B/entry,
// coalesced, the centre as an fp32 row in shared memory) forms the same fp32 sum K1 would, subtracts K1's rounding
// guard and keeps the column out of the exact pass iff that lower bound, with a 1e-6 relative
// margin (the reference's own fp64 rounding is 1e-13), exceeds the stored minimum.  The rest is evaluated by
// k_kpp_update in fp64 in the reference's order, so every stored value stays bit-identical to the reference's.
struct KppFilterParams {
    const int4    *sell;
    const int64_t *slice_ptr;
    int64_t        nslices, n;
    int            uniform, width2, p, boff;
    const float   *c32;            // [p+1] fp32 centre (scaled), c32[p] = 0; then cmax
    float          ga, gb_unit, ge_unit;
    const double  *mind;
    int32_t       *flagged;
    int           *nflag;
};

__device__ __forceinline__ int4 kpp_ld_stream(const int4 *p)
{
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

__global__ void __launch_bounds__(512) k_kpp_filter(const KppFilterParams P)
{
    extern __shared__ __align__(16) float kt[];
    for (int i = threadIdx.x; i <= P.p; i += blockDim.x) kt[i] = P.c32[i];
    __syncthreads();
    const float cm = P.c32[P.p + 1];
    const float gb = P.gb_unit * cm, ge = P.ge_unit * cm * cm + 1e-37f;
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t slice = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; slice < P.nslices; slice += warps_total) {
        int64_t base;
        int w2;
        if (P.uniform) { base = slice * (int64_t)P.width2 * 32; w2 = P.width2; }
        else { base = P.slice_ptr[slice]; w2 = (int)((P.slice_ptr[slice + 1] - base) >> 5); }
        const int4 *src = P.sell + base + lane;
        float acc = 0.f;
        auto one = [&](int r, float x) {
            if (P.boff && r >= P.boff) r -= P.boff;          // second table copy of the dual layouts
            if (r > P.p) r = P.p;                            // any pad row -> the zero row
            const float d = x - kt[r];
            acc = fmaf(d, d, acc);
        };
        int t2 = 0;
        int4 n0, n1, n2, n3;
        if (w2 >= 4) { n0 = kpp_ld_stream(src); n1 = kpp_ld_stream(src + 32); n2 = kpp_ld_stream(src + 64); n3 = kpp_ld_stream(src + 96); }
        for (; t2 + 4 <= w2; t2 += 4) {
            const int4 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
            if (t2 + 8 <= w2) {
                n0 = kpp_ld_stream(src + (t2 + 4) * 32); n1 = kpp_ld_stream(src + (t2 + 5) * 32);
                n2 = kpp_ld_stream(src + (t2 + 6) * 32); n3 = kpp_ld_stream(src + (t2 + 7) * 32);
            }
            one(q0.x, __int_as_float(q0.y)); one(q0.z, __int_as_float(q0.w));
            one(q1.x, __int_as_float(q1.y)); one(q1.z, __int_as_float(q1.w));
            one(q2.x, __int_as_float(q2.y)); one(q2.z, __int_as_float(q2.w));
            one(q3.x, __int_as_float(q3.y)); one(q3.z, __int_as_float(q3.w));
        }
        for (; t2 < w2; ++t2) {
            const int4 q = kpp_ld_stream(src + t2 * 32);
            one(q.x, __int_as_float(q.y)); one(q.z, __int_as_float(q.w));
        }
        const int64_t j = slice * SKM_SLICE + lane;
        if (j >= P.n) continue;
        const float E = P.ga * acc + gb * sqrtf(acc) + ge;
        const float lo = sqrtf(fmaxf(acc - E, 0.f)) * (1.f - 1.0e-6f);   // <= the distance to the new centre
        if (!((double)lo > P.mind[j])) {                                 // cannot be skipped (also NaN / Inf cases)
            const int slot = atomicAdd(P.nflag, 1);
            P.flagged[slot] = (int32_t)j;
        }
    }
}

// c32[r] = (float)c[r] (r < p), c32[p] = 0, c32[p+1] = max |c32| (NaN propagates)
__global__ void k_kpp_centre32(int64_t p, const double *__restrict__ c, float *__restrict__ c32)
{
    __shared__ int smax;
    if (threadIdx.x == 0) smax = 0;
    __syncthreads();
    int mi = 0;
    for (int64_t r = threadIdx.x; r <= p; r += blockDim.x) {
        const float v = r < p ? (float)c[r] : 0.f;
        c32[r] = v;
        float m = fabsf(v);
        if (v != v) m = __int_as_float(0x7fc00000);
        mi = max(mi, __float_as_int(m));
    }
    atomicMax(&smax, mi);
    __syncthreads();
    if (threadIdx.x == 0) c32[p + 1] = __int_as_float(smax);
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
                                                                      (const float *)ds->val, c_scaled, first, masked, mind, nullptr, nullptr);
    else
        k_kpp_update<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->n, ds->colptr, ds->rowidx,
                                                                       (const double *)ds->val, c_scaled, first, masked, mind, nullptr, nullptr);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

bool skm_kpp_filter_usable(const skm_ctx *ctx, const skm_dataset *ds)
{
    return ds->store_dtype == SKM_F32 && ds->sell && ds->sell_elems > 0 && ds->n >= (1 << 16) &&
           (size_t)(ds->p + 2) * sizeof(float) + 1024 <= (size_t)ctx->smem_optin && !getenv("SKM_NO_KPP_FILTER");
}

// fp32 filter + exact pass over what it could not skip; c32 is scratch of p + 2 floats, flagged of n int32
int skm_launch_kpp_update_filtered(skm_ctx *ctx, const skm_dataset *ds, const double *c_scaled, double *mind,
                                   float *c32, int32_t *flagged, int *nflag)
{
    if (ds->n == 0) return SKM_OK;
    k_kpp_centre32<<<1, 1024, 0, ctx->stream>>>(ds->p, c_scaled, c32);
    SKM_CHECK_LAUNCH(ctx);
    SKM_CUDA(cudaMemsetAsync(nflag, 0, sizeof(int), ctx->stream));
    const double u = 5.9604644775390625e-08;
    const double m = (double)(ds->max_col_nnz > 0 ? ds->max_col_nnz : 1);
    KppFilterParams P;
    P.sell = ds->sell; P.slice_ptr = ds->slice_ptr; P.nslices = ds->nslices; P.n = ds->n;
    P.uniform = ds->uniform_width ? 1 : 0; P.width2 = ds->sell_width2;
    P.p = (int)ds->p;
    P.boff = (!ds->sell_plain && ds->sell_mode >= 1) ? (int)skm_dual_boff(ds->p) : 0;
    P.c32 = c32;
    P.ga = (float)(1.01 * (m + 5.0) * u);
    P.gb_unit = (float)(2.02 * u * sqrt(m));
    P.ge_unit = (float)(2.1 * u * u * m);
    P.mind = mind; P.flagged = flagged; P.nflag = nflag;
    const size_t smem = (size_t)(ds->p + 2) * sizeof(float);
    SKM_CUDA(cudaFuncSetAttribute(k_kpp_filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_kpp_filter, 512, smem));
    if (per_sm < 1) { skm_set_error("kpp filter does not fit on an SM"); return SKM_ERR_UNSUPPORTED; }
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    const int64_t need = (ds->nslices * 32 + 511) / 512;
    if (blocks > need) blocks = need;
    k_kpp_filter<<<(unsigned)blocks, 512, smem, ctx->stream>>>(P);
    SKM_CHECK_LAUNCH(ctx);
    const int64_t eb = std::min<int64_t>((ds->n + 255) / 256, (int64_t)ctx->sm_count * 8);
    k_kpp_update<float><<<(unsigned)eb, 256, 0, ctx->stream>>>(ds->n, ds->colptr, ds->rowidx, (const float *)ds->val, c_scaled, 0, 0,
                                                              mind, flagged, nflag);
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
