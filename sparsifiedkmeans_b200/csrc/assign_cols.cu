// assign_cols.cu -- every centre for a LIST of columns, fp32 with the K1 guard: the re-evaluation of the columns a pruned
// or bounded assignment pass could not keep (api.cu: reevaluate_flagged), in front of the fp64 kernel of exact.cu.
//
// Same arithmetic as K1 (assign_fast.cu; private/SparseMatrixMinusCluster.c:169-182 followed by min,
// private/findClusterAssignments.m:168-171): d = x - fl32(c'), one FMA per term, fp32 sums, and the same rigorous guard
// decides whether the fp32 winner is certainly the reference's; what it cannot certify is appended to a list for fp64.
//
// Mapping: one warp per listed column, lanes over the centres (K <= 128: up to four per lane).  The column's entries are
// read from the column-major image (coalesced, 8 bytes per entry -- nothing but the useful bytes, unlike a lane-per-column
// walk of the SELL image, whose 16-byte loads 512 bytes apart cost 64-byte DRAM bursts: profiles/r2_prune.md) and handed
// round with shuffles; the centre values come from a row-major fp32 table [p + 1][kpad] in global memory, one coalesced
// 128-byte line per 32 centres and entry, which the L1 holds (262 KB at K = 64, p = 1024; the kernel uses no shared memory).
#include "common.cuh"
#include <math.h>

namespace {

struct ColsParams {
    const int64_t *colptr;
    const int32_t *rowidx;
    const float   *val;
    int64_t        n;
    const float   *table;      // [p + 1][kpad], row p zero, columns >= K zero
    int            kpad, K;
    const int32_t *list;
    int64_t        nlist;
    float          ga, gb_unit, ge_unit;
    const float   *cmax;
    int32_t       *assign;
    float         *dist, *lb;
    int32_t       *flagged;
    int           *nflag;
};

template <int KL>
__global__ void __launch_bounds__(256) k_assign_cols(const ColsParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float INF = __int_as_float(0x7f800000), QNAN = __int_as_float(0x7fc00000);
    const float cm = *P.cmax;
    const float gb = P.gb_unit * cm, ge = P.ge_unit * cm * cm + 1e-37f;
    for (int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < P.nlist; w += nwarps) {
        const int64_t j = P.list[w];
        if (j < 0 || j >= P.n) continue;
        const int64_t t0 = P.colptr[j], t1 = P.colptr[j + 1];
        float acc[KL];
#pragma unroll
        for (int q = 0; q < KL; ++q) acc[q] = 0.f;
        for (int64_t base = t0; base < t1; base += 32) {
            const int cnt = (int)min((int64_t)32, t1 - base);
            int my_r = 0;
            float my_x = 0.f;
            if (lane < cnt) { my_r = __ldg(P.rowidx + base + lane); my_x = __ldg(P.val + base + lane); }
            int e = 0;
            for (; e + 2 <= cnt; e += 2) {                         // two entries per trip: their table loads overlap
                const int r0 = __shfl_sync(0xffffffffu, my_r, e), r1 = __shfl_sync(0xffffffffu, my_r, e + 1);
                const float x0 = __shfl_sync(0xffffffffu, my_x, e), x1 = __shfl_sync(0xffffffffu, my_x, e + 1);
                const float *row0 = P.table + (int64_t)r0 * P.kpad + lane, *row1 = P.table + (int64_t)r1 * P.kpad + lane;
                float v0[KL], v1[KL];
#pragma unroll
                for (int q = 0; q < KL; ++q) { v0[q] = __ldg(row0 + 32 * q); v1[q] = __ldg(row1 + 32 * q); }
#pragma unroll
                for (int q = 0; q < KL; ++q) {
                    float d = x0 - v0[q]; acc[q] = fmaf(d, d, acc[q]);
                    d = x1 - v1[q]; acc[q] = fmaf(d, d, acc[q]);
                }
            }
            if (e < cnt) {
                const int r0 = __shfl_sync(0xffffffffu, my_r, e);
                const float x0 = __shfl_sync(0xffffffffu, my_x, e);
                const float *row0 = P.table + (int64_t)r0 * P.kpad + lane;
#pragma unroll
                for (int q = 0; q < KL; ++q) { const float d = x0 - __ldg(row0 + 32 * q); acc[q] = fmaf(d, d, acc[q]); }
            }
        }
        // ---- best / second best: this lane's centres, then across the lanes (smaller index wins ties, as K1's scan does) ----
        float b1 = INF, b2 = INF;
        int i1 = 0x7fffffff;
        bool bad = false;
#pragma unroll
        for (int q = 0; q < KL; ++q) {
            const int k = 32 * q + lane;
            if (k < P.K) {
                const float v = acc[q];
                if (!(v < INF)) bad = true;                        // NaN or overflow: cannot certify
                if (v < b1) { b2 = b1; b1 = v; i1 = k; }
                else if (v < b2) b2 = v;
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ob1 = __shfl_xor_sync(0xffffffffu, b1, o), ob2 = __shfl_xor_sync(0xffffffffu, b2, o);
            const int oi1 = __shfl_xor_sync(0xffffffffu, i1, o);
            const bool other = ob1 < b1 || (ob1 == b1 && oi1 < i1);
            const float loser = other ? b1 : ob1;
            b2 = fminf(fminf(b2, ob2), loser);
            if (other) { b1 = ob1; i1 = oi1; }
        }
        bad = __any_sync(0xffffffffu, bad);
        if (lane != 0) continue;
        if (i1 == 0x7fffffff) i1 = 0;                               // every sum NaN / inf
        if (bad) b2 = QNAN;
        // ---- guard: |fp32 sum - exact sum| <= E(s) = ga*s + gb*sqrt(s) + ge  (as K1) ----
        bool certified;
        if (P.K == 1) certified = (b1 < INF);
        else {
            const float E = P.ga * (b1 + b2) + gb * (sqrtf(b1) + sqrtf(b2)) + 2.f * ge;
            certified = (b2 - b1) > E;                             // false for NaN / inf
        }
        P.assign[j] = i1;
        P.dist[j] = sqrtf(b1);
        if (P.lb) {
            float lbv = 0.f;
            if (P.K == 1) lbv = INF;
            else if (b2 == b2) {
                const float e2 = P.ga * b2 + gb * sqrtf(b2) + ge;
                lbv = sqrtf(fmaxf(b2 - e2, 0.f)) * (1.f - 4.76837158203125e-07f);
            }
            P.lb[j] = lbv;
        }
        if (!certified) {
            const int slot = atomicAdd(P.nflag, 1);
            P.flagged[slot] = (int32_t)j;
        }
    }
}

__global__ void k_build_table_rm(int64_t p, int64_t K, int kpad, const double *__restrict__ ct /* [p + 1][K] */,
                                 float *__restrict__ table, float *__restrict__ cmax)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float m = 0.f;
    if (idx < (p + 1) * kpad) {
        const int64_t r = idx / kpad, k = idx % kpad;
        const float v = (r < p && k < K) ? (float)ct[r * K + k] : 0.f;
        table[idx] = v;
        m = fabsf(v);
        if (v != v) m = __int_as_float(0x7fc00000);
    }
    int mi = __float_as_int(m);
#pragma unroll
    for (int o = 16; o; o >>= 1) mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    if ((threadIdx.x & 31) == 0 && mi > 0) atomicMax(reinterpret_cast<int *>(cmax), mi);
}

}  // namespace

int skm_assign_cols_kpad(int64_t K) { return (int)((K + 31) / 32 * 32); }

bool skm_assign_cols_supported(const skm_dataset *ds, int64_t K)
{
    return ds->store_dtype == SKM_F32 && K >= 1 && K <= 128 && ds->colptr && ds->rowidx && ds->val;
}

int skm_launch_build_table_rm(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, float *table, float *cmax)
{
    SKM_CUDA(cudaMemsetAsync(cmax, 0, sizeof(float), ctx->stream));
    const int kpad = skm_assign_cols_kpad(K);
    const int64_t total = (p + 1) * kpad;
    k_build_table_rm<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p, K, kpad, ct, table, cmax);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_assign_cols(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const float *table_rm, const float *cmax,
                           const int32_t *list, int64_t nlist, int32_t *assign, float *dist, float *lb,
                           int32_t *flagged_out, int *nflag_out)
{
    SKM_CUDA(cudaMemsetAsync(nflag_out, 0, sizeof(int), ctx->stream));
    if (nlist <= 0 || ds->n == 0) return SKM_OK;
    if (!skm_assign_cols_supported(ds, K)) { skm_set_error("assign_cols: needs an SKM_F32 dataset and K <= 128"); return SKM_ERR_UNSUPPORTED; }
    const double u = 5.9604644775390625e-08;
    const double m = (double)(ds->max_col_nnz > 0 ? ds->max_col_nnz : 1);
    ColsParams P;
    P.colptr = ds->colptr; P.rowidx = ds->rowidx; P.val = (const float *)ds->val; P.n = ds->n;
    P.table = table_rm; P.kpad = skm_assign_cols_kpad(K); P.K = (int)K;
    P.list = list; P.nlist = nlist;
    P.ga = (float)(1.01 * (m + 5.0) * u);
    P.gb_unit = (float)(2.02 * u * sqrt(m));
    P.ge_unit = (float)(2.1 * u * u * m);
    P.cmax = cmax; P.assign = assign; P.dist = dist; P.lb = lb; P.flagged = flagged_out; P.nflag = nflag_out;
    int64_t blocks = (nlist * 32 + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    const int kl = P.kpad / 32;
    // no shared memory: leave the whole unified array to the L1, which holds the table
    auto launch = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        kern<<<(unsigned)blocks, 256, 0, ctx->stream>>>(P);
    };
    switch (kl) {
        case 1: launch(k_assign_cols<1>); break;
        case 2: launch(k_assign_cols<2>); break;
        case 3: launch(k_assign_cols<3>); break;
        default: launch(k_assign_cols<4>); break;
    }
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}
