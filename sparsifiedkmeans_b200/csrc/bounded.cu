// bounded.cu -- bounded assignment (opt-in): exact pruning of the K1 pass with bounds carried across
// Lloyd iterations (Hamerly 2010, adapted to the masked distance).
//
// Why it is needed: at K = 64 every stored entry (8 B from HBM) needs its row of 64 centre values,
// 256 B, through the shared-memory pipe -- 4.5 ms per 1.25e7-column shard against 0.8 ms of HBM time,
// whatever the kernel does.  Once the centres move little, almost every column
// keeps its centre, and proving that needs ONE centre value per entry.
//
// State per column j: its assignment a_j and lb_j <= min_{k != a_j} d_j(c_k), where d_j(c) =
// ||M_j (x_j - c)|| is the reference's masked distance (M_j = the column's support).  d_j is a seminorm
// of the centre, so when the centres move, |d_j(c_new) - d_j(c_old)| <= ||M_j (c_new - c_old)|| <=
// ||c_new - c_old||: with shift_k = ||c'_k,new - c'_k,old|| (c' = C/gamma, fp64, rounded up),
//       lb_j  <-  lb_j - max_{k != a_j} shift_k        stays a valid lower bound.
// Each pass streams the SELL image once, gathers the ONE value c'[row, a_j] per entry (4 B instead of
// 4K), gets u_j = d_j(c_{a_j}) exactly as K1 would (same fp32 sum, same rounding guard), and keeps the
// assignment iff u_j + guard < lb_j, with a relative margin (1e-6) that dwarfs the reference's own fp64
// rounding, so a kept assignment is the reference's argmin.  Columns that fail are flagged; the caller
// re-evaluates them against every centre (fp64 in the reference's order when few, the full K1 pass when
// many), which also refreshes their lb.  Distances returned for kept columns are the same fp32 sums K1
// returns.  Bound: HBM (one pass over the 8 B/entry stream); the table is one row per centre, resident
// in shared memory for as many centres as fit, the rest is read through L1/L2.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace {

struct BoundedParams {
    const int4    *sell;
    const int64_t *slice_ptr;
    int64_t        nslices, n;
    int            uniform, width2;
    int            p, boff, K, ksm;      // ksm: centres whose rows are staged in shared memory
    const float   *table_t;              // [K][p+1]
    float          ga, gb_unit, ge_unit;
    const float   *cmax;
    const float   *shift;                // [K] then max, second max, argmax (as float)
    const int32_t *assign;
    float         *lb, *dist;
    int32_t       *flagged;
    int           *nflag;
};

// PADTAIL: the last batch of a column is padded with (zero row, 0) entries instead of a serial tail.  It pays
// when part of the table is read through L2 (K = 64: 1.52 -> 1.39 ms) and costs when everything is in shared
// memory and the kernel already runs at the HBM rate (K = 10: 0.98 -> 1.07 ms), so the launcher picks.
// 128-bit streaming load that bypasses L1 allocation: whatever L1 the shared-memory carve-out leaves belongs to the
// table rows that do not fit in shared memory (K = 64: 9 of 64), which would otherwise be evicted by the stream
__device__ __forceinline__ int4 ld_stream(const int4 *p)
{
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int THREADS, bool PADTAIL>
__global__ void __launch_bounds__(THREADS) k_assign_bounded(const BoundedParams P)
{
    extern __shared__ __align__(16) float s_tab[];
    const int stride = P.p + 1;
    {
        const int64_t total = (int64_t)P.ksm * stride;
        for (int64_t i = threadIdx.x; i < total; i += THREADS) s_tab[i] = P.table_t[i];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * THREADS) >> 5;
    int64_t slice = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const float dmax = P.shift[P.K], dsec = P.shift[P.K + 1];
    const int imax = (int)P.shift[P.K + 2];
    const float cm = *P.cmax;
    const float gb = P.gb_unit * cm, ge = P.ge_unit * cm * cm + 1e-37f;

    for (; slice < P.nslices; slice += warps_total) {
        int64_t base;
        int w2;
        if (P.uniform) { base = slice * (int64_t)P.width2 * 32; w2 = P.width2; }
        else { base = P.slice_ptr[slice]; w2 = (int)((P.slice_ptr[slice + 1] - base) >> 5); }
        const int4 *src = P.sell + base + lane;
        const int64_t j = slice * SKM_SLICE + lane;
        const bool live = j < P.n;
        int a = live ? P.assign[j] : 0;
        if ((unsigned)a >= (unsigned)P.K) a = 0;
        const bool in_smem = a < P.ksm;
        const float *row = in_smem ? (s_tab + (int64_t)a * stride) : (P.table_t + (int64_t)a * stride);
        float acc = 0.f;
        auto one = [&](int r, float x) {
            if (P.boff && r >= P.boff) r -= P.boff;          // second table copy of the dual layouts
            if (r > P.p) r = P.p;                            // any pad row -> the zero row
            const float v = in_smem ? row[r] : __ldg(row + r);
            const float d = x - v;
            acc = fmaf(d, d, acc);
        };
        // software pipeline: the next four 16-byte loads are in flight while the current eight entries are
        // gathered (ncu before this: 41 cycles of long-scoreboard stall per issue at K=64, 32 resident warps --
        // a latency-bound kernel needs more bytes in flight per warp, not more warps)
        if (PADTAIL) {
            const int4 padq = make_int4(P.p, 0, P.p, 0);
            auto fetch = [&](int t) -> int4 { return t < w2 ? ld_stream(src + t * 32) : padq; };
            int4 n0 = fetch(0), n1 = fetch(1), n2 = fetch(2), n3 = fetch(3);
            for (int t2 = 0; t2 < w2; t2 += 4) {
                const int4 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
                n0 = fetch(t2 + 4); n1 = fetch(t2 + 5); n2 = fetch(t2 + 6); n3 = fetch(t2 + 7);
                one(q0.x, __int_as_float(q0.y)); one(q0.z, __int_as_float(q0.w));
                one(q1.x, __int_as_float(q1.y)); one(q1.z, __int_as_float(q1.w));
                one(q2.x, __int_as_float(q2.y)); one(q2.z, __int_as_float(q2.w));
                one(q3.x, __int_as_float(q3.y)); one(q3.z, __int_as_float(q3.w));
            }
        } else {
            int t2 = 0;
            int4 n0, n1, n2, n3;
            if (w2 >= 4) {
                n0 = ld_stream(src + 0 * 32); n1 = ld_stream(src + 1 * 32); n2 = ld_stream(src + 2 * 32); n3 = ld_stream(src + 3 * 32);
            }
            for (; t2 + 4 <= w2; t2 += 4) {
                const int4 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
                if (t2 + 8 <= w2) {
                    n0 = ld_stream(src + (t2 + 4) * 32); n1 = ld_stream(src + (t2 + 5) * 32);
                    n2 = ld_stream(src + (t2 + 6) * 32); n3 = ld_stream(src + (t2 + 7) * 32);
                }
                one(q0.x, __int_as_float(q0.y)); one(q0.z, __int_as_float(q0.w));
                one(q1.x, __int_as_float(q1.y)); one(q1.z, __int_as_float(q1.w));
                one(q2.x, __int_as_float(q2.y)); one(q2.z, __int_as_float(q2.w));
                one(q3.x, __int_as_float(q3.y)); one(q3.z, __int_as_float(q3.w));
            }
            for (; t2 < w2; ++t2) {
                const int4 q = ld_stream(src + t2 * 32);
                one(q.x, __int_as_float(q.y)); one(q.z, __int_as_float(q.w));
            }
        }
        if (!live) continue;
        // bound after this move of the centres (rounded down), then the test with K1's rounding guard
        const float mv = (a == imax) ? dsec : dmax;
        float lbn = (P.lb[j] - mv) * (1.f - 4.76837158203125e-07f);
        if (!(lbn > 0.f)) lbn = 0.f;                          // also NaN -> 0
        const float E = P.ga * acc + gb * sqrtf(acc) + ge;
        const float uhi = sqrtf(acc + E) * (1.f + 1.0e-6f);
        P.lb[j] = lbn;
        if (uhi < lbn) P.dist[j] = sqrtf(acc);                // assignment kept (false for NaN / inf)
        else {
            const int slot = atomicAdd(P.nflag, 1);
            P.flagged[slot] = (int32_t)j;
        }
    }
}

// shift[k] = || (C_new(:,k) - C_prev(:,k)) / gamma ||_2 rounded up; C_prev <- C_new
__global__ void k_center_shift(int64_t p, const double *__restrict__ cnew, double *__restrict__ cprev,
                               int has_gamma, double gamma, float *__restrict__ shift)
{
    const int64_t k = blockIdx.x;
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < p; r += blockDim.x) {
        const double a = cnew[k * p + r], b = cprev[k * p + r];
        double d = has_gamma ? (__ddiv_rn(a, gamma) - __ddiv_rn(b, gamma)) : (a - b);
        s += d * d;
        cprev[k * p + r] = a;
    }
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        const double v = sqrt(t) * (1.0 + 1e-9);
        float f = __double2float_ru(v);
        if (!(f == f)) f = __int_as_float(0x7f800000);       // NaN centre: nothing can be kept
        shift[k] = f;
    }
}

// A-priori count of the columns the bounded pass is certain to keep: the masked distance is a seminorm of the
// centre, so the new distance to the own centre is at most dist_j + shift[a_j]; if even that stays below the
// lowered bound the column keeps its centre.  One pass over 12 bytes per column (no entry is read) tells the
// caller whether the bounded pass is worth launching at all after this move of the centres.
__global__ void k_bound_predict(int64_t n, int K, const float *__restrict__ lb, const float *__restrict__ dist,
                                const int32_t *__restrict__ assign, const float *__restrict__ shift,
                                unsigned long long *__restrict__ count)
{
    const float dmax = shift[K], dsec = shift[K + 1];
    const int imax = (int)shift[K + 2];
    unsigned int local = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        int a = assign[j];
        if ((unsigned)a >= (unsigned)K) a = 0;
        const float mv = (a == imax) ? dsec : dmax;
        local += (dist[j] + shift[a]) * (1.f + 2.0e-6f) < (lb[j] - mv);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, (unsigned long long)local);
}

__global__ void k_shift_top2(int64_t K, float *__restrict__ shift)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float m1 = 0.f, m2 = 0.f;
    int i1 = 0;
    for (int64_t k = 0; k < K; ++k) {
        const float v = shift[k];
        if (v > m1 || k == 0) { if (k) m2 = m1; m1 = v; i1 = (int)k; }
        else if (v > m2) m2 = v;
    }
    shift[K] = m1; shift[K + 1] = m2; shift[K + 2] = (float)i1;
}

__global__ void k_build_table_t(int64_t p, int64_t K, const double *__restrict__ ct /* [p+1][K] */, float *__restrict__ tt,
                                float *__restrict__ cmax)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = K * (p + 1);
    float m = 0.f;
    if (idx < total) {
        const int64_t k = idx / (p + 1), r = idx % (p + 1);
        const float v = r < p ? (float)ct[r * K + k] : 0.f;
        tt[idx] = v;
        m = fabsf(v);
        if (v != v) m = __int_as_float(0x7fc00000);
    }
    int mi = __float_as_int(m);
#pragma unroll
    for (int o = 16; o; o >>= 1) mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    if ((threadIdx.x & 31) == 0 && mi > 0) atomicMax(reinterpret_cast<int *>(cmax), mi);
}

}  // namespace

int skm_launch_center_shift(skm_ctx *ctx, int64_t p, int64_t K, const double *centers, double *centers_prev,
                            int has_gamma, double gamma, float *shift)
{
    k_center_shift<<<(unsigned)K, 256, 0, ctx->stream>>>(p, centers, centers_prev, has_gamma, gamma, shift);
    SKM_CHECK_LAUNCH(ctx);
    k_shift_top2<<<1, 32, 0, ctx->stream>>>(K, shift);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_bound_predict(skm_ctx *ctx, int64_t n, int64_t K, const float *lb, const float *dist, const int32_t *assign,
                             const float *shift, unsigned long long *count_dev)
{
    SKM_CUDA(cudaMemsetAsync(count_dev, 0, sizeof(unsigned long long), ctx->stream));
    if (n == 0) return SKM_OK;
    const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_bound_predict<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, (int)K, lb, dist, assign, shift, count_dev);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_build_table_t(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, float *table_t, float *cmax)
{
    SKM_CUDA(cudaMemsetAsync(cmax, 0, sizeof(float), ctx->stream));
    const int64_t total = K * (p + 1);
    k_build_table_t<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p, K, ct, table_t, cmax);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_assign_bounded(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const float *table_t, const float *cmax,
                              const float *shift, const int32_t *assign, float *lb, float *dist,
                              int32_t *flagged, int *nflag)
{
    SKM_CUDA(cudaMemsetAsync(nflag, 0, sizeof(int), ctx->stream));
    if (ds->n == 0) return SKM_OK;
    if (ds->sell_mode < 0 && !ds->sell_plain) { skm_set_error("assign_bounded: the SELL image has not been filled"); return SKM_ERR_STATE; }
    const double u = 5.9604644775390625e-08;
    const double m = (double)(ds->max_col_nnz > 0 ? ds->max_col_nnz : 1);
    BoundedParams P;
    P.sell = ds->sell; P.slice_ptr = ds->slice_ptr; P.nslices = ds->nslices; P.n = ds->n;
    P.uniform = ds->uniform_width ? 1 : 0; P.width2 = ds->sell_width2;
    P.p = (int)ds->p; P.K = (int)K;
    P.boff = (!ds->sell_plain && ds->sell_mode >= 1) ? (int)skm_dual_boff(ds->p) : 0;
    const size_t row_bytes = (size_t)(ds->p + 1) * sizeof(float);
    const size_t budget = (size_t)ctx->smem_optin - 2048;
    int64_t ksm = (int64_t)(budget / row_bytes);
    if (ksm > K) ksm = K;
    // When the table does not fit, the rows left out are gathered through L1 (the stream bypasses it, ld_stream):
    // leave them the L1 that a 196 KB carve-out keeps free instead of filling shared memory to the brim
    // (K = 64, p = 1024: 48 rows in shared memory 1.08 ms, 55 rows 1.28 ms, 52 rows 1.46 ms; profiles/r2_bounded_ksm.md)
    if (ksm < K) ksm = std::min<int64_t>(ksm, (int64_t)((193 * 1024) / row_bytes));
    if (ksm < 1) ksm = 1;
    {
        static const char *e = getenv("SKM_BOUNDED_KSM");          // tuning knob: centres kept in shared memory
        if (e && atoi(e) > 0 && atoi(e) < ksm) ksm = atoi(e);
    }
    // leave room for two CTAs per SM when the whole table is small
    const size_t smem = (size_t)ksm * row_bytes;
    P.ksm = (int)ksm;
    P.table_t = table_t;
    P.ga = (float)(1.01 * (m + 5.0) * u);
    P.gb_unit = (float)(2.02 * u * sqrt(m));
    P.ge_unit = (float)(2.1 * u * u * m);
    P.cmax = cmax; P.shift = shift; P.assign = assign; P.lb = lb; P.dist = dist; P.flagged = flagged; P.nflag = nflag;
    const bool padtail = ksm < K;                             // part of the table is read through L2
    auto kern = padtail ? k_assign_bounded<1024, true> : k_assign_bounded<1024, false>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 1024, smem));
    if (per_sm < 1) { skm_set_error("assign_bounded does not fit on an SM (smem %zu)", smem); return SKM_ERR_UNSUPPORTED; }
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    const int64_t need = (ds->nslices * 32 + 1023) / 1024;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 1024, smem, ctx->stream>>>(P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}
