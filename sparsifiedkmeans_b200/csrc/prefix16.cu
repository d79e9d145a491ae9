// prefix16.cu -- the prefix launch of the pruned assignment pass (api.cu, "partial-distance pruning") with a
// HALF-PRECISION centre table: all K <= 64 centres in ONE launch.
//
// The pruned plan needs, per column, a candidate winner and a rigorous lower bound on the reference's masked distance
// (private/SparseMatrixMinusCluster.c:169-182, sqrt of the sum over the column's stored entries) to every OTHER centre.
// Both come from the partial sums over the first `max_pairs` entry pairs.  With the fp32 table this takes one launch
// per 16 centres (the 64-centre table, 262 KB, does not fit in shared memory), and the per-launch fixed cost (table
// staging, per-column epilogue, re-reading the prefix) dominates a pass that touches 8 of 52 entries.  In fp16 the whole
// table is 144 B per row (148 KB at p = 1024): one launch, half the gathered bytes.  The table is only a FILTER here --
// the candidate's distance is evaluated exactly on all entries afterwards (k_assign_bounded) -- so its rounding only
// has to be bounded, not avoided:
//
//   t = fp16(s c'), s a power of two with s cmax in [2^13, 2^14)  =>  |t - s c'| <= delta = 2^-11 (1.001) s cmax + 2^-24
//   A_k   = fp32 sum over the prefix of (s x - t_k)^2             (s x exact; one rounding per subtraction and per FMA)
//   true  sqrt(sum (s x - t_k)^2) >= sqrt(A_k (1 - 1.01 (q + 5) u)),  q = entries in the prefix, u = 2^-24
//   s d_j(c_k) >= s sqrt(prefix sum of (x - c'_k)^2) >= sqrt(A_k (1 - ga)) - sqrt(q) delta        (triangle inequality)
//
// lb_j = that bound for the SECOND-smallest A (the candidate is the smallest), divided by s and rounded down.  The low
// six mantissa bits of every A carry the centre's index during the min/second-min network (they are cleared first, which
// only lowers the bound), so the epilogue is four integer min/max per centre with no selects.
#include "common.cuh"
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace {

struct Prefix16Params {
    const int4    *sell;
    const int64_t *slice_ptr;
    int64_t        nslices, n;
    int            uniform, width2;
    const unsigned char *table;   // this chunk's table: [(p + 1)][row_bytes], fp16 of s c'
    uint32_t       table_bytes;
    int            row_bytes;
    int            K, k0, first, last;
    int            max_pairs;
    int            boff, p;       // boff > 0: rows >= boff are second-copy rows of a dual-table image; rows > p are pad rows
    const float   *scale;         // [0] s, [1] 1/s, [2] delta (scaled units)
    float          ga, sq;        // 1.01 (q + 5) u;  sqrt(q) rounded up
    int32_t       *assign;
    uint2         *best2;         // running (smallest, second smallest) keys between the chunks of K > 64
    float         *lb;
};

__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA bulk copies behind an mbarrier (1-D cp.async.bulk, SASS UBLKCP); a copy that never lands traps instead of hanging
__device__ __forceinline__ void stage_table(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    const uint32_t bar_a = smem_addr(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        const uint32_t CH = 32768;
        const uint32_t dst = smem_addr(smem_dst);
        const char *src = (const char *)gsrc;
        for (uint32_t off = 0; off < bytes; off += CH) {
            const uint32_t sz = min(CH, bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(dst + off), "l"(src + off), "r"(sz), "r"(bar_a) : "memory");
        }
    }
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar_a) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();
    }
}

// one stored entry against KC centres: KC/8 LDS.128, KC mixed-precision subtractions (fp16 operand, fp32 result: no
// unpack instruction), KC/2 packed FMAs
template <int KC>
__device__ __forceinline__ void step16(unsigned long long (&acc2)[KC / 2], const unsigned char *tab, int row_bytes, int r, float xs)
{
    const uint4 *row = reinterpret_cast<const uint4 *>(tab + (size_t)r * row_bytes);
#pragma unroll
    for (int c = 0; c < KC / 8; ++c) {
        const uint4 q = row[c];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned short lo, hi;
            asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w[i]));
            float d0, d1;
            asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(d0) : "h"(lo), "f"(xs));
            asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(d1) : "h"(hi), "f"(xs));
            unsigned long long dd;
            asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d0), "f"(d1));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc2[4 * c + i]) : "l"(dd));
        }
    }
}

template <int KC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_prefix16(const Prefix16Params P)
{
    extern __shared__ __align__(128) unsigned char s_tab[];
    __shared__ uint64_t bar;
    stage_table(s_tab, P.table, P.table_bytes, &bar);

    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * THREADS) >> 5;
    const int rb = P.row_bytes, boff = P.boff, prow = P.p;
    const float s = P.scale[0];
    auto fold = [&](int r) { if (boff && r >= boff) r -= boff; return r > prow ? prow : r; };

    for (int64_t slice = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5; slice < P.nslices; slice += warps_total) {
        int64_t base;
        int w2;
        if (P.uniform) { base = slice * (int64_t)P.width2 * 32; w2 = P.width2; }
        else { base = P.slice_ptr[slice]; w2 = (int)((P.slice_ptr[slice + 1] - base) >> 5); }
        if (w2 > P.max_pairs) w2 = P.max_pairs;
        const int4 *src = P.sell + base + lane;

        unsigned long long acc2[KC / 2];
#pragma unroll
        for (int k = 0; k < KC / 2; ++k) acc2[k] = 0ULL;

        int t2 = 0;
        for (; t2 + 4 <= w2; t2 += 4) {
            const int4 q0 = __ldcs(src + (t2 + 0) * 32);
            const int4 q1 = __ldcs(src + (t2 + 1) * 32);
            const int4 q2 = __ldcs(src + (t2 + 2) * 32);
            const int4 q3 = __ldcs(src + (t2 + 3) * 32);
            step16<KC>(acc2, s_tab, rb, fold(q0.x), __int_as_float(q0.y) * s);
            step16<KC>(acc2, s_tab, rb, fold(q0.z), __int_as_float(q0.w) * s);
            step16<KC>(acc2, s_tab, rb, fold(q1.x), __int_as_float(q1.y) * s);
            step16<KC>(acc2, s_tab, rb, fold(q1.z), __int_as_float(q1.w) * s);
            step16<KC>(acc2, s_tab, rb, fold(q2.x), __int_as_float(q2.y) * s);
            step16<KC>(acc2, s_tab, rb, fold(q2.z), __int_as_float(q2.w) * s);
            step16<KC>(acc2, s_tab, rb, fold(q3.x), __int_as_float(q3.y) * s);
            step16<KC>(acc2, s_tab, rb, fold(q3.z), __int_as_float(q3.w) * s);
        }
        for (; t2 < w2; ++t2) {
            const int4 q = __ldcs(src + t2 * 32);
            step16<KC>(acc2, s_tab, rb, fold(q.x), __int_as_float(q.y) * s);
            step16<KC>(acc2, s_tab, rb, fold(q.z), __int_as_float(q.w) * s);
        }

        // ---- smallest and second-smallest partial sum of this chunk.  Keys: the sum's bit image with the low six bits
        // replaced by the centre's index in the chunk, compared as unsigned (non-negative floats order like their
        // images; +inf, NaN and anything with the sign bit set sort on top) ----
        uint32_t m1 = 0xffffffffu, m2 = 0xffffffffu;
        const int kv = P.K - P.k0;
        auto scan = [&](auto masked) {
#pragma unroll
            for (int k2 = 0; k2 < KC / 2; ++k2) {
                float a0, a1;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc2[k2]));
                uint32_t key0 = (__float_as_uint(a0) & ~63u) | (uint32_t)(2 * k2);
                uint32_t key1 = (__float_as_uint(a1) & ~63u) | (uint32_t)(2 * k2 + 1);
                if (decltype(masked)::value) {
                    if (2 * k2 >= kv) key0 = 0xffffffffu;
                    if (2 * k2 + 1 >= kv) key1 = 0xffffffffu;
                }
                uint32_t t = max(key0, m1); m1 = min(key0, m1); m2 = min(m2, t);
                t = max(key1, m1); m1 = min(key1, m1); m2 = min(m2, t);
            }
        };
        if (kv >= KC) scan(std::false_type{}); else scan(std::true_type{});      // uniform branch: a full chunk needs no masks
        const int64_t j = slice * SKM_SLICE + lane;
        if (j >= P.n) continue;
        int i1 = P.k0 + (int)(m1 & 63u);
        uint32_t v1 = m1 & ~63u, v2 = m2 & ~63u;
        if (!P.first) {
            const uint2 r = P.best2[j];
            if (v1 < r.x) v2 = min(r.x, v2);
            else { v2 = min(r.y, v1); v1 = r.x; i1 = P.assign[j]; }
        }
        P.assign[j] = i1;
        if (!P.last) { P.best2[j] = make_uint2(v1, v2); continue; }
        float lbv = 0.f;
        if (v2 < 0x7f800000u) {                                   // finite second-smallest sum
            const float b2 = __uint_as_float(v2);
            const float lbs = sqrtf(b2 * (1.f - P.ga)) * (1.f - 2.4e-7f) - P.sq * P.scale[2];
            if (lbs > 0.f) lbv = lbs * P.scale[1] * (1.f - 4.8e-7f);   // false for a NaN delta
        }
        P.lb[j] = lbv;
    }
}

// s, 1/s, delta from cmax = max |fp32(c')| (written by k_build_table_t), then the table itself
__device__ __forceinline__ float prefix16_scale(float cm)
{
    int e = 0;
    if (cm > 0.f && cm < __int_as_float(0x7f800000)) {
        int ex;
        frexpf(cm, &ex);                                          // cm = f 2^ex, f in [0.5, 1)
        e = 14 - ex;
        e = max(-100, min(100, e));
    }
    return __int_as_float((127 + e) << 23);
}

__global__ void k_build_table16(int64_t p, int64_t K, const double *__restrict__ ct /* [p + 1][K] */, int kc, int row_halves,
                                int nchunks, const float *__restrict__ cmax, __half *__restrict__ table, float *__restrict__ scale)
{
    const float cm = *cmax;
    const float s = prefix16_scale(cm);
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx == 0) {
        scale[0] = s;
        scale[1] = 1.f / s;                                       // exact: s is a power of two in [2^-100, 2^100]
        scale[2] = 4.888e-4f * (s * cm) + 6.0e-8f;                // 2^-11 * 1.001 * s * cmax + 2^-24, rounded up; NaN stays NaN
    }
    const int64_t per_chunk = (p + 1) * row_halves;
    if (idx >= per_chunk * nchunks) return;
    const int64_t c = idx / per_chunk, rem = idx % per_chunk;
    const int64_t r = rem / row_halves, kk = rem % row_halves, k = c * kc + kk;
    float v = 0.f;
    if (r < p && kk < kc && k < K) v = (float)(ct[r * K + k] * (double)s);
    table[idx] = __float2half_rn(v);
}

}  // namespace

// chunk of 32 or 64 centres per launch, the largest whose table fits in shared memory; false: use the fp32 prefix
bool skm_prefix16_plan(const skm_ctx *ctx, int64_t p, int64_t K, Prefix16Plan *pl)
{
    const size_t budget = (size_t)ctx->smem_optin - 1024;
    int kc = K <= 32 ? 32 : 64;
    for (;; kc = 32) {
        const size_t row = (size_t)kc * 2 + 16;                    // odd number of 16-byte chunks per row spreads the banks
        if ((size_t)(p + 1) * row <= budget) {
            pl->kc = kc;
            pl->row_bytes = (int)row;
            pl->nchunks = (int)((K + kc - 1) / kc);
            pl->smem = (((size_t)(p + 1) * row) + 127) & ~(size_t)127;
            return true;
        }
        if (kc == 32) return false;
    }
}

size_t skm_prefix16_table_bytes(int64_t p, const Prefix16Plan &pl)
{
    return (size_t)(p + 1) * pl.row_bytes * pl.nchunks + 16;
}

int skm_launch_build_table16(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, const Prefix16Plan &pl, const float *cmax,
                             void *table, float *scale)
{
    const int64_t total = (p + 1) * (int64_t)(pl.row_bytes / 2) * pl.nchunks;
    k_build_table16<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(p, K, ct, pl.kc, pl.row_bytes / 2, pl.nchunks, cmax,
                                                                             reinterpret_cast<__half *>(table), scale);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

template <int KC, int THREADS, int MINB>
static int launch_prefix16(skm_ctx *ctx, const Prefix16Params &P, size_t smem)
{
    auto kern = k_prefix16<KC, THREADS, MINB>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) { skm_set_error("prefix16<%d> does not fit on an SM (smem %zu)", KC, smem); return SKM_ERR_UNSUPPORTED; }
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    const int64_t need = (P.nslices * 32 + THREADS - 1) / THREADS;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, THREADS, smem, ctx->stream>>>(P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// candidate (assign) and lower bound on the distance to every other centre (lb) from the first max_pairs entry pairs
int skm_launch_prefix16(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const Prefix16Plan &pl, const void *table,
                        const float *scale, int32_t *assign, float *best2, float *lb, int max_pairs)
{
    if (ds->n == 0) return SKM_OK;
    if (ds->sell_mode < 0 && !ds->sell_plain) { skm_set_error("prefix16: the SELL image has not been filled"); return SKM_ERR_STATE; }
    const double u = 5.9604644775390625e-08;
    const double q = 2.0 * max_pairs;
    Prefix16Params P;
    P.sell = ds->sell; P.slice_ptr = ds->slice_ptr; P.nslices = ds->nslices; P.n = ds->n;
    P.uniform = ds->uniform_width ? 1 : 0; P.width2 = ds->sell_width2;
    P.row_bytes = pl.row_bytes;
    P.table_bytes = (uint32_t)((((size_t)(ds->p + 1) * pl.row_bytes) + 15) & ~(size_t)15);
    P.K = (int)K;
    P.max_pairs = max_pairs;
    P.boff = (!ds->sell_plain && ds->sell_mode >= 1) ? (int)skm_dual_boff(ds->p) : 0;
    P.p = (int)ds->p;
    P.scale = scale;
    P.ga = (float)(1.01 * (q + 5.0) * u);
    P.sq = (float)(sqrt(q) * (1.0 + 1e-6));
    P.assign = assign; P.best2 = reinterpret_cast<uint2 *>(best2); P.lb = lb;
    for (int c = 0; c < pl.nchunks; ++c) {
        P.table = (const unsigned char *)table + (size_t)c * (ds->p + 1) * pl.row_bytes;
        P.k0 = c * pl.kc;
        P.first = (c == 0);
        P.last = (c == pl.nchunks - 1);
        const bool two = 2 * (pl.smem + 1024) <= (size_t)ctx->smem_optin;
        int rc;
        if (pl.kc == 32) rc = two ? launch_prefix16<32, 384, 2>(ctx, P, pl.smem) : launch_prefix16<32, 768, 1>(ctx, P, pl.smem);
        else rc = launch_prefix16<64, 512, 1>(ctx, P, pl.smem);
        if (rc != SKM_OK) return rc;
    }
    return SKM_OK;
}
