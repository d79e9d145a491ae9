// assign_fast.cu -- K1, the hot kernel: fused masked squared distance + argmin over the
// SELL-32 image of the sparsified matrix (replaces private/SparseMatrixMinusCluster.c:169-182
// followed by min, private/findClusterAssignments.m:168-171, without the K x n temporary).
//
// Mapping: one thread per column, one warp per 32-column slice, persistent warps striding over
// slices.  Lane l streams its column's (row, value) pairs with one coalesced 128-bit load per
// two entries (the slice is stored interleaved, so a warp reads 512 contiguous bytes), gathers
// the centroid row c'[row, 0:KC) from a padded fp32 table staged once per CTA into shared
// memory with a TMA bulk copy, and keeps KC running sums in registers.  No tensor cores: this
// is a sparse gather.  Algorithmic HBM bytes per point: 8 per stored entry + 4 (assignment)
// + 4 (distance).
//
// Exactness: sums are fp32 here, the reference's are fp64.  Each column's winner is certified
// with a rigorous rounding bound (see guard below); columns that cannot be certified are
// appended to `flagged` and re-evaluated in fp64 in the reference's order by exact.cu, so the
// assignments equal the reference's bit for bit.
#include "common.cuh"
#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace {

struct FastParams {
    const int4    *sell;
    const int64_t *slice_ptr;
    int64_t        nslices, n;
    int            uniform, width2;
    const float   *table;        // this chunk's table: [(p+1)][ks]
    uint32_t       table_bytes;
    int            ks;
    int            k0, kvalid, ktotal;   // first centre of the chunk, centres in it, K
    int            first, last;
    float          ga;           // guard: relative term  1.01*(m+5)*u
    float          gb_unit;      // guard: 2.02*u*sqrt(m)      (times cmax)
    float          ge_unit;      // guard: 2.1*u*u*m           (times cmax^2)
    const float   *cmax;
    const int     *m_dev;        // optional: max entries per column on the device (streamed path)
    int32_t       *assign;
    float         *dist;
    float2        *best2;
    int32_t       *flagged;
    int           *nflag;
    float         *lb;           // optional: lower bound on the distance to every centre but the winner (bounded.cu)
    int            max_pairs;    // > 0: only the first max_pairs entry pairs of every column (partial-distance pass)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Stage `bytes` (multiple of 16) from global to shared memory with TMA bulk copies tracked by
// an mbarrier; all threads of the CTA return once the data has landed.
__device__ __forceinline__ void tma_stage(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    const uint32_t bar_a = smem_u32(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        const uint32_t CH = 32768;
        uint32_t dst = smem_u32(smem_dst);
        const char *src = (const char *)gsrc;
        for (uint32_t off = 0; off < bytes; off += CH) {
            uint32_t sz = min(CH, bytes - off);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                ::"r"(dst + off), "l"(src + off), "r"(sz), "r"(bar_a) : "memory");
        }
    }
    uint32_t done = 0, spins = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar_a) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();       // a copy that never lands is a bug, not a hang
    }
}

// Per-column epilogue shared by the fast kernels: best / second best of this chunk, merge with
// earlier chunks, exactness guard, results.
template <int KC>
__device__ __forceinline__ void finish_column(const FastParams &P, const float (&acc)[KC], int64_t slice, int lane)
{
        // ---- per-column epilogue: best / second best of this chunk ----
    const float INF = __int_as_float(0x7f800000);
    const float QNAN = __int_as_float(0x7fc00000);
    float b1 = INF, b2 = INF;
    int i1 = P.k0;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        if (k < P.kvalid) {
            const float v = acc[k];
            if (!(v < INF)) bad = true;                    // NaN or overflow: cannot certify
            if (v < b1) { b2 = b1; b1 = v; i1 = P.k0 + k; }
            else if (v < b2) b2 = v;
        }
    }
    if (bad) b2 = QNAN;

    const int64_t j = slice * SKM_SLICE + lane;
    if (j >= P.n) return;
    if (!P.first) {                                        // merge with earlier chunks
        const float2 r = P.best2[j];
        const int ri = P.assign[j];
        const bool nan2 = (r.y != r.y) || (b2 != b2);
        if (b1 < r.x) { b2 = fminf(r.x, b2); }
        else { b2 = fminf(r.y, b1); b1 = r.x; i1 = ri; }
        if (nan2) b2 = QNAN;
    }
    if (!P.last) {
        P.best2[j] = make_float2(b1, b2);
        P.assign[j] = i1;
        return;
    }
    // ---- guard: |fp32 sum - exact sum| <= E(s) = ga*s + gb*sqrt(s) + ge   ----
    const float cm = *P.cmax;
    float ga = P.ga, gbu = P.gb_unit, geu = P.ge_unit;
    if (P.m_dev) {                                          // same constants, rounded up in fp32
        const float U = 5.9604644775390625e-08f, mm = (float)max(*P.m_dev, 1);
        ga = 1.02f * (mm + 5.f) * U; gbu = 2.04f * U * sqrtf(mm); geu = 2.2f * U * U * mm;
    }
    const float gb = gbu * cm, ge = geu * cm * cm + 1e-37f;
    bool certified;
    if (P.ktotal == 1) certified = (b1 < INF);
    else {
        const float E = ga * (b1 + b2) + gb * (sqrtf(b1) + sqrtf(b2)) + 2.f * ge;
        certified = (b2 - b1) > E;                         // false for NaN / inf
    }
    P.assign[j] = i1;
    P.dist[j] = sqrtf(b1);
    if (P.lb) {
        // sqrt(second-best sum minus its error bound), rounded down; uncertified columns get theirs from
        // the fp64 re-evaluation (NaN / single centre: 0 resp. +inf keep the bounded test conservative)
        float lbv = 0.f;
        if (P.ktotal == 1) lbv = INF;
        else if (b2 == b2) {
            const float e2 = ga * b2 + gb * sqrtf(b2) + ge;
            lbv = sqrtf(fmaxf(b2 - e2, 0.f)) * (1.f - 4.76837158203125e-07f);
        }
        P.lb[j] = lbv;
    }
    if (!certified) {
        const int slot = atomicAdd(P.nflag, 1);
        P.flagged[slot] = (int32_t)j;
    }
}

template <int KC>
__device__ __forceinline__ void step(float (&acc)[KC], const float *tab, int ks, int r, float x)
{
    const float4 *row = reinterpret_cast<const float4 *>(tab + (size_t)r * ks);
#pragma unroll
    for (int c = 0; c < KC / 4; ++c) {
        const float4 v = row[c];
        float d;
        d = x - v.x; acc[4 * c + 0] = fmaf(d, d, acc[4 * c + 0]);
        d = x - v.y; acc[4 * c + 1] = fmaf(d, d, acc[4 * c + 1]);
        d = x - v.z; acc[4 * c + 2] = fmaf(d, d, acc[4 * c + 2]);
        d = x - v.w; acc[4 * c + 3] = fmaf(d, d, acc[4 * c + 3]);
    }
}

// GTAB = true: the centroid table is too large for shared memory (e.g. p2 = 32768) and is gathered
// from global memory instead (it stays L2-resident; the kernel is then L2-bound, not HBM-bound).
template <int KC, int THREADS, int MINB, bool GTAB = false>
__global__ void __launch_bounds__(THREADS, MINB) k_assign_fast(const FastParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    const float *tab;
    if (GTAB) tab = P.table;
    else {
        float *stab = reinterpret_cast<float *>(smem_raw);
        tma_stage(stab, P.table, P.table_bytes, &bar);
        tab = stab;
    }

    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * THREADS) >> 5;
    int64_t slice = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;
    const int ks = P.ks;

    for (; slice < P.nslices; slice += warps_total) {
        int64_t base;
        int w2;
        if (P.uniform) { base = slice * (int64_t)P.width2 * 32; w2 = P.width2; }
        else { base = P.slice_ptr[slice]; w2 = (int)((P.slice_ptr[slice + 1] - base) >> 5); }
        if (P.max_pairs > 0 && w2 > P.max_pairs) w2 = P.max_pairs;     // partial pass: a prefix of every column's entries
        const int4 *src = P.sell + base + lane;

        float acc[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[k] = 0.f;

        int t2 = 0;
        for (; t2 + 4 <= w2; t2 += 4) {
            const int4 q0 = __ldcs(src + (t2 + 0) * 32);
            const int4 q1 = __ldcs(src + (t2 + 1) * 32);
            const int4 q2 = __ldcs(src + (t2 + 2) * 32);
            const int4 q3 = __ldcs(src + (t2 + 3) * 32);
            step<KC>(acc, tab, ks, q0.x, __int_as_float(q0.y));
            step<KC>(acc, tab, ks, q0.z, __int_as_float(q0.w));
            step<KC>(acc, tab, ks, q1.x, __int_as_float(q1.y));
            step<KC>(acc, tab, ks, q1.z, __int_as_float(q1.w));
            step<KC>(acc, tab, ks, q2.x, __int_as_float(q2.y));
            step<KC>(acc, tab, ks, q2.z, __int_as_float(q2.w));
            step<KC>(acc, tab, ks, q3.x, __int_as_float(q3.y));
            step<KC>(acc, tab, ks, q3.z, __int_as_float(q3.w));
        }
        for (; t2 < w2; ++t2) {
            const int4 q = __ldcs(src + t2 * 32);
            step<KC>(acc, tab, ks, q.x, __int_as_float(q.y));
            step<KC>(acc, tab, ks, q.z, __int_as_float(q.w));
        }

        finish_column<KC>(P, acc, slice, lane);
    }
}

// LDS.64 variant for K chunks with an odd number of 8-byte slots per table row (KC = 2, 6, 10, 14):
// the row is exactly KC floats (K = 10: 40 bytes instead of the 48 the 16-byte variant gathers),
// one LDS.64 per two centres.  A 64-bit shared load is served per half-warp; its 16 slots are
// conflict-free iff the 16 rows differ mod 16 (slot = row * KC/2 + c, KC/2 odd).  The SELL image
// built for this kernel (convert.cu, layout mode 1) guarantees that: the table is staged TWICE
// (copy B starts at row `boff`, boff = 1 mod 16, so an entry can be served from bank class
// row mod 16 or row + 1 mod 16), the per-class loads of every half-warp are balanced over the two
// copies and the entries are then scheduled by an exact bipartite edge colouring.  The row field
// of the image already holds the row of the copy to read; pad entries point at one of 16 zero rows.
template <int KC>
__device__ __forceinline__ void step64(float (&acc)[KC], const float *tab, int r, float x)
{
    const float2 *row = reinterpret_cast<const float2 *>(tab + (size_t)r * KC);
#pragma unroll
    for (int c = 0; c < KC / 2; ++c) {
        const float2 v = row[c];
        float d;
        d = x - v.x; acc[2 * c + 0] = fmaf(d, d, acc[2 * c + 0]);
        d = x - v.y; acc[2 * c + 1] = fmaf(d, d, acc[2 * c + 1]);
    }
}

template <int KC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_assign_fast64(const FastParams P)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    float *stab = reinterpret_cast<float *>(smem_raw);
    tma_stage(stab, P.table, P.table_bytes, &bar);
    const float *tab = stab;

    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * THREADS) >> 5;
    int64_t slice = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5;

    for (; slice < P.nslices; slice += warps_total) {
        int64_t base;
        int w2;
        if (P.uniform) { base = slice * (int64_t)P.width2 * 32; w2 = P.width2; }
        else { base = P.slice_ptr[slice]; w2 = (int)((P.slice_ptr[slice + 1] - base) >> 5); }
        if (P.max_pairs > 0 && w2 > P.max_pairs) w2 = P.max_pairs;     // partial pass: a prefix of every column's entries
        const int4 *src = P.sell + base + lane;

        float acc[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[k] = 0.f;

        int t2 = 0;
        for (; t2 + 4 <= w2; t2 += 4) {
            const int4 q0 = __ldcs(src + (t2 + 0) * 32);
            const int4 q1 = __ldcs(src + (t2 + 1) * 32);
            const int4 q2 = __ldcs(src + (t2 + 2) * 32);
            const int4 q3 = __ldcs(src + (t2 + 3) * 32);
            step64<KC>(acc, tab, q0.x, __int_as_float(q0.y));
            step64<KC>(acc, tab, q0.z, __int_as_float(q0.w));
            step64<KC>(acc, tab, q1.x, __int_as_float(q1.y));
            step64<KC>(acc, tab, q1.z, __int_as_float(q1.w));
            step64<KC>(acc, tab, q2.x, __int_as_float(q2.y));
            step64<KC>(acc, tab, q2.z, __int_as_float(q2.w));
            step64<KC>(acc, tab, q3.x, __int_as_float(q3.y));
            step64<KC>(acc, tab, q3.z, __int_as_float(q3.w));
        }
        for (; t2 < w2; ++t2) {
            const int4 q = __ldcs(src + t2 * 32);
            step64<KC>(acc, tab, q.x, __int_as_float(q.y));
            step64<KC>(acc, tab, q.z, __int_as_float(q.w));
        }
        finish_column<KC>(P, acc, slice, lane);
    }
}

// table rows: [0,p) the centres, [p,zrows_end) zeros, then (dual tables) a second copy at row boff
__global__ void k_build_table(int64_t p, int64_t K, const double *__restrict__ ct, int kc, int ks,
                              int nchunks, int64_t rows, int64_t boff, float *__restrict__ table,
                              float *__restrict__ cmax)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t per_chunk = rows * ks;
    int64_t total = per_chunk * nchunks;
    float m = 0.f;
    if (idx < total) {
        int64_t c = idx / per_chunk, rem = idx % per_chunk;
        int64_t r = rem / ks, kk = rem % ks;
        int64_t k = c * kc + kk;
        if (boff > 0 && r >= boff) r -= boff;              // second copy
        float v = 0.f;
        if (r < p && kk < kc && k < K) v = (float)ct[r * K + k];
        table[idx] = v;
        m = fabsf(v);
        if (v != v) m = __int_as_float(0x7fc00000);
    }
    // block max via warp shuffles on the int image (non-negative floats order like ints; NaN on top)
    int mi = __float_as_int(m);
#pragma unroll
    for (int o = 16; o; o >>= 1) mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    if ((threadIdx.x & 31) == 0 && mi > 0) atomicMax(reinterpret_cast<int *>(cmax), mi);
}

template <int KC, int THREADS, int MINB, bool GTAB = false>
int launch_fast(skm_ctx *ctx, const FastParams &P, size_t smem)
{
    auto kern = k_assign_fast<KC, THREADS, MINB, GTAB>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) {
        skm_set_error("assign_fast<%d>: kernel does not fit on an SM (smem %zu)", KC, smem);
        return SKM_ERR_UNSUPPORTED;
    }
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    int64_t need = (P.nslices * 32 + THREADS - 1) / THREADS;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, THREADS, smem, ctx->stream>>>(P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

template <int KC, int THREADS, int MINB>
int launch_fast64(skm_ctx *ctx, const FastParams &P, size_t smem)
{
    auto kern = k_assign_fast64<KC, THREADS, MINB>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
    if (per_sm < 1) {
        skm_set_error("assign_fast64<%d>: kernel does not fit on an SM (smem %zu)", KC, smem);
        return SKM_ERR_UNSUPPORTED;
    }
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    int64_t need = (P.nslices * 32 + THREADS - 1) / THREADS;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, THREADS, smem, ctx->stream>>>(P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

}  // namespace

static const int kKcOptions[] = {4, 8, 12, 16, 24, 32, 48, 64};

int skm_fast_stride(int kc) { return 4 * ((kc / 4) | 1); }

static int stride_for(int kc)
{
    int chunks = kc / 4;
    return 4 * (chunks | 1);          // odd number of 16-byte chunks per row spreads the banks
}

int64_t skm_dual_boff(int64_t p)
{
    // first row of the second table copy: leaves 16 zero rows [p, p+16) for pad entries and is
    // 1 mod 16, so copy B of a row sits one bank class above copy A
    const int64_t b = p + 16;
    return b + (((1 - b) % 16) + 16) % 16;
}

bool skm_fast_plan(const skm_ctx *ctx, int64_t p, int64_t K, FastPlan *plan, int64_t max_col_nnz)
{
    const size_t budget = (size_t)ctx->smem_optin - 1024;
    plan->mode64 = false;
    plan->boff = 0;
    plan->rows = p + 1;
    // LDS.64 kernel on a dual table: K <= 14 with an odd number of 8-byte slots per row, columns
    // short enough for the byte-sized scheduler state (max_col_nnz < 0: caller cannot use it)
    {
        const int kc = (int)((K + 1) & ~(int64_t)1);
        const int64_t boff = skm_dual_boff(p);
        const size_t bytes = (size_t)(boff + p) * kc * sizeof(float);
        static const bool off = getenv("SKM_NO_FAST64") != nullptr;
        if (!off && K <= 14 && ((kc / 2) & 1) && max_col_nnz >= 0 && max_col_nnz <= 254 && p < (1 << 20) &&
            bytes <= budget / 2) {
            plan->mode64 = true;
            plan->dual8 = false;
            plan->layout = 1;
            plan->boff = (int)boff;
            plan->rows = boff + p;
            plan->kc = kc;
            plan->ks = kc;
            plan->nchunks = 1;
            plan->smem = (bytes + 127 + 16) & ~(size_t)127;
            plan->threads = 512;
            plan->global_table = false;
            return true;
        }
    }
    plan->dual8 = false;
    plan->layout = 0;
    int best = -1;
    // largest chunk whose table fits; then the smallest chunk that needs no more launches
    // (K=64 with room for 48 centres takes two launches of 32, not 48 + 16)
    for (int kc : kKcOptions) {
        size_t bytes = (size_t)(p + 1) * stride_for(kc) * sizeof(float);
        if (bytes > budget) break;
        best = kc;
        if (kc >= K) break;
    }
    plan->global_table = false;
    if (best < 0) {
        // no chunk fits in shared memory: gather from global memory (L2) with a 16-centre chunk
        plan->global_table = true;
        plan->kc = K <= 4 ? 4 : (K <= 8 ? 8 : 16);
        plan->ks = stride_for(plan->kc);
        plan->nchunks = (int)((K + plan->kc - 1) / plan->kc);
        plan->smem = 0;
        plan->threads = 256;
        return true;
    }
    const int launches = (int)((K + best - 1) / best);
    for (int kc : kKcOptions) {
        if ((K + kc - 1) / kc <= launches) { best = kc; break; }
    }
    // Dual-table alternative (conflict-free schedule, SELL layout mode 2): the table is staged twice, so
    // chunks are smaller and there may be more launches.  Cost model in "centre units" per launch
    // (measured, K=64 p=1024 m=51: 0.096 ms per centre with the greedy order's ~23% conflicts, 0.084
    // conflict-free; a launch cannot beat its HBM pass, ~0.85 ms = 10.9 units).
    {
        static const bool off = getenv("SKM_NO_DUAL8") != nullptr;
        static const bool force = getenv("SKM_FORCE_DUAL8") != nullptr;
        const int64_t boff = skm_dual_boff(p);
        auto cost = [&](int kc, double per_centre) {
            double t = 0;
            for (int64_t k0 = 0; k0 < K; k0 += kc) t += std::max(per_centre * (double)kc, 10.9 * 0.078);
            return t;
        };
        // cheapest dual-table chunk among those whose doubled table fits (padding slots cost like centres)
        int bd = -1;
        double t_dual = 1e300;
        for (int kc : kKcOptions) {
            size_t bytes = (size_t)(boff + p) * stride_for(kc) * sizeof(float);
            if (bytes > budget) break;
            const double t = cost(kc, 0.084);
            if (t < t_dual - 1e-9) { t_dual = t; bd = kc; }
            if (kc >= K) break;
        }
        if (!off && bd > 0 && max_col_nnz >= 0 && max_col_nnz <= 254 && p < (1 << 20)) {
            const double t_single = cost(best, 0.096);
            if (force || t_dual < 0.97 * t_single) {
                plan->dual8 = true;
                plan->layout = 2;
                plan->boff = (int)boff;
                plan->rows = boff + p;
                plan->kc = bd;
                plan->ks = stride_for(bd);
                plan->nchunks = (int)((K + bd - 1) / bd);
                plan->smem = (((size_t)plan->rows * plan->ks * sizeof(float)) + 127) & ~(size_t)127;
                plan->threads = 256;
                return true;
            }
        }
    }
    plan->kc = best;
    plan->ks = stride_for(best);
    plan->nchunks = (int)((K + best - 1) / best);
    plan->smem = (((size_t)(p + 1) * plan->ks * sizeof(float)) + 127) & ~(size_t)127;
    plan->threads = 256;
    return true;
}

size_t skm_fast_table_floats(int64_t p, const FastPlan &pl)
{
    (void)p;
    return (size_t)pl.rows * pl.ks * pl.nchunks + 4;        // the staged size is rounded up to 16 bytes
}

int skm_launch_build_table(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, const FastPlan &pl,
                           float *table, float *cmax)
{
    SKM_CUDA(cudaMemsetAsync(cmax, 0, sizeof(float), ctx->stream));
    int64_t total = pl.rows * pl.ks * (int64_t)pl.nchunks;
    int64_t blocks = (total + 255) / 256;
    k_build_table<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, K, ct, pl.kc, pl.ks, pl.nchunks, pl.rows, pl.boff, table, cmax);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_assign_fast(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const FastPlan &pl,
                           const float *table, const float *cmax, int32_t *assign, float *dist,
                           float *best2, int32_t *flagged, int *nflag, const int *m_dev, float *lb, int max_pairs)
{
    SKM_CUDA(cudaMemsetAsync(nflag, 0, sizeof(int), ctx->stream));
    if (ds->n == 0) return SKM_OK;
    const double u = 5.9604644775390625e-08;   // 2^-24
    const double m = (double)(ds->max_col_nnz > 0 ? ds->max_col_nnz : 1);
    FastParams P;
    P.sell = ds->sell;
    P.slice_ptr = ds->slice_ptr;
    P.nslices = ds->nslices;
    P.n = ds->n;
    P.uniform = ds->uniform_width ? 1 : 0;
    P.width2 = ds->sell_width2;
    P.ks = pl.ks;
    P.table_bytes = (uint32_t)((((size_t)pl.rows * pl.ks * sizeof(float)) + 15) & ~(size_t)15);   // bulk copies move multiples of 16 bytes
    P.ktotal = (int)K;
    P.ga = (float)(1.01 * (m + 5.0) * u);
    P.gb_unit = (float)(2.02 * u * sqrt(m));
    P.ge_unit = (float)(2.1 * u * u * m);
    P.cmax = cmax;
    P.m_dev = m_dev;
    P.assign = assign;
    P.dist = dist;
    P.best2 = reinterpret_cast<float2 *>(best2);
    P.flagged = flagged;
    P.nflag = nflag;
    P.lb = lb;
    P.max_pairs = max_pairs;
    for (int c = 0; c < pl.nchunks; ++c) {
        P.table = table + (size_t)c * pl.rows * pl.ks;
        P.k0 = c * pl.kc;
        P.kvalid = (int)((K - P.k0) < pl.kc ? (K - P.k0) : pl.kc);
        P.first = (c == 0);
        P.last = (c == pl.nchunks - 1);
        int rc;
        if (!ds->sell_plain && ds->sell_mode != pl.layout) {
            skm_set_error("assign_fast: the SELL image is in layout %d, the plan reads layout %d", ds->sell_mode, pl.layout);
            return SKM_ERR_STATE;
        }
        if (pl.mode64) {
            const bool wide = getenv("SKM_FAST64_WIDE") != nullptr;      // tuning knob: one 1024-thread CTA per SM
            switch (pl.kc) {
                case 2:  rc = launch_fast64<2, 512, 2>(ctx, P, pl.smem); break;
                case 6:  rc = launch_fast64<6, 512, 2>(ctx, P, pl.smem); break;
                case 10: rc = wide ? launch_fast64<10, 1024, 1>(ctx, P, pl.smem) : launch_fast64<10, 512, 2>(ctx, P, pl.smem); break;
                case 14: rc = launch_fast64<14, 512, 2>(ctx, P, pl.smem); break;
                default: skm_set_error("assign_fast64: unsupported chunk %d", pl.kc); return SKM_ERR_UNSUPPORTED;
            }
            if (rc != SKM_OK) return rc;
            continue;
        }
        if (pl.global_table) {
            switch (pl.kc) {
                case 4:  rc = launch_fast<4, 256, 4, true>(ctx, P, 0); break;
                case 8:  rc = launch_fast<8, 256, 4, true>(ctx, P, 0); break;
                default: rc = launch_fast<16, 256, 4, true>(ctx, P, 0); break;
            }
            if (rc != SKM_OK) return rc;
            continue;
        }
        // CTA shape by table size: one wide CTA per SM when the table leaves room for nothing else
        const bool one = pl.smem > 110 * 1024, two = !one && pl.smem > 54 * 1024;
        switch (pl.kc) {
            case 4:  rc = one ? launch_fast<4, 1024, 1>(ctx, P, pl.smem) : two ? launch_fast<4, 512, 2>(ctx, P, pl.smem) : launch_fast<4, 256, 4>(ctx, P, pl.smem); break;
            case 8:  rc = one ? launch_fast<8, 1024, 1>(ctx, P, pl.smem) : two ? launch_fast<8, 512, 2>(ctx, P, pl.smem) : launch_fast<8, 256, 4>(ctx, P, pl.smem); break;
            case 12: rc = one ? launch_fast<12, 1024, 1>(ctx, P, pl.smem) : two ? launch_fast<12, 512, 2>(ctx, P, pl.smem) : launch_fast<12, 256, 4>(ctx, P, pl.smem); break;
            case 16: rc = one ? launch_fast<16, 1024, 1>(ctx, P, pl.smem) : two ? launch_fast<16, 512, 2>(ctx, P, pl.smem) : launch_fast<16, 256, 4>(ctx, P, pl.smem); break;
            case 24: rc = one ? launch_fast<24, 512, 1>(ctx, P, pl.smem) : launch_fast<24, 256, 3>(ctx, P, pl.smem); break;
            case 32: rc = one ? launch_fast<32, 512, 1>(ctx, P, pl.smem) : launch_fast<32, 256, 2>(ctx, P, pl.smem); break;
            case 48: rc = one ? launch_fast<48, 512, 1>(ctx, P, pl.smem) : launch_fast<48, 256, 2>(ctx, P, pl.smem); break;
            case 64: rc = one ? launch_fast<64, 512, 1>(ctx, P, pl.smem) : launch_fast<64, 256, 2>(ctx, P, pl.smem); break;
            default: skm_set_error("assign_fast: unsupported chunk %d", pl.kc); return SKM_ERR_UNSUPPORTED;
        }
        if (rc != SKM_OK) return rc;
    }
    return SKM_OK;
}
