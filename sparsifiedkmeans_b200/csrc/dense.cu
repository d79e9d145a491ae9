// dense.cu -- the second pass over the ORIGINAL dense data:
//   kmeans_sparsified.m:542-560 (in core) and private/recalculateAssignmentLargeFile.m:85-113
//   (out of core, the same two computations chunk by chunk):
//     pass 1   centers_twoPass(:,k) = mean( XFull(:, bestAssignments == k), 2 )
//     pass 2   [assignments_twoPass, distances_twoPass] = findClusters( full(XFull), bestCenters )
//              = the dense branch of private/findClusterAssignments.m:124-171 (plain Euclidean
//              nearest centre; gamma is not used there).
//
// Both passes read each column chunk of X once while it is on the device.  Points are columns
// of the p x n column-major matrix, so a point is p contiguous values.
//
// K-dense (k_dense_assign): one warp per G columns, lanes over rows; the centres are a fp32
// row-major table in shared memory (same padded layout as K1), each table row is loaded once per
// lane and used for G columns, G*KC running sums live in registers and are folded across the warp
// with shuffles.  sum (x - c)^2 is evaluated directly (no |x|^2 - 2x'c + |c|^2 cancellation), so
// the same rounding guard as K1 certifies the winner; uncertified columns go to k_dense_exact
// (fp64, rows in order, separately rounded multiply/add, first-index ties, NaN skipped).
// Bound: 4 bytes per matrix element from HBM against 2*K fp32 instructions per element: HBM-bound
// up to K ~ 8, fp32-issue-bound above.  No tensor cores: a tf32 product cannot certify a winner
// (2^-11 relative) and the fp32 CUDA-core rate already sits above the PCIe rate the data arrives at.
//
// K-sums (k_dense_sums): one thread per row of a 128-row tile, looping over the columns of its
// column range: the read of a column is a coalesced 512-byte segment, the assignment is uniform
// across the CTA, every thread owns its row of the [cluster][128] fp64 bins in shared memory (no
// conflicts, no atomics); bins leave the SM as one fp64 atomic per (row, cluster, CTA).  HBM-bound.
#include "common.cuh"
#include <algorithm>
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <vector>

namespace {

struct DenseParams {
    const float  *x;            // [n][p] fp32 (column-major p x n)
    int64_t       p, n;
    const float  *table;        // nchunks x [(p+1)][ks]
    int           ks, kc_unused, nchunks, K;
    int64_t       chunk_floats; // (p+1)*ks
    int           table_in_smem;
    float         ga, gb_unit, ge_unit;
    const float  *cmax;
    int32_t      *assign;       // 0-based
    float        *dist;
    int32_t      *flagged;
    int          *nflag;
};

template <int KC, int G>
__global__ void __launch_bounds__(256, 2) k_dense_assign(const DenseParams P)
{
    extern __shared__ __align__(16) float s_tab[];
    const float *tab = P.table;
    if (P.table_in_smem) {
        const int64_t total = P.chunk_floats * P.nchunks;
        for (int64_t i = threadIdx.x * 4; i < total; i += blockDim.x * 4)
            *reinterpret_cast<float4 *>(s_tab + i) = *reinterpret_cast<const float4 *>(P.table + i);
        __syncthreads();
        tab = s_tab;
    }
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t ngroups = (P.n + G - 1) / G;
    const float INF = __int_as_float(0x7f800000), QNAN = __int_as_float(0x7fc00000);
    const int64_t p = P.p;

    for (int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ngroups; g += nwarps) {
        const int64_t j0 = g * G;
        const float *xc[G];
#pragma unroll
        for (int q = 0; q < G; ++q) xc[q] = P.x + min(j0 + q, P.n - 1) * p;     // tail group re-reads the last column
        float b1[G], b2[G], xm[G];
        int i1[G];
        bool anybad[G];
#pragma unroll
        for (int q = 0; q < G; ++q) { b1[q] = INF; b2[q] = INF; i1[q] = 0; xm[q] = 0.f; anybad[q] = false; }

        for (int c = 0; c < P.nchunks; ++c) {
            const float *tc = tab + (int64_t)c * P.chunk_floats;
            float acc[G][KC];
#pragma unroll
            for (int q = 0; q < G; ++q)
#pragma unroll
                for (int k = 0; k < KC; ++k) acc[q][k] = 0.f;
            // U row-steps per trip: all U*G loads are issued before the arithmetic (bytes in flight)
            constexpr int U = (KC * G <= 48) ? 4 : 2;
            auto body = [&](const float (&xv)[G], int64_t r) {
                if (c == 0) {
#pragma unroll
                    for (int q = 0; q < G; ++q) xm[q] = fmaxf(xm[q], fabsf(xv[q]));
                }
                const float4 *row = reinterpret_cast<const float4 *>(tc + r * P.ks);
#pragma unroll
                for (int k4 = 0; k4 < KC / 4; ++k4) {
                    const float4 v = row[k4];
#pragma unroll
                    for (int q = 0; q < G; ++q) {
                        float d;
                        d = xv[q] - v.x; acc[q][4 * k4 + 0] = fmaf(d, d, acc[q][4 * k4 + 0]);
                        d = xv[q] - v.y; acc[q][4 * k4 + 1] = fmaf(d, d, acc[q][4 * k4 + 1]);
                        d = xv[q] - v.z; acc[q][4 * k4 + 2] = fmaf(d, d, acc[q][4 * k4 + 2]);
                        d = xv[q] - v.w; acc[q][4 * k4 + 3] = fmaf(d, d, acc[q][4 * k4 + 3]);
                    }
                }
            };
            int64_t r = lane;
            for (; r + 32 * (U - 1) < p; r += 32 * U) {
                float xv[U][G];
#pragma unroll
                for (int uu = 0; uu < U; ++uu)
#pragma unroll
                    for (int q = 0; q < G; ++q) xv[uu][q] = __ldcs(xc[q] + r + 32 * uu);
#pragma unroll
                for (int uu = 0; uu < U; ++uu) body(xv[uu], r + 32 * uu);
            }
            for (; r < p; r += 32) {
                float xv[G];
#pragma unroll
                for (int q = 0; q < G; ++q) xv[q] = __ldcs(xc[q] + r);
                body(xv, r);
            }
            // fold across the warp; every lane ends with every sum
#pragma unroll
            for (int q = 0; q < G; ++q)
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    float v = acc[q][k];
#pragma unroll
                    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                    acc[q][k] = v;
                }
            const int k0 = c * KC;
#pragma unroll
            for (int q = 0; q < G; ++q) {
                bool bad = false;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    if (k0 + k < P.K) {
                        const float v = acc[q][k];
                        if (!(v < INF)) bad = true;
                        if (v < b1[q]) { b2[q] = b1[q]; b1[q] = v; i1[q] = k0 + k; }
                        else if (v < b2[q]) b2[q] = v;
                    }
                }
                if (bad) anybad[q] = true;
            }
        }
        // ---- results + guard (lane q writes column j0+q) ----
        const float cm0 = *P.cmax;
#pragma unroll
        for (int q = 0; q < G; ++q) {
            float xmq = xm[q];
#pragma unroll
            for (int o = 16; o; o >>= 1) xmq = fmaxf(xmq, __shfl_xor_sync(0xffffffffu, xmq, o));
            if (lane == q && j0 + q < P.n) {
                const int64_t j = j0 + q;
                if (anybad[q]) b2[q] = QNAN;                 // NaN / overflow somewhere: never certified
                const float cm = cm0 + xmq;                  // both the centre and the point were rounded to fp32
                const float gb = P.gb_unit * cm, ge = P.ge_unit * cm * cm + 1e-37f;
                bool certified;
                if (P.K == 1) certified = (b1[q] < INF);
                else {
                    const float E = P.ga * (b1[q] + b2[q]) + gb * (sqrtf(b1[q]) + sqrtf(b2[q])) + 2.f * ge;
                    certified = (b2[q] - b1[q]) > E;
                }
                P.assign[j] = i1[q];
                P.dist[j] = sqrtf(b1[q]);
                if (!certified) {
                    const int slot = atomicAdd(P.nflag, 1);
                    P.flagged[slot] = (int32_t)j;
                }
            }
        }
    }
}

// fp64 re-evaluation of flagged columns: warp per column, lanes over centres, rows in order.
template <typename XT>
__global__ void k_dense_exact(int64_t p, int64_t K, const XT *__restrict__ xraw, double scale,
                              const double *__restrict__ ct /* row-major [p+1][K] */, const int32_t *__restrict__ flagged,
                              const int *__restrict__ nflag, int32_t *__restrict__ assign, float *__restrict__ dist32,
                              double *__restrict__ dist64, int64_t n_all)
{
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = flagged ? (int64_t)*nflag : n_all;
    for (; w < total; w += nwarps) {
        const int64_t j = flagged ? (int64_t)flagged[w] : w;
        const XT *xc = xraw + j * p;
        double bv = 0.0;
        int bk = -1;
        for (int64_t k = lane; k < K; k += 32) {
            double s = 0.0;
            for (int64_t r = 0; r < p; ++r) {
                const double d = __dsub_rn(__dmul_rn((double)xc[r], scale), ct[r * K + k]);
                s = __dadd_rn(s, __dmul_rn(d, d));
            }
            const double v = __dsqrt_rn(s);
            if (v != v) continue;
            if (bk < 0 || v < bv) { bv = v; bk = (int)k; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (ok >= 0 && (bk < 0 || ov < bv || (ov == bv && ok < bk))) { bv = ov; bk = ok; }
        }
        if (lane == 0) {
            if (bk < 0) { bv = __longlong_as_double(0x7ff8000000000000LL); bk = 0; }
            assign[j] = bk;
            if (dist32) dist32[j] = (float)bv;
            if (dist64) dist64[j] = bv;
        }
    }
}

// per-cluster sums of dense columns.  grid.x = row tiles of 128, grid.y = column ranges.
template <int UC>                                             // columns in flight per thread
__global__ void __launch_bounds__(128) k_dense_sums(int64_t p, int64_t n, int kb, int k0,
                                                    const float *__restrict__ x, const int32_t *__restrict__ assign1,
                                                    double *__restrict__ S /* [K][p] */)
{
    extern __shared__ double bins[];                       // [kb][128]
    const int tid = threadIdx.x;
    for (int i = tid; i < kb * 128; i += 128) bins[i] = 0.0;
    __syncthreads();
    const int64_t r = (int64_t)blockIdx.x * 128 + tid;
    const int64_t per = (n + gridDim.y - 1) / gridDim.y;
    const int64_t ja = (int64_t)blockIdx.y * per, jb = min(n, ja + per);
    const bool live = r < p;
    int64_t j = ja;
    for (; j + UC <= jb; j += UC) {
        int a[UC];
        float v[UC];
#pragma unroll
        for (int u = 0; u < UC; ++u) a[u] = assign1[j + u] - 1 - k0;
#pragma unroll
        for (int u = 0; u < UC; ++u) v[u] = live ? __ldcs(x + (j + u) * p + r) : 0.f;
#pragma unroll
        for (int u = 0; u < UC; ++u)
            if ((unsigned)a[u] < (unsigned)kb) bins[a[u] * 128 + tid] += (double)v[u];
    }
    for (; j < jb; ++j) {
        const int a0 = assign1[j] - 1 - k0;
        const float v0 = live ? __ldcs(x + j * p + r) : 0.f;
        if ((unsigned)a0 < (unsigned)kb) bins[a0 * 128 + tid] += (double)v0;
    }
    if (live)
        for (int k = 0; k < kb; ++k) {
            const double v = bins[k * 128 + tid];
            if (v != 0.0 || v != v) atomicAdd(&S[(int64_t)(k0 + k) * p + r], v);
        }
}

__global__ void k_count_labels(int64_t n, int64_t K, const int32_t *__restrict__ assign1, unsigned long long *__restrict__ counts,
                               int *__restrict__ bad)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < n; j += stride) {
        const int a = assign1[j];
        if (a >= 1 && a <= K) atomicAdd(&counts[a - 1], 1ULL);
        else if (a != 0) atomicOr(bad, 1);                   // 0 = unassigned (EmptyAction 'drop'), ignored
    }
}

template <typename T>
__global__ void k_cast_scaled(int64_t count, const T *__restrict__ x, double scale, float *__restrict__ y)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) y[i] = (float)((double)x[i] * scale);
}

template <int KC, int G>
int launch_dense_assign(skm_ctx *ctx, const DenseParams &P, size_t smem)
{
    auto kern = k_dense_assign<KC, G>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem));
    if (per_sm < 1) { skm_set_error("dense_assign<%d,%d> does not fit on an SM (smem %zu)", KC, G, smem); return SKM_ERR_UNSUPPORTED; }
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    const int64_t need = ((P.n + G - 1) / G * 32 + 255) / 256;
    if (blocks > need) blocks = need;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 256, smem, ctx->stream>>>(P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

struct DensePlan { int kc, ks, nchunks; bool smem_table; size_t smem; };

DensePlan dense_plan(const skm_ctx *ctx, int64_t p, int64_t K)
{
    DensePlan d;
    d.kc = K <= 4 ? 4 : (K <= 8 ? 8 : (K <= 12 ? 12 : 16));
    d.ks = skm_fast_stride(d.kc);
    d.nchunks = (int)((K + d.kc - 1) / d.kc);
    const size_t bytes = (size_t)(p + 1) * d.ks * d.nchunks * sizeof(float);
    d.smem_table = bytes <= (size_t)ctx->smem_optin - 2048;
    d.smem = d.smem_table ? ((bytes + 15) & ~(size_t)15) : 0;
    return d;
}

}  // namespace

// One chunk: x32 = device fp32 [nc][p]; results for local columns [0,nc) into assign/dist (device, offset by caller)
static int dense_chunk(skm_ctx *ctx, int64_t p, int64_t nc, int64_t K, const DensePlan &dp, const float *x32,
                       const void *xraw, int x_type, double scale, const float *table, const float *cmax,
                       const double *ct, int32_t *assign, float *dist, int32_t *flagged, int *nflag,
                       const int32_t *assign_in, double *S, void *tc_scratch, int64_t *tc_state)
{
    if (nc == 0) return SKM_OK;
    if (assign_in) {
        const size_t budget = (size_t)ctx->smem_optin - 1024;
        int kb = (int)std::min<int64_t>(K, (int64_t)(budget / (128 * sizeof(double))));
        if (kb > 96) kb = 96;                                      // keep a few CTAs resident
        const size_t smem = (size_t)kb * 128 * sizeof(double);
        // few CTAs fit beside many clusters' bins (K = 64: 64 KB each, 12 warps per SM): more columns in flight per thread then
        static const char *uce = getenv("SKM_DENSE_SUMS_UC");           // tuning knob
        const int uc = uce ? atoi(uce) : (kb > 24 ? 32 : 16);    // K = 10: 2.10 / 2.01 / 2.48 ms, K = 64: 5.67 / 4.55 / 3.89 ms at UC = 8 / 16 / 32 (n = 2e6)
        auto kern = uc >= 32 ? k_dense_sums<32> : (uc >= 16 ? k_dense_sums<16> : k_dense_sums<8>);
        SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned tiles = (unsigned)((p + 127) / 128);
        int64_t ranges = std::max<int64_t>(1, ((int64_t)ctx->sm_count * 16) / tiles);
        ranges = std::min<int64_t>(ranges, std::max<int64_t>(1, nc / 64));
        for (int64_t k0 = 0; k0 < K; k0 += kb) {
            const int kbb = (int)std::min<int64_t>(kb, K - k0);
            kern<<<dim3(tiles, (unsigned)ranges), 128, smem, ctx->stream>>>(p, nc, kbb, (int)k0, x32, assign_in, S);
            SKM_CHECK_LAUNCH(ctx);
        }
    }
    if (!assign) return SKM_OK;
    SKM_CUDA(cudaMemsetAsync(nflag, 0, sizeof(int), ctx->stream));
    bool done = false;
    if (tc_scratch && tc_state[0] >= 0) {
        // tensor-core filter + exact evaluation of its candidates (tcgemm.cu).  It pays off when the filter leaves
        // one or two candidates per point; on data without structure most points end up flagged, and after one
        // such chunk the CUDA-core kernel takes over for the rest of the call.
        SKM_TRY(skm_launch_dense_assign_tc(ctx, p, nc, K, x32, tc_scratch, cmax, assign, dist, flagged, nflag));
        int nf = 0;
        SKM_CUDA(cudaMemcpyAsync(&nf, nflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        SKM_CUDA(cudaStreamSynchronize(ctx->stream));
        tc_state[1] += 1;
        if (nf > nc / 8) {
            tc_state[0] = -1;                                       // stop using the filter for this call
            ctx->tc_chunks_dropped += 1;
            SKM_CUDA(cudaMemsetAsync(nflag, 0, sizeof(int), ctx->stream));
        } else { done = true; ctx->tc_chunks += 1; }
    }
    const double u = 5.9604644775390625e-08, m = (double)p;
    DenseParams P;
    P.x = x32; P.p = p; P.n = nc; P.table = table; P.ks = dp.ks; P.kc_unused = dp.kc; P.nchunks = dp.nchunks; P.K = (int)K;
    P.chunk_floats = (p + 1) * dp.ks;
    P.table_in_smem = dp.smem_table ? 1 : 0;
    P.ga = (float)(1.01 * (m + 5.0) * u);
    P.gb_unit = (float)(2.02 * u * sqrt(m));
    P.ge_unit = (float)(2.1 * u * u * m);
    P.cmax = cmax; P.assign = assign; P.dist = dist; P.flagged = flagged; P.nflag = nflag;
    int rc = SKM_OK;
    if (!done)
        switch (dp.kc) {
            case 4:  rc = launch_dense_assign<4, 4>(ctx, P, dp.smem); break;
            case 8:  rc = launch_dense_assign<8, 4>(ctx, P, dp.smem); break;
            case 12: rc = launch_dense_assign<12, 4>(ctx, P, dp.smem); break;
            default: rc = launch_dense_assign<16, 4>(ctx, P, dp.smem); break;
        }
    if (rc != SKM_OK) return rc;
    int64_t blocks = std::min<int64_t>((nc * 32 + 255) / 256, (int64_t)ctx->sm_count * 8);
    if (x_type == SKM_F32)
        k_dense_exact<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, K, (const float *)xraw, scale, ct, flagged, nflag, assign, dist, nullptr, nc);
    else
        k_dense_exact<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, K, (const double *)xraw, scale, ct, flagged, nflag, assign, dist, nullptr, nc);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

extern "C" int skm_second_pass(skm_ctx *ctx, int64_t p, int64_t n, const void *x, int x_type, int x_on_device,
                               double scale, const double *centers, int64_t K, const int32_t *assign_in,
                               double *centers_out, int64_t *counts_out, int32_t *assign_out, double *dist_out,
                               int64_t chunk_cols, int64_t *n_rechecked)
{
    SKM_REQUIRE(ctx, "ctx is NULL");
    SKM_CUDA(cudaSetDevice(ctx->device));
    SKM_REQUIRE(p >= 1 && n >= 0 && K >= 1, "bad dimensions");
    SKM_REQUIRE(x || n == 0, "x is NULL");
    SKM_REQUIRE(x_type == SKM_F32 || x_type == SKM_F64, "x_type must be SKM_F32 or SKM_F64");
    SKM_REQUIRE(!(assign_out || dist_out) || centers, "centers are needed for the assignment pass");
    SKM_REQUIRE(!centers_out || assign_in, "assign_in is needed for the centre pass");
    SKM_REQUIRE(scale == scale && scale != 0.0, "bad scale");
    const bool want_assign = assign_out || dist_out;
    const bool want_sums = centers_out != nullptr;
    if (n_rechecked) *n_rechecked = 0;
    const size_t xs = x_type == SKM_F32 ? 4 : 8;
    // chunks of ~256 MB when the data crosses PCIe (double-buffered staging); resident data needs no staging, and
    // larger launches cut the tail of the persistent tensor-core kernel (a wave of 148 CTAs x 128 points)
    if (chunk_cols <= 0) chunk_cols = std::max<int64_t>(1, (int64_t)((x_on_device ? 2048LL : 256LL) << 20) / (int64_t)(p * 4));
    chunk_cols = std::min<int64_t>(chunk_cols, std::max<int64_t>(n, 1));
    const bool need_cast = !(x_type == SKM_F32 && scale == 1.0 && x_on_device);

    DensePlan dp = dense_plan(ctx, p, K);
    DevBuf raw[2], x32, dcent, ct, table, cmax, d_assign, d_dist, flagged, nflag, d_in, S, counts, badflag, tc_scratch;
    int64_t tc_state[2] = {0, 0};                               // [0] < 0: filter switched off; [1] chunks it ran on
    if (want_assign) {
        SKM_TRY(dcent.alloc(sizeof(double) * p * K));
        SKM_TRY(ct.alloc(sizeof(double) * (p + 1) * K));
        SKM_TRY(table.alloc(sizeof(float) * ((size_t)(p + 1) * dp.ks * dp.nchunks + 4)));
        SKM_TRY(cmax.alloc(sizeof(float) * 4));
        SKM_TRY(d_assign.alloc(sizeof(int32_t) * std::max<int64_t>(n, 1)));
        SKM_TRY(d_dist.alloc(sizeof(float) * std::max<int64_t>(n, 1)));
        SKM_TRY(flagged.alloc(sizeof(int32_t) * std::max<int64_t>(chunk_cols, 1)));
        SKM_TRY(nflag.alloc(sizeof(int) * 4));
        SKM_CUDA(cudaMemcpyAsync(dcent.ptr, centers, sizeof(double) * p * K, cudaMemcpyHostToDevice, ctx->stream));
        SKM_TRY(skm_launch_prep_centers(ctx, p, K, dcent.as<double>(), 0, 1.0, ct.as<double>(), nullptr, nullptr));
        FastPlan fp;
        fp.kc = dp.kc; fp.ks = dp.ks; fp.nchunks = dp.nchunks; fp.smem = 0; fp.threads = 256; fp.global_table = false;
        fp.mode64 = false; fp.dual8 = false; fp.layout = 0; fp.boff = 0; fp.rows = p + 1;
        SKM_TRY(skm_launch_build_table(ctx, p, K, ct.as<double>(), fp, table.as<float>(), cmax.as<float>()));
        if (skm_tc_dense_usable(p, K)) {
            SKM_TRY(tc_scratch.alloc(skm_tc_scratch_bytes(p, K, chunk_cols)));
            SKM_TRY(skm_launch_tc_prep(ctx, p, K, ct.as<double>(), tc_scratch.ptr));
        }
    }
    if (want_sums) {
        SKM_TRY(d_in.alloc(sizeof(int32_t) * std::max<int64_t>(n, 1)));
        SKM_TRY(S.alloc(sizeof(double) * p * K));
        SKM_TRY(counts.alloc(sizeof(unsigned long long) * K));
        SKM_TRY(badflag.alloc(sizeof(int)));
        SKM_CUDA(cudaMemcpyAsync(d_in.ptr, assign_in, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx->stream));
        SKM_CUDA(cudaMemsetAsync(S.ptr, 0, sizeof(double) * p * K, ctx->stream));
        SKM_CUDA(cudaMemsetAsync(counts.ptr, 0, sizeof(unsigned long long) * K, ctx->stream));
        SKM_CUDA(cudaMemsetAsync(badflag.ptr, 0, sizeof(int), ctx->stream));
        if (n > 0) {
            k_count_labels<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
                n, K, d_in.as<int32_t>(), counts.as<unsigned long long>(), badflag.as<int>());
            SKM_CHECK_LAUNCH(ctx);
        }
    }
    if (!x_on_device) { SKM_TRY(raw[0].alloc(xs * p * chunk_cols)); SKM_TRY(raw[1].alloc(xs * p * chunk_cols)); }
    if (need_cast) SKM_TRY(x32.alloc(sizeof(float) * p * chunk_cols));

    const int64_t nchunks = n > 0 ? (n + chunk_cols - 1) / chunk_cols : 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    auto issue = [&](int64_t c) {
        const int64_t j0 = c * chunk_cols, nc = std::min(chunk_cols, n - j0);
        if (c >= 2) cudaStreamWaitEvent(copy_stream, freed[c & 1], 0);
        cudaMemcpyAsync(raw[c & 1].ptr, (const char *)x + (size_t)j0 * p * xs, xs * p * nc, cudaMemcpyHostToDevice, copy_stream);
        cudaEventRecord(up[c & 1], copy_stream);
    };
    int rc = SKM_OK;
    int64_t rechecked = 0;
    std::vector<int> h_nflag;
    DevBuf nflag_log;
    if (want_assign && nchunks > 0) { SKM_TRY(nflag_log.alloc(sizeof(int) * nchunks)); h_nflag.resize(nchunks); }
    // every allocation is done: from here on nothing returns before the stream and events are released
    if (!x_on_device) {
        SKM_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) { cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming); }
    }
    if (!x_on_device && nchunks > 0) issue(0);
    for (int64_t c = 0; c < nchunks && rc == SKM_OK; ++c) {
        if (!x_on_device && c + 1 < nchunks) issue(c + 1);
        const int64_t j0 = c * chunk_cols, nc = std::min(chunk_cols, n - j0);
        const void *xr = x_on_device ? (const void *)((const char *)x + (size_t)j0 * p * xs) : raw[c & 1].ptr;
        if (!x_on_device) cudaStreamWaitEvent(ctx->stream, up[c & 1], 0);
        const float *xf = (const float *)xr;
        if (need_cast) {
            const int64_t blocks = std::min<int64_t>((p * nc + 255) / 256, (int64_t)ctx->sm_count * 32);
            if (x_type == SKM_F32) k_cast_scaled<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p * nc, (const float *)xr, scale, x32.as<float>());
            else k_cast_scaled<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p * nc, (const double *)xr, scale, x32.as<float>());
            ctx->launches++;
            xf = x32.as<float>();
        }
        {
            SkmTimed t(ctx, SKM_T_ASSIGN);
            rc = dense_chunk(ctx, p, nc, K, dp, xf, xr, x_type, scale, table.as<float>(), cmax.as<float>(), ct.as<double>(),
                             want_assign ? d_assign.as<int32_t>() + j0 : nullptr, want_assign ? d_dist.as<float>() + j0 : nullptr,
                             flagged.as<int32_t>(), nflag.as<int>(), want_sums ? d_in.as<int32_t>() + j0 : nullptr, S.as<double>(),
                             tc_scratch.ptr, tc_state);
        }
        if (rc == SKM_OK && want_assign)
            cudaMemcpyAsync(nflag_log.as<int>() + c, nflag.ptr, sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream);
        if (!x_on_device) cudaEventRecord(freed[c & 1], ctx->stream);
    }
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (copy_stream) {
        cudaStreamSynchronize(copy_stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(up[i]); cudaEventDestroy(freed[i]); }
        cudaStreamDestroy(copy_stream);
    }
    if (rc != SKM_OK) return rc;
    if (e != cudaSuccess) { skm_set_error("second pass failed: %s", cudaGetErrorString(e)); return SKM_ERR_CUDA; }

    if (want_assign && n > 0) {
        SKM_CUDA(cudaMemcpy(h_nflag.data(), nflag_log.ptr, sizeof(int) * nchunks, cudaMemcpyDeviceToHost));
        for (int v : h_nflag) rechecked += v;
        if (assign_out) {
            SKM_CUDA(cudaMemcpy(assign_out, d_assign.ptr, sizeof(int32_t) * n, cudaMemcpyDeviceToHost));
            for (int64_t j = 0; j < n; ++j) assign_out[j] += 1;
        }
        if (dist_out) {
            std::vector<float> tmp(n);
            SKM_CUDA(cudaMemcpy(tmp.data(), d_dist.ptr, sizeof(float) * n, cudaMemcpyDeviceToHost));
            for (int64_t j = 0; j < n; ++j) dist_out[j] = (double)tmp[j];
        }
    }
    if (want_sums) {
        int bad = 0;
        SKM_CUDA(cudaMemcpy(&bad, badflag.ptr, sizeof(int), cudaMemcpyDeviceToHost));
        if (bad) { skm_set_error("assign_in holds labels outside 0..K"); return SKM_ERR_INVALID; }
        std::vector<unsigned long long> hc(K);
        SKM_CUDA(cudaMemcpy(hc.data(), counts.ptr, sizeof(unsigned long long) * K, cudaMemcpyDeviceToHost));
        SKM_CUDA(cudaMemcpy(centers_out, S.ptr, sizeof(double) * p * K, cudaMemcpyDeviceToHost));
        for (int64_t k = 0; k < K; ++k) {
            if (counts_out) counts_out[k] = (int64_t)hc[k];
            double *col = centers_out + k * p;
            if (hc[k] == 0) { for (int64_t r = 0; r < p; ++r) col[r] = 0.0; }        // centers_twoPass = zeros(p,K), kmeans_sparsified.m:545
            else { const double cnt = (double)hc[k]; for (int64_t r = 0; r < p; ++r) col[r] = col[r] / cnt; }
        }
    }
    if (n_rechecked) *n_rechecked = rechecked;
    return SKM_OK;
}
