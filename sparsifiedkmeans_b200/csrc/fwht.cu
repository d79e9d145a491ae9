// fwht.cu -- K4: fast Walsh-Hadamard precondition (replaces private/hadamard.c:57-92 and
// private/hadamard_pthreads.c:69-204, and the mix closure kmeans_sparsified.m:238-248,286-295).
//
// The transform is the unnormalised natural-order (Sylvester) WHT: radix-2 butterflies
// (a,b) -> (a+b, a-b) with strides 1,2,4,...  Only additions and subtractions occur, so as
// long as the stages are applied in that order the result is bit-identical to the
// reference's whatever the thread schedule -- the fp64 instantiation is the exact drop-in,
// the fp32 one is the fast path (with the sign flip fused on load, the 1/sqrt(p2) division
// fused on store, and optionally the fixed-count row sample fused on store).
#include "common.cuh"
#include <stdlib.h>
#include "philox.cuh"

namespace {

template <typename T> __device__ __forceinline__ T div_rn(T a, T b);
template <> __device__ __forceinline__ double div_rn<double>(double a, double b) { return __ddiv_rn(a, b); }
template <> __device__ __forceinline__ float div_rn<float>(float a, float b) { return __fdiv_rn(a, b); }

// all stages with stride < L on one contiguous segment of L elements held in shared memory
template <typename T>
__device__ __forceinline__ void smem_stages(T *s, int L)
{
    for (int h = 1; h < L; h <<= 1) {
        for (int q = threadIdx.x; q < (L >> 1); q += blockDim.x) {
            const int i = ((q / h) * (h << 1)) + (q % h);
            const T a = s[i], b = s[i + h];
            s[i] = a + b;
            s[i + h] = a - b;
        }
        __syncthreads();
    }
}

// grid: (m / L) segments x n columns, flattened.  Applies signs on load; applies the final
// division on store when the segment is the whole column.
template <typename T>
__global__ void k_fwht_segment(int64_t m, int64_t nseg_total, int L, T *__restrict__ x,
                               const T *__restrict__ signs, T divide_by, int whole)
{
    extern __shared__ __align__(16) unsigned char fw_raw[];
    T *s = reinterpret_cast<T *>(fw_raw);
    const int segs_per_col = (int)(m / L);
    for (int64_t seg = blockIdx.x; seg < nseg_total; seg += gridDim.x) {
        const int64_t col = seg / segs_per_col;
        const int64_t off = (seg % segs_per_col) * (int64_t)L;
        T *g = x + col * m + off;
        for (int i = threadIdx.x; i < L; i += blockDim.x) {
            T v = g[i];
            if (signs) v = v * signs[off + i];
            s[i] = v;
        }
        __syncthreads();
        smem_stages<T>(s, L);
        for (int i = threadIdx.x; i < L; i += blockDim.x) {
            T v = s[i];
            if (whole && divide_by != (T)0) v = div_rn<T>(v, divide_by);
            g[i] = v;
        }
        __syncthreads();
    }
}

// one butterfly stage with stride h >= L straight on global memory
template <typename T>
__global__ void k_fwht_stage(int64_t m, int64_t n, int64_t h, T *__restrict__ x)
{
    const int64_t half = m >> 1;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = half * n;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) {
        const int64_t col = idx / half, q = idx % half;
        const int64_t i = ((q / h) * (h << 1)) + (q % h);
        T *g = x + col * m;
        const T a = g[i], b = g[i + h];
        g[i] = a + b;
        g[i + h] = a - b;
    }
}

template <typename T>
__global__ void k_divide(int64_t total, T *__restrict__ x, T divide_by)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; idx < total; idx += stride) x[idx] = div_rn<T>(x[idx], divide_by);
}

template <typename T>
int fwht_any(skm_ctx *ctx, int64_t m, int64_t n, T *x, const T *signs, T divide_by)
{
    if (m < 2 || (m & (m - 1)) != 0) {
        skm_set_error("hadamard: number of rows must be a power of two >= 2 (got %lld)", (long long)m);
        return SKM_ERR_INVALID;
    }
    if (n == 0) return SKM_OK;
    const int64_t max_elems = (int64_t)(128 * 1024) / (int64_t)sizeof(T);
    int64_t L = m < max_elems ? m : max_elems;
    const size_t smem = (size_t)L * sizeof(T);
    auto kern = k_fwht_segment<T>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t nseg = (m / L) * n;
    int threads = (int)(L / 2 < 1024 ? (L / 2 < 32 ? 32 : L / 2) : 1024);
    int64_t blocks = nseg;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    const int whole = (L == m);
    kern<<<(unsigned)blocks, threads, smem, ctx->stream>>>(m, nseg, (int)L, x, signs, divide_by, whole);
    SKM_CHECK_LAUNCH(ctx);
    if (!whole) {
        int64_t total = (m >> 1) * n;
        int64_t b2 = (total + 255) / 256;
        if (b2 > cap * 4) b2 = cap * 4;
        for (int64_t h = L; h < m; h <<= 1) {
            k_fwht_stage<T><<<(unsigned)b2, 256, 0, ctx->stream>>>(m, n, h, x);
            SKM_CHECK_LAUNCH(ctx);
        }
        if (divide_by != (T)0) {
            k_divide<T><<<(unsigned)b2, 256, 0, ctx->stream>>>(m * n, x, divide_by);
            SKM_CHECK_LAUNCH(ctx);
        }
    }
    return SKM_OK;
}


// ---------------------------------------------------------------------------------------------
// fp32 fast path.  Index bits of a column of length p2 = 32*E*W are split as
//   [ warp w : log2 W | register slot j : log2 E | lane : 5 ]
// so every global and shared access of a warp is 32 consecutive words (coalesced, conflict-free).
// Stages over the j bits run in registers, stages over the lane bits with warp shuffles, and the
// stages over the warp bits after ONE exchange through shared memory.  The stage order differs
// from the reference's (1,2,4,...), which only matters for rounding: this path is compared to the
// reference to tolerance; the fp64 path above is the bit-exact one.
// ---------------------------------------------------------------------------------------------
template <int E>
__device__ __forceinline__ void wht_regs_and_lanes(float (&v)[E], int lane)
{
#pragma unroll
    for (int h = 1; h < E; h <<= 1) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
            if (!(j & h)) { const float a = v[j], b = v[j + h]; v[j] = a + b; v[j + h] = a - b; }
        }
    }
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        // upper lane of the pair: other - v, lower lane: v + other; as ONE fused multiply-add with a +-1
        // factor (exact: the product is v or -v, so the result is the same single rounding of the sum)
        const float sg = (lane & o) ? -1.f : 1.f;
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const float other = __shfl_xor_sync(0xffffffffu, v[j], o);
            v[j] = fmaf(sg, v[j], other);
        }
    }
}

// one warp per column, p2 = 32*E
template <int E>
__global__ void __launch_bounds__(256) k_fwht_warp(int64_t n, float *__restrict__ x, const float *__restrict__ signs, float inv_scale_div)
{
    const int lane = threadIdx.x & 31;
    int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t ncol_stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int P2 = 32 * E;
    for (; col < n; col += ncol_stride) {
        float *g = x + col * P2;
        float v[E];
#pragma unroll
        for (int j = 0; j < E; ++j) {
            const int idx = (j << 5) | lane;
            float t = __ldcs(g + idx);
            if (signs) t *= __ldg(signs + idx);
            v[j] = t;
        }
        wht_regs_and_lanes<E>(v, lane);
#pragma unroll
        for (int j = 0; j < E; ++j) {
            float t = v[j];
            if (inv_scale_div != 0.f) t = __fdiv_rn(t, inv_scale_div);
            __stcs(g + ((j << 5) | lane), t);
        }
    }
}

// W warps per column (W = 2..32), p2 = 1024*W, one CTA works on one column at a time
template <int W>
__global__ void __launch_bounds__(32 * W) k_fwht_cta(int64_t n, float *__restrict__ x, const float *__restrict__ signs, float divide_by)
{
    extern __shared__ __align__(16) unsigned char fc_raw[];
    float *s = reinterpret_cast<float *>(fc_raw);
    constexpr int T = 32 * W, P2 = 1024 * W, G = 32 / W;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t col = blockIdx.x; col < n; col += gridDim.x) {
        float *g = x + col * P2;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int idx = (w << 10) | (j << 5) | lane;
            float t = __ldcs(g + idx);
            if (signs) t *= __ldg(signs + idx);
            v[j] = t;
        }
        wht_regs_and_lanes<32>(v, lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) s[(w << 10) | (j << 5) | lane] = v[j];
        __syncthreads();
        // stages over the warp bits: thread handles G groups of W elements 1024 apart
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int jw = 0; jw < W; ++jw) v[i * W + jw] = s[(jw << 10) | (i * T + threadIdx.x)];
        }
#pragma unroll
        for (int h = 1; h < W; h <<= 1) {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                if (!((q % W) & h)) { const float a = v[q], b = v[q + h]; v[q] = a + b; v[q + h] = a - b; }
            }
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int jw = 0; jw < W; ++jw) {
                float t = v[i * W + jw];
                if (divide_by != 0.f) t = __fdiv_rn(t, divide_by);
                __stcs(g + ((jw << 10) | (i * T + threadIdx.x)), t);
            }
        }
        __syncthreads();
    }
}

// Variant without warp shuffles (an experiment, kept behind SKM_FWHT_X=1; NEGATIVE result, profiles/r2_fwht.md).  The five
// lane-bit stages of k_fwht_cta cost 160 SHFL + 160 FMA per thread and column; shuffles and shared-memory accesses share
// the same issue port, and the hypothesis was that this port, not HBM, bounds k_fwht_cta (ncu: l1tex 63 %, issue 55 %,
// DRAM 4.1 of 6.5 TB/s).  Here the warp's 32 x 32 block is TRANSPOSED through shared memory
// instead (padded stride 33: conflict-free both ways), so the lane bits become register-slot bits and all ten low stages
// run as plain register butterflies: 128 shared-memory instructions per thread and column instead of 224 port slots, and
// 160 fewer FMAs.  The padded transpose layout is exactly the padded natural layout the warp-bit exchange reads, so the
// values are written back in place and one __syncthreads separates the two uses.
template <int W>
__global__ void __launch_bounds__(32 * W) k_fwht_cta_x(int64_t n, float *__restrict__ x, const float *__restrict__ signs, float divide_by)
{
    extern __shared__ __align__(16) unsigned char fc_raw[];
    float *s = reinterpret_cast<float *>(fc_raw);          // W x 1056 floats: one pad word per 32
    constexpr int T = 32 * W, P2 = 1024 * W, G = 32 / W;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float *sw = s + w * 1056;
    for (int64_t col = blockIdx.x; col < n; col += gridDim.x) {
        float *g = x + col * P2;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int idx = (w << 10) | (j << 5) | lane;
            float t = __ldcs(g + idx);
            if (signs) t *= __ldg(signs + idx);
            v[j] = t;
        }
#pragma unroll
        for (int h = 1; h < 32; h <<= 1) {                  // index bits 5..9
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (!(j & h)) { const float a = v[j], b = v[j + h]; v[j] = a + b; v[j + h] = a - b; }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) sw[j * 33 + lane] = v[j];
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = sw[lane * 33 + j];          // (slot j, lane l) <- (slot l, lane j)
#pragma unroll
        for (int h = 1; h < 32; h <<= 1) {                  // index bits 0..4
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (!(j & h)) { const float a = v[j], b = v[j + h]; v[j] = a + b; v[j + h] = a - b; }
        }
        // element (w << 10) | (lane << 5) | j lives at padded address idx + (idx >> 5) = w * 1056 + lane * 33 + j
#pragma unroll
        for (int j = 0; j < 32; ++j) sw[lane * 33 + j] = v[j];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int jw = 0; jw < W; ++jw) {
                const int nidx = (jw << 10) | (i * T + threadIdx.x);
                v[i * W + jw] = s[nidx + (nidx >> 5)];
            }
        }
#pragma unroll
        for (int h = 1; h < W; h <<= 1) {                   // index bits 10..
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                if (!((q % W) & h)) { const float a = v[q], b = v[q + h]; v[q] = a + b; v[q + h] = a - b; }
            }
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int jw = 0; jw < W; ++jw) {
                float t = v[i * W + jw];
                if (divide_by != 0.f) t = __fdiv_rn(t, divide_by);
                __stcs(g + ((jw << 10) | (i * T + threadIdx.x)), t);
            }
        }
        __syncthreads();
    }
}

// Same transform with the NEXT column prefetched by the TMA while the current one is finished: a column of
// p2 >= 8192 fills the CTA's shared memory, so only one is resident per SM and, without this, nothing is in flight
// between the last load of a column and the first load of the next (ncu: profiles/r2_fwht.md).  The column buffer
// is only busy during the one exchange between the lane stages and the warp stages; as soon as every thread has
// read its operands back, one thread issues cp.async.bulk for the next column into the same buffer (mbarrier
// complete_tx), which then overlaps the last butterflies and the stores of the current column.
template <int W, int NBUF>
__global__ void __launch_bounds__(32 * W) k_fwht_cta_tma(int64_t n, float *__restrict__ x, const float *__restrict__ signs, float divide_by)
{
    // NBUF = 2 (columns of <= 32 KB): the column after the next one is prefetched into the buffer the current column has
    // just left, so a load is in flight during the whole of a column's processing, not only during its second half
    extern __shared__ __align__(128) unsigned char fc_raw[];
    float *sbase = reinterpret_cast<float *>(fc_raw);
    __shared__ uint64_t bar[NBUF];
    constexpr int T = 32 * W, P2 = 1024 * W, G = 32 / W;
    constexpr uint32_t BYTES = P2 * 4;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    auto prefetch = [&](int64_t col, int b) {                 // one thread
        const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[b]);
        const uint32_t s_a = (uint32_t)__cvta_generic_to_shared(sbase + (size_t)b * P2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic reads of the buffer come first
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(BYTES) : "memory");
        const char *src = reinterpret_cast<const char *>(x + col * P2);
        for (uint32_t off = 0; off < BYTES; off += 32768) {
            const uint32_t sz = BYTES - off < 32768u ? BYTES - off : 32768u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(s_a + off), "l"(src + off), "r"(sz), "r"(bar_a) : "memory");
        }
    };
    if (threadIdx.x == 0) {
        for (int b = 0; b < NBUF; ++b) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"((uint32_t)__cvta_generic_to_shared(&bar[b])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int b = 0; b < NBUF; ++b)
            if ((int64_t)blockIdx.x + (int64_t)b * gridDim.x < n) prefetch(blockIdx.x + (int64_t)b * gridDim.x, b);
    int64_t it = 0;
    for (int64_t col = blockIdx.x; col < n; col += gridDim.x, ++it) {
        float *g = x + col * P2;
        const int b = (int)(it % NBUF);
        float *s = sbase + (size_t)b * P2;
        {
            const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar[b]);
            const uint32_t parity = (uint32_t)((it / NBUF) & 1);
            uint32_t done = 0, spins = 0;
            while (!done) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar_a), "r"(parity) : "memory");
                if (!done && ++spins > (1u << 26)) __trap();
            }
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int idx = (w << 10) | (j << 5) | lane;
            float t = s[idx];
            if (signs) t *= __ldg(signs + idx);
            v[j] = t;
        }
        wht_regs_and_lanes<32>(v, lane);
        __syncthreads();                                       // everyone has read the raw column
#pragma unroll
        for (int j = 0; j < 32; ++j) s[(w << 10) | (j << 5) | lane] = v[j];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int jw = 0; jw < W; ++jw) v[i * W + jw] = s[(jw << 10) | (i * T + threadIdx.x)];
        }
        __syncthreads();                                       // the buffer is free again
        if (threadIdx.x == 0 && col + (int64_t)NBUF * gridDim.x < n) prefetch(col + (int64_t)NBUF * gridDim.x, b);
#pragma unroll
        for (int h = 1; h < W; h <<= 1) {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                if (!((q % W) & h)) { const float a = v[q], b2 = v[q + h]; v[q] = a + b2; v[q + h] = a - b2; }
            }
        }
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int jw = 0; jw < W; ++jw) {
                float t = v[i * W + jw];
                if (divide_by != 0.f) t = __fdiv_rn(t, divide_by);
                __stcs(g + ((jw << 10) | (i * T + threadIdx.x)), t);
            }
        }
    }
}

template <int E>
int launch_fwht_warp(skm_ctx *ctx, int64_t n, float *x, const float *signs, float divide_by)
{
    int64_t blocks = (n * 32 + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_fwht_warp<E><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, x, signs, divide_by);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

template <int W>
int launch_fwht_cta(skm_ctx *ctx, int64_t n, float *x, const float *signs, float divide_by)
{
    static const bool no_tma = getenv("SKM_FWHT_NO_TMA") != nullptr;
    static const char *xe = getenv("SKM_FWHT_X");           // 1: shuffle-free variant for every size, 0: never
    const bool use_x = xe ? atoi(xe) != 0 : false;          // measured 3-14 % SLOWER than the shuffle kernels at every size: opt-in only
    size_t smem = use_x ? (size_t)1056 * W * sizeof(float) : (size_t)1024 * W * sizeof(float);
    // TMA prefetch of the next column for every CTA size (p2 = 2048 / 4096: 0.71 / 0.67 -> 0.77 / 0.77 of the HBM peak).  A second
    // column buffer (SKM_FWHT_TMA_NBUF=2, W <= 8) halves the resident CTAs and measured SLOWER: 0.70 / 0.73 / 0.71 at 2048 /
    // 4096 / 8192 against 0.77 / 0.77 / 0.74 (profiles/r2_fwht.md), so it stays an experiment.
    static const char *nb = getenv("SKM_FWHT_TMA_NBUF");
    const bool tma = !use_x && !no_tma;
    const bool two = tma && W <= 8 && nb && atoi(nb) == 2;
    if (tma) smem = (size_t)(two ? 2 : 1) * 1024 * W * sizeof(float);
    auto kern = use_x ? k_fwht_cta_x<W> : (tma ? (two ? k_fwht_cta_tma<W, 2> : k_fwht_cta_tma<W, 1>) : k_fwht_cta<W>);
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * W, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    if (blocks > n) blocks = n;
    kern<<<(unsigned)blocks, 32 * W, smem, ctx->stream>>>(n, x, signs, divide_by);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// returns SKM_ERR_UNSUPPORTED when the size has no fast kernel (caller falls back)
int fwht_f32_fast(skm_ctx *ctx, int64_t m, int64_t n, float *x, const float *signs, float divide_by)
{
    switch (m) {
        case 32: return launch_fwht_warp<1>(ctx, n, x, signs, divide_by);
        case 64: return launch_fwht_warp<2>(ctx, n, x, signs, divide_by);
        case 128: return launch_fwht_warp<4>(ctx, n, x, signs, divide_by);
        case 256: return launch_fwht_warp<8>(ctx, n, x, signs, divide_by);
        case 512: return launch_fwht_warp<16>(ctx, n, x, signs, divide_by);
        case 1024: return launch_fwht_warp<32>(ctx, n, x, signs, divide_by);
        case 2048: return launch_fwht_cta<2>(ctx, n, x, signs, divide_by);
        case 4096: return launch_fwht_cta<4>(ctx, n, x, signs, divide_by);
        case 8192: return launch_fwht_cta<8>(ctx, n, x, signs, divide_by);
        case 16384: return launch_fwht_cta<16>(ctx, n, x, signs, divide_by);
        case 32768: return launch_fwht_cta<32>(ctx, n, x, signs, divide_by);
        default: return SKM_ERR_UNSUPPORTED;
    }
}

// ---- on-device row sampler -------------------------------------------------------------------
// Contract of private/randsample_fixedNumberEntries.m:44-63 / randsample_block.m:44-84: every
// column keeps exactly m distinct rows, uniformly among all m-subsets.  MATLAB's generator is
// closed source, so the draws come from Philox4x32-10 keyed by the seed and counted by
// (global column, draw index, attempt): the sample of a column does not depend on how columns
// are sharded over GPUs or scheduled over threads.
// Marks m distinct uniformly random rows of [0, P2) in `bits` (P2/32 words, already zero).
// `owner` is scratch of P2 ints shared by the T cooperating threads; barrier_any(pred) must
// synchronise those T threads and return whether pred holds for any of them.  Draw i proposes
// philox(seed; col, i, attempt) mod P2; among the draws proposing the same free row in a round
// the smallest i wins, the others redraw -- a pure function of (seed, col), whatever the timing.
template <class BarrierAny>
__device__ __forceinline__ void mark_random_rows(uint32_t *bits, int *owner, int P2, int m, uint64_t seed, int64_t col,
                                                 int tid, int T, BarrierAny barrier_any)
{
    for (int i = tid; i < P2; i += T) owner[i] = 0x7fffffff;
    int draw = tid;
    uint32_t attempt = 0;
    bool pending = draw < m;
    while (barrier_any(pending)) {
        int r = -1;
        if (pending) {
            r = (int)(skm_philox_draw(seed, (uint64_t)col, (uint32_t)draw, attempt) & (uint32_t)(P2 - 1));
            if ((bits[r >> 5] >> (r & 31)) & 1u) { r = -1; ++attempt; }     // taken in an earlier round
            else atomicMin(&owner[r], draw);
        }
        barrier_any(false);
        if (pending && r >= 0) {
            if (owner[r] == draw) {
                atomicOr(&bits[r >> 5], 1u << (r & 31));
                draw += T; attempt = 0; pending = draw < m;
            } else ++attempt;
        }
    }
}

// __syncwarp orders the lanes' shared-memory accesses (a vote alone converges the warp but is no memory barrier)
struct WarpBarrierAny { __device__ bool operator()(bool p) const { __syncwarp(); const bool r = __any_sync(0xffffffffu, p); __syncwarp(); return r; } };
struct CtaBarrierAny { __device__ bool operator()(bool p) const { return __syncthreads_or(p) != 0; } };

// rows_out[col*m .. +m) = the sampled rows of global column col0+col, ascending
__global__ void k_sample_rows(int P2, int64_t n, int m, uint64_t seed, int64_t col0, int32_t *__restrict__ rows_out)
{
    extern __shared__ __align__(16) unsigned char sr_raw[];
    int *owner = reinterpret_cast<int *>(sr_raw);
    const int nwords = P2 >> 5;
    uint32_t *bits = reinterpret_cast<uint32_t *>(owner + P2);
    int *scan = reinterpret_cast<int *>(bits + nwords);
    const int T = blockDim.x;
    for (int64_t col = blockIdx.x; col < n; col += gridDim.x) {
        for (int w = threadIdx.x; w < nwords; w += T) bits[w] = 0u;
        __syncthreads();
        mark_random_rows(bits, owner, P2, m, seed, col0 + col, threadIdx.x, T, CtaBarrierAny());
        __syncthreads();
        if (threadIdx.x == 0) {
            int run = 0;
            for (int w = 0; w < nwords; ++w) { scan[w] = run; run += __popc(bits[w]); }
        }
        __syncthreads();
        for (int w = threadIdx.x; w < nwords; w += T) {
            uint32_t b = bits[w];
            int64_t out = col * (int64_t)m + scan[w];
            while (b) { const int bit = __ffs(b) - 1; b &= b - 1; rows_out[out++] = (w << 5) + bit; }
        }
        __syncthreads();
    }
}

// ---- fused precondition + row sample on the fast transform --------------------------------
// After the transform the column sits in shared memory; the m sampled rows are marked in a
// bitmap (one 32-bit word per lane / thread), an exclusive scan of the popcounts gives every
// word its output offset, and the set bits are emitted in ascending row order as (row, value)
// with value = (h / sqrt(p2)) / (m / p2)   (randsample_fixedNumberEntries.m:30-31,62).
template <int E>
__global__ void __launch_bounds__(256) k_fwht_sample_warp(int64_t n, int m_keep, const float *__restrict__ x,
                                                          const float *__restrict__ signs, const int32_t *__restrict__ rows,
                                                          uint64_t seed, int64_t col0,
                                                          int64_t *__restrict__ colptr, int32_t *__restrict__ rowidx,
                                                          float *__restrict__ val, int *__restrict__ bad_flag)
{
    constexpr int P2 = 32 * E;
    __shared__ float s_col[8][P2];
    __shared__ uint32_t s_bits[8][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float *s = s_col[wib];
    uint32_t *bits = s_bits[wib];
    int64_t col = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float root = sqrtf((float)P2), level = __fdiv_rn((float)m_keep, (float)P2);
    for (; col < n; col += stride) {
        // issue the column's loads first: the sample does not depend on the data, so marking it (row list
        // reads or the Philox draws) runs while they are in flight instead of in front of them
        const float *g = x + col * P2;
        float v[E];
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = __ldcs(g + ((j << 5) | lane));
        bits[lane] = 0u;
        __syncwarp();
        if (rows) {
            const int32_t *rr = rows + col * (int64_t)m_keep;
            for (int i = lane; i < m_keep; i += 32) {
                const int r = rr[i];
                if ((unsigned)r < (unsigned)P2) atomicOr(&bits[r >> 5], 1u << (r & 31));
                else *bad_flag = 1;
            }
        } else {
            mark_random_rows(bits, reinterpret_cast<int *>(s), P2, m_keep, seed, col0 + col, lane, 32, WarpBarrierAny());
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] *= __ldg(signs + ((j << 5) | lane));
        wht_regs_and_lanes<E>(v, lane);
#pragma unroll
        for (int j = 0; j < E; ++j) s[(j << 5) | lane] = v[j];
        __syncwarp();
        uint32_t b = lane < E ? bits[lane] : 0u;
        const int cnt = __popc(b);
        int off = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += t; }
        const int total = __shfl_sync(0xffffffffu, off, 31);
        if (total != m_keep) *bad_flag = 1;                 // repeated rows inside a column
        // emit: output o is the (o - base_L)-th set bit of word L, where L is the first word whose
        // inclusive prefix exceeds o (binary search over the prefixes with shuffles).  One output per
        // lane, so the divisions and the global stores run 32 wide and the stores are coalesced.
        const int base_l = off - cnt;
        const int64_t out0 = col * (int64_t)m_keep;
        for (int ob = 0; ob < m_keep; ob += 32) {
            const int o = ob + lane;
            int pos = 0;
#pragma unroll
            for (int sft = 16; sft; sft >>= 1) {
                const int tt = __shfl_sync(0xffffffffu, off, pos + sft - 1);
                if (tt <= o) pos += sft;
            }
            uint32_t bl = __shfl_sync(0xffffffffu, b, pos);
            const int k = o - __shfl_sync(0xffffffffu, base_l, pos);
            if (o < m_keep && o < total) {
                for (int i = 0; i < k; ++i) bl &= bl - 1;
                const int r = (pos << 5) + __ffs(bl) - 1;
                rowidx[out0 + o] = r;
                val[out0 + o] = __fdiv_rn(__fdiv_rn(s[r], root), level);
            }
        }
        if (lane == 0) { colptr[col] = col * (int64_t)m_keep; if (col == n - 1) colptr[n] = n * (int64_t)m_keep; }
        __syncwarp();
    }
}

template <int W>
__global__ void __launch_bounds__(32 * W) k_fwht_sample_cta(int64_t n, int m_keep, const float *__restrict__ x,
                                                            const float *__restrict__ signs, const int32_t *__restrict__ rows,
                                                            uint64_t seed, int64_t col0,
                                                            int64_t *__restrict__ colptr, int32_t *__restrict__ rowidx,
                                                            float *__restrict__ val, int *__restrict__ bad_flag)
{
    extern __shared__ __align__(16) unsigned char fsc_raw[];
    constexpr int T = 32 * W, P2 = 1024 * W, G = 32 / W;
    float *s = reinterpret_cast<float *>(fsc_raw);
    uint32_t *bits = reinterpret_cast<uint32_t *>(s + P2);          // [T] one word per thread
    int *wsum = reinterpret_cast<int *>(bits + T);                  // [W] warp totals
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float root = sqrtf((float)P2), level = __fdiv_rn((float)m_keep, (float)P2);
    for (int64_t col = blockIdx.x; col < n; col += gridDim.x) {
        // loads first, marking while they are in flight (see k_fwht_sample_warp)
        const float *g = x + col * P2;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __ldcs(g + ((w << 10) | (j << 5) | lane));
        bits[threadIdx.x] = 0u;
        __syncthreads();
        if (rows) {
            const int32_t *rr = rows + col * (int64_t)m_keep;
            for (int i = threadIdx.x; i < m_keep; i += T) {
                const int r = rr[i];
                if ((unsigned)r < (unsigned)P2) atomicOr(&bits[r >> 5], 1u << (r & 31));
                else *bad_flag = 1;
            }
        } else {
            mark_random_rows(bits, reinterpret_cast<int *>(s), P2, m_keep, seed, col0 + col, threadIdx.x, T, CtaBarrierAny());
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= __ldg(signs + ((w << 10) | (j << 5) | lane));
        wht_regs_and_lanes<32>(v, lane);
#pragma unroll
        for (int j = 0; j < 32; ++j) s[(w << 10) | (j << 5) | lane] = v[j];
        __syncthreads();
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int jw = 0; jw < W; ++jw) v[i * W + jw] = s[(jw << 10) | (i * T + threadIdx.x)];
        }
#pragma unroll
        for (int h = 1; h < W; h <<= 1) {
#pragma unroll
            for (int q = 0; q < 32; ++q) {
                if (!((q % W) & h)) { const float a = v[q], b = v[q + h]; v[q] = a + b; v[q + h] = a - b; }
            }
        }
        __syncthreads();                                            // everyone has read the old contents
#pragma unroll
        for (int i = 0; i < G; ++i) {
#pragma unroll
            for (int jw = 0; jw < W; ++jw) s[(jw << 10) | (i * T + threadIdx.x)] = v[i * W + jw];
        }
        __syncthreads();
        uint32_t b = bits[threadIdx.x];
        const int cnt = __popc(b);
        int off = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += t; }
        if (lane == 31) wsum[w] = off;
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < W; ++q) { const int t = wsum[q]; if (q < w) before += t; total += t; }
        if (threadIdx.x == 0 && total != m_keep) *bad_flag = 1;
        int64_t out = col * (int64_t)m_keep + before + (off - cnt);
        while (b) {
            const int bit = __ffs(b) - 1;
            b &= b - 1;
            const int r = ((int)threadIdx.x << 5) + bit;
            rowidx[out] = r;
            val[out] = __fdiv_rn(__fdiv_rn(s[r], root), level);
            ++out;
        }
        if (threadIdx.x == 0) { colptr[col] = col * (int64_t)m_keep; if (col == n - 1) colptr[n] = n * (int64_t)m_keep; }
        __syncthreads();
    }
}

template <int E>
int launch_sample_warp(skm_ctx *ctx, int64_t n, int m, const float *x, const float *signs, const int32_t *rows,
                       uint64_t seed, int64_t col0, int64_t *colptr, int32_t *rowidx, float *val, int *bad)
{
    int64_t blocks = (n * 32 + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 8;
    if (blocks > cap) blocks = cap;
    k_fwht_sample_warp<E><<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, m, x, signs, rows, seed, col0, colptr, rowidx, val, bad);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

template <int W>
int launch_sample_cta(skm_ctx *ctx, int64_t n, int m, const float *x, const float *signs, const int32_t *rows,
                      uint64_t seed, int64_t col0, int64_t *colptr, int32_t *rowidx, float *val, int *bad)
{
    const size_t smem = (size_t)1024 * W * sizeof(float) + (size_t)32 * W * 4 + (size_t)W * 4 + 16;
    auto kern = k_fwht_sample_cta<W>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * W, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    if (blocks > n) blocks = n;
    kern<<<(unsigned)blocks, 32 * W, smem, ctx->stream>>>(n, m, x, signs, rows, seed, col0, colptr, rowidx, val, bad);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int fwht_sample_fast(skm_ctx *ctx, int64_t p2, int64_t n, int m, const float *x, const float *signs, const int32_t *rows,
                     uint64_t seed, int64_t col0, int64_t *colptr, int32_t *rowidx, float *val, int *bad)
{
#define SKM_SW(E) return launch_sample_warp<E>(ctx, n, m, x, signs, rows, seed, col0, colptr, rowidx, val, bad)
#define SKM_SC(W) return launch_sample_cta<W>(ctx, n, m, x, signs, rows, seed, col0, colptr, rowidx, val, bad)
    switch (p2) {
        case 32: SKM_SW(1); case 64: SKM_SW(2); case 128: SKM_SW(4); case 256: SKM_SW(8); case 512: SKM_SW(16);
        case 1024: SKM_SW(32);
        case 2048: SKM_SC(2); case 4096: SKM_SC(4); case 8192: SKM_SC(8); case 16384: SKM_SC(16); case 32768: SKM_SC(32);
        default: return SKM_ERR_UNSUPPORTED;
    }
#undef SKM_SW
#undef SKM_SC
}

// One CTA per column: sign flip, full FWHT in shared memory, then keep exactly m_keep rows
// (ascending) with value (h / sqrt(p2)) / (m_keep / p2).
__global__ void k_fwht_sample(int64_t p2, int64_t n, int m_keep, const float *__restrict__ x,
                              const float *__restrict__ signs, const int32_t *__restrict__ rows,
                              int64_t *__restrict__ colptr, int32_t *__restrict__ rowidx,
                              float *__restrict__ val, int *__restrict__ bad_flag)
{
    extern __shared__ __align__(16) unsigned char fs_raw[];
    float *s = reinterpret_cast<float *>(fs_raw);
    const int P2 = (int)p2;
    const int nwords = (P2 + 31) >> 5;
    uint32_t *bits = reinterpret_cast<uint32_t *>(s + P2);
    int *scan = reinterpret_cast<int *>(bits + nwords);          // [blockDim.x + 1]
    const float root = sqrtf((float)P2);
    const float level = __fdiv_rn((float)m_keep, (float)P2);

    for (int64_t col = blockIdx.x; col < n; col += gridDim.x) {
        const float *g = x + col * p2;
        for (int i = threadIdx.x; i < P2; i += blockDim.x) s[i] = g[i] * signs[i];
        for (int w = threadIdx.x; w < nwords; w += blockDim.x) bits[w] = 0u;
        __syncthreads();
        smem_stages<float>(s, P2);
        const int32_t *rr = rows + col * (int64_t)m_keep;
        for (int i = threadIdx.x; i < m_keep; i += blockDim.x) {
            const int r = rr[i];
            if ((unsigned)r < (unsigned)P2) atomicOr(&bits[r >> 5], 1u << (r & 31));
            else *bad_flag = 1;                        // reported by the caller; never corrupts memory
        }
        __syncthreads();
        // ordered compaction of the set bits
        const int wpt = (nwords + blockDim.x - 1) / blockDim.x;
        const int w0 = threadIdx.x * wpt, w1 = min(nwords, w0 + wpt);
        int cnt = 0;
        for (int w = w0; w < w1; ++w) cnt += __popc(bits[w]);
        scan[threadIdx.x + 1] = cnt;
        if (threadIdx.x == 0) scan[0] = 0;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int t = 1; t <= (int)blockDim.x; ++t) scan[t] += scan[t - 1];
        __syncthreads();
        int64_t out = col * (int64_t)m_keep + scan[threadIdx.x];
        for (int w = w0; w < w1; ++w) {
            uint32_t b = bits[w];
            while (b) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                const int r = (w << 5) + bit;
                rowidx[out] = r;
                val[out] = __fdiv_rn(__fdiv_rn(s[r], root), level);
                ++out;
            }
        }
        if (threadIdx.x == 0) {
            colptr[col] = col * (int64_t)m_keep;
            if (col == n - 1) colptr[n] = n * (int64_t)m_keep;
        }
        __syncthreads();
    }
}

}  // namespace

int skm_launch_fwht_f64(skm_ctx *ctx, int64_t m, int64_t n, double *x, const double *signs, double divide_by)
{
    return fwht_any<double>(ctx, m, n, x, signs, divide_by);
}

int skm_launch_fwht_f32(skm_ctx *ctx, int64_t m, int64_t n, float *x, const float *signs, float divide_by)
{
    if (n > 0) {
        SkmTimed t(ctx, SKM_T_FWHT);
        const int rc = fwht_f32_fast(ctx, m, n, x, signs, divide_by);
        if (rc != SKM_ERR_UNSUPPORTED) return rc;
    }
    return fwht_any<float>(ctx, m, n, x, signs, divide_by);
}

int skm_launch_sample_rows(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, uint64_t seed, int64_t col0, int32_t *rows_out)
{
    if (p2 < 32 || (p2 & (p2 - 1)) != 0) { skm_set_error("sample_rows: p2 must be a power of two >= 32"); return SKM_ERR_INVALID; }
    if (m < 1 || m > p2) { skm_set_error("sample_rows: need 1 <= m <= p2"); return SKM_ERR_INVALID; }
    if (n == 0) return SKM_OK;
    const int threads = (int)(p2 / 32 < 1024 ? (p2 / 32 < 32 ? 32 : p2 / 32) : 1024);
    const size_t smem = (size_t)p2 * 4 + (size_t)(p2 / 32) * 8 + 16;
    if (smem > (size_t)ctx->smem_optin) { skm_set_error("sample_rows: p2=%lld does not fit in shared memory", (long long)p2); return SKM_ERR_UNSUPPORTED; }
    SKM_CUDA(cudaFuncSetAttribute(k_sample_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sample_rows, threads, smem));
    int64_t blocks = (int64_t)ctx->sm_count * (per_sm < 1 ? 1 : per_sm);
    if (blocks > n) blocks = n;
    k_sample_rows<<<(unsigned)blocks, threads, smem, ctx->stream>>>((int)p2, n, (int)m, seed, col0, rows_out);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_fwht_sample_f32(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, const float *x,
                               const float *signs, const int32_t *rows, uint64_t seed, int64_t col0,
                               int64_t *colptr, int32_t *rowidx, float *val)
{
    if (p2 < 2 || (p2 & (p2 - 1)) != 0) {
        skm_set_error("fwht_sample: p2 must be a power of two >= 2");
        return SKM_ERR_INVALID;
    }
    if (m < 1 || m > p2) { skm_set_error("fwht_sample: need 1 <= m <= p2"); return SKM_ERR_INVALID; }
    int threads = (int)(p2 / 2 < 1024 ? (p2 / 2 < 32 ? 32 : p2 / 2) : 1024);
    size_t smem = (size_t)p2 * sizeof(float) + (size_t)((p2 + 31) / 32) * 4 + (size_t)(threads + 1) * 4 + 16;
    if (smem > (size_t)ctx->smem_optin) {
        skm_set_error("fwht_sample: p2=%lld does not fit in shared memory", (long long)p2);
        return SKM_ERR_UNSUPPORTED;
    }
    if (n == 0) return SKM_OK;
    SKM_CUDA(cudaFuncSetAttribute(k_fwht_sample, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fwht_sample, threads, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    if (blocks > n) blocks = n;
    SKM_CUDA(cudaMemsetAsync(ctx->d_flag + 8, 0, sizeof(int), ctx->stream));
    {
        SkmTimed timed(ctx, SKM_T_FWHT);
        int rc = fwht_sample_fast(ctx, p2, n, (int)m, x, signs, rows, seed, col0, colptr, rowidx, val, ctx->d_flag + 8);
        if (rc == SKM_ERR_UNSUPPORTED && !rows) {
            skm_set_error("fwht_sample: on-device row sampling needs 32 <= p2 <= 32768");
            return SKM_ERR_UNSUPPORTED;
        }
        if (rc == SKM_ERR_UNSUPPORTED) {
            k_fwht_sample<<<(unsigned)blocks, threads, smem, ctx->stream>>>(p2, n, (int)m, x, signs, rows, colptr,
                                                                           rowidx, val, ctx->d_flag + 8);
            SKM_CHECK_LAUNCH(ctx);
        } else if (rc != SKM_OK) return rc;
    }
    SKM_CUDA(cudaMemcpyAsync(ctx->h_flag + 8, ctx->d_flag + 8, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_flag[8]) {
        skm_set_error("fwht_sample: a sampled row index is outside [0, p2) (or rows repeat within a column)");
        return SKM_ERR_INVALID;
    }
    return SKM_OK;
}
