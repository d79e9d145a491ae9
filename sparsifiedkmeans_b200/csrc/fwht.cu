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

// One CTA per column: sign flip, full FWHT in shared memory, then keep exactly m_keep rows
// (ascending) with value (h / sqrt(p2)) / (m_keep / p2).
__global__ void k_fwht_sample(int64_t p2, int64_t n, int m_keep, const float *__restrict__ x,
                              const float *__restrict__ signs, const int32_t *__restrict__ rows,
                              int64_t *__restrict__ colptr, int32_t *__restrict__ rowidx,
                              float *__restrict__ val)
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
            atomicOr(&bits[r >> 5], 1u << (r & 31));
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
    return fwht_any<float>(ctx, m, n, x, signs, divide_by);
}

int skm_launch_fwht_sample_f32(skm_ctx *ctx, int64_t p2, int64_t n, int64_t m, const float *x,
                               const float *signs, const int32_t *rows, int64_t *colptr,
                               int32_t *rowidx, float *val)
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
    k_fwht_sample<<<(unsigned)blocks, threads, smem, ctx->stream>>>(p2, n, (int)m, x, signs, rows, colptr,
                                                                   rowidx, val);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}
