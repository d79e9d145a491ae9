// dct.cu -- the DCT preconditioner: what kmeans_sparsified.m picks
// when p is not a power of two (`SketchType` 'auto' -> 'DCT', :226-231; H = dct, Ht = idct, :256-258),
// e.g. MNIST's p = 784.
//
//   mix(X)   = dct( D * X )          unmix(C) = D * idct( C )         (kmeans_sparsified.m:295-296)
//
// MATLAB's dct is the orthonormal DCT-II along columns, y = T x with
//   T[k][i] = w_k cos(pi (2i+1) k / (2p)),  w_0 = sqrt(1/p), w_k = sqrt(2/p);  idct is T'.
// For p without a fast power-of-two structure the transform is applied as ONE dense product with
// the p x p matrix M = T * diag(d) * (1+2eps) (signs and the reference's *(1+2eps), :292, folded in
// on the host in fp64).  The data pipeline runs it on the tensor cores: tcgen05.mma kind::tf32 with both
// operands split into tf32-exact halves and three products per tile (3xTF32, fp32 accuracy; k_tc_dct in
// tcgemm.cu, TMA-fed, TMEM accumulators); the small fp64 products of mix/unmix on p x K centre matrices use
// a plain tiled fp64 kernel.  No library GEMM is involved.  What follows the product is the fixed-count row
// sampler for arbitrary p (Philox, same contract as the Hadamard path) fused with the gather that writes CSC
// directly, so the mixed dense chunk never leaves the device.
#include "common.cuh"
#include "philox.cuh"
#include <algorithm>
#include <math.h>
#include <string.h>
#include <vector>

namespace {

// M (column-major p x p) = T * diag(d) * scale (forward) or its transpose-inverse diag(d) * T' (inverse)
void dct_matrix(int64_t p, const double *signs, double scale, bool inverse, std::vector<double> &M)
{
    M.resize((size_t)p * p);
    const double w0 = sqrt(1.0 / (double)p), w = sqrt(2.0 / (double)p), pi = 3.14159265358979323846;
    for (int64_t i = 0; i < p; ++i) {
        const double di = signs ? signs[i] : 1.0;
        for (int64_t k = 0; k < p; ++k) {
            const double t = (k == 0 ? w0 : w) * cos(pi * (double)(2 * i + 1) * (double)k / (double)(2 * p));
            // forward: Y = M X with M[k][i] = t * d_i * scale ; inverse: X = M Y with M[i][k] = d_i * t
            if (!inverse) M[(size_t)i * p + k] = t * di * scale;
            else          M[(size_t)k * p + i] = di * t;
        }
    }
}

// ---- general-p row sampler + gather ------------------------------------------------------------
// One CTA per column: m distinct rows of [0,p), uniform over m-subsets (rejection; draw i proposes
// floor(philox * p / 2^32), the smallest draw index wins a contested row, the others redraw -- a pure
// function of (seed, global column)), then the rows are emitted ascending with value y / (m/p)
// (private/randsample_fixedNumberEntries.m:30-31,62).  rows_in != NULL: use those rows instead.
__global__ void k_sample_gather(int p, int64_t n, int m, const float *__restrict__ y, const int32_t *__restrict__ rows_in,
                                uint64_t seed, int64_t col0, int64_t *__restrict__ colptr, int32_t *__restrict__ rowidx,
                                float *__restrict__ val, int32_t *__restrict__ rows_out, int *__restrict__ bad_flag)
{
    extern __shared__ __align__(16) unsigned char sg_raw[];
    int *owner = reinterpret_cast<int *>(sg_raw);
    const int nwords = (p + 31) >> 5;
    uint32_t *bits = reinterpret_cast<uint32_t *>(owner + p);
    int *scan = reinterpret_cast<int *>(bits + nwords);
    const int T = blockDim.x, tid = threadIdx.x;
    const float level = __fdiv_rn((float)m, (float)p);
    for (int64_t col = blockIdx.x; col < n; col += gridDim.x) {
        for (int w = tid; w < nwords; w += T) bits[w] = 0u;
        for (int i = tid; i < p; i += T) owner[i] = 0x7fffffff;
        __syncthreads();
        if (rows_in) {
            for (int i = tid; i < m; i += T) {
                const int r = rows_in[col * (int64_t)m + i];
                if (r < 0 || r >= p) { atomicOr(bad_flag, 1); continue; }
                const uint32_t old = atomicOr(&bits[r >> 5], 1u << (r & 31));
                if ((old >> (r & 31)) & 1u) atomicOr(bad_flag, 1);          // repeated row
            }
            __syncthreads();
        } else {
            int draw = tid;
            uint32_t attempt = 0;
            bool pending = draw < m;
            while (__syncthreads_or(pending)) {
                int r = -1;
                if (pending) {
                    r = (int)__umulhi(skm_philox_draw(seed, (uint64_t)(col0 + col), (uint32_t)draw, attempt), (uint32_t)p);
                    if ((bits[r >> 5] >> (r & 31)) & 1u) { r = -1; ++attempt; }
                    else atomicMin(&owner[r], draw);
                }
                __syncthreads();
                if (pending && r >= 0) {
                    if (owner[r] == draw) { atomicOr(&bits[r >> 5], 1u << (r & 31)); draw += T; attempt = 0; pending = draw < m; }
                    else ++attempt;
                }
            }
        }
        if (tid == 0) {
            int run = 0;
            for (int w = 0; w < nwords; ++w) { scan[w] = run; run += __popc(bits[w]); }
        }
        __syncthreads();
        const float *yc = y ? y + col * (int64_t)p : nullptr;
        for (int w = tid; w < nwords; w += T) {
            uint32_t b = bits[w];
            int64_t out = col * (int64_t)m + scan[w];
            while (b) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                const int r = (w << 5) + bit;
                if (rows_out) rows_out[out] = r;
                if (yc) { rowidx[out] = r; val[out] = __fdiv_rn(yc[r], level); }
                ++out;
            }
        }
        if (tid == 0 && colptr) { colptr[col] = col * (int64_t)m; if (col == n - 1) colptr[n] = n * (int64_t)m; }
        __syncthreads();
    }
}

int launch_sample_gather(skm_ctx *ctx, int64_t p, int64_t n, int64_t m, const float *y, const int32_t *rows_in,
                         uint64_t seed, int64_t col0, int64_t *colptr, int32_t *rowidx, float *val, int32_t *rows_out, int *bad)
{
    if (n == 0) return SKM_OK;
    const size_t smem = (size_t)p * 4 + (size_t)((p + 31) / 32) * 8 + 16;
    if (smem > (size_t)ctx->smem_optin) { skm_set_error("row sampler: p=%lld does not fit in shared memory", (long long)p); return SKM_ERR_UNSUPPORTED; }
    int threads = 32;
    while (threads < 256 && threads < m) threads <<= 1;
    SKM_CUDA(cudaFuncSetAttribute(k_sample_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sample_gather, threads, smem));
    int64_t blocks = (int64_t)ctx->sm_count * (per_sm < 1 ? 1 : per_sm);
    if (blocks > n) blocks = n;
    k_sample_gather<<<(unsigned)blocks, threads, smem, ctx->stream>>>((int)p, n, (int)m, y, rows_in, seed, col0, colptr, rowidx, val, rows_out, bad);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// Y (p x n) = M (p x p) X (p x n), all column-major fp64: 32 x 32 output tile per CTA, one output per thread
__global__ void __launch_bounds__(1024) k_dgemm_tile(int64_t p, int64_t n, const double *__restrict__ M,
                                                     const double *__restrict__ X, double *__restrict__ Y)
{
    __shared__ double Ms[32][33], Xs[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t k0 = (int64_t)blockIdx.x * 32, j0 = (int64_t)blockIdx.y * 32;
    double acc = 0.0;
    for (int64_t i0 = 0; i0 < p; i0 += 32) {
        // Ms[ii][kk] = M[k0+kk][i0+ii] (coalesced along k), Xs[jj][ii] = X[i0+ii][j0+jj] (coalesced along i)
        Ms[ty][tx] = (k0 + tx < p && i0 + ty < p) ? M[(i0 + ty) * p + k0 + tx] : 0.0;
        Xs[ty][tx] = (i0 + tx < p && j0 + ty < n) ? X[(j0 + ty) * p + i0 + tx] : 0.0;
        __syncthreads();
#pragma unroll 8
        for (int ii = 0; ii < 32; ++ii) acc = fma(Ms[ii][tx], Xs[ty][ii], acc);
        __syncthreads();
    }
    if (k0 + tx < p && j0 + ty < n) Y[(j0 + ty) * p + k0 + tx] = acc;
}

// tf32 (10 explicit mantissa bits) rounding of a float on the host
float tf32_round_host(float f)
{
    uint32_t b;
    memcpy(&b, &f, 4);
    if ((b & 0x7f800000u) != 0x7f800000u) b = (b + 0x00000fffu + ((b >> 13) & 1u)) & 0xffffe000u;
    memcpy(&f, &b, 4);
    return f;
}

}  // namespace

// y = dct(D .* x) (inverse == 0) or y = D .* idct(x) (inverse != 0) for a dense p x n HOST matrix,
// fp64 (k_dgemm_tile).  signs may be NULL.
extern "C" int skm_dct_mix(skm_ctx *ctx, int64_t p, int64_t n, const double *x, const double *signs, int inverse, double *y)
{
    SKM_REQUIRE(ctx && (n == 0 || (x && y)), "NULL argument");
    SKM_CUDA(cudaSetDevice(ctx->device));
    SKM_REQUIRE(p >= 1 && p <= 46340 && n >= 0 && n <= 2147483647LL, "bad dimensions");
    if (n == 0) return SKM_OK;
    std::vector<double> M;
    dct_matrix(p, signs, 1.0, inverse != 0, M);
    DevBuf dM, dX, dY;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(256LL << 20) / (8 * p)));
    SKM_TRY(dM.alloc(sizeof(double) * p * p));
    SKM_TRY(dX.alloc(sizeof(double) * p * chunk));
    SKM_TRY(dY.alloc(sizeof(double) * p * chunk));
    SKM_CUDA(cudaMemcpyAsync(dM.ptr, M.data(), sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    for (int64_t j0 = 0; j0 < n; j0 += chunk) {
        const int64_t nc = std::min(chunk, n - j0);
        SKM_CUDA(cudaMemcpyAsync(dX.ptr, x + j0 * p, sizeof(double) * p * nc, cudaMemcpyHostToDevice, ctx->stream));
        for (int64_t c0 = 0; c0 < nc; c0 += 65535 * 32) {            // grid.y limit
            const int64_t cc = std::min<int64_t>(nc - c0, 65535 * 32);
            k_dgemm_tile<<<dim3((unsigned)((p + 31) / 32), (unsigned)((cc + 31) / 32)), 1024, 0, ctx->stream>>>(
                p, cc, dM.as<double>(), dX.as<double>() + c0 * p, dY.as<double>() + c0 * p);
            SKM_CHECK_LAUNCH(ctx);
        }
        SKM_CUDA(cudaMemcpyAsync(y + j0 * p, dY.ptr, sizeof(double) * p * nc, cudaMemcpyDeviceToHost, ctx->stream));
        SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return SKM_OK;
}

// The precondition + sample stage of kmeans_sparsified.m:286-334 with the DCT sketch, from a dense
// HOST matrix (p x n column-major, points are columns) to a resident SKM_F32 dataset: column chunks
// cross PCIe on a copy stream, Y = (T diag(d) (1+2eps)) X in fp32 (one GEMM per chunk), then the
// fused row sampler + gather writes the CSC arrays.  rows_host (may be NULL): int32[m*n] explicit rows.
extern "C" int skm_dataset_from_dense_host_dct(skm_ctx *ctx, int64_t p, int64_t n, const void *x, int x_type,
                                               const double *signs, int64_t m, uint64_t seed, int64_t col0,
                                               const int32_t *rows_host, int64_t chunk_cols, skm_dataset **out)
{
    SKM_REQUIRE(ctx && out && (x || n == 0), "NULL argument");
    *out = nullptr;
    SKM_CUDA(cudaSetDevice(ctx->device));
    SKM_REQUIRE(x_type == SKM_F32 || x_type == SKM_F64, "x_type must be SKM_F32 or SKM_F64");
    SKM_REQUIRE(p >= 1 && p <= 32768 && n >= 0, "need 1 <= p <= 32768");
    SKM_REQUIRE(m >= 1 && m <= p, "need 1 <= m <= p");
    const size_t xs = x_type == SKM_F32 ? 4 : 8;
    if (chunk_cols <= 0) chunk_cols = std::max<int64_t>(1, (int64_t)(256LL << 20) / (int64_t)(p * 4));
    chunk_cols = std::min<int64_t>(chunk_cols, std::max<int64_t>(n, 1));
    chunk_cols = std::min<int64_t>(chunk_cols, 2147483647LL / std::max<int64_t>(p, 1));

    std::vector<double> M;
    dct_matrix(p, signs, 1.0 + 2.0 * 2.220446049250313e-16, false, M);
    // tf32-exact halves of M, row k (output row) contiguous over the input index, rows padded to a multiple of 4
    // floats (TMA needs 16-byte row strides): M = M_hi + M_lo to ~22 bits
    const int64_t p_pad = (p + 3) & ~(int64_t)3;
    std::vector<float> Mh((size_t)p * p_pad, 0.f), Ml((size_t)p * p_pad, 0.f);
    for (int64_t i = 0; i < p; ++i)
        for (int64_t k = 0; k < p; ++k) {
            const double v = M[(size_t)i * p + k];
            const float h = tf32_round_host((float)v);
            Mh[(size_t)k * p_pad + i] = h;
            Ml[(size_t)k * p_pad + i] = tf32_round_host((float)(v - (double)h));
        }

    int64_t *colptr = nullptr; int32_t *rowidx = nullptr; float *val = nullptr;
    DevBuf dMh, dMl, raw[2], xhi, xlo, y32, drows, bad;
    int rc = SKM_OK;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    do {
        if (cudaMalloc((void **)&colptr, sizeof(int64_t) * (n + 1)) != cudaSuccess ||
            cudaMalloc((void **)&rowidx, sizeof(int32_t) * std::max<int64_t>(n * m, 1)) != cudaSuccess ||
            cudaMalloc((void **)&val, sizeof(float) * std::max<int64_t>(n * m, 1)) != cudaSuccess) {
            skm_set_error("out of device memory for the sampled matrix"); cudaGetLastError(); rc = SKM_ERR_NOMEM; break;
        }
        if ((rc = dMh.alloc(sizeof(float) * p * p_pad)) || (rc = dMl.alloc(sizeof(float) * p * p_pad)) ||
            (rc = raw[0].alloc(xs * p * chunk_cols)) || (rc = raw[1].alloc(xs * p * chunk_cols)) ||
            (rc = xhi.alloc(sizeof(float) * p_pad * chunk_cols)) || (rc = xlo.alloc(sizeof(float) * p_pad * chunk_cols)) ||
            (rc = y32.alloc(sizeof(float) * p * chunk_cols)) || (rc = bad.alloc(sizeof(int)))) break;
        if (rows_host && (rc = drows.alloc(sizeof(int32_t) * m * chunk_cols))) break;
        cudaMemcpyAsync(dMh.ptr, Mh.data(), sizeof(float) * p * p_pad, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(dMl.ptr, Ml.data(), sizeof(float) * p * p_pad, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemsetAsync(bad.ptr, 0, sizeof(int), ctx->stream);
        cudaMemsetAsync(colptr, 0, sizeof(int64_t), ctx->stream);
        if (cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking) != cudaSuccess) { skm_set_error("cudaStreamCreate failed"); rc = SKM_ERR_CUDA; break; }
        for (int i = 0; i < 2; ++i) { cudaEventCreateWithFlags(&up[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&freed[i], cudaEventDisableTiming); }
        const int64_t nchunks = n > 0 ? (n + chunk_cols - 1) / chunk_cols : 0;
        auto issue = [&](int64_t c) {
            const int64_t j0 = c * chunk_cols, nc = std::min(chunk_cols, n - j0);
            if (c >= 2) cudaStreamWaitEvent(copy_stream, freed[c & 1], 0);
            cudaMemcpyAsync(raw[c & 1].ptr, (const char *)x + (size_t)j0 * p * xs, xs * p * nc, cudaMemcpyHostToDevice, copy_stream);
            cudaEventRecord(up[c & 1], copy_stream);
        };
        if (nchunks > 0) issue(0);
        for (int64_t c = 0; c < nchunks && rc == SKM_OK; ++c) {
            if (c + 1 < nchunks) issue(c + 1);
            const int64_t j0 = c * chunk_cols, nc = std::min(chunk_cols, n - j0);
            cudaStreamWaitEvent(ctx->stream, up[c & 1], 0);
            {
                SkmTimed t(ctx, SKM_T_FWHT);
                // fl32(x) = x_hi + x_lo (tf32-exact halves), then Y = M_hi X_hi + M_hi X_lo + M_lo X_hi on the tensor cores
                if ((rc = skm_launch_cast_split(ctx, nc, p, p_pad, raw[c & 1].ptr, x_type, 1.0, xhi.as<float>(), xlo.as<float>()))) break;
                cudaEventRecord(freed[c & 1], ctx->stream);
                if ((rc = skm_launch_tc_dct(ctx, p, p_pad, nc, xhi.as<float>(), xlo.as<float>(), dMh.as<float>(), dMl.as<float>(),
                                            y32.as<float>(), p))) break;
                if (rows_host) cudaMemcpyAsync(drows.ptr, rows_host + j0 * m, sizeof(int32_t) * m * nc, cudaMemcpyHostToDevice, ctx->stream);
                rc = launch_sample_gather(ctx, p, nc, m, y32.as<float>(), rows_host ? drows.as<int32_t>() : nullptr, seed, col0 + j0,
                                          nullptr, rowidx + j0 * m, val + j0 * m, nullptr, bad.as<int>());
            }
        }
    } while (0);
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (copy_stream) {
        cudaStreamSynchronize(copy_stream);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(up[i]); cudaEventDestroy(freed[i]); }
        cudaStreamDestroy(copy_stream);
    }
    if (rc == SKM_OK && e != cudaSuccess) { skm_set_error("DCT pipeline failed: %s", cudaGetErrorString(e)); rc = SKM_ERR_CUDA; }
    if (rc == SKM_OK) {
        int hb = 0;
        cudaMemcpy(&hb, bad.ptr, sizeof(int), cudaMemcpyDeviceToHost);
        if (hb) { skm_set_error("a sampled row index is outside [0, p) or repeats within a column"); rc = SKM_ERR_INVALID; }
    }
    if (rc == SKM_OK) {
        std::vector<int64_t> hc(n + 1);
        for (int64_t j = 0; j <= n; ++j) hc[j] = j * m;
        if (cudaMemcpy(colptr, hc.data(), sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice) != cudaSuccess) { skm_set_error("colptr upload failed"); rc = SKM_ERR_CUDA; }
    }
    if (rc == SKM_OK)
        rc = skm_dataset_create_csc(ctx, p, n, colptr, SKM_I64, rowidx, SKM_I32, val, SKM_F32, SKM_F32, 1, out);
    cudaFree(colptr); cudaFree(rowidx); cudaFree(val);
    return rc;
}

// The row sets skm_dataset_from_dense_host_dct draws for (seed, col0): rows_dev int32[m*n], ascending per column.
extern "C" int skm_sample_rows_general(skm_ctx *ctx, int64_t p, int64_t n, int64_t m, uint64_t seed, int64_t col0, int32_t *rows_dev)
{
    SKM_REQUIRE(ctx && (rows_dev || n == 0), "NULL argument");
    SKM_CUDA(cudaSetDevice(ctx->device));
    SKM_REQUIRE(p >= 1 && p <= 32768 && m >= 1 && m <= p && n >= 0, "need 1 <= m <= p <= 32768");
    SKM_TRY(launch_sample_gather(ctx, p, n, m, nullptr, nullptr, seed, col0, nullptr, nullptr, nullptr, rows_dev, nullptr));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    return SKM_OK;
}
