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
// on the host in fp64): a plain library GEMM (cuBLAS, loaded lazily with dlopen so nothing else in
// the library depends on it), fp32 for the data pipeline, fp64 for the small centre matrices.
// The hand-written part is what follows the product: the fixed-count row sampler for arbitrary p
// (Philox, same contract as the Hadamard path) fused with the gather that writes CSC directly, so
// the mixed dense chunk never leaves the device.
#include "common.cuh"
#include "philox.cuh"
#include <algorithm>
#include <dlfcn.h>
#include <math.h>
#include <string.h>
#include <vector>

namespace {

// ---- minimal cuBLAS binding (v2 API) ----------------------------------------------------------
typedef void *blas_handle;
typedef int (*fn_create)(blas_handle *);
typedef int (*fn_destroy)(blas_handle);
typedef int (*fn_set_stream)(blas_handle, cudaStream_t);
typedef int (*fn_sgemm)(blas_handle, int, int, int, int, int, const float *, const float *, int, const float *, int,
                        const float *, float *, int);
typedef int (*fn_dgemm)(blas_handle, int, int, int, int, int, const double *, const double *, int, const double *, int,
                        const double *, double *, int);
struct Blas {
    void *lib = nullptr;
    blas_handle h = nullptr;
    fn_create create = nullptr; fn_destroy destroy = nullptr; fn_set_stream set_stream = nullptr;
    fn_sgemm sgemm = nullptr; fn_dgemm dgemm = nullptr;
};

void blas_free(void *p)
{
    Blas *b = (Blas *)p;
    if (!b) return;
    if (b->h && b->destroy) b->destroy(b->h);
    delete b;                                             // the library handle stays loaded for the process
}

int blas_get(skm_ctx *ctx, Blas **out)
{
    if (ctx->blas) { *out = (Blas *)ctx->blas; return SKM_OK; }
    Blas *b = new Blas();
    const char *names[] = {"libcublas.so.12", "/usr/local/cuda/lib64/libcublas.so.12", "libcublas.so"};
    for (const char *nm : names) { b->lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (b->lib) break; }
    if (!b->lib) { delete b; skm_set_error("the DCT sketch needs cuBLAS (libcublas.so.12 not found: %s)", dlerror()); return SKM_ERR_UNSUPPORTED; }
    b->create = (fn_create)dlsym(b->lib, "cublasCreate_v2");
    b->destroy = (fn_destroy)dlsym(b->lib, "cublasDestroy_v2");
    b->set_stream = (fn_set_stream)dlsym(b->lib, "cublasSetStream_v2");
    b->sgemm = (fn_sgemm)dlsym(b->lib, "cublasSgemm_v2");
    b->dgemm = (fn_dgemm)dlsym(b->lib, "cublasDgemm_v2");
    if (!b->create || !b->destroy || !b->set_stream || !b->sgemm || !b->dgemm) { delete b; skm_set_error("cuBLAS symbols missing"); return SKM_ERR_UNSUPPORTED; }
    if (b->create(&b->h) != 0) { delete b; skm_set_error("cublasCreate failed"); return SKM_ERR_CUDA; }
    if (b->set_stream(b->h, ctx->stream) != 0) { blas_free(b); skm_set_error("cublasSetStream failed"); return SKM_ERR_CUDA; }
    ctx->blas = b;
    ctx->blas_free = blas_free;
    *out = b;
    return SKM_OK;
}

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

template <typename T>
__global__ void k_cast_f32(int64_t count, const T *__restrict__ x, float *__restrict__ y)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) y[i] = (float)x[i];
}

}  // namespace

// y = dct(D .* x) (inverse == 0) or y = D .* idct(x) (inverse != 0) for a dense p x n HOST matrix,
// fp64 (cublasDgemm).  signs may be NULL.
extern "C" int skm_dct_mix(skm_ctx *ctx, int64_t p, int64_t n, const double *x, const double *signs, int inverse, double *y)
{
    SKM_REQUIRE(ctx && (n == 0 || (x && y)), "NULL argument");
    SKM_CUDA(cudaSetDevice(ctx->device));
    SKM_REQUIRE(p >= 1 && p <= 46340 && n >= 0 && n <= 2147483647LL, "bad dimensions");
    if (n == 0) return SKM_OK;
    Blas *b;
    SKM_TRY(blas_get(ctx, &b));
    std::vector<double> M;
    dct_matrix(p, signs, 1.0, inverse != 0, M);
    DevBuf dM, dX, dY;
    const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(n, (int64_t)(256LL << 20) / (8 * p)));
    SKM_TRY(dM.alloc(sizeof(double) * p * p));
    SKM_TRY(dX.alloc(sizeof(double) * p * chunk));
    SKM_TRY(dY.alloc(sizeof(double) * p * chunk));
    SKM_CUDA(cudaMemcpyAsync(dM.ptr, M.data(), sizeof(double) * p * p, cudaMemcpyHostToDevice, ctx->stream));
    const double one = 1.0, zero = 0.0;
    for (int64_t j0 = 0; j0 < n; j0 += chunk) {
        const int64_t nc = std::min(chunk, n - j0);
        SKM_CUDA(cudaMemcpyAsync(dX.ptr, x + j0 * p, sizeof(double) * p * nc, cudaMemcpyHostToDevice, ctx->stream));
        if (b->dgemm(b->h, 0, 0, (int)p, (int)nc, (int)p, &one, dM.as<double>(), (int)p, dX.as<double>(), (int)p, &zero,
                     dY.as<double>(), (int)p) != 0) { skm_set_error("cublasDgemm failed"); return SKM_ERR_CUDA; }
        ctx->launches++;
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
    Blas *b;
    SKM_TRY(blas_get(ctx, &b));
    const size_t xs = x_type == SKM_F32 ? 4 : 8;
    if (chunk_cols <= 0) chunk_cols = std::max<int64_t>(1, (int64_t)(256LL << 20) / (int64_t)(p * 4));
    chunk_cols = std::min<int64_t>(chunk_cols, std::max<int64_t>(n, 1));
    chunk_cols = std::min<int64_t>(chunk_cols, 2147483647LL / std::max<int64_t>(p, 1));

    std::vector<double> M;
    dct_matrix(p, signs, 1.0 + 2.0 * 2.220446049250313e-16, false, M);
    std::vector<float> M32((size_t)p * p);
    for (size_t i = 0; i < M32.size(); ++i) M32[i] = (float)M[i];

    int64_t *colptr = nullptr; int32_t *rowidx = nullptr; float *val = nullptr;
    DevBuf dM, raw[2], x32, y32, drows, bad;
    int rc = SKM_OK;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t up[2] = {nullptr, nullptr}, freed[2] = {nullptr, nullptr};
    do {
        if (cudaMalloc((void **)&colptr, sizeof(int64_t) * (n + 1)) != cudaSuccess ||
            cudaMalloc((void **)&rowidx, sizeof(int32_t) * std::max<int64_t>(n * m, 1)) != cudaSuccess ||
            cudaMalloc((void **)&val, sizeof(float) * std::max<int64_t>(n * m, 1)) != cudaSuccess) {
            skm_set_error("out of device memory for the sampled matrix"); cudaGetLastError(); rc = SKM_ERR_NOMEM; break;
        }
        if ((rc = dM.alloc(sizeof(float) * p * p)) || (rc = raw[0].alloc(xs * p * chunk_cols)) || (rc = raw[1].alloc(xs * p * chunk_cols)) ||
            (rc = y32.alloc(sizeof(float) * p * chunk_cols)) || (rc = bad.alloc(sizeof(int)))) break;
        if (x_type == SKM_F64 && (rc = x32.alloc(sizeof(float) * p * chunk_cols))) break;
        if (rows_host && (rc = drows.alloc(sizeof(int32_t) * m * chunk_cols))) break;
        cudaMemcpyAsync(dM.ptr, M32.data(), sizeof(float) * p * p, cudaMemcpyHostToDevice, ctx->stream);
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
        const float one = 1.f, zero = 0.f;
        for (int64_t c = 0; c < nchunks && rc == SKM_OK; ++c) {
            if (c + 1 < nchunks) issue(c + 1);
            const int64_t j0 = c * chunk_cols, nc = std::min(chunk_cols, n - j0);
            cudaStreamWaitEvent(ctx->stream, up[c & 1], 0);
            const float *xin = (const float *)raw[c & 1].ptr;
            if (x_type == SKM_F64) {
                const int64_t blocks = std::min<int64_t>((p * nc + 255) / 256, (int64_t)ctx->sm_count * 32);
                k_cast_f32<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(p * nc, (const double *)raw[c & 1].ptr, x32.as<float>());
                ctx->launches++;
                xin = x32.as<float>();
            }
            {
                SkmTimed t(ctx, SKM_T_FWHT);
                if (b->sgemm(b->h, 0, 0, (int)p, (int)nc, (int)p, &one, dM.as<float>(), (int)p, xin, (int)p, &zero, y32.as<float>(), (int)p) != 0) {
                    skm_set_error("cublasSgemm failed"); rc = SKM_ERR_CUDA; break;
                }
                ctx->launches++;
                cudaEventRecord(freed[c & 1], ctx->stream);
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
