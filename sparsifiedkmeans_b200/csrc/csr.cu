// csr.cu -- upload-time transposition of the sparsified matrix into a row-major image for K2.
//
// K2 needs, for every row r and cluster k, the sum of X(r, j) over the columns j assigned to
// k (kmeans_sparsified.m:447-448).  Scattering from the column-major stream costs two L2
// atomics per stored entry (the first version of K2: 12.6 ms per iteration at config 2).
// Walking the matrix by ROWS instead lets a warp keep lane-private bins for one row in
// shared memory and touch global memory once per (row, cluster); the only gather left is the
// 1-byte assignment of each entry's column.
//
// Layout: csr[rowptr[r] .. rowptr[r+1]) holds (column, value-bits) pairs of row r, ordered by
// column tile (tiles of CSR_TILE = 512 consecutive columns; order inside a tile is unspecified).
#include "common.cuh"
#include <cub/device/device_scan.cuh>
#include <chrono>
#include <string.h>
#include <vector>

// small tiles keep a row's entries nearly column-sorted, so the 32 assignment gathers of a
// warp in K2 touch a handful of cache lines instead of up to 32
#define CSR_TILE 512

namespace {

// counts[cell_base + row * ntc + (tile - tile0)] = entries of `row` in column tile `tile`, tiles [tile0, tile1) of one chunk
__global__ void k_csr_count(int64_t p, int64_t n, int64_t tile0, int64_t tile1, int64_t cell_base, const int64_t *__restrict__ colptr,
                            const int32_t *__restrict__ rowidx, int64_t *__restrict__ counts)
{
    extern __shared__ int hist[];
    const int64_t ntc = tile1 - tile0;
    for (int64_t tile = tile0 + blockIdx.x; tile < tile1; tile += gridDim.x) {
        for (int i = threadIdx.x; i < p; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const int64_t j0 = tile * CSR_TILE, j1 = min(n, j0 + CSR_TILE);
        const int64_t t0 = colptr[j0], t1 = colptr[j1];
        for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) {
            const int r = rowidx[t];
            if ((unsigned)r < (unsigned)p) atomicAdd(&hist[r], 1);      // an invalid row is reported by the validation pass
        }
        __syncthreads();
        for (int i = threadIdx.x; i < p; i += blockDim.x) counts[cell_base + (int64_t)i * ntc + (tile - tile0)] = hist[i];
        __syncthreads();
    }
}

template <typename VT>
__global__ void k_csr_scatter(int64_t p, int64_t n, int64_t tile0, int64_t tile1, int64_t cell_base, int64_t ebase,
                              const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                              const int64_t *__restrict__ offsets, int2 *__restrict__ csr)
{
    // per row: the tile's base offset (64-bit) with the cursor in the same word, so one shared-memory atomic
    // returns the final position (the first version chained shared atomic -> global load of the offset -> store)
    extern __shared__ unsigned long long cur[];
    const int64_t ntc = tile1 - tile0;
    for (int64_t tile = tile0 + blockIdx.x; tile < tile1; tile += gridDim.x) {
        for (int i = threadIdx.x; i < p; i += blockDim.x)
            cur[i] = (unsigned long long)(ebase + offsets[cell_base + (int64_t)i * ntc + (tile - tile0)]);
        __syncthreads();
        const int64_t j0 = tile * CSR_TILE, j1 = min(n, j0 + CSR_TILE);
        // one warp per column keeps the column index available without a search; two columns per trip for
        // independent work in flight
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int64_t j = j0 + warp; j < j1; j += 2 * nwarps) {
            const int64_t jb = j + nwarps;
            const int64_t t0 = colptr[j], t1 = colptr[j + 1];
            int64_t u0 = 0, u1 = 0;
            if (jb < j1) { u0 = colptr[jb]; u1 = colptr[jb + 1]; }
            int64_t t = t0 + lane, u = u0 + lane;
            while (t < t1 || u < u1) {
                int ra = -1, rb = -1, xa = 0, xb = 0;
                if (t < t1) { ra = rowidx[t]; xa = __float_as_int((float)val[t]); }
                if (u < u1) { rb = rowidx[u]; xb = __float_as_int((float)val[u]); }
                // (an out-of-range row is reported by the validation pass; it must not write anywhere here)
                if ((unsigned)ra < (unsigned)p) csr[atomicAdd(&cur[ra], 1ULL)] = make_int2((int)j, xa);
                if ((unsigned)rb < (unsigned)p) csr[atomicAdd(&cur[rb], 1ULL)] = make_int2((int)jb, xb);
                t += 32; u += 32;
            }
        }
        __syncthreads();
    }
}

// fallback for very long columns' worth of rows (p * 8 bytes would not fit in shared memory): 32-bit cursors and the
// tile offsets read from global memory per entry
template <typename VT>
__global__ void k_csr_scatter_big(int64_t p, int64_t n, int64_t tile0, int64_t tile1, int64_t cell_base, int64_t ebase,
                                  const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                                  const int64_t *__restrict__ offsets, int2 *__restrict__ csr)
{
    extern __shared__ int hist[];
    const int64_t ntc = tile1 - tile0;
    for (int64_t tile = tile0 + blockIdx.x; tile < tile1; tile += gridDim.x) {
        for (int i = threadIdx.x; i < p; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const int64_t j0 = tile * CSR_TILE, j1 = min(n, j0 + CSR_TILE);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int64_t j = j0 + warp; j < j1; j += nwarps) {
            const int64_t t0 = colptr[j], t1 = colptr[j + 1];
            for (int64_t t = t0 + lane; t < t1; t += 32) {
                const int r = rowidx[t];
                if ((unsigned)r >= (unsigned)p) continue;
                const int slot = atomicAdd(&hist[r], 1);
                csr[ebase + offsets[cell_base + (int64_t)r * ntc + (tile - tile0)] + slot] = make_int2((int)j, __float_as_int((float)val[t]));
            }
        }
        __syncthreads();
    }
}

// rowptr[r] = first entry of row r inside this chunk's region (r < p), rowptr[p] = end of the region
__global__ void k_csr_rowptr(int64_t p, int64_t ntc, int64_t cell_base, int64_t ebase, int64_t eend,
                             const int64_t *__restrict__ offsets, int64_t *__restrict__ rowptr)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < p) rowptr[r] = ebase + offsets[cell_base + r * ntc];
    if (r == p) rowptr[p] = eend;
}

double csr_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

}  // namespace

// The image is built chunk by chunk (a chunk = a run of whole 512-column tiles): the region of a chunk starts at the
// chunk's first stored entry (colptr[j0]) and holds, row after row, the chunk's entries ordered by tile -- so a chunk can
// be counted, scanned and scattered as soon as its columns are on the device, while the next one is still crossing
// PCIe.  K2 only sees work units (row, start, end), so it does not care that a row's entries sit in one run per chunk.
void skm_csr_abort(SkmCsrBuild *b)
{
    if (!b) return;
    cudaFree(b->counts); cudaFree(b->offsets); cudaFree(b->scan_tmp);
    b->counts = b->offsets = b->scan_tmp = nullptr; b->active = false;
}

// step 1: decide whether the row-major image is built at all, allocate it and the per-chunk scratch
int skm_csr_begin(skm_dataset *ds, SkmCsrBuild *b, int64_t max_chunk_cols, int64_t nchunks)
{
    skm_ctx *ctx = ds->ctx;
    const int64_t p = ds->p, n = ds->n, nnz = ds->nnz;
    memset(b, 0, sizeof *b);
    ds->csr = nullptr;
    ds->rowptr = nullptr;
    ds->h_rowptr = nullptr;
    ds->unit_row = nullptr;
    ds->unit_start = nullptr;
    ds->unit_counter = nullptr;
    ds->nunits = 0;
    if (n == 0 || p == 0 || nnz == 0 || ds->store_dtype != SKM_F32) return SKM_OK;
    if (n >= 2147483647LL) { skm_set_error("a shard may hold at most 2^31-1 columns"); return SKM_ERR_UNSUPPORTED; }
    const size_t smem = (size_t)p * sizeof(int);
    if (smem > (size_t)ctx->smem_optin) return SKM_OK;          // K2 falls back to the atomic kernel
    b->max_tiles = (max_chunk_cols + CSR_TILE - 1) / CSR_TILE;
    b->nchunks = nchunks;
    const int64_t cells = p * b->max_tiles;
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const int64_t *)nullptr, (int64_t *)nullptr, cells);
    b->scan_bytes = tmp_bytes;
    cudaError_t e = cudaMalloc(&b->counts, sizeof(int64_t) * cells);
    if (e == cudaSuccess) e = cudaMalloc(&b->offsets, sizeof(int64_t) * cells);
    if (e == cudaSuccess) e = cudaMalloc(&b->scan_tmp, tmp_bytes ? tmp_bytes : 16);
    void *d = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&d, sizeof(int64_t) * (size_t)(p + 1) * (size_t)nchunks);
    if (e != cudaSuccess) { cudaGetLastError(); skm_csr_abort(b); skm_set_error("cudaMalloc(csr scratch) failed: %s", cudaGetErrorString(e)); return SKM_ERR_NOMEM; }
    ds->rowptr = (int64_t *)d;
    const int rc = skm_big_alloc(ctx, &d, sizeof(int2) * (size_t)(nnz > 0 ? nnz : 1), "csr");
    if (rc != SKM_OK) { skm_csr_abort(b); return rc; }
    ds->csr = (int2 *)d;
    ds->device_bytes += (int64_t)sizeof(int2) * nnz + (int64_t)sizeof(int64_t) * (p + 1) * nchunks;
    SKM_CUDA(cudaFuncSetAttribute(k_csr_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    b->active = true;
    return SKM_OK;
}

// step 2: chunk `c` = columns [j0, j1) (j0 a multiple of 512), stored entries [e0, e1): count, scan, scatter, row pointers
int skm_csr_chunk(skm_dataset *ds, SkmCsrBuild *b, int64_t c, int64_t j0, int64_t j1, int64_t e0, int64_t e1)
{
    if (!b->active || j1 <= j0) return SKM_OK;
    skm_ctx *ctx = ds->ctx;
    const int64_t p = ds->p;
    const int64_t tile0 = j0 / CSR_TILE, tile1 = (j1 + CSR_TILE - 1) / CSR_TILE, nt = tile1 - tile0;
    if (nt > b->max_tiles || c >= b->nchunks) { skm_set_error("csr chunk larger than announced"); return SKM_ERR_INVALID; }
    const int64_t cells = p * nt;
    const int64_t blocks = nt < (int64_t)ctx->sm_count * 4 ? nt : (int64_t)ctx->sm_count * 4;
    int64_t *counts = (int64_t *)b->counts, *offsets = (int64_t *)b->offsets;
    k_csr_count<<<(unsigned)blocks, 512, (size_t)p * sizeof(int), ctx->stream>>>(p, ds->n, tile0, tile1, 0, ds->colptr, ds->rowidx, counts);
    SKM_CHECK_LAUNCH(ctx);
    size_t tb = b->scan_bytes;
    SKM_CUDA(cub::DeviceScan::ExclusiveSum(b->scan_tmp, tb, counts, offsets, cells, ctx->stream));
    ctx->launches++;
    k_csr_rowptr<<<(unsigned)((p + 1 + 255) / 256), 256, 0, ctx->stream>>>(p, nt, 0, e0, e1, offsets, ds->rowptr + c * (p + 1));
    SKM_CHECK_LAUNCH(ctx);
    const size_t smem = (size_t)p * sizeof(int), smem_sc = (size_t)p * sizeof(unsigned long long);
    if (smem_sc <= (size_t)ctx->smem_optin) {
        SKM_CUDA(cudaFuncSetAttribute(k_csr_scatter<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sc));
        k_csr_scatter<float><<<(unsigned)blocks, 512, smem_sc, ctx->stream>>>(p, ds->n, tile0, tile1, 0, e0, ds->colptr, ds->rowidx,
                                                                          (const float *)ds->val, offsets, ds->csr);
    } else {
        SKM_CUDA(cudaFuncSetAttribute(k_csr_scatter_big<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_csr_scatter_big<float><<<(unsigned)blocks, 512, smem, ctx->stream>>>(p, ds->n, tile0, tile1, 0, e0, ds->colptr, ds->rowidx,
                                                                           (const float *)ds->val, offsets, ds->csr);
    }
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_build_csr(skm_dataset *ds)
{
    SkmCsrBuild b;
    SKM_TRY(skm_csr_begin(ds, &b, ds->n, 1));
    if (!b.active) return SKM_OK;
    int rc = skm_csr_chunk(ds, &b, 0, 0, ds->n, 0, ds->nnz);
    if (rc == SKM_OK) rc = skm_csr_finish(ds, &b);
    else skm_csr_abort(&b);
    return rc;
}

// step 3: K2's work list from the chunks' row pointers
static int csr_finish_impl(skm_dataset *ds, SkmCsrBuild *b);
int skm_csr_finish(skm_dataset *ds, SkmCsrBuild *b)
{
    if (!b->active) return SKM_OK;
    const int rc = csr_finish_impl(ds, b);
    cudaStreamSynchronize(ds->ctx->stream);
    skm_csr_abort(b);
    return rc;
}

static int csr_finish_impl(skm_dataset *ds, SkmCsrBuild *b)
{
    skm_ctx *ctx = ds->ctx;
    const int64_t p = ds->p, nnz = ds->nnz, nch = b->nchunks;
    const double tcsr0 = csr_now();
    // host copy of the row pointers: K2's work list is built from it
    ds->h_rowptr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(p + 1) * (size_t)nch);
    if (!ds->h_rowptr) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    SKM_CUDA(cudaMemcpyAsync(ds->h_rowptr, ds->rowptr, sizeof(int64_t) * (p + 1) * nch, cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (getenv("SKM_TRACE")) fprintf(stderr, "[skm trace]   csr: drain + row pointers   %8.2f ms\n", 1e3 * (csr_now() - tcsr0));

    // work list: rows cut into chunks sized so every resident warp gets several units
    // (about 8 units per resident warp keeps the tail of the last wave short; K2 pulls units
    // from an atomic counter, so long and short rows balance out)
    int64_t chunk = nnz / ((int64_t)ctx->sm_count * 64 * 8);
    chunk = chunk < 1024 ? 1024 : (chunk > 16384 ? 16384 : chunk);
    chunk = (chunk + 31) & ~(int64_t)31;
    std::vector<int32_t> urow;
    std::vector<int64_t> ustart, uend;
    for (int64_t r = 0; r < p; ++r) {
        for (int64_t c = 0; c < nch; ++c) {
            const int64_t *rp = ds->h_rowptr + c * (p + 1);
            for (int64_t s = rp[r]; s < rp[r + 1]; s += chunk) {
                urow.push_back((int32_t)r);
                ustart.push_back(s);
                uend.push_back(s + chunk < rp[r + 1] ? s + chunk : rp[r + 1]);
            }
        }
    }
    ds->nunits = (int64_t)urow.size();
    if (ds->nunits > 0) {
        SKM_CUDA(cudaMalloc((void **)&ds->unit_counter, sizeof(unsigned long long)));
        SKM_CUDA(cudaMalloc((void **)&ds->unit_row, sizeof(int32_t) * ds->nunits));
        SKM_CUDA(cudaMalloc((void **)&ds->unit_start, sizeof(int64_t) * 2 * ds->nunits));
        SKM_CUDA(cudaMemcpyAsync(ds->unit_row, urow.data(), sizeof(int32_t) * ds->nunits, cudaMemcpyHostToDevice, ctx->stream));
        SKM_CUDA(cudaMemcpyAsync(ds->unit_start, ustart.data(), sizeof(int64_t) * ds->nunits, cudaMemcpyHostToDevice, ctx->stream));
        SKM_CUDA(cudaMemcpyAsync(ds->unit_start + ds->nunits, uend.data(), sizeof(int64_t) * ds->nunits, cudaMemcpyHostToDevice, ctx->stream));
        SKM_CUDA(cudaStreamSynchronize(ctx->stream));
        ds->device_bytes += (int64_t)ds->nunits * 20;
    }
    return SKM_OK;
}
