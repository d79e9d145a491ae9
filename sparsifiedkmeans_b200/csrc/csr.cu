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
#include <vector>

// small tiles keep a row's entries nearly column-sorted, so the 32 assignment gathers of a
// warp in K2 touch a handful of cache lines instead of up to 32
#define CSR_TILE 512

namespace {

// counts[row * ntiles + tile] = entries of `row` in column tile `tile`
__global__ void k_csr_count(int64_t p, int64_t n, int64_t ntiles, const int64_t *__restrict__ colptr,
                            const int32_t *__restrict__ rowidx, int64_t *__restrict__ counts)
{
    extern __shared__ int hist[];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < p; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const int64_t j0 = tile * CSR_TILE, j1 = min(n, j0 + CSR_TILE);
        const int64_t t0 = colptr[j0], t1 = colptr[j1];
        for (int64_t t = t0 + threadIdx.x; t < t1; t += blockDim.x) atomicAdd(&hist[rowidx[t]], 1);
        __syncthreads();
        for (int i = threadIdx.x; i < p; i += blockDim.x) counts[(int64_t)i * ntiles + tile] = hist[i];
        __syncthreads();
    }
}

template <typename VT>
__global__ void k_csr_scatter(int64_t p, int64_t n, int64_t ntiles, const int64_t *__restrict__ colptr,
                              const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                              const int64_t *__restrict__ offsets, int2 *__restrict__ csr)
{
    // per row: the tile's base offset (64-bit) with the cursor in the same word, so one shared-memory atomic
    // returns the final position (the first version chained shared atomic -> global load of the offset -> store)
    extern __shared__ unsigned long long cur[];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < p; i += blockDim.x) cur[i] = (unsigned long long)offsets[(int64_t)i * ntiles + tile];
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
                if (ra >= 0) csr[atomicAdd(&cur[ra], 1ULL)] = make_int2((int)j, xa);
                if (rb >= 0) csr[atomicAdd(&cur[rb], 1ULL)] = make_int2((int)jb, xb);
                t += 32; u += 32;
            }
        }
        __syncthreads();
    }
}

// fallback for very long columns' worth of rows (p * 8 bytes would not fit in shared memory): 32-bit cursors and the
// tile offsets read from global memory per entry
template <typename VT>
__global__ void k_csr_scatter_big(int64_t p, int64_t n, int64_t ntiles, const int64_t *__restrict__ colptr,
                                  const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                                  const int64_t *__restrict__ offsets, int2 *__restrict__ csr)
{
    extern __shared__ int hist[];
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < p; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        const int64_t j0 = tile * CSR_TILE, j1 = min(n, j0 + CSR_TILE);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
        for (int64_t j = j0 + warp; j < j1; j += nwarps) {
            const int64_t t0 = colptr[j], t1 = colptr[j + 1];
            for (int64_t t = t0 + lane; t < t1; t += 32) {
                const int r = rowidx[t];
                const int slot = atomicAdd(&hist[r], 1);
                csr[offsets[(int64_t)r * ntiles + tile] + slot] = make_int2((int)j, __float_as_int((float)val[t]));
            }
        }
        __syncthreads();
    }
}

__global__ void k_csr_rowptr(int64_t p, int64_t ntiles, int64_t nnz, const int64_t *__restrict__ offsets,
                             int64_t *__restrict__ rowptr)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < p) rowptr[r] = offsets[r * ntiles];
    if (r == p) rowptr[p] = nnz;
}

}  // namespace

int skm_build_csr(skm_dataset *ds)
{
    skm_ctx *ctx = ds->ctx;
    const int64_t p = ds->p, n = ds->n, nnz = ds->nnz;
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
    const int64_t ntiles = (n + CSR_TILE - 1) / CSR_TILE;
    const int64_t cells = p * ntiles;
    DevBuf counts, offsets, tmp;
    SKM_TRY(counts.alloc(sizeof(int64_t) * cells));
    SKM_TRY(offsets.alloc(sizeof(int64_t) * cells));
    SKM_CUDA(cudaFuncSetAttribute(k_csr_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = ntiles < (int64_t)ctx->sm_count * 4 ? ntiles : (int64_t)ctx->sm_count * 4;
    k_csr_count<<<(unsigned)blocks, 512, smem, ctx->stream>>>(p, n, ntiles, ds->colptr, ds->rowidx, counts.as<int64_t>());
    SKM_CHECK_LAUNCH(ctx);
    size_t tmp_bytes = 0;
    SKM_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, counts.as<int64_t>(), offsets.as<int64_t>(), cells, ctx->stream));
    SKM_TRY(tmp.alloc(tmp_bytes));
    SKM_CUDA(cub::DeviceScan::ExclusiveSum(tmp.ptr, tmp_bytes, counts.as<int64_t>(), offsets.as<int64_t>(), cells, ctx->stream));
    ctx->launches++;

    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(int64_t) * (p + 1));
    if (e != cudaSuccess) { skm_set_error("cudaMalloc(rowptr) failed: %s", cudaGetErrorString(e)); return SKM_ERR_NOMEM; }
    ds->rowptr = (int64_t *)d;
    e = cudaMalloc(&d, sizeof(int2) * (size_t)(nnz > 0 ? nnz : 1));
    if (e != cudaSuccess) { skm_set_error("cudaMalloc(csr, %lld entries) failed: %s", (long long)nnz, cudaGetErrorString(e)); return SKM_ERR_NOMEM; }
    ds->csr = (int2 *)d;
    ds->device_bytes += (int64_t)sizeof(int2) * nnz + (int64_t)sizeof(int64_t) * (p + 1);
    k_csr_rowptr<<<(unsigned)((p + 1 + 255) / 256), 256, 0, ctx->stream>>>(p, ntiles, nnz, offsets.as<int64_t>(), ds->rowptr);
    SKM_CHECK_LAUNCH(ctx);
    const size_t smem_sc = (size_t)p * sizeof(unsigned long long);
    if (smem_sc <= (size_t)ctx->smem_optin) {
        SKM_CUDA(cudaFuncSetAttribute(k_csr_scatter<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_sc));
        k_csr_scatter<float><<<(unsigned)blocks, 512, smem_sc, ctx->stream>>>(p, n, ntiles, ds->colptr, ds->rowidx,
                                                                          (const float *)ds->val, offsets.as<int64_t>(), ds->csr);
    } else {
        SKM_CUDA(cudaFuncSetAttribute(k_csr_scatter_big<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_csr_scatter_big<float><<<(unsigned)blocks, 512, smem, ctx->stream>>>(p, n, ntiles, ds->colptr, ds->rowidx,
                                                                           (const float *)ds->val, offsets.as<int64_t>(), ds->csr);
    }
    SKM_CHECK_LAUNCH(ctx);
    // host copy of the row pointers: K2's work list is built from it
    ds->h_rowptr = (int64_t *)malloc(sizeof(int64_t) * (p + 1));
    if (!ds->h_rowptr) { skm_set_error("out of host memory"); return SKM_ERR_NOMEM; }
    SKM_CUDA(cudaMemcpyAsync(ds->h_rowptr, ds->rowptr, sizeof(int64_t) * (p + 1), cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));

    // work list: rows cut into chunks sized so every resident warp gets several units
    // (about 8 units per resident warp keeps the tail of the last wave short; K2 pulls units
    // from an atomic counter, so long and short rows balance out)
    int64_t chunk = nnz / ((int64_t)ctx->sm_count * 64 * 8);
    chunk = chunk < 1024 ? 1024 : (chunk > 16384 ? 16384 : chunk);
    chunk = (chunk + 31) & ~(int64_t)31;
    std::vector<int32_t> urow;
    std::vector<int64_t> ustart;
    for (int64_t r = 0; r < p; ++r) {
        for (int64_t s = ds->h_rowptr[r]; s < ds->h_rowptr[r + 1]; s += chunk) {
            urow.push_back((int32_t)r);
            ustart.push_back(s);
        }
    }
    ds->nunits = (int64_t)urow.size();
    // unit u ends at min(start+chunk, end of its row): store explicit ends as ustart2
    std::vector<int64_t> uend(ds->nunits);
    for (int64_t u = 0; u < ds->nunits; ++u) {
        int64_t e2 = ustart[u] + chunk, re = ds->h_rowptr[urow[u] + 1];
        uend[u] = e2 < re ? e2 : re;
    }
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
