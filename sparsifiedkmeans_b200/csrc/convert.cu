// convert.cu -- upload-time conversions: index/value narrowing, CSC validation and the
// SELL-32 image the fast assignment kernel streams (layout described in common.cuh).
#include "common.cuh"
#include <stdlib.h>
#include <vector>
#include <cub/device/device_scan.cuh>
#include "sched16.cuh"

namespace {

template <typename S, typename D>
__global__ void k_convert(const S *__restrict__ src, D *__restrict__ dst, int64_t count)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) dst[i] = (D)src[i];
}

template <typename S, typename D>
int convert(skm_ctx *ctx, const void *src, void *dst, int64_t count)
{
    if (count == 0) return SKM_OK;
    int threads = 256;
    int64_t blocks = (count + threads - 1) / threads;
    int64_t cap = (int64_t)ctx->sm_count * 32;
    if (blocks > cap) blocks = cap;
    k_convert<S, D><<<(unsigned)blocks, threads, 0, ctx->stream>>>((const S *)src, (D *)dst, count);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// flag[0] |= 1 if colptr is not non-decreasing / does not start at 0 / does not end at nnz
// flag[0] |= 2 if a row index is outside [0,p)
// flag[1]  = max column length
__global__ void k_validate_cols(int64_t n, int64_t nnz, const int64_t *__restrict__ colptr, int *flag)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0, mx = 0;
    if (j == 0 && (colptr[0] != 0 || colptr[n] != nnz)) bad = 1;
    for (; j < n; j += stride) {
        int64_t a = colptr[j], b = colptr[j + 1];
        if (b < a || a < 0 || b > nnz) bad = 1;
        else if (b - a > mx) mx = (int)min((int64_t)INT32_MAX, b - a);
    }
    if (bad) atomicOr(&flag[0], 1);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 8));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(&flag[1], mx);
}

__global__ void k_validate_rows(int64_t p, int64_t nnz, const int32_t *__restrict__ rowidx, int *flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int bad = 0;
    for (; i < nnz; i += stride) {
        int32_t r = rowidx[i];
        if (r < 0 || (int64_t)r >= p) bad = 1;
    }
    if (bad) atomicOr(&flag[0], 2);
}

// pairs per column of each slice (max column length in the slice, rounded up to even, / 2)
__global__ void k_slice_width(int64_t n, int64_t nslices, const int64_t *__restrict__ colptr,
                              int32_t *__restrict__ width2)
{
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= nslices) return;
    int64_t j = warp * SKM_SLICE + lane;
    int len = 0;
    if (j < n) len = (int)(colptr[j + 1] - colptr[j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) width2[warp] = (len + 1) >> 1;
}

template <typename VT>
__global__ void k_fill_sell(int64_t p, int64_t n, int64_t nslices, const int64_t *__restrict__ colptr,
                            const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                            const int64_t *__restrict__ slice_ptr, int4 *__restrict__ sell)
{
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (warp >= nslices) return;
    int64_t j = warp * SKM_SLICE + lane;
    int64_t a = 0, b = 0;
    if (j < n) { a = colptr[j]; b = colptr[j + 1]; }
    int64_t base = slice_ptr[warp];
    int w2 = (int)((slice_ptr[warp + 1] - base) >> 5);
    const int pad_row = (int)p;
    for (int t2 = 0; t2 < w2; ++t2) {
        int4 q;
        int64_t t = a + 2 * (int64_t)t2;
        if (t < b) { q.x = rowidx[t]; q.y = __float_as_int((float)val[t]); }
        else       { q.x = pad_row;   q.y = 0; }
        if (t + 1 < b) { q.z = rowidx[t + 1]; q.w = __float_as_int((float)val[t + 1]); }
        else           { q.z = pad_row;       q.w = 0; }
        sell[base + (int64_t)t2 * 32 + lane] = q;
    }
}


// Bank-aware variant of k_fill_sell.  The fast kernel gathers one 16-byte chunk of a centroid
// row per lane per LDS.128; the hardware serves a quarter-warp (8 lanes) per wavefront when the
// 8 chunks fall in 8 different 16-byte bank groups.  With an odd number of chunks per table
// row the bank group of a chunk is a bijection of (row mod 8), so the order in which a column's
// entries are visited decides the conflicts (measured 2.46 wavefronts per quarter-warp in
// stored order, profiles/r1_a_first_path.md).  The order is free -- the fast kernel's guard
// bound does not depend on it and the fp64 re-evaluation reads the CSC copy -- so each slice
// is scheduled greedily here: at every step the 8 lanes of a quarter pick, in rotating order,
// the residue class (row mod 8) with the most remaining entries among the classes no lane of
// the quarter has taken yet.
template <typename VT>
__global__ void k_fill_sell_banked(int64_t p, int64_t n, int64_t nslices, int wmax,
                                   const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                   const VT *__restrict__ val, const int64_t *__restrict__ slice_ptr,
                                   int4 *__restrict__ sell)
{
    extern __shared__ unsigned short s_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (warp >= nslices) return;
    unsigned short *ord = s_all + (size_t)wib * wmax * 32;        // [t][lane]: entry ids sorted by class
    const int64_t j = warp * SKM_SLICE + lane;
    int64_t a = 0, b = 0;
    if (j < n) { a = colptr[j]; b = colptr[j + 1]; }
    const int len = (int)(b - a);
    const int64_t base = slice_ptr[warp];
    const int w2 = (int)((slice_ptr[warp + 1] - base) >> 5);
    const int pad_row = (int)p;

    // counting sort of this lane's entries by (row & 7)
    int cnt[8], start[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) cnt[c] = 0;
    for (int t = 0; t < len; ++t) {
        const int c = rowidx[a + t] & 7;
#pragma unroll
        for (int q = 0; q < 8; ++q) cnt[q] += (q == c);
    }
    int run = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) { start[c] = run; run += cnt[c]; }
    {
        int fill[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) fill[c] = start[c];
        for (int t = 0; t < len; ++t) {
            const int c = rowidx[a + t] & 7;
            int pos = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { if (q == c) { pos = fill[q]; fill[q] = pos + 1; } }
            ord[pos * 32 + lane] = (unsigned short)t;
        }
    }
    __syncwarp();

    int rem = len;
    int4 q4 = make_int4(pad_row, 0, pad_row, 0);
    const int qbase = lane & ~7, me = lane & 7;
    for (int step = 0; step < 2 * w2; ++step) {
        unsigned load = 0;                       // 4-bit load per class, shared by the quarter
        int pick = -1;
        for (int i = 0; i < 8; ++i) {
            const int chooser = (step + i) & 7;
            if (me == chooser && rem > 0) {
                int best = -2147483647 - 1, bestc = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int l = (load >> (4 * c)) & 15;
                    const int sc = (l == 0 ? (1 << 20) : 0) - l * 1024 + cnt[c];
                    if (cnt[c] > 0 && sc > best) { best = sc; bestc = c; }
                }
                int pos = 0;
#pragma unroll
                for (int c = 0; c < 8; ++c) { if (c == bestc) { pos = start[c]; start[c] = pos + 1; cnt[c] -= 1; } }
                pick = ord[pos * 32 + lane];
                rem -= 1;
                const unsigned l = (load >> (4 * bestc)) & 15;
                if (l < 15) load += 1u << (4 * bestc);
            }
            load = __shfl_sync(0xffffffffu, load, qbase | chooser);
        }
        int r = pad_row, xb = 0;
        if (pick >= 0) { r = rowidx[a + pick]; xb = __float_as_int((float)val[a + pick]); }
        if (step & 1) {
            q4.z = r; q4.w = xb;
            sell[base + (int64_t)(step >> 1) * 32 + lane] = q4;
        } else { q4.x = r; q4.y = xb; q4.z = pad_row; q4.w = 0; }
    }
}


// ---------------------------------------------------------------------------------------------
// Layout modes 1 / 2 (dual table): exact conflict-free schedule per half- / quarter-warp, see
// sched16.cuh for the algorithm (balance over the two table copies, then a Birkhoff-von Neumann
// decomposition of the padded lanes x classes matrix into permutations = steps).
//
// k_sched_dual: one THREAD per problem (the decomposition is sequential), ~1.1 KB of scratch per
// problem in shared memory, interleaved across the threads of the block.  It writes one code byte per
// (lane, step) into the head of the slice's own SELL region; k_fill_sell_sched then materialises
// the slice (warp per slice).
struct SchedMem {
    unsigned char *b;     // byte arrays, element i of thread t at b[i * nt + t]
    uint32_t      *w;     // word arrays
    int nt, t;
    __device__ unsigned char &B(int i) const { return b[(size_t)i * nt + t]; }
    __device__ uint32_t &Wd(int i) const { return w[(size_t)i * nt + t]; }
};
struct SchedOut {
    unsigned char *dst; int W;
    __device__ void operator()(int l, int t, unsigned char code) const { dst[(size_t)l * W + t] = code; }
};

template <int NL>
__global__ void k_sched_dual(int64_t n, int64_t slice0, int64_t nslices, int wmax, const int64_t *__restrict__ colptr,
                             const int32_t *__restrict__ rowidx, const int64_t *__restrict__ slice_ptr,
                             int4 *__restrict__ sell, unsigned long long *__restrict__ overflow_total)
{
    constexpr int PARTS = 32 / NL;                        // problems per slice: 2 half-warps or 4 quarter-warps
    extern __shared__ __align__(16) unsigned char sraw[];
    const int nt = blockDim.x, tid = threadIdx.x;
    SchedMem M;
    M.b = sraw; M.nt = nt; M.t = tid;
    M.w = nullptr;                                        // the Birkhoff-von Neumann scheduler needs bytes only
    (void)wmax;

    const int64_t prob = (int64_t)blockIdx.x * nt + tid;  // half- / quarter-warp index
    if (prob >= PARTS * nslices) return;
    const int64_t slice = slice0 + prob / PARTS;           // slices [slice0, slice0 + nslices)
    const int half = (int)(prob % PARTS);
    const int64_t base = slice_ptr[slice];
    const int W = (int)((slice_ptr[slice + 1] - base) >> 5) * 2;
    if (W == 0) return;
    for (int i = 0; i < NL * NL; ++i) M.B(i) = 0;
    for (int l = 0; l < NL; ++l) {
        const int64_t j = slice * SKM_SLICE + half * NL + l;
        int64_t a = 0, b = 0;
        if (j < n) { a = colptr[j]; b = colptr[j + 1]; }
        for (int64_t t = a; t < b; ++t) M.B(l * NL + (rowidx[t] & (NL - 1))) += 1;
    }
    SchedOut out;
    out.dst = reinterpret_cast<unsigned char *>(sell + base) + (size_t)half * NL * W;
    out.W = W;
    const int ovf = skm_sched_bvn<NL>(M, W, out);
    if (ovf) atomicAdd(overflow_total, (unsigned long long)ovf);
}

template <typename VT>
__global__ void k_fill_sell_sched(int64_t p, int64_t n, int64_t slice0, int64_t nslices, int wmax, int boff, int nl,
                                  const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                                  const VT *__restrict__ val, const int64_t *__restrict__ slice_ptr,
                                  int4 *__restrict__ sell)
{
    extern __shared__ unsigned short s_all[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nwb = blockDim.x >> 5;
    if ((int64_t)blockIdx.x * nwb + wib >= nslices) return;
    const int64_t warp = slice0 + (int64_t)blockIdx.x * nwb + wib;      // slice index
    // per warp: ord[wmax][32] (u16), start[16][32] (u16), code[wmax][32] (u8)
    unsigned short *ord = s_all + (size_t)wib * (wmax * 32 + 16 * 32 + wmax * 16);
    unsigned short *st = ord + (size_t)wmax * 32;
    unsigned char *code = reinterpret_cast<unsigned char *>(st + 16 * 32);
    const int64_t j = warp * SKM_SLICE + lane;
    int64_t a = 0, b = 0;
    if (j < n) { a = colptr[j]; b = colptr[j + 1]; }
    const int len = (int)(b - a);
    const int64_t base = slice_ptr[warp];
    const int w2 = (int)((slice_ptr[warp + 1] - base) >> 5);
    const int W = 2 * w2;
    const unsigned char *sched = reinterpret_cast<const unsigned char *>(sell + base) + (size_t)lane * W;
    for (int t = 0; t < W; ++t) code[t * 32 + lane] = sched[t];
    // counting sort of this lane's entries by (row & (nl - 1))
    const int mk = nl - 1;
    for (int g = 0; g < 16; ++g) st[g * 32 + lane] = 0;
    for (int t = 0; t < len; ++t) st[(rowidx[a + t] & mk) * 32 + lane] += 1;
    int run = 0;
    for (int g = 0; g < 16; ++g) { const int c = st[g * 32 + lane]; st[g * 32 + lane] = (unsigned short)run; run += c; }
    {
        unsigned short fill[16];
#pragma unroll
        for (int g = 0; g < 16; ++g) fill[g] = st[g * 32 + lane];
        for (int t = 0; t < len; ++t) {
            const int g = rowidx[a + t] & mk;
            int pos = 0;
#pragma unroll
            for (int q = 0; q < 16; ++q) { if (q == g) { pos = fill[q]; fill[q] = (unsigned short)(pos + 1); } }
            ord[pos * 32 + lane] = (unsigned short)t;
        }
    }
    __syncwarp();                                            // every lane has read its codes: the region may be overwritten
    int4 q4 = make_int4((int)p, 0, (int)p, 0);
    for (int step = 0; step < W; ++step) {
        const int cd = code[step * 32 + lane];
        int r, xb = 0;
        if (cd & 0x80) r = (int)p + (((cd & 15) - (int)(p & mk)) & mk);
        else {
            const int g = cd & 15;
            const int pos = st[g * 32 + lane];
            st[g * 32 + lane] = (unsigned short)(pos + 1);
            const int e = ord[pos * 32 + lane];
            r = rowidx[a + e] + ((cd & 0x10) ? boff : 0);
            xb = __float_as_int((float)val[a + e]);
        }
        if (step & 1) {
            q4.z = r; q4.w = xb;
            sell[base + (int64_t)(step >> 1) * 32 + lane] = q4;
        } else { q4.x = r; q4.y = xb; }
    }
}

// verification of a SELL image against the CSC image + wavefront count of its schedule
__global__ void k_sell_check(int64_t p, int64_t n, int64_t nslices, int mode, int boff,
                             const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                             const float *__restrict__ val, const int64_t *__restrict__ slice_ptr,
                             const int4 *__restrict__ sell, unsigned long long *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (warp >= nslices) return;
    const int64_t j = warp * SKM_SLICE + lane;
    int64_t a = 0, b = 0;
    if (j < n) { a = colptr[j]; b = colptr[j + 1]; }
    const int64_t base = slice_ptr[warp];
    const int w2 = (int)((slice_ptr[warp + 1] - base) >> 5);
    // order-independent fingerprints of the column: count, sum of rows, xor/sum of mixed (row, value) hashes
    unsigned long long c0 = 0, s0 = 0, h0 = 0, c1 = 0, s1 = 0, h1 = 0;
    auto mix = [](unsigned r, unsigned v) -> unsigned long long {
        unsigned long long z = ((unsigned long long)r << 32 | v) + 0x9e3779b97f4a7c15ULL;
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
        return z ^ (z >> 31);
    };
    for (int64_t t = a; t < b; ++t) { c0++; s0 += (unsigned)rowidx[t]; h0 += mix((unsigned)rowidx[t], (unsigned)__float_as_int(val[t])); }
    unsigned long long steps = 0, waves = 0;
    const int group = mode == 1 ? 16 : 8, mod = group;
    bool bad = false;
    for (int t = 0; t < 2 * w2; ++t) {
        const int4 q = sell[base + (int64_t)(t >> 1) * 32 + lane];
        int r = (t & 1) ? q.z : q.x;
        const int xb = (t & 1) ? q.w : q.y;
        const int cls = r & (mod - 1);
        // wavefronts of this step: max multiplicity of a class within each group of lanes
        int mult = 0;
        for (int o = 0; o < group; ++o) {
            const int oc = __shfl_sync(0xffffffffu, cls, (lane & ~(group - 1)) | o);
            mult += (oc == cls);
        }
        int mx = mult;
        for (int o = group >> 1; o; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        if ((lane & (group - 1)) == 0) { steps += 1; waves += mx; }
        if (mode >= 1 && r >= boff) { r -= boff; if (r >= p) bad = true; }
        if (r >= p) {                                        // pad: a zero row, value 0
            if (xb != 0 || r >= p + (mode >= 1 ? 16 : 1)) bad = true;
            continue;
        }
        c1++; s1 += (unsigned)r; h1 += mix((unsigned)r, (unsigned)xb);
    }
    if (c0 != c1 || s0 != s1 || h0 != h1) bad = true;
    if (j < n && bad) atomicAdd(&out[0], 1ULL);
    if (steps) { atomicAdd(&out[1], steps); atomicAdd(&out[2], waves); }
}

}  // namespace


namespace {
template <typename S>
__global__ void k_rebase(const S *__restrict__ src, int64_t count, int64_t base, int64_t *__restrict__ dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) dst[i] = (int64_t)src[i] - base;
}
__global__ void k_width_to_elems(int64_t nslices, const int32_t *__restrict__ w2, int64_t *__restrict__ elems)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nslices) elems[i] = (int64_t)w2[i] * 32;
    if (i == nslices) elems[i] = 0;
}
}  // namespace

int skm_launch_validate_async(skm_ctx *ctx, int64_t p, int64_t n, int64_t nnz, const int64_t *colptr,
                              const int32_t *rowidx, int *flags_dev)
{
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (n > 0) {
        int64_t blocks = (n + 255) / 256;
        if (blocks > cap) blocks = cap;
        k_validate_cols<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, nnz, colptr, flags_dev);
        SKM_CHECK_LAUNCH(ctx);
    }
    if (nnz > 0) {
        int64_t blocks = (nnz + 255) / 256;
        if (blocks > cap) blocks = cap;
        k_validate_rows<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, nnz, rowidx, flags_dev);
        SKM_CHECK_LAUNCH(ctx);
    }
    return SKM_OK;
}

// the two halves of the validation on their own (pipelined dataset creation): flags_dev as above
int skm_launch_validate_cols_async(skm_ctx *ctx, int64_t n, int64_t nnz, const int64_t *colptr, int *flags_dev)
{
    if (n <= 0) return SKM_OK;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_validate_cols<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, nnz, colptr, flags_dev);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}
int skm_launch_validate_rows_async(skm_ctx *ctx, int64_t p, int64_t count, const int32_t *rowidx, int *flags_dev)
{
    if (count <= 0) return SKM_OK;
    int64_t blocks = (count + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    k_validate_rows<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, count, rowidx, flags_dev);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_rebase_colptr(skm_ctx *ctx, const void *src, int src_type, int64_t count, int64_t base, int64_t *dst)
{
    if (count == 0) return SKM_OK;
    int64_t blocks = (count + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (src_type == SKM_I64) k_rebase<int64_t><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const int64_t *)src, count, base, dst);
    else k_rebase<int32_t><<<(unsigned)blocks, 256, 0, ctx->stream>>>((const int32_t *)src, count, base, dst);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

size_t skm_sell_scan_tmp_bytes(int64_t nslices)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int64_t *)nullptr, (int64_t *)nullptr, nslices + 1);
    return bytes;
}

// plain (stored-order) SELL image of a view whose colptr/rowidx/val are on the device; slice_ptr and
// sell must be preallocated (slice_ptr: nslices+1).  No host synchronisation.
int skm_build_sell_async(skm_ctx *ctx, skm_dataset *ds, int32_t *w2, int64_t *elems, void *cub_tmp, size_t cub_tmp_bytes)
{
    const int64_t n = ds->n, nslices = (n + SKM_SLICE - 1) / SKM_SLICE;
    ds->nslices = nslices;
    ds->uniform_width = false;
    ds->sell_mode = 0;
    ds->sell_plain = true;
    if (nslices == 0) return SKM_OK;
    int64_t blocks = (nslices * 32 + 255) / 256;
    k_slice_width<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, nslices, ds->colptr, w2);
    SKM_CHECK_LAUNCH(ctx);
    k_width_to_elems<<<(unsigned)((nslices + 1 + 255) / 256), 256, 0, ctx->stream>>>(nslices, w2, elems);
    SKM_CHECK_LAUNCH(ctx);
    SKM_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp, cub_tmp_bytes, elems, ds->slice_ptr, nslices + 1, ctx->stream));
    ctx->launches++;
    k_fill_sell<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->p, n, nslices, ds->colptr, ds->rowidx,
                                                                 (const float *)ds->val, ds->slice_ptr, ds->sell);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_convert_index(skm_ctx *ctx, const void *src, int src_type, int64_t count, void *dst,
                             int dst_is_i64)
{
    if (src_type == SKM_I64) {
        return dst_is_i64 ? convert<int64_t, int64_t>(ctx, src, dst, count)
                          : convert<int64_t, int32_t>(ctx, src, dst, count);
    } else if (src_type == SKM_I32) {
        return dst_is_i64 ? convert<int32_t, int64_t>(ctx, src, dst, count)
                          : convert<int32_t, int32_t>(ctx, src, dst, count);
    } else if (src_type == SKM_U16) {
        return dst_is_i64 ? convert<uint16_t, int64_t>(ctx, src, dst, count)
                          : convert<uint16_t, int32_t>(ctx, src, dst, count);
    }
    skm_set_error("index type must be SKM_I32, SKM_I64 or SKM_U16");
    return SKM_ERR_INVALID;
}

int skm_launch_convert_value(skm_ctx *ctx, const void *src, int src_type, int64_t count, void *dst,
                             int dst_type)
{
    if (src_type == SKM_F64 && dst_type == SKM_F64) return convert<double, double>(ctx, src, dst, count);
    if (src_type == SKM_F64 && dst_type == SKM_F32) return convert<double, float>(ctx, src, dst, count);
    if (src_type == SKM_F32 && dst_type == SKM_F64) return convert<float, double>(ctx, src, dst, count);
    if (src_type == SKM_F32 && dst_type == SKM_F32) return convert<float, float>(ctx, src, dst, count);
    skm_set_error("value type must be SKM_F32 or SKM_F64");
    return SKM_ERR_INVALID;
}

int skm_validate_csc(skm_ctx *ctx, int64_t p, int64_t n, int64_t nnz, const int64_t *colptr,
                     const int32_t *rowidx, int64_t *max_col_nnz)
{
    SKM_CUDA(cudaMemsetAsync(ctx->d_flag, 0, 4 * sizeof(int), ctx->stream));
    int64_t cap = (int64_t)ctx->sm_count * 16;
    if (n > 0) {
        int64_t blocks = (n + 255) / 256;
        if (blocks > cap) blocks = cap;
        k_validate_cols<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, nnz, colptr, ctx->d_flag);
        SKM_CHECK_LAUNCH(ctx);
    }
    if (nnz > 0) {
        int64_t blocks = (nnz + 255) / 256;
        if (blocks > cap) blocks = cap;
        k_validate_rows<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, nnz, rowidx, ctx->d_flag);
        SKM_CHECK_LAUNCH(ctx);
    }
    SKM_CUDA(cudaMemcpyAsync(ctx->h_flag, ctx->d_flag, 4 * sizeof(int), cudaMemcpyDeviceToHost,
                             ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_flag[0] & 1) {
        skm_set_error("invalid CSC column pointers (must start at 0, be non-decreasing, end at nnz)");
        return SKM_ERR_INVALID;
    }
    if (ctx->h_flag[0] & 2) {
        skm_set_error("invalid CSC row index (must lie in [0,p))");
        return SKM_ERR_INVALID;
    }
    *max_col_nnz = ctx->h_flag[1];
    return SKM_OK;
}

int skm_build_sell(skm_dataset *ds)
{
    skm_ctx *ctx = ds->ctx;
    int64_t n = ds->n;
    int64_t nslices = (n + SKM_SLICE - 1) / SKM_SLICE;
    ds->nslices = nslices;
    ds->sell = nullptr;
    ds->slice_ptr = nullptr;
    ds->sell_elems = 0;
    if (nslices == 0) return SKM_OK;

    DevBuf w2;
    SKM_TRY(w2.alloc(sizeof(int32_t) * nslices));
    {
        int64_t threads_total = nslices * 32;
        int64_t blocks = (threads_total + 255) / 256;
        k_slice_width<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, nslices, ds->colptr, w2.as<int32_t>());
        SKM_CHECK_LAUNCH(ctx);
    }
    std::vector<int32_t> hw(nslices);
    SKM_CUDA(cudaMemcpyAsync(hw.data(), w2.ptr, sizeof(int32_t) * nslices, cudaMemcpyDeviceToHost,
                             ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<int64_t> hp(nslices + 1);
    hp[0] = 0;
    bool uniform = true;
    for (int64_t s = 0; s < nslices; ++s) {
        hp[s + 1] = hp[s] + (int64_t)hw[s] * 32;
        if (hw[s] != hw[0]) uniform = false;
    }
    ds->uniform_width = uniform;
    ds->sell_width2 = hw[0];
    ds->sell_elems = hp[nslices];

    void *d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(int64_t) * (nslices + 1));
    if (e != cudaSuccess) { skm_set_error("cudaMalloc(slice_ptr) failed: %s", cudaGetErrorString(e)); return SKM_ERR_NOMEM; }
    ds->slice_ptr = (int64_t *)d;
    SKM_CUDA(cudaMemcpyAsync(ds->slice_ptr, hp.data(), sizeof(int64_t) * (nslices + 1),
                             cudaMemcpyHostToDevice, ctx->stream));
    size_t sell_bytes = sizeof(int4) * (size_t)(ds->sell_elems > 0 ? ds->sell_elems : 1);
    SKM_TRY(skm_big_alloc(ctx, &d, sell_bytes, "sell"));
    ds->sell = (int4 *)d;
    ds->device_bytes += (int64_t)sell_bytes + (int64_t)sizeof(int64_t) * (nslices + 1);
    int wmax = 0;
    for (int64_t s2 = 0; s2 < nslices; ++s2) wmax = hw[s2] * 2 > wmax ? hw[s2] * 2 : wmax;
    ds->sell_wmax = wmax;
    ds->sell_mode = -1;                 // filled lazily, in the entry order of the first kernel family that reads it
    ds->sell_plain = false;
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));   // hp goes out of scope
    return SKM_OK;
}

// Dual-table entry order (layout 1 or 2) of slices [slice0, slice0 + nsl): scheduler + fill, asynchronous on the context
// stream.  The CSC entries of those slices' columns must be on the device; used slice range by slice range while an
// upload is still in flight (api.cu, pipelined dataset creation) and for the whole image by skm_sell_ensure_layout.
int skm_sell_layout_range(skm_dataset *ds, int mode, int64_t slice0, int64_t nsl, unsigned long long *ovf_dev)
{
    skm_ctx *ctx = ds->ctx;
    if (nsl <= 0) return SKM_OK;
    const int wmax = ds->sell_wmax;
    if (wmax <= 0 || wmax > 254) { skm_set_error("dual-table layout needs columns of at most 254 entries"); return SKM_ERR_UNSUPPORTED; }
    const int nl = mode == 1 ? 16 : 8;
    // scheduler: one thread per half-/quarter-warp, as many as the scratch allows (interleaved shared memory)
    const size_t per = (size_t)skm_bvn_bytes(nl);
    int nt = (int)(((size_t)ctx->smem_optin - 1024) / per);
    nt = nt >= 256 ? 256 : (nt & ~31);
    if (nt < 32) { skm_set_error("dual-table scheduler does not fit in shared memory"); return SKM_ERR_UNSUPPORTED; }
    const size_t smem = per * nt;
    const int64_t probs = (32 / nl) * nsl;
    if (nl == 16) {
        SKM_CUDA(cudaFuncSetAttribute(k_sched_dual<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sched_dual<16><<<(unsigned)((probs + nt - 1) / nt), nt, smem, ctx->stream>>>(
            ds->n, slice0, nsl, wmax, ds->colptr, ds->rowidx, ds->slice_ptr, ds->sell, ovf_dev);
    } else {
        SKM_CUDA(cudaFuncSetAttribute(k_sched_dual<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_sched_dual<8><<<(unsigned)((probs + nt - 1) / nt), nt, smem, ctx->stream>>>(
            ds->n, slice0, nsl, wmax, ds->colptr, ds->rowidx, ds->slice_ptr, ds->sell, ovf_dev);
    }
    SKM_CHECK_LAUNCH(ctx);
    int warps = 8;
    const size_t per_warp = (size_t)wmax * 64 + 16 * 64 + (size_t)wmax * 32;
    while (warps > 1 && (size_t)warps * per_warp > (size_t)ctx->smem_optin - 1024) warps >>= 1;
    const size_t smem2 = (size_t)warps * per_warp;
    SKM_CUDA(cudaFuncSetAttribute(k_fill_sell_sched<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    k_fill_sell_sched<float><<<(unsigned)((nsl + warps - 1) / warps), warps * 32, smem2, ctx->stream>>>(
        ds->p, ds->n, slice0, nsl, wmax, (int)skm_dual_boff(ds->p), nl, ds->colptr, ds->rowidx, (const float *)ds->val,
        ds->slice_ptr, ds->sell);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// (Re)write the SELL image in the entry order of kernel family `mode`; widths and slice_ptr are
// unchanged, so this is an in-place rewrite from the CSC image.
int skm_sell_ensure_layout(skm_dataset *ds, int mode)
{
    skm_ctx *ctx = ds->ctx;
    if (!ds->sell || ds->sell_elems == 0 || ds->nslices == 0) { ds->sell_mode = mode; return SKM_OK; }
    if (ds->sell_mode == mode) return SKM_OK;
    const int64_t n = ds->n, nslices = ds->nslices;
    const int wmax = ds->sell_wmax;
    if (ds->store_dtype != SKM_F32) {
        int64_t blocks = (nslices * 32 + 255) / 256;
        k_fill_sell<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            ds->p, n, nslices, ds->colptr, ds->rowidx, (const double *)ds->val, ds->slice_ptr, ds->sell);
        SKM_CHECK_LAUNCH(ctx);
        ds->sell_mode = mode; ds->sell_plain = true;
        return SKM_OK;
    }
    if (mode == 1 || mode == 2) {
        DevBuf ovf;
        SKM_TRY(ovf.alloc(sizeof(unsigned long long)));
        SKM_CUDA(cudaMemsetAsync(ovf.ptr, 0, sizeof(unsigned long long), ctx->stream));
        SKM_TRY(skm_sell_layout_range(ds, mode, 0, nslices, ovf.as<unsigned long long>()));
        SKM_CUDA(cudaStreamSynchronize(ctx->stream));          // ovf dies here
        ds->sell_mode = mode; ds->sell_plain = false;
        return SKM_OK;
    }
    const bool banked = wmax > 0 && wmax <= 1024 && !getenv("SKM_NO_BANKED");
    if (banked) {
        int warps = 8;
        while (warps > 1 && (size_t)warps * wmax * 64 > (size_t)ctx->smem_optin - 1024) warps >>= 1;
        const size_t smem = (size_t)warps * wmax * 64;
        SKM_CUDA(cudaFuncSetAttribute(k_fill_sell_banked<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int64_t blocks = (nslices + warps - 1) / warps;
        k_fill_sell_banked<float><<<(unsigned)blocks, warps * 32, smem, ctx->stream>>>(
            ds->p, n, nslices, wmax, ds->colptr, ds->rowidx, (const float *)ds->val, ds->slice_ptr, ds->sell);
        SKM_CHECK_LAUNCH(ctx);
        ds->sell_plain = false;
    } else {
        int64_t blocks = (nslices * 32 + 255) / 256;
        k_fill_sell<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            ds->p, n, nslices, ds->colptr, ds->rowidx, (const float *)ds->val, ds->slice_ptr, ds->sell);
        SKM_CHECK_LAUNCH(ctx);
        ds->sell_plain = true;
    }
    ds->sell_mode = 0;
    return SKM_OK;
}

// Any valid entry order will do (the bounded pass gathers one value per entry, bank conflicts do not matter):
// keep what is there, or fill the image in stored order -- the cheapest fill.
int skm_sell_ensure_any(skm_dataset *ds)
{
    skm_ctx *ctx = ds->ctx;
    if (!ds->sell || ds->sell_elems == 0 || ds->nslices == 0 || ds->sell_mode >= 0) return SKM_OK;
    if (ds->store_dtype != SKM_F32) return skm_sell_ensure_layout(ds, 0);
    const int64_t blocks = (ds->nslices * 32 + 255) / 256;
    k_fill_sell<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
        ds->p, ds->n, ds->nslices, ds->colptr, ds->rowidx, (const float *)ds->val, ds->slice_ptr, ds->sell);
    SKM_CHECK_LAUNCH(ctx);
    ds->sell_plain = true;
    ds->sell_mode = 3;                   // stored order: no kernel family asks for it, so a later bind re-lays it out
    return SKM_OK;
}

int skm_sell_check(skm_dataset *ds, int64_t out[3])
{
    skm_ctx *ctx = ds->ctx;
    out[0] = out[1] = out[2] = 0;
    if (!ds->sell || ds->nslices == 0 || ds->store_dtype != SKM_F32) return SKM_OK;
    if (ds->sell_mode < 0) SKM_TRY(skm_sell_ensure_layout(ds, 0));
    DevBuf r;
    SKM_TRY(r.alloc(3 * sizeof(unsigned long long)));
    SKM_CUDA(cudaMemsetAsync(r.ptr, 0, 3 * sizeof(unsigned long long), ctx->stream));
    int64_t blocks = (ds->nslices * 32 + 255) / 256;
    k_sell_check<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->p, ds->n, ds->nslices, (ds->sell_mode >= 1 && !ds->sell_plain) ? ds->sell_mode : 0,
                                                           (int)skm_dual_boff(ds->p), ds->colptr, ds->rowidx,
                                                           (const float *)ds->val, ds->slice_ptr, ds->sell,
                                                           r.as<unsigned long long>());
    SKM_CHECK_LAUNCH(ctx);
    unsigned long long h[3];
    SKM_CUDA(cudaMemcpyAsync(h, r.ptr, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));
    out[0] = (int64_t)h[0]; out[1] = (int64_t)h[1]; out[2] = (int64_t)h[2];
    return SKM_OK;
}
