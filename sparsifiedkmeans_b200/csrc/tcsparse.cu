// tcsparse.cu -- K1 for many centres (K >= 24) on the 5th-generation tensor cores.
//
// The masked distance of private/SparseMatrixMinusCluster.c:169-182 expands, per column j with support W_j, into
//     d_j(k)^2 = |x_j|^2 + s_j(k),      s_j(k) = sum_{r in W_j} ( c_rk^2 - 2 x_jr c_rk ),
// i.e. a product of the DENSIFIED column [y_j | m_j] (values and 0/1 mask, interleaved) with the table [-2c ; c^2].
// The gather kernels of assign_fast.cu pay 4 bytes of shared-memory bandwidth per (stored entry, centre): 256 B per
// entry at K = 64, 4.5 ms per 1.25e7-column shard whatever the instruction mix (profiles/r2_k64_probe.md).  Here the
// column is densified ON THE FLY into a swizzled fp16 operand tile in shared memory -- one 4-byte store per stored
// entry, 64x less shared-memory traffic -- and the 20x redundant dense product runs on tcgen05.mma, which has the
// throughput to spare (2 x 128 x N x p2 MACs per 128-column tile).
//
// fp16 operands make this a FILTER, not the answer: the epilogue keeps the three best scores, turns the fourth and
// the second into rigorous lower bounds (rounding analysis below), and the exact evaluation of the candidate(s)
// -- the same fp32 sum and rounding guard as everywhere else (k_assign_bounded, then k_tcs_resolve for the columns
// whose runner-up is too close, then the fp64 kernel in the reference's order for what is left) -- decides.  A winner
// is therefore the reference's winner (first index on exact ties included), never the tensor cores'.
//
// Image ("TSB": tile / stripe blocks).  Tile = 128 consecutive columns (UMMA M), stripe = 64 consecutive rows
// (128 fp16 reduction elements = 2 swizzle atoms of 128 B).  The entries of block (tile, stripe) are contiguous,
// sorted by column, 8 bytes each: key = byte offset of the (value, mask) pair inside the operand tile | column-in-tile
// << 16 | row-in-stripe << 24, and the fp32 bits of sigma * value (sigma: the power of two that brings max |x| below
// 64, so the product is exact).  blk_ptr[tile * S + stripe] is the block's first entry.
//
// k_tcs_filter<BN>  (one persistent CTA per SM, 12 warps at N = 64)
//   warps 0..3  epilogue: thread = column (TMEM lane), tcgen05.ld, top-4 of the scores, bounds, 16 B out per column
//   warp 4      stages the centre image of the next stripe (BN x 128 fp16, already swizzled in global memory) with
//               cp.async.bulk into a ring behind full/empty mbarriers
//   warp 5      MMA issuer: 8 x tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = BN, K = 16) per block into the
//               tile's accumulator; T = 256/BN tiles share one staged stripe, two sets of accumulators (2 x 256
//               TMEM columns) let the epilogue of one super-tile overlap the products of the next
//   warps 6..   densifiers, one per operand stage (6 stages at N = 64, 4 at N = 128): the entries of the stage's next block are prefetched into
//               registers (coalesced 8-byte loads), the previous block's pairs are zeroed again once its MMAs have
//               completed (tcgen05.commit -> mbarrier), the new pairs are stored (cvt.rn.f16x2.f32 packs the value
//               with the mask's 1.0), fence.proxy.async, arrive
// Bound -- MEASURED, and the reason this plan is opt-in (profiles/r2_tcsparse.md): not the tensor pipe (32 cycles of math
// per MMA at N = 64, 4096 per tile) but the shared-memory operand fetch of the MMAs, ~64 B/clk/SM: the densified tile is
// 512 KB per 128 columns + 256 KB of centre image = 12 K cycles per tile, 739 measured per block of 48 KB whatever the
// number of stages or densifier warps (N = 128: 64 KB per block, 911 cycles).  4.2 ms per 1.25e7-column shard against
// 5.35 ms for the gather kernels: the 20x redundant dense operand costs as much shared-memory bandwidth as the gathers.
#include "common.cuh"
#include <cub/cub.cuh>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

namespace {

constexpr int TS_TILE = 128;                         // columns per tile (UMMA M)
constexpr int TS_SR = 64;                            // rows per stripe
constexpr int TS_A_BYTES = TS_TILE * TS_SR * 2 * 2;  // 128 rows x 128 fp16
constexpr int TS_DW = 1;                             // densifier warps per operand stage.  ONE: the zeros of the previous block
                                                     // and the pairs of the next may share positions, and only program order
                                                     // inside a warp orders them (two warps per stage raced; the kernel is bound
                                                     // by the operand fetch of the MMAs, not by the densifiers, see below)
constexpr int TS_REGE = 16;                          // entries per lane held in registers per block (16 x 32 = 512 per block)
// operand stages: the stage round trip (MMAs complete -> zeros back -> new pairs -> fence -> MMA issue) is ~1500 cycles
// against 256 (N = 64) or 512 (N = 128) cycles of MMA per block, so as many stages as shared memory holds
// HAMMER > 0 (micro-benchmark only, SKM_TC_HAMMER=1): that many extra warps stream LDS.128 from a 16 KB region for as
// long as the filter runs, to measure whether the MMAs' operand fetch and the LSU share their shared-memory bandwidth
__host__ __device__ constexpr int ts_stages(int BN, int HAMMER = 0) { return BN <= 64 ? (HAMMER ? 5 : 6) : 4; }
__host__ __device__ constexpr int ts_threads(int BN, int HAMMER = 0) { return 32 * (6 + ts_stages(BN, HAMMER) * TS_DW + HAMMER); }
constexpr int TS_HAMMER_BYTES = 16384;
constexpr int TS_MAXS = 64;                          // stripes the builder keeps counters for (p <= 4096)

// ---------------------------------------------------------------- PTX helpers (same idioms as tcgemm.cu)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0, spins = 0;
    const uint32_t a = s32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity), "r"(2000u) : "memory");     // suspend-time hint: idle warps leave the issue slots alone
        if (!done && ++spins > (1u << 24)) __trap();          // a lost arrival is a bug, not a hang
    }
}
// waits on the critical path of an operand stage: plain polling (a suspended warp wakes up hundreds of cycles late)
__device__ __forceinline__ void mbar_spin(uint64_t *bar, uint32_t parity)
{
    uint32_t done = 0, spins = 0;
    const uint32_t a = s32(bar);
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 28)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// K-major operand, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart, descriptor version 1
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint2 ld_stream8(const uint2 *p)
{
    uint2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

// byte offset of the (value, mask) fp16 pair of (column-in-tile j, row-in-stripe r) inside one operand stage:
// k-block r / 32 (16 KB each), then the canonical K-major 128-byte-swizzle image of a 128 x 64-element tile
__host__ __device__ __forceinline__ uint32_t tsb_offset(uint32_t j, uint32_t r)
{
    const uint32_t kb = r >> 5, rr = r & 31;
    return kb * 16384u + (j >> 3) * 1024u + (j & 7) * 128u + ((((rr >> 2) ^ (j & 7)) & 7) << 4) + (rr & 3) * 4u;
}

// ---------------------------------------------------------------- image builder
// pass 1: entries per (tile, stripe) block, max |x|
__global__ void __launch_bounds__(TS_TILE) k_tsb_count(int64_t n, int S, const int64_t *__restrict__ colptr,
                                                       const int32_t *__restrict__ rowidx, const float *__restrict__ val,
                                                       int64_t *__restrict__ blk_cnt /* [ntiles*S] */,
                                                       int *__restrict__ xmax_bits, int *__restrict__ max_blk)
{
    __shared__ int cnt[TS_MAXS];
    const int64_t tile = blockIdx.x;
    const int64_t j = tile * TS_TILE + threadIdx.x;
    for (int s = threadIdx.x; s < S; s += TS_TILE) cnt[s] = 0;
    __syncthreads();
    float xm = 0.f;
    if (j < n) {
        const int64_t e0 = colptr[j], e1 = colptr[j + 1];
        for (int64_t e = e0; e < e1; ++e) {
            const float v = val[e];
            xm = fmaxf(xm, fabsf(v));
            if (!(v == v)) xm = __int_as_float(0x7fc00000);
            atomicAdd(&cnt[rowidx[e] >> 6], 1);
        }
    }
    int mi = __float_as_int(xm);
#pragma unroll
    for (int o = 16; o; o >>= 1) mi = max(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    if ((threadIdx.x & 31) == 0 && mi > 0) atomicMax(xmax_bits, mi);
    __syncthreads();
    int mb = 0;
    for (int s = threadIdx.x; s < S; s += TS_TILE) { blk_cnt[tile * S + s] = cnt[s]; mb = max(mb, cnt[s]); }
    if (mb > 0) atomicMax(max_blk, mb);
}

// pass 2: write the entries of every block, sorted by column, values multiplied by sigma (a power of two: exact),
// and the per-column |sigma x|^2 (fp32)
__global__ void __launch_bounds__(TS_TILE) k_tsb_fill(int64_t n, int S, float sigma, const int64_t *__restrict__ colptr,
                                                      const int32_t *__restrict__ rowidx, const float *__restrict__ val,
                                                      const int64_t *__restrict__ blk_ptr, uint2 *__restrict__ tsb,
                                                      float *__restrict__ colnorm2)
{
    extern __shared__ unsigned short pos[];          // [S][128]: count, then running position inside the block
    const int64_t tile = blockIdx.x;
    const int tj = threadIdx.x;
    const int64_t j = tile * TS_TILE + tj;
    for (int s = 0; s < S; ++s) pos[s * TS_TILE + tj] = 0;
    int64_t e0 = 0, e1 = 0;
    if (j < n) {
        e0 = colptr[j]; e1 = colptr[j + 1];
        for (int64_t e = e0; e < e1; ++e) pos[(rowidx[e] >> 6) * TS_TILE + tj] += 1;
    }
    __syncthreads();
    // exclusive prefix over the columns of each stripe: one warp per stripe, four columns per lane
    const int warp = tj >> 5, lane = tj & 31;
    for (int s = warp; s < S; s += TS_TILE / 32) {
        unsigned short *row = pos + s * TS_TILE;
        int c0 = row[lane * 4], c1 = row[lane * 4 + 1], c2 = row[lane * 4 + 2], c3 = row[lane * 4 + 3];
        int tot = c0 + c1 + c2 + c3, inc = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        int ex = inc - tot;
        row[lane * 4] = (unsigned short)ex; ex += c0;
        row[lane * 4 + 1] = (unsigned short)ex; ex += c1;
        row[lane * 4 + 2] = (unsigned short)ex; ex += c2;
        row[lane * 4 + 3] = (unsigned short)ex;
    }
    __syncthreads();
    float acc = 0.f;
    for (int64_t e = e0; e < e1; ++e) {
        const int r = rowidx[e];
        const int s = r >> 6, rl = r & 63;
        const int at = pos[s * TS_TILE + tj]++;
        const float v = val[e] * sigma;
        acc = fmaf(v, v, acc);
        const uint32_t key = tsb_offset((uint32_t)tj, (uint32_t)rl) | ((uint32_t)tj << 16) | ((uint32_t)rl << 24);
        tsb[blk_ptr[tile * S + s] + at] = make_uint2(key, __float_as_uint(v));
    }
    if (j < n) colnorm2[j] = acc;
}

// ---------------------------------------------------------------- per-iteration centre image
// scale[0] = sigma (power of two fixed when the image was built: sigma max|x| < 64), scale[1] = 1/sigma,
// scale[2] != 0: the filter is off for this call (a centre entry that is not finite, or so much larger than the data
// that its square would leave the fp16 range; every column then goes to the exact kernels)
__global__ void k_tcs_scale(float sigma, const float *__restrict__ cmax, float *__restrict__ scale)
{
    const float cm = *cmax * sigma;
    scale[0] = sigma;
    scale[1] = 1.f / sigma;
    scale[2] = (cm < 128.f) ? 0.f : 1.f;                          // false for NaN
    scale[3] = cm;
}

// bimg[stripe][k-block][centre row][64 fp16, swizzled]: element 2i = -2 fl16(sigma c'), element 2i+1 = fl16((sigma c')^2)
// for row stripe*64 + kblock*32 + i of the stripe; one thread per 16-byte chunk (4 rows)
__global__ void k_tcs_centres(int64_t p, int64_t K, int BN, int S, const double *__restrict__ ct /* [p+1][K] */,
                              const float *__restrict__ scale, uint4 *__restrict__ bimg)
{
    const int64_t total = (int64_t)S * 2 * BN * 8;
    const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= total) return;
    const int c = (int)(id & 7);
    const int64_t rowid = id >> 3;
    const int k = (int)(rowid % BN);
    const int kb = (int)((rowid / BN) & 1);
    const int s = (int)(rowid / (2 * BN));
    const double sg = (double)scale[0];
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t r = (int64_t)s * TS_SR + kb * 32 + c * 4 + i;
        double cv = 0.0;
        if (r < p && k < K) cv = ct[r * K + k] * sg;
        const __half h = __double2half(cv);
        const __half m2 = __hmul(h, __float2half_rn(-2.f));
        const __half sq = __double2half(cv * cv);
        w[i] = (uint32_t)__half_as_ushort(m2) | ((uint32_t)__half_as_ushort(sq) << 16);
    }
    const size_t byte = ((size_t)s * 2 + kb) * (size_t)BN * 128 + (size_t)(k >> 3) * 1024 + (size_t)(k & 7) * 128 +
                        (size_t)(((c ^ (k & 7)) & 7) << 4);
    bimg[byte >> 4] = make_uint4(w[0], w[1], w[2], w[3]);
}

// ---------------------------------------------------------------- the filter
struct TcsParams {
    const uint2   *tsb;
    const int64_t *blk_ptr;
    int64_t        ntiles, n;
    int            S, K;
    const unsigned char *bimg;
    const float   *scale;
    const float   *colnorm2;       // |sigma x_j|^2
    float          u_eff;          // relative rounding of one fp16-operand term (see the analysis at skm_launch_tcs_filter)
    float          abs_unit;       // absolute part, scaled units, for a column of max_col_nnz entries
    float          ga;             // relative rounding of the stored fp32 |x|^2
    float          sqrt_m;         // sqrt(max_col_nnz)
    int32_t       *assign;         // best centre by score
    float         *lb;             // lower bound on the distance to every other centre
    uint32_t      *cand;           // second | third << 16
    float         *lb4;            // lower bound on the distance to every centre outside the best three
    float         *dbg_scores;     // optional [n][BN] raw scores (tests)
    unsigned long long *hammer_out; // micro-benchmark counters (HAMMER instantiation only)
};

template <int BN, int HAMMER>
__global__ void __launch_bounds__(ts_threads(BN, HAMMER), 1) k_tcs_filter(const TcsParams P)
{
    constexpr int T = 256 / BN;                       // tiles per super-tile: one set of accumulators = 256 TMEM columns
    constexpr int LOG_T = (T == 2) ? 1 : 2;
    constexpr int NB = 2;                             // stages of the centre image
    constexpr int TS_STAGES = ts_stages(BN, HAMMER);
    constexpr int TS_THREADS = ts_threads(BN, HAMMER);
    constexpr uint32_t B_BYTES = BN * 256;            // 2 k-blocks x BN rows x 128 bytes
    static_assert(BN == 64 || BN == 128, "BN");
    extern __shared__ __align__(1024) unsigned char ts_smem[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(ts_smem) + 1023) & ~(uintptr_t)1023);
    unsigned char *smA = base, *smB = base + TS_STAGES * TS_A_BYTES;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smB + NB * B_BYTES);
    uint64_t *a_done = a_full + TS_STAGES, *b_full = a_done + TS_STAGES, *b_empty = b_full + NB;
    uint64_t *acc_full = b_empty + NB, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nst = (P.ntiles + T - 1) / T;
    const int64_t my_st = (nst > (int64_t)blockIdx.x) ? (nst - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == 4 && lane == 0) {
        for (int s = 0; s < TS_STAGES; ++s) { mbar_init(&a_full[s], TS_DW); mbar_init(&a_done[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (HAMMER > 0 && threadIdx.x == 0) { tmem_slot[2] = 0; tmem_slot[3] = 0; }
    {   // operand stages start as zero tiles; the densifiers keep them zero outside the block in flight
        int4 *z = reinterpret_cast<int4 *>(smA);
        for (int i = threadIdx.x; i < TS_STAGES * TS_A_BYTES / 16; i += TS_THREADS) z[i] = make_int4(0, 0, 0, 0);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // ===== epilogue: thread = column; everything in the scaled units of the operands =====
        const float INF = __int_as_float(0x7f800000);
        const float inv_sigma = P.scale[1];
        const bool off = P.scale[2] != 0.f;
        const float cmaxn = P.sqrt_m * P.scale[3] * 1.0001f;
        for (int64_t i = 0; i < my_st; ++i) {
            const uint32_t ab = (uint32_t)(i & 1);
            const int64_t stg = (int64_t)blockIdx.x + i * gridDim.x;
            mbar_wait(&acc_full[ab], (uint32_t)((i >> 1) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int t = 0; t < T; ++t) {
                const int64_t j = (stg * T + t) * TS_TILE + warp * 32 + lane;
                // the four smallest scores with their centres: the centre index replaces the low 7 mantissa bits of the
                // score (a 2^-16 relative perturbation, part of u_eff), so one min/max network sorts both
                float b1 = INF, b2 = INF, b3 = INF, b4 = INF, zs = 0.f;
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + ab * 256 + t * BN + c0, v);
                    if (P.dbg_scores && j < P.n) {
#pragma unroll
                        for (int c = 0; c < 32; ++c) P.dbg_scores[j * BN + c0 + c] = __uint_as_float(v[c]) * inv_sigma * inv_sigma;
                    }
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        const int k = c0 + c;
                        if (k < P.K) {
                            zs = fmaf(__uint_as_float(v[c]), 0.f, zs);                  // NaN / Inf anywhere -> NaN
                            const float key = __uint_as_float((v[c] & 0xffffff80u) | (uint32_t)k);
                            const float c1 = fmaxf(b1, key); b1 = fminf(b1, key);
                            const float c2 = fmaxf(b2, c1);  b2 = fminf(b2, c1);
                            const float c3 = fmaxf(b3, c2);  b3 = fminf(b3, c2);
                            b4 = fminf(b4, c3);
                        }
                    }
                }
                if (j < P.n) {
                    const float xn2 = P.colnorm2[j];
                    const float xn = sqrtf(xn2) * 1.000001f;
                    // s - E(s), E = u_eff (4 |x| rq + rq^2) + abs, rq >= the masked norm of that centre:
                    // from s >= q - 2|x|sqrt(q) - E:  sqrt(q) <= 1.01 (|x| + sqrt(|x|^2 + max(s, 0))); also <= sqrt(m) max|c'|
                    auto lower = [&](float s) -> float {
                        if (!(s < INF)) return INF;
                        float rq = 1.01f * (xn + sqrtf(xn2 + fmaxf(s, 0.f)));
                        rq = fminf(rq, cmaxn);
                        const float E = P.u_eff * (4.f * xn * rq + rq * rq) + P.abs_unit;
                        const float lbsq = xn2 * (1.f - P.ga) + s - E;
                        return lbsq > 0.f ? sqrtf(lbsq) * (1.f - 1.0e-6f) * inv_sigma : 0.f;
                    };
                    float l2 = lower(b2), l4 = lower(b4);
                    if (!(zs == 0.f) || off || !(l2 == l2) || !(l4 == l4)) { l2 = 0.f; l4 = 0.f; }
                    P.assign[j] = (int)(__float_as_uint(b1) & 127u);
                    P.lb[j] = l2;
                    P.cand[j] = (__float_as_uint(b2) & 127u) | ((__float_as_uint(b3) & 127u) << 16);
                    P.lb4[j] = l4;
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
    } else if (warp == 4) {
        // ===== centre image producer =====
        if (elect_one()) {
            uint32_t bq = 0;
            for (int64_t i = 0; i < my_st; ++i)
                for (int s = 0; s < P.S; ++s, ++bq) {
                    const uint32_t bs = bq % NB;
                    mbar_wait(&b_empty[bs], ((bq / NB) & 1) ^ 1);
                    mbar_expect_tx(&b_full[bs], B_BYTES);
                    bulk_g2s(smB + bs * B_BYTES, P.bimg + (size_t)s * B_BYTES, B_BYTES, &b_full[bs]);
                }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        // instruction descriptor: D = F32 (bit 4), A = B = F16 (0 at bits 7 and 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TS_TILE >> 4) << 24);
        uint32_t q = 0, bq = 0;
        for (int64_t i = 0; i < my_st; ++i) {
            const uint32_t ab = (uint32_t)(i & 1);
            mbar_wait(&acc_empty[ab], (uint32_t)(((i >> 1) & 1) ^ 1));
            tc_fence_after();
            for (int s = 0; s < P.S; ++s, ++bq) {
                const uint32_t bs = bq % NB;
                mbar_spin(&b_full[bs], (bq / NB) & 1);
                for (int t = 0; t < T; ++t, ++q) {
                    const uint32_t st = q % TS_STAGES;
                    mbar_spin(&a_full[st], (q / TS_STAGES) & 1);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t d = tmem_base + ab * 256 + t * BN;
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
                            const uint64_t da = umma_desc_k_sw128(s32(smA + st * TS_A_BYTES + kb * 16384));
                            const uint64_t db = umma_desc_k_sw128(s32(smB + bs * B_BYTES + kb * BN * 128));
#pragma unroll
                            for (int k = 0; k < 4; ++k)      // K = 16 fp16 values (32 bytes) per instruction
                                umma_f16(d, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (uint32_t)((s | kb | k) != 0));
                        }
                        umma_commit(&a_done[st]);            // the densifiers may clear the stage once these have read it
                        if (t == T - 1) umma_commit(&b_empty[bs]);
                        if (t == T - 1 && s == P.S - 1) umma_commit(&acc_full[ab]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (HAMMER > 0 && warp >= 6 + TS_STAGES * TS_DW) {
        // ===== micro-benchmark: LSU shared-memory reads while the MMAs fetch their operands =====
        volatile int *stop = reinterpret_cast<volatile int *>(tmem_slot + 2);
        const int4 *hb = reinterpret_cast<const int4 *>(reinterpret_cast<unsigned char *>(tmem_slot) + 64);
        int4 a0 = make_int4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
        unsigned long long iters = 0;
        const uint32_t hba = s32(hb) + lane * 16;
        const long long t0 = clock64();
        for (;;) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {                      // volatile asm: the loads stay in the loop
                int4 v;
                asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                             : "r"(hba + (uint32_t)((r & 31) * 512)));
                a0.x ^= v.x; a1.y ^= v.y; a2.z ^= v.z; a3.w ^= v.w;
            }
            iters += 32;                                      // LDS.128 warp-instructions per trip
            if ((iters & 255) == 0 && *stop) break;
        }
        const long long t1 = clock64();
        if (lane == 0 && P.hammer_out) {
            atomicAdd(P.hammer_out, iters);
            atomicMax(P.hammer_out + 1, (unsigned long long)(t1 - t0));
            if ((a0.x ^ a1.y ^ a2.z ^ a3.w) == 0x12345678) P.hammer_out[2] = 1;      // keep the loads alive
        }
    } else {
        // ===== densifiers: TS_DW warps per operand stage; a stage takes every TS_STAGES-th block of the CTA's sequence
        // (super-tile, stripe, tile); warp h of the stage takes entries h*32 + lane, + 32 TS_DW, ... of the block.
        // Only the stores, the fence and the arrive sit between "the stage's MMAs have completed" and "the stage is
        // full again": the new pairs are packed, and the next block's entries requested, outside that window. =====
        const int dw = warp - 6, st = dw % TS_STAGES, h = dw / TS_STAGES;
        unsigned char *A = smA + st * TS_A_BYTES;
        const int ST = P.S * T;                                  // blocks per super-tile
        // position of a block in the sequence: (super-tile iteration, remainder = stripe * T + tile)
        struct Pos { int64_t i; int rem; };
        auto advance = [&](Pos &x) { x.rem += TS_STAGES; while (x.rem >= ST) { x.rem -= ST; x.i += 1; } };
        auto issue_ptr = [&](const Pos &x, int64_t &o0, int64_t &o1) {
            o0 = 0; o1 = 0;
            if (x.i >= my_st) return;
            const int s = x.rem >> LOG_T, t = x.rem & (T - 1);
            const int64_t tile = ((int64_t)blockIdx.x + x.i * gridDim.x) * T + t;
            if (tile >= P.ntiles) return;
            const int64_t *bp = P.blk_ptr + tile * P.S + s;
            o0 = __ldg(bp); o1 = __ldg(bp + 1);
        };
        Pos pc{0, st};
        while (pc.rem >= ST) { pc.rem -= ST; pc.i += 1; }
        Pos pn = pc;
        advance(pn);
        int64_t o0c, o1c, o0n, o1n;
        issue_ptr(pc, o0c, o1c);
        issue_ptr(pn, o0n, o1n);
        uint32_t prev_off[TS_REGE], off[TS_REGE], pk[TS_REGE];
        uint2 nxt[TS_REGE];
        const uint2 *b_cur = P.tsb + o0c + h * 32 + lane, *b_prev = b_cur;
        int c_cur = (int)(o1c - o0c) - h * 32 - lane, c_prev = 0;       // entries t * 32 * TS_DW < c_cur are this lane's
#pragma unroll
        for (int t = 0; t < TS_REGE; ++t) if (t * (32 * TS_DW) < c_cur) nxt[t] = ld_stream8(b_cur + t * (32 * TS_DW));
        for (int64_t it = 0; pc.i < my_st; ++it) {
            // pack the (value, mask = 1) pairs of this block while the stage is still busy with the previous one
#pragma unroll
            for (int t = 0; t < TS_REGE; ++t) {
                off[t] = nxt[t].x & 0x7fffu;
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk[t]) : "f"(1.0f), "f"(__uint_as_float(nxt[t].y)));
            }
            if (it > 0) {
                // the previous block's MMAs have read the stage: put its zeros back
                mbar_spin(&a_done[st], (uint32_t)((it - 1) & 1));
#pragma unroll
                for (int t = 0; t < TS_REGE; ++t)
                    if (t * (32 * TS_DW) < c_prev) *reinterpret_cast<uint32_t *>(A + prev_off[t]) = 0u;
                for (int e = TS_REGE * 32 * TS_DW; e < c_prev; e += 32 * TS_DW)
                    *reinterpret_cast<uint32_t *>(A + (b_prev[e].x & 0x7fffu)) = 0u;
            }
            // a zero of the old block and a pair of the new one may target the same position from different lanes:
            // order them (compute-sanitizer racecheck flags the pair otherwise; profiles/r2_tcsparse.md)
            __syncwarp();
#pragma unroll
            for (int t = 0; t < TS_REGE; ++t)
                if (t * (32 * TS_DW) < c_cur) *reinterpret_cast<uint32_t *>(A + off[t]) = pk[t];
            for (int e = TS_REGE * 32 * TS_DW; e < c_cur; e += 32 * TS_DW) {
                const uint2 en = b_cur[e];
                uint32_t pe;
                asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pe) : "f"(1.0f), "f"(__uint_as_float(en.y)));
                *reinterpret_cast<uint32_t *>(A + (en.x & 0x7fffu)) = pe;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[st]);
            // next block of this stage: its pointers were requested one iteration ago, its entries are requested now
#pragma unroll
            for (int t = 0; t < TS_REGE; ++t) prev_off[t] = off[t];
            b_prev = b_cur; c_prev = c_cur;
            pc = pn;
            b_cur = P.tsb + o0n + h * 32 + lane;
            c_cur = (int)(o1n - o0n) - h * 32 - lane;
#pragma unroll
            for (int t = 0; t < TS_REGE; ++t) if (t * (32 * TS_DW) < c_cur) nxt[t] = ld_stream8(b_cur + t * (32 * TS_DW));
            advance(pn);
            issue_ptr(pn, o0n, o1n);
        }
    }
    if (HAMMER > 0 && warp < 4) {
        // the epilogue warps finish last: the last one to arrive stops the hammer warps
        __syncwarp();
        if (lane == 0 && atomicAdd(reinterpret_cast<int *>(tmem_slot + 3), 1) == 3) *reinterpret_cast<volatile int *>(tmem_slot + 2) = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------- exact evaluation of the runner-ups
// Columns the bounded pass could not keep (runner-up within the filter's error of the winner): a warp per column
// over the CSC image evaluates the best three candidates exactly (fp32 sums, the usual rounding guard); the winner
// stands if the guard separates it from the other two and the filter's bound excludes everything else.  What is left
// goes to the fp64 kernel.
struct ResolveParams {
    const int64_t *colptr;
    const int32_t *rowidx;
    const float   *val;
    int64_t        p;
    int            K;
    const float   *table_t;        // [K][p+1]
    float          ga, gb_unit, ge_unit;
    const float   *cmax;
    const int32_t *flagged_in;
    const int     *nflag_in;
    const uint32_t *cand;
    const float   *lb4;
    int32_t       *assign;
    float         *dist, *lb;
    int32_t       *flagged_out;
    int           *nflag_out;
};

__global__ void __launch_bounds__(256) k_tcs_resolve(const ResolveParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int total = *P.nflag_in;
    const float cm = *P.cmax;
    const float gb = P.gb_unit * cm, ge = P.ge_unit * cm * cm + 1e-37f;
    const int64_t stride = P.p + 1;
    for (int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; f < total; f += nwarps) {
        const int64_t j = P.flagged_in[f];
        const uint32_t cd = P.cand[j];
        int k[3] = {P.assign[j], (int)(cd & 0xffffu), (int)(cd >> 16)};
        const int nc = P.K < 3 ? P.K : 3;
        const float *r0 = P.table_t + (int64_t)k[0] * stride;
        const float *r1 = P.table_t + (int64_t)k[nc > 1 ? 1 : 0] * stride;
        const float *r2 = P.table_t + (int64_t)k[nc > 2 ? 2 : 0] * stride;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        const int64_t e0 = P.colptr[j], e1 = P.colptr[j + 1];
        for (int64_t e = e0 + lane; e < e1; e += 32) {
            const int r = P.rowidx[e];
            const float x = P.val[e];
            float d;
            d = x - __ldg(r0 + r); a0 = fmaf(d, d, a0);
            d = x - __ldg(r1 + r); a1 = fmaf(d, d, a1);
            d = x - __ldg(r2 + r); a2 = fmaf(d, d, a2);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            a0 += __shfl_xor_sync(0xffffffffu, a0, o);
            a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane) continue;
        const float INF = __int_as_float(0x7f800000);
        float a[3] = {a0, nc > 1 ? a1 : INF, nc > 2 ? a2 : INF};
        int w = 0;
        if (a[1] < a[w]) w = 1;
        if (a[2] < a[w]) w = 2;
        auto guard = [&](float s) { return P.ga * s + gb * sqrtf(s) + ge; };
        const float aw = a[w], Ew = guard(aw);
        bool ok = aw < INF;
        float lo = P.lb4[j];                                   // bound on everything outside the three
        for (int c = 0; c < 3 && ok; ++c) {
            if (c == w || c >= nc) continue;
            if (!(a[c] - aw > Ew + guard(a[c]))) ok = false;   // not separated (or NaN): the fp64 kernel decides
            else lo = fminf(lo, sqrtf(fmaxf(a[c] - guard(a[c]), 0.f)) * (1.f - 1.0e-6f));
        }
        if (ok && !(sqrtf(aw + Ew) * (1.f + 1.0e-6f) < P.lb4[j])) ok = false;
        if (ok) {
            P.assign[j] = k[w];
            P.dist[j] = sqrtf(aw);
            P.lb[j] = lo;
        } else {
            const int slot = atomicAdd(P.nflag_out, 1);
            P.flagged_out[slot] = (int32_t)j;
        }
    }
}

template <int BN, int HAMMER>
int launch_filter_h(skm_ctx *ctx, const TcsParams &P)
{
    constexpr int NB = 2;
    const size_t smem = 1024 + (size_t)ts_stages(BN, HAMMER) * TS_A_BYTES + (size_t)NB * BN * 256 + 256 + (HAMMER ? TS_HAMMER_BYTES : 0);
    SKM_CUDA(cudaFuncSetAttribute(k_tcs_filter<BN, HAMMER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    constexpr int T = 256 / BN;
    const int64_t nst = (P.ntiles + T - 1) / T;
    const int64_t blocks = std::min<int64_t>(ctx->sm_count, nst);
    if (blocks < 1) return SKM_OK;
    k_tcs_filter<BN, HAMMER><<<(unsigned)blocks, ts_threads(BN, HAMMER), smem, ctx->stream>>>(P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

template <int BN>
int launch_filter(skm_ctx *ctx, const TcsParams &P)
{
    static const char *h = getenv("SKM_TC_HAMMER");
    if (BN == 64 && h && atoi(h) > 0) {
        // micro-benchmark: 8 extra warps read shared memory with LDS.128 for as long as the filter runs
        TcsParams Q = P;
        DevBuf cnt;
        SKM_TRY(cnt.alloc(4 * sizeof(unsigned long long)));
        SKM_CUDA(cudaMemsetAsync(cnt.ptr, 0, 4 * sizeof(unsigned long long), ctx->stream));
        Q.hammer_out = cnt.as<unsigned long long>();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, ctx->stream);
        const int rc = launch_filter_h<64, 8>(ctx, Q);
        cudaEventRecord(e1, ctx->stream);
        unsigned long long hc[4] = {0, 0, 0, 0};
        cudaMemcpyAsync(hc, cnt.ptr, sizeof hc, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        const double sms = (double)std::min<int64_t>(ctx->sm_count, (P.ntiles + 3) / 4);
        fprintf(stderr, "[skm tc hammer] filter %.3f ms with 8 LDS.128 warps per SM: %.1f B/clk/SM of LSU shared-memory reads (%llu warp-loads, %llu cycles)\n",
                ms, hc[1] ? (double)hc[0] * 512.0 / sms / (double)hc[1] : 0.0, hc[0], hc[1]);
        return rc;
    }
    return launch_filter_h<BN, 0>(ctx, P);
}

}  // namespace

// ---------------------------------------------------------------- host side
bool skm_tcs_supported(const skm_ctx *ctx, const skm_dataset *ds, int64_t K)
{
    if (!ds || ds->store_dtype != SKM_F32) return false;
    if (K < 2 || K > 128) return false;
    if (ds->p < 1 || (ds->p + TS_SR - 1) / TS_SR > TS_MAXS) return false;
    if (ds->n < 1 || ds->nnz < 1) return false;
    if (ds->max_col_nnz > 4096) return false;
    if (ctx->smem_optin < 230656) return false;          // the N = 64 kernel: 6 operand stages + 2 centre stages
    return true;
}

int skm_tcs_bn(int64_t K) { return K <= 64 ? 64 : 128; }

size_t skm_tcs_bimg_bytes(int64_t p, int64_t K)
{
    const int64_t S = (p + TS_SR - 1) / TS_SR;
    return (size_t)S * 2 * skm_tcs_bn(K) * 128;
}

void skm_tsb_free(skm_dataset *ds)
{
    skm_big_free(ds->ctx, ds->tsb); ds->tsb = nullptr;
    cudaFree(ds->tsb_ptr); ds->tsb_ptr = nullptr;
    cudaFree(ds->colnorm2); ds->colnorm2 = nullptr;
    cudaFree(ds->xmax_bits); ds->xmax_bits = nullptr;
}

// build the tile/stripe image of the dataset (one-off, on the device)
int skm_tsb_build(skm_dataset *ds)
{
    if (ds->tsb) return SKM_OK;
    skm_ctx *ctx = ds->ctx;
    const int S = (int)((ds->p + TS_SR - 1) / TS_SR);
    const int64_t ntiles = (ds->n + TS_TILE - 1) / TS_TILE;
    const int64_t nblk = ntiles * S;
    auto fail = [&](int rc) { skm_tsb_free(ds); return rc; };
    cudaError_t e;
    if (skm_big_alloc(ctx, (void **)&ds->tsb, sizeof(uint2) * (size_t)std::max<int64_t>(ds->nnz, 1), "tile/stripe image") != SKM_OK) return fail(SKM_ERR_NOMEM);
    if ((e = cudaMalloc((void **)&ds->tsb_ptr, sizeof(int64_t) * (size_t)(nblk + 1))) != cudaSuccess ||
        (e = cudaMalloc((void **)&ds->colnorm2, sizeof(float) * (size_t)ds->n)) != cudaSuccess ||
        (e = cudaMalloc((void **)&ds->xmax_bits, sizeof(int) * 4)) != cudaSuccess) {
        cudaGetLastError();
        skm_set_error("cudaMalloc of the tile/stripe image failed: %s", cudaGetErrorString(e));
        return fail(SKM_ERR_NOMEM);
    }
    DevBuf cnt, tmp;
    if (cnt.alloc(sizeof(int64_t) * (size_t)(nblk + 1)) != SKM_OK) return fail(SKM_ERR_NOMEM);
    if (cudaMemsetAsync(ds->xmax_bits, 0, sizeof(int) * 4, ctx->stream) != cudaSuccess ||
        cudaMemsetAsync(cnt.as<int64_t>() + nblk, 0, sizeof(int64_t), ctx->stream) != cudaSuccess) {
        skm_set_error("memset failed");
        return fail(SKM_ERR_CUDA);
    }
    k_tsb_count<<<(unsigned)ntiles, TS_TILE, 0, ctx->stream>>>(ds->n, S, ds->colptr, ds->rowidx, (const float *)ds->val,
                                                                cnt.as<int64_t>(), ds->xmax_bits, ds->xmax_bits + 1);
    ctx->launches++;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.as<int64_t>(), ds->tsb_ptr, nblk + 1, ctx->stream);
    if (tmp.alloc(tb) != SKM_OK) return fail(SKM_ERR_NOMEM);
    if (cub::DeviceScan::ExclusiveSum(tmp.ptr, tb, cnt.as<int64_t>(), ds->tsb_ptr, nblk + 1, ctx->stream) != cudaSuccess) {
        skm_set_error("scan of the block counts failed");
        return fail(SKM_ERR_CUDA);
    }
    ctx->launches++;
    // sigma: the power of two that brings max |x| below 64 (fp16 then neither overflows on squares nor loses small
    // values to subnormals); fixed for the life of the image, the stored values are already multiplied by it
    int hx[2] = {0, 0};
    if (cudaMemcpyAsync(hx, ds->xmax_bits, sizeof hx, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
        cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
        skm_set_error("tile/stripe image build failed (count pass): %s", cudaGetErrorString(cudaGetLastError()));
        return fail(SKM_ERR_CUDA);
    }
    float xmax;
    memcpy(&xmax, &hx[0], sizeof xmax);
    float sigma = 1.f;
    if (!(xmax < INFINITY)) { skm_set_error("the tensor-core filter needs finite values"); return fail(SKM_ERR_UNSUPPORTED); }
    if (xmax > 0.f) {
        int ex;
        frexpf(xmax, &ex);                                           // xmax < 2^ex
        if (ex > 100 || ex < -100) { skm_set_error("the tensor-core filter needs values within 2^+-100"); return fail(SKM_ERR_UNSUPPORTED); }
        sigma = ldexpf(1.f, 6 - ex);
    }
    ds->tsb_sigma = sigma;
    ds->tsb_max_block = hx[1];
    const size_t sm = (size_t)S * TS_TILE * sizeof(unsigned short);
    k_tsb_fill<<<(unsigned)ntiles, TS_TILE, sm, ctx->stream>>>(ds->n, S, sigma, ds->colptr, ds->rowidx, (const float *)ds->val,
                                                               ds->tsb_ptr, ds->tsb, ds->colnorm2);
    ctx->launches++;
    e = cudaStreamSynchronize(ctx->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { skm_set_error("tile/stripe image build failed: %s", cudaGetErrorString(e)); return fail(SKM_ERR_CUDA); }
    ds->tsb_ntiles = ntiles; ds->tsb_stripes = S;
    ds->device_bytes += (int64_t)sizeof(uint2) * ds->nnz + (int64_t)sizeof(int64_t) * (nblk + 1) + (int64_t)sizeof(float) * ds->n;
    return SKM_OK;
}

int skm_launch_tcs_centres(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const double *ct, const float *cmax,
                           float *scale, void *bimg)
{
    const int BN = skm_tcs_bn(K);
    const int S = ds->tsb_stripes;
    k_tcs_scale<<<1, 1, 0, ctx->stream>>>(ds->tsb_sigma, cmax, scale);
    SKM_CHECK_LAUNCH(ctx);
    const int64_t total = (int64_t)S * 2 * BN * 8;
    k_tcs_centres<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(ds->p, K, BN, S, ct, scale, (uint4 *)bimg);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// Rounding analysis of one score (scaled units, u = 2^-11, eta = 2^-25 the fp16 subnormal half-spacing):
//   x^ = fl16(x), c^ = fl16(c), the table holds -2c^ (exact) and fl16(c^2); products of fp16 values are exact in fp32;
//   |x^ c^ - x c| <= (2u + u^2)|x c| + eta (1 + u)(|x| + |c|) + eta^2,   |fl16(c^2) - c^2| <= u c^2 + eta
//   => per column |s^ - s| <= (4u + 2u^2) sum|x c| + u sum c^2 + m eta (2|x|max + 2|c|max + 2)     (|.|max <= 64)
//   the fp32 accumulation in the tensor core (truncating adds, 2p/16 dependent steps) is budgeted at 0.2 u of
//   sum |terms|;  sum|x c| <= |x| q^1/2, sum c^2 = q  with q the masked squared norm of the centre, bounded in the
//   epilogue.  u_eff = 1.25 u covers all relative parts; abs = m * 2^-25 * 260.
int skm_launch_tcs_filter(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const void *bimg, const float *scale,
                          const float *cmax, int32_t *assign, float *lb, uint32_t *cand, float *lb4, float *dbg_scores)
{
    TcsParams P;
    P.tsb = ds->tsb; P.blk_ptr = ds->tsb_ptr; P.ntiles = ds->tsb_ntiles; P.n = ds->n;
    P.S = ds->tsb_stripes; P.K = (int)K;
    P.bimg = (const unsigned char *)bimg; P.scale = scale; P.colnorm2 = ds->colnorm2;
    (void)cmax;
    const double m = (double)(ds->max_col_nnz > 0 ? ds->max_col_nnz : 1);
    const double u16 = 1.0 / 2048.0, u32 = 5.9604644775390625e-08;
    P.u_eff = (float)(1.25 * u16 + 1.0 / 65536.0);          // + the centre index in the low mantissa bits of the score
    P.abs_unit = (float)(m * 260.0 / 33554432.0);
    P.ga = (float)(1.01 * (m + 5.0) * u32);
    P.sqrt_m = (float)(sqrt(m) * (1.0 + 1e-6));
    P.assign = assign; P.lb = lb; P.cand = cand; P.lb4 = lb4; P.dbg_scores = dbg_scores; P.hammer_out = nullptr;
    if (skm_tcs_bn(K) == 64) return launch_filter<64>(ctx, P);
    return launch_filter<128>(ctx, P);
}

int skm_launch_tcs_resolve(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const float *table_t, const float *cmax,
                           const int32_t *flagged_in, const int *nflag_in, int64_t nflag_host, const uint32_t *cand,
                           const float *lb4, int32_t *assign, float *dist, float *lb, int32_t *flagged_out, int *nflag_out)
{
    SKM_CUDA(cudaMemsetAsync(nflag_out, 0, sizeof(int), ctx->stream));
    if (nflag_host <= 0) return SKM_OK;
    const double u = 5.9604644775390625e-08;
    const double m = (double)(ds->max_col_nnz > 0 ? ds->max_col_nnz : 1);
    ResolveParams P;
    P.colptr = ds->colptr; P.rowidx = ds->rowidx; P.val = (const float *)ds->val; P.p = ds->p; P.K = (int)K;
    P.table_t = table_t;
    P.ga = (float)(1.01 * (m + 5.0) * u);
    P.gb_unit = (float)(2.02 * u * sqrt(m));
    P.ge_unit = (float)(2.1 * u * u * m);
    P.cmax = cmax; P.flagged_in = flagged_in; P.nflag_in = nflag_in; P.cand = cand; P.lb4 = lb4;
    P.assign = assign; P.dist = dist; P.lb = lb; P.flagged_out = flagged_out; P.nflag_out = nflag_out;
    const int64_t blocks = std::min<int64_t>((nflag_host * 32 + 255) / 256, (int64_t)ctx->sm_count * 8);
    k_tcs_resolve<<<(unsigned)blocks, 256, 0, ctx->stream>>>(P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}
