// tcgemm.cu -- the dense contraction of the second pass on the 5th-generation tensor cores.
//
// The reference's dense branch of findClusterAssignments is literally a matrix product:
//     distances(k,:) = nrm2 - 2*(X'*centers(:,k))' + norm(centers(:,k))^2        (private/findClusterAssignments.m:154-163)
// followed by min over k (:168-171).  For K >= 16 the CUDA-core kernel of dense.cu (2*K fp32 instructions per
// matrix element) is issue-bound far below the HBM rate the data streams at, so here the product runs where it
// belongs, as tcgen05.mma with TMEM accumulators and TMA-fed shared-memory operands -- but only as a FILTER, because
// a tf32 product cannot decide a near-tie: per point the epilogue keeps the two best scores, their centres and the
// third-best score; k_dense_verify then evaluates the candidate(s) exactly (sum (x-c)^2 in fp32, the same rounding
// guard as everywhere else) and proves with a bound on the tf32 error that no other centre can win.  Points it
// cannot prove go to the fp64 kernel, as before.  Winners are therefore the reference's, not the tensor cores'.
//
// k_tc_scores<BN>  (one CTA per SM, persistent over 128-point tiles, 6 warps)
//   warp 0      TMA producer: A tile = 128 points x 32 floats of X (K-major: a point's p values are contiguous),
//               B tile = BN centres x 32 floats of C' -- cp.async.bulk.tensor.2d, 128-byte swizzle, 4-stage ring
//               of full/empty mbarriers
//   warp 1      MMA issuer (one elected lane): 4 x tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=BN, K=8) per
//               stage into a double-buffered TMEM accumulator (2 x BN columns); tcgen05.commit releases the stage
//               and, after the last k-block, hands the accumulator to the epilogue
//   warps 2..5  epilogue: thread = point (TMEM lane), tcgen05.ld 32 columns at a time, score = |c|^2 - 2 x.c,
//               running top-3, 24 bytes out per point
// Bound: HBM (4 bytes per matrix element, read once); the MMA work is 2*p*K flops per point, ~0.3 ms per 2e6 points
// at p = 1024, K = 64 against 1.3 ms of HBM time.
#include "common.cuh"
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <algorithm>

namespace {

constexpr int TC_BM = 128;          // points per tile (UMMA M)
constexpr int TC_BK = 32;           // floats per k-block: one 128-byte swizzle row
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 192;

// ---------------------------------------------------------------- PTX helpers
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
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 27)) __trap();          // a lost arrival is a bug, not a hang
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *tm, int c0, int c1, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(s32(bar)) : "memory");
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
// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes
// apart (SBO), version 1 (sm_100), layout type 2 (SWIZZLE_128B); the tile base is 1024-byte aligned
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major): 1
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
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

struct TcParams {
    int64_t n;                 // points
    int     kblocks;           // ceil(p / 32)
    int     K;                 // real centres (<= BN)
    const float *cnorm2;       // [K] |c_k|^2
    int2   *cand;              // [n] best / second-best centre by score
    float4 *score;             // [n] (s1, s2, s3, -): |c|^2 - 2 x.c of the best three
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_scores(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams P)
{
    constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4;
    // two accumulators BN columns apart; the epilogue reads 32 columns at a time, so the second one is padded to 32
    constexpr uint32_t NEED_COLS = BN + ((BN + 31) / 32) * 32;
    constexpr uint32_t TMEM_COLS = (NEED_COLS <= 32) ? 32 : (NEED_COLS <= 64) ? 64 : (NEED_COLS <= 128) ? 128 : (NEED_COLS <= 256) ? 256 : 512;
    static_assert(NEED_COLS <= 512, "TMEM has 512 columns");
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M = 128: a multiple of 16 up to 256");
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    unsigned char *smA = base, *smB = base + TC_STAGES * A_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smB + TC_STAGES * B_BYTES);
    uint64_t *full = bars, *empty = bars + TC_STAGES, *acc_full = bars + 2 * TC_STAGES, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntiles = (P.n + TC_BM - 1) / TC_BM;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {                                             // TMEM allocation: one whole warp
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kb = 0; kb < P.kblocks; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], A_BYTES + B_BYTES);
                    tma_load_2d(smA + stage * A_BYTES, &tmA, kb * TC_BK, (int)(tile * TC_BM), &full[stage]);
                    tma_load_2d(smB + stage * B_BYTES, &tmB, kb * TC_BK, 0, &full[stage]);
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        // instruction descriptor: D = F32 (bit 4), A = B = TF32 (2 at bits 7 and 10), both K-major, N >> 3 at bit 17,
        // M >> 4 at bit 24
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&acc_empty[as], aphase ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < P.kblocks; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = umma_desc_k_sw128(s32(smA + stage * A_BYTES));
                    const uint64_t db = umma_desc_k_sw128(s32(smB + stage * B_BYTES));
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k)      // K = 8 tf32 values (32 bytes) per instruction
                        umma_tf32(tmem_base + as * BN, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit(&empty[stage]);              // the stage is free once these MMAs have read it
                    if (kb == P.kblocks - 1) umma_commit(&acc_full[as]);
                }
                __syncwarp();
                if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    } else {
        // ===== epilogue: thread = point =====
        const int quarter = warp & 3;                            // the TMEM lanes this warp may read
        const float INF = __int_as_float(0x7f800000);
        uint32_t as = 0, aphase = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&acc_full[as], aphase);
            tc_fence_after();
            float b1 = INF, b2 = INF, b3 = INF;
            int i1 = 0, i2 = 0;
            bool bad = false;
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + c0, v);
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const int k = c0 + c;
                    if (k < P.K) {
                        const float s = fmaf(-2.f, __uint_as_float(v[c]), __ldg(P.cnorm2 + k));
                        if (s < b1) { b3 = b2; b2 = b1; i2 = i1; b1 = s; i1 = k; }
                        else if (s < b2) { b3 = b2; b2 = s; i2 = k; }
                        else if (s < b3) b3 = s;
                        else if (!(s == s)) bad = true;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);          // this warp has drained its quarter
            const int64_t j = tile * TC_BM + quarter * 32 + lane;
            if (j < P.n) {
                if (bad) b3 = __int_as_float(0x7fc00000);       // a NaN score poisons the third one: nothing is certified
                P.cand[j] = make_int2(i1, i2);
                P.score[j] = make_float4(b1, b2, b3, 0.f);
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------- exact evaluation of the candidates
struct VerifyParams {
    const float *x;            // [n][p]
    int64_t      p, n;
    int          K;
    const float *ct;           // [K][p] fp32 centres (the values the guard's centre-rounding term refers to)
    const int2  *cand;
    const float4 *score;
    float        eps_tc;       // bound on |score - (|c|^2 - 2 x.c)| / (|x| max_k |c_k|)
    const float *cnorm_max;    // max_k |c_k| (device)
    float        ga, gb_unit, ge_unit;
    const float *cmax;
    int32_t     *assign;
    float       *dist;
    int32_t     *flagged;
    int         *nflag;
};

// one warp per point, lanes over rows (float4), the point stays in registers for the optional second candidate
template <int NV>   // float4 per lane kept in registers: p <= 128 * NV
__global__ void __launch_bounds__(256) k_dense_verify(const VerifyParams P)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t p4 = P.p >> 2;
    const float INF = __int_as_float(0x7f800000);
    for (int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < P.n; j += nwarps) {
        const float4 *xr = reinterpret_cast<const float4 *>(P.x + j * P.p);
        const int2 cd = P.cand[j];
        const float4 sc = P.score[j];
        const float4 *c1 = reinterpret_cast<const float4 *>(P.ct + (int64_t)cd.x * P.p);
        float4 xv[NV];
        float x2 = 0.f, d1 = 0.f, xm = 0.f;
#pragma unroll
        for (int t = 0; t < NV; ++t) {
            const int64_t i = lane + 32 * t;
            xv[t] = i < p4 ? __ldcs(xr + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < NV; ++t) {
            const int64_t i = lane + 32 * t;
            if (i < p4) {
                const float4 c = __ldg(c1 + i), x = xv[t];
                float d;
                d = x.x - c.x; d1 = fmaf(d, d, d1); x2 = fmaf(x.x, x.x, x2);
                d = x.y - c.y; d1 = fmaf(d, d, d1); x2 = fmaf(x.y, x.y, x2);
                d = x.z - c.z; d1 = fmaf(d, d, d1); x2 = fmaf(x.z, x.z, x2);
                d = x.w - c.w; d1 = fmaf(d, d, d1); x2 = fmaf(x.w, x.w, x2);
                xm = fmaxf(xm, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            x2 += __shfl_xor_sync(0xffffffffu, x2, o);
            xm = fmaxf(xm, __shfl_xor_sync(0xffffffffu, xm, o));
        }
        // rounding guard of an fp32 sum of p squares, centre and point both rounded to fp32
        const float cm = *P.cmax + xm;
        const float gb = P.gb_unit * cm, ge = P.ge_unit * cm * cm + 1e-37f;
        auto guard = [&](float s) { return P.ga * s + gb * sqrtf(s) + ge; };
        // what the tensor-core scores can be off by, and the fp32 rounding of |x|^2 itself
        const float Etc = P.eps_tc * sqrtf(x2) * __ldg(P.cnorm_max) * 1.0001f + P.ga * x2;
        const float hi1 = d1 + guard(d1);                       // upper bound on the true squared distance to cand 1
        bool certified = false;
        int win = cd.x;
        float dwin = d1;
        const bool others_out = (x2 + sc.z - Etc) > hi1;       // no centre beyond the best two can win (false for NaN)
        if (P.K == 1) certified = d1 < INF;
        else if (others_out) {
            if ((x2 + sc.y - Etc) > hi1) certified = true;      // the runner-up cannot win either
            else {
                const float4 *c2 = reinterpret_cast<const float4 *>(P.ct + (int64_t)cd.y * P.p);
                float d2 = 0.f;
#pragma unroll
                for (int t = 0; t < NV; ++t) {
                    const int64_t i = lane + 32 * t;
                    if (i < p4) {
                        const float4 c = __ldg(c2 + i), x = xv[t];
                        float d;
                        d = x.x - c.x; d2 = fmaf(d, d, d2);
                        d = x.y - c.y; d2 = fmaf(d, d, d2);
                        d = x.z - c.z; d2 = fmaf(d, d, d2);
                        d = x.w - c.w; d2 = fmaf(d, d, d2);
                    }
                }
#pragma unroll
                for (int o = 16; o; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
                const float E = guard(d1) + guard(d2);
                if (d2 - d1 > E) certified = true;
                else if (d1 - d2 > E) { certified = true; win = cd.y; dwin = d2; }
            }
        }
        if (lane == 0) {
            P.assign[j] = win;
            P.dist[j] = sqrtf(dwin);
            if (!certified) {
                const int slot = atomicAdd(P.nflag, 1);
                P.flagged[slot] = (int32_t)j;
            }
        }
    }
}

// [K][p] fp32 centres, their tf32-rounded copy for the tensor cores, |c_k|^2 and max_k |c_k|
__global__ void k_tc_prep_centers(int64_t p, int64_t K, const double *__restrict__ ct /* row-major [p+1][K] */,
                                  float *__restrict__ c32, float *__restrict__ ctf, float *__restrict__ cnorm2,
                                  float *__restrict__ cnorm_max)
{
    const int64_t k = blockIdx.x;
    double s = 0.0;
    for (int64_t r = threadIdx.x; r < p; r += blockDim.x) {
        const double c = ct[r * K + k];
        const float f = (float)c;
        c32[k * p + r] = f;
        // round to tf32 (10 explicit mantissa bits) to nearest: the tensor core then reads the value exactly
        uint32_t b = __float_as_uint(f);
        if ((b & 0x7f800000u) != 0x7f800000u) b = (b + 0x00000fffu + ((b >> 13) & 1u)) & 0xffffe000u;
        ctf[k * p + r] = __uint_as_float(b);
        s += c * c;
    }
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        cnorm2[k] = (float)t;
        const float nm = __double2float_ru(sqrt(t) * (1.0 + 1e-7));
        atomicMax(reinterpret_cast<int *>(cnorm_max), __float_as_int(nm == nm ? nm : __int_as_float(0x7f800000)));
    }
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

encode_tiled_fn get_encode()
{
    static encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (encode_tiled_fn)p;
        else cudaGetLastError();
    }
    return fn;
}

// 2-D fp32 tensor [rows][cols] (cols contiguous), box [box_rows][32 floats], 128-byte swizzle, zero fill out of bounds
int make_map(CUtensorMap *tm, const float *ptr, int64_t rows, int64_t cols, int box_rows)
{
    encode_tiled_fn enc = get_encode();
    if (!enc) { skm_set_error("cuTensorMapEncodeTiled is not available from this driver"); return SKM_ERR_UNSUPPORTED; }
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { skm_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return SKM_ERR_CUDA; }
    return SKM_OK;
}

template <int BN>
int launch_scores(skm_ctx *ctx, const CUtensorMap &ta, const CUtensorMap &tb, const TcParams &P)
{
    const size_t smem = 1024 + (size_t)TC_STAGES * (TC_BM + BN) * TC_BK * 4 + 256;
    auto kern = k_tc_scores<BN>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = (P.n + TC_BM - 1) / TC_BM;
    const int64_t blocks = std::min<int64_t>(ntiles, ctx->sm_count);
    kern<<<(unsigned)blocks, TC_THREADS, smem, ctx->stream>>>(ta, tb, P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}


// ---------------------------------------------------------------- the DCT sketch as a 3xTF32 product
// Y = M X with M = T diag(d) (1+2eps) (csrc/dct.cu; kmeans_sparsified.m:256-258,292-295) needs fp32 accuracy, which one
// tf32 product does not give.  Both operands are split into tf32-exact halves, x = x_hi + x_lo, and three products
// are accumulated into the same TMEM accumulator:  M_hi X_hi + M_hi X_lo + M_lo X_hi  (the dropped M_lo X_lo term and
// the truncation of the *_lo halves are ~2^-21 relative).  Same warp roles as k_tc_scores; a stage holds the four
// operand tiles (X_hi, X_lo: 128 points x 32; M_hi, M_lo: BN output rows x 32), twelve tcgen05.mma per stage; the
// epilogue stores the fp32 tile (thread = point, 32 output rows at a time).  Tiles: (point tile, output-row tile),
// output-row tile fastest so the X tiles of a point tile are re-read from L2.
constexpr int DCT_STAGES = 3;

struct DctParams {
    int64_t n;                 // points in the chunk
    int     p;                 // output rows (= input rows of the square transform)
    int     kblocks;           // ceil(p_pad / 32)
    int     ntn;               // output-row tiles
    float  *y;                 // [n][ldy]
    int64_t ldy;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_tc_dct(const __grid_constant__ CUtensorMap tmXh, const __grid_constant__ CUtensorMap tmXl,
         const __grid_constant__ CUtensorMap tmMh, const __grid_constant__ CUtensorMap tmMl, const DctParams P)
{
    constexpr uint32_t A_BYTES = TC_BM * TC_BK * 4, B_BYTES = BN * TC_BK * 4, STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr uint32_t NEED_COLS = BN + ((BN + 31) / 32) * 32;
    constexpr uint32_t TMEM_COLS = (NEED_COLS <= 32) ? 32 : (NEED_COLS <= 64) ? 64 : (NEED_COLS <= 128) ? 128 : (NEED_COLS <= 256) ? 256 : 512;
    static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "UMMA N for M = 128: a multiple of 16 up to 256");
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + DCT_STAGES * STAGE_BYTES);
    uint64_t *full = bars, *empty = bars + DCT_STAGES, *acc_full = bars + 2 * DCT_STAGES, *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t ntm = (P.n + TC_BM - 1) / TC_BM, ntiles = ntm * P.ntn;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmXh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmXl)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmMh)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmMl)) : "memory");
        for (int s = 0; s < DCT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = (int)((tile / P.ntn) * TC_BM), n0 = (int)((tile % P.ntn) * BN);
                for (int kb = 0; kb < P.kblocks; ++kb) {
                    unsigned char *st = base + stage * STAGE_BYTES;
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(st, &tmXh, kb * TC_BK, m0, &full[stage]);
                    tma_load_2d(st + A_BYTES, &tmXl, kb * TC_BK, m0, &full[stage]);
                    tma_load_2d(st + 2 * A_BYTES, &tmMh, kb * TC_BK, n0, &full[stage]);
                    tma_load_2d(st + 2 * A_BYTES + B_BYTES, &tmMl, kb * TC_BK, n0, &full[stage]);
                    if (++stage == DCT_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
        uint32_t stage = 0, phase = 0, as = 0, aphase = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&acc_empty[as], aphase ^ 1);
            tc_fence_after();
            for (int kb = 0; kb < P.kblocks; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    unsigned char *st = base + stage * STAGE_BYTES;
                    const uint64_t xh = umma_desc_k_sw128(s32(st)), xl = umma_desc_k_sw128(s32(st + A_BYTES));
                    const uint64_t mh = umma_desc_k_sw128(s32(st + 2 * A_BYTES)), ml = umma_desc_k_sw128(s32(st + 2 * A_BYTES + B_BYTES));
                    const uint32_t d = tmem_base + as * BN;
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t o = (uint64_t)(k * 2);
                        umma_tf32(d, xh + o, mh + o, idesc, (kb | k) != 0);      // small terms need no special order:
                        umma_tf32(d, xl + o, mh + o, idesc, 1);                   // the accumulator is fp32
                        umma_tf32(d, xh + o, ml + o, idesc, 1);
                    }
                    umma_commit(&empty[stage]);
                    if (kb == P.kblocks - 1) umma_commit(&acc_full[as]);
                }
                __syncwarp();
                if (++stage == DCT_STAGES) { stage = 0; phase ^= 1; }
            }
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    } else {
        const int quarter = warp & 3;
        const bool vec = (P.ldy & 3) == 0;
        uint32_t as = 0, aphase = 0;
        for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int64_t j = (tile / P.ntn) * TC_BM + quarter * 32 + lane;
            const int n0 = (int)((tile % P.ntn) * BN);
            mbar_wait(&acc_full[as], aphase);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + c0, v);
                if (j < P.n) {
                    float *dst = P.y + j * P.ldy + n0 + c0;
                    if (vec && n0 + c0 + 32 <= P.p) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            reinterpret_cast<float4 *>(dst)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                             __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    } else {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            if (n0 + c0 + c < P.p) dst[c] = __uint_as_float(v[c]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[as]);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

__device__ __forceinline__ float tf32_round(float f)
{
    uint32_t b = __float_as_uint(f);
    if ((b & 0x7f800000u) != 0x7f800000u) b = (b + 0x00000fffu + ((b >> 13) & 1u)) & 0xffffe000u;
    return __uint_as_float(b);
}

// x (fp32 or fp64, [n][p]) -> tf32-exact halves hi + lo = fl32(x * scale) in a row-padded layout [n][p_pad]
template <typename T>
__global__ void k_cast_split(int64_t n, int64_t p, int64_t p_pad, const T *__restrict__ x, double scale,
                             float *__restrict__ hi, float *__restrict__ lo)
{
    const int64_t total = n * p_pad, stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t j = i / p_pad, r = i - j * p_pad;
        float v = 0.f;
        if (r < p) v = scale == 1.0 ? (float)x[j * p + r] : (float)((double)x[j * p + r] * scale);
        const float h = tf32_round(v);
        hi[i] = h;
        lo[i] = v - h;                                          // exact: h is v rounded to fewer bits
    }
}

template <int BN>
int launch_dct(skm_ctx *ctx, const CUtensorMap &xh, const CUtensorMap &xl, const CUtensorMap &mh, const CUtensorMap &ml, const DctParams &P)
{
    const size_t smem = 1024 + (size_t)DCT_STAGES * (2 * TC_BM + 2 * BN) * TC_BK * 4 + 256;
    auto kern = k_tc_dct<BN>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t ntiles = ((P.n + TC_BM - 1) / TC_BM) * P.ntn;
    const int64_t blocks = std::min<int64_t>(ntiles, ctx->sm_count);
    kern<<<(unsigned)blocks, TC_THREADS, smem, ctx->stream>>>(xh, xl, mh, ml, P);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

}  // namespace

bool skm_tc_dense_usable(int64_t p, int64_t K)
{
    static const bool off = getenv("SKM_NO_TC") != nullptr;
    // TMA needs 16-byte row strides; below K = 16 the CUDA-core kernel is already HBM-bound
    return !off && K >= 16 && K <= 256 && (p % 4) == 0 && p >= 32 && p <= 2048;
}

size_t skm_tc_scratch_bytes(int64_t p, int64_t K, int64_t nc)
{
    return sizeof(float) * (size_t)(2 * p * K + K + 64) + sizeof(int2) * (size_t)nc + sizeof(float4) * (size_t)nc + 1024;
}

// centres -> [K][p] fp32 + tf32 copy + norms, once per call of skm_second_pass
int skm_launch_tc_prep(skm_ctx *ctx, int64_t p, int64_t K, const double *ct, void *scratch)
{
    float *c32 = (float *)scratch, *ctf = c32 + p * K, *cn2 = ctf + p * K, *cnmax = cn2 + K;
    SKM_CUDA(cudaMemsetAsync(cnmax, 0, sizeof(float) * 4, ctx->stream));
    k_tc_prep_centers<<<(unsigned)K, 256, 0, ctx->stream>>>(p, K, ct, c32, ctf, cn2, cnmax);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// filter (tensor cores) + exact evaluation of the candidates for one resident chunk x32 = [nc][p]
int skm_launch_dense_assign_tc(skm_ctx *ctx, int64_t p, int64_t nc, int64_t K, const float *x32, void *scratch,
                               const float *cmax, int32_t *assign, float *dist, int32_t *flagged, int *nflag)
{
    float *c32 = (float *)scratch, *ctf = c32 + p * K, *cn2 = ctf + p * K, *cnmax = cn2 + K;
    unsigned char *tail = (unsigned char *)(cnmax + 64);
    tail = (unsigned char *)(((uintptr_t)tail + 255) & ~(uintptr_t)255);
    int2 *cand = (int2 *)tail;
    float4 *score = (float4 *)(((uintptr_t)(cand + nc) + 255) & ~(uintptr_t)255);
    const int BN = (int)((K + 15) / 16 * 16);
    CUtensorMap ta, tb;
    SKM_TRY(make_map(&ta, x32, nc, p, TC_BM));
    TcParams P;
    P.n = nc; P.kblocks = (int)((p + TC_BK - 1) / TC_BK); P.K = (int)K; P.cnorm2 = cn2; P.cand = cand; P.score = score;
    int rc;
#define SKM_TC_CASE(bn) case bn: SKM_TRY(make_map(&tb, ctf, K, p, bn)); rc = launch_scores<bn>(ctx, ta, tb, P); break;
    switch (BN) {
        SKM_TC_CASE(16) SKM_TC_CASE(32) SKM_TC_CASE(48) SKM_TC_CASE(64) SKM_TC_CASE(80) SKM_TC_CASE(96) SKM_TC_CASE(112)
        SKM_TC_CASE(128) SKM_TC_CASE(144) SKM_TC_CASE(160) SKM_TC_CASE(176) SKM_TC_CASE(192) SKM_TC_CASE(208) SKM_TC_CASE(224)
        SKM_TC_CASE(240) SKM_TC_CASE(256)
        default: skm_set_error("dense_assign_tc: K = %lld out of range", (long long)K); return SKM_ERR_UNSUPPORTED;
    }
#undef SKM_TC_CASE
    if (rc != SKM_OK) return rc;
    const double u = 5.9604644775390625e-08, m = (double)p;
    VerifyParams V;
    V.x = x32; V.p = p; V.n = nc; V.K = (int)K; V.ct = c32; V.cand = cand; V.score = score;
    // tf32 operands: the centres are rounded to tf32 here (2^-11), the hardware converts the points (at worst a
    // truncation, 2^-10), the products are accumulated in fp32: |x~.c~ - x.c| <= 2^-9 |x||c| leaves a factor 2.6
    // for the accumulation; the score doubles it
    V.eps_tc = (float)(2.0 * 0.001953125);
    V.cnorm_max = cnmax;
    V.ga = (float)(1.01 * (m + 5.0) * u);
    V.gb_unit = (float)(2.02 * u * sqrt(m));
    V.ge_unit = (float)(2.1 * u * u * m);
    V.cmax = cmax; V.assign = assign; V.dist = dist; V.flagged = flagged; V.nflag = nflag;
    const int64_t blocks = std::min<int64_t>((nc * 32 + 255) / 256, (int64_t)ctx->sm_count * 8);
    const int64_t p4 = p / 4;
    if (p4 <= 32 * 2)       k_dense_verify<2><<<(unsigned)blocks, 256, 0, ctx->stream>>>(V);
    else if (p4 <= 32 * 4)  k_dense_verify<4><<<(unsigned)blocks, 256, 0, ctx->stream>>>(V);
    else if (p4 <= 32 * 8)  k_dense_verify<8><<<(unsigned)blocks, 256, 0, ctx->stream>>>(V);
    else                    k_dense_verify<16><<<(unsigned)blocks, 256, 0, ctx->stream>>>(V);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// ---- DCT sketch on the tensor cores (called from dct.cu) -----------------------------------------------------
// split + pad a raw chunk (fp32 or fp64 [nc][p]) into the tf32-exact halves [nc][p_pad]
int skm_launch_cast_split(skm_ctx *ctx, int64_t nc, int64_t p, int64_t p_pad, const void *x, int x_type, double scale,
                          float *hi, float *lo)
{
    if (nc == 0) return SKM_OK;
    const int64_t blocks = std::min<int64_t>((nc * p_pad + 255) / 256, (int64_t)ctx->sm_count * 32);
    if (x_type == SKM_F32) k_cast_split<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(nc, p, p_pad, (const float *)x, scale, hi, lo);
    else k_cast_split<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(nc, p, p_pad, (const double *)x, scale, hi, lo);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// y[nc][ldy] (first p entries of a row) = sum_i M[k][i] x[j][i] with both operands given as tf32-exact halves:
// xh/xl [nc][p_pad], mh/ml [p][p_pad] (row k of M contiguous)
int skm_launch_tc_dct(skm_ctx *ctx, int64_t p, int64_t p_pad, int64_t nc, const float *xh, const float *xl,
                      const float *mh, const float *ml, float *y, int64_t ldy)
{
    if (nc == 0) return SKM_OK;
    CUtensorMap txh, txl, tmh, tml;
    SKM_TRY(make_map(&txh, xh, nc, p_pad, TC_BM));
    SKM_TRY(make_map(&txl, xl, nc, p_pad, TC_BM));
    DctParams P;
    P.n = nc; P.p = (int)p; P.kblocks = (int)((p_pad + TC_BK - 1) / TC_BK); P.y = y; P.ldy = ldy;
    if (p <= 32) {
        SKM_TRY(make_map(&tmh, mh, p, p_pad, 32)); SKM_TRY(make_map(&tml, ml, p, p_pad, 32));
        P.ntn = 1;
        return launch_dct<32>(ctx, txh, txl, tmh, tml, P);
    }
    if (p <= 64) {
        SKM_TRY(make_map(&tmh, mh, p, p_pad, 64)); SKM_TRY(make_map(&tml, ml, p, p_pad, 64));
        P.ntn = 1;
        return launch_dct<64>(ctx, txh, txl, tmh, tml, P);
    }
    if (p % 112 == 0) {                                      // p = 784 (MNIST): seven exact tiles
        SKM_TRY(make_map(&tmh, mh, p, p_pad, 112)); SKM_TRY(make_map(&tml, ml, p, p_pad, 112));
        P.ntn = (int)(p / 112);
        return launch_dct<112>(ctx, txh, txl, tmh, tml, P);
    }
    SKM_TRY(make_map(&tmh, mh, p, p_pad, 128)); SKM_TRY(make_map(&tml, ml, p, p_pad, 128));
    P.ntn = (int)((p + 127) / 128);
    return launch_dct<128>(ctx, txh, txl, tmh, tml, P);
}
