// k64_probe.cu -- micro-benchmarks behind DESIGN.md section 4 "K = 64": which resource bounds the
// fused masked-distance + argmin pass when every stored entry needs 64 centre values?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o k64_probe k64_probe.cu
//   ./k64_probe [n_columns] [rows_mode]      rows_mode 0 = uniform random rows (stored order),
//                                            1 = rows arranged conflict-free per quarter-warp
//
// Shape: p = 1024 rows, m = 52 stored entries per column (26 int4 pairs), SELL-32 image like the
// library's (lane l's pair t at sell[(slice*W2 + t)*32 + l]).  Variants:
//   f32x16   fp32 table, 16 centres per pass, 4 passes, scalar FADD+FFMA   (shape of k_assign_fast<16> x4)
//   f32x16p  same, packed FADD2 + FFMA2 (sm_100 f32x2)
//   f32x32p  fp32 table, 32 centres per pass, 2 passes, packed math
//   h16x64   fp16 table, all 64 centres in ONE pass, mixed-precision FHADD (f16 operand, fp32 result, no
//            unpack instruction) + packed FFMA2 accumulation in fp32: a FILTER (table rounded to 11 bits)
//   h16x32   fp16 table, 32 centres per pass, 2 passes
//   atoms    shared-memory atomics scatter (u64 add + u32 add per stored entry into [784 x 10] bins): the
//            cost a K2 epilogue fused into K1 would add at config 2
// Reported: ms per full assignment pass over n columns (all passes), and the implied fraction of the
// HBM roofline for 416 algorithmic bytes per column.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static const int P = 1024, W2 = 26;

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

__global__ void k_gen(int4 *sell, int64_t nslices, int mode)
{
    const int64_t total = nslices * W2 * 32;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int lane = (int)(i & 31);
        uint32_t h0 = hash32((uint32_t)(i * 2 + 1)), h1 = hash32((uint32_t)(i * 2 + 2) ^ 0x9e3779b9U);
        int r0 = h0 & (P - 1), r1 = h1 & (P - 1);
        if (mode == 1) { r0 = (r0 & ~7) | (lane & 7); r1 = (r1 & ~7) | ((lane + 3) & 7); }
        const float v0 = ((hash32(h0) >> 8) * (1.0f / 8388608.0f) - 1.0f) * 20.f;
        const float v1 = ((hash32(h1) >> 8) * (1.0f / 8388608.0f) - 1.0f) * 20.f;
        sell[i] = make_int4(r0, __float_as_int(v0), r1, __float_as_int(v1));
    }
}

// ---------------------------------------------------------------- fp32 table variants
template <int KC, bool PACKED>
__device__ __forceinline__ void step_f32(float (&acc)[KC], const float *tab, int ks, int r, float x)
{
    const float4 *row = reinterpret_cast<const float4 *>(tab + (size_t)r * ks);
    if (!PACKED) {
#pragma unroll
        for (int c = 0; c < KC / 4; ++c) {
            const float4 v = row[c];
            float d;
            d = x - v.x; acc[4 * c + 0] = fmaf(d, d, acc[4 * c + 0]);
            d = x - v.y; acc[4 * c + 1] = fmaf(d, d, acc[4 * c + 1]);
            d = x - v.z; acc[4 * c + 2] = fmaf(d, d, acc[4 * c + 2]);
            d = x - v.w; acc[4 * c + 3] = fmaf(d, d, acc[4 * c + 3]);
        }
    } else {
        unsigned long long xx;
        asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
#pragma unroll
        for (int c = 0; c < KC / 4; ++c) {
            const float4 v = row[c];
            unsigned long long v01, v23, d01, d23, a01, a23;
            asm("mov.b64 %0, {%1, %2};" : "=l"(v01) : "f"(v.x), "f"(v.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(v23) : "f"(v.z), "f"(v.w));
            asm("mov.b64 %0, {%1, %2};" : "=l"(a01) : "f"(acc[4 * c + 0]), "f"(acc[4 * c + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(a23) : "f"(acc[4 * c + 2]), "f"(acc[4 * c + 3]));
            asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(xx), "l"(v01));
            asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(xx), "l"(v23));
            asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(a01) : "l"(d01), "l"(a01));
            asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(a23) : "l"(d23), "l"(a23));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[4 * c + 0]), "=f"(acc[4 * c + 1]) : "l"(a01));
            asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[4 * c + 2]), "=f"(acc[4 * c + 3]) : "l"(a23));
        }
    }
}

template <int KC, bool PACKED, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
probe_f32(const int4 *__restrict__ sell, int64_t nslices, const float *__restrict__ table, int ks, int first,
          float *__restrict__ best, int *__restrict__ arg, int k0)
{
    extern __shared__ __align__(16) float s_tab[];
    for (int i = threadIdx.x; i < (P + 1) * ks; i += THREADS) s_tab[i] = table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * THREADS) >> 5;
    for (int64_t slice = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5; slice < nslices; slice += warps_total) {
        const int4 *src = sell + slice * W2 * 32 + lane;
        float acc[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) acc[k] = 0.f;
        for (int t = 0; t + 2 <= W2; t += 2) {
            const int4 q0 = __ldcs(src + t * 32), q1 = __ldcs(src + (t + 1) * 32);
            step_f32<KC, PACKED>(acc, s_tab, ks, q0.x, __int_as_float(q0.y));
            step_f32<KC, PACKED>(acc, s_tab, ks, q0.z, __int_as_float(q0.w));
            step_f32<KC, PACKED>(acc, s_tab, ks, q1.x, __int_as_float(q1.y));
            step_f32<KC, PACKED>(acc, s_tab, ks, q1.z, __int_as_float(q1.w));
        }
        float b = acc[0]; int bi = k0;
#pragma unroll
        for (int k = 1; k < KC; ++k) if (acc[k] < b) { b = acc[k]; bi = k0 + k; }
        const int64_t j = slice * 32 + lane;
        if (!first) { const float ob = best[j]; if (ob <= b) { b = ob; bi = arg[j]; } }
        best[j] = b; arg[j] = bi;
    }
}

// ---------------------------------------------------------------- fp16 table variants (filter)
template <int KC>
__device__ __forceinline__ void step_h16(unsigned long long (&acc2)[KC / 2], const unsigned char *tab, int row_bytes, int r, float x)
{
    const uint4 *row = reinterpret_cast<const uint4 *>(tab + (size_t)r * row_bytes);
#pragma unroll
    for (int c = 0; c < KC / 8; ++c) {
        const uint4 q = row[c];
        const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned short lo, hi;
            asm("mov.b32 {%0, %1}, %2;" : "=h"(lo), "=h"(hi) : "r"(w[i]));
            float d0, d1;
            asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(d0) : "h"(lo), "f"(x));      // FHADD: fp16 operand, fp32 result
            asm("sub.rn.f32.f16 %0, %1, %2;" : "=f"(d1) : "h"(hi), "f"(x));
            unsigned long long dd;
            asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d0), "f"(d1));
            asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc2[4 * c + i]) : "l"(dd));  // FFMA2
        }
    }
}

template <int KC, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
probe_h16(const int4 *__restrict__ sell, int64_t nslices, const unsigned char *__restrict__ table, int row_bytes, int first,
          float *__restrict__ best, int *__restrict__ arg, int k0)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    {
        const uint4 *g = reinterpret_cast<const uint4 *>(table);
        uint4 *s = reinterpret_cast<uint4 *>(s_raw);
        const int total = (P + 1) * row_bytes / 16;
        for (int i = threadIdx.x; i < total; i += THREADS) s[i] = g[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * THREADS) >> 5;
    for (int64_t slice = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5; slice < nslices; slice += warps_total) {
        const int4 *src = sell + slice * W2 * 32 + lane;
        unsigned long long acc2[KC / 2];
#pragma unroll
        for (int k = 0; k < KC / 2; ++k) acc2[k] = 0ULL;
        for (int t = 0; t + 2 <= W2; t += 2) {
            const int4 q0 = __ldcs(src + t * 32), q1 = __ldcs(src + (t + 1) * 32);
            step_h16<KC>(acc2, s_raw, row_bytes, q0.x, __int_as_float(q0.y));
            step_h16<KC>(acc2, s_raw, row_bytes, q0.z, __int_as_float(q0.w));
            step_h16<KC>(acc2, s_raw, row_bytes, q1.x, __int_as_float(q1.y));
            step_h16<KC>(acc2, s_raw, row_bytes, q1.z, __int_as_float(q1.w));
        }
        float b = __int_as_float(0x7f800000); int bi = k0;
#pragma unroll
        for (int k = 0; k < KC / 2; ++k) {
            float a0, a1;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc2[k]));
            if (a0 < b) { b = a0; bi = k0 + 2 * k; }
            if (a1 < b) { b = a1; bi = k0 + 2 * k + 1; }
        }
        const int64_t j = slice * 32 + lane;
        if (!first) { const float ob = best[j]; if (ob <= b) { b = ob; bi = arg[j]; } }
        best[j] = b; arg[j] = bi;
    }
}

// ---------------------------------------------------------------- shared-memory atomics scatter (fused K2 epilogue cost)
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
probe_atoms(const int4 *__restrict__ sell, int64_t nslices, int pbins, int kbins, unsigned long long *__restrict__ out)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned long long *binS = reinterpret_cast<unsigned long long *>(s_raw);
    unsigned int *binN = reinterpret_cast<unsigned int *>(binS + (size_t)pbins * kbins);
    for (int i = threadIdx.x; i < pbins * kbins; i += THREADS) { binS[i] = 0ULL; binN[i] = 0U; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warps_total = ((int64_t)gridDim.x * THREADS) >> 5;
    for (int64_t slice = ((int64_t)blockIdx.x * THREADS + threadIdx.x) >> 5; slice < nslices; slice += warps_total) {
        const int4 *src = sell + slice * W2 * 32 + lane;
        const int a = (int)(hash32((uint32_t)(slice * 32 + lane)) % (uint32_t)kbins);     // the column's cluster
        for (int t = 0; t < W2; ++t) {
            const int4 q = __ldcs(src + t * 32);
            const int r0 = q.x % pbins, r1 = q.z % pbins;
            const long long f0 = (long long)llrintf(__int_as_float(q.y) * 1048576.f), f1 = (long long)llrintf(__int_as_float(q.w) * 1048576.f);
            atomicAdd(&binS[r0 * kbins + a], (unsigned long long)f0); atomicAdd(&binN[r0 * kbins + a], 1U);
            atomicAdd(&binS[r1 * kbins + a], (unsigned long long)f1); atomicAdd(&binN[r1 * kbins + a], 1U);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < pbins * kbins; i += THREADS) if (binN[i]) atomicAdd(&out[i], binS[i]);
}

// ---------------------------------------------------------------- host
struct Timer {
    cudaEvent_t a, b;
    Timer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    void start() { cudaEventRecord(a); }
    float stop() { cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); return ms; }
};

static bool g_once = false;       // profile mode: every variant runs exactly once (ncu captures)
template <typename F> static float bench(F &&f, int reps = 5)
{
    if (g_once) { Timer t; t.start(); f(); float ms = t.stop(); CK(cudaGetLastError()); return ms; }
    f(); CK(cudaDeviceSynchronize());
    Timer t; t.start();
    for (int i = 0; i < reps; ++i) f();
    float ms = t.stop() / reps;
    CK(cudaGetLastError());
    return ms;
}

template <typename K> static int blocks_for(K kern, int threads, size_t smem, int sms)
{
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, threads, smem));
    if (per < 1) { printf("kernel does not fit (smem %zu)\n", smem); return 0; }
    return per * sms;
}

int main(int argc, char **argv)
{
    const int64_t n = argc > 1 ? atoll(argv[1]) : 12500000;
    const int mode = argc > 2 ? atoi(argv[2]) : 0;
    g_once = argc > 3 && atoi(argv[3]) != 0;
    const int64_t nslices = n / 32;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int4 *sell; CK(cudaMalloc(&sell, sizeof(int4) * nslices * W2 * 32));
    k_gen<<<sms * 8, 256>>>(sell, nslices, mode);
    CK(cudaDeviceSynchronize());
    float *best, *best_ref; int *arg, *arg_ref;
    CK(cudaMalloc(&best, 4 * n)); CK(cudaMalloc(&arg, 4 * n)); CK(cudaMalloc(&best_ref, 4 * n)); CK(cudaMalloc(&arg_ref, 4 * n));
    // centre table c[r][k], N(0,1)-like, scaled like the data
    std::vector<float> c((size_t)P * 64);
    srand(1);
    for (auto &v : c) v = ((rand() / (float)RAND_MAX) * 2.f - 1.f) * 6.f;
    const double alg = (double)n * 416.0;
    printf("# n=%lld columns, 52 entries each, p=%d, K=64, rows_mode=%d, %d SMs\n", (long long)n, P, mode, sms);
    printf("%-10s %10s %12s %10s\n", "variant", "ms/pass", "alg GB/s", "notes");

    auto fill_f32 = [&](int kc, int ks, std::vector<float> &h) {       // [chunk][(P+1)][ks]
        const int nch = 64 / kc;
        h.assign((size_t)nch * (P + 1) * ks, 0.f);
        for (int ch = 0; ch < nch; ++ch)
            for (int r = 0; r < P; ++r)
                for (int k = 0; k < kc; ++k) h[((size_t)ch * (P + 1) + r) * ks + k] = c[(size_t)r * 64 + ch * kc + k];
    };
    auto run_f32 = [&](const char *name, int kc, auto kern, int threads) {
        const int ks = 4 * ((kc / 4) | 1);
        std::vector<float> h; fill_f32(kc, ks, h);
        float *d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(P + 1) * ks * 4;
        const int blocks = blocks_for(kern, threads, smem, sms);
        if (!blocks) return;
        const int nch = 64 / kc;
        const float ms = bench([&] {
            for (int ch = 0; ch < nch; ++ch)
                kern<<<blocks, threads, smem>>>(sell, nslices, d + (size_t)ch * (P + 1) * ks, ks, ch == 0, best, arg, ch * kc);
        });
        printf("%-10s %10.3f %12.1f   %d passes x %d centres, %zu B smem, %d CTAs/SM x %d thr\n", name, ms, alg / ms * 1e-6, nch, kc, smem,
               blocks / sms, threads);
        CK(cudaFree(d));
    };
    run_f32("f32x16", 16, probe_f32<16, false, 512, 2>, 512);
    CK(cudaMemcpy(best_ref, best, 4 * n, cudaMemcpyDeviceToDevice)); CK(cudaMemcpy(arg_ref, arg, 4 * n, cudaMemcpyDeviceToDevice));
    run_f32("f32x16p", 16, probe_f32<16, true, 512, 2>, 512);
    {   // packed math must give the same bits
        std::vector<float> a(1 << 16), b(1 << 16);
        CK(cudaMemcpy(a.data(), best, a.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), best_ref, b.size() * 4, cudaMemcpyDeviceToHost));
        size_t diff = 0; for (size_t i = 0; i < a.size(); ++i) diff += a[i] != b[i];
        printf("#   f32x16p vs f32x16: %zu of %zu best values differ\n", diff, a.size());
    }
    run_f32("f32x32p", 32, probe_f32<32, true, 512, 1>, 512);
    run_f32("f32x32", 32, probe_f32<32, false, 512, 1>, 512);

    auto run_h16 = [&](const char *name, int kc, auto kern, int threads) {
        const int row_bytes = kc * 2 + 16;                           // odd number of 16-byte chunks
        const int nch = 64 / kc;
        std::vector<unsigned char> h((size_t)nch * (P + 1) * row_bytes, 0);
        for (int ch = 0; ch < nch; ++ch)
            for (int r = 0; r < P; ++r)
                for (int k = 0; k < kc; ++k) {
                    __half v = __float2half_rn(c[(size_t)r * 64 + ch * kc + k]);
                    memcpy(&h[((size_t)ch * (P + 1) + r) * row_bytes + 2 * k], &v, 2);
                }
        unsigned char *d; CK(cudaMalloc(&d, h.size())); CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(P + 1) * row_bytes;
        const int blocks = blocks_for(kern, threads, smem, sms);
        if (!blocks) return;
        const float ms = bench([&] {
            for (int ch = 0; ch < nch; ++ch)
                kern<<<blocks, threads, smem>>>(sell, nslices, d + (size_t)ch * (P + 1) * row_bytes, row_bytes, ch == 0, best, arg, ch * kc);
        });
        printf("%-10s %10.3f %12.1f   %d passes x %d centres (fp16 table), %zu B smem, %d CTAs/SM x %d thr\n", name, ms, alg / ms * 1e-6,
               nch, kc, smem, blocks / sms, threads);
        std::vector<int> a(1 << 16), b(1 << 16);
        CK(cudaMemcpy(a.data(), arg, a.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), arg_ref, b.size() * 4, cudaMemcpyDeviceToHost));
        size_t diff = 0; for (size_t i = 0; i < a.size(); ++i) diff += a[i] != b[i];
        printf("#   %s vs f32x16: %zu of %zu winners differ (fp16 table = filter, not the answer)\n", name, diff, a.size());
        CK(cudaFree(d));
    };
    run_h16("h16x64", 64, probe_h16<64, 512, 1>, 512);
    run_h16("h16x64w", 64, probe_h16<64, 384, 1>, 384);
    run_h16("h16x32", 32, probe_h16<32, 512, 2>, 512);
    run_h16("h16x32w", 32, probe_h16<32, 1024, 1>, 1024);

    {   // shared-memory atomics: config-2 shaped bins, 2 atomics per stored entry
        const int pb = 784, kb = 10;
        unsigned long long *out; CK(cudaMalloc(&out, 8 * pb * kb)); CK(cudaMemset(out, 0, 8 * pb * kb));
        const size_t smem = (size_t)pb * kb * 12;
        auto kern = probe_atoms<512>;
        const int blocks = blocks_for(kern, 512, smem, sms);
        if (blocks) {
            const float ms = bench([&] { kern<<<blocks, 512, smem>>>(sell, nslices, pb, kb, out); }, 3);
            printf("%-10s %10.3f %12.1f   u64+u32 shared atomics per entry, %lld entries: %.2f ns/entry/SM-parallel, %.2f cycles per lane-atomic per SM at 1.9 GHz\n",
                   "atoms", ms, (double)n * 416.0 / ms * 1e-6, (long long)(n * 52), ms * 1e6 / (double)(n * 52),
                   ms * 1e-3 * 1.9e9 * sms / (double)(n * 52 * 2));
        }
        CK(cudaFree(out));
    }
    return 0;
}
