// update.cu -- K2 (per-cluster row sums S, support counts N, member counts, sum of squared
// distances) and K3 (centre finalisation, centre change) of the Lloyd iteration.
//   K2 replaces kmeans_sparsified.m:430-453  (sum(X(:,ind),2), sum(NormalizationMatrix(:,ind),2))
//   K3 replaces kmeans_sparsified.m:448,450,470-471
// The partials buffer is [S (p*K) | N (p*K) | counts (K) | sumsq (1)] in doubles, column-major
// per cluster (S[k*p + r]); it is what the multi-GPU all-reduce sums.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace {

__device__ __forceinline__ void ordered_block_sum(double s_block, double *__restrict__ scratch, unsigned int *__restrict__ ticket,
                                                  double *__restrict__ out);

// v1: one warp per column, lanes over the column's stored entries, fp64 atomics to L2.
template <typename VT>
__global__ void k_accumulate_csc(int64_t p, int64_t n, int64_t K, const int64_t *__restrict__ colptr,
                                 const int32_t *__restrict__ rowidx, const VT *__restrict__ val,
                                 const int32_t *__restrict__ assign, const float *__restrict__ dist32,
                                 const double *__restrict__ dist64, double *__restrict__ partials,
                                 double *__restrict__ red_scratch, unsigned int *__restrict__ red_ticket)
{
    double *S = partials, *N = partials + p * K, *counts = partials + 2 * p * K;
    double *sumsq = counts + K;
    const int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double local_sq = 0.0;
    for (int64_t j = warp; j < n; j += nwarps) {
        const int64_t k = assign[j];
        if (k < 0 || k >= K) continue;
        const int64_t t0 = colptr[j], t1 = colptr[j + 1];
        for (int64_t t = t0 + lane; t < t1; t += 32) {
            const int64_t r = rowidx[t];
            atomicAdd(&S[k * p + r], (double)val[t]);
            atomicAdd(&N[k * p + r], 1.0);
        }
        if (lane == 0) {
            atomicAdd(&counts[k], 1.0);
            const double d = dist64 ? dist64[j] : (double)dist32[j];
            local_sq += d * d;
        }
    }
    // block reduction of the squared distances (lane 0 of each warp holds a partial)
    __shared__ double red[32];
    if (lane == 0) red[threadIdx.x >> 5] = local_sq;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    ordered_block_sum(s, red_scratch, red_ticket, sumsq);
}


// Sum of one double per block, independent of the order the blocks finish in: every block deposits its partial and takes
// a ticket; the block that draws the last ticket adds the partials up in a fixed pattern (thread t takes t, t + T, ...;
// then a fixed tree).  The objective (sum of squared distances) decides between replicates that reach the same clustering,
// so it must not depend on scheduling (an fp64 atomicAdd per block made two identical runs differ in the last bits).
__device__ __forceinline__ void ordered_block_sum(double s_block /* valid in thread 0 */, double *__restrict__ scratch,
                                                  unsigned int *__restrict__ ticket, double *__restrict__ out)
{
    __shared__ bool last;
    __shared__ double tree[32];
    if (threadIdx.x == 0) {
        scratch[blockIdx.x] = s_block;
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double t = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) t += __ldcg(scratch + b);
#pragma unroll
    for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if ((threadIdx.x & 31) == 0) tree[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += tree[w];
        *out += tot;
        *ticket = 0;                                          // ready for the next launch on this stream
    }
}

// members per cluster, sum of squared distances, and the compact copy of the assignments that
// K2's row-major pass gathers from (1 byte per column when K <= 256).
template <typename AT>
__global__ void k_count_sumsq(int64_t n, int64_t K, const int32_t *__restrict__ assign,
                              const float *__restrict__ dist32, const double *__restrict__ dist64,
                              AT *__restrict__ assign_c, double *__restrict__ counts, double *__restrict__ sumsq,
                              double *__restrict__ red_scratch, unsigned int *__restrict__ red_ticket)
{
    extern __shared__ int hist[];
    for (int k = threadIdx.x; k < K; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    double local = 0.0;
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < n; j += stride) {
        const int a = assign[j];
        assign_c[j] = (AT)a;
        if (a >= 0 && a < K) {
            atomicAdd(&hist[a], 1);
            const double d = dist64 ? dist64[j] : (double)dist32[j];
            local += d * d;
        }
    }
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (hist[k]) atomicAdd(&counts[k], (double)hist[k]);      // integers: exact in any order
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    ordered_block_sum(s, red_scratch, red_ticket, sumsq);
}

// K2 over the row-major image: one warp per work unit (a run of one row's entries).  Lanes own
// private bin columns [cluster][BW] in shared memory (fp64 sums, int counts): with BW = 32
// every lane has its own column and the read-modify-writes are conflict-free without atomics;
// for larger K the columns are shared by 32/BW lanes, which take turns (BW shrinks so that
// enough warps stay resident to hide the stream/gather latency).  The bins are folded across
// columns with a rotated read and leave the SM as one fp64 atomic per (row, cluster, unit).
template <typename AT, int BW>
__global__ void k_accumulate_csr(int64_t p, int k0, int kb, int64_t nunits,
                                 const int32_t *__restrict__ unit_row, const int64_t *__restrict__ unit_start,
                                 const int2 *__restrict__ csr, const AT *__restrict__ assign_c,
                                 double *__restrict__ S, double *__restrict__ N, unsigned long long *__restrict__ next_unit)
{
    extern __shared__ __align__(16) unsigned char acc_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double *binS = reinterpret_cast<double *>(acc_raw) + (size_t)warp * kb * BW;
    int *binN = reinterpret_cast<int *>(reinterpret_cast<double *>(acc_raw) + (size_t)nw * kb * BW) + (size_t)warp * kb * BW;
    const int64_t *unit_end = unit_start + nunits;
    const int col = lane & (BW - 1);
    constexpr int PHASES = 32 / BW;
    const int my_phase = lane / BW;

    auto add = [&](int a, int xbits) {
        if (PHASES == 1) {
            if ((unsigned)a < (unsigned)kb) { binS[a * BW + col] += (double)__int_as_float(xbits); binN[a * BW + col] += 1; }
        } else {
#pragma unroll
            for (int ph = 0; ph < PHASES; ++ph) {
                if (my_phase == ph && (unsigned)a < (unsigned)kb) {
                    binS[a * BW + col] += (double)__int_as_float(xbits);
                    binN[a * BW + col] += 1;
                }
                __syncwarp();
            }
        }
    };

    for (;;) {
        unsigned long long uu = 0;
        if (lane == 0) uu = atomicAdd(next_unit, 1ULL);          // dynamic schedule: one unit per fetch
        const int64_t u = (int64_t)__shfl_sync(0xffffffffu, uu, 0);
        if (u >= nunits) break;
        for (int i = lane; i < kb * BW; i += 32) { binS[i] = 0.0; binN[i] = 0; }
        __syncwarp();
        const int64_t r = unit_row[u], s = unit_start[u], e = unit_end[u];
        int64_t base = s;                                        // warp-uniform trip counts throughout
        if (base + 128 <= e) {
            // software pipeline: the next 4 stream loads are in flight while the current 4 entries
            // are gathered and binned; entries of one lane that hit the same bin are merged in
            // registers first, so the 4 read-modify-writes are independent (loads, adds, stores)
            int2 n0 = __ldcs(csr + base + lane), n1 = __ldcs(csr + base + lane + 32);
            int2 n2 = __ldcs(csr + base + lane + 64), n3 = __ldcs(csr + base + lane + 96);
            for (; base + 128 <= e; base += 128) {
                const int2 q0 = n0, q1 = n1, q2 = n2, q3 = n3;
                int a0 = (int)__ldg(assign_c + q0.x) - k0, a1 = (int)__ldg(assign_c + q1.x) - k0;
                int a2 = (int)__ldg(assign_c + q2.x) - k0, a3 = (int)__ldg(assign_c + q3.x) - k0;
                if (base + 256 <= e) {
                    const int64_t i2 = base + 128 + lane;
                    n0 = __ldcs(csr + i2); n1 = __ldcs(csr + i2 + 32); n2 = __ldcs(csr + i2 + 64); n3 = __ldcs(csr + i2 + 96);
                }
                double x0 = (double)__int_as_float(q0.y), x1 = (double)__int_as_float(q1.y);
                double x2 = (double)__int_as_float(q2.y), x3 = (double)__int_as_float(q3.y);
                int c0 = 1, c1 = 1, c2 = 1, c3 = 1;
                if ((unsigned)a0 >= (unsigned)kb) a0 = -1;
                if ((unsigned)a1 >= (unsigned)kb) a1 = -1;
                if ((unsigned)a2 >= (unsigned)kb) a2 = -1;
                if ((unsigned)a3 >= (unsigned)kb) a3 = -1;
                if (a3 >= 0 && a3 == a2) { x2 += x3; c2 += c3; a3 = -1; }
                if (a3 >= 0 && a3 == a1) { x1 += x3; c1 += c3; a3 = -1; }
                if (a3 >= 0 && a3 == a0) { x0 += x3; c0 += c3; a3 = -1; }
                if (a2 >= 0 && a2 == a1) { x1 += x2; c1 += c2; a2 = -1; }
                if (a2 >= 0 && a2 == a0) { x0 += x2; c0 += c2; a2 = -1; }
                if (a1 >= 0 && a1 == a0) { x0 += x1; c0 += c1; a1 = -1; }
                const int i0 = a0 * BW + col, i1 = a1 * BW + col, i2b = a2 * BW + col, i3 = a3 * BW + col;
#pragma unroll
                for (int ph = 0; ph < PHASES; ++ph) {               // lanes sharing a bin column take turns
                    if (PHASES == 1 || my_phase == ph) {
                        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                        int m0 = 0, m1 = 0, m2 = 0, m3 = 0;
                        if (a0 >= 0) { s0 = binS[i0]; m0 = binN[i0]; }
                        if (a1 >= 0) { s1 = binS[i1]; m1 = binN[i1]; }
                        if (a2 >= 0) { s2 = binS[i2b]; m2 = binN[i2b]; }
                        if (a3 >= 0) { s3 = binS[i3]; m3 = binN[i3]; }
                        if (a0 >= 0) { binS[i0] = s0 + x0; binN[i0] = m0 + c0; }
                        if (a1 >= 0) { binS[i1] = s1 + x1; binN[i1] = m1 + c1; }
                        if (a2 >= 0) { binS[i2b] = s2 + x2; binN[i2b] = m2 + c2; }
                        if (a3 >= 0) { binS[i3] = s3 + x3; binN[i3] = m3 + c3; }
                    }
                    if (PHASES > 1) __syncwarp();
                }
            }
        }
        for (; base + 128 <= e; base += 128) {
            const int64_t i = base + lane;
            const int2 q0 = __ldcs(csr + i), q1 = __ldcs(csr + i + 32), q2 = __ldcs(csr + i + 64), q3 = __ldcs(csr + i + 96);
            const int a0 = (int)__ldg(assign_c + q0.x) - k0, a1 = (int)__ldg(assign_c + q1.x) - k0;
            const int a2 = (int)__ldg(assign_c + q2.x) - k0, a3 = (int)__ldg(assign_c + q3.x) - k0;
            add(a0, q0.y); add(a1, q1.y); add(a2, q2.y); add(a3, q3.y);
        }
        for (; base < e; base += 32) {
            const int64_t ii = base + lane;
            int a = -1, xb = 0;
            if (ii < e) { const int2 q = __ldcs(csr + ii); a = (int)__ldg(assign_c + q.x) - k0; xb = q.y; }
            add(a, xb);
        }
        __syncwarp();
        for (int kblk = 0; kblk < kb; kblk += 32) {
            const int k = kblk + lane;
            double sum = 0.0;
            int cnt = 0;
            if (k < kb) {
#pragma unroll 8
                for (int t = 0; t < BW; ++t) {
                    const int idx = k * BW + ((t + lane) & (BW - 1));
                    sum += binS[idx];
                    cnt += binN[idx];
                }
                if (cnt) {
                    atomicAdd(&S[(int64_t)(k0 + k) * p + r], sum);
                    atomicAdd(&N[(int64_t)(k0 + k) * p + r], (double)cnt);
                }
            }
        }
        __syncwarp();
    }
}

// ---- incremental K2 (opt-in): only the columns whose assignment changed move their entries ----
// k_diff_assign: list of changed columns, member-count deltas, and the sum of squared distances
// (that one changes for every column, so it is recomputed in full here).
__global__ void k_diff_assign(int64_t n, int64_t K, const int32_t *__restrict__ assign, const int32_t *__restrict__ prev,
                              const float *__restrict__ dist32, const double *__restrict__ dist64,
                              int32_t *__restrict__ changed, int *__restrict__ nchanged, double *__restrict__ sumsq,
                              double *__restrict__ red_scratch, unsigned int *__restrict__ red_ticket)
{
    double local = 0.0;
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < n; j += stride) {
        const int a = assign[j];
        if (a != prev[j]) changed[atomicAdd(nchanged, 1)] = (int32_t)j;
        if (a >= 0 && a < K) {
            const double d = dist64 ? dist64[j] : (double)dist32[j];
            local += d * d;
        }
    }
    __shared__ double red[32];
#pragma unroll
    for (int o = 16; o; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    ordered_block_sum(s, red_scratch, red_ticket, sumsq);
}

// one warp per changed column: its entries leave the old cluster's sums and join the new one's
template <typename VT>
__global__ void k_move_changed(int64_t p, int64_t K, const int64_t *__restrict__ colptr, const int32_t *__restrict__ rowidx,
                               const VT *__restrict__ val, const int32_t *__restrict__ assign, int32_t *__restrict__ prev,
                               const int32_t *__restrict__ changed, const int *__restrict__ nchanged,
                               double *__restrict__ acc /* [S | N | counts] */)
{
    double *S = acc, *N = acc + p * K, *counts = acc + 2 * p * K;
    const int lane = threadIdx.x & 31;
    int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t total = *nchanged;
    for (; w < total; w += nwarps) {
        const int64_t j = changed[w];
        const int64_t ko = prev[j], kn = assign[j];
        const bool has_o = ko >= 0 && ko < K, has_n = kn >= 0 && kn < K;
        const int64_t t0 = colptr[j], t1 = colptr[j + 1];
        for (int64_t t = t0 + lane; t < t1; t += 32) {
            const int64_t r = rowidx[t];
            const double x = (double)val[t];
            if (has_o) { atomicAdd(&S[ko * p + r], -x); atomicAdd(&N[ko * p + r], -1.0); }
            if (has_n) { atomicAdd(&S[kn * p + r], x); atomicAdd(&N[kn * p + r], 1.0); }
        }
        __syncwarp();
        if (lane == 0) {
            if (has_o) atomicAdd(&counts[ko], -1.0);
            if (has_n) atomicAdd(&counts[kn], 1.0);
            prev[j] = (int32_t)kn;
        }
    }
}

// stats[0] += sum (old-new)^2 ; stats[1] = 1 if any NaN in new centres
__global__ void k_finalize(int64_t p, int64_t K, const double *__restrict__ partials, double gamma,
                           int ml, double *__restrict__ centers, double *__restrict__ centers_old,
                           double *__restrict__ stats)
{
    const double *S = partials, *N = partials + p * K, *counts = partials + 2 * p * K;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d2 = 0.0;
    int nan = 0;
    if (idx < p * K) {
        const int64_t k = idx / p;
        const double old = centers[idx];
        double nw = old;
        const double cnt = counts[k];
        if (cnt > 0.0) {
            // N (a sum of ones) is exact in fp64, also after the incremental +/- updates and the all-reduce;
            // a cell no member touches is a structural zero: the reference's S is exactly 0 there and
            // gamma*0/(0+1e-16) = 0.  The incremental path can leave a rounding residue in S for such a
            // cell ((a+b)-a-b != 0), which must not be divided by 1e-16.
            const double s = (N[idx] == 0.0) ? 0.0 : S[idx];
            if (ml) nw = __ddiv_rn(__dmul_rn(gamma, s), __dadd_rn(N[idx], 1e-16));
            else    nw = __ddiv_rn(s, cnt);
        }
        centers_old[idx] = old;
        centers[idx] = nw;
        const double d = old - nw;
        d2 = d * d;
        nan = (nw != nw);
    }
    __shared__ double red[32];
    __shared__ int rnan;
    if (threadIdx.x == 0) rnan = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
    if (nan) rnan = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(&stats[0], s);
        if (rnan) stats[1] = 1.0;
    }
}

// Multi-device K3 (multi.cu): the all-reduce of the partials is fused into the finalisation.  parts[g] points at
// device g's [S | N | counts | sumsq] -- peer memory over NVLink for g != this device -- and the sums run in
// device order, so every device computes bit-identical centres whatever the timing.  tail receives the reduced
// [counts | sumsq] for the statistics read-back.
__global__ void k_finalize_peers(int64_t p, int64_t K, int ndev, const double *const *__restrict__ parts, double gamma,
                                 int ml, double *__restrict__ centers, double *__restrict__ centers_old,
                                 double *__restrict__ stats, double *__restrict__ tail)
{
    const int64_t pk = p * K;
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d2 = 0.0;
    int nan = 0;
    if (idx < pk) {
        const int64_t k = idx / p;
        double S = 0.0, N = 0.0, cnt = 0.0;
        for (int g = 0; g < ndev; ++g) {
            const double *q = parts[g];
            S += q[idx]; N += q[pk + idx]; cnt += q[2 * pk + k];
        }
        const double old = centers[idx];
        double nw = old;
        if (cnt > 0.0) {
            const double s = (N == 0.0) ? 0.0 : S;                       // structural zero, see k_finalize
            if (ml) nw = __ddiv_rn(__dmul_rn(gamma, s), __dadd_rn(N, 1e-16));
            else    nw = __ddiv_rn(s, cnt);
        }
        centers_old[idx] = old;
        centers[idx] = nw;
        const double d = old - nw;
        d2 = d * d;
        nan = (nw != nw);
    }
    if (blockIdx.x == 0)
        for (int64_t i = threadIdx.x; i <= K; i += blockDim.x) {
            double t = 0.0;
            for (int g = 0; g < ndev; ++g) t += parts[g][2 * pk + i];
            tail[i] = t;
        }
    __shared__ double red[32];
    __shared__ int rnan;
    if (threadIdx.x == 0) rnan = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
    if (nan) rnan = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(&stats[0], s);
        if (rnan) stats[1] = 1.0;
    }
}

// centre change without an update (after the host patched columns for EmptyAction)
__global__ void k_diff(int64_t total, const double *__restrict__ centers,
                       const double *__restrict__ centers_old, double *__restrict__ stats)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double d2 = 0.0;
    int nan = 0;
    if (idx < total) {
        const double d = centers_old[idx] - centers[idx];
        d2 = d * d;
        nan = (centers[idx] != centers[idx]);
    }
    __shared__ double red[32];
    __shared__ int rnan;
    if (threadIdx.x == 0) rnan = 0;
    __syncthreads();
#pragma unroll
    for (int o = 16; o; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = d2;
    if (nan) rnan = 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(&stats[0], s);
        if (rnan) stats[1] = 1.0;
    }
}

// first index attaining the maximum (MATLAB max): per-block candidates, then one block
template <typename T>
__global__ void k_argmax_blocks(int64_t n, const T *__restrict__ d, double *__restrict__ bval,
                                int64_t *__restrict__ bidx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    double bv = 0.0;
    int64_t bi = -1;
    for (; i < n; i += stride) {
        const double v = (double)d[i];
        if (v != v) continue;                                  // max skips NaN
        if (bi < 0 || v > bv) { bv = v; bi = i; }
    }
    __shared__ double sv[256];
    __shared__ int64_t si[256];
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x >> 1; o; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const double ov = sv[threadIdx.x + o];
            const int64_t oi = si[threadIdx.x + o];
            const double mv = sv[threadIdx.x];
            const int64_t mi = si[threadIdx.x];
            if (oi >= 0 && (mi < 0 || ov > mv || (ov == mv && oi < mi))) { sv[threadIdx.x] = ov; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { bval[blockIdx.x] = sv[0]; bidx[blockIdx.x] = si[0]; }
}

__global__ void k_argmax_final(int nb, const double *__restrict__ bval, const int64_t *__restrict__ bidx,
                               double *__restrict__ out_val, int64_t *__restrict__ out_idx)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double bv = 0.0;
    int64_t bi = -1;
    for (int b = 0; b < nb; ++b) {
        const double ov = bval[b];
        const int64_t oi = bidx[b];
        if (oi >= 0 && (bi < 0 || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
    }
    *out_val = bi >= 0 ? bv : __longlong_as_double(0x7ff8000000000000LL);
    *out_idx = bi >= 0 ? bi : 0;
}

}  // namespace

template <typename AT>
static int accumulate_csr_bins(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const AT *assign_c, double *S, double *N);

template <typename AT>
static int accumulate_csr(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const int32_t *assign, void *assign_c,
                          const float *dist32, const double *dist64, double *partials)
{
    const int64_t p = ds->p, n = ds->n;
    double *S = partials, *N = partials + p * K, *counts = partials + 2 * p * K, *sumsq = counts + K;
    {
        int64_t blocks = (n + 255) / 256;
        const int64_t cap = std::min<int64_t>((int64_t)ctx->sm_count * 8, SKM_RED_BLOCKS);   // ordered_block_sum: one scratch slot per block
        if (blocks > cap) blocks = cap;
        const size_t sm = sizeof(int) * (size_t)K;
        if (sm > 48 * 1024) SKM_CUDA(cudaFuncSetAttribute(k_count_sumsq<AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        k_count_sumsq<AT><<<(unsigned)blocks, 256, sm, ctx->stream>>>(n, K, assign, dist32, dist64, (AT *)assign_c, counts, sumsq,
                                                                     ctx->red_scratch, ctx->red_ticket);
        SKM_CHECK_LAUNCH(ctx);
    }
    return accumulate_csr_bins<AT>(ctx, ds, K, (const AT *)assign_c, S, N);
}

template <typename AT, int BW>
static int launch_csr_bw(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const AT *assign_c, double *S, double *N)
{
    // bins: 12 bytes x BW columns per cluster per warp
    const int64_t p = ds->p;
    const size_t budget = (size_t)ctx->smem_optin - 2048;
    const size_t per_k = (size_t)12 * BW;
    int warps = 8;
    {
        static const char *e = getenv("SKM_K2_WARPS");              // tuning knob
        if (e && atoi(e) > 0 && atoi(e) <= 32) warps = atoi(e);
    }
    while (warps > 1 && (size_t)warps * per_k * (size_t)(K < 32 ? K : 32) > budget) warps >>= 1;
    int64_t kb = (int64_t)(budget / ((size_t)warps * per_k));
    if (kb > K) kb = K;
    const size_t smem = (size_t)warps * kb * per_k;
    auto kern = k_accumulate_csr<AT, BW>;
    SKM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    SKM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)ctx->sm_count * per_sm;
    const int64_t need = (ds->nunits + warps - 1) / warps;
    if (blocks > need) blocks = need;
    for (int64_t k0 = 0; k0 < K; k0 += kb) {
        const int kbb = (int)((K - k0) < kb ? (K - k0) : kb);
        SKM_CUDA(cudaMemsetAsync(ds->unit_counter, 0, sizeof(unsigned long long), ctx->stream));
        kern<<<(unsigned)blocks, warps * 32, smem, ctx->stream>>>(p, (int)k0, kbb, ds->nunits, ds->unit_row, ds->unit_start,
                                                               ds->csr, assign_c, S, N, ds->unit_counter);
        SKM_CHECK_LAUNCH(ctx);
    }
    return SKM_OK;
}

template <typename AT>
static int accumulate_csr_bins(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const AT *assign_c, double *S, double *N)
{
    // shrink the bin width until roughly 32 warps fit on an SM
    {
        static const char *e = getenv("SKM_K2_BW");                 // tuning knob: bin columns per cluster per warp
        const int bw = e ? atoi(e) : 0;
        if (bw == 32) return launch_csr_bw<AT, 32>(ctx, ds, K, assign_c, S, N);
        if (bw == 16) return launch_csr_bw<AT, 16>(ctx, ds, K, assign_c, S, N);
        if (bw == 8) return launch_csr_bw<AT, 8>(ctx, ds, K, assign_c, S, N);
        if (bw == 4) return launch_csr_bw<AT, 4>(ctx, ds, K, assign_c, S, N);
    }
    if (K <= 16) return launch_csr_bw<AT, 32>(ctx, ds, K, assign_c, S, N);
    // The kernel sits on the shared-memory pipe (ncu at K = 64, BW = 8: l1tex 93 %, mio_throttle the top stall): every
    // turn-taking phase is four shared-memory instructions, so fewer phases beat more resident warps down to 16 warps
    // per SM (K = 64, n = 1.25e7: BW 8 / 16 / 32 = 1.91 / 1.57 / 3.00 ms, profiles/r2_k2_k64.md)
    if (K <= 64) return launch_csr_bw<AT, 16>(ctx, ds, K, assign_c, S, N);
    if (K <= 128) return launch_csr_bw<AT, 8>(ctx, ds, K, assign_c, S, N);
    return launch_csr_bw<AT, 4>(ctx, ds, K, assign_c, S, N);
}

int skm_launch_accumulate(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const int32_t *assign, void *assign_c,
                          const float *dist32, const double *dist64, double *partials, bool zero_first)
{
    const int64_t p = ds->p, n = ds->n;
    if (zero_first)
        SKM_CUDA(cudaMemsetAsync(partials, 0, sizeof(double) * (size_t)(2 * p * K + K + 1), ctx->stream));
    if (n == 0) return SKM_OK;
    if (ds->csr && ds->nunits > 0 && assign_c) {
        if (K <= 256) return accumulate_csr<uint8_t>(ctx, ds, K, assign, assign_c, dist32, dist64, partials);
        if (K <= 65536) return accumulate_csr<uint16_t>(ctx, ds, K, assign, assign_c, dist32, dist64, partials);
        return accumulate_csr<int32_t>(ctx, ds, K, assign, assign_c, dist32, dist64, partials);
    }
    int64_t blocks = (n * 32 + 255) / 256;
    int64_t cap = std::min<int64_t>((int64_t)ctx->sm_count * 8, SKM_RED_BLOCKS);
    if (blocks > cap) blocks = cap;
    if (ds->store_dtype == SKM_F32)
        k_accumulate_csc<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            p, n, K, ds->colptr, ds->rowidx, (const float *)ds->val, assign, dist32, dist64, partials, ctx->red_scratch, ctx->red_ticket);
    else
        k_accumulate_csc<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(
            p, n, K, ds->colptr, ds->rowidx, (const double *)ds->val, assign, dist32, dist64, partials, ctx->red_scratch, ctx->red_ticket);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_finalize(skm_ctx *ctx, int64_t p, int64_t K, const double *partials, double gamma,
                        int ml_correction, double *centers, double *centers_old, double *stats)
{
    SKM_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 8, ctx->stream));
    int64_t total = p * K;
    if (total == 0) return SKM_OK;
    int64_t blocks = (total + 255) / 256;
    if (partials)
        k_finalize<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, K, partials, gamma, ml_correction, centers,
                                                             centers_old, stats);
    else
        k_diff<<<(unsigned)blocks, 256, 0, ctx->stream>>>(total, centers, centers_old, stats);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_finalize_peers(skm_ctx *ctx, int64_t p, int64_t K, int ndev, const double *const *parts_dev,
                              double gamma, int ml_correction, double *centers, double *centers_old,
                              double *stats, double *tail)
{
    SKM_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 8, ctx->stream));
    const int64_t total = p * K;
    const int64_t blocks = total > 0 ? (total + 255) / 256 : 1;
    k_finalize_peers<<<(unsigned)blocks, 256, 0, ctx->stream>>>(p, K, ndev, parts_dev, gamma, ml_correction, centers,
                                                               centers_old, stats, tail);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

namespace {
__global__ void k_export(int64_t n, const int32_t *__restrict__ a, int32_t *__restrict__ a_out, const float *__restrict__ d,
                         double *__restrict__ d_out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        if (a_out) a_out[j] = a[j] + 1;
        if (d_out) d_out[j] = (double)d[j];
    }
}
}  // namespace

int skm_launch_export(skm_ctx *ctx, int64_t n, const int32_t *a, int32_t *a_out, const float *d, double *d_out)
{
    if (n <= 0 || (!a_out && !d_out)) return SKM_OK;
    const int64_t blocks = std::min<int64_t>((n + 255) / 256, (int64_t)ctx->sm_count * 16);
    k_export<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, a_out ? a : nullptr, a_out, d_out ? d : nullptr, d_out);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

int skm_launch_argmax(skm_ctx *ctx, int64_t n, const float *dist32, const double *dist64,
                      double *out_val, int64_t *out_idx)
{
    const int nb = 256;
    DevBuf bv, bi;
    SKM_TRY(bv.alloc(sizeof(double) * nb));
    SKM_TRY(bi.alloc(sizeof(int64_t) * nb));
    if (dist64)
        k_argmax_blocks<double><<<nb, 256, 0, ctx->stream>>>(n, dist64, bv.as<double>(), bi.as<int64_t>());
    else
        k_argmax_blocks<float><<<nb, 256, 0, ctx->stream>>>(n, dist32, bv.as<double>(), bi.as<int64_t>());
    SKM_CHECK_LAUNCH(ctx);
    k_argmax_final<<<1, 32, 0, ctx->stream>>>(nb, bv.as<double>(), bi.as<int64_t>(), out_val, out_idx);
    SKM_CHECK_LAUNCH(ctx);
    SKM_CUDA(cudaStreamSynchronize(ctx->stream));   // temporaries are freed on return
    return SKM_OK;
}

// Incremental K2, step 1: changed[] / *nchanged and the fresh sum of squared distances (into *sumsq, zeroed here).
int skm_launch_diff_assign(skm_ctx *ctx, int64_t n, int64_t K, const int32_t *assign, const int32_t *prev,
                           const float *dist32, const double *dist64, int32_t *changed, int *nchanged, double *sumsq)
{
    SKM_CUDA(cudaMemsetAsync(nchanged, 0, sizeof(int), ctx->stream));
    SKM_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double), ctx->stream));
    if (n == 0) return SKM_OK;
    int64_t blocks = (n + 255) / 256;
    const int64_t cap = std::min<int64_t>((int64_t)ctx->sm_count * 16, SKM_RED_BLOCKS);
    if (blocks > cap) blocks = cap;
    k_diff_assign<<<(unsigned)blocks, 256, 0, ctx->stream>>>(n, K, assign, prev, dist32, dist64, changed, nchanged, sumsq,
                                                             ctx->red_scratch, ctx->red_ticket);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}

// Incremental K2, step 2: move the entries of the changed columns between clusters in acc = [S | N | counts].
int skm_launch_move_changed(skm_ctx *ctx, const skm_dataset *ds, int64_t K, const int32_t *assign, int32_t *prev,
                            const int32_t *changed, const int *nchanged, int64_t nchanged_host, double *acc)
{
    if (nchanged_host <= 0) return SKM_OK;
    int64_t blocks = (nchanged_host * 32 + 255) / 256;
    const int64_t cap = (int64_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (ds->store_dtype == SKM_F32)
        k_move_changed<float><<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->p, K, ds->colptr, ds->rowidx, (const float *)ds->val,
                                                                        assign, prev, changed, nchanged, acc);
    else
        k_move_changed<double><<<(unsigned)blocks, 256, 0, ctx->stream>>>(ds->p, K, ds->colptr, ds->rowidx, (const double *)ds->val,
                                                                         assign, prev, changed, nchanged, acc);
    SKM_CHECK_LAUNCH(ctx);
    return SKM_OK;
}
