// philox.cuh -- the counter-based generator behind the on-device row samplers (fwht.cu, dct.cu).
// Philox4x32-10 keyed by the 64-bit seed and counted by (draw index, attempt, global column): a draw is a
// pure function of those, so a column's sample does not depend on thread timing or on how columns are
// sharded over GPUs (the contract of private/randsample_block.m:44-84 is distributional; MATLAB's own
// generator is closed source).
#pragma once
#include <stdint.h>

__device__ __forceinline__ uint32_t skm_philox_draw(uint64_t seed, uint64_t col, uint32_t draw, uint32_t attempt)
{
    uint32_t c0 = draw, c1 = attempt, c2 = (uint32_t)col, c3 = (uint32_t)(col >> 32);
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}
