// sampler.cuh -- counter-based restatement of gen_uniform_sampling (point2d.jl:71-75,
// point3d.jl:82-86, arm22.jl:214-218).  Julia's task-local RNG stream cannot be reproduced
// outside Julia, so try t of (seed, agent) is defined as Philox4x32-10 of the counter
// (t_lo, t_hi, agent, block) under the key (seed_lo, seed_hi): any try can be generated
// anywhere, in any order, on any GPU.
#pragma once
#include "models.cuh"

namespace mrmp {

// ---- Philox4x32-10 (Salmon et al., SC'11), counter = (t_lo, t_hi, agent, block) ----------
__device__ __forceinline__ void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                           uint32_t k0, uint32_t k1, uint32_t *out) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0;
        c1 = lo1;
        c2 = n2;
        c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0;
    out[1] = c1;
    out[2] = c2;
    out[3] = c3;
}

__device__ __forceinline__ double u01_53(uint32_t hi, uint32_t lo) {
    // exact: 27 + 26 bits, then a power-of-two scale
    return dmul(dadd(dmul((double)(hi >> 5), 67108864.0), (double)(lo >> 6)),
                1.0 / 9007199254740992.0);
}

// gen_uniform_sampling: point2d.jl:71-75, point3d.jl:82-86, arm22.jl:214-218
template <class M>
__device__ __forceinline__ void sample_state(uint64_t seed, int agent, uint64_t t, double *q) {
    // uniform k of a try = pair (k % 2) of Philox block k / 2
#pragma unroll
    for (int b = 0; 2 * b < M::SD; ++b) {
        uint32_t r[4];
        philox4x32((uint32_t)t, (uint32_t)(t >> 32), (uint32_t)agent, (uint32_t)b, (uint32_t)seed,
                   (uint32_t)(seed >> 32), r);
        q[2 * b] = u01_53(r[0], r[1]);
        if (2 * b + 1 < M::SD) q[2 * b + 1] = u01_53(r[2], r[3]);
    }
    M::scale_sample(q);
}

}  // namespace mrmp
