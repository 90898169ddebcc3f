// philox.cuh -- counter-based Philox4x32-10 in registers + the draw-index contract.
//
// Replaces the reference's `rand::rng()` + `Uniform`/`Normal` samplers
// (particle_swarm.rs:367,373,407-408; cross_entropy.rs:270), which are unseedable.  One Philox block per
// (particle, dimension) yields exactly the two uniforms (rp, rg) the reference draws per dimension.
//
//   key = (seed lo32, seed hi32)
//   ctr = (gidx lo32, gidx hi32, stream << 28 | j, iteration)       j = dimension (PSO) or dim pair (CE)
//   u   = ((w_hi mod 2^20) * 2^32 + w_lo) * 2^-52   in [0,1)        (w_lo, w_hi) = words (0,1) or (2,3)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "devmath.cuh"

namespace eqs {

enum : uint32_t { STREAM_PSO_INIT = 0u, STREAM_PSO_STEP = 1u, STREAM_CE = 2u };

constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;

// the ten round keys, expanded once on the host: as kernel parameters they are constant-bank operands of the
// LOP3s, so a round is exactly 2 IMAD.WIDE.U32 + 2 LOP3 with no key arithmetic on the device
struct PhiloxKey {
    uint32_t k0[10], k1[10];
};

__host__ __device__ inline PhiloxKey philox_key(uint32_t lo, uint32_t hi)
{
    PhiloxKey k;
    for (int r = 0; r < 10; ++r) {
        k.k0[r] = lo + (uint32_t)r * PHILOX_W0;
        k.k1[r] = hi + (uint32_t)r * PHILOX_W1;
    }
    return k;
}

__host__ __device__ inline PhiloxKey philox_key(uint64_t seed) { return philox_key((uint32_t)seed, (uint32_t)(seed >> 32)); }

// EQS_PHILOX_ROUNDS exists for ONE purpose: measuring what the ten rounds cost (profiles/r02_sweep.md builds a 7-round
// variant, Random123's smallest Crush-resistant one).  The draw contract, the oracle and every parity test are 10 rounds.
#ifndef EQS_PHILOX_ROUNDS
#define EQS_PHILOX_ROUNDS 10
#endif
__host__ __device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey &key)
{
#pragma unroll
    for (int r = 0; r < EQS_PHILOX_ROUNDS; ++r) {
        const uint32_t k0 = key.k0[r], k1 = key.k1[r];
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return make_uint4(c0, c1, c2, c3);
}

__host__ __device__ __forceinline__ uint4 philox_draw(const PhiloxKey &key, uint32_t stream, uint32_t t, int64_t gidx, uint32_t j)
{
    return philox4x32_10((uint32_t)(uint64_t)gidx, (uint32_t)((uint64_t)gidx >> 32), (stream << 28) | j, t, key);
}

// 52 random mantissa bits pasted under exponent 0 give a double in [1,2); minus one is exact, so this equals
// (double)k * 2^-52 bit for bit (k = (w_hi mod 2^20) * 2^32 + w_lo).  One LOP3 + one DADD; it is also how
// rand's Uniform<f64> turns bits into a float (into_float_with_exponent(0) - 1.0).
__device__ __forceinline__ double u01(uint32_t w_lo, uint32_t w_hi)
{
    return __dadd_rn(__hiloint2double((int)((w_hi & 0xFFFFFu) | 0x3FF00000u), (int)w_lo), -1.0);
}

#ifdef EQS_DEVMATH_TABLES
// Box-Muller on one block per dimension PAIR: z0 = r cos(2 pi u2), z1 = r sin(2 pi u2), r = sqrt(-2 ln(1-u1)).
// Stands in for rand_distr's StandardNormal (cross_entropy.rs:270).  The logarithm reads devmath.cuh's shared-memory
// tables: the calling kernel runs devmath_tables_load() first.
__device__ __forceinline__ void normal_pair(uint4 w, double &z0, double &z1)
{
    // 1 - u1 = 2 - (1.mantissa): the same value as 1.0 - u01(w.x, w.y) (both differences are exact) in one DADD
    const double u1c = __dsub_rn(2.0, __hiloint2double((int)((w.y & 0xFFFFFu) | 0x3FF00000u), (int)w.x));
    const double r = sqrt_boxmuller(-2.0 * log_unit_tab(u1c));
    double s, c;
    sincos2pi_mantissa(__hiloint2double((int)((w.w & 0xFFFFFu) | 0x3FF00000u), (int)w.z), s, c); // u2 = this - 1
    z0 = r * c;
    z1 = r * s;
}
#endif // EQS_DEVMATH_TABLES

} // namespace eqs
