// Philox4x32-10 counter-based generator (Salmon et al., SC'11), device + host.
//
// The reference draws from an unseeded NumPy PCG64 (basicrta/gibbs.py:17); a device
// sampler needs a stream addressable by (seed, chain, iteration, datum) so that the
// result does not depend on how a chain is split over CTAs or GPUs.  Host mirror:
// oracle/philox.py (pinned by the Random123 known-answer vectors).
//
//   key     = (seed & 0xffffffff, seed >> 32)
//   counter = (x, iteration, chain_id, purpose)
//   purpose 0        indicator uniforms, x = datum >> 2, datum i uses word (i & 3)
//   purpose 1 + 4k   Dirichlet gamma of component k, x = rejection trial
//   purpose 2 + 4k   rate gamma of component k,      x = rejection trial
#pragma once
#include <stdint.h>

namespace brta {

struct Words4 { uint32_t x, y, z, w; };

constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;

// 32 x 32 -> 64-bit product as (hi, lo): one IMAD.WIDE.U32.
__host__ __device__ __forceinline__ void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo)
{
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .u64 t;\n\tmul.wide.u32 t, %2, %3;\n\tmov.b64 {%0, %1}, t;\n\t}"
        : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
#else
    const uint64_t p = (uint64_t)a * b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
#endif
}

__host__ __device__ __forceinline__ Words4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                         uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(PHILOX_M0, c0, hi0, lo0);
        mulhilo(PHILOX_M1, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ k0;
        c2 = hi0 ^ c3 ^ k1;
        c1 = lo1;
        c3 = lo0;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    return Words4{c0, c1, c2, c3};
}

// Same generator with the 10 round keys precomputed (rk[2r], rk[2r+1] = key + r * Weyl).
// In the sweep kernel rk lives in the kernel-parameter constant bank, so each round is
// 2 IMAD.WIDE + 2 LOP3 with no key-schedule arithmetic in the loop.
struct RoundKeys { uint32_t k[20]; };

__host__ __device__ inline RoundKeys philox_round_keys(uint64_t seed)
{
    RoundKeys rk;
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) {
        rk.k[2 * r] = k0;
        rk.k[2 * r + 1] = k1;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    return rk;
}

__device__ __forceinline__ Words4 philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                   const RoundKeys& rk)
{
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        uint32_t hi0, lo0, hi1, lo1;
        mulhilo(PHILOX_M0, c0, hi0, lo0);
        mulhilo(PHILOX_M1, c2, hi1, lo1);
        c0 = hi1 ^ c1 ^ rk.k[2 * round];
        c2 = hi0 ^ c3 ^ rk.k[2 * round + 1];
        c1 = lo1;
        c3 = lo0;
    }
    return Words4{c0, c1, c2, c3};
}

// uint32 -> float32 in [0,1): the top 23 bits become the mantissa of a float in [1,2).
// Exact, no I2F (keeps the XU pipe for MUFU.EX2).  Mirror: oracle/philox.py word_to_uniform.
__device__ __forceinline__ float word_to_12(uint32_t w)      // 1 + u, in [1, 2)
{
    return __uint_as_float(0x3f800000u | (w >> 9));
}
__device__ __forceinline__ float word_to_unit(uint32_t w)
{
    return word_to_12(w) - 1.0f;
}

// (0,1] and (0,1) variants for the posterior draws (24 bits).
__device__ __forceinline__ float word_to_unit_open_low(uint32_t w)   // (0, 1]
{
    return __fmul_rn((float)((w >> 8) + 1u), 5.9604644775390625e-08f);
}
__device__ __forceinline__ float word_to_unit_open(uint32_t w)       // (0, 1)
{
    return __fmul_rn(__fadd_rn((float)(w >> 8), 0.5f), 5.9604644775390625e-08f);
}

}  // namespace brta
