// Scalar math of the sampler: exp2 in two flavours, warp reductions, and the
// Marsaglia-Tsang log-gamma draw used for the Dirichlet / Gamma posterior update
// (basicrta/gibbs.py:210-211 calls numpy's Generator.dirichlet / .gamma there).
#pragma once
#include <stdint.h>
#include "brta_rng.cuh"

namespace brta {

constexpr float LOG2E = 1.4426950408889634f;
constexpr unsigned FULL = 0xffffffffu;

// ---- exp2 ---------------------------------------------------------------------------
// FAST: one MUFU.EX2 (the roofline unit of this kernel).
__device__ __forceinline__ float fast_exp2(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// EXACT: IEEE-only, reproduced operation by operation in oracle/gibbs_oracle.py
// (soft_exp2).  x <= 0 expected; x < -125 (also -inf, NaN) gives exactly 0.
__device__ __forceinline__ float soft_exp2(float x)
{
    const bool alive = x >= -125.0f;
    const float xc = alive ? x : 0.0f;
    const float z = __fadd_rn(xc, 12582912.0f);          // 1.5*2^23: rounds xc to an integer
    const float nf = __fsub_rn(z, 12582912.0f);
    const float f = __fsub_rn(xc, nf);                   // exact, |f| <= 0.5
    float p = __uint_as_float(0x377fe5feu);              // (ln2)^7/7!
    p = __fadd_rn(__fmul_rn(p, f), __uint_as_float(0x39218489u));
    p = __fadd_rn(__fmul_rn(p, f), __uint_as_float(0x3aaec3ffu));
    p = __fadd_rn(__fmul_rn(p, f), __uint_as_float(0x3c1d955bu));
    p = __fadd_rn(__fmul_rn(p, f), __uint_as_float(0x3d635847u));
    p = __fadd_rn(__fmul_rn(p, f), __uint_as_float(0x3e75fdf0u));
    p = __fadd_rn(__fmul_rn(p, f), __uint_as_float(0x3f317218u));
    p = __fadd_rn(__fmul_rn(p, f), 1.0f);
    const int n = __float_as_int(z) - 0x4b400000;        // integer value of nf
    const float r = __int_as_float(__float_as_int(p) + (n << 23));
    return alive ? r : 0.0f;
}

// ---- warp reductions (xor butterfly: every lane ends with the same bits) -------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// ---- Gamma(shape, 1) variates ---------------------------------------------------------
// Marsaglia & Tsang (2000).  One Philox call per trial: words 0,1 -> Box-Muller normal,
// 2 -> accept uniform, 3 -> boost uniform.  The squeeze test accepts ~92 % of the trials
// without a logarithm; the exact test needs log(v) accurately (v is within 1e-3 of 1 for
// the large shapes 1 + n_k), hence log1pf there and MUFU-grade intrinsics elsewhere.
// Returns d*v with d = a - 1/3 for the (possibly boosted) shape a >= 1.
__device__ __forceinline__ float gamma_core(float a, uint32_t iter, uint32_t chain, uint32_t purpose,
                                            uint32_t k0, uint32_t k1, uint32_t& boost_word)
{
    const float d = a - (1.0f / 3.0f);
    const float c = rsqrtf(9.0f * d);
    float result = d;                                     // fallback: never reached in practice
    boost_word = 0x80000000u;
    for (uint32_t trial = 0; trial < 64u; ++trial) {
        const Words4 w = philox4x32_10(trial, iter, chain, purpose, k0, k1);
        const float u1 = word_to_unit_open_low(w.x);
        const float u2 = word_to_unit(w.y);
        const float z = sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
        const float x = c * z;
        if (x <= -1.0f) continue;
        const float t = 1.0f + x;
        const float v = t * t * t;
        const float u = word_to_unit_open(w.z);
        const float z2 = z * z;
        bool ok = u < 1.0f - 0.0331f * z2 * z2;
        if (!ok) ok = __logf(u) < 0.5f * z2 + d * (1.0f - v + 3.0f * log1pf(x));
        if (ok) {
            result = d * v;
            boost_word = w.w;
            break;
        }
    }
    return result;
}

// Gamma(shape, 1) in linear space (rates: shape = 1 + n_k >= 1 with the default prior).
__device__ __forceinline__ float gamma_draw(float shape, uint32_t iter, uint32_t chain, uint32_t purpose,
                                            uint32_t k0, uint32_t k1)
{
    const bool boost = shape < 1.0f;
    uint32_t bw;
    float g = gamma_core(boost ? shape + 1.0f : shape, iter, chain, purpose, k0, k1, bw);
    if (boost) g *= fast_exp2(__log2f(word_to_unit_open_low(bw)) / shape);
    return g;
}

// log2 of a Gamma(shape, 1) variate.  The shape < 1 boost G(a) = G(a+1) U^(1/a) stays in
// log space: an empty component has Dirichlet shape 1/K (gibbs.py:173), where U^K
// underflows float32; the sampler only ever needs log w_k.
__device__ __forceinline__ float gamma_log2_draw(float shape, uint32_t iter, uint32_t chain,
                                                 uint32_t purpose, uint32_t k0, uint32_t k1)
{
    const bool boost = shape < 1.0f;
    uint32_t bw;
    const float g = gamma_core(boost ? shape + 1.0f : shape, iter, chain, purpose, k0, k1, bw);
    float l2 = __log2f(g);
    if (boost) l2 += __log2f(word_to_unit_open_low(bw)) / shape;
    return l2;
}

}  // namespace brta
