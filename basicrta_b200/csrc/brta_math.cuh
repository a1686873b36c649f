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
// Marsaglia & Tsang (2000), returned as log2 of the variate (the sampler only ever needs
// log2 w_k and log2 r_k).  One Philox call per trial: words 0,1 -> Box-Muller normal,
// 2 -> accept uniform, 3 -> boost uniform.  A trial for the (possibly boosted) shape a >= 1 yields
// d*v with d = a - 1/3, v = (1 + x)^3, x = z / sqrt(9 d); the shape < 1 boost G(a) = G(a+1) U^(1/a)
// is applied in log2 space by the caller: an empty component has Dirichlet shape 1/K
// (gibbs.py:173), where U^K underflows float32.
//
// The part of a trial that does not depend on the shape (TrialRandoms) is computed while the team
// exchange is still in flight, so only the ~25 dependent instructions of trial_finish remain on the
// critical path once n_k is known.
struct TrialRandoms {
    float z;          // standard normal (Box-Muller, words 0 and 1)
    float half_zz;    // z^2 / 2
    float u;          // accept uniform in (0,1) (word 2)
    float log_u;      // ln u
    float squeeze;    // 1 - 0.0331 z^4
    float l2_boost;   // log2 of the boost uniform in (0,1] (word 3)
};

// Every floating-point operation below is spelled out (__fmul_rn / __fadd_rn / __fmaf_rn): the compiler may
// neither contract a multiply and an add into an FMA nor keep one apart, so every build of the kernel
// (any K, 3 or 4 CTAs per SM) draws bit-identical variates from the same Philox words.
constexpr float LN2 = 0.6931471805599453f;

__device__ __forceinline__ TrialRandoms trial_randoms(const Words4& w)
{
    TrialRandoms r;
    const float u1 = word_to_unit_open_low(w.x);
    const float u2 = word_to_unit(w.y);
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(__fmul_rn(-2.0f * LN2, __log2f(u1))));
    r.z = __fmul_rn(rad, __cosf(__fmul_rn(6.283185307179586f, u2)));
    r.u = word_to_unit_open(w.z);
    r.log_u = __fmul_rn(LN2, __log2f(r.u));
    const float z2 = __fmul_rn(r.z, r.z);
    r.half_zz = __fmul_rn(0.5f, z2);
    r.squeeze = __fmaf_rn(-0.0331f, __fmul_rn(z2, z2), 1.0f);
    r.l2_boost = __log2f(word_to_unit_open_low(w.w));
    return r;
}

struct GammaTrial { float l2g; bool ok; };               // log2(d v), accepted?

// Finish a trial for shape a >= 1: straight-line, both acceptance tests evaluated.
// Exact test: ln u < z^2/2 + d (1 - v + 3 ln(1 + x)).  For the large shapes 1 + n_k, x is tiny and the
// right-hand side is the small remainder of two cancellations (3 ln(1+x) against 1 - v up to x^3, and
// d * 4.5 x^2 against z^2/2 exactly, because 9 d c^2 = 1): for |x| < 1/8 it is evaluated as the series
//     d x^4 (-3/4 + 3/5 x - 3/6 x^2 + ... + 3/11 x^7)        (truncation < 1e-7 relative)
// -- no logarithm, no cancellation; otherwise 1 + x is far from 1 and the plain form is accurate.
__device__ __forceinline__ GammaTrial trial_finish(float a, const TrialRandoms& r)
{
    const float d = __fadd_rn(a, -1.0f / 3.0f);
    const float c = rsqrtf(__fmul_rn(9.0f, d));
    const float x = __fmul_rn(c, r.z);
    const float t = __fadd_rn(1.0f, x);
    const float l2t = __log2f(t);                         // NaN for t < 0: such trials are rejected
    const float x2 = __fmul_rn(x, x);
    float p = 3.0f / 11.0f;
    p = __fmaf_rn(p, x, -3.0f / 10.0f);
    p = __fmaf_rn(p, x, 3.0f / 9.0f);
    p = __fmaf_rn(p, x, -3.0f / 8.0f);
    p = __fmaf_rn(p, x, 3.0f / 7.0f);
    p = __fmaf_rn(p, x, -3.0f / 6.0f);
    p = __fmaf_rn(p, x, 3.0f / 5.0f);
    p = __fmaf_rn(p, x, -3.0f / 4.0f);
    const float rhs_small = __fmul_rn(__fmul_rn(d, __fmul_rn(x2, x2)), p);
    const float v = __fmul_rn(__fmul_rn(t, t), t);
    const float rhs_large = __fmaf_rn(d, __fmaf_rn(3.0f * LN2, l2t, __fadd_rn(1.0f, -v)), r.half_zz);
    const float rhs = fabsf(x) < 0.125f ? rhs_small : rhs_large;
    const bool ok = (t > 0.0f) && (r.u < r.squeeze || r.log_u < rhs);
    return GammaTrial{__fmaf_rn(3.0f, l2t, __log2f(d)), ok};
}

// log2 of a Gamma(shape, 1) variate: the first accepted trial in counter order 0, 1, 2, ... of the
// Philox stream (x = trial, iteration, chain, purpose).  Trials 0 .. NPRE-1 arrive precomputed.
template <int NPRE>
__device__ __forceinline__ float log2_gamma(float shape, const TrialRandoms (&pre)[NPRE], uint32_t iter,
                                            uint32_t chain, uint32_t purpose, uint32_t k0, uint32_t k1, bool live)
{
    const float a = shape < 1.0f ? __fadd_rn(shape, 1.0f) : shape;
    float l2g = 0.0f, l2b = 0.0f;
    bool ok = false;
#pragma unroll
    for (int i = NPRE - 1; i >= 0; --i) {                 // lowest accepted trial wins
        const GammaTrial t = trial_finish(a, pre[i]);
        if (t.ok) { l2g = t.l2g; l2b = pre[i].l2_boost; ok = true; }
    }
    if (live && !ok) {                                    // all precomputed trials rejected (< 1 %): go on serially
        for (uint32_t trial = NPRE; trial < NPRE + 64u; ++trial) {
            const TrialRandoms r = trial_randoms(philox4x32_10(trial, iter, chain, purpose, k0, k1));
            const GammaTrial t = trial_finish(a, r);
            if (t.ok) { l2g = t.l2g; l2b = r.l2_boost; break; }
        }
    }
    if (shape < 1.0f) l2g = __fadd_rn(l2g, __fdividef(l2b, shape));
    return l2g;
}

}  // namespace brta
