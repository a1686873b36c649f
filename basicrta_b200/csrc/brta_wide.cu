// Gibbs sweep for 32 < ncomp <= 255 (the reference stores labels as uint8, basicrta/gibbs.py:167-168, so it
// accepts up to 255 components; its default is 15).
//
// The main kernel (brta_sweep.cuh) maps one warp lane to one component, which ends at 32.  Wider mixtures are
// rare, so this kernel trades speed for generality: ONE CTA per chain (no team, no exchange), ticks read from
// global memory (L2-resident) every iteration, K a run-time value, the per-datum cumulative sums never stored
// -- three passes over the components per datum (max, total, inverse CDF), each repeating exactly the same
// IEEE operations in the same order, so the label is the one the oracle's single pass gives:
//   EXACT: l_k = c_k - (a_k * tick), m = max_k l_k, cum_k = cum_{k-1} + soft_exp2(l_k - m), thr = u * cum_{K-1},
//          label = #{k : cum_k <= thr} clamped to K - 1          (oracle.gibbs_oracle.draw_indicators_f32)
//   FAST:  the same with fma(-a_k, tick, c_k) and MUFU.EX2.
// Statistics are exact integers (shared-memory atomics, 64-bit tick sums), thread k draws the Dirichlet and
// rate gammas of component k with the device functions of the main kernel (brta_math.cuh) on the same Philox
// stream layout (brta_rng.cuh: purpose 1 + 4k / 2 + 4k), rows are stored as in gibbs.py:214-217.  Same ABI,
// same flags (teacher forcing, injected uniforms, traces), same canonical order / perm convention.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/basicrta_b200.h"
#include "brta_host.h"
#include "brta_math.cuh"
#include "brta_rng.cuh"

namespace brta {

constexpr int WIDE_THREADS = 256;
constexpr int WIDE_MAXK = BRTA_MAX_NCOMP;

template <bool EXACT>
__device__ __forceinline__ float wide_logit(float2 ca, float tick)
{
    return EXACT ? __fsub_rn(ca.x, __fmul_rn(ca.y, tick)) : __fmaf_rn(-ca.y, tick, ca.x);
}
template <bool EXACT>
__device__ __forceinline__ float wide_term(float l, float m)
{
    return EXACT ? soft_exp2(__fsub_rn(l, m)) : fast_exp2(__fsub_rn(l, m));
}

template <bool EXACT>
__device__ __forceinline__ int wide_label(float tick, float f12, const float2* __restrict__ s_coef, int K)
{
    float m = -INFINITY;
    for (int k = 0; k < K; ++k) m = fmaxf(m, wide_logit<EXACT>(s_coef[k], tick));
    float total = 0.0f;
    for (int k = 0; k < K; ++k) total = __fadd_rn(total, wide_term<EXACT>(wide_logit<EXACT>(s_coef[k], tick), m));
    const float thr = __fmaf_rn(f12, total, -total);      // u * total rounded once (f12 = 1 + u exactly)
    float cum = 0.0f;
    int label = 0;
    for (int k = 0; k < K; ++k) {
        cum = __fadd_rn(cum, wide_term<EXACT>(wide_logit<EXACT>(s_coef[k], tick), m));
        label += (cum <= thr) ? 1 : 0;
    }
    return min(label, K - 1);
}

__device__ __forceinline__ bool wide_coef_ok(float2 ca)
{
    return (ca.x == ca.x) && (ca.x < INFINITY) && (ca.y >= 0.0f) && (ca.y < INFINITY);
}

template <bool EXACT>
__global__ void __launch_bounds__(WIDE_THREADS) gibbs_wide_kernel(const __grid_constant__ brta_batch b)
{
    __shared__ float2 s_coef[WIDE_MAXK + 1];
    __shared__ unsigned s_cnt[WIDE_MAXK + 1];
    __shared__ unsigned long long s_sum[WIDE_MAXK + 1];
    __shared__ float s_l2y[WIDE_MAXK + 1];
    __shared__ float s_red[WIDE_THREADS / 32];
    __shared__ unsigned s_bad;

    const int tid = threadIdx.x;
    const int r = blockIdx.x;
    if (r >= b.n_chains) return;
    const int K = b.ncomp;
    const int niter = b.niter, thin = b.thin;
    const int rows = (niter + 1) / thin;
    const int j_begin = b.iter_begin;
    const int j_end = b.iter_end > 0 ? b.iter_end : niter;
    const bool inject_coef = (b.flags & BRTA_FLAG_INJECT_COEF) != 0;
    const bool inject_u = (b.flags & BRTA_FLAG_INJECT_U) != 0;
    const bool trace = (b.flags & BRTA_FLAG_TRACE) != 0;
    const uint32_t key0 = (uint32_t)b.seed, key1 = (uint32_t)(b.seed >> 32);
    const int n_data = b.n_data[r];
    const int nq = (n_data + 3) / 4;
    const uint32_t chain_id = b.chain_id[r];
    const float ts = b.ts[r];
    const int64_t tick_off = b.tick_offset[r];
    const int ind_stride = b.ind_stride[r];
    uint8_t* const ind_base = b.indicator + b.ind_offset[r];
    const int32_t* const perm = b.perm ? b.perm + b.perm_offset[r] : nullptr;
    const float* const inj_u_base = inject_u ? b.inj_u + b.inj_u_offset[r] : nullptr;
    const size_t u_pitch = (size_t)nq * 4;
    const bool own = tid < K;

    bool bad = false;
    if (own) {
        float2 ca;
        if (inject_coef) {
            const size_t o = ((size_t)r * niter + j_begin) * K + tid;
            ca = make_float2(b.inj_c[o], b.inj_a[o]);
            bad |= !wide_coef_ok(ca);
        } else {
            ca = make_float2(b.init_c[(size_t)r * K + tid], b.init_a[(size_t)r * K + tid]);
        }
        s_coef[tid] = ca;
    }
    if (tid == 0) s_bad = 0;
    const float wh = own ? b.whyper[(size_t)r * K + tid] : 1.0f;
    const float rh_a = own ? b.rhyper[((size_t)r * K + tid) * 2 + 0] : 1.0f;
    const float rh_b = own ? b.rhyper[((size_t)r * K + tid) * 2 + 1] : 1.0f;

    for (int j = j_begin + 1; j <= j_end; ++j) {
        if (tid <= K) { s_cnt[tid] = 0; s_sum[tid] = 0; }
        __syncthreads();
        const bool save = (j % thin == 0);
        const int row = j / thin - 1;
        uint8_t* const ind_row = ind_base + (size_t)row * ind_stride;

        // ---- indicator draws + sufficient statistics (gibbs.py:196-207) ----
        for (int q = tid; q < nq; q += WIDE_THREADS) {
            unsigned tk[4];
            if (b.tick_bytes == 2) {
                const ushort4 raw = reinterpret_cast<const ushort4*>(static_cast<const uint16_t*>(b.ticks) + tick_off)[q];
                tk[0] = raw.x; tk[1] = raw.y; tk[2] = raw.z; tk[3] = raw.w;
            } else {
                const uint4 raw = reinterpret_cast<const uint4*>(static_cast<const uint32_t*>(b.ticks) + tick_off)[q];
                tk[0] = raw.x; tk[1] = raw.y; tk[2] = raw.z; tk[3] = raw.w;
            }
            float f12[4];
            if (inject_u) {
                const float4 uu = reinterpret_cast<const float4*>(inj_u_base + (size_t)(j - 1) * u_pitch)[q];
                f12[0] = uu.x + 1.0f; f12[1] = uu.y + 1.0f; f12[2] = uu.z + 1.0f; f12[3] = uu.w + 1.0f;
            } else {
                const Words4 w = philox4x32_10((uint32_t)q, (uint32_t)j, chain_id, 0u, key0, key1);
                f12[0] = word_to_12(w.x); f12[1] = word_to_12(w.y); f12[2] = word_to_12(w.z); f12[3] = word_to_12(w.w);
            }
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                const int i = 4 * q + d;
                if (i >= n_data) break;
                const int label = wide_label<EXACT>((float)tk[d], f12[d], s_coef, K);
                atomicAdd(&s_cnt[label], 1u);
                atomicAdd(&s_sum[label], (unsigned long long)tk[d]);
                if (save) ind_row[perm ? perm[i] : i] = (uint8_t)label;
            }
        }
        __syncthreads();

        // ---- posterior update (gibbs.py:210-211): thread k = component k ----
        const unsigned cnt = own ? s_cnt[tid] : 0u;
        const unsigned long long sum = own ? s_sum[tid] : 0ull;
        if (trace && own) {
            const size_t o = ((size_t)r * niter + (j - 1)) * K + tid;
            b.trace_nk[o] = (int64_t)cnt;
            b.trace_tk[o] = (int64_t)sum;
        }
        float2 ca = make_float2(-INFINITY, 0.0f);
        float l2y = -INFINITY, rate = 0.0f;
        if (inject_coef) {
            if (own && j < j_end) {
                const size_t o = ((size_t)r * niter + j) * K + tid;
                ca = make_float2(b.inj_c[o], b.inj_a[o]);
                bad |= !wide_coef_ok(ca);
            }
        } else if (own) {
            const float fcnt = (float)cnt;
            float l2g[2];
#pragma unroll
            for (int ty = 0; ty < 2; ++ty) {
                const uint32_t purpose = (ty == 0 ? 1u : 2u) + 4u * (uint32_t)tid;
                TrialRandoms pre[2];
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    pre[t] = trial_randoms(philox4x32_10((uint32_t)t, (uint32_t)j, chain_id, purpose, key0, key1));
                l2g[ty] = log2_gamma<2>(__fadd_rn(ty == 0 ? wh : rh_a, fcnt), pre, (uint32_t)j, chain_id, purpose, key0, key1, true);
            }
            l2y = l2g[0];
            const float l2r = __fsub_rn(l2g[1], __log2f(__fmaf_rn((float)sum, ts, rh_b)));
            rate = fast_exp2(l2r);
            ca = make_float2(__fadd_rn(l2y, l2r), __fmul_rn(rate, __fmul_rn(ts, LOG2E)));
            bad |= !wide_coef_ok(ca);
        }
        __syncthreads();                                   // everyone is done reading s_coef of iteration j
        if (own) s_coef[tid] = ca;
        if (!inject_coef && save && row < rows) {          // normalised weights of the stored row (block-wide log-sum-exp)
            float mx = warp_max(l2y);
            if ((tid & 31) == 0) s_red[tid >> 5] = mx;
            __syncthreads();
            mx = s_red[0];
#pragma unroll
            for (int w = 1; w < WIDE_THREADS / 32; ++w) mx = fmaxf(mx, s_red[w]);
            __syncthreads();
            if (own) s_l2y[tid] = l2y;
            __syncthreads();
            // sequential sum in component order: one well-defined float32 total for any K
            float tot = 0.0f;
            for (int k = 0; k < K; ++k) tot = __fadd_rn(tot, fast_exp2(__fsub_rn(s_l2y[k], mx)));
            if (own) {
                const size_t o = ((size_t)r * rows + row) * K + tid;
                b.mcweights[o] = exp2((double)__fsub_rn(__fsub_rn(l2y, mx), __log2f(tot)));
                b.mcrates[o] = (double)rate;
            }
        }
        __syncthreads();
    }
    if (b.final_c && b.final_a && own) {
        b.final_c[(size_t)r * K + tid] = s_coef[tid].x;
        b.final_a[(size_t)r * K + tid] = s_coef[tid].y;
    }
    // a usable state needs at least one live component
    bool alive = own && s_coef[tid].x > -INFINITY;
    if (__syncthreads_or(alive) == 0 && !(inject_coef)) bad = true;
    if (bad) atomicOr(&s_bad, 1u);
    __syncthreads();
    if (tid == 0 && s_bad) atomicOr(reinterpret_cast<unsigned*>(&b.status[r]), (unsigned)BRTA_STATUS_NONFINITE);
}

int launch_wide(const brta_batch& b, cudaStream_t stream)
{
    if (b.n_shards > 1) return fail(BRTA_E_RANGE, "ncomp > 32 does not support chains sharded over GPUs");
    void (*fn)(const brta_batch) = (b.flags & BRTA_FLAG_EXACT) ? gibbs_wide_kernel<true> : gibbs_wide_kernel<false>;
    fn<<<(unsigned)b.n_chains, WIDE_THREADS, 0, stream>>>(b);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "gibbs_wide_kernel");
    return 0;
}

int wide_launch_info(uint32_t flags, brta_launch_info* info)
{
    const void* fn = (flags & BRTA_FLAG_EXACT) ? (const void*)gibbs_wide_kernel<true> : (const void*)gibbs_wide_kernel<false>;
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, fn);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncGetAttributes");
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, WIDE_THREADS, 0);
    if (e != cudaSuccess) return cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    info->ctas_per_sm = per_sm;
    info->regs_per_thread = fa.numRegs;
    info->static_smem = (int32_t)fa.sharedSizeBytes;
    return 0;
}

}  // namespace brta
