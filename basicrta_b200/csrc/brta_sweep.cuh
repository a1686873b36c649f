// B200 (sm_100a) Gibbs sweep for basicrta's exponential-mixture residence-time model.
//
// Replaces the loop body of basicrta/gibbs.py:191-217 (reference: NumPy, one process per
// residue) by ONE persistent cooperative launch per batch of residues:
//
//   * a chain (residue) is owned by a *team* of CTAs chosen by the host schedule
//     (basicrta_b200/plan.py); each CTA keeps its slice of the residence times in shared
//     memory for the whole run, as the integer ticks they are (times are integer multiples
//     of ts, basicrta/contacts.py:222-229): 2 B per datum if the chain's ticks fit 16 bits;
//   * the host passes a chain in ascending-tick order (labels go back through `perm`).  Once
//     per iteration the CTA builds the K cumulative sums  sum_{k'<=k} 2^(c_k' - a_k' tick)
//     for every tick value of a window of its slice ("memoised rows", shared memory);
//   * per iteration every thread handles "quads" of 4 data: one Philox4x32-10 call gives
//     the 4 uniforms; a quad inside the window finds its labels by a binary search directly
//     on the rows (two quads in flight per thread), any other quad recomputes its row in
//     registers with the same arithmetic (MUFU.EX2; the oracle's max-subtracted IEEE-only
//     form in EXACT mode); inverse-CDF label either way (gibbs.py:196-200);
//   * sufficient statistics (n_k, sum of ticks) are exact integers (gibbs.py:203-207):
//     shared-memory atomics for every label except the one that dominates the CTA's own slice, whose
//     statistics follow by subtraction from the slice totals; branch-free (a dominant label adds to a
//     per-lane dummy slot), and in the memoised prefix count and tick offset share ONE 32-bit atomic;
//   * team members post their partials into the chain's tagged mailboxes in L2 (one
//     64-bit word per value, valid once it carries the iteration number -- no fence, no
//     flag) and every warp gathers a share of them; teams wider than 32 CTAs fall back to
//     L2 atomics + a monotonic arrive counter; one chain sharded over several GPUs adds a
//     second level over NVLink.  Every member then draws the same Dirichlet / Gamma update
//     (gibbs.py:210-211) from the same Philox key (lane k = component k, one Marsaglia-Tsang
//     trial per warp in parallel), so no broadcast is needed;
//   * every `thin`-th iteration the labels and, from team rank 0, the post-update
//     weights/rates are written (gibbs.py:214-217).
//
// Everything random is keyed by (chain, iteration, datum or component) and every statistic is an
// integer, so the result does not depend on the schedule: any team size, slicing, wave order or
// number of GPUs gives the same bits.
//
// No tensor cores: nothing here is a contraction.  The stated roofline is the MUFU (XU) pipe, one
// ex2 per (datum, component) pair; with the memoised rows the kernel is issue/latency-bound.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/basicrta_b200.h"
#include "brta_math.cuh"
#include "brta_rng.cuh"

namespace brta {

constexpr int THREADS = BRTA_THREADS;
constexpr int WARPS = THREADS / 32;
constexpr int NTRIALS = 2;                                // Marsaglia-Tsang trials evaluated straight-line per draw

// exchange workspace of one chain (BRTA_EXCH_STRIDE bytes):
//   sum[3][32] u64 @ 0, cnt[3][32] u32 @ 768, arrive u32 @ 1152
constexpr int EXCH_SUM_OFF = 0;
constexpr int EXCH_CNT_OFF = 768;
constexpr int EXCH_ARRIVE_OFF = 1152;
static_assert(EXCH_ARRIVE_OFF + 4 <= BRTA_EXCH_BYTES(BRTA_MAILBOX_MAX_TEAM + 1), "exchange layout");

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ ulonglong2 ld_relaxed_v2(const ulonglong2* p)
{
    ulonglong2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_v2(ulonglong2* p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}
// system scope: words exchanged between GPUs over NVLink (peer-mapped memory)
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Branch-free binary search in a sorted register array: the number of k with cum[k] <= thr.
// cum is non-decreasing (sums of non-negative terms), so this equals the linear count the
// oracle performs.  Level l compares against one pivot chosen from 2^(l-1) candidates by
// the earlier predicates; the latest predicate drives the outermost select so only one
// FSEL sits behind each FSETP.  K compares + K adds become ~log2(K) compares + K selects.
template <int K, int P, int NB, int OFF, int STR>
__device__ __forceinline__ float pivot_mux(const float (&cum)[K], const bool (&p)[5])
{
    if constexpr (NB == 0) {
        if constexpr (OFF < P) return cum[OFF]; else return INFINITY;
    } else {
        const float hi = pivot_mux<K, P, NB - 1, OFF + STR, 2 * STR>(cum, p);
        const float lo = pivot_mux<K, P, NB - 1, OFF, 2 * STR>(cum, p);
        return p[NB - 1] ? hi : lo;
    }
}

// level LVL of LEVELS: candidate j (bits p[0..LVL-2], p[0] most significant) sits at
// index (2j+1) * 2^(LEVELS-LVL) - 1.
template <int K, int P, int LEVELS, int LVL>
__device__ __forceinline__ void search_level(const float (&cum)[K], float thr, bool (&p)[5], int& c)
{
    if constexpr (LVL <= LEVELS) {
        const float pv = pivot_mux<K, P, LVL - 1, (1 << (LEVELS - LVL)) - 1, 1 << (LEVELS - LVL + 1)>(cum, p);
        p[LVL - 1] = pv <= thr;
        if (p[LVL - 1]) c |= (1 << (LEVELS - LVL));
        search_level<K, P, LEVELS, LVL + 1>(cum, thr, p, c);
    }
}

template <int K>
__device__ __forceinline__ int count_le(const float (&cum)[K], float thr)
{
    // tree over the first P = 2^LEVELS - 1 (padded with +inf) entries; a power-of-two K
    // keeps its last entry out of the tree and tests it directly.
    constexpr bool POW2 = (K & (K - 1)) == 0;
    constexpr int P = POW2 ? K - 1 : K;
    constexpr int LEVELS = (P >= 16) ? 5 : (P >= 8) ? 4 : (P >= 4) ? 3 : (P >= 2) ? 2 : (P >= 1) ? 1 : 0;
    bool p[5] = {false, false, false, false, false};
    int c = 0;
    search_level<K, P, LEVELS, 1>(cum, thr, p, c);
    if constexpr (POW2) c += (cum[K - 1] <= thr) ? 1 : 0;
    return c;
}

// Label of ONE datum with the max-subtracted logits of the oracle: the EXACT arithmetic, and
// the rarely-taken safe path of the FAST mode.
template <int K, bool EXACT>
__device__ __forceinline__ int draw_label_maxsub(float tick, float f12, const float2* __restrict__ s_coef, int kmax)
{
    float l[K];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float2 ca = s_coef[k];
        l[k] = EXACT ? __fsub_rn(ca.x, __fmul_rn(ca.y, tick)) : fmaf(-ca.y, tick, ca.x);
        m = fmaxf(m, l[k]);
    }
    float cum = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        cum = EXACT ? __fadd_rn(cum, soft_exp2(__fsub_rn(l[k], m))) : cum + fast_exp2(l[k] - m);
        l[k] = cum;
    }
    // u = f12 - 1 is exact, so fma(f12, total, -total) is u * total rounded once: the same bits as
    // the oracle's float32 product
    const float thr = __fmaf_rn(f12, cum, -cum);
    return min(count_le<K>(l, thr), kmax);
}

template <int K>
__device__ __noinline__ int draw_label_safe(float tick, float f12, const float2* __restrict__ s_coef, int kmax)
{
    return draw_label_maxsub<K, false>(tick, f12, s_coef, kmax);
}

// Labels of D data at once.  l[d][k] holds the logit, then the running cumulative sum.
//
// EXACT: the oracle's arithmetic (max-subtracted logits, IEEE-only operations).
// FAST : no max subtraction.  logit_k = log2(w_k r_k) - r_k t log2(e) is bounded above by
//        log2(max rate) (w <= 1, t >= 0), far from float32 overflow, so the only hazard is
//        underflow of EVERY term of a datum -- a state in which the datum would be impossible
//        under all components (it occurs, if at all, in the first burn-in sweeps).  That is
//        caught by one compare on the total and redone with the max-subtracted form.  Dropping
//        the max removes a subtract per (datum, component) pair and the max tree: 92 of the
//        476 instructions of a quad at K = 15.  Terms keep full float32 relative precision;
//        the label differs from EXACT only where u lands within rounding of a CDF boundary.
template <int K, int D, bool EXACT>
__device__ __forceinline__ unsigned draw_labels(const float (&tick)[D], const float (&f12)[D],
                                                const float2* __restrict__ s_coef, int kmax,
                                                int (&lab)[D])
{
    unsigned redo = 0u;                                    // bit d: datum d underflowed, caller redoes it
    if constexpr (EXACT) {
#pragma unroll
        for (int d = 0; d < D; ++d) lab[d] = draw_label_maxsub<K, true>(tick[d], f12[d], s_coef, kmax);
    } else {
        float l[D][K];
        float cum[D];
#pragma unroll
        for (int d = 0; d < D; ++d) cum[d] = 0.0f;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float2 ca = s_coef[k];                   // broadcast LDS
#pragma unroll
            for (int d = 0; d < D; ++d) {
                cum[d] += fast_exp2(fmaf(-ca.y, tick[d], ca.x));
                l[d][k] = cum[d];
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            const float thr = __fmaf_rn(f12[d], cum[d], -cum[d]);
            lab[d] = min(count_le<K>(l[d], thr), kmax);
            if (!(cum[d] > 8.0779357e-28f)) redo |= 1u << d;   // total < 2^-90 (or NaN): TOTAL_FLOOR
        }
    }
    return redo;
}

// data of one quad processed together by a thread (register pressure vs ILP)
#ifndef BRTA_D_SMALLK
#define BRTA_D_SMALLK 4
#endif
__host__ __device__ constexpr int data_in_flight(int k) { return k <= 16 ? BRTA_D_SMALLK : 2; }

// Kernel parameters: the caller's batch plus the Philox key schedule, both in the constant bank.
struct SweepParams {
    brta_batch b;
    RoundKeys rk;
};


// Shared-memory reductions of one datum: n_k += 1 and, 128 bytes further, the tick
// accumulator, addressed in the shared window.
__device__ __forceinline__ void red_shared_stats(uint32_t cnt_addr, uint32_t tick_bits)
{
    asm volatile("red.shared.add.u32 [%0], 1;\n\t"
                 "red.shared.add.u32 [%0+128], %1;"
                 :: "r"(cnt_addr), "r"(tick_bits) : "memory");
}
// The same, for a datum whose label offset may be the dominant one (which is not counted).
// BRTA_STATS_MODE 0: `if (off != dom_off) red_shared_stats(...)`.  ptxas turns that into a divergent branch per
//   datum (BSSY, BRA, the address rematerialised inside the branch, two ATOMS, BSYNC): nine issue slots per datum
//   whenever one lane of the warp is active, four otherwise.  (Explicitly predicated `@p red.shared` is no way
//   out: ptxas wraps EACH predicated shared-memory atomic in its own branch region.)
// BRTA_STATS_MODE 2: no branch -- a dominant label adds to a per-lane dummy slot (block 2 of s_stat, 32
//   different banks, never read) instead of being skipped: select + two unconditional atomics.
#ifndef BRTA_STATS_MODE
#define BRTA_STATS_MODE 2
#endif
__device__ __forceinline__ void red_shared_stats_ne(uint32_t off, uint32_t dom_off, uint32_t stat_addr, uint32_t tick_bits)
{
#if BRTA_STATS_MODE == 2
    const uint32_t dummy = (stat_addr & ~256u) + 512u + 4u * (threadIdx.x & 31u);
    red_shared_stats(off != dom_off ? stat_addr + off : dummy, tick_bits);
#else
    if (off != dom_off) red_shared_stats(stat_addr + off, tick_bits);
#endif
}
// Statistics of the SERVED prefix: one atomic per datum instead of two.  Every tick of the prefix lies in the
// window [lo, lo + rows) of the memoised rows, rows <= PACKED_MAX_ROWS = 256, so `tick - lo` fits 8 bits and
// (1 << 20 | tick - lo) added to one 32-bit word counts the datum in bits 20-31 and sums its offset in bits
// 0-19.  Eight accumulator sets per CTA keep both fields from overflowing: a set sees an eighth of the served
// data, i.e. <= 3968 (< 4096) for slices of up to PACKED_MAX_QUADS quads, and 3968 * 255 < 2^20.  The lead warp unpacks: n += sum of counts,
// sum of ticks += sum of offsets + lo * counts.  Shared-memory atomics are what limits the branch-free form
// (BRTA_STATS_MODE 2), so halving them is what pays.  s_stat layout (512-byte aligned, 64 words per block):
// blocks 0, 1 = (n_k, sum tick_k) of the two parities, block 2 = dummy slots, blocks 3-6 = the eight packed sets.
#ifndef BRTA_PACKED_STATS
#define BRTA_PACKED_STATS 1
#endif
constexpr int PACKED_MAX_ROWS = 256;
constexpr int PACKED_MAX_QUADS = 7680;
constexpr int STAT_BLOCKS = 7;
// BRTA_PACKED_BY_LANE 0 (default): set = 2 * warp + position in the unrolled pair.  1: set = lane & 7 (every set
// still sees an eighth of the data), so the lanes of a warp that add to the same component spread over eight
// addresses -- measured: +3 % at 50 chains per GPU, -5 % at 100, -0.5 % on C2 (profiles/r2q_*): not adopted.
#ifndef BRTA_PACKED_BY_LANE
#define BRTA_PACKED_BY_LANE 0
#endif
__device__ __forceinline__ void red_shared_packed(uint32_t off, uint32_t dom_off, uint32_t set_addr, uint32_t dummy, uint32_t rel)
{
#if BRTA_STATS_MODE == 2
    const uint32_t a = off != dom_off ? set_addr + off : dummy;
    asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(a), "r"(rel | (1u << 20)) : "memory");
#else
    if (off != dom_off) asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(set_addr + off), "r"(rel | (1u << 20)) : "memory");
#endif
}
// ---- memoised cumulative rows ---------------------------------------------------------------
// The cumulative sums of a datum depend on its tick only, and residence times are small
// integers with huge multiplicities (half of a typical slice holds fewer than a dozen distinct
// ticks once the chain is in ascending-tick order).  Once per iteration the CTA therefore
// computes the K cumulative sums for every tick value lo, lo+1, ... of its slice (up to
// TABLE_FLOATS / KP rows) with exactly the per-datum arithmetic, and a quad whose ticks all fall
// in that range reads its rows instead of recomputing them: the same bits, without the K
// FFMA + MUFU + FADD per datum.  Quads beyond the table (the sparse tail) take the direct path.
constexpr int TABLE_FLOATS = 4096;                          // 16 KB of shared memory per CTA
__host__ __device__ constexpr int table_row_floats(int k) { return (k + 3) / 4 * 4; }
// rows are laid out with an odd stride so that the same entry of neighbouring rows (what the lanes
// of a warp read during the search) falls into different banks
__host__ __device__ constexpr int table_row_stride(int k) { return table_row_floats(k) + 1; }
// rows of the table: what fits, at most 256 (tick - lo of a served datum fits 8 bits: packed statistics)
__host__ __device__ constexpr int table_rows_max(int k) { return TABLE_FLOATS / table_row_stride(k) < 256 ? TABLE_FLOATS / table_row_stride(k) : 256; }

struct TableView {
    uint32_t addr;           // shared-window byte address of row 0; row r = cumulative sums of tick lo + r
    uint32_t lo;             // first tick of the table
    uint32_t limit;          // quads with every tick < limit are served from the table (0: no table)
};

// Shared-memory accesses of the hot loop by explicit 32-bit address.  The compiler otherwise
// re-derives the shared-window base of every access from SR_CgaCtaId (an S2UR + uniform-datapath
// chain ahead of each load / reduction: a sixth of the stall samples of the first memoised kernel);
// the kernels read the base once, through opaque_u32, and keep it in a register.
__device__ __forceinline__ uint32_t opaque_u32(uint32_t x)
{
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}
template <int OFF>
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(OFF));
    return v;
}
__device__ __forceinline__ uint2 lds_v2(uint32_t addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint4 lds_v4(uint32_t addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
// ticks of quad q of a staged slice: 8 B per quad (16-bit ticks) or 16 B
template <bool T16>
__device__ __forceinline__ uint4 lds_quad(uint32_t slice_addr, int q)
{
    if constexpr (T16) {
        const uint2 raw = lds_v2(slice_addr + 8u * (uint32_t)q);
        return make_uint4(raw.x & 0xffffu, raw.x >> 16, raw.y & 0xffffu, raw.y >> 16);
    } else {
        return lds_v4(slice_addr + 16u * (uint32_t)q);
    }
}

template <int K, bool EXACT>
__device__ __forceinline__ float build_table_row(float tick, const float2* __restrict__ s_coef, float* __restrict__ row)
{
    constexpr int KP = table_row_floats(K);
    float l[KP];
    float cum = 0.0f;
    if constexpr (EXACT) {
        float m = -INFINITY;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float2 ca = s_coef[k];
            l[k] = __fsub_rn(ca.x, __fmul_rn(ca.y, tick));
            m = fmaxf(m, l[k]);
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            cum = __fadd_rn(cum, soft_exp2(__fsub_rn(l[k], m)));
            l[k] = cum;
        }
    } else {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float2 ca = s_coef[k];
            cum += fast_exp2(fmaf(-ca.y, tick, ca.x));
            l[k] = cum;
        }
    }
#pragma unroll
    for (int k = K; k < KP; ++k) l[k] = cum;
#pragma unroll
    for (int k = 0; k < KP; ++k) row[k] = l[k];
    return cum;                                            // the row total
}

// label of one datum from its memoised row; returns true if the row underflowed (FAST only).
// The inverse-CDF search runs directly on the row in shared memory: log2(KP) dependent 4-byte
// loads instead of fetching the whole row and searching it in registers.  Entries past K hold the
// total, which never compares <= thr (thr < total), so they act as +inf padding; within a warp the
// data are neighbours in tick order, hence mostly the same row and a handful of distinct addresses.
// Returns 4 * (number of row entries <= u * total), i.e. the label as a byte offset into the
// statistics array, and the row total.
template <int K>
__device__ __forceinline__ uint32_t search_row(uint32_t tick, float f12, const TableView& tab, float& total)
{
    constexpr int KP = table_row_floats(K);
    constexpr int TOP = (KP > 16) ? 32 : (KP > 8) ? 16 : (KP > 4) ? 8 : 4;      // power of two >= KP
    const uint32_t row = tab.addr + (tick - tab.lo) * (uint32_t)(4 * table_row_stride(K));
    total = lds_f32<4 * (KP - 1)>(row);
    const float thr = __fmaf_rn(f12, total, -total);
    uint32_t p = row;                                      // row + 4 * (number of entries <= thr found so far)
#define BRTA_PROBE(STEP)                                                                          \
    if constexpr (TOP / 2 >= STEP) {                                                                  \
        const float v = lds_f32<4 * (STEP - 1)>(p);                                                   \
        /* rows are KP long: probes past the end count as +inf */                                     \
        const bool in_row = (TOP == KP) || (p + 4u * (STEP - 1) < row + 4u * KP);                     \
        if (in_row && v <= thr) p += 4u * STEP;                                                       \
    }
    BRTA_PROBE(16) BRTA_PROBE(8) BRTA_PROBE(4) BRTA_PROBE(2) BRTA_PROBE(1)
#undef BRTA_PROBE
    return p - row;
}

// FAST mode: a total below 2^-90 (every term underflowed) sends the datum to draw_label_safe
constexpr float TOTAL_FLOOR = 8.0779357e-28f;

template <int K>
__device__ __forceinline__ bool label_from_table(uint32_t tick, float f12, const TableView& tab, int kmax, int& lab)
{
    float total;
    lab = min((int)(search_row<K>(tick, f12, tab, total) >> 2), kmax);
    return !(total > TOTAL_FLOOR);
}

// One quad (4 data) of the sweep: labels, statistics, optional label store.
//
// Statistics: shared-memory atomics for every label except the currently dominant one
// (`dom`, the most populated component of the previous iteration), whose statistics follow
// by subtraction from the slice totals -- exact integer arithmetic, so any choice of `dom`
// gives the same result; skipping it removes most of the same-address contention.
// integer tick -> float32 without the XU pipe: 2^23 + tick is exact for tick < 2^23
__device__ __forceinline__ float tick_to_float(uint32_t tick)
{
    return __uint_as_float(0x4B000000u | tick) - 8388608.0f;
}

template <int K, bool EXACT, bool SAVE>
__device__ __forceinline__ void sweep_quad(const uint4 tk, const float4 f12, const float2* __restrict__ s_coef,
                                           int kmax, int dom, int i0, int n_data, bool partial,
                                           uint32_t stat_addr, uint8_t* ind_row, const int32_t* __restrict__ perm,
                                           const TableView& tab)
{
    constexpr int D = data_in_flight(K);
    const uint32_t tis[4] = {tk.x, tk.y, tk.z, tk.w};
    float tks[4];
    const float fs[4] = {f12.x, f12.y, f12.z, f12.w};      // uniforms as floats in [1,2)
    int labs[4];
    unsigned redo = 0u;
    if (!partial && max(max(tk.x, tk.y), max(tk.z, tk.w)) < tab.limit) {
#pragma unroll
        for (int d = 0; d < 4; ++d)                        // memoised rows: same bits as the direct path
            if (label_from_table<K>(tis[d], fs[d], tab, kmax, labs[d])) redo |= 1u << d;
    } else {
#pragma unroll
        for (int d = 0; d < 4; ++d) tks[d] = tick_to_float(tis[d]);
#pragma unroll
        for (int h = 0; h < 4; h += D) {
            float td[D], fd[D];
            int ld[D];
#pragma unroll
            for (int d = 0; d < D; ++d) { td[d] = tks[h + d]; fd[d] = fs[h + d]; }
            redo |= draw_labels<K, D, EXACT>(td, fd, s_coef, kmax, ld) << h;
#pragma unroll
            for (int d = 0; d < D; ++d) labs[h + d] = ld[d];
        }
    }
    if (redo) {                                            // cold: after the quad, few registers are live
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (redo & (1u << d)) labs[d] = draw_label_safe<K>(tick_to_float(tis[d]), fs[d], s_coef, kmax);
    }
    if (partial) {                                         // the chain's last quad: padding is never counted
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (i0 + d >= n_data) labs[d] = dom;
    }
#pragma unroll
    for (int d = 0; d < 4; ++d)                            // ATOMS cost scales with the active lanes
        red_shared_stats_ne(4u * (uint32_t)labs[d], 4u * (uint32_t)dom, stat_addr, tis[d]);
    if (SAVE) {                                            // 1 iteration in `thin`: dense [S,N] bytes
#pragma unroll
        for (int d = 0; d < 4; ++d)
            if (i0 + d < n_data) ind_row[perm ? perm[i0 + d] : i0 + d] = (uint8_t)labs[d];
    }
}

// U quads of the served prefix at once: U Philox blocks and 4U row searches in flight per thread.
// The chain of dependent shared-memory loads of one search is what the warp waits for; with
// only four CTAs of four warps per SM the extra independent work is what fills those slots.
// Precondition (checked once per iteration when the rows are built): no row underflowed, i.e. every
// total is > TOTAL_FLOOR, so no datum of the prefix needs the safe path.
#ifndef BRTA_SERVED_UNROLL
#define BRTA_SERVED_UNROLL 2
#endif
template <int K, bool SAVE, bool T16, int U, int STRIDE = THREADS>
__device__ __forceinline__ void sweep_served(uint32_t slice_addr, int q0, int qb, int dom, uint32_t j,
                                             uint32_t chain_id, const RoundKeys& rk, uint32_t stat_addr,
                                             uint8_t* ind_row, const int32_t* __restrict__ perm, const TableView& tab)
{
    uint32_t tis[U][4];
    float fs[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int q = q0 + u * STRIDE;
        const uint4 tk = lds_quad<T16>(slice_addr, q);
        tis[u][0] = tk.x; tis[u][1] = tk.y; tis[u][2] = tk.z; tis[u][3] = tk.w;
        const Words4 w = philox4x32_10_rk((uint32_t)(qb + q), j, chain_id, 0u, rk);
        fs[u][0] = word_to_12(w.x); fs[u][1] = word_to_12(w.y); fs[u][2] = word_to_12(w.z); fs[u][3] = word_to_12(w.w);
    }
    // label as byte offset 4 * k.  With a usable row (0 < total < inf) the last real entry equals the
    // total and u * total < total, so the count never passes K - 1 and needs no clamp.
    uint32_t off[U][4];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int d = 0; d < 4; ++d) {
            float total;
            off[u][d] = search_row<K>(tis[u][d], fs[u][d], tab, total);
        }
    const uint32_t dom_off = 4u * (uint32_t)dom;
#if BRTA_PACKED_STATS
    static_assert(STRIDE == THREADS && U <= 2 && WARPS == 4, "packed statistics: eight sets, set = 2 * warp + u");
    const uint32_t stat_base = stat_addr & ~256u;
    const uint32_t dummy = stat_base + 512u + 4u * (threadIdx.x & 31u);
#if BRTA_PACKED_BY_LANE
    const uint32_t set0 = stat_base + 768u + 128u * (threadIdx.x & 7u);
    constexpr uint32_t SET_STEP = 0u;
#else
    const uint32_t set0 = stat_base + 768u + 256u * (threadIdx.x >> 5);
    constexpr uint32_t SET_STEP = 128u;
#endif
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int d = 0; d < 4; ++d)
            red_shared_packed(off[u][d], dom_off, set0 + SET_STEP * u, dummy, tis[u][d] - tab.lo);
#else
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int d = 0; d < 4; ++d)
            red_shared_stats_ne(off[u][d], dom_off, stat_addr, tis[u][d]);
#endif
    if (SAVE) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i0 = (qb + q0 + u * STRIDE) * 4;         // served quads are full quads
#pragma unroll
            for (int d = 0; d < 4; ++d) ind_row[perm ? perm[i0 + d] : i0 + d] = (uint8_t)(off[u][d] >> 2);
        }
    }
}

// The served prefix of one iteration with a FIXED share per thread (BRTA_DYNAMIC = 0): quads [0, n_served) of the
// staged slice, U at a time and the rest one at a time.  (As a separate, not inlined function -- tried to keep
// the twenty Philox round keys out of the kernel's register allocation -- ptxas serialises the eight row
// searches of a loop iteration: 0.93x.)
#ifndef BRTA_SERVED_ATTR
#define BRTA_SERVED_ATTR __forceinline__
#endif
template <int K, bool SAVE, bool T16>
__device__ BRTA_SERVED_ATTR void served_prefix_sweep(uint32_t slice_addr, int n_served, int qb, int dom, uint32_t j,
                                                 uint32_t chain_id, const RoundKeys& rk_param, uint32_t stat_addr,
                                                 uint8_t* ind_row, const int32_t* __restrict__ perm,
                                                 uint32_t tab_addr, uint32_t tab_lo)
{
    constexpr int U = BRTA_SERVED_UNROLL;
    const RoundKeys& rk = rk_param;
    TableView tab;
    tab.addr = tab_addr;
    tab.lo = tab_lo;
    tab.limit = 0;
    const int q_done = n_served / (U * THREADS) * (U * THREADS);
    for (int q = threadIdx.x; q < q_done; q += U * THREADS)
        sweep_served<K, SAVE, T16, U>(slice_addr, q, qb, dom, j, chain_id, rk, stat_addr, ind_row, perm, tab);
    if constexpr (U > 1) {
        const int q_one = n_served / THREADS * THREADS;
        for (int q = q_done + threadIdx.x; q < q_one; q += THREADS)
            sweep_served<K, SAVE, T16, 1>(slice_addr, q, qb, dom, j, chain_id, rk, stat_addr, ind_row, perm, tab);
    }
}

// One iteration's sweep over a staged slice (`slice_addr`: its shared-window address).
// `n_served`: the first n_served quads are full quads with every tick inside the memoised rows
// (with the chain in ascending-tick order: all the served quads); they run U at a time without
// any per-quad case distinction, the rest through sweep_quad.
// Work distribution inside a CTA.  Default: every thread walks a fixed share of the slice.  BRTA_DYNAMIC = 1
// (developer builds) lets the warps claim chunks of 32 * U quads from a counter in shared memory instead, so
// that a warp on a busier SM sub-partition claims less; measured 0.75x -- the claim (an atomic with a return
// value plus a shuffle per chunk) costs more than the imbalance it removes.
#ifndef BRTA_DYNAMIC
#define BRTA_DYNAMIC 0
#endif
__device__ __forceinline__ int claim_chunk(uint32_t counter_addr)
{
    unsigned v = 0;
    if ((threadIdx.x & 31) == 0)
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(v) : "r"(counter_addr) : "memory");
    return (int)__shfl_sync(FULL, v, 0);
}

template <int K, bool EXACT, bool SAVE, bool T16>
__device__ __forceinline__ void sweep_slice_t(uint32_t slice_addr, const float2* __restrict__ s_coef,
                                              int nq, int qb, int n_data, int kmax, int dom, uint32_t j,
                                              uint32_t chain_id, const RoundKeys& rk, uint32_t next_addr,
                                              const float4* __restrict__ u_row, uint32_t stat_addr,
                                              uint8_t* ind_row, const int32_t* __restrict__ perm, const TableView& tab,
                                              int n_served)
{
    // at most one quad of the whole chain is partial; find out once whether it is in this slice
    const int tail_q = ((n_data & 3) != 0) ? (n_data >> 2) - qb : -1;
#if BRTA_DYNAMIC
    constexpr int U = BRTA_SERVED_UNROLL;
    const int lane = threadIdx.x & 31;
    // served prefix: chunks of 32 * U quads, lane handles quads lane, lane + 32, ... of its chunk
    const int n_chunks = (u_row == nullptr) ? n_served / (32 * U) : 0;   // injected uniforms (tests): general loop only
    if (n_chunks > 0) {
        int c = claim_chunk(next_addr);
        while (c < n_chunks) {
            const int c_next = claim_chunk(next_addr);
            sweep_served<K, SAVE, T16, U, 32>(slice_addr, c * (32 * U) + lane, qb, dom, j, chain_id, rk, stat_addr,
                                              ind_row, perm, tab);
            c = c_next;
        }
    }
    // the rest (end of the prefix, recomputed tail, partial quad): chunks of 32 quads through the general path
    const int rest0 = n_chunks * (32 * U);
    const int n_rest = (nq - rest0 + 31) / 32;
    int c = n_rest > 0 ? claim_chunk(next_addr + 4u) : 0;
    while (c < n_rest) {
        const int c_next = claim_chunk(next_addr + 4u);
        const int q = rest0 + c * 32 + lane;
        if (q < nq) {
            const uint4 tk = lds_quad<T16>(slice_addr, q);
            float4 f12;
            if (u_row != nullptr) {
                const float4 uu = u_row[qb + q];
                f12 = make_float4(uu.x + 1.0f, uu.y + 1.0f, uu.z + 1.0f, uu.w + 1.0f);
            } else {
                const Words4 w = philox4x32_10_rk((uint32_t)(qb + q), j, chain_id, 0u, rk);
                f12 = make_float4(word_to_12(w.x), word_to_12(w.y), word_to_12(w.z), word_to_12(w.w));
            }
            sweep_quad<K, EXACT, SAVE>(tk, f12, s_coef, kmax, dom, (qb + q) * 4, n_data, q == tail_q,
                                       stat_addr, ind_row, perm, tab);
        }
        c = c_next;
    }
#else
    int q_done = 0;
    if (u_row == nullptr && n_served >= THREADS) {         // injected uniforms (tests) take the general loop
        served_prefix_sweep<K, SAVE, T16>(slice_addr, n_served, qb, dom, j, chain_id, rk, stat_addr, ind_row, perm,
                                          tab.addr, tab.lo);
        q_done = n_served / THREADS * THREADS;
    }
    for (int q = q_done + threadIdx.x; q < nq; q += THREADS) {
        const uint4 tk = lds_quad<T16>(slice_addr, q);
        float4 f12;
        if (u_row != nullptr) {
            const float4 uu = u_row[qb + q];
            f12 = make_float4(uu.x + 1.0f, uu.y + 1.0f, uu.z + 1.0f, uu.w + 1.0f);
        } else {
            const Words4 w = philox4x32_10_rk((uint32_t)(qb + q), j, chain_id, 0u, rk);
            f12 = make_float4(word_to_12(w.x), word_to_12(w.y), word_to_12(w.z), word_to_12(w.w));
        }
        sweep_quad<K, EXACT, SAVE>(tk, f12, s_coef, kmax, dom, (qb + q) * 4, n_data, q == tail_q,
                                   stat_addr, ind_row, perm, tab);
    }
#endif
}

template <int K, bool EXACT, bool SAVE>
__device__ __forceinline__ void sweep_slice(uint32_t slice_addr, bool ticks16, const float2* __restrict__ s_coef,
                                            int nq, int qb, int n_data, int kmax, int dom, uint32_t j,
                                            uint32_t chain_id, const RoundKeys& rk, uint32_t next_addr,
                                            const float4* __restrict__ u_row, uint32_t stat_addr,
                                            uint8_t* ind_row, const int32_t* __restrict__ perm, const TableView& tab,
                                            int n_served)
{
    if (ticks16)
        sweep_slice_t<K, EXACT, SAVE, true>(slice_addr, s_coef, nq, qb, n_data, kmax, dom, j, chain_id, rk, next_addr, u_row,
                                            stat_addr, ind_row, perm, tab, n_served);
    else
        sweep_slice_t<K, EXACT, SAVE, false>(slice_addr, s_coef, nq, qb, n_data, kmax, dom, j, chain_id, rk, next_addr, u_row,
                                             stat_addr, ind_row, perm, tab, n_served);
}

// length of the served prefix of a staged slice: the first quad that is partial or has a tick
// outside [lo, limit) ends it.  Block-wide; `s_scratch` holds WARPS words.
__device__ __forceinline__ int served_prefix(const unsigned char* __restrict__ s_ticks, bool ticks16, int nq, int qb,
                                             int n_data, unsigned limit, unsigned* s_scratch)
{
    unsigned first = (unsigned)nq;
    for (int q = threadIdx.x; q < nq; q += THREADS) {
        unsigned hi;
        if (ticks16) {
            const uint2 raw = reinterpret_cast<const uint2*>(s_ticks)[q];
            hi = max(max(raw.x & 0xffffu, raw.x >> 16), max(raw.y & 0xffffu, raw.y >> 16));
        } else {
            const uint4 raw = reinterpret_cast<const uint4*>(s_ticks)[q];
            hi = max(max(raw.x, raw.y), max(raw.z, raw.w));
        }
        if ((qb + q) * 4 + 3 >= n_data || hi >= limit) { first = (unsigned)q; break; }
    }
    first = __reduce_min_sync(FULL, first);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_scratch[threadIdx.x >> 5] = first;
    __syncthreads();
    unsigned r = s_scratch[0];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) r = min(r, s_scratch[w]);
    __syncthreads();
    return (int)r;
}

// Developer-only phase timing (-DBRTA_PHASE_TIMING builds a debug library): thread 0 of every CTA
// accumulates clock64 deltas of phases 0-4, lane 0 of the lead warp those of phases 5-7, into
// g_phase[blockIdx.x][8].
#ifdef BRTA_PHASE_TIMING
__device__ unsigned long long* g_phase = nullptr;
#define PHASE_DECL long long ph_t = clock64(); unsigned long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define PHASE_MARK(i) do { if (threadIdx.x == 0) { const long long now = clock64(); ph_acc[i] += (unsigned long long)(now - ph_t); ph_t = now; } } while (0)
#define PHASE_LEAD_ARG , ph_acc
#define PHASE_LEAD_PARAM , unsigned long long* ph_acc
#define PHASE_LEAD_BEGIN long long ph_t = clock64();
#define PHASE_LEAD_MARK(i) do { if ((threadIdx.x & 31) == 0) { const long long now = clock64(); ph_acc[i] += (unsigned long long)(now - ph_t); ph_t = now; } } while (0)
#define PHASE_FLUSH do { if (g_phase) { \
        if (threadIdx.x == 0) for (int i = 0; i < 5; ++i) atomicAdd(&g_phase[blockIdx.x * 8 + i], ph_acc[i]); \
        if (warp == lead && lane == 0) for (int i = 5; i < 8; ++i) atomicAdd(&g_phase[blockIdx.x * 8 + i], ph_acc[i]); } } while (0)
#else
#define PHASE_DECL
#define PHASE_MARK(i)
#define PHASE_LEAD_ARG
#define PHASE_LEAD_PARAM
#define PHASE_LEAD_BEGIN
#define PHASE_LEAD_MARK(i)
#define PHASE_FLUSH
#endif

__device__ __forceinline__ bool coef_ok(float2 ca)
{
    return (ca.x == ca.x) && (ca.x < INFINITY) && (ca.y >= 0.0f) && (ca.y < INFINITY);
}

__device__ __forceinline__ void named_barrier_sync(int id, int nthreads)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// co-resident CTAs per SM the register budget is capped for: the serial section of one
// CTA (exchange + posterior draw) is hidden by the sweeps of the others
#ifndef BRTA_MIN_CTAS_SMALLK
#define BRTA_MIN_CTAS_SMALLK (512 / THREADS)
#endif
#ifndef BRTA_MIN_CTAS_LARGEK
#define BRTA_MIN_CTAS_LARGEK (384 / THREADS)
#endif
constexpr int min_ctas(int k) { return k <= 16 ? BRTA_MIN_CTAS_SMALLK : BRTA_MIN_CTAS_LARGEK; }

// ---- waiting on other CTAs / GPUs -------------------------------------------------------------
// A waiting warp sleeps a few tens of nanoseconds between polls: a tight poll loop keeps the SM's
// memory pipe busy and slows the co-resident CTAs that are sweeping (measured: -20 % on the whole
// run).  Long waits (a team mate still finishing its previous wave) back off to 0.5 us.  The
// watchdog reads the nanosecond timer once per 4096 polls and gives up after `limit_ns`
// (brta_batch.watchdog_ns; the host scales it from the schedule's makespan), so a teammate that
// never posts cannot hang the device.
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct Watchdog {
    unsigned spins = 0;
    unsigned long long t0 = 0;
    // one unsuccessful poll: sleep `short_ns` (the first 64 times) or 500 ns, then check the clock now and then
    __device__ __forceinline__ bool expired(unsigned long long limit_ns, unsigned short_ns = 20u)
    {
        __nanosleep(spins < 64u ? short_ns : 500u);
        if ((++spins & 4095u) != 0u) return false;
        const unsigned long long now = global_timer_ns();
        if (t0 == 0ull) { t0 = now; return false; }
        return now - t0 > limit_ns;
    }
};

__device__ __forceinline__ ulonglong2 ld_relaxed_sys_v2(const unsigned long long* p)
{
    ulonglong2 v;
    asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_v2(unsigned long long* p, unsigned long long a, unsigned long long b)
{
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" :: "l"(p), "l"(a), "l"(b) : "memory");
}

// ---- the serial part of an iteration: ONE warp (lane = component) -----------------------------------
// This CTA's partial statistics -> team (and cross-GPU) exchange -> totals -> Dirichlet / Gamma update
// (gibbs.py:210-211) -> coefficients of the next iteration into shared memory.  Every member of a team
// (and every GPU of a sharded chain) computes the same update from the same Philox key, so nothing is
// broadcast.  Not inlined: its registers are not the sweep's.
#ifdef BRTA_SERIAL_NOINLINE                               // measured: a call here costs 5-7 % of the whole run
#define BRTA_SERIAL_ATTR __noinline__
#else
#define BRTA_SERIAL_ATTR __forceinline__
#endif
template <int K>
__device__ BRTA_SERIAL_ATTR void serial_part(const SweepParams& prm, const brta_task task, const int j, const int par,
                                             int dom, const unsigned slice_n, const unsigned long long slice_t,
                                             const bool save, const int row, const int rows, const unsigned tick_lo,
                                             const uint32_t key0, const uint32_t key1,
                                             unsigned (*s_stat)[64], float2* s_coef, float (*s_hyp)[32], const float* s_rhb,
                                             int* s_dom, unsigned* s_bad, unsigned* s_abort PHASE_LEAD_PARAM)
{
    constexpr bool HALF = K <= 16;
    constexpr int PASSES = HALF ? 1 : 2;
    const brta_batch& b = prm.b;
    PHASE_LEAD_BEGIN
    const int lane = threadIdx.x & 31;
    const int r = task.chain;
    const int team = task.team_size;
    const int kreal = b.ncomp;
    const int niter = b.niter;
    const int j_begin = b.iter_begin;
    const int j_end = b.iter_end > 0 ? b.iter_end : niter;
    const int n_shards = b.n_shards > 1 ? b.n_shards : 1;
    const bool sharded = n_shards > 1;                     // the chain continues on other GPUs
    const bool mailbox = !sharded && team > 1 && team <= BRTA_MAILBOX_MAX_TEAM;
    const bool inject_coef = (b.flags & BRTA_FLAG_INJECT_COEF) != 0;
    const bool trace = (b.flags & BRTA_FLAG_TRACE) != 0;
    const unsigned long long wd_limit = b.watchdog_ns ? b.watchdog_ns : 60000000000ull;
    const uint32_t chain_id = b.chain_id[r];
    const int comp = HALF ? (lane & 15) : lane;            // component this lane draws for
    const bool own = lane < kreal;                         // lane holds the statistics / coefficients of component `lane`
    const bool live = comp < kreal;
    unsigned char* const exch = static_cast<unsigned char*>(b.exchange) + b.exch_offset[r];
    // mailbox layout: slot[parity][member][32] of {tag<<32 | n_k, tag<<32 | tick sum}
    ulonglong2* const mbox = reinterpret_cast<ulonglong2*>(exch);
    // atomics layout (teams larger than the mailbox limit)
    unsigned long long* const ex_sum = reinterpret_cast<unsigned long long*>(exch + EXCH_SUM_OFF);
    unsigned* const ex_cnt = reinterpret_cast<unsigned*>(exch + EXCH_CNT_OFF);
    unsigned* const ex_arrive = reinterpret_cast<unsigned*>(exch + EXCH_ARRIVE_OFF);
    bool abort = false;

    // ---- this CTA's partials; the uncounted dominant label follows by subtraction --
    unsigned cnt = s_stat[par][lane];
    unsigned long long sum = s_stat[par][32 + lane];
    s_stat[par][lane] = 0;                                 // ready for iteration j + 2
    s_stat[par][32 + lane] = 0;
#if BRTA_PACKED_STATS
    {   // the served prefix's packed sets: bits 20-31 count, bits 0-19 sum of (tick - tick_lo)
        unsigned pc = 0, ps = 0;
#pragma unroll
        for (int set = 0; set < 8; ++set) {
            const unsigned w = s_stat[3 + (set >> 1)][(set & 1) * 32 + lane];
            s_stat[3 + (set >> 1)][(set & 1) * 32 + lane] = 0;
            pc += w >> 20;
            ps += w & 0xfffffu;
        }
        cnt += pc;
        sum += (unsigned long long)ps + (unsigned long long)pc * tick_lo;
    }
#endif
    {
        const unsigned oc = __reduce_add_sync(FULL, cnt);
        const unsigned os = __reduce_add_sync(FULL, (unsigned)sum);
        if (lane == dom) { cnt = slice_n - oc; sum = (unsigned)(slice_t - os); }
    }
    // Live progress (brta_batch.progress): the rows saved up to iteration j - 1 are published to the host once
    // every member's stores are visible system-wide.  Member side: a system fence before this iteration's post
    // (the label stores of the CTA's other threads precede it through barrier A); the publishing side follows
    // the gather below.
    const bool publish = b.progress_rows > 0 && b.progress != nullptr && j > 1 &&
                         (j - 1) % (b.thin * b.progress_rows) == 0;
    if (publish) __threadfence_system();
    Watchdog wd;
    if (mailbox) {
        if (own) {
            const unsigned long long tag = (unsigned long long)(unsigned)j << 32;
            st_relaxed_v2(&mbox[((size_t)par * team + task.team_rank) * 32 + lane], tag | cnt, tag | sum);
        }
    } else if (team > 1 || sharded) {
        // large team: L2 atomics + monotonic arrive counter (3 rotating buffers).  On a sharded chain
        // the CTA that arrives LAST holds the GPU's totals and sends them to every GPU's mailbox over
        // NVLink as tagged 64-bit words {iteration | n_k}, {iteration | sum lo}, {iteration | sum hi}.
        const int buf = j % 3;
        if (task.team_rank == 0) {
            // recycle the buffer of iteration j + 1 BEFORE arriving for j.  Nobody adds to it earlier: a member
            // starts j + 1 only after the arrive count of j is complete, which includes this arrive (ordered
            // after these stores by the fence below).  Nobody still reads its old contents (iteration j - 2):
            // this CTA is past iteration j - 1, so every member has posted j - 1, i.e. finished j - 2.
            const int nxt = (j + 1) % 3;
            ex_cnt[nxt * 32 + lane] = 0u;
            ex_sum[nxt * 32 + lane] = 0ull;
        }
        if (cnt != 0u) {
            atomicAdd(&ex_cnt[buf * 32 + lane], cnt);
            atomicAdd(&ex_sum[buf * 32 + lane], sum);
        }
        __syncwarp();
        unsigned prev = 0;
        if (lane == 0) {
            __threadfence();
            prev = atomicAdd(ex_arrive, 1u);
        }
        if (sharded) {
            prev = __shfl_sync(FULL, prev, 0);
            if (prev + 1u == (unsigned)team * (unsigned)(j - j_begin)) {   // the counter starts at 0 every launch
                __threadfence();
                const unsigned g_c = ld_relaxed_u32(&ex_cnt[buf * 32 + lane]);
                const unsigned long long g_s = ld_relaxed_u64(&ex_sum[buf * 32 + lane]);
                const unsigned long long tag = (unsigned long long)(unsigned)j << 32;
                const size_t slot = (((size_t)par * n_shards + b.shard_rank) * 32 + lane) * 4;
                if (own) {
                    for (int g = 0; g < n_shards; ++g) {
                        unsigned long long* const dst = static_cast<unsigned long long*>(b.shard_mailbox[g]) + slot;
                        st_relaxed_sys_v2(dst, tag | g_c, tag | (g_s & 0xffffffffull));
                        st_relaxed_sys_u64(dst + 2, tag | (g_s >> 32));
                    }
                }
            }
        }
    }
    // shape-independent half of the Marsaglia-Tsang trials, overlapped with the exchange
    TrialRandoms rnd[PASSES][NTRIALS];
    if (!inject_coef) {
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const uint32_t purpose = ((HALF ? (lane >> 4) : p) == 0 ? 1u : 2u) + 4u * (uint32_t)comp;
#pragma unroll
            for (int t = 0; t < NTRIALS; ++t)
                rnd[p][t] = trial_randoms(philox4x32_10_rk((uint32_t)t, (uint32_t)j, chain_id, purpose, prm.rk));
        }
    }
    PHASE_LEAD_MARK(5);
    // ---- team totals -------------------------------------------------------------------------
    unsigned tot_c = cnt;
    unsigned long long tot_s = sum;
    if (mailbox) {
        // lane = component; K <= 16: the upper half-warp takes the odd members.  All loads of a
        // round are in flight together; a word is valid once it carries this iteration's tag.
        constexpr int SPLIT = HALF ? 2 : 1;
        constexpr int ROUND = 16;
        tot_c = 0;
        tot_s = 0;
        if (live) {
            const ulonglong2* const base = &mbox[(size_t)par * team * 32 + comp];
            for (int m0 = HALF ? (lane >> 4) : 0; m0 < team && !abort; m0 += ROUND * SPLIT) {
                unsigned pend = 0;
#pragma unroll
                for (int i = 0; i < ROUND; ++i)
                    if (m0 + i * SPLIT < team) pend |= 1u << i;
                while (pend) {
                    ulonglong2 v[ROUND];
#pragma unroll
                    for (int i = 0; i < ROUND; ++i)
                        if (pend & (1u << i)) v[i] = ld_relaxed_v2(&base[(size_t)(m0 + i * SPLIT) * 32]);
#pragma unroll
                    for (int i = 0; i < ROUND; ++i) {
                        if ((pend & (1u << i)) && (unsigned)(v[i].x >> 32) == (unsigned)j &&
                            (unsigned)(v[i].y >> 32) == (unsigned)j) {
                            tot_c += (unsigned)v[i].x;
                            tot_s += (unsigned)v[i].y;
                            pend &= ~(1u << i);
                        }
                    }
                    if (pend && wd.expired(wd_limit)) { abort = true; break; }
                }
            }
        }
        if constexpr (HALF) {
            tot_c += __shfl_xor_sync(FULL, tot_c, 16);
            tot_s += __shfl_xor_sync(FULL, tot_s, 16);
        }
    } else if (sharded) {
        // every CTA of every GPU reads the G shards' totals from its GPU's own mailbox
        const unsigned long long* const mine = static_cast<const unsigned long long*>(b.shard_mailbox[b.shard_rank]);
        tot_c = 0;
        tot_s = 0;
        if (own) {
            for (int g = 0; g < n_shards && !abort; ++g) {
                const unsigned long long* const src = mine + (((size_t)par * n_shards + g) * 32 + lane) * 4;
                for (;;) {
                    const ulonglong2 w01 = ld_relaxed_sys_v2(src);
                    const unsigned long long w2 = ld_relaxed_sys_u64(src + 2);
                    if ((unsigned)(w01.x >> 32) == (unsigned)j && (unsigned)(w01.y >> 32) == (unsigned)j &&
                        (unsigned)(w2 >> 32) == (unsigned)j) {
                        tot_c += (unsigned)w01.x;
                        tot_s += (w01.y & 0xffffffffull) | (w2 << 32);
                        break;
                    }
                    if (wd.expired(wd_limit)) { abort = true; break; }
                }
            }
        }
    } else if (team > 1) {
        const int buf = j % 3;
        const unsigned target = (unsigned)team * (unsigned)(j - j_begin);
        while (ld_acquire_u32(ex_arrive) < target)
            if (wd.expired(wd_limit)) { abort = true; break; }
        tot_c = ld_relaxed_u32(&ex_cnt[buf * 32 + lane]);
        tot_s = ld_relaxed_u64(&ex_sum[buf * 32 + lane]);
    }
    if (__any_sync(FULL, abort)) {                         // rendezvous watchdog: give the chain up
        if (lane == 0) *s_abort = 1u;
        return;
    }
    if (publish && task.team_rank == 0) {                  // every member has posted iteration j: rows <= (j-1)/thin are complete
        __threadfence_system();
        if (lane == 0)
            asm volatile("st.relaxed.sys.global.u32 [%0], %1;" :: "l"(b.progress + r), "r"((unsigned)((j - 1) / b.thin)) : "memory");
    }
    PHASE_LEAD_MARK(6);

    if (trace && task.team_rank == 0 && own) {
        const size_t o = ((size_t)r * niter + (j - 1)) * kreal + lane;
        b.trace_nk[o] = (int64_t)tot_c;
        b.trace_tk[o] = (int64_t)tot_s;
    }
    // next iteration skips, in the atomics, the label that is most populated IN THIS CTA'S SLICE (the chain is in
    // ascending-tick order: later slices are dominated by slower components than the chain as a whole, and a
    // warp whose lanes all add to one shared-memory address serialises).  Any choice gives the same integers.
#ifndef BRTA_LOCAL_DOM
#define BRTA_LOCAL_DOM 1
#endif
    {
        const unsigned dom_src = BRTA_LOCAL_DOM ? cnt : tot_c;
        const unsigned keyv = own ? ((dom_src << 5) | (unsigned)(31 - lane)) : 0u;
        dom = 31 - (int)(__reduce_max_sync(FULL, keyv) & 31u);
        if (lane == 0) *s_dom = dom;
    }

    // ---- posterior update (gibbs.py:210-211): lane = component ---------------------------------
    float2 ca = make_float2(-INFINITY, 0.0f);
    bool bad = false;
    if (inject_coef) {
        if (own && j < j_end) {
            const size_t o = ((size_t)r * niter + j) * kreal + lane;
            ca = make_float2(b.inj_c[o], b.inj_a[o]);
            bad = !coef_ok(ca);
        }
    } else {
        // explicit __f*_rn operations: no contraction freedom, so every build of the kernel gives the same bits
        const float ts = b.ts[r];
        const float fcnt = (float)tot_c;
        float l2g[PASSES];
#pragma unroll
        for (int p = 0; p < PASSES; ++p) {
            const float n_c = HALF ? __shfl_sync(FULL, fcnt, comp) : fcnt;
            const uint32_t purpose = ((HALF ? (lane >> 4) : p) == 0 ? 1u : 2u) + 4u * (uint32_t)comp;
            l2g[p] = log2_gamma<NTRIALS>(__fadd_rn(s_hyp[p][lane], n_c), rnd[p], (uint32_t)j, chain_id, purpose, key0, key1, live);
        }
        // log2 of the weight gamma y_k and of the rate r_k = G_k / (b + T_k), for component `lane`
        const float l2den = __log2f(__fmaf_rn((float)tot_s, ts, s_rhb[lane]));
        const float l2y = l2g[0];
        const float l2r = __fsub_rn(HALF ? __shfl_down_sync(FULL, l2g[0], 16) : l2g[PASSES - 1], l2den);
        const float rate = fast_exp2(l2r);
        // The indicator draw is invariant to a common factor of the weights (the inverse CDF is
        // taken at u * total), so the sweep runs on the UNNORMALISED Dirichlet gammas:
        // c_k = log2(y_k r_k); the weights are normalised only for the stored rows.
        if (own) ca = make_float2(__fadd_rn(l2y, l2r), __fmul_rn(rate, __fmul_rn(ts, LOG2E)));
        // a usable row: every slope finite and >= 0, every intercept finite or -inf (a dead
        // component), at least one alive
        bad = (own && !coef_ok(ca)) || !__any_sync(FULL, own && ca.x > -INFINITY);
        if (save && task.team_rank == 0 && row < rows) {
            const float y = own ? l2y : -INFINITY;
            const float mx = warp_max(y);
            const float tot = warp_sum(own ? fast_exp2(__fsub_rn(y, mx)) : 0.0f);
            if (own) {
                const size_t o = ((size_t)r * rows + row) * kreal + lane;
                b.mcweights[o] = exp2((double)__fsub_rn(__fsub_rn(y, mx), __log2f(tot)));
                b.mcrates[o] = (double)rate;
            }
        }
    }
    s_coef[lane] = ca;
    if (__any_sync(FULL, bad) && lane == 0) *s_bad = 1u;
    PHASE_LEAD_MARK(7);
}

// Phases of one iteration of one CTA (PHASE_MARK indices of the developer build; 0-4 are timed by thread 0,
// 5-7 by lane 0 of the lead warp):
//   0 table build   1 wait C   2 sweep   3 wait A   4 serial part + wait B (as seen by thread 0)
//   5 partials + post + trial randoms   6 gather (waits for the team)   7 posterior + store
//
// Per iteration a CTA passes THREE block-wide barriers: C (memoised rows complete), A (statistics complete)
// and B (new coefficients published).  Between A and B only the LEAD warp works -- partials, team exchange,
// posterior draw (lane = component; Dirichlet and rate gammas in the two half-warps for K <= 16), stored row,
// new coefficients -- while the other warps wait at B without taking issue slots from the co-resident CTAs
// that are sweeping.  (Generation 2 let every warp run that part redundantly to save the barrier: the
// serial part got shorter, but its four copies slowed every co-resident sweep down by 45 %; 0.80x overall.)
// All of the lead warp's state lives in shared memory between iterations, so the sweep's register
// allocation carries none of it.
//
// CTAS = co-resident CTAs per SM the register budget is capped for.  K <= 16 is built twice: 4 CTAs (128
// registers per thread; best when the launch has several waves of work) and 3 CTAs (168 registers, a third
// more shared memory per slice: fewer, larger slices per chain, i.e. less per-iteration fixed work -- best for
// small batches, +5-8 % at 50 chains per GPU).  BRTA_FLAG_CTAS3 selects the latter; the host measures both.
template <int K, bool EXACT, int CTAS = min_ctas(K)>
__global__ void __launch_bounds__(THREADS, CTAS)
gibbs_sweep_kernel(const __grid_constant__ SweepParams prm)
{
    // K <= 16: lanes 0..15 draw the Dirichlet gammas, lanes 16..31 the rate gammas of component lane & 15
    constexpr bool HALF = K <= 16;
    constexpr int PASSES = HALF ? 1 : 2;
    const brta_batch& b = prm.b;
    extern __shared__ __align__(16) unsigned char smem_raw[];

    __shared__ __align__(16) float2 s_coef[32];            // {coef_c, coef_a} of the running iteration
    __shared__ float s_hyp[2][32];                         // prior shapes of the gammas a lane draws (per pass)
    __shared__ float s_rhb[32];                            // prior rate of the rate gamma (gibbs.py:174)
    // this CTA's statistics, [parity][0..31] n_k, [parity][32..63] tick sums (fit 32 bits: host-checked)
    __shared__ __align__(512) unsigned s_stat[STAT_BLOCKS][64];   // layout: see red_shared_packed
    __shared__ __align__(16) float s_table[TABLE_FLOATS + 32];  // memoised cumulative rows of the running iteration (+ probe overrun)
    const uint32_t stat_addr0 = opaque_u32((uint32_t)__cvta_generic_to_shared(&s_stat[0][0]));
    const uint32_t slice_addr = opaque_u32((uint32_t)__cvta_generic_to_shared(smem_raw));
    const uint32_t table_addr = opaque_u32((uint32_t)__cvta_generic_to_shared(s_table));
    __shared__ unsigned s_red_lo[WARPS], s_red_hi[WARPS];  // smallest / largest tick of the slice
    __shared__ unsigned long long s_red_t[WARPS];
    __shared__ unsigned s_red_n[WARPS];
    __shared__ unsigned s_bad;
    __shared__ unsigned s_abort;                           // rendezvous watchdog tripped
    __shared__ int s_dom;                                  // label left out of the atomics in the next sweep
    __shared__ unsigned s_next[2];                         // chunk counters of the running sweep (served prefix, rest)
    const uint32_t next_addr = opaque_u32((uint32_t)__cvta_generic_to_shared(&s_next[0]));

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    // The warp that runs the serial part.  Warp w of every CTA sits on SM sub-partition w % 4; rotating the
    // role over the co-resident CTAs spreads it over the four schedulers.
    const int lead = (int)((blockIdx.x + blockIdx.x / max(gridDim.x >> 2, 1u)) & (WARPS - 1));
    const uint32_t key0 = (uint32_t)b.seed;
    const uint32_t key1 = (uint32_t)(b.seed >> 32);
    const int kreal = b.ncomp;
    const int niter = b.niter;
    const int thin = b.thin;
    const int rows = (niter + 1) / thin;
    const int j_begin = b.iter_begin;                        // this launch runs iterations j_begin + 1 .. j_end
    const int j_end = b.iter_end > 0 ? b.iter_end : niter;
    const bool inject_coef = (b.flags & BRTA_FLAG_INJECT_COEF) != 0;
    const bool inject_u = (b.flags & BRTA_FLAG_INJECT_U) != 0;

    const int task_end = b.cta_task_begin[blockIdx.x + 1];
    for (int ti = b.cta_task_begin[blockIdx.x]; ti < task_end; ++ti) {
        const brta_task task = b.tasks[ti];
        const int r = task.chain;
        const int n_data = b.n_data[r];
        const int nq = task.quad_count;
        const int qb = task.quad_begin;
        const uint32_t chain_id = b.chain_id[r];
        // a chain whose largest tick fits 16 bits keeps its slice in shared memory at 8 B per quad
        const bool ticks16 = b.max_tick[r] < 65536u;
        const int64_t tick_off = b.tick_offset[r];
        const int ind_stride = b.ind_stride[r];
        uint8_t* const ind_base = b.indicator + b.ind_offset[r];
        const int32_t* const perm = b.perm ? b.perm + b.perm_offset[r] : nullptr;
        const float* const inj_u_base = inject_u ? b.inj_u + b.inj_u_offset[r] : nullptr;
        const size_t u_pitch = (size_t)((n_data + 3) / 4) * 4;

        // ---- stage the slice: integer ticks into shared memory, slice totals, tick range ----
        unsigned my_n = 0, my_lo = 0xffffffffu, my_hi = 0u;
        unsigned long long my_t = 0;
        for (int q = tid; q < nq; q += THREADS) {
            const int i0 = (qb + q) * 4;
            unsigned t0, t1, t2, t3;
            if (b.tick_bytes == 2) {
                const ushort4 raw = reinterpret_cast<const ushort4*>(
                    static_cast<const uint16_t*>(b.ticks) + tick_off)[qb + q];
                t0 = raw.x; t1 = raw.y; t2 = raw.z; t3 = raw.w;
            } else {
                const uint4 raw = reinterpret_cast<const uint4*>(
                    static_cast<const uint32_t*>(b.ticks) + tick_off)[qb + q];
                t0 = raw.x; t1 = raw.y; t2 = raw.z; t3 = raw.w;
            }
            if (i0 + 3 < n_data) {                           // full quads only: the partial quad never uses the table
                my_lo = min(my_lo, min(min(t0, t1), min(t2, t3)));
                my_hi = max(my_hi, max(max(t0, t1), max(t2, t3)));
            }
            if (i0 + 0 >= n_data) t0 = 0;
            if (i0 + 1 >= n_data) t1 = 0;
            if (i0 + 2 >= n_data) t2 = 0;
            if (i0 + 3 >= n_data) t3 = 0;
            my_n += (unsigned)min(4, max(0, n_data - i0));
            my_t += (unsigned long long)t0 + t1 + t2 + t3;
            if (ticks16) reinterpret_cast<uint2*>(smem_raw)[q] = make_uint2(t0 | (t1 << 16), t2 | (t3 << 16));
            else reinterpret_cast<uint4*>(smem_raw)[q] = make_uint4(t0, t1, t2, t3);
        }
        my_n = __reduce_add_sync(FULL, my_n);
        my_t = warp_sum_u64(my_t);
        my_lo = __reduce_min_sync(FULL, my_lo);
        my_hi = __reduce_max_sync(FULL, my_hi);
        if (lane == 0) { s_red_n[warp] = my_n; s_red_t[warp] = my_t; s_red_lo[warp] = my_lo; s_red_hi[warp] = my_hi; }
        for (int x = tid; x < STAT_BLOCKS * 64; x += THREADS) (&s_stat[0][0])[x] = 0;
        if (tid == 0) { s_bad = 0; s_abort = 0; s_dom = 0; s_next[0] = 0; s_next[1] = 0; }
        if (warp == 0) {
            // coefficients of the first iteration of this launch and the priors: lane = component
            float2 ca = make_float2(-INFINITY, 0.0f);
            if (lane < kreal) {
                if (inject_coef) {
                    const size_t o = ((size_t)r * niter + j_begin) * kreal + lane;
                    ca = make_float2(b.inj_c[o], b.inj_a[o]);
                    if (!coef_ok(ca)) atomicOr(&s_bad, 1u);
                } else {
                    ca = make_float2(b.init_c[(size_t)r * kreal + lane], b.init_a[(size_t)r * kreal + lane]);
                }
            }
            s_coef[lane] = ca;
            // s_hyp[p][lane]: prior shape of the gamma this lane draws in pass p (HALF: Dirichlet below lane 16, rate above)
            const int comp = HALF ? (lane & 15) : lane;
#pragma unroll
            for (int p = 0; p < PASSES; ++p) {
                const int type = HALF ? (lane >> 4) : p;
                float h = 1.0f;
                if (comp < kreal) h = type == 0 ? b.whyper[(size_t)r * kreal + comp] : b.rhyper[((size_t)r * kreal + comp) * 2 + 0];
                s_hyp[p][lane] = h;
            }
            s_rhb[lane] = lane < kreal ? b.rhyper[((size_t)r * kreal + lane) * 2 + 1] : 1.0f;
        }
        __syncthreads();

        // slice totals and tick range (every warp)
        unsigned slice_n = 0;
        unsigned long long slice_t = 0;
        unsigned tick_lo = 0xffffffffu, tick_hi = 0u;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            slice_n += s_red_n[w];
            slice_t += s_red_t[w];
            tick_lo = min(tick_lo, s_red_lo[w]);
            tick_hi = max(tick_hi, s_red_hi[w]);
        }
        // memoised rows cover ticks lo .. lo + table_rows - 1 of this slice
        int table_rows = 0;
        if (!(b.flags & BRTA_FLAG_NO_TABLE) && tick_lo <= tick_hi)
            table_rows = (int)min((unsigned)table_rows_max(K), tick_hi - tick_lo + 1u);
        TableView tab;
        tab.addr = table_addr;
        tab.lo = tick_lo;
        tab.limit = table_rows > 0 ? tick_lo + (unsigned)table_rows : 0u;
        // (a slice too long for the packed statistics -- developer builds with fewer CTAs per SM -- takes the general loop)
        const int n_served = (!BRTA_PACKED_STATS || nq <= PACKED_MAX_QUADS)
                                 ? served_prefix(smem_raw, ticks16, nq, qb, n_data, tab.limit, s_red_lo) : 0;

        unsigned long long busy_cycles = 0;                // schedule feedback: iteration start -> statistics complete
        PHASE_DECL
        for (int j = j_begin + 1; j <= j_end; ++j) {
            const long long iter_t0 = b.task_cycles ? clock64() : 0;
            // ---- this thread's share of the memoised rows ---------------------------------------
            bool row_bad = false;                          // FAST mode: a row whose every term underflowed
            for (int rr = tid; rr < table_rows; rr += THREADS) {
                const float total = build_table_row<K, EXACT>((float)(tick_lo + (unsigned)rr), s_coef,
                                                              s_table + (size_t)rr * table_row_stride(K));
                row_bad |= !(total > TOTAL_FLOOR);
            }
            PHASE_MARK(0);
            // C: rows complete.  If any row underflowed (early burn-in states at most) no quad takes the
            // unchecked served loop in this iteration; the general loop redoes such data with max subtraction.
            const int n_served_j = __syncthreads_or(!EXACT && row_bad) ? 0 : n_served;
            PHASE_MARK(1);
            const int par = j & 1;
            const int dom = s_dom;
            const uint32_t stat_addr = opaque_u32(stat_addr0 + 256u * (uint32_t)par);   // opaque: not rematerialised per datum
            const bool save = (j % thin == 0);
            const int row = j / thin - 1;
            const float4* const u_row = inject_u
                ? reinterpret_cast<const float4*>(inj_u_base + (size_t)(j - 1) * u_pitch) : nullptr;

            // ---- indicator draws + sufficient statistics (gibbs.py:196-207) -------------
            if (save)
                sweep_slice<K, EXACT, true>(slice_addr, ticks16, s_coef, nq, qb, n_data, kreal - 1, dom, (uint32_t)j, chain_id,
                                            prm.rk, next_addr, u_row, stat_addr, ind_base + (size_t)row * ind_stride, perm, tab, n_served_j);
            else
                sweep_slice<K, EXACT, false>(slice_addr, ticks16, s_coef, nq, qb, n_data, kreal - 1, dom, (uint32_t)j, chain_id,
                                             prm.rk, next_addr, u_row, stat_addr, nullptr, nullptr, tab, n_served_j);
            PHASE_MARK(2);
            __syncthreads();                               // A: this CTA's statistics are complete
            PHASE_MARK(3);
            if (tid == 0) { s_next[0] = 0; s_next[1] = 0; }  // for the next sweep (two barriers away)
            if (b.task_cycles && 2 * (j - j_begin) > j_end - j_begin) busy_cycles += (unsigned long long)(clock64() - iter_t0);

            if (warp == lead)
                serial_part<K>(prm, task, j, par, dom, slice_n, slice_t, save, row, rows, tick_lo, key0, key1,
                               s_stat, s_coef, s_hyp, s_rhb, &s_dom, &s_bad, &s_abort PHASE_LEAD_ARG);
            __syncthreads();                               // B: coefficients of iteration j + 1 are published
            PHASE_MARK(4);
            if (s_abort) break;                            // uniform: written before the barrier
        }
        PHASE_FLUSH;
        if (b.task_cycles && tid == 0) b.task_cycles[ti] = busy_cycles;
        if (b.final_c && b.final_a && task.team_rank == 0 && tid < kreal) {   // state for a following launch
            b.final_c[(size_t)r * kreal + tid] = s_coef[tid].x;
            b.final_a[(size_t)r * kreal + tid] = s_coef[tid].y;
        }
        if (tid == 0 && s_bad) atomicOr(reinterpret_cast<unsigned*>(&b.status[r]), (unsigned)BRTA_STATUS_NONFINITE);
        if (tid == 0 && s_abort) atomicOr(reinterpret_cast<unsigned*>(&b.status[r]), (unsigned)BRTA_STATUS_TIMEOUT);
        __syncthreads();
    }
}

typedef void (*kernel_fn)(const SweepParams);

struct Variant {
    int k;
    kernel_fn fast, exact;
    kernel_fn fast3;                                       // FAST arithmetic at 3 CTAs per SM (nullptr: `fast` already is)
    kernel_fn pick(uint32_t flags) const
    {
        if (flags & BRTA_FLAG_EXACT) return exact;
        return ((flags & BRTA_FLAG_CTAS3) && fast3) ? fast3 : fast;
    }
};

template <int K>
kernel_fn three_cta_kernel()
{
    if constexpr (min_ctas(K) > 3) return gibbs_sweep_kernel<K, false, 3>;
    else return nullptr;
}

}  // namespace brta
