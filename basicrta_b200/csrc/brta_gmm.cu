// Gaussian-mixture clustering of the posterior samples, batched over residues and restarts: SURVEY.md 8(f-4).
//
// Gibbs.process_gibbs fits sklearn.mixture.GaussianMixture(n_init=117, n_components=lmode) to the retained
// (log weight, log rate) samples of one residue (basicrta/gibbs.py:255-257, 296) -- 117 k-means
// initialisations and EM runs on a few thousand 2-D points, one residue after the other, fanned out over a
// process pool by ProcessProtein.reprocess (basicrta/cluster.py:54-76).  Once the sampler takes seconds this is
// the pipeline's bottleneck (about 2 s per residue on a host core).  Here every (residue, restart) pair is one
// CTA of one launch: 400 residues x 117 restarts = 46 800 independent fits in flight.
//
// Algorithm = scikit-learn's, for covariance_type='full' in two dimensions (oracle/gmm_oracle.py restates it
// and is pinned against scikit-learn itself):
//   * initialisation: greedy k-means++ (2 + int(log k) local trials) and Lloyd iterations until the squared
//     centre shift is <= tol * mean(var(X)), then an M step on the one-hot responsibilities;
//     the random choices come from the sampler's Philox stream keyed by (seed; draw, restart, problem id)
//     instead of NumPy's RandomState -- the reference never seeds it (random_state=None), so there is no
//     stream to share; parity of the EM itself is checked from injected initial parameters;
//   * EM: responsibilities from the precision-Cholesky form, nk = sum r + 10 eps, centred covariances +
//     reg_covar, lower bound = mean log-sum-exp of the parameters before the M step, stop at |change| < tol.
//
// Arithmetic is float64 (the reference's), written so that the order of operations is fixed: thread t owns the
// points t, t + 128, ...; per-thread sums are sequential; a block sum is a butterfly over the lanes of each warp
// and then the four warps in order.  This file is compiled with -fmad=false: products and sums round
// separately, as NumPy's do, which makes the k-means part reproducible bit for bit by the oracle (every sampled
// candidate, chosen candidate and nearest centre); the EM part agrees to the accuracy of exp/log (1e-12).
//
// Second moments are accumulated around the current mean in one pass,
//     sum r (x - mu_new)(x - mu_new)^T = S2 - d S1^T - S1 d^T + R d d^T,   d = mu_new - mu_old,
// instead of a second pass over the data with the new mean.
//
// Not tensor-core work (2 x 2 matrices); float64 pipe + exp/log; the data of a residue (48 KB) stay in L1/L2.
#include <math.h>
#include <stdint.h>

#include "../../include/basicrta_b200.h"
#include "brta_host.h"
#include "brta_rng.cuh"

namespace brta {

constexpr int GM_THREADS = BRTA_GMM_THREADS;
constexpr int GM_WARPS = GM_THREADS / 32;
constexpr int GM_KMAX = BRTA_GMM_MAX_COMPONENTS;
constexpr int GM_PSTRIDE = 6;                               // (weight, mean x, mean y, cov xx, cov xy, cov yy)
constexpr uint32_t GM_PURPOSE = 0x474D4D00u;
constexpr double GM_EPS10 = 10.0 * 2.220446049250313e-16;
constexpr double GM_LOG_2PI = 1.8378770664093453;
constexpr unsigned GM_FULL = 0xffffffffu;

struct GmmShared {
    double red[GM_WARPS][GM_KMAX * 6 + 1];                  // warp partials of a block sum
    double mean[GM_KMAX][2];
    double cov[GM_KMAX][3];
    double pchol[GM_KMAX][3];                               // p00, p01, p11 of the upper-triangular factor
    double logc[GM_KMAX];                                   // log w + log det(pchol) - log(2 pi)
    double nk[GM_KMAX];
    double tot[GM_THREADS];                                 // k-means++ scan: per-thread totals, then prefixes
    double prefix[GM_THREADS];
    double scalar[4];
    int cand[4];
    int flag;
};

__device__ __forceinline__ double warp_butterfly(double v)
{
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v = v + __shfl_xor_sync(GM_FULL, v, off);
    return v;
}

// block sum of NV per-thread values; every thread returns with the totals.  Two barriers.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], int nv, GmmShared& sh)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nv) {
            const double s = warp_butterfly(v[i]);
            if (lane == 0) sh.red[warp][i] = s;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (i < nv) {
            double s = sh.red[0][i];
#pragma unroll
            for (int w = 1; w < GM_WARPS; ++w) s = s + sh.red[w][i];
            v[i] = s;
        }
    }
    __syncthreads();
}

__device__ __forceinline__ double block_sum1(double v, GmmShared& sh)
{
    double a[1] = {v};
    block_sum<1>(a, 1, sh);
    return a[0];
}

__device__ __forceinline__ double uniform53(uint64_t seed, uint32_t problem, uint32_t restart, uint32_t draw)
{
    const Words4 w = philox4x32_10(draw, restart, problem, GM_PURPOSE, (uint32_t)seed, (uint32_t)(seed >> 32));
    return ((double)(w.x >> 5) * 67108864.0 + (double)(w.y >> 6)) / 9007199254740992.0;
}

__device__ __forceinline__ double dist2(double x0, double x1, double c0, double c1)
{
    const double d0 = x0 - c0, d1 = x1 - c1;
    return d0 * d0 + d1 * d1;
}

// Cholesky factor of the precision of component k from sh.cov[k]; false if the covariance is not positive
// definite (scikit-learn raises "ill-defined empirical covariance").
__device__ __forceinline__ bool precision_cholesky(int k, GmmShared& sh)
{
    const double cxx = sh.cov[k][0], cxy = sh.cov[k][1], cyy = sh.cov[k][2];
    if (!(cxx > 0.0)) return false;
    const double l00 = sqrt(cxx);
    const double l10 = cxy / l00;
    const double d = cyy - l10 * l10;
    if (!(d > 0.0)) return false;
    const double l11 = sqrt(d);
    sh.pchol[k][0] = 1.0 / l00;
    sh.pchol[k][1] = -l10 / (l00 * l11);
    sh.pchol[k][2] = 1.0 / l11;
    return true;
}

// M step of component k = threadIdx.x from the block totals tot = (R, S1x, S1y, S2xx, S2xy, S2yy) taken
// around the shift point (sx, sy): _estimate_gaussian_parameters of sklearn/mixture/_gaussian_mixture.py.
__device__ __forceinline__ void m_step_component(int k, const double* tot, double sx, double sy, double reg,
                                                 GmmShared& sh)
{
    const double R = tot[0];
    const double nk = R + GM_EPS10;
    const double dx = (tot[1] - GM_EPS10 * sx) / nk, dy = (tot[2] - GM_EPS10 * sy) / nk;
    sh.nk[k] = nk;
    sh.mean[k][0] = sx + dx;
    sh.mean[k][1] = sy + dy;
    sh.cov[k][0] = (tot[3] - 2.0 * dx * tot[1] + R * dx * dx) / nk + reg;
    sh.cov[k][1] = (tot[4] - dx * tot[2] - dy * tot[1] + R * dx * dy) / nk;
    sh.cov[k][2] = (tot[5] - 2.0 * dy * tot[2] + R * dy * dy) / nk + reg;
}

// One pass over the points: responsibilities (HARD: one-hot of the nearest of the K points sh.mean) and the
// six moments per component around sh.mean; returns, in every thread, sum of log-sum-exp (0 if HARD).
// Afterwards thread k < K holds the totals of component k in tot[0..5].
template <int KT, bool HARD>
__device__ __forceinline__ double moments_pass(const double* __restrict__ x, int M, int K, GmmShared& sh, double* tot)
{
    double mu0[KT], mu1[KT], p00[KT], p01[KT], p11[KT], lc[KT];
    double acc[KT][6];
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        if (k < K) {
            mu0[k] = sh.mean[k][0]; mu1[k] = sh.mean[k][1];
            if (!HARD) { p00[k] = sh.pchol[k][0]; p01[k] = sh.pchol[k][1]; p11[k] = sh.pchol[k][2]; lc[k] = sh.logc[k]; }
        }
#pragma unroll
        for (int v = 0; v < 6; ++v) acc[k][v] = 0.0;
    }
    double lse_acc = 0.0;
    for (int i = threadIdx.x; i < M; i += GM_THREADS) {
        const double2 p = __ldg(reinterpret_cast<const double2*>(x) + i);
        double r[KT];
        if (HARD) {
            double best = dist2(p.x, p.y, mu0[0], mu1[0]);
            int lab = 0;
#pragma unroll
            for (int k = 1; k < KT; ++k)
                if (k < K) {
                    const double d = dist2(p.x, p.y, mu0[k], mu1[k]);
                    if (d < best) { best = d; lab = k; }
                }
#pragma unroll
            for (int k = 0; k < KT; ++k) r[k] = (k == lab) ? 1.0 : 0.0;
        } else {
            double wlp[KT];
            double m = -INFINITY;
#pragma unroll
            for (int k = 0; k < KT; ++k)
                if (k < K) {
                    const double d0 = p.x - mu0[k], d1 = p.y - mu1[k];
                    const double y0 = d0 * p00[k];
                    const double y1 = d0 * p01[k] + d1 * p11[k];
                    wlp[k] = lc[k] - 0.5 * (y0 * y0 + y1 * y1);
                    m = fmax(m, wlp[k]);
                }
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < KT; ++k)
                if (k < K) s = s + exp(wlp[k] - m);
            const double lse = m + log(s);
            lse_acc = lse_acc + lse;
#pragma unroll
            for (int k = 0; k < KT; ++k)
                if (k < K) r[k] = exp(wlp[k] - lse);
        }
#pragma unroll
        for (int k = 0; k < KT; ++k)
            if (k < K) {
                const double d0 = p.x - mu0[k], d1 = p.y - mu1[k];
                const double r0 = r[k] * d0, r1 = r[k] * d1;
                acc[k][0] = acc[k][0] + r[k];
                acc[k][1] = acc[k][1] + r0;
                acc[k][2] = acc[k][2] + r1;
                acc[k][3] = acc[k][3] + r0 * d0;
                acc[k][4] = acc[k][4] + r0 * d1;
                acc[k][5] = acc[k][5] + r1 * d1;
            }
    }
    // block sums: warp butterflies into shared memory, thread k adds the four warps of component k
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < KT; ++k)
        if (k < K) {
#pragma unroll
            for (int v = 0; v < 6; ++v) {
                const double s = warp_butterfly(acc[k][v]);
                if (lane == 0) sh.red[warp][k * 6 + v] = s;
            }
        }
    const double l = warp_butterfly(lse_acc);
    if (lane == 0) sh.red[warp][GM_KMAX * 6] = l;
    __syncthreads();
    if ((int)threadIdx.x < K) {
#pragma unroll
        for (int v = 0; v < 6; ++v) {
            double s = sh.red[0][threadIdx.x * 6 + v];
#pragma unroll
            for (int w = 1; w < GM_WARPS; ++w) s = s + sh.red[w][threadIdx.x * 6 + v];
            tot[v] = s;
        }
    }
    double lse_tot = sh.red[0][GM_KMAX * 6];
#pragma unroll
    for (int w = 1; w < GM_WARPS; ++w) lse_tot = lse_tot + sh.red[w][GM_KMAX * 6];
    return lse_tot;                                         // caller synchronises before sh.red is reused
}

// k-means++ seeding and Lloyd iterations; leaves the centres in sh.mean.  closest: M doubles of shared memory.
template <int KT>
__device__ __forceinline__ void kmeans_init(const double* __restrict__ x, int M, int K, const brta_gmm_batch& b,
                                            uint32_t pid, uint32_t restart, double* __restrict__ closest, GmmShared& sh)
{
    const int tid = threadIdx.x;
    const double2* const pts = reinterpret_cast<const double2*>(x);
    // tolerance of scikit-learn's KMeans: tol * mean of the per-feature variances
    double a[2] = {0.0, 0.0};
    for (int i = tid; i < M; i += GM_THREADS) { const double2 p = __ldg(pts + i); a[0] = a[0] + p.x; a[1] = a[1] + p.y; }
    block_sum<2>(a, 2, sh);
    const double m0 = a[0] / (double)M, m1 = a[1] / (double)M;
    a[0] = 0.0; a[1] = 0.0;
    for (int i = tid; i < M; i += GM_THREADS) {
        const double2 p = __ldg(pts + i);
        a[0] = a[0] + (p.x - m0) * (p.x - m0);
        a[1] = a[1] + (p.y - m1) * (p.y - m1);
    }
    block_sum<2>(a, 2, sh);
    const double tol = 0.5 * (a[0] / (double)M + a[1] / (double)M) * b.kmeans_tol;

    // ---- greedy k-means++ -------------------------------------------------------------------------------
    const int n_trials = 2 + (int)log((double)K);
    {
        int first = (int)(uniform53(b.seed, pid, restart, 0u) * (double)M);
        first = first < M - 1 ? first : M - 1;
        const double2 c = __ldg(pts + first);
        if (tid == 0) { sh.mean[0][0] = c.x; sh.mean[0][1] = c.y; }
        for (int i = tid; i < M; i += GM_THREADS) { const double2 p = __ldg(pts + i); closest[i] = dist2(p.x, p.y, c.x, c.y); }
    }
    for (int c = 1; c < K; ++c) {
        // scan of the squared distances in thread-major order: per-thread totals, prefixes by thread 0
        double run = 0.0;
        for (int i = tid; i < M; i += GM_THREADS) run = run + closest[i];
        sh.tot[tid] = run;
        if (tid < 4) sh.cand[tid] = M - 1;
        __syncthreads();
        if (tid == 0) {
            double pre = 0.0;
            for (int t = 0; t < GM_THREADS; ++t) { sh.prefix[t] = pre; pre = pre + sh.tot[t]; }
            sh.scalar[0] = pre;
        }
        __syncthreads();
        const double pot = sh.scalar[0];
        const double my_prefix = sh.prefix[tid], my_end = my_prefix + sh.tot[tid];
        for (int trial = 0; trial < n_trials; ++trial) {
            const double target = uniform53(b.seed, pid, restart, (uint32_t)(8 * c + trial)) * pot;
            // the owner of the first position whose running sum reaches the target
            const bool before = tid > 0 && my_prefix >= target;
            if (!before && my_end >= target && tid < M) {
                double w = 0.0;
                for (int i = tid; i < M; i += GM_THREADS) {
                    w = w + closest[i];
                    if (my_prefix + w >= target) { sh.cand[trial] = i; break; }
                }
            }
        }
        __syncthreads();
        double best_pot = INFINITY;
        int best = sh.cand[0];
        for (int trial = 0; trial < n_trials; ++trial) {
            const int cand = sh.cand[trial];
            const double2 cc = __ldg(pts + cand);
            double s = 0.0;
            for (int i = tid; i < M; i += GM_THREADS) {
                const double2 p = __ldg(pts + i);
                s = s + fmin(closest[i], dist2(p.x, p.y, cc.x, cc.y));
            }
            s = block_sum1(s, sh);
            if (s < best_pot) { best_pot = s; best = cand; }
        }
        const double2 cc = __ldg(pts + best);
        if (tid == 0) { sh.mean[c][0] = cc.x; sh.mean[c][1] = cc.y; }
        for (int i = tid; i < M; i += GM_THREADS) {
            const double2 p = __ldg(pts + i);
            closest[i] = fmin(closest[i], dist2(p.x, p.y, cc.x, cc.y));
        }
        __syncthreads();                                    // sh.cand / sh.tot are rewritten in the next round
    }
    __syncthreads();

    // ---- Lloyd ------------------------------------------------------------------------------------------
    for (int it = 0; it < b.kmeans_max_iter; ++it) {
        double c0[KT], c1[KT], acc[KT * 3];
#pragma unroll
        for (int k = 0; k < KT; ++k) {
            if (k < K) { c0[k] = sh.mean[k][0]; c1[k] = sh.mean[k][1]; }
            acc[3 * k] = 0.0; acc[3 * k + 1] = 0.0; acc[3 * k + 2] = 0.0;
        }
        for (int i = tid; i < M; i += GM_THREADS) {
            const double2 p = __ldg(pts + i);
            double best = dist2(p.x, p.y, c0[0], c1[0]);
            int lab = 0;
#pragma unroll
            for (int k = 1; k < KT; ++k)
                if (k < K) {
                    const double d = dist2(p.x, p.y, c0[k], c1[k]);
                    if (d < best) { best = d; lab = k; }
                }
#pragma unroll
            for (int k = 0; k < KT; ++k)
                if (k < K) {
                    const bool mine = lab == k;
                    acc[3 * k] = acc[3 * k] + (mine ? 1.0 : 0.0);
                    acc[3 * k + 1] = acc[3 * k + 1] + (mine ? p.x : 0.0);
                    acc[3 * k + 2] = acc[3 * k + 2] + (mine ? p.y : 0.0);
                }
        }
        block_sum<KT * 3>(acc, 3 * K, sh);                  // ends with a barrier: every thread has read sh.mean
        if (tid == 0) {
            double shift = 0.0;
            for (int k = 0; k < K; ++k) {
                double n0 = sh.mean[k][0], n1 = sh.mean[k][1];
                if (acc[3 * k] > 0.0) { n0 = acc[3 * k + 1] / acc[3 * k]; n1 = acc[3 * k + 2] / acc[3 * k]; }
                const double d0 = n0 - sh.mean[k][0], d1 = n1 - sh.mean[k][1];
                shift = shift + (d0 * d0 + d1 * d1);
                sh.mean[k][0] = n0; sh.mean[k][1] = n1;
            }
            sh.flag = shift <= tol ? 1 : 0;
        }
        __syncthreads();
        const int stop = sh.flag;
        __syncthreads();
        if (stop) break;
    }
}

template <int KT>
__device__ __forceinline__ void gmm_fit_cta(const brta_gmm_batch& b, int p, int restart, int M, int K,
                                            const double* __restrict__ x, double* closest, GmmShared& sh)
{
    const int tid = threadIdx.x;
    const size_t slot = (size_t)p * b.n_init + restart;
    double tot[6];
    bool ok = true;
    if (b.init_params) {                                    // parity mode: injected initial parameters
        if (tid < K) {
            const double* q = b.init_params + (slot * GM_KMAX + tid) * GM_PSTRIDE;
            sh.nk[tid] = q[0];
            sh.mean[tid][0] = q[1]; sh.mean[tid][1] = q[2];
            sh.cov[tid][0] = q[3]; sh.cov[tid][1] = q[4]; sh.cov[tid][2] = q[5];
        }
    } else {
        const uint32_t pid = b.problem_id ? b.problem_id[p] : (uint32_t)p;
        kmeans_init<KT>(x, M, K, b, pid, (uint32_t)restart, closest, sh);
        moments_pass<KT, true>(x, M, K, sh, tot);
        __syncthreads();
        if (tid < K) {
            m_step_component(tid, tot, sh.mean[tid][0], sh.mean[tid][1], b.reg_covar, sh);
            sh.nk[tid] = sh.nk[tid] / (double)M;            // GaussianMixture._initialize: weights = nk / n_samples
        }
    }
    __syncthreads();
    if (b.init_out && tid < K) {
        double* q = b.init_out + (slot * GM_KMAX + tid) * GM_PSTRIDE;
        q[0] = sh.nk[tid]; q[1] = sh.mean[tid][0]; q[2] = sh.mean[tid][1];
        q[3] = sh.cov[tid][0]; q[4] = sh.cov[tid][1]; q[5] = sh.cov[tid][2];
    }
    // sh.nk holds the weights from here on
    double lower = -INFINITY;
    int n_iter = 0, converged = 0;
    for (int it = 1; it <= b.max_iter + 1; ++it) {
        // parameters -> Cholesky factors and log constants
        if (tid < K) {
            const bool pd = precision_cholesky(tid, sh);
            if (pd) sh.logc[tid] = (log(sh.pchol[tid][0]) + log(sh.pchol[tid][2])) + log(sh.nk[tid]) - GM_LOG_2PI;
            if (!pd || !isfinite(sh.logc[tid]) || !isfinite(sh.mean[tid][0]) || !isfinite(sh.mean[tid][1])) sh.flag = -1;
        }
        __syncthreads();
        if (sh.flag < 0) { ok = false; break; }
        if (converged || it > b.max_iter) break;
        n_iter = it;
        const double lse_tot = moments_pass<KT, false>(x, M, K, sh, tot);
        __syncthreads();
        if (tid < K) m_step_component(tid, tot, sh.mean[tid][0], sh.mean[tid][1], b.reg_covar, sh);
        __syncthreads();
        if (tid < K) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s = s + sh.nk[k];
            tot[0] = sh.nk[tid] / s;
        }
        const double lb = lse_tot / (double)M;
        const double change = lb - lower;
        lower = lb;
        if (fabs(change) < b.tol) converged = 1;
        __syncthreads();
        if (tid < K) sh.nk[tid] = tot[0];
    }
    __syncthreads();
    if (tid < K) {
        double* q = b.params + (slot * GM_KMAX + tid) * GM_PSTRIDE;
        q[0] = sh.nk[tid]; q[1] = sh.mean[tid][0]; q[2] = sh.mean[tid][1];
        q[3] = sh.cov[tid][0]; q[4] = sh.cov[tid][1]; q[5] = sh.cov[tid][2];
    }
    if (tid == 0) {
        b.lower_bound[slot] = ok ? lower : NAN;
        b.n_iter[slot] = n_iter;
        b.status[slot] = !ok ? BRTA_GMM_ILL_DEFINED : converged ? BRTA_GMM_CONVERGED : BRTA_GMM_NOT_CONVERGED;
    }
}

// One instantiation per register-array size KT (components 1-4, 5-8, 9-16); a CTA whose problem belongs to
// another class returns at once (the host launches the classes brta_gmm_batch.class_mask names).
template <int KT, int KLO>
__global__ void __launch_bounds__(GM_THREADS)
gmm_fit_kernel(const brta_gmm_batch b)
{
    extern __shared__ __align__(16) double gm_closest[];
    __shared__ GmmShared sh;
    const int p = blockIdx.x / b.n_init, restart = blockIdx.x - p * b.n_init;
    const int64_t o = b.offsets[p];
    const int M = (int)(b.offsets[p + 1] - o);
    const int K = b.n_components[p];
    if (K < KLO || K > KT) return;
    const double* x = b.x + 2 * o;
    if (threadIdx.x == 0) sh.flag = 0;
    __syncthreads();
    gmm_fit_cta<KT>(b, p, restart, M, K, x, gm_closest, sh);
}

// GaussianMixture.predict: argmax_k of log w_k + log N(x | mu_k, Sigma_k), first maximum on ties.
__global__ void __launch_bounds__(256)
gmm_predict_kernel(const double* __restrict__ x, const int64_t* __restrict__ offsets, const int32_t* __restrict__ n_components,
                   const double* __restrict__ params, uint8_t* __restrict__ labels)
{
    __shared__ double s_mu[GM_KMAX][2], s_p[GM_KMAX][3], s_c[GM_KMAX];
    const int p = blockIdx.y;
    const int K = n_components[p];
    const int64_t o = offsets[p];
    const int M = (int)(offsets[p + 1] - o);
    if ((int)threadIdx.x < K) {
        const double* q = params + ((size_t)p * GM_KMAX + threadIdx.x) * GM_PSTRIDE;
        const double l00 = sqrt(q[3]), l10 = q[4] / l00, l11 = sqrt(q[5] - l10 * l10);
        s_mu[threadIdx.x][0] = q[1]; s_mu[threadIdx.x][1] = q[2];
        s_p[threadIdx.x][0] = 1.0 / l00; s_p[threadIdx.x][1] = -l10 / (l00 * l11); s_p[threadIdx.x][2] = 1.0 / l11;
        s_c[threadIdx.x] = (log(1.0 / l00) + log(1.0 / l11)) + log(q[0]) - GM_LOG_2PI;
    }
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const double2 pt = __ldg(reinterpret_cast<const double2*>(x) + o + i);
        double best = -INFINITY;
        int lab = 0;
        for (int k = 0; k < K; ++k) {
            const double d0 = pt.x - s_mu[k][0], d1 = pt.y - s_mu[k][1];
            const double y0 = d0 * s_p[k][0], y1 = d0 * s_p[k][1] + d1 * s_p[k][2];
            const double v = s_c[k] - 0.5 * (y0 * y0 + y1 * y1);
            if (v > best) { best = v; lab = k; }
        }
        labels[o + i] = (uint8_t)lab;
    }
}

}  // namespace brta

extern "C" int brta_gmm_fit_batch(const brta_gmm_batch* b, void* stream)
{
    if (!b) return brta::fail(BRTA_E_NULL, "brta_gmm_fit_batch: NULL batch");
    if (!b->x || !b->offsets || !b->n_components || !b->lower_bound || !b->n_iter || !b->status || !b->params)
        return brta::fail(BRTA_E_NULL, "brta_gmm_fit_batch: NULL pointer in brta_gmm_batch");
    if (b->n_problems < 0 || b->n_init < 1 || b->max_iter < 0 || b->kmeans_max_iter < 0 || b->max_points < 1)
        return brta::fail(BRTA_E_RANGE, "brta_gmm_fit_batch: n_problems >= 0, n_init >= 1, max_iter >= 0, max_points >= 1 required");
    if (b->max_points > BRTA_GMM_MAX_POINTS)
        return brta::fail(BRTA_E_RANGE, "brta_gmm_fit_batch: more than BRTA_GMM_MAX_POINTS points in one problem");
    if (b->n_problems == 0) return 0;
    if ((long long)b->n_problems * b->n_init > 2147483647LL)
        return brta::fail(BRTA_E_RANGE, "brta_gmm_fit_batch: n_problems * n_init exceeds the grid limit");
    const size_t smem = b->init_params ? 0 : (size_t)b->max_points * sizeof(double);
    typedef void (*fit_fn)(const brta_gmm_batch);
    const fit_fn fns[3] = {brta::gmm_fit_kernel<4, 1>, brta::gmm_fit_kernel<8, 5>, brta::gmm_fit_kernel<brta::GM_KMAX, 9>};
    const int mask = (b->class_mask & 7) ? (b->class_mask & 7) : 7;
    for (int c = 0; c < 3; ++c) {
        if (!(mask & (1 << c))) continue;
        cudaError_t e = cudaFuncSetAttribute(fns[c], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return brta::cuda_fail(e, "cudaFuncSetAttribute(gmm_fit_kernel)");
        fns[c]<<<(unsigned)(b->n_problems * b->n_init), brta::GM_THREADS, smem, (cudaStream_t)stream>>>(*b);
        e = cudaGetLastError();
        if (e != cudaSuccess) return brta::cuda_fail(e, "gmm_fit_kernel");
    }
    return 0;
}

extern "C" int brta_gmm_predict(const double* x, const int64_t* offsets, int32_t n_problems, int32_t max_points,
                                const int32_t* n_components, const double* params, uint8_t* labels, void* stream)
{
    if (!x || !offsets || !n_components || !params || !labels)
        return brta::fail(BRTA_E_NULL, "brta_gmm_predict: NULL pointer");
    if (n_problems < 0 || n_problems > 65535 || max_points < 0)
        return brta::fail(BRTA_E_RANGE, "brta_gmm_predict: 0 <= n_problems <= 65535 required");
    if (n_problems == 0 || max_points == 0) return 0;
    int bx = (max_points + 255) / 256;
    bx = bx > 64 ? 64 : bx;
    brta::gmm_predict_kernel<<<dim3((unsigned)bx, (unsigned)n_problems), 256, 0, (cudaStream_t)stream>>>(
        x, offsets, n_components, params, labels);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return brta::cuda_fail(e, "gmm_predict_kernel");
    return 0;
}
