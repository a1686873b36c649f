/*
 * basicrta_b200.h — C ABI of the B200-native Gibbs sampler for basicrta's
 * exponential-mixture residence-time model.
 *
 * The reference (orbeckst/basicrta) is pure Python and has no FFI of its own; the
 * boundary this library sits behind is the public Python API of
 * basicrta/gibbs.py.  Each entry point below cites the reference lines it replaces.
 * A maintainer binds it with ctypes (basicrta_b200/_cabi.py is
 * that binding).
 *
 * Conventions
 *   - plain C: pointers and sizes only, no C++/torch types;
 *   - every pointer inside brta_batch is a DEVICE pointer owned by the caller;
 *   - calls are asynchronous on the given CUDA stream (a cudaStream_t passed as
 *     void*); the library allocates nothing (brta_shard_mailbox_create excepted) and keeps
 *     no mutable global state apart from a per-thread error string;
 *   - return value: 0 = ok, < 0 = argument error (BRTA_E_*), > 0 = cudaError_t;
 *   - nothing throws across the boundary; there is no CPU fallback.
 */
#ifndef BASICRTA_B200_H
#define BASICRTA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BRTA_ABI_VERSION 6

/* compile-time geometry of the sampler kernel */
#ifndef BRTA_THREADS
#define BRTA_THREADS          128   /* threads per CTA                                  */
#endif
#define BRTA_MAX_NCOMP        255   /* labels are uint8 (gibbs.py:167-168)              */
#define BRTA_LANE_MAX_NCOMP    32   /* up to here: the team kernel (one warp lane per component); wider mixtures run
                                       one CTA per chain through a general, slower kernel (no schedule, no sharding) */
#define BRTA_TICK_LIMIT  (1u << 23) /* ticks must be < 2^23 (exact in float32)          */
#define BRTA_MAILBOX_MAX_TEAM   32   /* teams up to this size exchange through tagged mailboxes */
/* bytes of exchange workspace of a chain run by `team` CTAs (zeroed by the caller):
 * mailbox teams: 2 parities x team x 32 components x 16 B; larger teams: 3 x 32 x (8+4) B
 * of L2 atomics accumulators + an arrive counter */
#define BRTA_EXCH_BYTES(team) ((team) <= BRTA_MAILBOX_MAX_TEAM ? 2 * (team) * 32 * 16 : 1280)

/* cross-GPU exchange of a sharded chain: 2 parities x G sources x 32 components x 32 B */
#define BRTA_SHARD_MAILBOX_BYTES(g) (2 * (g) * 32 * 32)
#define BRTA_MAX_SHARDS 16
#define BRTA_IPC_HANDLE_BYTES 64    /* sizeof(cudaIpcMemHandle_t) */

/* brta_batch.flags */
#define BRTA_FLAG_EXACT        1u   /* IEEE-only arithmetic (bit-exact vs oracle/gibbs_oracle.py) */
#define BRTA_FLAG_INJECT_COEF  2u   /* teacher forcing: per-iteration (coef_c, coef_a) rows given  */
#define BRTA_FLAG_INJECT_U     4u   /* indicator uniforms given instead of the Philox stream       */
#define BRTA_FLAG_TRACE        8u   /* write (n_k, sum tick_k) of every iteration                  */
#define BRTA_FLAG_NO_TABLE    16u   /* recompute every datum's cumulative row instead of sharing the
                                       rows of equal ticks (same results; for tests / measurements) */
#define BRTA_FLAG_CTAS3       32u   /* ncomp <= 16: run the build of the kernel that is resident 3 times per SM
                                       (more registers and shared memory per CTA) instead of 4; same results.
                                       brta_gibbs_launch_info with the same flags sizes the schedule for it */

/* argument errors */
#define BRTA_E_NULL      -1
#define BRTA_E_NCOMP     -2
#define BRTA_E_RANGE     -3
#define BRTA_E_PLAN      -4
#define BRTA_E_DEVICE    -5

/* per-chain status words written by the kernel */
#define BRTA_STATUS_OK        0
#define BRTA_STATUS_NONFINITE 1   /* a coefficient row was not finite / had no live component */
#define BRTA_STATUS_TIMEOUT   2   /* a team rendezvous timed out (watchdog); results invalid   */

typedef struct brta_caps {
    int32_t abi_version;
    int32_t cc_major, cc_minor;
    int32_t sm_count;
    int32_t max_smem_per_cta;      /* opt-in dynamic shared memory, bytes */
    int32_t threads_per_cta;       /* BRTA_THREADS */
    int32_t max_ncomp;             /* BRTA_MAX_NCOMP */
    int32_t mailbox_max_team;      /* BRTA_MAILBOX_MAX_TEAM */
} brta_caps;

/* How many CTAs of the sampler kernel are co-resident per SM for (ncomp, flags,
 * slice capacity); the schedule is built for grid = sm_count * ctas_per_sm. */
typedef struct brta_launch_info {
    int32_t ctas_per_sm;
    int32_t regs_per_thread;
    int32_t static_smem;           /* bytes */
    int32_t kernel_ncomp;          /* the instantiated K the call is routed to (>= ncomp) */
} brta_launch_info;

/* One CTA's share of one chain: a contiguous range of "quads" (4 data each). A chain
 * split over team_size CTAs exchanges integer (n_k, sum tick_k) partials through its
 * exchange workspace once per iteration. Every CTA walks its tasks in ascending
 * `order`, and the members of a team all hold the chain at the same order, which is
 * what makes the in-kernel rendezvous deadlock-free. */
typedef struct brta_task {
    int32_t chain;                 /* index into the per-chain arrays                  */
    int32_t team_size;             /* number of CTAs sharing this chain                */
    int32_t team_rank;             /* 0 .. team_size-1; rank 0 writes mcweights/mcrates */
    int32_t quad_begin;            /* first quad of this CTA's slice                   */
    int32_t quad_count;            /* quads in the slice                               */
    int32_t order;                 /* global position of the chain in the schedule     */
} brta_task;

/* A batch of independent chains (residues) run by ONE persistent cooperative launch.
 *
 * Replaces the loop of basicrta/gibbs.py:191-217 for every chain of the batch, i.e.
 * what ParallelGibbs.run (gibbs.py:42-88) fans out to a multiprocessing pool.
 *
 * S = (niter + 1) / thin saved rows (gibbs.py:167-170); row j/thin - 1 holds the
 * post-update (weights, rates) of iteration j and the indicators drawn in iteration j
 * (gibbs.py:214-217).
 */
typedef struct brta_batch {
    int32_t  n_chains;             /* R                                                 */
    int32_t  ncomp;                /* K (gibbs.py:133 `ncomp`)                          */
    int32_t  niter;                /* gibbs.py:133 `niter`                              */
    int32_t  thin;                 /* gibbs.py:140 `g`                                  */
    int32_t  tick_bytes;           /* 2: uint16 ticks, 4: uint32 ticks                  */
    uint32_t flags;                /* BRTA_FLAG_*                                       */
    uint64_t seed;                 /* Philox key                                        */

    /* inputs, per chain r */
    const void*     ticks;         /* residence times as integer multiples of ts; chain r
                                      occupies [tick_offset[r], tick_offset[r]+n_data[r]) ,
                                      tick_offset a multiple of 8 elements, storage padded
                                      to a multiple of 4 elements per chain               */
    const int64_t*  tick_offset;   /* [R] element offsets                                */
    const int32_t*  perm;          /* canonical order: the sampler walks a chain's data in ASCENDING
                                      TICK order (stable) -- Philox word p belongs to the datum at
                                      canonical position p, and equal ticks sit together so their
                                      cumulative rows can be shared.  `ticks` holds each chain in
                                      that order; perm[perm_offset[r] + p] is the original index of
                                      position p and routes the stored labels back, so `indicator`
                                      is in the caller's original datum order.  NULL: identity
                                      (labels stay in canonical order)                          */
    const int64_t*  perm_offset;   /* [R] element offsets into perm                        */
    const int32_t*  n_data;        /* [R] N_r                                            */
    const uint32_t* max_tick;      /* [R] largest tick of the chain: below 65536 the chain's slices sit
                                      in shared memory at 8 B per quad, else 16 B                  */
    const uint32_t* chain_id;      /* [R] Philox counter word identifying the chain      */
    const float*    ts;            /* [R] time step (gibbs.py:147-151)                   */
    const float*    whyper;        /* [R,K] Dirichlet prior (gibbs.py:173)               */
    const float*    rhyper;        /* [R,K,2] Gamma prior shape, rate (gibbs.py:174)     */
    const float*    init_c;        /* [R,K] log2(w_k r_k) of the initial state (gibbs.py:186-188) */
    const float*    init_a;        /* [R,K] r_k * ts * log2(e) of the initial state      */

    /* outputs */
    double*         mcweights;     /* [R,S,K] float64 (gibbs.py:169)                     */
    double*         mcrates;       /* [R,S,K] float64 (gibbs.py:170)                     */
    uint8_t*        indicator;     /* chain r: S rows of ind_stride[r] bytes at ind_offset[r];
                                      first N_r bytes of a row are the labels (gibbs.py:167-168);
                                      ind_stride[r] = N_r gives the reference's dense [S,N] */
    const int64_t*  ind_offset;    /* [R] byte offsets (any alignment)                   */
    const int32_t*  ind_stride;    /* [R] row pitch in bytes, >= N_r                     */
    int32_t*        status;        /* [R] BRTA_STATUS_*                                  */

    /* parity hooks (may be NULL unless the matching flag is set) */
    const float*    inj_c;         /* [R,niter,K] coefficients used in iteration j at row j-1 */
    const float*    inj_a;         /* [R,niter,K]                                        */
    const float*    inj_u;         /* chain r: niter rows of 4*ceil(N_r/4) floats at inj_u_offset[r];
                                      values in [0,1) on the 2^-23 grid the Philox path produces */
    const int64_t*  inj_u_offset;  /* [R] element offsets, multiples of 4                */
    int64_t*        trace_nk;      /* [R,niter,K] n_k of every iteration                 */
    int64_t*        trace_tk;      /* [R,niter,K] sum of ticks per component             */

    /* schedule */
    const brta_task* tasks;        /* tasks of CTA b: tasks[cta_task_begin[b] .. cta_task_begin[b+1]) */
    const int32_t*   cta_task_begin; /* [grid_ctas + 1]                                  */
    int32_t          grid_ctas;    /* must equal sm_count * ctas_per_sm or less          */
    int32_t          slice_cap_quads; /* dynamic shared memory per CTA in 16-byte units: the largest
                                         task slice (a quad takes 16 B, or 8 B if max_tick < 65536)   */
    /* one chain sharded over several GPUs (config C4): n_shards > 1 requires n_chains == 1.  Each GPU
     * sweeps a contiguous range of the chain's quads (tasks carry GLOBAL quad indices; tick_offset /
     * ind_offset are shifted by the caller so that global indices address the local shard) and the GPUs
     * exchange (n_k, sum tick_k) once per iteration through tagged words written into each other's
     * memory over NVLink: the CTA of a GPU that completes its GPU's totals stores them into EVERY GPU's
     * mailbox, and every CTA reads all G shards' totals from its own GPU's mailbox.  The mailboxes are
     * peer-mapped by the caller: one process driving all GPUs (brta_enable_peer_access) or one process
     * per GPU (brta_shard_mailbox_create / _open: CUDA IPC).  The G launches must be able to run
     * concurrently; iteration numbers (the tags) must not repeat on a mailbox without zeroing it. */
    int32_t          n_shards;     /* G; 0 or 1 = not sharded                             */
    int32_t          shard_rank;   /* this GPU's position 0..G-1                          */
    void* const*     shard_mailbox;/* device array of G pointers: mailbox of every GPU as mapped into
                                      this process, each BRTA_SHARD_MAILBOX_BYTES(G) bytes, zeroed  */

    void*            exchange;     /* zeroed by the caller; chain r owns BRTA_EXCH_BYTES(team_size)
                                      bytes at exch_offset[r]                             */
    const int64_t*   exch_offset;  /* [R] byte offsets, multiples of 128                 */

    /* schedule feedback (may be NULL): task i gets the clock cycles its CTA spent from the start of an
     * iteration to the post of its partial statistics, summed over the second half of the launch's iterations
     * (the first ones start from the initial state and are not typical).  A short launch
     * with this set tells the scheduler how long each slice really takes (see engine.calibrate). */
    uint64_t*        task_cycles;  /* [number of tasks]                                  */

    /* one chain run as several consecutive launches (so that the schedule can be re-cut in between):
     * this launch runs iterations iter_begin + 1 .. iter_end of the niter (iter_end = 0 means niter),
     * starting from (init_c, init_a); iteration numbers, Philox counters and row indices stay those of
     * the whole run, so any segmentation gives the bits of a single launch.  The exchange workspace
     * must be zeroed before every launch. */
    int32_t          iter_begin;   /* iterations already done, 0 <= iter_begin < iter_end */
    int32_t          iter_end;     /* last iteration of this launch, <= niter; 0 = niter  */
    float*           final_c;      /* [R,K] may be NULL: coefficients after iteration iter_end, */
    float*           final_a;      /* [R,K] the (init_c, init_a) of the next launch             */

    /* rendezvous watchdog: a wait for team mates (or peer GPUs) longer than this many nanoseconds of
     * the device's global timer ends the chain with BRTA_STATUS_TIMEOUT instead of hanging the GPU.
     * Legitimate waits are as long as a team mate needs to finish its previous wave, so the host
     * scales it from the schedule (engine.py); 0 = 60 s. */
    uint64_t         watchdog_ns;
    int32_t          device;       /* CUDA ordinal the pointers live on; the call runs there and restores
                                      the caller's current device.  -1 = the current device */
    /* live progress (optional; ncomp <= 32, not sharded): every `progress_rows` saved rows the kernel
     * stores, per chain, the number of rows that are COMPLETE -- labels of every team member, weights and
     * rates written and visible system-wide (fence.sys on both sides of the team exchange) -- into
     * progress[r].  With `progress` in mapped pinned host memory the caller can copy finished row blocks
     * out (cudaMemcpyAsync on another stream) and write them to disk while the sweep is still running,
     * and drive a progress bar (the reference has one tqdm bar per chain, gibbs.py:191-193).  The rows after
     * the last multiple of progress_rows are complete when the launch is. */
    int32_t          progress_rows; /* 0 = off */
    int32_t*         progress;      /* [R], may be NULL */
} brta_batch;

/* Device capabilities.  Python side: Gibbs.run needs it to size the schedule. */
int brta_query(int device, brta_caps* caps);

/* Occupancy of the sampler kernel for this (ncomp, flags, slice capacity). */
int brta_gibbs_launch_info(int device, int ncomp, uint32_t flags, int slice_cap_quads,
                           brta_launch_info* info);

/* Run every chain of the batch for niter iterations: gibbs.py:186-217 (loop body) for
 * each chain; asynchronous on `stream`.  With BRTA_FLAG_INJECT_COEF and thin = 1 the same
 * call is Gibbs._sample_indicator (gibbs.py:321-334): labels drawn from stored
 * (mcweights, mcrates) rows without a parameter update. */
int brta_gibbs_run_batch(const brta_batch* batch, void* stream);

/* Fill out[4*n] with Philox4x32-10 words of counters (x0+i, c1, c2, c3) under key
 * `seed`: exposes the device generator to the known-answer tests. */
int brta_philox_fill(uint32_t* out_dev, int64_t n, uint32_t x0, uint32_t c1, uint32_t c2,
                     uint32_t c3, uint64_t seed, void* stream);

/* Let `device` read and write `peer`'s memory (cudaDeviceEnablePeerAccess; already enabled is
 * not an error).  Needed once per ordered pair before a sharded launch. */
int brta_enable_peer_access(int device, int peer);

/* MUFU.EX2 throughput probe: `blocks` CTAs of 256 threads each execute 8*iters
 * ex2.approx (plus as many FADD).  Timed by the caller with CUDA events it yields the
 * measured peak the sampler's roofline fraction is quoted against.
 * sink_dev: blocks*256 floats (never written in practice). */
int brta_mufu_probe(float* sink_dev, int blocks, int iters, void* stream);

/* Cluster membership counts of every datum from the stored label rows -- the accumulation of
 * Gibbs.cluster (basicrta/gibbs.py:264-268): for every row j < n_rows and datum i < n_data,
 *     c = cluster_of[j * ncomp + indicator[j * row_stride + i]];  if (c >= 0) counts[i * n_clusters + c] += 1;
 * indicator: uint8 rows of row_stride >= n_data bytes (the `indicator` of brta_batch, or a copy of
 * Gibbs.indicator[burnin/g:]); cluster_of: int8 [n_rows, ncomp], -1 for (row, component) pairs below the
 * weight cutoff (gibbs.py:233, 247), else the mixture label (gibbs.py:257); counts: int32
 * [n_data, n_clusters], ADDED to (zero it first).  All pointers are device pointers.  The caller
 * normalises the rows (gibbs.py:270). */
#define BRTA_PINDICATOR_MAX_CLUSTERS 32
int brta_pindicator_counts(const uint8_t* indicator, int64_t row_stride, int32_t n_rows, int32_t n_data,
                           const int8_t* cluster_of, int32_t ncomp, int32_t n_clusters,
                           int32_t* counts, void* stream);

/* Gaussian-mixture clustering of the retained (log weight, log rate) posterior samples -- the
 * sklearn.mixture.GaussianMixture(n_init=117, n_components=lmode).fit of Gibbs.cluster / process_gibbs
 * (basicrta/gibbs.py:255-256, 296), batched over residues ("problems") and restarts: one CTA per
 * (problem, restart).  covariance_type='full', two features, float64.  Initialisation = greedy k-means++ and
 * Lloyd iterations (scikit-learn's rule: squared centre shift <= kmeans_tol * mean(var(X))) on the Philox
 * stream (seed; draw, restart, problem_id), then an M step on the one-hot responsibilities; or, with
 * init_params, injected initial parameters (parity tests).  EM as scikit-learn's BaseMixture.fit_predict:
 * lower bound = mean log-sum-exp before the M step, stop at |change| < tol or after max_iter iterations.
 * The caller picks the restart (first strictly largest lower_bound; a restart with status
 * BRTA_GMM_ILL_DEFINED is where scikit-learn raises ValueError).  Parameter blocks are
 * [BRTA_GMM_MAX_COMPONENTS][6] doubles per (problem, restart): weight, mean x, mean y, cov xx, cov xy, cov yy.
 * All pointers are device pointers. */
#define BRTA_GMM_MAX_COMPONENTS 16
#define BRTA_GMM_THREADS       128
#define BRTA_GMM_MAX_POINTS  26624   /* per problem: 8 B of shared memory per point during the seeding */
#define BRTA_GMM_CONVERGED       0
#define BRTA_GMM_NOT_CONVERGED   1   /* max_iter reached (scikit-learn warns) */
#define BRTA_GMM_ILL_DEFINED     2   /* a covariance lost positive definiteness (scikit-learn raises ValueError) */
typedef struct brta_gmm_batch {
    int32_t n_problems, n_init, max_iter, kmeans_max_iter;   /* sklearn defaults: n_init 1 (the reference: 117), 100, 300 */
    int32_t max_points;                                      /* largest point count of a problem                        */
    int32_t class_mask;              /* which component-count classes occur: bit 0: 1-4, bit 1: 5-8, bit 2: 9-16; 0: all */
    double tol, reg_covar, kmeans_tol;                       /* sklearn defaults: 1e-3, 1e-6, 1e-4                      */
    uint64_t seed;
    const double*   x;               /* [sum M_p][2] points of all problems, concatenated                      */
    const int64_t*  offsets;         /* [n_problems + 1] first point of each problem                           */
    const int32_t*  n_components;    /* [n_problems], 1..BRTA_GMM_MAX_COMPONENTS, <= M_p                       */
    const uint32_t* problem_id;      /* [n_problems] Philox counter word of each problem; NULL: its index      */
    const double*   init_params;     /* optional [n_problems][n_init][16][6]: skip the k-means initialisation  */
    double*  lower_bound;            /* [n_problems][n_init]; NaN for an ill-defined restart                    */
    int32_t* n_iter;                 /* [n_problems][n_init] EM iterations run                                  */
    int32_t* status;                 /* [n_problems][n_init] BRTA_GMM_*                                         */
    double*  params;                 /* [n_problems][n_init][16][6] fitted parameters                           */
    double*  init_out;               /* optional, same shape: the initial parameters the EM started from        */
} brta_gmm_batch;
int brta_gmm_fit_batch(const brta_gmm_batch* batch, void* stream);

/* GaussianMixture.predict (gibbs.py:257): labels[i] = argmax_k log w_k + log N(x_i | mu_k, Sigma_k), first
 * maximum on ties; params: [n_problems][16][6] (one block per problem), labels: uint8 [sum M_p]. */
int brta_gmm_predict(const double* x, const int64_t* offsets, int32_t n_problems, int32_t max_points,
                     const int32_t* n_components, const double* params, uint8_t* labels, void* stream);

/* Test hook for the posterior update (gibbs.py:210-211 calls numpy's Generator.dirichlet / .gamma):
 * out[i] = log2 of a Gamma(shapes[i % n_shapes], 1) variate drawn by the sampler's own device
 * functions (Marsaglia-Tsang trials on the Philox stream (trial, i, chain, purpose), first accepted
 * trial, shape < 1 boost in log2 space).  All pointers are device pointers. */
int brta_gamma_fill(float* out_dev, int64_t n, const float* shapes_dev, int n_shapes, uint32_t chain,
                    uint32_t purpose, uint64_t seed, void* stream);

/* Mailbox of one GPU of a sharded chain when every GPU is driven by its OWN process (torchrun: one
 * rank per GPU).  create: cudaMalloc + zero BRTA_SHARD_MAILBOX_BYTES(n_shards) on `device` and export
 * a CUDA IPC handle (BRTA_IPC_HANDLE_BYTES bytes, to be sent to the peers by any host transport);
 * open: map a peer's mailbox into this process (peer access is enabled lazily); close / destroy undo
 * them.  These are the only calls of the library that own device memory. */
int brta_shard_mailbox_create(int device, int n_shards, void** dev_ptr, unsigned char* handle64);
int brta_shard_mailbox_open(int device, const unsigned char* handle64, void** dev_ptr);
int brta_shard_mailbox_clear(int device, void* dev_ptr, int n_shards, void* stream);   /* zero it: before re-running */
int brta_shard_mailbox_close(int device, void* dev_ptr);
int brta_shard_mailbox_destroy(int device, void* dev_ptr);

/* Last error message of the calling thread ("" if none). */
const char* brta_last_error(void);

int brta_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BASICRTA_B200_H */
