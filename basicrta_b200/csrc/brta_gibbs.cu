// Host side of libbrta_gibbs.so: the C ABI of include/basicrta_b200.h (argument checks, launches) and the
// small auxiliary kernels.  The sweep kernel itself lives in brta_sweep.cuh and is instantiated once per
// supported K by brta_sweep_inst.cu.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "../../include/basicrta_b200.h"
#include "brta_host.h"
#include "brta_sweep.cuh"

namespace brta {

__global__ void philox_fill_kernel(uint32_t* out, int64_t n, uint32_t x0, uint32_t c1, uint32_t c2,
                                   uint32_t c3, uint32_t k0, uint32_t k1)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Words4 w = philox4x32_10(x0 + (uint32_t)i, c1, c2, c3, k0, k1);
    reinterpret_cast<uint4*>(out)[i] = make_uint4(w.x, w.y, w.z, w.w);
}

// MUFU.EX2 throughput probe: the measured denominator of the sampler's roofline.  Eight
// independent ex2 chains per thread, no memory traffic; the result is stored so the loop
// cannot be removed.
__global__ void __launch_bounds__(256) mufu_probe_kernel(float* sink, int iters)
{
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = -1.0e-3f * (float)(threadIdx.x + i + 1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = fast_exp2(x[i]) - 1.0f;       // 1 MUFU + 1 FADD
    }
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += x[i];
    if (acc == 12345.678f) sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// ---- host side ----------------------------------------------------------------------
thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* detail)
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
int cuda_fail(cudaError_t e, const char* where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

// Gamma-sampler probe: draw i takes shape shapes[i % n_shapes] and returns log2 of a Gamma(shape, 1)
// variate produced by exactly the device functions of the sampler's posterior update (trial_randoms,
// trial_finish, log2_gamma incl. the shape < 1 boost) on the Philox stream (trial, i, chain, purpose).
__global__ void gamma_fill_kernel(float* out, int64_t n, const float* shapes, int n_shapes, uint32_t chain,
                                  uint32_t purpose, uint32_t k0, uint32_t k1)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    TrialRandoms pre[NTRIALS];
#pragma unroll
    for (int t = 0; t < NTRIALS; ++t)
        pre[t] = trial_randoms(philox4x32_10((uint32_t)t, (uint32_t)i, chain, purpose, k0, k1));
    out[i] = log2_gamma<NTRIALS>(shapes[i % n_shapes], pre, (uint32_t)i, chain, purpose, k0, k1, true);
}

#define BRTA_FOR_EACH_K(X) X(2) X(3) X(4) X(5) X(6) X(8) X(10) X(12) X(15) X(16) X(20) X(24) X(30) X(32)
#ifdef BRTA_ONLY_K15                                       // developer builds: one instantiation
#undef BRTA_FOR_EACH_K
#define BRTA_FOR_EACH_K(X) X(15)
#endif
#define BRTA_DECLARE(KK) Variant variant_k##KK();
BRTA_FOR_EACH_K(BRTA_DECLARE)

const Variant* pick_variant(int ncomp)
{
#define BRTA_ENTRY(KK) variant_k##KK(),
    static const Variant variants[] = {BRTA_FOR_EACH_K(BRTA_ENTRY)};
    for (const Variant& v : variants)
        if (v.k >= ncomp) return &v;
    return nullptr;
}

// brta_wide.cu: 32 < ncomp <= 255
int launch_wide(const brta_batch& b, cudaStream_t stream);
int wide_launch_info(uint32_t flags, brta_launch_info* info);

// the ABI never leaves the calling thread on another device
struct DeviceGuard {
    int prev = -1;
    cudaError_t err;
    explicit DeviceGuard(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
    }
    ~DeviceGuard()
    {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

}  // namespace brta

extern "C" {

int brta_abi_version(void) { return BRTA_ABI_VERSION; }

const char* brta_last_error(void) { return brta::g_err; }

int brta_query(int device, brta_caps* caps)
{
    if (!caps) return brta::fail(BRTA_E_NULL, "brta_query: caps is NULL");
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaGetDeviceProperties");
    caps->abi_version = BRTA_ABI_VERSION;
    caps->cc_major = p.major;
    caps->cc_minor = p.minor;
    caps->sm_count = p.multiProcessorCount;
    caps->max_smem_per_cta = (int32_t)p.sharedMemPerBlockOptin;
    caps->threads_per_cta = BRTA_THREADS;
    caps->max_ncomp = BRTA_MAX_NCOMP;
    caps->mailbox_max_team = BRTA_MAILBOX_MAX_TEAM;
    return 0;
}

int brta_gibbs_launch_info(int device, int ncomp, uint32_t flags, int slice_cap_quads,
                           brta_launch_info* info)
{
    if (!info) return brta::fail(BRTA_E_NULL, "brta_gibbs_launch_info: info is NULL");
    if (ncomp < 1 || ncomp > BRTA_MAX_NCOMP) return brta::fail(BRTA_E_NCOMP, "ncomp must be in 1..255");
    if (ncomp > BRTA_LANE_MAX_NCOMP) {
        brta::DeviceGuard guard(device);
        if (guard.err != cudaSuccess) return brta::cuda_fail(guard.err, "cudaSetDevice");
        info->kernel_ncomp = ncomp;
        return brta::wide_launch_info(flags, info);
    }
    const brta::Variant* v = brta::pick_variant(ncomp);
    if (!v) return brta::fail(BRTA_E_NCOMP, "ncomp must be in 1..255");
    brta::kernel_fn fn = v->pick(flags);
    brta::DeviceGuard guard(device);
    cudaError_t e = guard.err;
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaSetDevice");
    const size_t smem = (size_t)slice_cap_quads * 16;
    e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaFuncSetAttribute(smem)");
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, (const void*)fn);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaFuncGetAttributes");
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, BRTA_THREADS, smem);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaOccupancyMaxActiveBlocksPerMultiprocessor");
    info->ctas_per_sm = per_sm;
    info->regs_per_thread = fa.numRegs;
    info->static_smem = (int32_t)fa.sharedSizeBytes;
    info->kernel_ncomp = v->k;
    return 0;
}

int brta_gibbs_run_batch(const brta_batch* batch, void* stream)
{
    if (!batch) return brta::fail(BRTA_E_NULL, "brta_gibbs_run_batch: batch is NULL");
    const brta_batch& b = *batch;
    if (b.ncomp < 1 || b.ncomp > BRTA_MAX_NCOMP) return brta::fail(BRTA_E_NCOMP, "ncomp must be in 1..255");
    const bool wide = b.ncomp > BRTA_LANE_MAX_NCOMP;       // one CTA per chain, no schedule
    const brta::Variant* v = wide ? nullptr : brta::pick_variant(b.ncomp);
    if (!wide && !v) return brta::fail(BRTA_E_NCOMP, "ncomp must be in 1..255");
    if (b.n_chains < 1 || b.niter < 1 || b.thin < 1 || (!wide && (b.grid_ctas < 1 || b.slice_cap_quads < 1)))
        return brta::fail(BRTA_E_RANGE, "n_chains, niter, thin, grid_ctas, slice_cap_quads must be >= 1");
    if (b.iter_begin < 0 || b.iter_end < 0 || b.iter_end > b.niter ||
        b.iter_begin >= (b.iter_end > 0 ? b.iter_end : b.niter))
        return brta::fail(BRTA_E_RANGE, "0 <= iter_begin < iter_end <= niter required (iter_end = 0: niter)");
    if (b.tick_bytes != 2 && b.tick_bytes != 4)
        return brta::fail(BRTA_E_RANGE, "tick_bytes must be 2 or 4");
    if (!b.max_tick) return brta::fail(BRTA_E_NULL, "max_tick is required");
    if (!b.ticks || !b.tick_offset || !b.n_data || !b.chain_id || !b.ts || !b.whyper || !b.rhyper ||
        !b.indicator || !b.ind_offset || !b.ind_stride || !b.status ||
        (!wide && (!b.tasks || !b.cta_task_begin || !b.exchange || !b.exch_offset)))
        return brta::fail(BRTA_E_NULL, "brta_gibbs_run_batch: a required pointer is NULL");
    if (b.flags & BRTA_FLAG_INJECT_COEF) {
        if (!b.inj_c || !b.inj_a) return brta::fail(BRTA_E_NULL, "INJECT_COEF needs inj_c and inj_a");
    } else if (!b.init_c || !b.init_a || !b.mcweights || !b.mcrates) {
        return brta::fail(BRTA_E_NULL, "init_c, init_a, mcweights, mcrates are required");
    }
    if ((b.flags & BRTA_FLAG_INJECT_U) && (!b.inj_u || !b.inj_u_offset))
        return brta::fail(BRTA_E_NULL, "INJECT_U needs inj_u and inj_u_offset");
    if (b.n_shards > 1) {
        if (b.n_chains != 1 || b.n_shards > BRTA_MAX_SHARDS || b.shard_rank < 0 || b.shard_rank >= b.n_shards)
            return brta::fail(BRTA_E_RANGE, "sharded launch: n_chains must be 1 and 0 <= shard_rank < n_shards <= 16");
        if (!b.shard_mailbox)
            return brta::fail(BRTA_E_NULL, "sharded launch needs shard_mailbox");
    }
    if ((b.flags & BRTA_FLAG_TRACE) && (!b.trace_nk || !b.trace_tk))
        return brta::fail(BRTA_E_NULL, "TRACE needs trace_nk and trace_tk");

    int device = b.device;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return brta::fail(BRTA_E_DEVICE, "no current CUDA device");
    brta::DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return brta::cuda_fail(guard.err, "cudaSetDevice");
    if (wide) return brta::launch_wide(b, (cudaStream_t)stream);
    brta::kernel_fn fn = v->pick(b.flags);
    const size_t smem = (size_t)b.slice_cap_quads * 16;
    cudaError_t e = cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaFuncSetAttribute(smem)");
    brta::SweepParams prm;
    prm.b = b;
    prm.rk = brta::philox_round_keys(b.seed);
    void* args[] = {(void*)&prm};
    e = cudaLaunchCooperativeKernel((const void*)fn, dim3((unsigned)b.grid_ctas), dim3(BRTA_THREADS), args,
                                    smem, (cudaStream_t)stream);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaLaunchCooperativeKernel");
    return 0;
}

int brta_enable_peer_access(int device, int peer)
{
    brta::DeviceGuard guard(device);
    cudaError_t e = guard.err;
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaSetDevice");
    int can = 0;
    e = cudaDeviceCanAccessPeer(&can, device, peer);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaDeviceCanAccessPeer");
    if (!can) return brta::fail(BRTA_E_DEVICE, "devices cannot access each other's memory (no NVLink / P2P)");
    e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaDeviceEnablePeerAccess");
    return 0;
}

int brta_mufu_probe(float* sink_dev, int blocks, int iters, void* stream)
{
    if (!sink_dev) return brta::fail(BRTA_E_NULL, "brta_mufu_probe: sink is NULL");
    if (blocks < 1 || iters < 1) return brta::fail(BRTA_E_RANGE, "brta_mufu_probe: blocks, iters >= 1");
    brta::mufu_probe_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(sink_dev, iters);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return brta::cuda_fail(e, "mufu_probe_kernel");
    return 0;
}


int brta_philox_fill(uint32_t* out_dev, int64_t n, uint32_t x0, uint32_t c1, uint32_t c2, uint32_t c3,
                     uint64_t seed, void* stream)
{
    if (!out_dev) return brta::fail(BRTA_E_NULL, "brta_philox_fill: out is NULL");
    if (n <= 0) return 0;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    brta::philox_fill_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        out_dev, n, x0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return brta::cuda_fail(e, "philox_fill_kernel");
    return 0;
}

int brta_gamma_fill(float* out_dev, int64_t n, const float* shapes_dev, int n_shapes, uint32_t chain,
                    uint32_t purpose, uint64_t seed, void* stream)
{
    if (!out_dev || !shapes_dev) return brta::fail(BRTA_E_NULL, "brta_gamma_fill: a pointer is NULL");
    if (n_shapes < 1) return brta::fail(BRTA_E_RANGE, "brta_gamma_fill: n_shapes >= 1");
    if (n <= 0) return 0;
    const int threads = 256;
    const unsigned blocks = (unsigned)((n + threads - 1) / threads);
    brta::gamma_fill_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(
        out_dev, n, shapes_dev, n_shapes, chain, purpose, (uint32_t)seed, (uint32_t)(seed >> 32));
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return brta::cuda_fail(e, "gamma_fill_kernel");
    return 0;
}

int brta_shard_mailbox_create(int device, int n_shards, void** dev_ptr, unsigned char* handle64)
{
    if (!dev_ptr || !handle64) return brta::fail(BRTA_E_NULL, "brta_shard_mailbox_create: a pointer is NULL");
    if (n_shards < 1 || n_shards > BRTA_MAX_SHARDS) return brta::fail(BRTA_E_RANGE, "n_shards must be in 1..16");
    static_assert(sizeof(cudaIpcMemHandle_t) == BRTA_IPC_HANDLE_BYTES, "IPC handle size");
    brta::DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return brta::cuda_fail(guard.err, "cudaSetDevice");
    void* p = nullptr;
    const size_t bytes = BRTA_SHARD_MAILBOX_BYTES(n_shards);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaMalloc(shard mailbox)");
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return brta::cuda_fail(e, "cudaIpcGetMemHandle"); }
    memcpy(handle64, &h, sizeof(h));
    *dev_ptr = p;
    return 0;
}

int brta_shard_mailbox_open(int device, const unsigned char* handle64, void** dev_ptr)
{
    if (!dev_ptr || !handle64) return brta::fail(BRTA_E_NULL, "brta_shard_mailbox_open: a pointer is NULL");
    brta::DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return brta::cuda_fail(guard.err, "cudaSetDevice");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaIpcOpenMemHandle");
    return 0;
}

int brta_shard_mailbox_clear(int device, void* dev_ptr, int n_shards, void* stream)
{
    if (!dev_ptr) return brta::fail(BRTA_E_NULL, "brta_shard_mailbox_clear: pointer is NULL");
    brta::DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return brta::cuda_fail(guard.err, "cudaSetDevice");
    cudaError_t e = cudaMemsetAsync(dev_ptr, 0, BRTA_SHARD_MAILBOX_BYTES(n_shards), (cudaStream_t)stream);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaMemsetAsync(shard mailbox)");
    return 0;
}

int brta_shard_mailbox_close(int device, void* dev_ptr)
{
    brta::DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return brta::cuda_fail(guard.err, "cudaSetDevice");
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaIpcCloseMemHandle");
    return 0;
}

int brta_shard_mailbox_destroy(int device, void* dev_ptr)
{
    brta::DeviceGuard guard(device);
    if (guard.err != cudaSuccess) return brta::cuda_fail(guard.err, "cudaSetDevice");
    cudaError_t e = cudaFree(dev_ptr);
    if (e != cudaSuccess) return brta::cuda_fail(e, "cudaFree(shard mailbox)");
    return 0;
}

}  // extern "C"
