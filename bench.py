#!/usr/bin/env python
"""Benchmark of the Gibbs-sampler hot path on BASELINE.json's configurations.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C1|C2|C3|C4|C5]

Default workload ("C2", the configuration BASELINE.json's metric is quoted on): the
full-membrane sweep -- 400 residues x 1e4-1e5 residence times each, ncomp = 15, niter = 110 000,
synthetic multi-exponential times on the 0.1 ns grid, seeds 1000 + r.  One *step* = the whole sweep:
every residue's 110 000 Gibbs iterations.  With N > 1 (torchrun, one rank per GPU) the residues are
sharded over the ranks (strong scaling, no collective on the data path: chains are independent).
``--config`` selects the other configurations BASELINE.json names (C1 single residue, C3 the
2 000-chain multi-cutoff ensemble, C4 the giant single residue -- with N > 1 its times are sharded over
the ranks and the kernels exchange (n_k, sum tick_k) over NVLink every iteration --, C5 the ncomp = 30
stress case); they print the same JSON line for their workload.

Metric: Gibbs indicator draws / s = sum_r N_r * K * niter / seconds (whole job).

* ``value``        device-timed (CUDA events around each step's launch, inputs resident in HBM)
* ``e2e``          the same through the host-buffer path: H2D of the tick arrays from pinned memory +
                   kernel + D2H of mcweights / mcrates / indicator into pinned memory
* ``e2e_api``      (C2 only) wall clock of the PUBLIC API: ``basicrta_b200.gibbs.dispatch`` from NumPy
                   ``times`` arrays to the reference's per-residue pickles on disk -- sorting, planning,
                   calibration launches, upload, sweep, download, pickling, everything -- on N GPUs
                   driven by one process (one host thread per GPU), to tmpfs and to the box's disk
* ``c4``           (C2 runs only) the giant single residue (N = 1e6, K = 15) on the same N GPUs, one rank
                   per GPU: microseconds per iteration and whether the result equals the 1-GPU run bit for bit
* ``roofline``     the sampler's unit is one ex2 per (datum, component) pair: MUFU (XU) pipe;
                   the peak is MEASURED here with ``brta_mufu_probe``.  ``issue_frac`` = executed warp
                   instructions / s over the SM's issue rate (the limit the kernel actually runs into)
* ``cpu_baseline`` the reference's arithmetic (oracle.gibbs_oracle.run_reference_order, the bit-exact NumPy
                   restatement of basicrta/gibbs.py:191-217) on all host cores, on a bounded sample

``--impl reference`` times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NCOMP = 15
NITER = 110000
THIN = 100
N_RESIDUES = 400
TS = 0.1
METRIC = 'gibbs_indicator_draws_per_s'
UNIT = 'N*K*iter/s'
C4_LEG_NITER = 11000                 # the c4 leg of the default run: a tenth of the chain (cost per iteration is stationary)


# ---- synthetic workloads ---------------------------------------------------------
def residue_times(r, seed_offset=0, n_scale=1.0):
    """Residue r of C2: seed 1000+r; N = round(10^U(4,5)); 2-4 true components with rates
    log-uniform in [1e-3, 10] /ns separated by >= x5; weights ~ Dirichlet(1) sorted so the
    fastest rate has the largest weight; ceil-quantised to the 0.1 ns grid.  C3 (the same
    residues at five contact cutoffs) shifts the seed by 10 000 per cutoff and scales N."""
    rng = np.random.default_rng(1000 + r + seed_offset)
    n = int(round(n_scale * 10 ** rng.uniform(4, 5)))
    m = int(rng.integers(2, 5))
    while True:
        rates = np.sort(10 ** rng.uniform(-3, 1, size=m))[::-1]
        if np.all(rates[:-1] / rates[1:] >= 5):
            break
    w = np.sort(rng.dirichlet(np.ones(m)))[::-1]
    comp = rng.choice(m, size=n, p=w)
    x = rng.exponential(1.0 / rates[comp])
    ticks = np.maximum(np.ceil(x / TS), 1.0).astype(np.int64)
    return ticks


def c5_residue_times(r):
    """Residue r of the stress configuration C5: seed 5000+r, N = 20 000, five true
    components with rates over four decades; run with ncomp = 30."""
    rng = np.random.default_rng(5000 + r)
    rates = np.array([10, 1, 0.1, 0.01, 0.001])
    comp = rng.choice(5, size=20000, p=[0.6, 0.25, 0.1, 0.04, 0.01])
    x = rng.exponential(1.0 / rates[comp])
    return np.maximum(np.ceil(x / TS), 1.0).astype(np.int64)


def three_exponential(n, seed):
    """C1 / C4: the reference's test mixture (tests/test_functions.py:43-44), quantised to the grid."""
    rng = np.random.default_rng(seed)
    comp = rng.choice(3, size=n, p=[0.90, 0.09, 0.01])
    x = rng.exponential(1.0 / np.array([5, 0.05, 0.001])[comp])
    return np.maximum(np.ceil(x / TS), 1.0).astype(np.int64)


def workload(indices):
    return [residue_times(r) for r in indices]


C3_SCALES = (0.6, 0.8, 1.0, 1.25, 1.5)


def config_workload(name, n_res=None, niter=None):
    """(description, ncomp, niter, list of tick arrays) of a named configuration."""
    if name == 'C1':
        return ('C1 single residue: synthetic 3-exponential residence times N=5000', 15, niter or 10000,
                [three_exponential(5000, 20241109)])
    if name == 'C2':
        n = n_res or N_RESIDUES
        return (f'C2 full-membrane sweep: {n} residues x 1e4-1e5 times', NCOMP, niter or NITER, workload(range(n)))
    if name == 'C3':
        n = n_res or N_RESIDUES
        chains = [residue_times(r, seed_offset=10000 * c, n_scale=f) for c, f in enumerate(C3_SCALES) for r in range(n)]
        return (f'C3 multi-cutoff ensemble: {n} residues x 5 contact cutoffs = {len(chains)} chains in one batch per GPU',
                NCOMP, niter or NITER, chains)
    if name == 'C4':
        return ('C4 giant single residue N=1000000', 15, niter or NITER, [three_exponential(1_000_000, 4)])
    if name == 'C5':
        n = n_res or 100
        return (f'C5 stress: {n} residues N=20000, rates over 4 decades, ncomp=30', 30, niter or NITER,
                [c5_residue_times(r) for r in range(n)])
    raise SystemExit(f'unknown config {name}')


# ---- CPU reference arm ---------------------------------------------------------------------------
def _cpu_chain(args):
    ticks, niter, seed = args
    from oracle import gibbs_oracle as O
    t0 = time.perf_counter()
    O.run_reference_order(ticks * TS, NCOMP, niter, np.random.default_rng(seed), g=THIN)
    return time.perf_counter() - t0


def cpu_reference(n_chains, niter, cores, repeats=1):
    """Time the reference arithmetic on `cores` processes: residues 0..n_chains-1 of the
    workload, truncated to `niter` iterations (per-iteration cost is stationary)."""
    from multiprocessing import get_context
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    os.environ.setdefault('OPENBLAS_NUM_THREADS', '1')
    os.environ.setdefault('MKL_NUM_THREADS', '1')
    chains = workload(range(n_chains))
    units = float(sum(len(c) for c in chains)) * NCOMP * niter
    walls = []
    with get_context('fork').Pool(cores) as pool:
        for rep in range(repeats):
            t0 = time.perf_counter()
            pool.map(_cpu_chain, [(c, niter, 7 + i) for i, c in enumerate(chains)], chunksize=1)
            walls.append(time.perf_counter() - t0)
    return units, walls


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_chains, niter = cores, 300
    units, walls = cpu_reference(n_chains, niter, cores, repeats=args.warmup + args.steps)
    timed = walls[args.warmup:]
    sec = float(np.mean(timed))
    value = units / sec
    sample = (f'EXTRAPOLATED from a bounded sample: residues 0..{n_chains - 1} of the 400-residue workload, niter '
              f'truncated to {niter} (of {NITER}); oracle.gibbs_oracle.run_reference_order = basicrta/gibbs.py:191-217 '
              f'arithmetic, multiprocessing.Pool({cores})')
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'C2 full-membrane sweep: {N_RESIDUES} residues x 1e4-1e5 times, ncomp={NCOMP}, '
                               f'niter={NITER}, thin={THIN} (CPU arm: bounded sample of it, see cpu_baseline.sample)'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'residues_per_hour': n_chains * (niter / NITER) * 3600.0 / sec,
    }))


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap,power.draw')

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.gpu), f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                 '-lms', '200'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for row in self.rows:
            try:
                sm.append(float(row[0]))
                mx = float(row[1])
                for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                     row[2:6]):
                    if val.lower().startswith('active'):
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ---- evidence from the committed ncu captures (profiles/README.md) ---------------------------------
PROFILE_TRAFFIC = os.path.join(ROOT, 'profiles', 'r2z_dram_bytes_bench.csv')
PROFILE_FULL = os.path.join(ROOT, 'profiles', 'r2y_ncu_full_selected_c2.csv')
PROFILE_FULL_NITER = 300             # iterations of the launch captured in PROFILE_FULL (tools/perf.py 400 300)


def profiled_traffic(config, n_res, niter, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the sweep kernel on the default
    workload, from the committed ncu capture; None for any other workload."""
    if (config, n_res, niter, world) != ('C2', N_RESIDUES, NITER, 1) or not os.path.exists(PROFILE_TRAFFIC):
        return None
    total = 0
    with open(PROFILE_TRAFFIC) as f:
        for row in f:
            cells = [c.strip('"') for c in row.strip().split('","')]
            if len(cells) > 3 and cells[-3] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                total += int(cells[-1])
    return float(total) if total else None


def profiled_inst_per_unit(config, n_res):
    """Executed warp instructions per algorithmic unit of the sweep kernel on C2 (ncu --set full capture)."""
    if (config, n_res) != ('C2', N_RESIDUES) or not os.path.exists(PROFILE_FULL):
        return None
    with open(PROFILE_FULL) as f:
        for row in f:
            if row.startswith('smsp__inst_executed.sum,'):
                inst = float(row.strip().split(',')[-1])
                sum_n = float(sum(len(t) for t in workload(range(N_RESIDUES))))
                return inst / (sum_n * NCOMP * PROFILE_FULL_NITER)
    return None


def measure_mufu_peak(torch, lib, device, sm_count):
    """ex2/s the chip sustains (brta_mufu_probe): 8 resident CTAs of 256 threads per SM."""
    import ctypes as C
    blocks, iters = sm_count * 8, 20000
    sink = torch.zeros(blocks * 256, dtype=torch.float32, device=f'cuda:{device}')
    s = torch.cuda.current_stream()
    best = 0.0
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.brta_mufu_probe(C.c_void_p(sink.data_ptr()), blocks, iters, C.c_void_p(s.cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        assert rc == 0
        if rep:
            best = max(best, blocks * 256 * 8.0 * iters / (e0.elapsed_time(e1) * 1e-3))
    return best


# ---- extra legs of the default run -------------------------------------------------------------------
def api_leg(all_ticks, ncomp, niter, n_gpu, where):
    """Wall clock of the public API on this process: Gibbs objects from NumPy times -> dispatch over n_gpu GPUs
    (one host thread per GPU) -> the reference's pickles under `where`.  Everything inside the timed region."""
    from basicrta_b200.gibbs import Gibbs, dispatch
    times = [t * TS for t in all_ticks]
    root = tempfile.mkdtemp(prefix='brta_bench_', dir=where)
    cwd = os.getcwd()
    os.chdir(root)
    try:
        t0 = time.perf_counter()
        gibbs = [Gibbs(t, f'X{r}', r % n_gpu, ncomp=ncomp, niter=niter, cutoff=7.0) for r, t in enumerate(times)]
        t_ctor = time.perf_counter() - t0
        dispatch(gibbs, n_gpu, seed=20241109)
        wall = time.perf_counter() - t0
        n_files = sum(os.path.exists(f'basicrta-7.0/X{r}/gibbs_{niter}.pkl') for r in range(len(times)))
        n_bytes = sum(os.path.getsize(f'basicrta-7.0/X{r}/gibbs_{niter}.pkl') for r in range(len(times)))
        ok = all(g.mcweights is not None and abs(float(g.mcweights[-1].sum()) - 1.0) < 1e-4 for g in gibbs)
    finally:
        os.chdir(cwd)
        shutil.rmtree(root, ignore_errors=True)
    return {'wall_s': wall, 'constructors_s': t_ctor, 'pickles': n_files, 'pickle_bytes': n_bytes, 'ok': bool(ok),
            'dir': where}


def posterior_cloud(r, rows=1000):
    """Retained (log weight, log rate) samples shaped like residue r of C2 after sampling: 2-4 clusters, one per
    true component, `rows` posterior rows each (what Gibbs.cluster feeds the mixture model, gibbs.py:238-252)."""
    rng = np.random.default_rng(77000 + r)
    m = int(rng.integers(2, 5))
    lr = np.sort(rng.uniform(np.log(1e-3), np.log(10.0), m))[::-1] + np.arange(m)[::-1] * 0.5
    w = np.sort(rng.dirichlet(np.ones(m)))[::-1]
    pts = []
    for k in range(m):
        a = np.log(w[k]) + 0.05 / np.sqrt(w[k]) * rng.standard_normal(rows)
        b = lr[k] + 0.04 / np.sqrt(w[k]) * rng.standard_normal(rows) + 0.3 * (a - np.log(w[k]))
        pts.append(np.stack((a, b), axis=1))
    x = np.concatenate(pts)
    return x[rng.permutation(len(x))], m


def gmm_leg(device, n_res=N_RESIDUES, n_init=117, cpu_residues=2):
    """the mixture fits of process_gibbs for every residue of the sweep (n_init = 117,
    gibbs.py:296) as one batch on the GPU, host arrays in, fitted parameters + labels out (wall clock), beside
    scikit-learn -- the reference's own implementation -- on a bounded sample of the same problems."""
    from basicrta_b200 import gmm
    clouds = [posterior_cloud(r) for r in range(n_res)]
    xs, ks = [c[0] for c in clouds], [c[1] for c in clouds]
    gmm.fit_batch(xs[:8], ks[:8], n_init=4, seed=1, device=device)                 # module load, allocator
    best = 1e30
    for rep in range(2):
        t0 = time.perf_counter()
        fits = gmm.fit_batch(xs, ks, n_init=n_init, seed=20241109, device=device, problem_ids=np.arange(n_res))
        labels = gmm.predict_batch(xs, fits, device=device)
        best = min(best, time.perf_counter() - t0)
    out = {'residues': n_res, 'n_init': n_init, 'points': int(sum(len(x) for x in xs)), 'wall_s': best,
           'fits_per_s': n_res * n_init / best, 'failed': int(sum(f.error is not None for f in fits)),
           'not_converged': int(sum(not f.converged for f in fits)),
           'what': 'gmm.fit_batch + predict_batch: NumPy points in, parameters and labels out; one CTA per (residue, restart)'}
    try:
        from sklearn.mixture import GaussianMixture
        t0 = time.perf_counter()
        worst, agree = 0.0, 1.0
        for r in range(cpu_residues):
            sk = GaussianMixture(n_components=ks[r], n_init=n_init, random_state=r).fit(xs[r])
            worst = max(worst, abs(sk.lower_bound_ - fits[r].lower_bound))
            a, b = np.argsort(fits[r].means[:, 1]), np.argsort(sk.means_[:, 1])
            ra, rb = np.empty(ks[r], int), np.empty(ks[r], int)
            ra[a], rb[b] = np.arange(ks[r]), np.arange(ks[r])
            agree = min(agree, float(np.mean(ra[labels[r]] == rb[sk.predict(xs[r])])))
        cpu_s = (time.perf_counter() - t0) / cpu_residues
        out['cpu_baseline'] = {'kind': 'reference', 'cores': 1, 's_per_residue': cpu_s,
                               'sample': f'scikit-learn GaussianMixture(n_init={n_init}) on residues 0..{cpu_residues - 1} '
                                         'of the same problems, one core (the reference fans residues over a process pool)',
                               'extrapolated_wall_s_one_core': cpu_s * n_res,
                               'max_lower_bound_difference': worst, 'min_label_agreement': agree}
    except ImportError:
        pass
    return out


def c4_leg(torch, dist, rank, world, local, niter):
    """The giant single residue on `world` GPUs (one rank per GPU): its times sharded over the ranks, integer
    (n_k, sum tick_k) exchanged inside the persistent kernels over NVLink every iteration.  Returns on rank 0:
    microseconds per iteration (device-timed, max over ranks) and whether the run equals the 1-GPU run bit for bit."""
    from basicrta_b200.engine import ChainInput, ShardedChain, get_engine
    ticks = three_exponential(1_000_000, 4)
    chain = ChainInput(ticks=ticks, ts=TS, chain_id=4)
    dev = f'cuda:{local}'
    out = {'n_data': int(len(ticks)), 'ncomp': 15, 'niter': int(niter), 'n_gpus': world}
    ms_multi, crc_multi, status = None, None, 0
    if world > 1:
        sc = ShardedChain(chain, 15, niter, thin=THIN, seed=1, device=local)
        try:
            best = 1e30
            for rep in range(3):
                torch.cuda.synchronize(local)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                sc.launch(before=e0.record, after=e1.record)
                torch.cuda.synchronize(local)
                best = min(best, e0.elapsed_time(e1))
            res = sc.fetch()
            t = torch.tensor([best], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_multi = float(t[0])
            crc = zlib.crc32(np.ascontiguousarray(res.indicator).tobytes())
            st = torch.tensor([float(res.status), float(crc), float(sc.data_begin), float(sc.data_end)],
                              dtype=torch.float64, device=dev)
            gathered = [torch.zeros_like(st) for _ in range(world)]
            dist.all_gather(gathered, st)
            status = int(max(float(g[0]) for g in gathered))
            crc_multi = [(int(g[1]), int(g[2]), int(g[3])) for g in gathered]
            w_multi, r_multi, order = res.mcweights, res.mcrates, sc.order
        finally:
            sc.close()
    if rank == 0:
        eng = get_engine(local)
        db = eng.prepare([chain], 15, niter, thin=THIN, seed=1)
        best = 1e30
        for rep in range(3):
            eng.reset(db)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.launch(db)
            e1.record()
            torch.cuda.synchronize(local)
            best = min(best, e0.elapsed_time(e1))
        one = eng.fetch(db)[0]
        out['us_per_iter_1gpu'] = best * 1e3 / niter
        out['status'] = int(one.status) | status
        if world > 1:
            out['us_per_iter'] = ms_multi * 1e3 / niter
            canonical = one.indicator[:, order]
            same = np.array_equal(one.mcweights, w_multi) and np.array_equal(one.mcrates, r_multi)
            for crc, a, b in crc_multi:
                same = same and crc == zlib.crc32(np.ascontiguousarray(canonical[:, a:b]).tobytes())
            out['identical_to_1gpu'] = bool(same)
            out['units_per_s'] = len(ticks) * 15.0 * niter / (ms_multi * 1e-3)
            out['exchange'] = ('in-kernel: every GPU stores its 2K integers into every peer\'s mailbox over NVLink '
                               '(CUDA IPC peer memory), no host call per iteration')
        else:
            out['us_per_iter'] = out['us_per_iter_1gpu']
            out['units_per_s'] = len(ticks) * 15.0 * niter / (best * 1e-3)
        del db
    return out


# ---- B200 arm ------------------------------------------------------------------------------------------
def run_b200_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    config = args.config
    default_run = (config == 'C2' and args.residues == N_RESIDUES and args.niter is None)

    # CPU baseline first (rank 0, N = 1 only): fork before CUDA is initialised in this process
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and config == 'C2':
        cores = os.cpu_count() or 1
        n_chains, niter_cpu = cores, 300
        units, walls = cpu_reference(n_chains, niter_cpu, cores)
        cpu = {'value': units / walls[0], 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': f'EXTRAPOLATED from a bounded sample: residues 0..{n_chains - 1} of the workload, niter truncated '
                         f'to {niter_cpu}; oracle.gibbs_oracle.run_reference_order (basicrta/gibbs.py:191-217 arithmetic, '
                         f'fp64 NumPy), multiprocessing.Pool({cores}); {walls[0]:.1f} s wall'}

    import torch
    import torch.distributed as dist
    from basicrta_b200.engine import ChainInput, GibbsEngine, ShardedChain
    from basicrta_b200.plan import shard_chains

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (B200); there is no CPU fallback')
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        cpu_group = dist.new_group(backend='gloo')          # host-side barriers that leave the GPUs alone

    label, ncomp, niter, all_ticks = config_workload(config, None if config != 'C2' and args.residues == N_RESIDUES
                                                     else args.residues, args.niter)
    sizes = np.array([len(t) for t in all_ticks])
    total_units = float(sizes.sum()) * ncomp * niter
    sharded_chain = (config == 'C4' and world > 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f'cuda:{local}')   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    eng = GibbsEngine(local)
    if sharded_chain:
        sc = ShardedChain(ChainInput(ticks=all_ticks[0], ts=TS, chain_id=4), ncomp, niter, thin=THIN, seed=20241109,
                          device=local)
        db = sc.db

        def device_step():
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sc.launch(before=e0.record, after=e1.record)
            return e0, e1
    else:
        mine = shard_chains(sizes, world)[rank]
        chains = [ChainInput(ticks=all_ticks[i], ts=TS, chain_id=int(i)) for i in mine]
        db = eng.prepare(chains, ncomp, niter, thin=THIN, seed=20241109)

        def device_step():
            eng.reset(db)
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.launch(db)
            e1.record()
            return e0, e1

    label_bytes = int(db.tensors['indicator'].numel())
    pinned_e2e = label_bytes <= (24 << 30)                   # page-locking more than that is not a sane host buffer
    host_out = eng.alloc_host_outputs(db) if pinned_e2e else None

    def e2e_step():
        if not sharded_chain:
            eng.reset(db)
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sharded_chain:
            sc.launch(before=lambda: (e0.record(), eng.upload(db)))
            h2d = db.h2d_bytes
        else:
            e0.record()
            h2d = eng.upload(db)
            eng.launch(db)
        d2h = eng.download(db, host_out) if pinned_e2e else eng.download_through_ring(db)
        e1.record()
        return e0, e1, h2d, d2h

    for _ in range(args.warmup):
        device_step()
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    events = [device_step() for _ in range(args.steps)]
    barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b in events)

    # end-to-end leg (host buffers): one warm-up, then `steps` timed
    e2e_step()
    barrier()
    e2e_events = [e2e_step() for _ in range(args.steps)]
    barrier()
    e2e_ms = sum(a.elapsed_time(b) for a, b, _, _ in e2e_events)
    h2d_bytes, d2h_bytes = e2e_events[0][2], e2e_events[0][3]
    status = db.tensors['status'].cpu().numpy()
    if int(np.abs(status).max()) != 0:
        raise SystemExit(f'sampler reported status {int(np.abs(status).max())}')
    w_last = db.tensors['mcweights'][:, -1, :].sum(dim=1).cpu().numpy()
    if not np.allclose(w_last, 1.0, atol=1e-4):
        raise SystemExit('weights do not sum to one: the kernel did not run correctly')

    t = torch.tensor([ms, e2e_ms, float(h2d_bytes), float(d2h_bytes)], dtype=torch.float64, device=f'cuda:{local}')
    per_rank_ms = [ms / args.steps]
    if world > 1:
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        per_rank_ms = [float(g[0]) / args.steps for g in gathered]
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(tmax[0]), float(tmax[1])
        h2d_bytes, d2h_bytes = float(tsum[2]), float(tsum[3])

    n_launch = len(db.segments) if db.segments and len(db.segments) > 2 else 1
    plan_note = (f'{n_launch} back-to-back persistent cooperative launch(es) per step per GPU, grid {db.plan.grid} x 128 '
                 f'threads, {db.plan.n_waves} waves, teams {int(db.plan.team_size.min())}-{int(db.plan.team_size.max())} CTAs'
                 + (f'; kernel build chosen by measurement: {db.kernel_choice}' if db.kernel_choice else ''))
    ex2_share = db.executed_ex2_share
    peak = measure_mufu_peak(torch, eng.lib, local, eng.caps.sm_count) if rank == 0 else None

    # ---- extra legs of the default run: free the main batch first --------------------------------------
    extras = {}
    if sharded_chain:
        sc.close()
    del db, host_out
    torch.cuda.empty_cache()
    if default_run and not args.skip_legs:
        barrier()
        extras['c4'] = c4_leg(torch, dist, rank, world, local, C4_LEG_NITER)
        torch.cuda.empty_cache()
        barrier()
        if rank == 0:
            # the API leg drives ALL `world` GPUs from this process (one host thread per GPU); the other ranks
            # wait on a host-side (gloo) barrier with idle GPUs
            api = {}
            # one small job first (8 residues, 2000 iterations): CUDA contexts on all GPUs of this process, kernel
            # modules, the pinned staging rings -- what a long-lived process pays once, not per job
            api_leg(all_ticks[:8 * world], ncomp, 2000, world, tempfile.gettempdir())
            for name, where in (('tmpfs', '/dev/shm'), ('disk', tempfile.gettempdir())):
                if os.path.isdir(where) and shutil.disk_usage(where).free > 3 * label_bytes * world:
                    api[name] = api_leg(all_ticks, ncomp, niter, world, where)
            extras['e2e_api'] = api
            extras['gmm'] = gmm_leg(local)
        if world > 1:
            dist.barrier(group=cpu_group)

    if rank == 0:
        sec = ms * 1e-3 / args.steps
        value = total_units / sec
        peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        hbm = json.load(open(peaks_file))['hbm_gbs'] if os.path.exists(peaks_file) else 6650.0
        per_gpu = value / world
        rows = (niter + 1) // THIN
        algo_bytes = (float(sizes.sum()) * 2 + float(sizes.sum()) * rows) / world        # ticks in + labels out
        ipu = profiled_inst_per_unit(config, len(all_ticks))
        clk_hz = (clk.get('sm_mhz') or 1965.0) * 1e6
        issue_rate = 4.0 * eng.caps.sm_count * clk_hz
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{label} (sum N = {int(sizes.sum())}), ncomp={ncomp}, niter={niter}, thin={THIN}; '
                                   + ('times of the one chain sharded over' if sharded_chain else 'residues sharded over')
                                   + f' {world} GPU(s), '
                                   + ('per-iteration in-kernel exchange of (n_k, sum tick_k) over NVLink' if sharded_chain
                                      else 'no collective'),
                       'l2': 'flushed between steps (256 MiB write); chain data lives in shared memory',
                       'launch': plan_note},
            'residues_per_hour': len(all_ticks) * 3600.0 / sec,
            'e2e': {'value': total_units / (e2e_ms * 1e-3 / args.steps), 'unit': UNIT,
                    'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes,
                    'ms_per_step': e2e_ms / args.steps,
                    'note': 'engine level: pinned H2D of the packed ticks + sweep + D2H of every result into '
                            + ('one pinned host buffer' if pinned_e2e else 'a ring of pinned staging buffers')
                            + '; e2e_api is the public-API figure'},
            'gpu_launches': args.steps * world * n_launch,
            'per_rank_ms_per_step': [round(x, 3) for x in per_rank_ms],
            'roofline': {'bound': 'mufu', 'achieved': per_gpu / 1e9, 'peak': peak / 1e9, 'unit': 'G ex2/s',
                         'frac': per_gpu / peak, 'frac_executed': per_gpu * ex2_share / peak,
                         'issue_frac': None if ipu is None else ipu * per_gpu / issue_rate,
                         'traffic': profiled_traffic(config, len(all_ticks), niter, world),
                         'traffic_unit': 'bytes per launch (ncu dram read + write, profiles/r2z_dram_bytes_bench.csv); '
                                         f'algorithmic: {algo_bytes:.4g}',
                         'executed_ex2_share': ex2_share,
                         'note': 'per GPU.  frac: algorithmic units (1 ex2 per (datum, component) pair per iteration, SURVEY '
                                 '8d) over the MUFU peak measured in this run by brta_mufu_probe (nominal 148 SM x 16/clk x '
                                 '1.965 GHz = 4654 G/s): how fast the work is retired relative to doing every ex2 at XU '
                                 'peak.  Data with equal ticks share memoised cumulative rows, so only executed_ex2_share of '
                                 'those ex2 are issued (frac_executed = the XU pipe\'s real load).  The limit the kernel '
                                 'runs into is instruction issue and dependent-instruction latency: issue_frac = executed '
                                 'warp instructions per unit (ncu, profiles/r2y_ncu_full_selected_c2.csv) x units/s over '
                                 '4 x SMs x clock.',
                         'hbm': {'achieved': algo_bytes / sec / 1e9, 'peak': hbm, 'unit': 'GB/s',
                                 'frac': algo_bytes / sec / 1e9 / hbm,
                                 'note': 'algorithmic bytes = ticks in (2 B/datum) + labels out (1 B/datum/saved row); '
                                         'HBM is not the bound of this kernel'}},
            'clocks': clk,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if 'e2e_api' in extras and extras['e2e_api']:
            api = extras['e2e_api']
            first = api.get('tmpfs') or next(iter(api.values()))
            line['e2e_api'] = {'value': total_units / first['wall_s'], 'unit': UNIT, 'n_gpus': world,
                               'what': 'basicrta_b200.gibbs.dispatch: NumPy times -> reference pickles on '
                                       + first['dir'] + '; one process, one host thread per GPU; sorting, planning, '
                                       'calibration launches, H2D, sweep, D2H, pickling all inside the wall clock (after one small '
                                       'warm-up job of 8 residues per GPU: contexts, modules, pinned staging rings)',
                               'runs': api}
        if 'c4' in extras:
            line['c4'] = extras['c4']
        if 'gmm' in extras:
            line['gmm'] = extras['gmm']
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', default='C2', choices=['C1', 'C2', 'C3', 'C4', 'C5'])
    ap.add_argument('--residues', type=int, default=N_RESIDUES, help='developer knob; default = the named config')
    ap.add_argument('--niter', type=int, default=None, help='developer knob; default = the named config')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--skip-legs', action='store_true', help='developer knob: no e2e_api / c4 / gmm legs')
    ap.add_argument('--only-gmm', action='store_true', help='developer knob: run the gmm leg alone and print it')
    args = ap.parse_args()
    if args.only_gmm:
        print(json.dumps({'gmm': gmm_leg(0)}))
    elif args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
