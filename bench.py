#!/usr/bin/env python
"""Benchmark of the Gibbs-sampler hot path on BASELINE.json's headline configuration.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload ("C2"): the full-membrane sweep -- 400 residues x 1e4-1e5 residence
times each, ncomp = 15, niter = 110 000, synthetic multi-exponential times on the 0.1 ns
grid, seeds 1000 + r.  One *step* = the whole sweep: every residue's 110 000 Gibbs
iterations.  With N > 1 (torchrun, one rank per GPU) the 400 residues are sharded over
the ranks (strong scaling, no collective on the data path: chains are independent).

Metric: Gibbs indicator draws / s = sum_r N_r * K * niter / seconds (whole job).

* ``value``        device-timed (CUDA events around each step's launch, inputs resident in HBM)
* ``e2e``          the same through the host-buffer path: H2D of the tick arrays from pinned
                   memory + kernel + D2H of mcweights / mcrates / indicator into pinned memory
* ``roofline``     the sampler is bound by the MUFU (XU) pipe, one ex2 per (datum, component)
                   pair; the peak is MEASURED here with ``brta_mufu_probe``
* ``cpu_baseline`` the reference's arithmetic (oracle.gibbs_oracle.run_reference_order, the
                   bit-exact NumPy restatement of basicrta/gibbs.py:191-217) on all host cores,
                   on a bounded sample of the same workload

``--impl reference`` times that CPU path alone (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

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


# ---- synthetic workload ---------------------------------------
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


def workload(indices):
    return [residue_times(r) for r in indices]


# ---- CPU reference arm ---------------------------------------------------------------------
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
    sample = (f'residues 0..{n_chains - 1} of the 400-residue workload, niter truncated to {niter} '
              f'(of {NITER}); oracle.gibbs_oracle.run_reference_order = basicrta/gibbs.py:191-217 arithmetic, '
              f'multiprocessing.Pool({cores})')
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


# ---- clocks --------------------------------------------------------------------------------
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


# ---- B200 arm ----------------------------------------------------------------------------------
def profiled_traffic(n_res, niter, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the sweep kernel on the default
    workload, from the committed ncu capture (profiles/README.md); None for any other workload."""
    path = os.path.join(ROOT, 'profiles', 'r1e_dram_bytes_bench.csv')
    if (n_res, niter, world) != (N_RESIDUES, NITER, 1) or not os.path.exists(path):
        return None
    total = 0
    with open(path) as f:
        for row in f:
            cells = [c.strip('"') for c in row.strip().split('","')]
            if len(cells) > 3 and cells[-3] in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
                total += int(cells[-1])
    return float(total) if total else None


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


def run_b200_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    # CPU baseline first (rank 0, N = 1 only): fork before CUDA is initialised in this process
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_chains, niter = cores, 300
        units, walls = cpu_reference(n_chains, niter, cores)
        cpu = {'value': units / walls[0], 'unit': UNIT, 'cores': cores, 'kind': 'port',
               'sample': f'residues 0..{n_chains - 1} of the workload, niter truncated to {niter}; '
                         f'oracle.gibbs_oracle.run_reference_order (basicrta/gibbs.py:191-217 arithmetic, fp64 NumPy), '
                         f'multiprocessing.Pool({cores}); {walls[0]:.1f} s wall'}

    import torch
    import torch.distributed as dist
    from basicrta_b200.engine import ChainInput, GibbsEngine
    from basicrta_b200.plan import shard_chains

    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (B200); there is no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))

    n_res = args.residues
    niter = args.niter
    all_ticks = workload(range(n_res))
    sizes = np.array([len(t) for t in all_ticks])
    mine = shard_chains(sizes, world)[rank]
    chains = [ChainInput(ticks=all_ticks[i], ts=TS, chain_id=int(i)) for i in mine]
    total_units = float(sizes.sum()) * NCOMP * niter

    eng = GibbsEngine(local)
    db = eng.prepare(chains, NCOMP, niter, thin=THIN, seed=20241109)
    host_out = eng.alloc_host_outputs(db)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f'cuda:{local}')   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        eng.reset(db)
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.launch(db)
        e1.record()
        return e0, e1

    def e2e_step():
        eng.reset(db)
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h2d = eng.upload(db)
        eng.launch(db)
        d2h = eng.download(db, host_out)
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
    status = host_out['status'].numpy()
    if int(np.abs(status).max()) != 0:
        raise SystemExit('sampler reported a non-finite likelihood')
    w_last = host_out['mcweights'].numpy()[:, -1, :].sum(axis=1)
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

    if rank == 0:
        sec = ms * 1e-3 / args.steps
        value = total_units / sec
        peak = measure_mufu_peak(torch, eng.lib, local, eng.caps.sm_count)
        peaks_file = os.path.join(ROOT, 'MEASURED_PEAKS.json')
        hbm = json.load(open(peaks_file))['hbm_gbs'] if os.path.exists(peaks_file) else 6650.0
        per_gpu = value / world
        algo_bytes = (float(sizes.sum()) * 2 + float(sizes.sum()) * ((niter + 1) // THIN)) / world  # ticks in + labels out
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'C2 full-membrane sweep: {n_res} residues x 1e4-1e5 times (sum N = {int(sizes.sum())}), '
                                   f'ncomp={NCOMP}, niter={niter}, thin={THIN}; residues sharded over {world} GPU(s), '
                                   f'no collective',
                       'l2': 'flushed between steps (256 MiB write); chain data lives in shared memory',
                       'launch': f'{len(db.segments) if db.segments and len(db.segments) > 2 else 1} back-to-back persistent cooperative launch(es) per step per GPU '
                                 f'(slices re-cut in between from measured cycles), grid {db.plan.grid} x 128 threads, '
                                 f'{db.plan.n_waves} waves, teams {int(db.plan.team_size.min())}-{int(db.plan.team_size.max())} CTAs'},
            'residues_per_hour': n_res * 3600.0 / sec,
            'e2e': {'value': total_units / (e2e_ms * 1e-3 / args.steps), 'unit': UNIT,
                    'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': d2h_bytes,
                    'ms_per_step': e2e_ms / args.steps},
            'gpu_launches': args.steps * world * (len(db.segments) if db.segments and len(db.segments) > 2 else 1),
            'per_rank_ms_per_step': [round(x, 3) for x in per_rank_ms],
            'roofline': {'bound': 'mufu', 'achieved': per_gpu / 1e9, 'peak': peak / 1e9, 'unit': 'G ex2/s',
                         'frac': per_gpu / peak, 'traffic': profiled_traffic(n_res, niter, world),
                         'traffic_unit': 'bytes per launch (ncu dram read + write, profiles/r1e_dram_bytes_bench.csv); '
                                         f'algorithmic: {algo_bytes:.4g}',
                         'executed_ex2_share': db.executed_ex2_share,
                         'note': 'per GPU; algorithmic unit = 1 ex2 per (datum, component) pair per iteration (SURVEY 8d); '
                                 'peak = brta_mufu_probe measured in this run (nominal 148 SM x 16/clk x 1.965 GHz = 4654 G/s). '
                                 'Data with equal ticks share memoised cumulative rows, so only executed_ex2_share of the '
                                 'algorithmic ex2 are issued to the XU pipe (rank 0 figure); frac counts algorithmic units.',
                         'hbm': {'achieved': algo_bytes / sec / 1e9, 'peak': hbm, 'unit': 'GB/s',
                                 'frac': algo_bytes / sec / 1e9 / hbm,
                                 'note': 'algorithmic bytes = ticks in (2 B/datum) + labels out (1 B/datum/saved row); '
                                         'HBM is not the bound of this kernel'}},
            'clocks': clk,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=2)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--residues', type=int, default=N_RESIDUES, help='developer knob; default = the named config')
    ap.add_argument('--niter', type=int, default=NITER, help='developer knob; default = the named config')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == '__main__':
    main()
