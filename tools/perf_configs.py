"""Developer tool: device-timed throughput of the sampler on every configuration BASELINE.json names
(short chains -- the per-iteration cost is stationary).  One GPU; C4's multi-GPU form is tools/perf_c4.py.

    python tools/perf_configs.py [NITER]
"""
import sys

import numpy as np

sys.path.insert(0, '.')
import torch  # noqa: E402
import bench  # noqa: E402
from basicrta_b200.engine import ChainInput, GibbsEngine  # noqa: E402

niter = int(sys.argv[1]) if len(sys.argv) > 1 else 300


def synth(rng, n, rates, weights, ts=0.1):
    comp = rng.choice(len(rates), size=n, p=weights)
    x = rng.exponential(1.0 / np.asarray(rates)[comp])
    return np.maximum(np.ceil(x / ts), 1).astype(np.int64)


def configs():
    rng = np.random.default_rng(7)
    # C1: one residue, three exponentials, N = 5 000
    yield 'C1 single residue N=5e3 K=15', 15, [synth(rng, 5000, [5, 0.05, 0.001], [0.9, 0.09, 0.01])]
    # C2: the bench workload
    c2 = bench.workload(range(bench.N_RESIDUES))
    yield 'C2 400 residues K=15', 15, c2
    # C3: the same residues at five contact cutoffs -- a larger cutoff keeps more (and longer) contacts
    c3 = []
    for f in (0.6, 0.8, 1.0, 1.25, 1.5):
        for t in c2:
            keep = max(64, int(len(t) * min(f, 1.0)))
            c3.append(np.maximum(1, np.rint(t[:keep] * max(f, 1.0)).astype(np.int64)))
    yield 'C3 2000 chains (5 cutoffs) K=15', 15, c3
    # C4 on one GPU: one giant chain
    yield 'C4 giant residue N=1e6 K=15 (1 GPU)', 15, [synth(rng, 1_000_000, [5, 0.5, 0.05, 0.005], [0.6, 0.25, 0.1, 0.05])]
    # C5: K = 30, true rates over four decades, 100 residues
    c5 = []
    for r in range(100):
        n = int(round(10 ** rng.uniform(4, 5)))
        c5.append(synth(rng, n, [10, 1, 0.1, 0.01, 0.001], [0.5, 0.25, 0.15, 0.07, 0.03]))
    yield 'C5 100 residues K=30, rates over 4 decades', 30, c5


eng = GibbsEngine(0)
for name, K, ticks in configs():
    chains = [ChainInput(ticks=t, ts=0.1, chain_id=i) for i, t in enumerate(ticks)]
    db = eng.prepare(chains, K, niter, thin=100, seed=1)
    best = 1e30
    for rep in range(3):
        eng.reset(db)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.launch(db); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    st = int(db.tensors['status'].max().item())
    p = db.plan
    n_tot = int(sum(len(t) for t in ticks))
    print(f'{name:46s} sum N {n_tot:9d}  {best / niter * 1e3:8.1f} us/iter  {db.units / best / 1e6:8.1f} G units/s  '
          f'waves {p.n_waves}, teams {p.team_size.min()}-{p.team_size.max()}, ex2 share {db.executed_ex2_share:.3f}, '
          f'status {st}', flush=True)
    del db
    torch.cuda.empty_cache()
