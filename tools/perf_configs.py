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
    """The synthetic inputs of SURVEY.md 8(d)."""
    c1 = [0.90, 0.09, 0.01], [5, 0.05, 0.001]
    yield 'C1 single residue N=5e3 K=15', 15, [synth(np.random.default_rng(20241109), 5000, c1[1], c1[0])]
    yield 'C2 400 residues K=15', 15, bench.workload(range(bench.N_RESIDUES))
    c3 = [bench.residue_times(r, seed_offset=10000 * c, n_scale=f)
          for c, f in enumerate((0.6, 0.8, 1.0, 1.25, 1.5)) for r in range(bench.N_RESIDUES)]
    yield 'C3 2000 chains (5 cutoffs) K=15', 15, c3
    yield 'C4 giant residue N=1e6 K=15 (1 GPU)', 15, [synth(np.random.default_rng(4), 1_000_000, c1[1], c1[0])]
    c5 = [synth(np.random.default_rng(5000 + r), 20000, [10, 1, 0.1, 0.01, 0.001], [0.6, 0.25, 0.1, 0.04, 0.01])
          for r in range(100)]
    yield 'C5 100 residues N=2e4 K=30, rates over 4 decades', 30, c5


eng = GibbsEngine(0)
for name, K, ticks in configs():
    chains = [ChainInput(ticks=t, ts=0.1, chain_id=i) for i, t in enumerate(ticks)]
    db = eng.prepare(chains, K, niter, thin=100, seed=1, calibrate=True)
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
