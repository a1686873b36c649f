"""Quick free-running sanity run on the GPU (developer tool; prints throughput and posterior)."""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from basicrta_b200.engine import ChainInput, get_engine  # noqa: E402

n_chains = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
niter = int(sys.argv[3]) if len(sys.argv) > 3 else 10000
K = int(sys.argv[4]) if len(sys.argv) > 4 else 15

rng = np.random.default_rng(20241109)
chains = []
for r in range(n_chains):
    comp = rng.choice(3, size=n, p=[0.9, 0.09, 0.01])
    x = rng.exponential(1.0 / np.array([5, 0.05, 0.001])[comp])
    ticks = np.maximum(np.ceil(x / 0.1), 1).astype(np.int64)
    chains.append(ChainInput(ticks=ticks, ts=0.1, chain_id=r))

import torch  # noqa: E402
eng = get_engine(0)
db = eng.prepare(chains, K, niter, thin=100, seed=1)
print('plan: waves', db.plan.n_waves, 'grid', db.plan.grid, 'cap', db.plan.slice_cap_quads,
      'teams', db.plan.team_size.min(), db.plan.team_size.max(), 'eff', round(db.plan.est_efficiency, 3))
for rep in range(3):
    eng.reset(db)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.launch(db)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'rep {rep}: {ms:.2f} ms  {db.units / ms / 1e6:.2f} G units/s')
t0 = time.time()
res = eng.fetch(db)
print('fetch s', round(time.time() - t0, 3))
r0 = res[0]
burn = r0.mcweights.shape[0] // 5
w, rt = r0.mcweights[burn:], r0.mcrates[burn:]
keep = w > 10.0 / n
print('status', [r.status for r in res][:8])
print('mean #comp above cutoff', keep.sum(1).mean())
order = np.argsort(-rt, axis=1)
for j in range(4):
    sel = np.take_along_axis(keep, order, 1)
    rs = np.take_along_axis(rt, order, 1)
    ws = np.take_along_axis(w, order, 1)
# crude summary: weight-weighted log-rate histogram peaks
allr, allw = rt[keep], w[keep]
for lo, hi in [(1, 20), (0.01, 0.3), (0.0002, 0.005)]:
    m = (allr > lo) & (allr < hi)
    if m.any():
        print(f'rates in ({lo},{hi}): mean rate {allr[m].mean():.5f} mean weight {allw[m].mean():.4f} n {m.sum()}')
print('weights sum', r0.mcweights[-1].sum(), 'last rates', np.sort(r0.mcrates[-1])[::-1][:5])
