"""Developer tool: per-phase cycle breakdown of the sweep kernel (needs the -DBRTA_PHASE_TIMING build).

    BRTA_LIB=basicrta_b200/libbrta_gibbs_dbg.so python tools/phase_timing.py 40 30000 1000
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, '.')
import torch  # noqa: E402
from basicrta_b200 import _cabi  # noqa: E402
from basicrta_b200.engine import ChainInput, get_engine  # noqa: E402

n_chains, n, niter = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
K = int(sys.argv[4]) if len(sys.argv) > 4 else 15
rng = np.random.default_rng(1)
chains = []
if n == 0:                                                # the bench workload (C2): first n_chains residues
    import bench
    chains = [ChainInput(ticks=t, ts=0.1, chain_id=i) for i, t in enumerate(bench.workload(range(n_chains)))]
for r in range(n_chains if n else 0):
    comp = rng.choice(3, size=n, p=[0.9, 0.09, 0.01])
    x = rng.exponential(1.0 / np.array([5, 0.05, 0.001])[comp])
    chains.append(ChainInput(ticks=np.maximum(np.ceil(x / 0.1), 1).astype(np.int64), ts=0.1, chain_id=r))
from basicrta_b200.engine import GibbsEngine
cps = os.environ.get('BRTA_CTAS_PER_SM')
eng = GibbsEngine(0, ctas_per_sm=int(cps) if cps else None)
nw = os.environ.get('BRTA_WAVES')
db = eng.prepare(chains, K, niter, thin=100, seed=1, n_waves=int(nw) if nw else None, calibrate=bool(int(os.environ.get('BRTA_CALIBRATE', '0'))))
print('waves', db.plan.n_waves, 'est eff', round(db.plan.est_efficiency, 3))
buf = torch.zeros(db.plan.grid * 8, dtype=torch.int64, device='cuda')
lib = _cabi.load()
assert lib.brta_debug_set_phase_buffer(C.c_void_p(buf.data_ptr())) == 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.launch(db); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
ph = buf.cpu().numpy().reshape(-1, 8).astype(np.float64)
ph = ph[ph.sum(1) > 0]
names = ['table build', 'wait C', 'sweep', 'wait A', 'serial+wait B', '(lead) post', '(lead) gather', '(lead) posterior']
tot = ph[:, :5].sum(1).mean()          # phases 5-7 are the lead warp's view of phase 4
print(f'{ms:.2f} ms, {db.units / ms / 1e6:.1f} G units/s, grid {db.plan.grid}, teams {db.plan.team_size.min()}-{db.plan.team_size.max()}, '
      f'slice {db.plan.slice_cap_quads} quads; cycles/iter {tot / niter:.0f}')
for i, nme in enumerate(names):
    print(f'  {nme:14s} {ph[:, i].mean() / niter:9.0f} cyc/iter  {100 * ph[:, i].mean() / tot:5.1f} %   (min {ph[:, i].min() / niter:.0f} max {ph[:, i].max() / niter:.0f})')

# regression: sweep cycles per iteration of a CTA ~ a * served quads + b * recomputed quads + c * tasks
try:
    from basicrta_b200.memo import ChainCost
    srt = [np.sort(np.asarray(c.ticks)) for c in chains]
    costs = [ChainCost(t, K) for t in srt]
    A = np.zeros((db.plan.grid, 3))
    for bidx in range(db.plan.grid):
        for t in db.plan.tasks_of_cta(bidx):
            s0, e0 = int(t['quad_begin']), int(t['quad_begin'] + t['quad_count'])
            k = costs[int(t['chain'])].served(s0, e0)
            A[bidx] += (k, e0 - s0 - k, 1)
    full = buf.cpu().numpy().reshape(-1, 8).astype(np.float64)
    y = full[:, 2] / niter
    sel = A[:, 2] > 0
    coef, *_ = np.linalg.lstsq(A[sel], y[sel], rcond=None)
    print(f'sweep cycles/iter ~ {coef[0]:.2f} * served + {coef[1]:.2f} * recomputed + {coef[2]:.0f} * tasks  '
          f'(ratio served/recomputed = {coef[0] / coef[1]:.3f}); served share {A[:, 0].sum() / A[:, :2].sum():.3f}')
except Exception as e:
    print('regression failed:', e)

# within-team spread of the sweep phase (single-wave plans: one task per CTA)
if db.plan.n_waves == 1:
    allph = buf.cpu().numpy().reshape(-1, 8).astype(np.float64) / niter
    spread, wait = [], []
    for r in range(len(chains)):
        ctas = [b for b in range(db.plan.grid) if len(db.plan.tasks_of_cta(b)) and db.plan.tasks_of_cta(b)[0]['chain'] == r]
        sw = allph[ctas, 2]
        spread.append((sw.max() - sw.min()) / sw.mean())
        wait.append((sw.max() - sw.mean()))
    print(f'within-team sweep spread (max-min)/mean: median {np.median(spread):.2f}, max {np.max(spread):.2f}; '
          f'mean wait for the slowest member {np.mean(wait):.0f} cyc/iter')
    tm = np.array([allph[[b for b in range(db.plan.grid) if len(db.plan.tasks_of_cta(b)) and db.plan.tasks_of_cta(b)[0]['chain'] == r], 2].max() for r in range(len(chains))])
    print(f'slowest-member sweep per team: min {tm.min():.0f} mean {tm.mean():.0f} max {tm.max():.0f} cyc/iter')
