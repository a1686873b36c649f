"""Developer tool: device-timed throughput of one launch on the bench workload.

    python tools/perf.py N_CHAINS NITER               # first N_CHAINS residues of bench.py's C2 workload
    python tools/perf.py N_CHAINS NITER N_DATA        # N_CHAINS synthetic chains of N_DATA data (team tests)
"""
import os
import sys

import numpy as np

sys.path.insert(0, '.')
import torch  # noqa: E402
import bench  # noqa: E402
from basicrta_b200 import _cabi  # noqa: E402
from basicrta_b200.engine import ChainInput, GibbsEngine  # noqa: E402

n_chains, niter = int(sys.argv[1]), int(sys.argv[2])
nd = [int(a) for a in sys.argv[3:] if a.isdigit()]
print('NO TABLE' if 'notable' in sys.argv else 'table', end=' ')
if nd:
    rng = np.random.default_rng(1)
    ticks = []
    for r in range(n_chains):
        comp = rng.choice(3, size=nd[0], p=[0.9, 0.09, 0.01])
        x = rng.exponential(1.0 / np.array([5, 0.05, 0.001])[comp])
        ticks.append(np.maximum(np.ceil(x / 0.1), 1).astype(np.int64))
elif any(a.startswith('shard=') for a in sys.argv):      # shard=G:r -> what rank r of G GPUs gets from bench.py
    from basicrta_b200.plan import shard_chains
    G, r = (int(x) for x in [a for a in sys.argv if a.startswith('shard=')][0][6:].split(':'))
    full = bench.workload(range(bench.N_RESIDUES))
    ticks = [full[i] for i in shard_chains(np.array([len(t) for t in full]), G)[r]]
    n_chains = len(ticks)
else:
    ticks = bench.workload(range(n_chains))
chains = [ChainInput(ticks=t, ts=0.1, chain_id=i) for i, t in enumerate(ticks)]
cs = os.environ.get('BRTA_COST_SERVED')
if cs:
    from basicrta_b200 import memo
    memo.COST_SERVED = float(cs)
cps = os.environ.get('BRTA_CTAS_PER_SM')
nw = os.environ.get('BRTA_WAVES')
ov = os.environ.get('BRTA_OVERHEAD')
eng = GibbsEngine(0, ctas_per_sm=int(cps) if cps else None, overhead_quads=float(ov) if ov else None)
seg = os.environ.get('BRTA_SEGMENTS')
db = eng.prepare(chains, 15, niter, thin=100, seed=1, n_waves=int(nw) if nw else None, calibrate=bool(int(os.environ.get('BRTA_CALIBRATE', '0'))),
                 segments=tuple(float(x) for x in seg.split(',')) if seg else None,
                 flags=(_cabi.FLAG_NO_TABLE if 'notable' in sys.argv else 0) | int(os.environ.get('BRTA_FLAGS', '0')))
best = 1e30
for rep in range(3):
    eng.reset(db)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.launch(db); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
st = db.tensors['status'].cpu().numpy()
p = db.plan
print(f'chains {n_chains} niter {niter}: {best:8.2f} ms  {db.units / best / 1e6:8.1f} G units/s  '
      f'grid {p.grid}, waves {p.n_waves}, teams {p.team_size.min()}-{p.team_size.max()}, '
      f'slice {p.slice_cap_quads} quads, plan eff {p.est_efficiency:.3f}, status max {int(st.max())}, choice {db.kernel_choice}')
