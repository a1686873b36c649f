"""Developer tool: the giant single residue (C4: N = 1e6, K = 15) on 1..G GPUs of one box.

    python tools/perf_c4.py NITER [N]
"""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
import torch  # noqa: E402
from basicrta_b200.engine import ChainInput, get_engine, run_sharded  # noqa: E402

niter = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
rng = np.random.default_rng(4)
comp = rng.choice(3, size=n, p=[0.9, 0.09, 0.01])
ticks = np.maximum(np.ceil(rng.exponential(1.0 / np.array([5, 0.05, 0.001])[comp]) / 0.1), 1).astype(np.int64)
chain = ChainInput(ticks=ticks, ts=0.1, chain_id=4)
ref = None
for G in (1, 2, 4, 8):
    if G > torch.cuda.device_count():
        break
    best = 1e30
    for rep in range(2):
        for d in range(G):
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        res = get_engine(0).run([chain], 15, niter, thin=100, seed=1)[0] if G == 1 else \
            run_sharded(chain, 15, niter, devices=list(range(G)), thin=100, seed=1)
        best = min(best, time.perf_counter() - t0)      # includes H2D/D2H of this short run
    if ref is None:
        ref = res
    same = np.array_equal(ref.mcrates, res.mcrates) and np.array_equal(ref.indicator, res.indicator)
    print(f'G={G}: {best * 1e3:8.1f} ms wall for {niter} iterations ({best / niter * 1e6:6.2f} us/iter incl. copies), '
          f'{n * 15.0 * niter / best / 1e9:8.1f} G units/s, identical to G=1: {same}, status {res.status}')
