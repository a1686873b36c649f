"""Developer tool: where the wall clock of the public API goes (C2 on one GPU, pickles to /dev/shm).

    python tools/api_breakdown.py [N_RESIDUES] [NITER]
"""
import os
import shutil
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, '.')
import bench  # noqa: E402
from basicrta_b200 import engine as E  # noqa: E402
from basicrta_b200 import gibbs as G  # noqa: E402

n_res = int(sys.argv[1]) if len(sys.argv) > 1 else 400
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 110000
ticks = bench.workload(range(n_res))
times = [t * bench.TS for t in ticks]
marks = []


def timed(obj, name):
    fn = getattr(obj, name)

    def wrap(*a, **k):
        t0 = time.perf_counter()
        out = fn(*a, **k)
        marks.append((name, time.perf_counter() - t0))
        return out
    setattr(obj, name, wrap)


for name in ('prepare', 'launch', 'stream_results', '_calibrate', '_choose_kernel', '_plan'):
    timed(E.GibbsEngine, name)
for name in ('_canonical_order',):
    timed(E, name)
timed(G.Gibbs, '_prepare')
timed(G.Gibbs, '_chain_input')
timed(G.Gibbs, 'save')

for rep in range(2):
    marks.clear()
    root = tempfile.mkdtemp(prefix='brta_api_', dir='/dev/shm')
    cwd = os.getcwd()
    os.chdir(root)
    t0 = time.perf_counter()
    gl = [G.Gibbs(t, f'X{r}', 0, ncomp=15, niter=niter, cutoff=7.0) for r, t in enumerate(times)]
    t1 = time.perf_counter()
    G.dispatch(gl, 1, seed=1)
    t2 = time.perf_counter()
    os.chdir(cwd)
    shutil.rmtree(root)
    print(f'rep {rep}: constructors {t1 - t0:.2f} s, dispatch {t2 - t1:.2f} s')
    agg = {}
    for k, v in marks:
        agg.setdefault(k, [0, 0.0])
        agg[k][0] += 1
        agg[k][1] += v
    for k, (n, v) in agg.items():
        print(f'   {k:18s} x{n:4d}  {v:7.3f} s (summed over threads for save)')
