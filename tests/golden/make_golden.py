"""Generate golden vectors by running the UNMODIFIED reference sampler.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # everything (several minutes on 8 cores)
    python tests/golden/make_golden.py small      # bit-exact replay vectors only

Outputs (committed):

``ref_replay.npz``
    Seeded runs of the reference ``basicrta.gibbs.Gibbs.run()`` (gibbs.py:176-219) with
    the module generator replaced by ``default_rng(seed)`` after import (the loop looks
    ``rng`` up at call time; the reference itself never seeds, gibbs.py:17).  For each
    case: ``times``, the seed and the full ``mcweights`` / ``mcrates`` / ``indicator``
    plus ``ts``, ``t``, ``s``, ``whypers``, ``rhypers`` and the pickle's attribute names.
    ``oracle.gibbs_oracle.run_reference_order`` must reproduce these bit for bit.

``ref_c1_posterior.npz``
    Config 1 of BASELINE.json (N=5000 three-exponential data, K=15) run with the
    reference for niter=110000 under 4 seeds: the thinned ``mcweights`` / ``mcrates``
    samples (float32).  This is the statistical truth the CUDA sampler's free-running
    chains are compared with (2 % on posterior means, KS p > 0.01).
"""
import os
import sys
import tempfile
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refstubs import import_reference  # noqa: E402
from oracle import gibbs_oracle as O    # noqa: E402

C1 = dict(n=5000, weights=[0.90, 0.09, 0.01], rates=[5, 0.05, 0.001], seed=20241109)

REPLAY_CASES = [
    # name, n, true weights, true rates, data seed, ncomp, niter, g, rng seed
    ('k4_n300', 300, [0.9, 0.09, 0.01], [5, 0.05, 0.001], 1, 4, 40, 5, 7),
    ('k15_n500', 500, [0.9, 0.09, 0.01], [5, 0.05, 0.001], 2, 15, 30, 10, 11),
    ('k2_n64', 64, [0.5, 0.5], [2.0, 0.1], 3, 2, 25, 1, 13),
    ('k7_n1000_g100', 1000, [0.6, 0.3, 0.1], [3.0, 0.3, 0.02], 4, 7, 200, 100, 17),
]


def _run_reference(times, ncomp, niter, g, rng_seed, residue='X1'):
    gibbs, _ = import_reference()
    gibbs.rng = np.random.default_rng(rng_seed)
    gibbs.tqdm = lambda it, **kw: it                      # silence the progress bar only
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            G = gibbs.Gibbs(times, residue, 0, ncomp=ncomp, niter=niter, cutoff=7.0)
            G.g = g
            G.run()
            assert os.path.exists(f'basicrta-7.0/{residue}/gibbs_{niter}.pkl')
        finally:
            os.chdir(cwd)
    return G


def make_replay():
    out = {}
    for name, n, w, r, dseed, ncomp, niter, g, rseed in REPLAY_CASES:
        times = O.synth_times(n, w, r, dseed)
        G = _run_reference(times, ncomp, niter, g, rseed)
        out[f'{name}/times'] = times
        out[f'{name}/meta'] = np.array([ncomp, niter, g, rseed], dtype=np.int64)
        for attr in ('mcweights', 'mcrates', 'indicator', 't', 's', 'whypers', 'rhypers'):
            out[f'{name}/{attr}'] = getattr(G, attr)
        out[f'{name}/ts'] = np.float64(G.ts)
        out[f'{name}/attrs'] = np.array(sorted(G.__dict__.keys()))
    np.savez_compressed(os.path.join(HERE, 'ref_replay.npz'), **out)
    print('wrote ref_replay.npz', len(out), 'arrays')


def _c1_worker(rng_seed):
    times = O.synth_times(C1['n'], C1['weights'], C1['rates'], C1['seed'])
    G = _run_reference(times, 15, 110000, 100, rng_seed, residue=f'X{rng_seed}')
    return G.mcweights.astype(np.float32), G.mcrates.astype(np.float32)


def make_c1():
    seeds = [101, 202, 303, 404]
    with Pool(len(seeds)) as p:
        res = p.map(_c1_worker, seeds)
    np.savez_compressed(os.path.join(HERE, 'ref_c1_posterior.npz'),
                        seeds=np.array(seeds),
                        mcweights=np.stack([r[0] for r in res]),
                        mcrates=np.stack([r[1] for r in res]))
    print('wrote ref_c1_posterior.npz')


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in ('all', 'small'):
        make_replay()
    if what in ('all', 'c1'):
        make_c1()
