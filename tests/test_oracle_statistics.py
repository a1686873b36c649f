"""The two CPU oracles agree statistically ("O2 is validated against O1 by the same 2 % / KS
criteria the kernel must meet").

O1 = ``run_reference_order``: the reference's arithmetic and NumPy generator calls, pinned bit for bit to
outputs of the unmodified reference (tests/test_oracle_golden.py).  O2 = ``run_teacher_forced``: the device
arithmetic (float32 log2-space logits, max subtraction, ``soft_exp2``, ONE Philox uniform per datum and an
inverse CDF in component order, integer tick sums) with the posterior drawn by NumPy -- free running, it is
a complete sampler.  O2 replaces NumPy's conditional-binomial ``multinomial`` (basicrta/gibbs.py:200) by a
single-uniform inverse CDF: same distribution, different random numbers.  This test runs both on the same
data and compares label-invariant posterior functionals: means within 2 %, two-sample KS p > 0.01 on
thinned samples.  Everything is seeded, so the outcome is deterministic.
"""
from multiprocessing import get_context

import numpy as np
import pytest
from scipy import stats

from oracle import gibbs_oracle as O

N, K, NITER, THIN, BURN_ROWS = 1500, 6, 16000, 100, 20
WEIGHTS, RATES = [0.80, 0.15, 0.05], [4.0, 0.2, 0.01]
REL_TOL = 0.02


def _chain(job):
    kind, seed = job
    times = O.synth_times(N, WEIGHTS, RATES, seed=99)
    if kind == 'O1':
        out = O.run_reference_order(times, K, NITER, np.random.default_rng(seed), g=THIN)
    else:
        ts = O.time_step(times)
        out = O.run_teacher_forced(O.to_ticks(times, ts), ts, K, NITER, seed=seed, chain_id=seed,
                                   rng=np.random.default_rng(1000 + seed), g=THIN)
    return O.posterior_functionals(out['mcweights'][BURN_ROWS:], out['mcrates'][BURN_ROWS:], times)


@pytest.fixture(scope='module')
def samples():
    jobs = [('O1', 1), ('O1', 2), ('O2', 3), ('O2', 4)]
    with get_context('fork').Pool(len(jobs)) as pool:
        res = pool.map(_chain, jobs, chunksize=1)
    return np.stack(res[:2]), np.stack(res[2:])               # [chain, sample, functional]


def test_means_within_2_percent(samples):
    o1, o2 = samples
    for i, name in enumerate(O.FUNCTIONAL_NAMES):
        if not O.well_determined(N)[i]:
            continue
        a, b = o1[..., i].mean(), o2[..., i].mean()
        assert abs(b / a - 1) < REL_TOL, (name, a, b)


def test_ks_on_thinned_samples(samples):
    o1, o2 = samples
    for i, name in enumerate(O.FUNCTIONAL_NAMES[:-1]):        # the last one is an integer count (ties)
        if not O.well_determined(N)[i]:
            continue
        p = stats.ks_2samp(o1[:, ::4, i].ravel(), o2[:, ::4, i].ravel()).pvalue
        assert p > 0.01, (name, p)


def test_component_count_distribution(samples):
    o1, o2 = samples
    h1 = np.bincount(o1[..., -1].astype(int).ravel(), minlength=K + 1) / o1[..., -1].size
    h2 = np.bincount(o2[..., -1].astype(int).ravel(), minlength=K + 1) / o2[..., -1].size
    assert h1.argmax() == h2.argmax()
    assert np.abs(h1 - h2).max() < 0.1, (h1, h2)
