"""Statistical parity of the free-running CUDA sampler (FAST mode, own Philox stream) with
the REFERENCE sampler: config C1 of BASELINE.json (N = 5000 three-exponential data, K = 15),
reference posterior samples from tests/golden/ref_c1_posterior.npz (4 seeds x 110 000
iterations of the unmodified basicrta.gibbs.Gibbs.run).

North-star criterion: posterior means within 2 % relative, two-sample KS p > 0.01 on
thinned samples.  Compared functionals are invariant under label switching:
the mixture survival S(t) = sum_k w_k exp(-r_k t) on a time grid, the rate / weight of the
heaviest component, and the slowest rate among components with weight > 10/N (what
estimate_tau is built on, gibbs.py:284-296, 691-715).
"""
import os

import numpy as np
import pytest
from scipy import stats

from basicrta_b200.engine import ChainInput
from oracle import gibbs_oracle as O

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
N, K, NITER, THIN, BURN_ROWS = 5000, 15, 110000, 100, 100
REL_TOL = 0.02


def functionals(w, r):
    out = {}
    for t in (0.1, 1.0, 10.0, 100.0, 1000.0):
        out[f'S({t:g})'] = (w * np.exp(-r * t)).sum(axis=1)
    idx = np.arange(len(w))
    out['rate of heaviest'] = r[idx, w.argmax(axis=1)]
    out['weight of heaviest'] = w.max(axis=1)
    out['slowest significant rate'] = np.where(w > 10.0 / N, r, np.inf).min(axis=1)
    return out


@pytest.fixture(scope='module')
def chains(engine):
    gold = np.load(os.path.join(HERE, 'golden', 'ref_c1_posterior.npz'))
    ref_w = gold['mcweights'].astype(np.float64)[:, BURN_ROWS:]
    ref_r = gold['mcrates'].astype(np.float64)[:, BURN_ROWS:]
    times = O.synth_times(N, [0.90, 0.09, 0.01], [5, 0.05, 0.001], seed=20241109)
    ts = O.time_step(times)
    ticks = O.to_ticks(times, ts)
    n_chains = 16                                             # independent chains: distinct Philox chain ids
    res = engine.run([ChainInput(ticks=ticks, ts=ts, chain_id=100 + c) for c in range(n_chains)],
                     K, NITER, thin=THIN, seed=777)
    assert all(x.status == 0 for x in res)
    got_w = np.stack([x.mcweights[BURN_ROWS:] for x in res])
    got_r = np.stack([x.mcrates[BURN_ROWS:] for x in res])
    return (ref_w, ref_r), (got_w, got_r), res


def test_posterior_means_within_2_percent(chains):
    (ref_w, ref_r), (got_w, got_r), _ = chains
    fr = functionals(ref_w.reshape(-1, K), ref_r.reshape(-1, K))
    fg = functionals(got_w.reshape(-1, K), got_r.reshape(-1, K))
    for name in fr:
        a, b = fr[name].mean(), fg[name].mean()
        assert abs(b / a - 1) < REL_TOL, (name, a, b)


def test_ks_on_thinned_samples(chains):
    (ref_w, ref_r), (got_w, got_r), _ = chains
    # every 5th stored sample (500 iterations apart) of every chain
    fr = functionals(ref_w[:, ::5].reshape(-1, K), ref_r[:, ::5].reshape(-1, K))
    fg = functionals(got_w[:4, ::5].reshape(-1, K), got_r[:4, ::5].reshape(-1, K))
    for name in fr:
        p = stats.ks_2samp(fr[name], fg[name]).pvalue
        assert p > 0.01, (name, p)


def test_component_count_distribution(chains):
    """#components above the 10/N weight cut-off per sample (lmode of process_gibbs)."""
    (ref_w, _), (got_w, _), _ = chains
    hr = np.bincount((ref_w.reshape(-1, K) > 10.0 / N).sum(1), minlength=K + 1) / ref_w.reshape(-1, K).shape[0]
    hg = np.bincount((got_w.reshape(-1, K) > 10.0 / N).sum(1), minlength=K + 1) / got_w.reshape(-1, K).shape[0]
    assert hr.argmax() == hg.argmax()
    assert np.abs(hr - hg).max() < 0.05, (hr, hg)


def test_indicator_consistent_with_parameters(chains):
    """Stored labels are draws from the pre-update parameters: the label histogram of a saved
    row matches the posterior weights to Monte-Carlo accuracy, and labels stay below K."""
    _, _, res = chains
    r0 = res[0]
    assert r0.indicator.shape == ((NITER + 1) // THIN, N) and r0.indicator.dtype == np.uint8
    assert r0.indicator.max() < K
    frac = np.stack([np.bincount(row, minlength=K) for row in r0.indicator[BURN_ROWS:]]) / N
    wmean = r0.mcweights[BURN_ROWS:].mean(axis=0)
    assert np.abs(np.sort(frac.mean(axis=0)) - np.sort(wmean)).max() < 0.01


def test_clustered_tau_matches_reference_posterior(chains):
    """End of the pipeline.  (1) The slowest significant tau = 1/rate per sample -- the quantity the
    reference's tau estimate is built on -- has the reference's median within 2 %, chain by chain.
    (2) process_gibbs (host GaussianMixture, restated from gibbs.py:221-308) runs on a GPU chain and
    its 95 % interval for the slowest tau covers the reference median.  The clustered point estimate
    itself is not compared: on this data set the reference's own chains split 44 % / 38 % between 3
    and 4 components above the weight cut-off, so the cluster count (and with it the histogram-mode
    estimate) flips from chain to chain for the reference as well."""
    from basicrta_b200.gibbs import Gibbs
    (ref_w, ref_r), (got_w, got_r), res = chains
    slow_ref = 1.0 / np.where(ref_w > 10.0 / N, ref_r, np.inf).min(axis=-1)
    slow_got = 1.0 / np.where(got_w > 10.0 / N, got_r, np.inf).min(axis=-1)
    for c in range(slow_got.shape[0]):
        assert abs(np.median(slow_got[c]) / np.median(slow_ref) - 1) < 0.02, c
    times = O.synth_times(N, [0.90, 0.09, 0.01], [5, 0.05, 0.001], seed=20241109)
    gb = Gibbs(times, 'X1', 0, ncomp=K, niter=NITER, cutoff=7.0)
    gb._prepare()
    gb.mcweights, gb.mcrates, gb.indicator = res[0].mcweights, res[0].mcrates, res[0].indicator
    gb.process_gibbs(save=False)
    lo, tau, hi = gb.estimate_tau()
    assert lo < np.median(slow_ref) < hi and lo < tau < hi
    assert gb.processed_results.ncomp in (3, 4)
