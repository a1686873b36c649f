"""32 < ncomp <= 255: the reference stores labels as uint8 (basicrta/gibbs.py:167-168) and therefore accepts up
to 255 components.  Such mixtures run through the general kernel (csrc/brta_wide.cu: one CTA per chain, three
passes over the components per datum) behind the same C ABI; it must meet the same bar as the team kernel:
EXACT mode teacher-forced = the oracle bit for bit, free running = a working sampler with the reference's
output layout."""
import os

import numpy as np
import pytest

from basicrta_b200 import _cabi
from basicrta_b200.engine import ChainInput
from oracle import gibbs_oracle as O

pytestmark = pytest.mark.gpu


def _oracle_chain(n, K, niter, thin, seed, chain_id, uniforms=False, ts_scale=1.0):
    times = O.synth_times(n, [0.9, 0.09, 0.01], [5, 0.05, 0.001], seed=seed) * ts_scale
    ts = O.time_step(times)
    ticks = O.to_ticks(times, ts)
    u = None
    if uniforms:
        u = (np.random.default_rng(seed + 2000).integers(0, 1 << 23, size=(niter, n)).astype(np.float32)
             * np.float32(2.0 ** -23))
    ref = O.run_teacher_forced(ticks, ts, K, niter, seed=99, chain_id=chain_id, rng=np.random.default_rng(seed + 1000),
                               g=thin, uniforms=u)
    return ticks, ts, ref, u


@pytest.mark.parametrize('K,n', [(33, 1000), (64, 777), (255, 300)])
def test_exact_teacher_forced_philox(engine, K, n):
    niter, thin = 8, 2
    chains, refs = [], []
    for cid in range(2):
        ticks, ts, ref, _ = _oracle_chain(n + 13 * cid, K, niter, thin, seed=20 + cid, chain_id=cid + 3)
        chains.append(ChainInput(ticks=ticks, ts=ts, chain_id=cid + 3))
        refs.append(ref)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE
    res = engine.run(chains, K, niter, thin=thin, seed=99, flags=flags,
                     inject={'coef_c': [x['coef_c'] for x in refs], 'coef_a': [x['coef_a'] for x in refs]})
    for got, ref in zip(res, refs):
        assert got.status == 0
        np.testing.assert_array_equal(got.trace_nk, ref['nk'])
        np.testing.assert_array_equal(got.trace_tk, ref['tk'])
        np.testing.assert_array_equal(got.indicator, ref['indicator'])


def test_exact_injected_uniforms_and_uint32_ticks(engine):
    K, n, niter, thin = 40, 901, 6, 1
    ticks, ts, ref, u = _oracle_chain(n, K, niter, thin, seed=5, chain_id=1, uniforms=True)
    big = ticks.copy()
    big[::50] += 70000                                        # forces the uint32 tick layout
    ref = O.run_teacher_forced(big, ts, K, niter, seed=99, chain_id=1, rng=np.random.default_rng(7), g=thin, uniforms=u)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE | _cabi.FLAG_INJECT_U
    got = engine.run([ChainInput(ticks=big, ts=ts, chain_id=1)], K, niter, thin=thin, seed=0, flags=flags,
                     inject={'coef_c': [ref['coef_c']], 'coef_a': [ref['coef_a']], 'u': [u]})[0]
    assert got.status == 0
    np.testing.assert_array_equal(got.trace_nk, ref['nk'])
    np.testing.assert_array_equal(got.trace_tk, ref['tk'])
    np.testing.assert_array_equal(got.indicator, ref['indicator'])


def test_free_running_through_the_api(tmp_path, monkeypatch):
    """Gibbs(ncomp = 40).run(): the reference's layout, and the sampler finds the three components."""
    from basicrta_b200.gibbs import Gibbs
    monkeypatch.chdir(tmp_path)
    times = O.synth_times(4000, [0.90, 0.09, 0.01], [5, 0.05, 0.001], seed=11)
    g = Gibbs(times, 'X7', 0, ncomp=40, niter=4000, cutoff=7.0)
    g.seed = 3
    g.run()
    assert os.path.exists('basicrta-7.0/X7/gibbs_4000.pkl')
    assert g.mcweights.shape == (40, 40) and g.mcrates.shape == (40, 40)
    assert g.indicator.shape == (40, 4000) and g.indicator.dtype == np.uint8 and g.indicator.max() < 40
    assert np.allclose(g.mcweights.sum(axis=1), 1.0, atol=1e-5)
    w, r = g.mcweights[20:], g.mcrates[20:]
    fast = r[w > 0.5]
    assert len(fast) and 3.0 < fast.mean() < 5.0, fast.mean()     # ~3.9: the 0.1 ns ceiling quantisation, as for K = 15
    slow = np.where(w > 10 / 4000, r, np.inf).min(axis=1)
    assert 0.0004 < np.median(slow) < 0.003
    # label histogram of a stored row follows the weights
    frac = np.bincount(g.indicator[-1], minlength=40) / 4000
    assert abs(frac.max() - w[-1].max()) < 0.05


def test_rejects_what_it_cannot_do(engine):
    ticks = np.arange(1, 200)
    with pytest.raises(ValueError):
        engine.prepare([ChainInput(ticks=ticks, ts=0.1)], 256, 10)
    with pytest.raises(ValueError):
        engine.prepare([ChainInput(ticks=ticks, ts=0.1)], 64, 10, flags=_cabi.FLAG_CTAS3)
