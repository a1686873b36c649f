"""Bit-exact parity of the CUDA kernel (EXACT mode, through the C ABI) with the oracle.

Teacher forcing: the oracle's per-iteration coefficient rows are injected, so indicators
and (n_k, sum tick_k) of EVERY iteration must be identical -- the north star's criterion
"given identical injected uniforms and gamma draws ... bit-exact".
"""
import numpy as np
import pytest

from basicrta_b200 import _cabi
from basicrta_b200.engine import ChainInput
from oracle import gibbs_oracle as O

pytestmark = pytest.mark.gpu

CASES = [
    # n, true weights, true rates, ncomp, niter, thin
    (1000, [0.9, 0.09, 0.01], [5, 0.05, 0.001], 15, 30, 10),
    (3001, [0.6, 0.3, 0.1], [3.0, 0.3, 0.02], 4, 24, 1),
    (257, [0.5, 0.5], [2.0, 0.1], 2, 20, 5),
    (5000, [0.9, 0.09, 0.01], [5, 0.05, 0.001], 30, 12, 4),
    (2500, [0.7, 0.2, 0.1], [4.0, 0.4, 0.01], 7, 16, 4),      # routed to the K=8 kernel (padded)
]


def _oracle_chain(n, w, r, K, niter, thin, seed, chain_id, uniforms=False):
    times = O.synth_times(n, w, r, seed=seed)
    ts = O.time_step(times)
    ticks = O.to_ticks(times, ts)
    rng = np.random.default_rng(seed + 1000)
    u = None
    if uniforms:
        # on the 2^-23 grid the Philox path produces (brta_batch.inj_u contract)
        u = (np.random.default_rng(seed + 2000).integers(0, 1 << 23, size=(niter, n)).astype(np.float32)
             * np.float32(2.0 ** -23))
    ref = O.run_teacher_forced(ticks, ts, K, niter, seed=99, chain_id=chain_id, rng=rng, g=thin,
                               uniforms=u)
    return ticks, ts, ref, u


@pytest.mark.parametrize('case', CASES)
def test_exact_teacher_forced_philox(engine, case):
    n, w, r, K, niter, thin = case
    chains, refs = [], []
    for cid in range(3):
        ticks, ts, ref, _ = _oracle_chain(n + 17 * cid, w, r, K, niter, thin, seed=10 + cid, chain_id=cid + 5)
        chains.append(ChainInput(ticks=ticks, ts=ts, chain_id=cid + 5))
        refs.append(ref)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE
    res = engine.run(chains, K, niter, thin=thin, seed=99, flags=flags,
                     inject={'coef_c': [x['coef_c'] for x in refs], 'coef_a': [x['coef_a'] for x in refs]})
    for got, ref in zip(res, refs):
        assert got.status == 0
        np.testing.assert_array_equal(got.trace_nk, ref['nk'])
        np.testing.assert_array_equal(got.trace_tk, ref['tk'])
        np.testing.assert_array_equal(got.indicator, ref['indicator'])


def test_exact_injected_uniforms(engine):
    n, K, niter, thin = 1500, 5, 10, 2
    ticks, ts, ref, u = _oracle_chain(n, [0.8, 0.15, 0.05], [5, 0.5, 0.01], K, niter, thin, seed=3,
                                      chain_id=1, uniforms=True)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE | _cabi.FLAG_INJECT_U
    res = engine.run([ChainInput(ticks=ticks, ts=ts, chain_id=1)], K, niter, thin=thin, seed=0, flags=flags,
                     inject={'coef_c': [ref['coef_c']], 'coef_a': [ref['coef_a']], 'u': [u]})[0]
    np.testing.assert_array_equal(res.trace_nk, ref['nk'])
    np.testing.assert_array_equal(res.trace_tk, ref['tk'])
    np.testing.assert_array_equal(res.indicator, ref['indicator'])


def test_team_size_does_not_change_results(engine):
    """Philox is keyed by the datum index, so any split of a chain over CTAs gives the same bits."""
    n, K, niter, thin = 20000, 15, 20, 5
    ticks, ts, ref, _ = _oracle_chain(n, [0.9, 0.09, 0.01], [5, 0.05, 0.001], K, niter, thin, seed=8, chain_id=2)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE
    inj = {'coef_c': [ref['coef_c']], 'coef_a': [ref['coef_a']]}
    db = engine.prepare([ChainInput(ticks=ticks, ts=ts, chain_id=2)], K, niter, thin=thin, seed=99,
                        flags=flags, inject=inj)
    assert db.plan.team_size[0] > 1
    engine.launch(db)
    got = engine.fetch(db)[0]
    np.testing.assert_array_equal(got.trace_nk, ref['nk'])
    np.testing.assert_array_equal(got.trace_tk, ref['tk'])
    np.testing.assert_array_equal(got.indicator, ref['indicator'])


def test_measured_slicing_changes_the_schedule_not_the_results(engine):
    """engine._calibrate times every slice in a short launch and moves the slice boundaries inside the
    teams; Philox is keyed by the datum, statistics are integers: the output must not change by a bit."""
    rng = np.random.default_rng(12)
    chains = []
    for r, n in enumerate((30000, 9000, 52001, 700)):
        comp = rng.choice(3, size=n, p=[0.8, 0.15, 0.05])
        x = rng.exponential(1.0 / np.array([4.0, 0.1, 0.002])[comp])
        chains.append(ChainInput(ticks=np.maximum(np.ceil(x / 0.1), 1).astype(np.int64), ts=0.1, chain_id=40 + r))
    K, niter, thin = 15, 60, 20
    plain = engine.prepare(chains, K, niter, thin=thin, seed=5, calibrate=False)
    tuned = engine.prepare(chains, K, niter, thin=thin, seed=5, calibrate=True, choose_kernel=False)
    t0, t1 = plain.plan.tasks, tuned.plan.tasks
    assert len(t0) == len(t1) and np.array_equal(t0['chain'], t1['chain']) and np.array_equal(t0['team_rank'], t1['team_rank'])
    assert not np.array_equal(t0['quad_begin'], t1['quad_begin'])          # boundaries moved
    for r, ch in enumerate(chains):                                        # and still tile every chain
        mine = np.sort(t1[t1['chain'] == r], order='team_rank')
        assert mine['quad_begin'][0] == 0 and np.all(mine['quad_count'] >= 1)
        assert np.array_equal(mine['quad_begin'][1:], (mine['quad_begin'] + mine['quad_count'])[:-1])
        assert mine['quad_begin'][-1] + mine['quad_count'][-1] == (len(ch.ticks) + 3) // 4
    engine.launch(plain)
    engine.launch(tuned)
    for a, b in zip(engine.fetch(plain), engine.fetch(tuned)):
        assert a.status == 0 and b.status == 0
        np.testing.assert_array_equal(a.mcweights, b.mcweights)
        np.testing.assert_array_equal(a.mcrates, b.mcrates)
        np.testing.assert_array_equal(a.indicator, b.indicator)


def test_kernel_build_choice_changes_the_schedule_not_the_results(engine):
    """K <= 16 exists in two builds (4 CTAs per SM / 3 CTAs per SM with larger slices, BRTA_FLAG_CTAS3);
    ``prepare`` times a short launch of each and keeps the faster schedule.  Either build, and the measured
    choice, must give the same bits."""
    rng = np.random.default_rng(33)
    chains = []
    for r, n in enumerate((61000, 9000, 33001, 1200)):
        comp = rng.choice(3, size=n, p=[0.8, 0.15, 0.05])
        x = rng.exponential(1.0 / np.array([4.0, 0.1, 0.002])[comp])
        chains.append(ChainInput(ticks=np.maximum(np.ceil(x / 0.1), 1).astype(np.int64), ts=0.1, chain_id=90 + r))
    K, niter, thin = 15, 80, 20
    four = engine.prepare(chains, K, niter, thin=thin, seed=6, calibrate=False)
    three = engine.prepare(chains, K, niter, thin=thin, seed=6, calibrate=False, flags=_cabi.FLAG_CTAS3)
    auto = engine.prepare(chains, K, niter, thin=thin, seed=6, calibrate=True, choose_kernel=True)
    sm = engine.caps.sm_count
    assert four.plan.grid == 4 * sm and three.plan.grid == 3 * sm
    assert auto.kernel_choice['ctas_per_sm'] in (3, 4) and auto.plan.grid == auto.kernel_choice['ctas_per_sm'] * sm
    assert auto.kernel_choice['ms_3'] > 0 and auto.kernel_choice['ms_4'] > 0
    results = []
    for db in (four, three, auto):
        engine.launch(db)
        results.append(engine.fetch(db))
    for other in results[1:]:
        for a, b in zip(results[0], other):
            assert a.status == 0 and b.status == 0
            np.testing.assert_array_equal(a.mcweights, b.mcweights)
            np.testing.assert_array_equal(a.mcrates, b.mcrates)
            np.testing.assert_array_equal(a.indicator, b.indicator)


def test_segmented_run_is_bit_identical_to_one_launch(engine):
    """Long calibrated runs go as several back-to-back launches (iter_begin / iter_end, state handed over in
    final_c / final_a) with the slices re-cut in between from measured cycles.  Iteration numbers, Philox
    counters and row indices are those of the whole run: same bits as a single launch, also when the same
    batch is launched a second time."""
    rng = np.random.default_rng(21)
    chains = []
    for r, n in enumerate((26000, 7001, 41000, 900, 15000)):
        comp = rng.choice(3, size=n, p=[0.7, 0.2, 0.1])
        x = rng.exponential(1.0 / np.array([3.0, 0.2, 0.004])[comp])
        chains.append(ChainInput(ticks=np.maximum(np.ceil(x / 0.1), 1).astype(np.int64), ts=0.1, chain_id=70 + r))
    K, niter, thin = 15, 600, 20
    one = engine.prepare(chains, K, niter, thin=thin, seed=9, calibrate=False)
    engine.launch(one)
    ref = engine.fetch(one)
    cut = engine.prepare(chains, K, niter, thin=thin, seed=9, calibrate=True, segments=(0.1, 0.2, 0.45, 0.7, 0.9),
                         choose_kernel=False)
    assert cut.segments == [60, 120, 280, 420, 540, 600]
    for _ in range(2):
        engine.reset(cut)
        engine.launch(cut)
        got = engine.fetch(cut)
        for a, b in zip(ref, got):
            assert b.status == 0
            np.testing.assert_array_equal(a.mcweights, b.mcweights)
            np.testing.assert_array_equal(a.mcrates, b.mcrates)
            np.testing.assert_array_equal(a.indicator, b.indicator)


def test_fast_mode_flip_rate(engine):
    """FAST (MUFU.EX2 + FMA contraction) differs from EXACT only where a uniform lands within
    float rounding of a CDF boundary: < 1e-4 of the labels, and the statistics stay close."""
    n, K, niter, thin = 20000, 15, 20, 1
    ticks, ts, ref, _ = _oracle_chain(n, [0.9, 0.09, 0.01], [5, 0.05, 0.001], K, niter, thin, seed=4, chain_id=3)
    flags = _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE
    inj = {'coef_c': [ref['coef_c']], 'coef_a': [ref['coef_a']]}
    got = engine.run([ChainInput(ticks=ticks, ts=ts, chain_id=3)], K, niter, thin=thin, seed=99,
                     flags=flags, inject=inj)[0]
    flips = np.mean(got.indicator != ref['indicator'])
    assert flips < 1e-4, flips
    assert np.abs(got.trace_nk - ref['nk']).max() <= 8


def test_fast_mode_underflow_takes_the_safe_path(engine):
    """FAST mode drops the max subtraction; a datum whose every term underflows float32 (possible only
    in wild transient states) must be redone with the max-subtracted form.  Coefficients with slopes of
    ~1 log2 unit per tick push the logits of most data below -126: labels must still match EXACT."""
    K, niter, n = 3, 4, 4096
    rng = np.random.default_rng(5)
    ticks = rng.integers(1, 3000, size=n).astype(np.int64)
    coef_c = np.tile(np.array([0.0, -1.0, -2.0], np.float32), (niter, 1))
    coef_a = np.tile(np.array([1.0, 0.9, 0.8], np.float32), (niter, 1))
    out = {}
    for name, flags in (('exact', _cabi.FLAG_EXACT), ('fast', 0)):
        out[name] = engine.run([ChainInput(ticks=ticks, ts=0.1, chain_id=1)], K, niter, thin=1, seed=3,
                               flags=flags | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE,
                               inject={'coef_c': [coef_c], 'coef_a': [coef_a]})[0]
        assert out[name].status == 0
    assert np.mean(out['fast'].indicator[:niter] != out['exact'].indicator[:niter]) < 1e-3
    assert np.abs(out['fast'].trace_nk - out['exact'].trace_nk).max() <= 4
    big = ticks > 200                                      # every term below 2^-126 without the max
    assert big.sum() > 3000 and np.all(out['fast'].indicator[0][big] == 2)


@pytest.mark.parametrize('K', [3, 15, 30])
def test_memoised_rows_are_bit_identical_to_recomputation(engine, K):
    """Sharing the cumulative rows of equal ticks is a pure memoisation: free-running FAST chains
    with and without the table produce identical weights, rates and labels."""
    times = O.synth_times(30000, [0.7, 0.2, 0.1], [5, 0.3, 0.004], seed=12)
    ticks = O.to_ticks(times, 0.1)
    chains = [ChainInput(ticks=ticks, ts=0.1, chain_id=7), ChainInput(ticks=ticks[:1237], ts=0.1, chain_id=8)]
    a = engine.run(chains, K, 400, thin=50, seed=21)
    b = engine.run(chains, K, 400, thin=50, seed=21, flags=_cabi.FLAG_NO_TABLE)
    for x, y in zip(a, b):
        assert x.status == 0 and y.status == 0
        assert np.array_equal(x.mcrates, y.mcrates) and np.array_equal(x.mcweights, y.mcweights)
        assert np.array_equal(x.indicator, y.indicator)
