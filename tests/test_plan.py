"""Host schedule: every quad of every chain is owned by exactly one CTA, team members share a
wave, no CTA holds two tasks of one wave, slices fit the capacity."""
import numpy as np
import pytest

from basicrta_b200.plan import MIN_SLICE_QUADS, build_plan, shard_chains


def _check(plan, n_quads, grid, cap):
    n_quads = np.asarray(n_quads)
    R = len(n_quads)
    covered = [np.zeros(q, dtype=np.int32) for q in n_quads]
    ranks = [set() for _ in range(R)]
    for b in range(grid):
        waves = []
        for t in plan.tasks_of_cta(b):
            r = int(t['chain'])
            covered[r][t['quad_begin']:t['quad_begin'] + t['quad_count']] += 1
            assert t['team_size'] == plan.team_size[r]
            assert 0 <= t['team_rank'] < t['team_size']
            assert t['team_rank'] not in ranks[r]
            ranks[r].add(int(t['team_rank']))
            assert t['order'] == plan.wave_of_chain[r]
            assert 1 <= t['quad_count'] <= cap
            waves.append(int(t['order']))
        assert waves == sorted(set(waves)), 'a CTA walks its waves in ascending order, one task per wave'
    for r in range(R):
        assert np.all(covered[r] == 1)
        assert ranks[r] == set(range(int(plan.team_size[r])))
    assert plan.cta_task_begin[0] == 0 and plan.cta_task_begin[-1] == len(plan.tasks)
    assert plan.slice_cap_quads == plan.tasks['quad_count'].max()


@pytest.mark.parametrize('seed,R,grid,cap', [(0, 400, 592, 3000), (1, 50, 592, 3000), (2, 2000, 296, 7000),
                                            (3, 1, 592, 3000), (4, 7, 8, 100000), (5, 1000, 16, 500)])
def test_plan_partitions_every_chain(seed, R, grid, cap):
    rng = np.random.default_rng(seed)
    n = np.round(10 ** rng.uniform(2, 5, R)).astype(np.int64)
    q = np.minimum((n + 3) // 4, cap * grid // max(1, R) + 50)
    plan = build_plan(q, grid, cap)
    _check(plan, q, grid, cap)


def test_plan_c2_efficiency_and_giant_chain():
    rng = np.random.default_rng(0)
    q = (np.round(10 ** rng.uniform(4, 5, 400)).astype(np.int64) + 3) // 4
    plan = build_plan(q, 592, 3000)
    assert plan.est_efficiency > 0.85
    giant = build_plan([250000], 592, 3000)                  # config C4 on one GPU
    assert giant.team_size[0] == 592 and giant.n_waves == 1
    _check(giant, [250000], 592, 3000)
    tiny = build_plan([5], 592, 3000)                        # N = 20: one CTA, no exchange
    assert tiny.team_size[0] == 1
    small = build_plan([1250], 592, 3000)                    # config C1: slices stay >= MIN_SLICE_QUADS
    assert small.tasks['quad_count'].min() >= MIN_SLICE_QUADS - 1


def test_plan_rejects_bad_input():
    with pytest.raises(ValueError):
        build_plan([], 8, 100)
    with pytest.raises(ValueError):
        build_plan([0, 3], 8, 100)
    with pytest.raises(ValueError):
        build_plan([10 ** 6], 4, 100)                        # does not fit 4 CTAs x 100 quads


def test_rebalance_team_equalises_measured_times():
    """Measured slicing (engine._calibrate): cutting the measured time density into equal parts
    converges to slices of equal time even when the cost model (uniform here) is wrong."""
    from basicrta_b200.plan import cut_density, density_from_times, rebalance_team
    nq, c = 4000, 7
    true = np.where(np.arange(nq) < 1500, 1.0, 3.0) + 0.5 * np.sin(np.arange(nq) / 300.0)
    cum = np.concatenate(([0.0], np.cumsum(true)))

    def measure(b):
        return cum[b[1:]] - cum[b[:-1]] + 40.0             # + a fixed per-slice overhead

    bounds = np.linspace(0, nq, c + 1).astype(np.int64)
    spread0 = np.ptp(measure(bounds)) / measure(bounds).mean()
    for _ in range(3):
        bounds = rebalance_team(bounds, measure(bounds))
        assert bounds[0] == 0 and bounds[-1] == nq and np.all(np.diff(bounds) >= 1)
    t = measure(bounds)
    assert spread0 > 0.8 and np.ptp(t) / t.mean() < 0.05
    # a slice longer than the shared memory allows: the old boundaries stay
    eq = np.linspace(0, nq, c + 1).astype(np.int64)
    np.testing.assert_array_equal(rebalance_team(eq, measure(eq), cap=600), eq)
    # unusable measurements and degenerate teams: unchanged
    np.testing.assert_array_equal(rebalance_team(eq, np.zeros(c)), eq)
    np.testing.assert_array_equal(rebalance_team([0, 5], [3.0]), [0, 5])
    np.testing.assert_array_equal(rebalance_team([0, 1, 2, 3], [5.0, 1.0, 1.0]), [0, 1, 2, 3])
    # model weights shape the density inside a slice
    d = density_from_times([0, 4], [8.0], quad_weight=[1, 1, 3, 3])
    np.testing.assert_allclose(d, [1, 1, 3, 3])
    np.testing.assert_array_equal(cut_density(d, 2), [0, 3, 4])
    assert cut_density(np.ones(10), 2, cap=4) is None


def test_shard_chains_lpt():
    costs = np.array([100, 90, 50, 40, 30, 20, 10, 5])
    shards = shard_chains(costs, 3)
    assert sorted(np.concatenate(shards).tolist()) == list(range(8))
    loads = [costs[s].sum() for s in shards]
    assert max(loads) - min(loads) <= 20
    assert all(len(s) == 0 for s in shard_chains([], 2))


def test_engine_rebalance_host_logic():
    """GibbsEngine._rebalanced (pure NumPy): slices of a team follow the measured task cycles, optionally
    damped; balanced teams and single-CTA chains are left alone; every chain stays tiled."""
    from types import SimpleNamespace

    from basicrta_b200.engine import GibbsEngine
    from basicrta_b200.plan import TASK_DTYPE
    # chain 0: 4 slices of 1000 quads, chain 1: one slice, chain 2: 2 slices
    tasks = np.array([(0, 4, 0, 0, 1000, 0), (0, 4, 1, 1000, 1000, 0), (0, 4, 2, 2000, 1000, 0), (0, 4, 3, 3000, 1000, 0),
                      (1, 1, 0, 0, 500, 0), (2, 2, 0, 0, 300, 0), (2, 2, 1, 300, 300, 0)], dtype=TASK_DTYPE)
    plan = SimpleNamespace(team_size=np.array([4, 1, 2]), cap_quads=np.array([4000, 4000, 4000]), cap_units=2000,
                           tick_total={})
    ticks = np.ones(4 * (4000 + 500 + 600), dtype=np.uint16)
    tick_offset = np.array([0, 16000, 18000])
    times = np.array([1000.0, 1000.0, 2000.0, 4000.0, 700.0, 500.0, 520.0])

    def tiles(t):
        for r, nq in ((0, 4000), (1, 500), (2, 600)):
            mine = np.sort(t[t['chain'] == r], order='team_rank')
            assert mine['quad_begin'][0] == 0 and mine['quad_begin'][-1] + mine['quad_count'][-1] == nq
            assert np.array_equal(mine['quad_begin'][1:], (mine['quad_begin'] + mine['quad_count'])[:-1])

    full, units = GibbsEngine._rebalanced(plan, tasks, times, None, ticks, tick_offset)
    tiles(full)
    assert tasks['quad_count'][3] == 1000                                 # the input is not modified
    c0 = full[full['chain'] == 0]['quad_count']
    assert c0[3] < 600 and c0[0] > 1500 and units == (int(full['quad_count'].max()) + 1) // 2
    half, _ = GibbsEngine._rebalanced(plan, tasks, times, None, ticks, tick_offset, damping=0.5)
    tiles(half)
    h0 = half[half['chain'] == 0]['quad_count']
    assert c0[3] < h0[3] < 1000 and 1000 < h0[0] < c0[0]                  # half of the way
    lazy, _ = GibbsEngine._rebalanced(plan, tasks, times, None, ticks, tick_offset, min_spread=0.15)
    assert np.array_equal(lazy[lazy['chain'] == 2], tasks[tasks['chain'] == 2])   # 4 % spread: untouched
    assert not np.array_equal(lazy[lazy['chain'] == 0], tasks[tasks['chain'] == 0])
    assert np.array_equal(full[full['chain'] == 1], tasks[tasks['chain'] == 1])
