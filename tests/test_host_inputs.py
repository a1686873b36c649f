"""Host-side input handling: the device tick grid (``ts`` of the reference is the FIRST gap,
not the trajectory step), continuous data, the survival function without a billion-bin histogram,
per-residue failure collection, shard boundaries.  No GPU needed."""
import numpy as np
import pytest

from basicrta_b200 import util
from basicrta_b200.engine import shard_bounds, tick_grid, times_to_ticks
from basicrta_b200.gibbs import Gibbs, GibbsBatchError, run_batch


def _ref_ts(x):
    srt = np.sort(x)
    d = srt[1:] - srt[:-1]
    nz = d[d != 0]
    return nz[0] if len(nz) else x.min()


def test_sparse_residues_get_the_trajectory_grid():
    """{0.3, 0.7, 0.8} has first gap 0.4 but lives on dt = 0.1: 73 % of random residues with 3-11 contacts
    were rejected in round 1."""
    ticks, grid = tick_grid(np.array([0.3, 0.7, 0.8]), 0.4)
    assert np.isclose(grid, 0.1) and ticks.tolist() == [3, 7, 8]
    rng = np.random.default_rng(0)
    for _ in range(500):
        n = int(rng.integers(2, 12))
        x = rng.integers(1, 60, n) * 0.1
        ticks, grid = tick_grid(x, _ref_ts(x))
        assert np.allclose(ticks * grid, x, atol=1e-9)
        k = _ref_ts(x) / grid
        assert abs(k - round(k)) < 1e-6 and ticks.min() >= 1
    with pytest.raises(ValueError):
        times_to_ticks(np.array([0.3, 0.7, 0.8]), 0.4)       # the strict form still refuses


def test_grid_data_keep_their_grid_and_float32_input_works():
    x = np.maximum(np.ceil(np.random.default_rng(1).exponential(3.0, 20000) / 0.1), 1) * 0.1
    ticks, grid = tick_grid(x, _ref_ts(x))
    assert np.isclose(grid, 0.1) and np.array_equal(ticks, np.rint(x / 0.1).astype(np.int64))
    t32, g32 = tick_grid(x.astype(np.float32), np.float32(0.1))
    assert np.array_equal(t32, ticks)


def test_continuous_times_go_to_fixed_point():
    """The reference's own example data (util.simulate_hn, util.py:596-608) are continuous."""
    x = util.simulate_hn(1e4, [0.9, 0.09, 0.01], [5, 0.05, 0.001], seed=3)
    ticks, grid = tick_grid(x, _ref_ts(x))
    assert ticks.max() < (1 << 23) and ticks.min() >= 0
    assert np.abs(ticks * grid - x).max() <= 0.5 * grid * (1 + 1e-9)
    assert np.sort(ticks)[-(1 << 17):].sum() < (1 << 32)
    # data whose largest values would overflow a 32-bit slice sum get a coarser grid
    y = np.random.default_rng(2).uniform(0.5, 1.0, 200000) + 1e-7 * np.arange(200000)
    t2, g2 = tick_grid(y, _ref_ts(y))
    assert np.sort(t2)[-(1 << 17):].sum() < (1 << 32) and np.abs(t2 * g2 - y).max() <= 0.5 * g2 * (1 + 1e-9)


def test_tick_grid_rejects_garbage():
    for bad in (np.array([]), np.array([1.0, np.nan]), np.array([-1.0, 2.0]), np.zeros(4)):
        with pytest.raises(ValueError):
            tick_grid(bad, 0.1)


def test_survival_function_sparse_equals_dense(monkeypatch):
    rng = np.random.default_rng(0)
    for trial in range(100):
        n = int(rng.integers(3, 300))
        x = rng.exponential(3.0, n) if trial % 2 else np.maximum(np.ceil(rng.exponential(3.0, n) / 0.1), 1) * 0.1
        ts = _ref_ts(x)
        if int(x.max() // ts) + 2 > 3e6:
            continue
        dense = util.make_surv(np.histogram(x, bins=util.get_bins(x, ts)))
        monkeypatch.setattr(util, 'DENSE_BINS_MAX', 0)
        sparse = util.get_s(x, ts)
        monkeypatch.setattr(util, 'DENSE_BINS_MAX', 1 << 22)
        assert np.array_equal(dense[0], sparse[0]) and np.array_equal(dense[1], sparse[1])


def test_prepare_on_continuous_data_is_cheap():
    """the reference needs 43 s and 8.6 GB here (1e9 histogram bins)."""
    import time
    x = util.simulate_hn(5000, [0.9, 0.09, 0.01], [5, 0.05, 0.001], seed=1)
    gb = Gibbs(x, 'X1', 0, ncomp=15, niter=1000)
    t0 = time.perf_counter()
    gb._prepare()
    assert time.perf_counter() - t0 < 2.0
    assert gb.t[0] == 0 and gb.s[0] == 1 and abs(gb.s[-1]) < 1e-12 and np.all(np.diff(gb.s) < 0)
    ci = gb._chain_input()
    assert ci.ticks.max() < (1 << 23)


def test_bad_residues_fail_alone_before_any_gpu_work():
    """A batch whose every member is unusable reports them all and never touches the device."""
    a = Gibbs(np.array([0.1, np.nan, 0.3]), 'A1', 0, ncomp=3, niter=100, cutoff=7.0)
    b = Gibbs(np.array([-0.1, 0.2, 0.3]), 'B2', 0, ncomp=3, niter=100, cutoff=7.0)
    with pytest.raises(GibbsBatchError) as ei:
        run_batch([a, b])
    assert [r for r, _ in ei.value.failures] == ['A1', 'B2']
    assert 'A1' in str(ei.value) and 'B2' in str(ei.value)


def test_shard_bounds():
    b = shard_bounds(1_000_000, 8)
    assert b[0] == 0 and b[-1] == 250000 and len(b) == 9
    assert all((x % 2) == 0 for x in b[:-1]) and np.all(np.diff(b) > 0)
    assert max(np.diff(b)) - min(np.diff(b)) <= 2 * 8
    assert shard_bounds(60001, 2) == [0, 7502, 15001]
    with pytest.raises(ValueError):
        shard_bounds(9, 8)


def test_memo_table_geometry_matches_the_kernel_source():
    """memo.py mirrors compile-time constants of csrc/brta_sweep.cuh (the host cuts slices with them)."""
    import os
    import re
    from basicrta_b200 import memo
    src = open(os.path.join(os.path.dirname(memo.__file__), 'csrc', 'brta_sweep.cuh')).read()
    assert int(re.search(r'constexpr int TABLE_FLOATS = (\d+);', src).group(1)) == memo.TABLE_FLOATS
    assert int(re.search(r'constexpr int PACKED_MAX_ROWS = (\d+);', src).group(1)) == memo.MAX_ROWS
    assert 'TABLE_FLOATS / table_row_stride(k) < 256' in src and memo.MAX_ROWS == 256
    # rows = floats / odd row stride, capped: K = 15 -> 240, K = 30 -> 124, K = 2 -> capped at 256
    assert [memo.table_rows(k) for k in (15, 16, 30, 2)] == [240, 240, 124, 256]


def test_gmm_host_packing_and_argument_checks():
    """Host side of basicrta_b200.gmm (no GPU): parameter blocks, problem packing, sklearn-style errors."""
    from basicrta_b200 import gmm
    w = np.array([0.5, 0.3, 0.2])
    mu = np.arange(6.0).reshape(3, 2)
    cov = np.array([[[2.0, 0.3], [0.3, 1.0]], [[1.0, 0.0], [0.0, 4.0]], [[0.5, -0.1], [-0.1, 0.7]]])
    block = gmm.pack_params(w, mu, cov)
    assert block.shape == (gmm.KMAX, 6) and np.all(block[3:] == 0)
    w2, mu2, cov2 = gmm._unpack_params(block, 3)
    assert np.array_equal(w2, w) and np.array_equal(mu2, mu) and np.array_equal(cov2, cov)
    p = gmm._precisions_cholesky(cov)
    assert np.allclose(np.einsum('kij,klj->kil', p, p) @ cov, np.eye(2))          # P P^T = cov^-1
    assert np.all(p[:, 1, 0] == 0)                                                 # upper triangular, like sklearn's
    flat, offsets, sizes = gmm._pack([np.zeros((5, 2)), np.ones((3, 2))])
    assert flat.shape == (8, 2) and offsets.tolist() == [0, 5, 8] and sizes.tolist() == [5, 3]
    with pytest.raises(ValueError):
        gmm._pack([np.zeros((5, 3))])
    with pytest.raises(ValueError):
        gmm.GaussianMixture(n_components=2, init_params='random')
    assert gmm.fit_batch([], 3, device=0) == []                                    # nothing to do: no GPU touched


def test_packed_statistics_fields_cannot_overflow():
    """csrc/brta_sweep.cuh packs (count, sum of tick offsets) of a served datum into one 32-bit atomic: 12 + 20 bits,
    eight accumulator sets per CTA.  The compile-time limits must keep both fields from overflowing."""
    import os
    import re
    from basicrta_b200 import memo
    src = open(os.path.join(os.path.dirname(memo.__file__), 'csrc', 'brta_sweep.cuh')).read()
    max_quads = int(re.search(r'constexpr int PACKED_MAX_QUADS = (\d+);', src).group(1))
    max_rows = int(re.search(r'constexpr int PACKED_MAX_ROWS = (\d+);', src).group(1))
    # a set (warp, position in the unrolled pair) sees 128 data per loop iteration of 256 quads, plus one batch of
    # the single-quad remainder loop
    per_set = (max_quads // 256 + 1) * 128
    assert per_set < (1 << 12)
    assert per_set * (max_rows - 1) < (1 << 20)
    assert '(1u << 20)' in src and 'w >> 20' in src and '0xfffffu' in src


def test_reprocess_batch_skips_residues_that_cannot_be_clustered():
    """cluster.py:44-52: a residue whose clustering raises is skipped, the others go on.  Here every residue is
    unusable (no samples after burn-in), so nothing reaches the GPU."""
    import warnings
    from basicrta_b200 import postprocess

    class Stub(object):
        pass
    g = Stub()
    g.burnin, g.g, g.times, g.residue, g.cutoff = 10000, 100, np.ones(10), 'X1', 7.0
    g.mcweights, g.mcrates = np.zeros((5, 3)), np.ones((5, 3))
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        assert postprocess.reprocess_batch([g], device=0) == []
