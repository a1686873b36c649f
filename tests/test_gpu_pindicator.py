"""the per-datum cluster-membership accumulation of ``Gibbs.cluster``
(basicrta/gibbs.py:264-268) as a device kernel, against the reference's own double loop and the
NumPy restatement.  Integer work: the bar is exact equality."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _reference_loop(indicator, rows, comps, labels, n_data, n_clusters):
    """gibbs.py:264-268, verbatim semantics (small cases only)."""
    pind = np.zeros((n_data, n_clusters))
    for j in np.unique(rows):
        mapinds = labels[rows == j]
        for i, indx in enumerate(comps[rows == j]):
            tmpind = np.where(indicator[j] == indx)[0]
            pind[tmpind, mapinds[i]] += 1
    return pind


def _case(rng, S, N, K, C, keep=0.6):
    indicator = rng.integers(0, K, size=(S, N)).astype(np.uint8)
    active = rng.random((S, K)) < keep
    rows, comps = np.where(active)
    labels = rng.integers(0, C, size=len(rows))
    return indicator, rows, comps, labels


@pytest.mark.parametrize('S,N,K,C', [(1, 1, 2, 1), (7, 257, 15, 3), (129, 5000, 15, 6), (300, 1031, 30, 32),
                                     (1000, 20000, 15, 4),
                                     # every row alignment class (N mod 4), more rows than one chunk of byte
                                     # counters, and 1 / 2 / 4 counter words per datum in the aligned-class kernel
                                     (1300, 4096, 15, 4), (1001, 4097, 15, 4), (1001, 4098, 30, 6), (517, 4099, 15, 12),
                                     (260, 3000, 30, 12), (9, 3, 64, 4), (2, 70001, 15, 2)])
def test_counts_match_numpy_and_reference_loop(S, N, K, C):
    from basicrta_b200.engine import pindicator_counts
    from basicrta_b200.postprocess import pindicator_counts_host
    rng = np.random.default_rng(S * 31 + N)
    indicator, rows, comps, labels = _case(rng, S, N, K, C)
    host = pindicator_counts_host(indicator, rows, comps, labels, C, K)
    dev = pindicator_counts_host(indicator, rows, comps, labels, C, K, device=0)
    assert dev.dtype == np.int32 and dev.shape == (N, C)
    np.testing.assert_array_equal(dev, host)
    if S * N <= 129 * 5000:
        np.testing.assert_array_equal(dev, _reference_loop(indicator, rows, comps, labels, N, C))
    # the same from a strided tensor that already lives on the GPU (rows longer than N)
    import torch
    big = torch.zeros((S, N + 37), dtype=torch.uint8, device='cuda:0')
    big[:, :N] = torch.from_numpy(indicator).cuda()
    lut = np.full((S, K), -1, dtype=np.int8)
    lut[rows, comps] = labels
    np.testing.assert_array_equal(pindicator_counts(big[:, :N], lut, C), host)


def test_both_kernels_give_the_same_counts(monkeypatch):
    """The aligned-class kernel (default where the table fits) against the generic one (developer knob)."""
    from basicrta_b200.engine import pindicator_counts
    rng = np.random.default_rng(5)
    indicator, rows, comps, labels = _case(rng, 1003, 30011, 15, 4)
    indicator[::7, ::5] = 77                                               # labels beyond ncomp: never counted
    lut = np.full((1003, 15), -1, dtype=np.int8)
    lut[rows, comps] = labels
    fast = pindicator_counts(indicator, lut, 4)
    monkeypatch.setenv('BRTA_PINDICATOR_GENERIC', '1')
    np.testing.assert_array_equal(pindicator_counts(indicator, lut, 4), fast)
    assert fast.sum() == (lut[np.arange(1003)[:, None], np.minimum(indicator, 14)] >= 0)[indicator < 15].sum()


def test_out_of_range_labels_and_empty_rows_are_ignored():
    from basicrta_b200.engine import pindicator_counts
    indicator = np.array([[0, 1, 2, 200], [3, 3, 3, 3]], dtype=np.uint8)
    lut = np.array([[0, -1, 1, 1], [-1, -1, -1, -1]], dtype=np.int8)
    got = pindicator_counts(indicator, lut, 2)
    np.testing.assert_array_equal(got, [[1, 0], [0, 0], [0, 1], [0, 0]])
    assert pindicator_counts(np.zeros((0, 5), np.uint8), np.zeros((0, 3), np.int8), 2).shape == (5, 2)
    with pytest.raises(ValueError):
        pindicator_counts(indicator.astype(np.int32), lut, 2)
    with pytest.raises(ValueError):
        pindicator_counts(indicator, lut[:1], 2)
    with pytest.raises(ValueError):
        pindicator_counts(indicator, lut, 33)


def test_cluster_on_device_equals_cluster_on_host(tmp_path, monkeypatch):
    """End to end through the reference-facing API: Gibbs.run -> cluster(device=0) gives the same
    processed_results.indicator as the NumPy accumulation."""
    from basicrta_b200.gibbs import Gibbs
    from oracle import gibbs_oracle as O
    monkeypatch.chdir(tmp_path)
    times = O.synth_times(4000, [0.8, 0.15, 0.05], [4.0, 0.1, 0.003], seed=3)
    g = Gibbs(times, 'P1', 0, ncomp=8, niter=6000, cutoff=7.0)
    g.burnin = 2000
    g.seed = 17
    g.run()
    g.cluster(n_components=3, n_init=3, random_state=0)
    host = g.processed_results.indicator.copy()
    g.cluster(n_components=3, n_init=3, random_state=0, device=0)
    np.testing.assert_array_equal(g.processed_results.indicator, host)
    assert np.allclose(np.nansum(host, axis=1)[~np.isnan(host).any(axis=1)], 1.0)


def test_labels_kept_on_the_device_reduce_to_the_same_counts(tmp_path, monkeypatch):
    """SURVEY.md 8 f-1 end to end: ``run(keep_indicator_on_device=True)`` leaves the label rows in HBM (the
    pickle holds ``indicator = None``, which the reference's ``cluster`` understands, gibbs.py:259-262);
    ``process_gibbs`` then reduces them with brta_pindicator_counts and only [N, clusters] integers cross
    PCIe.  Same seed => same chain, so the result must equal the host loop on the downloaded labels."""
    import pickle
    from basicrta_b200.gibbs import Gibbs
    from oracle import gibbs_oracle as O
    monkeypatch.chdir(tmp_path)
    times = O.synth_times(6000, [0.8, 0.15, 0.05], [4.0, 0.1, 0.003], seed=5)

    def make(name):
        g = Gibbs(times, name, 0, ncomp=8, niter=6000, cutoff=7.0)
        g.burnin, g.seed = 2000, 23
        return g
    host = make('H1')
    host.run()
    dev = make('H1')                                           # same residue name => same Philox chain id
    dev.run(keep_indicator_on_device=True)
    assert dev.indicator is None and dev.device_indicator() is not None
    assert tuple(dev.device_indicator().shape) == (60, 6000) and dev.device_indicator().is_cuda
    np.testing.assert_array_equal(dev.device_indicator().cpu().numpy(), host.indicator)
    with open('basicrta-7.0/H1/gibbs_6000.pkl', 'rb') as f:
        from basicrta_b200.gibbs import load_reference_pickle
        assert load_reference_pickle(f).indicator is None      # the pickle does not carry the labels
    host.cluster(n_components=3, n_init=3, random_state=0)
    dev.cluster(n_components=3, n_init=3, random_state=0)       # picks the resident rows up by itself
    np.testing.assert_array_equal(dev.processed_results.indicator, host.processed_results.indicator)
    np.testing.assert_array_equal(dev.processed_results.labels, host.processed_results.labels)
