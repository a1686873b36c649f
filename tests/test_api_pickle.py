"""The drop-in boundary on the host: constructor / attributes / pickle layout of ``Gibbs``,
round trips through OUR loader and through the REFERENCE's ``Gibbs.load`` + ``cluster``
(the latter only where /root/reference exists, i.e. in the build container)."""
import os
import pickle

import numpy as np
import pytest

from basicrta_b200 import postprocess
from basicrta_b200.gibbs import Gibbs, ParallelGibbs, load_reference_pickle
from basicrta_b200.cluster import ProcessCluster
from oracle import gibbs_oracle as O

from _refstubs import have_reference, import_reference


def _filled_gibbs(niter=3000, g=10, n=600, K=4, burnin=500):
    """A Gibbs instance whose arrays come from the oracle (test-only; the product has no CPU sampler)."""
    times = O.synth_times(n, [0.7, 0.3], [3.0, 0.05], seed=3)
    gb = Gibbs(times, 'W313', 2, ncomp=K, niter=niter, cutoff=7.0)
    gb.g, gb.burnin = g, burnin
    gb._prepare()
    out = O.run_reference_order(times, K, niter, np.random.default_rng(5), g=g)
    gb.mcweights, gb.mcrates, gb.indicator = out['mcweights'], out['mcrates'], out['indicator']
    return gb


def test_constructor_defaults_and_ts():
    gb = Gibbs()
    assert (gb.times, gb.residue, gb.loc, gb.ncomp, gb.niter, gb.cutoff) == (None, None, 0, 15, 110000, None)
    assert (gb.g, gb.burnin, gb._noise_cutoff, gb.ts) == (100, 10000, 0.4, None)
    assert gb.keys == {'times', 'residue', 'loc', 'ncomp', 'niter', 'g', 'burnin', 'processed_results', 'ts',
                       'mcweights', 'mcrates', 't', 's', 'cutoff', 'indicator'}
    gb = Gibbs(np.array([0.3, 0.1, 0.1, 0.5]), 'X1')
    assert np.isclose(gb.ts, 0.2) and gb['residue'] == 'X1'
    gb = Gibbs(np.array([0.4, 0.4]), 'X1')                 # no non-zero gap: falls back to min (gibbs.py:150-151)
    assert gb.ts == 0.4


def test_save_load_roundtrip_and_bak_rotation(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    gb = _filled_gibbs()
    with pytest.raises(OSError):
        gb.save()                                           # directory missing (gibbs.py:348-349)
    os.makedirs('basicrta-7.0/W313')
    gb.save()
    gb.save()
    assert os.path.exists('basicrta-7.0/W313/gibbs_3000.pkl.bak')
    back = Gibbs.load('basicrta-7.0/W313/gibbs_3000.pkl')
    for attr in ('times', 'mcweights', 'mcrates', 'indicator', 't', 's', 'whypers', 'rhypers'):
        assert np.array_equal(getattr(back, attr), getattr(gb, attr)), attr
    assert (back.residue, back.ncomp, back.niter, back.g, back.burnin, back.cutoff) == ('W313', 4, 3000, 10, 500, 7.0)
    raw = load_reference_pickle(open('basicrta-7.0/W313/gibbs_3000.pkl', 'rb'))
    assert raw.indicator.dtype == np.uint8 and raw.indicator.flags['C_CONTIGUOUS'] and raw.indicator.flags.writeable
    assert raw.mcweights.dtype == np.float64 and raw.mcweights.shape == (300, 4)


def test_pickle_stream_names_the_reference_classes(tmp_path, monkeypatch):
    """the pickle must import as ``basicrta.gibbs.Gibbs`` (and its ``processed_results`` as
    ``MDAnalysis.analysis.base.Results``) so that a stock basicrta -- cluster.ProcessProtein,
    cluster.py:44-46 -- loads it on a machine that has never seen this package."""
    import io
    import pickletools
    import sys
    monkeypatch.chdir(tmp_path)
    gb = _filled_gibbs()
    gb._device_indicator = object()                          # run-time handle: must not be pickled
    os.makedirs('basicrta-7.0/W313')
    gb.save()
    raw = open('basicrta-7.0/W313/gibbs_3000.pkl', 'rb').read()
    strings = {a for op, a, _ in pickletools.genops(io.BytesIO(raw)) if op.name in ('SHORT_BINUNICODE', 'BINUNICODE')}
    assert 'basicrta.gibbs' in strings and 'MDAnalysis.analysis.base' in strings
    assert not any('basicrta_b200' in str(x) for x in strings) and '_device_indicator' not in strings
    # a foreign environment: modules of those names that know nothing about basicrta_b200
    import types

    class ForeignGibbs:
        def __getitem__(self, k):
            return getattr(self, k)

    class ForeignResults:
        def __setstate__(self, state):
            self.data = state

    saved = {k: sys.modules.get(k) for k in ('basicrta', 'basicrta.gibbs', 'MDAnalysis', 'MDAnalysis.analysis',
                                             'MDAnalysis.analysis.base')}
    try:
        for name in saved:
            sys.modules[name] = types.ModuleType(name)
        sys.modules['basicrta.gibbs'].Gibbs = ForeignGibbs
        sys.modules['MDAnalysis.analysis.base'].Results = ForeignResults
        obj = pickle.loads(raw)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    assert type(obj) is ForeignGibbs and type(obj.processed_results) is ForeignResults
    assert obj.processed_results.data == {} and np.array_equal(obj['indicator'], gb.indicator)
    assert obj.niter == 3000 and obj.residue == 'W313'


def test_deferred_pickle_equals_a_saved_object(tmp_path, monkeypatch):
    """The overlapped output path creates ``gibbs_{niter}.pkl`` before the chain has finished: the label array is
    a hole of the file that is filled in place (``indicator`` = a memory map of it), weights and rates are
    patched in at the end (gibbs.DeferredPickle).  The result must unpickle to exactly what ``save`` of the
    finished object gives -- also through the plain ``pickle.load`` a stock basicrta uses --, rotate ``.bak``
    like gibbs.py:343-344, and vanish (previous file restored) if the chain fails."""
    import io
    import pickletools
    from basicrta_b200.gibbs import DeferredPickle, dump_reference_pickle
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(3)
    for n, niter, thin in ((4000, 3000, 10), (7, 300, 100), (3, 50, 100)):
        rows = (niter + 1) // thin
        times = np.ceil(rng.exponential(1.0, n) / 0.1) * 0.1
        gb = Gibbs(times, 'D1', 0, ncomp=5, niter=niter, cutoff=7.0)
        gb.g = thin
        gb._prepare(allocate_indicator=False)
        os.makedirs(gb._savedir(), exist_ok=True)
        dp = DeferredPickle(gb, rows)
        assert gb.indicator is None and dp.shape == (rows, n) and (dp.sink is None) == (rows * n == 0)
        ind = rng.integers(0, 5, size=(rows, n)).astype(np.uint8)
        w, r = rng.random((rows, 5)), rng.random((rows, 5))
        if dp.sink is not None:                               # what the staging threads do, block by block
            fd, base = dp.sink
            flat = ind.reshape(-1)
            for a in range(0, flat.size, 4096):
                os.pwrite(fd, flat[a:a + 4096].data, base + a)
        dp.complete(w, r)
        assert gb.indicator is dp.indicator and np.array_equal(gb.mcweights, w)
        raw = open(dp.path, 'rb').read()
        ops = [(op.name, a) for op, a, _ in pickletools.genops(io.BytesIO(raw))]
        assert ('SHORT_BINUNICODE', 'basicrta.gibbs') in ops and not any('basicrta_b200' in str(a) for _, a in ops)
        back = Gibbs.load(dp.path)
        ref = io.BytesIO()
        dump_reference_pickle(gb, ref)
        plain = Gibbs.load.__func__ if hasattr(Gibbs.load, '__func__') else Gibbs.load
        for name in ('indicator', 'mcweights', 'mcrates', 'times', 't', 's', 'whypers', 'rhypers'):
            assert np.array_equal(getattr(back, name), {'indicator': ind, 'mcweights': w, 'mcrates': r}.get(name, getattr(gb, name)))
        assert back.niter == niter and back.g == thin and back.residue == 'D1'
        # same object graph as a plain save of the finished object (framing aside)
        def ops_of(b):
            return [(op.name, a if not isinstance(a, (bytes, bytearray)) else len(a))
                    for op, a, _ in pickletools.genops(io.BytesIO(b)) if op.name != 'FRAME']
        assert ops_of(raw) == ops_of(ref.getvalue())
        # .bak rotation and failure handling
        gb2 = Gibbs(times, 'D1', 0, ncomp=5, niter=niter, cutoff=7.0)
        gb2.g = thin
        gb2._prepare(allocate_indicator=False)
        dp2 = DeferredPickle(gb2, rows)
        assert os.path.exists(dp.path + '.bak')
        dp2.abandon()
        assert not os.path.exists(dp.path + '.bak') and open(dp.path, 'rb').read() == raw
        os.remove(dp.path)


def test_results_stand_in_behaves_like_the_mdanalysis_class():
    from basicrta_b200.results import Results
    r = Results()
    r.labels = np.arange(3)
    r['ncomp'] = 2
    assert r.ncomp == 2 and r['labels'][2] == 2 and set(r.keys()) == {'labels', 'ncomp'}
    with pytest.raises(AttributeError):
        r.missing
    del r.ncomp
    assert 'ncomp' not in r
    back = pickle.loads(pickle.dumps(r))
    assert type(back) is Results and np.array_equal(back.labels, r.labels)
    assert r.__getstate__() is r.data                        # pickle state = the plain dict, as MDAnalysis does


def test_postprocess_recovers_components(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    gb = _filled_gibbs()
    os.makedirs('basicrta-7.0/W313')
    gb.process_gibbs()
    pr = gb.processed_results
    assert pr.ncomp == 2 and pr.indicator.shape == (600, 2)
    assert np.allclose(pr.indicator.sum(axis=1), 1.0)
    assert 2.0 < pr.parameters[0, 1] < 4.5 and 0.03 < pr.parameters[1, 1] < 0.08     # sorted fastest first
    lo, tau, hi = gb.estimate_tau()
    assert lo < tau < hi and 12 < tau < 35                  # 1 / 0.05 = 20 ns


@pytest.mark.skipif(not have_reference(), reason='/root/reference not present (GPU box)')
def test_reference_loads_our_pickle_and_clusters_identically(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    gb = _filled_gibbs()
    os.makedirs('basicrta-7.0/W313')
    gb.save()
    ref_gibbs, _ = import_reference()
    rg = ref_gibbs.Gibbs().load('basicrta-7.0/W313/gibbs_3000.pkl')
    assert type(rg).__module__ == 'basicrta.gibbs'
    for attr in ('times', 'mcweights', 'mcrates', 'indicator', 't', 's', 'whypers', 'rhypers'):
        assert np.array_equal(getattr(rg, attr), getattr(gb, attr)), attr
    rg.cluster(n_init=2, n_components=2, random_state=0)
    postprocess.cluster(gb, n_init=2, n_components=2, random_state=0)
    assert np.array_equal(rg.processed_results.labels, gb.processed_results.labels)
    assert np.allclose(rg.processed_results.indicator, gb.processed_results.indicator, equal_nan=True)
    # and our loader reads a pickle written by the reference class
    ref_gibbs.rng = np.random.default_rng(1)
    ref_gibbs.tqdm = lambda it, **kw: it
    r2 = ref_gibbs.Gibbs(gb.times, 'A1', 0, ncomp=3, niter=50, cutoff=7.0)
    r2.g = 5
    r2.run()
    ours = Gibbs.load('basicrta-7.0/A1/gibbs_50.pkl')
    assert np.array_equal(ours.mcrates, r2.mcrates) and ours.indicator.shape == (10, 600)


def test_parallel_gibbs_surface(tmp_path):
    pg = ParallelGibbs('some/dir/contacts_7.0.pkl', nproc=4, ncomp=7, niter=1000)
    assert (pg.cutoff, pg.nproc, pg.ncomp, pg.niter) == (7.0, 4, 7, 1000)
    assert issubclass(ProcessCluster, ParallelGibbs)
    # residue slicing: column 0 = protein resid, column 3 = duration (contacts.py:227-229)
    rng = np.random.default_rng(0)
    contacts = np.zeros((1000, 4))
    contacts[:, 0] = rng.integers(10, 15, 1000)
    contacts[:, 3] = rng.integers(1, 50, 1000) * 0.1
    path = tmp_path / 'contacts_7.0.pkl'
    with open(path, 'wb') as f:
        pickle.dump(contacts, f)
    pg = ParallelGibbs(str(path), nproc=1)
    resids, names, times = pg._load(None)
    assert list(resids) == [10, 11, 12, 13, 14] and names[0] == 'X10'
    for r, t in zip(resids, times):
        assert np.array_equal(np.sort(t), np.sort(contacts[contacts[:, 0] == r][:, 3]))
    resids, names, times = pg._load(12)
    assert list(resids) == [12] and len(times) == 1


def test_compat_shim_exposes_reference_import_paths():
    import sys
    from basicrta_b200 import compat
    saved = {k: v for k, v in sys.modules.items() if k == 'basicrta' or k.startswith('basicrta.')}
    for k in saved:
        del sys.modules[k]
    try:
        compat.install(force=True)
        from basicrta.gibbs import Gibbs as G2, ParallelGibbs as P2
        from basicrta.util import run_residue
        from basicrta.cluster import ProcessCluster as C2
        assert G2 is Gibbs and P2 is ParallelGibbs and C2 is ProcessCluster and callable(run_residue)
    finally:
        for k in [k for k in sys.modules if k == 'basicrta' or k.startswith('basicrta.')]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_compact_side_car_round_trips_and_halves_the_labels(tmp_path):
    """SURVEY.md 8 f-2, second step: the opt-in compact form of a finished chain (bit-packed labels) restores
    the same object; the reference pickle stays the default."""
    from basicrta_b200.gibbs import Gibbs
    from basicrta_b200.util import label_bits, pack_labels, unpack_labels
    rng = np.random.default_rng(3)
    for K in (2, 3, 4, 15, 16, 17, 200):
        for N in (1, 7, 8, 9, 1001):
            ind = rng.integers(0, K, size=(5, N)).astype(np.uint8)
            packed = pack_labels(ind, K)
            assert packed.shape == (5, -(-N * label_bits(K) // 8))
            assert np.array_equal(unpack_labels(packed, K, N), ind)
    with pytest.raises(ValueError):
        pack_labels(np.full((1, 4), 9, np.uint8), 4)
    times = np.round(rng.exponential(2.0, size=800) / 0.1 + 1) * 0.1
    g = Gibbs(times, 'W12', 0, ncomp=15, niter=2000, cutoff=7.0)
    g._prepare(allocate_indicator=False)
    g.indicator = rng.integers(0, 15, size=(20, 800)).astype(np.uint8)
    g.mcweights, g.mcrates = rng.random((20, 15)), rng.random((20, 15))
    path = g.save_compact(str(tmp_path / 'chain.compact.npz'))
    h = Gibbs.load_compact(path)
    for name in ('indicator', 'mcweights', 'mcrates', 'times', 't', 's', 'whypers', 'rhypers'):
        assert np.array_equal(getattr(h, name), getattr(g, name)), name
    for name in ('residue', 'loc', 'ncomp', 'niter', 'g', 'burnin', 'cutoff', 'ts', '_noise_cutoff'):
        assert getattr(h, name) == getattr(g, name), name
    assert h.indicator.dtype == np.uint8
    with np.load(path) as z:
        assert z['packed'].nbytes * 2 == g.indicator.nbytes                  # 4 bits per label at K = 15
