"""The dispatcher side of the drop-in on the GPU: several cutoff files in one batch, the rerun rule, the CLI,
failure isolation, the worker shim, and inputs the reference accepts that are not on its ``ts`` grid."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

from basicrta_b200.gibbs import Gibbs, GibbsBatchError, MultiCutoffGibbs, ParallelGibbs, run_batch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _contacts(path, seed, resids=(11, 12, 15), sizes=(800, 2500, 60)):
    """A ``contacts_{cutoff}.pkl`` as contacts.ProcessContacts writes it (contacts.py:227-229): rows of
    (protein resid, lipid resid, start time, duration in ns), durations on the 0.1 ns frame grid."""
    rng = np.random.default_rng(seed)
    rows = [np.stack([np.full(n, r), rng.integers(1, 40, n), rng.random(n) * 100,
                      np.ceil(rng.exponential(2.0, n) / 0.1) * 0.1], axis=1) for r, n in zip(resids, sizes)]
    arr = np.concatenate(rows)
    rng.shuffle(arr)
    with open(path, 'wb') as f:
        pickle.dump(arr, f)
    return {r: np.sort(arr[arr[:, 0] == r][:, 3]) for r in resids}


def test_multi_cutoff_ensemble_is_one_batch_and_skip_existing(tmp_path, monkeypatch):
    """BASELINE.json config 3 in miniature: two cutoff files -> every residue of both goes to the GPU in one
    batch and lands in its own basicrta-{cutoff}/{residue}/ (gibbs.py:34-40, 183-184, 340-347)."""
    monkeypatch.chdir(tmp_path)
    truth = {5.0: _contacts('contacts_5.0.pkl', 1), 7.0: _contacts('contacts_7.0.pkl', 2)}
    out = MultiCutoffGibbs(['contacts_5.0.pkl', 'contacts_7.0.pkl'], nproc=1, ncomp=4, niter=500).run()
    assert sorted((g.cutoff, g.residue) for g in out) == [(c, f'X{r}') for c in (5.0, 7.0) for r in (11, 12, 15)]
    stamps = {}
    for g in out:
        path = f'basicrta-{g.cutoff}/{g.residue}/gibbs_500.pkl'
        back = Gibbs.load(path)
        resid = int(g.residue[1:])
        np.testing.assert_array_equal(np.sort(back.times), truth[g.cutoff][resid])
        assert back.indicator.shape == (5, len(back.times)) and back.mcrates.shape == (5, 4)
        assert back.cutoff == g.cutoff and np.all(back.mcrates > 0)
        stamps[path] = os.stat(path).st_mtime_ns
    # the reference's rerun rule (scripts/get_rerun_residues.py:22-28): finished residues are not sampled again
    again = MultiCutoffGibbs(['contacts_5.0.pkl', 'contacts_7.0.pkl'], nproc=1, ncomp=4, niter=500).run(skip_existing=True)
    assert again == []
    os.remove('basicrta-7.0/X12/gibbs_500.pkl')
    again = ParallelGibbs('contacts_7.0.pkl', nproc=1, ncomp=4, niter=500).run(skip_existing=True)
    assert [g.residue for g in again] == ['X12']
    for path, stamp in stamps.items():
        if path != 'basicrta-7.0/X12/gibbs_500.pkl':
            assert os.stat(path).st_mtime_ns == stamp and not os.path.exists(path + '.bak')


def test_command_line_interface(tmp_path):
    """``python -m basicrta.gibbs --contacts ... --nproc --niter --ncomp [--resid]`` (gibbs.py:781-795)."""
    _contacts(tmp_path / 'contacts_6.5.pkl', 3)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    res = subprocess.run([sys.executable, '-m', 'basicrta_b200.gibbs', '--contacts', 'contacts_6.5.pkl', '--nproc', '1',
                          '--niter', '300', '--ncomp', '3', '--resid', '12'], cwd=tmp_path, env=env,
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    assert os.listdir(tmp_path / 'basicrta-6.5') == ['X12']
    back = Gibbs.load(str(tmp_path / 'basicrta-6.5' / 'X12' / 'gibbs_300.pkl'))
    assert back.ncomp == 3 and back.niter == 300 and back.indicator.shape == (3, 2500)


def test_a_bad_residue_fails_alone(tmp_path, monkeypatch):
    """One unusable residue must not take the batch down (the reference's pool workers are independent,
    gibbs.py:80-88): the others are sampled and saved, the failure is reported at the end."""
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(4)
    good = [np.ceil(rng.exponential(1.0, n) / 0.1) * 0.1 for n in (500, 2000)]
    gl = [Gibbs(good[0], 'A1', 0, ncomp=3, niter=200, cutoff=4.0),
          Gibbs(np.zeros(10), 'BAD', 0, ncomp=3, niter=200, cutoff=4.0),
          Gibbs(good[1], 'A2', 0, ncomp=3, niter=200, cutoff=4.0)]
    with pytest.raises(GibbsBatchError) as ei:
        run_batch(gl, device=0, seed=1, save=True)
    assert [res for res, _ in ei.value.failures] == ['BAD']
    for name in ('A1', 'A2'):
        assert os.path.exists(f'basicrta-4.0/{name}/gibbs_200.pkl')
    assert not os.path.exists('basicrta-4.0/BAD/gibbs_200.pkl')
    assert gl[0].mcrates.shape == (2, 3) and gl[2].indicator.shape == (2, 2000)


def test_worker_shim(tmp_path, monkeypatch):
    """util.run_residue(residue, time, proc, ncomp, niter, cutoff) (util.py:475-485)."""
    from basicrta_b200.util import run_residue
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(6)
    run_residue('W313', list(np.ceil(rng.exponential(1.0, 700) / 0.1) * 0.1), 0, 3, 400, 7.0)
    back = Gibbs.load('basicrta-7.0/W313/gibbs_400.pkl')
    assert back.residue == 'W313' and back.indicator.shape == (4, 700)


def test_times_that_are_not_on_the_reference_ts_grid(tmp_path, monkeypatch):
    """(1) Continuous times, as the reference's own example data (util.simulate_hn, util.py:596-608;
    tests/test_functions.py:43-44): fixed-point grid, cheap survival function, sensible posterior.
    (2) A sparse residue whose first gap is a multiple of the frame step: {0.3, 0.7, 0.8} has ts = 0.4 in the
    reference's definition (gibbs.py:147-151) but lives on the 0.1 grid."""
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(8)
    comp = rng.choice(3, size=6000, p=[0.90, 0.09, 0.01])
    x = rng.exponential(1.0 / np.array([5.0, 0.05, 0.001])[comp])             # continuous
    g = Gibbs(x, 'C1', 0, ncomp=8, niter=3000, cutoff=7.0)
    g.seed = 2
    g.run()
    srt = np.sort(x)
    assert g.ts == (srt[1:] - srt[:-1])[(srt[1:] - srt[:-1]) != 0][0]          # the pickle keeps the reference's ts
    assert len(g.t) == len(g.s) and g.s[0] == 1.0 and abs(g.s[-1]) < 1e-12
    w, r = g.mcweights[10:], g.mcrates[10:]
    fast = r[np.arange(len(w)), w.argmax(axis=1)]
    assert 4.0 < fast.mean() < 6.0                                             # no ceiling quantisation here: ~5
    slow = np.where(w > 10 / 6000, r, np.inf).min(axis=1)
    assert 0.0004 < np.median(slow) < 0.003
    sparse = Gibbs(np.array([0.3, 0.7, 0.8, 0.8, 1.5, 2.3, 0.3]), 'S1', 0, ncomp=2, niter=300, cutoff=7.0)
    assert abs(sparse.ts - 0.4) < 1e-12
    sparse.run()
    assert sparse.indicator.shape == (3, 7) and np.all(np.isfinite(sparse.mcrates))


def test_overlapped_output_path_equals_the_plain_one(tmp_path, monkeypatch):
    """A batch large enough for the live path (>= 64 MB of labels): rows are copied out while the sweep runs and
    written into holes of the pickle files (engine.LiveStream, gibbs.DeferredPickle).  Same seed, same residue
    names => the arrays in memory, and what Gibbs.load reads back from disk, must equal the plain path's
    (labels fetched after the launch, pickle.dump)."""
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(12)
    times = []
    for n in (90000, 61000, 30000, 45000, 12000, 70001):
        comp = rng.choice(3, size=n, p=[0.8, 0.15, 0.05])
        times.append(np.ceil(rng.exponential(1.0 / np.array([4.0, 0.1, 0.002])[comp]) / 0.1) * 0.1)

    def make(cutoff):
        gl = [Gibbs(t, f'L{i}', 0, ncomp=15, niter=3000, cutoff=cutoff) for i, t in enumerate(times)]
        for g in gl:
            g.g = 10                                           # 300 saved rows -> 92 MB of labels
        return gl
    plain, live = make(1.0), make(1.0)
    run_batch(plain, device=0, seed=5, save=False, live=False)
    seen = []
    run_batch(live, device=0, seed=5, save=True, progress=lambda done, total: seen.append(done))
    assert seen[0] == 0 and seen[-1] == 3000 and seen == sorted(seen) and len(seen) > 2
    for a, b in zip(plain, live):
        np.testing.assert_array_equal(a.mcweights, b.mcweights)
        np.testing.assert_array_equal(a.mcrates, b.mcrates)
        np.testing.assert_array_equal(a.indicator, b.indicator)
        back = Gibbs.load(f'basicrta-1.0/{b.residue}/gibbs_3000.pkl')
        np.testing.assert_array_equal(back.indicator, a.indicator)
        np.testing.assert_array_equal(back.mcrates, a.mcrates)
        np.testing.assert_array_equal(back.times, a.times)
        assert back.g == 10 and back.indicator.dtype == np.uint8
