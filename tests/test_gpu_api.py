"""The public API on the GPU: Gibbs.run / _sample_indicator / ParallelGibbs.run, edge cases."""
import os
import pickle

import numpy as np
import pytest

from basicrta_b200 import _cabi
from basicrta_b200.engine import ChainInput
from basicrta_b200.gibbs import Gibbs, ParallelGibbs, run_batch
from oracle import gibbs_oracle as O

pytestmark = pytest.mark.gpu


def test_gibbs_run_writes_reference_layout(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    times = O.synth_times(4000, [0.8, 0.2], [4.0, 0.05], seed=2)
    g = Gibbs(times, 'W313', 0, ncomp=5, niter=5000, cutoff=7.0)
    g.g, g.seed = 50, 3                                      # attribute override after construction
    g.run()
    assert os.path.exists('basicrta-7.0/W313/gibbs_5000.pkl')
    back = Gibbs.load('basicrta-7.0/W313/gibbs_5000.pkl')
    assert back.mcweights.shape == (100, 5) and back.indicator.shape == (100, 4000)
    assert back.indicator.dtype == np.uint8 and back.mcrates.dtype == np.float64
    assert np.allclose(back.mcweights.sum(1), 1, atol=1e-5)
    w, r = back.mcweights[40:], back.mcrates[40:]
    assert abs(r[w > 0.5].mean() / 3.2 - 1) < 0.15           # 4 /ns on a 0.1 ns ceil grid reads ~3.2
    # same seed -> same chain; different seed -> different chain
    g2 = Gibbs(times, 'W313', 0, ncomp=5, niter=5000, cutoff=7.0)
    g2.g, g2.seed = 50, 3
    g2._prepare()
    run_batch([g2], prepared=True)
    assert np.array_equal(g2.mcrates, back.mcrates) and np.array_equal(g2.indicator, back.indicator)
    g2.seed = 4
    run_batch([g2], prepared=True)
    assert not np.array_equal(g2.mcrates, back.mcrates)


def test_hyperparameters_are_read_at_run_time():
    times = O.synth_times(2000, [1.0], [1.0], seed=1)
    g = Gibbs(times, 'X1', 0, ncomp=3, niter=2000, cutoff=1.0)
    g.seed = 1
    g._prepare()
    g.rhypers = np.ones((3, 2)) * [1.0, 3.0e6]               # a crushing prior on the rates
    run_batch([g], prepared=True)
    assert g.mcrates[10:].max() < 0.01


def test_sample_indicator_redraws_labels():
    times = O.synth_times(3000, [0.7, 0.3], [3.0, 0.05], seed=4)
    g = Gibbs(times, 'X2', 0, ncomp=4, niter=2000, cutoff=1.0)
    g.g, g.burnin, g.seed = 20, 400, 5
    g._prepare()
    run_batch([g], prepared=True)
    kept = g.indicator.copy()
    g.indicator = None
    tail = g._sample_indicator()
    assert g.indicator.shape == kept.shape and tail.shape == (80, 3000)
    # new draws from the stored post-update parameters: same label distribution, not the same labels
    f_old = np.bincount(kept[50:].ravel(), minlength=4) / kept[50:].size
    f_new = np.bincount(g.indicator[50:].ravel(), minlength=4) / kept[50:].size
    assert np.abs(np.sort(f_old) - np.sort(f_new)).max() < 0.01


def test_parallel_gibbs_runs_residues(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    rng = np.random.default_rng(0)
    rows = []
    for resid, n in ((101, 900), (102, 3001), (105, 37)):
        rows.append(np.stack([np.full(n, resid), rng.integers(1, 30, n), rng.random(n),
                              np.ceil(rng.exponential(2.0, n) / 0.1) * 0.1], axis=1))
    contacts = np.concatenate(rows)
    rng.shuffle(contacts)
    with open('contacts_7.0.pkl', 'wb') as f:
        pickle.dump(contacts, f)
    out = ParallelGibbs('contacts_7.0.pkl', nproc=1, ncomp=4, niter=1000).run()
    assert sorted(g.residue for g in out) == ['X101', 'X102', 'X105']
    for g in out:
        back = Gibbs.load(f'basicrta-7.0/{g.residue}/gibbs_1000.pkl')
        assert back.indicator.shape == (10, len(g.times)) and back.mcrates.shape == (10, 4)
        assert np.all(back.mcrates > 0) and back.indicator.max() < 4
    only = ParallelGibbs('contacts_7.0.pkl', nproc=1, ncomp=4, niter=200).run(run_resids=102)
    assert [g.residue for g in only] == ['X102']


def test_dispatch_over_all_gpus_equals_one_gpu(tmp_path, monkeypatch):
    """``dispatch`` (what ParallelGibbs.run calls) shards residues over GPUs with one host thread per GPU.
    With a fixed seed the result must not depend on the number of GPUs: chains are keyed by residue name.
    Runs on however many GPUs the box has (1: the threads path with a single shard)."""
    import torch
    from basicrta_b200.gibbs import dispatch
    monkeypatch.chdir(tmp_path)
    n_gpu = torch.cuda.device_count()
    rng = np.random.default_rng(5)
    times = [np.ceil(rng.exponential(1.5, n) / 0.1) * 0.1 for n in (4000, 1500, 9000, 700, 2500, 6000, 30, 12000)]

    def make():
        return [Gibbs(t, f'R{i}', 0, ncomp=5, niter=600, cutoff=3.5) for i, t in enumerate(times)]
    one, many = make(), make()
    dispatch(one, 1, seed=77, save=False)
    dispatch(many, n_gpu, seed=77, save=True)
    assert {g.loc for g in many} == set(range(min(n_gpu, len(times))))       # every GPU got residues
    for a, b in zip(one, many):
        np.testing.assert_array_equal(a.mcweights, b.mcweights)
        np.testing.assert_array_equal(a.mcrates, b.mcrates)
        np.testing.assert_array_equal(a.indicator, b.indicator)
        assert os.path.exists(f'basicrta-3.5/{b.residue}/gibbs_600.pkl')


@pytest.mark.parametrize('n', [1, 2, 3, 5, 127, 129, 513])
def test_ragged_sizes_exact(engine, n):
    """N not a multiple of 4 / smaller than a CTA: padding must never be counted."""
    K, niter = 3, 6
    times = O.synth_times(n, [0.6, 0.4], [2.0, 0.1], seed=n)
    ts = 0.1
    ticks = O.to_ticks(times, ts)
    ref = O.run_teacher_forced(ticks, ts, K, niter, seed=1, chain_id=n, rng=np.random.default_rng(n), g=1)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE
    got = engine.run([ChainInput(ticks=ticks, ts=ts, chain_id=n)], K, niter, thin=1, seed=1, flags=flags,
                     inject={'coef_c': [ref['coef_c']], 'coef_a': [ref['coef_a']]})[0]
    assert np.array_equal(got.trace_nk, ref['nk']) and got.trace_nk.sum(axis=1).tolist() == [n] * niter
    assert np.array_equal(got.trace_tk, ref['tk']) and np.array_equal(got.indicator, ref['indicator'])


def test_uint32_ticks_and_large_values(engine):
    """Ticks >= 65536 take the uint32 path; sums stay exact."""
    K, niter = 4, 8
    rng = np.random.default_rng(3)
    ticks = np.concatenate([rng.integers(1, 50, 3000), rng.integers(70000, 4000000, 40)]).astype(np.int64)
    ref = O.run_teacher_forced(ticks, 0.01, K, niter, seed=2, chain_id=1, rng=np.random.default_rng(1), g=2)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE
    got = engine.run([ChainInput(ticks=ticks, ts=0.01, chain_id=1)], K, niter, thin=2, seed=2, flags=flags,
                     inject={'coef_c': [ref['coef_c']], 'coef_a': [ref['coef_a']]})[0]
    assert np.array_equal(got.trace_nk, ref['nk']) and np.array_equal(got.trace_tk, ref['tk'])
    assert np.array_equal(got.indicator, ref['indicator'])


def test_stress_k30_four_decades(engine):
    """Config C5 in miniature: K = 30 (initial rates down to 5e-29, gibbs.py:186), true rates over
    4 decades.  The reference's linear-space likelihood can produce NaN here; the log-space kernel
    must finish with status 0 and find the fast and the slow end."""
    n = 20000
    times = O.synth_times(n, [0.6, 0.25, 0.1, 0.04, 0.01], [10, 1, 0.1, 0.01, 0.001], seed=5000)
    ticks = O.to_ticks(times, 0.1)
    res = engine.run([ChainInput(ticks=ticks, ts=0.1, chain_id=7)], 30, 20000, thin=100, seed=9)[0]
    assert res.status == 0 and np.isfinite(res.mcrates).all() and np.isfinite(res.mcweights).all()
    w, r = res.mcweights[100:], res.mcrates[100:]
    sig = w > 10.0 / n
    assert r[sig].max() > 2.0 and r[sig].min() < 0.003
    assert res.indicator.max() < 30


def test_large_team_atomics_path_exact(engine):
    """A chain wide enough for a team > BRTA_MAILBOX_MAX_TEAM uses the L2-atomics exchange."""
    K, niter = 6, 6
    times = O.synth_times(40000, [0.9, 0.1], [5, 0.05], seed=11)
    ticks = O.to_ticks(times, 0.1)
    ref = O.run_teacher_forced(ticks, 0.1, K, niter, seed=4, chain_id=3, rng=np.random.default_rng(2), g=3)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE
    db = engine.prepare([ChainInput(ticks=ticks, ts=0.1, chain_id=3)], K, niter, thin=3, seed=4, flags=flags,
                        inject={'coef_c': [ref['coef_c']], 'coef_a': [ref['coef_a']]})
    assert db.plan.team_size[0] > _cabi.MAILBOX_MAX_TEAM
    engine.launch(db)
    got = engine.fetch(db)[0]
    assert np.array_equal(got.trace_nk, ref['nk']) and np.array_equal(got.trace_tk, ref['tk'])
    assert np.array_equal(got.indicator, ref['indicator'])


def test_argument_validation(engine):
    with pytest.raises(ValueError):
        engine.prepare([ChainInput(ticks=np.array([1, 2, 3]), ts=0.1)], 256, 10)
    with pytest.raises(ValueError):
        engine.prepare([ChainInput(ticks=np.array([], dtype=np.int64), ts=0.1)], 3, 10)
    with pytest.raises(ValueError):
        engine.prepare([ChainInput(ticks=np.array([1 << 23]), ts=0.1)], 3, 10)
    from basicrta_b200.engine import times_to_ticks
    with pytest.raises(ValueError):
        times_to_ticks(np.array([0.1, 0.25, 0.3]), 0.1)      # not on the grid


def _two_gpus():
    import torch
    return torch.cuda.device_count() >= 2


@pytest.mark.skipif(not _two_gpus(), reason='needs 2 GPUs (gpurun --gpus 2)')
def test_chain_sharded_over_two_gpus_is_bit_exact():
    """Config C4 in miniature: one chain, its times split over 2 GPUs, per-iteration exchange of the
    integer (n_k, sum tick_k) over peer memory inside the kernels.  Teacher-forced EXACT mode must match
    the oracle bit for bit, and the free-running chain must equal the single-GPU chain."""
    from basicrta_b200.engine import get_engine, run_sharded
    K, niter, thin = 15, 12, 4
    times = O.synth_times(60001, [0.9, 0.09, 0.01], [5, 0.05, 0.001], seed=4)
    ticks = O.to_ticks(times, 0.1)
    ref = O.run_teacher_forced(ticks, 0.1, K, niter, seed=6, chain_id=2, rng=np.random.default_rng(3), g=thin)
    flags = _cabi.FLAG_EXACT | _cabi.FLAG_INJECT_COEF | _cabi.FLAG_TRACE
    chain = ChainInput(ticks=ticks, ts=0.1, chain_id=2)
    got = run_sharded(chain, K, niter, devices=[0, 1], thin=thin, seed=6, flags=flags,
                      inject={'coef_c': [ref['coef_c']], 'coef_a': [ref['coef_a']]})
    assert got.status == 0
    assert np.array_equal(got.trace_nk, ref['nk']) and np.array_equal(got.trace_tk, ref['tk'])
    assert np.array_equal(got.indicator, ref['indicator'])
    # free running: 2 GPUs == 1 GPU, bit for bit
    one = get_engine(0).run([chain], K, 300, thin=50, seed=9)[0]
    two = run_sharded(chain, K, 300, devices=[0, 1], thin=50, seed=9)
    assert two.status == 0
    assert np.array_equal(one.mcrates, two.mcrates) and np.array_equal(one.mcweights, two.mcweights)
    assert np.array_equal(one.indicator, two.indicator)


def _sharded_rank(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from basicrta_b200.engine import run_sharded_dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    ticks = O.to_ticks(O.synth_times(80003, [0.9, 0.09, 0.01], [5, 0.05, 0.001], seed=4), 0.1)
    res, (a, b), order = run_sharded_dist(ChainInput(ticks=ticks, ts=0.1, chain_id=2), 15, 300, thin=50, seed=9, device=rank)
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), w=res.mcweights, r=res.mcrates, ind=res.indicator,
             span=np.array([a, b]), order=order, status=res.status)
    dist.destroy_process_group()


@pytest.mark.skipif(not _two_gpus(), reason='needs 2 GPUs (gpurun --gpus 2)')
def test_chain_sharded_over_two_processes_is_bit_exact(tmp_path):
    """The same with ONE PROCESS PER GPU (how torchrun launches bench.py): mailboxes exchanged as CUDA IPC
    handles by one all_gather, then the ranks only meet inside the kernels.  Every rank's shard of the labels
    and its copy of the weights / rates must equal the single-GPU run."""
    import socket
    import torch.multiprocessing as mp
    from basicrta_b200.engine import get_engine
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    mp.spawn(_sharded_rank, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ticks = O.to_ticks(O.synth_times(80003, [0.9, 0.09, 0.01], [5, 0.05, 0.001], seed=4), 0.1)
    one = get_engine(0).run([ChainInput(ticks=ticks, ts=0.1, chain_id=2)], 15, 300, thin=50, seed=9)[0]
    covered = 0
    for rank in range(2):
        z = np.load(tmp_path / f'rank{rank}.npz')
        assert int(z['status']) == 0
        a, b = z['span']
        np.testing.assert_array_equal(z['w'], one.mcweights)
        np.testing.assert_array_equal(z['r'], one.mcrates)
        np.testing.assert_array_equal(z['ind'], one.indicator[:, z['order']][:, a:b])
        covered += b - a
    assert covered == len(ticks)


def test_full_size_batch_invariants(engine):
    """BASELINE.json's sizes are out of the oracle's reach, so the full-size checks are properties that do
    not depend on size: every iteration's component counts add up to N and its tick sums to the chain's
    total (a checksum of checksums over the integer statistics); the stored label rows reproduce the
    traced counts; the launch is deterministic; and neither the batch a chain runs in nor the schedule
    (waves, measured slicing) changes one bit of its output.  Sizes: the C2 range (1e4 .. 1e5 data),
    K = 15, free-running FAST mode."""
    import bench
    ids = [22, 385, 74, 329, 231, 5]                       # N from 1.0e4 to 1.0e5
    ticks = bench.workload(ids)
    assert max(len(t) for t in ticks) > 90000 and min(len(t) for t in ticks) < 16000
    chains = [ChainInput(ticks=t, ts=0.1, chain_id=i) for i, t in zip(ids, ticks)]
    K, niter, thin = 15, 400, 100
    first = engine.run(chains, K, niter, thin=thin, seed=77, flags=_cabi.FLAG_TRACE)
    for ch, res in zip(chains, first):
        n = len(ch.ticks)
        assert res.status == 0
        assert np.array_equal(res.trace_nk.sum(axis=1), np.full(niter, n))
        assert np.array_equal(res.trace_tk.sum(axis=1), np.full(niter, int(np.sum(ch.ticks))))
        for row in range(niter // thin):
            hist = np.bincount(res.indicator[row], minlength=K)
            assert np.array_equal(hist, res.trace_nk[(row + 1) * thin - 1])
            tsum = np.bincount(res.indicator[row], weights=ch.ticks, minlength=K)
            assert np.array_equal(tsum.astype(np.int64), res.trace_tk[(row + 1) * thin - 1])
        assert np.allclose(res.mcweights.sum(axis=1), 1.0, atol=1e-5) and (res.mcrates > 0).all()
    again = engine.run(chains, K, niter, thin=thin, seed=77, flags=_cabi.FLAG_TRACE)
    waves3 = engine.run(chains, K, niter, thin=thin, seed=77, n_waves=3)
    alone = engine.run(chains[1:2], K, niter, thin=thin, seed=77)
    db = engine.prepare(chains, K, niter, thin=thin, seed=77, calibrate=True)
    engine.launch(db)
    tuned = engine.fetch(db)
    for a, b, c, d in zip(first, again, waves3, tuned):
        for other in (b, c, d):
            assert np.array_equal(a.mcweights, other.mcweights) and np.array_equal(a.mcrates, other.mcrates)
            assert np.array_equal(a.indicator, other.indicator)
        assert np.array_equal(a.trace_nk, b.trace_nk)
    assert np.array_equal(first[1].indicator, alone[0].indicator) and np.array_equal(first[1].mcrates, alone[0].mcrates)


def test_many_chains_per_launch(engine):
    """Config C3 in miniature: far more chains than CTAs (several waves, every CTA walks a list of tasks).
    Each chain's output equals what it produces when launched on its own."""
    rng = np.random.default_rng(31)
    chains = []
    for r in range(1500):
        n = int(rng.integers(40, 2500))
        comp = rng.choice(2, size=n, p=[0.85, 0.15])
        x = rng.exponential(1.0 / np.array([3.0, 0.05])[comp])
        chains.append(ChainInput(ticks=np.maximum(np.ceil(x / 0.1), 1).astype(np.int64), ts=0.1, chain_id=r))
    K, niter, thin = 5, 120, 40
    db = engine.prepare(chains, K, niter, thin=thin, seed=3)
    assert db.plan.n_waves >= 2 and len(db.plan.tasks) >= 1500
    engine.launch(db)
    res = engine.fetch(db)
    assert all(x.status == 0 for x in res)
    for r in (0, 1, 613, 1499):
        solo = engine.run([chains[r]], K, niter, thin=thin, seed=3)[0]
        assert np.array_equal(solo.indicator, res[r].indicator)
        assert np.array_equal(solo.mcweights, res[r].mcweights) and np.array_equal(solo.mcrates, res[r].mcrates)


def test_live_stream_brings_rows_home_during_the_run(engine):
    """Output path overlapped with the sweep: the kernel publishes per chain how many saved
    rows are complete (brta_batch.progress, mapped pinned memory); a poller copies finished row blocks out
    while the launch is still running.  The result must equal the plain fetch, for multi-CTA teams, ragged
    sizes and a row count that is not a multiple of the publication interval."""
    from concurrent.futures import ThreadPoolExecutor
    rng = np.random.default_rng(17)
    chains = []
    for r, n in enumerate((52000, 9001, 700, 33, 18000)):
        comp = rng.choice(3, size=n, p=[0.8, 0.15, 0.05])
        x = rng.exponential(1.0 / np.array([4.0, 0.1, 0.002])[comp])
        chains.append(ChainInput(ticks=np.maximum(np.ceil(x / 0.1), 1).astype(np.int64), ts=0.1, chain_id=60 + r))
    K, niter, thin = 15, 3050, 10                                           # 305 rows, published every 16
    ref = engine.run(chains, K, niter, thin=thin, seed=4)
    db = engine.prepare(chains, K, niter, thin=thin, seed=4, progress_rows=16)
    assert db.progress is not None
    got = {}
    reports = []
    with ThreadPoolExecutor(max_workers=4) as pool:
        live = engine.start_live_stream(db, lambda r, res: got.__setitem__(r, res), pool,
                                        progress=lambda done, total: reports.append((done, total)))
        live.MIN_FLUSH_BYTES = 1 << 16                                      # small batch: stream anyway
        engine.launch(db)
        for fut in live.finish():
            fut.result()
    assert sorted(got) == list(range(len(chains)))
    assert int(db.progress.numpy().min()) == 304                            # 19 publications of 16 rows
    for r, a in enumerate(ref):
        b = got[r]
        assert a.status == 0 and b.status == 0
        np.testing.assert_array_equal(a.mcweights, b.mcweights)
        np.testing.assert_array_equal(a.mcrates, b.mcrates)
        np.testing.assert_array_equal(a.indicator, b.indicator)
    assert reports and all(t == niter for _, t in reports)
