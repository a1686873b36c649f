"""Generate golden vectors by running the UNMODIFIED reference sampler.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py            # everything (several minutes on 8 cores)
    python tests/golden/make_golden.py small      # bit-exact replay vectors only

Outputs (committed):

``ref_replay.npz``
    Seeded runs of the reference ``basicrta.gibbs.Gibbs.run()`` (gibbs.py:176-219) with
    the module generator replaced by ``default_rng(seed)`` after import (the loop looks
    ``rng`` up at call time; the reference itself never seeds, gibbs.py:17).  For each
    case: ``times``, the seed and the full ``mcweights`` / ``mcrates`` / ``indicator``
    plus ``ts``, ``t``, ``s``, ``whypers``, ``rhypers`` and the pickle's attribute names.
    ``oracle.gibbs_oracle.run_reference_order`` must reproduce these bit for bit.

``ref_c1_posterior.npz``
    Config 1 of BASELINE.json (N=5000 three-exponential data, K=15) run with the
    reference for niter=110000 under 4 seeds: the thinned ``mcweights`` / ``mcrates``
    samples (float32).  This is the statistical truth the CUDA sampler's free-running
    chains are compared with (2 % on posterior means, KS p > 0.01).

``ref_c2_posterior.npz`` / ``ref_c5_posterior.npz``
    The same for residues of the headline configuration C2 (``bench.residue_times(r)`` for the
    residues in ``C2_RESIDUES``: N ~ 1e4, K = 15, 2 / 3 / 4 true components) and of the stress
    configuration C5 (``bench.c5_residue_times(r)``: N = 2e4, K = 30, five components over four
    decades), 110 000 iterations of the unmodified reference under two seeds each, plus the END of
    the reference's pipeline per run: ``Gibbs.cluster(n_init=117, n_components=lmode)``
    (gibbs.py:221-273, unmodified), the label sort of ``util.mixture_and_plot``
    (util.py:738-756, restated here because the function is ~450 lines of matplotlib),
    ``Gibbs._estimate_params`` and ``Gibbs.estimate_tau`` (gibbs.py:667-715, unmodified):
    cluster count, [CI low, tau, CI high], the slowest cluster's tau samples summarised.
    A reference run that dies with NumPy's ``ValueError`` (0/0 responsibilities,
    gibbs.py:196-200) is recorded as ``failed`` with the message.

``ref_c2_functionals.npz`` / ``ref_c5_functionals.npz``
    The same residues under EIGHT seeds (the two above plus six more), stored compactly: per run the
    label-invariant functionals of every stored sample (``oracle.gibbs_oracle.posterior_functionals``:
    mixture survival on a data-quantile grid, mean rate, mean time, slowest significant rate, ...)
    and the end-of-pipeline summary (cluster count, tau, CI).  The reference's chains mix slowly
    between configurations with a different number of live components, so two chains are not
    enough to tell the sampler's distribution from chain-to-chain variation; eight are.
"""
import os
import sys
import tempfile
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from _refstubs import import_reference  # noqa: E402
from oracle import gibbs_oracle as O    # noqa: E402

C1 = dict(n=5000, weights=[0.90, 0.09, 0.01], rates=[5, 0.05, 0.001], seed=20241109)

REPLAY_CASES = [
    # name, n, true weights, true rates, data seed, ncomp, niter, g, rng seed
    ('k4_n300', 300, [0.9, 0.09, 0.01], [5, 0.05, 0.001], 1, 4, 40, 5, 7),
    ('k15_n500', 500, [0.9, 0.09, 0.01], [5, 0.05, 0.001], 2, 15, 30, 10, 11),
    ('k2_n64', 64, [0.5, 0.5], [2.0, 0.1], 3, 2, 25, 1, 13),
    ('k7_n1000_g100', 1000, [0.6, 0.3, 0.1], [3.0, 0.3, 0.02], 4, 7, 200, 100, 17),
]


def _run_reference(times, ncomp, niter, g, rng_seed, residue='X1'):
    gibbs, _ = import_reference()
    gibbs.rng = np.random.default_rng(rng_seed)
    gibbs.tqdm = lambda it, **kw: it                      # silence the progress bar only
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            G = gibbs.Gibbs(times, residue, 0, ncomp=ncomp, niter=niter, cutoff=7.0)
            G.g = g
            G.run()
            assert os.path.exists(f'basicrta-7.0/{residue}/gibbs_{niter}.pkl')
        finally:
            os.chdir(cwd)
    return G


def make_replay():
    out = {}
    for name, n, w, r, dseed, ncomp, niter, g, rseed in REPLAY_CASES:
        times = O.synth_times(n, w, r, dseed)
        G = _run_reference(times, ncomp, niter, g, rseed)
        out[f'{name}/times'] = times
        out[f'{name}/meta'] = np.array([ncomp, niter, g, rseed], dtype=np.int64)
        for attr in ('mcweights', 'mcrates', 'indicator', 't', 's', 'whypers', 'rhypers'):
            out[f'{name}/{attr}'] = getattr(G, attr)
        out[f'{name}/ts'] = np.float64(G.ts)
        out[f'{name}/attrs'] = np.array(sorted(G.__dict__.keys()))
    np.savez_compressed(os.path.join(HERE, 'ref_replay.npz'), **out)
    print('wrote ref_replay.npz', len(out), 'arrays')


def _c1_worker(rng_seed):
    times = O.synth_times(C1['n'], C1['weights'], C1['rates'], C1['seed'])
    G = _run_reference(times, 15, 110000, 100, rng_seed, residue=f'X{rng_seed}')
    return G.mcweights.astype(np.float32), G.mcrates.astype(np.float32)


def make_c1():
    seeds = [101, 202, 303, 404]
    with Pool(len(seeds)) as p:
        res = p.map(_c1_worker, seeds)
    np.savez_compressed(os.path.join(HERE, 'ref_c1_posterior.npz'),
                        seeds=np.array(seeds),
                        mcweights=np.stack([r[0] for r in res]),
                        mcrates=np.stack([r[1] for r in res]))
    print('wrote ref_c1_posterior.npz')


C2_RESIDUES = [22, 4, 249, 329]          # the smallest N with 3 / 2 / 4 / 3 true components
C5_RESIDUES = [0, 1]
POSTERIOR_SEEDS = [11, 22]


def _label_sort(G):
    """util.py:738-756 without the figures: clusters ordered by mean rate, fastest first, noise
    clusters (no datum assigned with probability >= _noise_cutoff) last."""
    burn = G.burnin // G.g
    weights, rates = G.mcweights[burn:], G.mcrates[burn:]
    arates = rates[np.where(weights > 10 / len(G.times))]
    all_labels = G.processed_results.labels
    uniq = np.unique(all_labels)
    imaxs = G.processed_results.indicator.max(axis=0)
    noise = np.where(imaxs < G._noise_cutoff)[0]
    means = np.array([arates[all_labels == i].mean() for i in uniq])
    vsorts = means[np.delete(uniq, noise)].argsort()[::-1]
    nsorts = means[noise].argsort()[::-1]
    presorts = np.concatenate([np.delete(uniq, noise)[vsorts], noise[nsorts]]).astype(int)
    sorts = np.array([np.where(presorts == i)[0][0] for i in uniq])
    return sorts[all_labels], presorts


def _reference_pipeline_end(G):
    """process_gibbs (gibbs.py:275-308) with mixture_and_plot replaced by its label sort."""
    from scipy import stats
    wcutoff = 10 / len(G.times)
    burn = G.burnin // G.g
    inds = np.where(G.mcweights[burn:] > wcutoff)
    weights, rates = G.mcweights[burn:], G.mcrates[burn:]
    lens = [len(row[row > wcutoff]) for row in weights]
    lmode = stats.mode(lens).mode
    G.cluster(n_init=117, n_components=lmode)
    labels, presorts = _label_sort(G)
    pr = G.processed_results
    pr.labels = labels
    pr.indicator = pr.indicator[:, presorts]
    pr.weights, pr.rates, pr.ncomp = weights[inds], rates[inds], int(lmode)
    G._estimate_params()
    tau = G.estimate_tau()
    imaxs = pr.indicator.max(axis=0)
    noise = np.where(imaxs < G._noise_cutoff)[0]
    valid = np.delete(np.unique(pr.labels), noise)
    index = pr.parameters[valid, 1].argmin()
    taus = 1 / pr.rates[pr.labels == index]
    return dict(ncomp=int(lmode), n_noise=int(len(noise)), tau=np.array(tau, dtype=np.float64),
                parameters=pr.parameters, tau_median=float(np.median(taus)), tau_mean=float(taus.mean()),
                tau_binwidth=float((taus.max() - taus.min()) / 15), n_tau=int(len(taus)))


def _posterior_worker(job):
    kind, r, rng_seed = job
    import bench
    if kind == 'c2':
        ticks, ncomp = bench.residue_times(r), 15
    else:
        ticks, ncomp = bench.c5_residue_times(r), 30
    times = ticks * bench.TS
    try:
        G = _run_reference(times, ncomp, 110000, 100, rng_seed, residue=f'{kind}r{r}s{rng_seed}')
    except ValueError as e:                               # 0/0 responsibilities -> NumPy raises
        return job, dict(failed=str(e))
    end = _reference_pipeline_end(G)
    end.update(mcweights=G.mcweights.astype(np.float32), mcrates=G.mcrates.astype(np.float32), n=len(times))
    return job, end


def make_posteriors(kind, residues, nproc=6):
    jobs = [(kind, r, s) for r in residues for s in POSTERIOR_SEEDS]
    with Pool(min(nproc, len(jobs))) as p:
        res = p.map(_posterior_worker, jobs, chunksize=1)
    out = {'residues': np.array(residues), 'seeds': np.array(POSTERIOR_SEEDS)}
    for (_, r, s), d in res:
        for k, v in d.items():
            out[f'r{r}/s{s}/{k}'] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, f'ref_{kind}_posterior.npz'), **out)
    print(f'wrote ref_{kind}_posterior.npz', len(out), 'arrays')


EXTRA_SEEDS = [33, 44, 55, 66, 77, 88]


def _functionals_worker(job):
    kind, r, rng_seed = job
    import bench
    ticks, ncomp = (bench.residue_times(r), 15) if kind == 'c2' else (bench.c5_residue_times(r), 30)
    times = ticks * bench.TS
    try:
        G = _run_reference(times, ncomp, 110000, 100, rng_seed, residue=f'{kind}r{r}s{rng_seed}')
    except ValueError as e:
        return job, dict(failed=str(e))
    end = _reference_pipeline_end(G)
    end.update(functionals=O.posterior_functionals(G.mcweights, G.mcrates, times).astype(np.float32), n=len(times))
    return job, end


def make_functionals(kind, residues, nproc=7):
    """Six more seeds per residue, plus the functionals of the two raw runs already stored."""
    import bench
    raw = np.load(os.path.join(HERE, f'ref_{kind}_posterior.npz'))
    jobs = [(kind, r, s) for r in residues for s in EXTRA_SEEDS]
    with Pool(min(nproc, len(jobs))) as p:
        res = p.map(_functionals_worker, jobs, chunksize=1)
    out = {'residues': np.array(residues), 'seeds': np.array(POSTERIOR_SEEDS + EXTRA_SEEDS),
           'names': np.array(O.FUNCTIONAL_NAMES)}
    for r in residues:
        ticks = bench.residue_times(r) if kind == 'c2' else bench.c5_residue_times(r)
        times = ticks * bench.TS
        for s in POSTERIOR_SEEDS:
            pre = f'r{r}/s{s}/'
            if pre + 'failed' in raw.files:
                out[pre + 'failed'] = raw[pre + 'failed']
                continue
            out[pre + 'functionals'] = O.posterior_functionals(raw[pre + 'mcweights'], raw[pre + 'mcrates'],
                                                               times).astype(np.float32)
            for k in ('ncomp', 'n_noise', 'tau', 'tau_median', 'tau_mean', 'tau_binwidth', 'n_tau', 'n'):
                out[pre + k] = raw[pre + k]
    for (_, r, s), d in res:
        for k, v in d.items():
            if k != 'parameters':
                out[f'r{r}/s{s}/{k}'] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, f'ref_{kind}_functionals.npz'), **out)
    print(f'wrote ref_{kind}_functionals.npz', len(out), 'arrays')


if __name__ == '__main__':
    what = sys.argv[1] if len(sys.argv) > 1 else 'all'
    if what in ('all', 'small'):
        make_replay()
    if what in ('all', 'c1'):
        make_c1()
    if what in ('all', 'c2'):
        make_posteriors('c2', C2_RESIDUES)
    if what in ('all', 'c5'):
        make_posteriors('c5', C5_RESIDUES)
    if what in ('all', 'c2f'):
        make_functionals('c2', C2_RESIDUES)
    if what in ('all', 'c5f'):
        make_functionals('c5', C5_RESIDUES)
