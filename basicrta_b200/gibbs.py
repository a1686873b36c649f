"""Drop-in ``Gibbs`` / ``ParallelGibbs`` with the sweep on a B200.

Mirrors the public surface of the reference's ``basicrta/gibbs.py`` for the hot path:

* ``Gibbs(times, residue, loc, ncomp, niter, cutoff)`` with the same attributes
  (gibbs.py:133-157), ``run()`` filling ``mcweights`` / ``mcrates`` / ``indicator`` thinned
  every ``g`` steps and pickling the instance to ``basicrta-{cutoff}/{residue}/
  gibbs_{niter}.pkl`` (gibbs.py:176-219, 336-349), ``load`` (gibbs.py:351-381),
  ``__getitem__`` (gibbs.py:159-160), ``_sample_indicator`` (gibbs.py:321-334);
* ``ParallelGibbs(contacts, nproc, ncomp, niter).run(run_resids)`` (gibbs.py:20-88):
  residues are sharded over ``nproc`` GPUs instead of a multiprocessing pool.

``g``, ``burnin``, ``whypers``, ``rhypers`` are read at ``run()`` time, so the reference's
"assign after construction" idiom (tests/test_functions.py:11-12) keeps working.

The sweep itself (gibbs.py:191-217) is ``brta_gibbs_run_batch`` of the C ABI; there is no
CPU implementation in this package.  Posterior processing (``cluster``, ``process_gibbs``,
``estimate_tau``) stays on the host, plot-free, in :mod:`basicrta_b200.postprocess`.
"""
import os
import pickle
import threading
import zlib

import numpy as np

from . import _cabi
from .engine import ChainInput, coefficients, get_engine, times_to_ticks
from .plan import shard_chains
from .util import get_s

try:                                                   # same class as the reference when available
    from MDAnalysis.analysis.base import Results       # gibbs.py:9
except Exception:                                      # MDAnalysis is optional here
    from .results import Results


def _fresh_seed():
    return int(np.random.SeedSequence().entropy) & 0xFFFFFFFFFFFFFFFF


def _chain_id(residue, cutoff):
    return zlib.crc32(f'{residue}|{cutoff}'.encode()) & 0xFFFFFFFF


class Gibbs(object):
    """Gibbs sampler for an exponential mixture of residence times (one residue).

    Same constructor and attributes as ``basicrta.gibbs.Gibbs`` (gibbs.py:133-157).  Two
    attributes are new and optional: ``seed`` (None = fresh entropy per run; the reference
    never seeds, gibbs.py:17) and ``device`` (CUDA ordinal; default ``loc % device_count``).
    """

    def __init__(self, times=None, residue=None, loc=0, ncomp=15, niter=110000, cutoff=None):
        self.times = times
        self.residue = residue
        self.niter = niter
        self.loc = loc
        self.ncomp = ncomp
        self.g = 100
        self.burnin = 10000
        self.cutoff = cutoff
        self.processed_results = Results()
        self._noise_cutoff = 0.4

        if times is not None:
            srt = np.sort(times)
            gaps = srt[1:] - srt[:-1]
            nonzero = gaps[gaps != 0]
            self.ts = nonzero[0] if len(nonzero) else times.min()
        else:
            self.ts = None

        self.keys = {'times', 'residue', 'loc', 'ncomp', 'niter', 'g', 'burnin',
                     'processed_results', 'ts', 'mcweights', 'mcrates', 't',
                     's', 'cutoff', 'indicator'}

    def __getitem__(self, item):
        return getattr(self, item)

    # ---- host preparation (gibbs.py:162-174) ---------------------------------------------
    def _prepare(self):
        self.t, self.s = get_s(self.times, self.ts)
        rows = (self.niter + 1) // self.g
        self.indicator = np.zeros((rows, self.times.shape[0]), dtype=np.uint8)
        self.mcweights = np.zeros((rows, self.ncomp))
        self.mcrates = np.zeros((rows, self.ncomp))
        self.whypers = np.ones(self.ncomp) / [self.ncomp]
        self.rhypers = np.ones((self.ncomp, 2)) * [1, 3]

    def _chain_input(self):
        times = np.asarray(self.times, dtype=np.float64)
        return ChainInput(ticks=times_to_ticks(times, self.ts), ts=float(self.ts),
                          chain_id=_chain_id(self.residue, self.cutoff),
                          whypers=np.asarray(self.whypers, dtype=np.float64),
                          rhypers=np.asarray(self.rhypers, dtype=np.float64))

    def _savedir(self):
        return f'basicrta-{self.cutoff}/{self.residue}'

    # ---- the sampler (gibbs.py:176-219) ---------------------------------------------------
    def run(self):
        """Run the sampler on the GPU and pickle the instance, like the reference's ``run``."""
        self._prepare()
        os.makedirs(self._savedir(), exist_ok=True)
        run_batch([self], device=getattr(self, 'device', None), prepared=True)
        self.save()

    def _sample_indicator(self):
        """Re-draw the labels of every stored (mcweights, mcrates) row without a parameter
        update (gibbs.py:321-334): the sweep kernel in teacher-forced mode, thin = 1."""
        rows = self.mcweights.shape[0]
        coef = [coefficients(w, r, self.ts) for w, r in zip(self.mcweights, self.mcrates)]
        eng = get_engine(_pick_device(getattr(self, 'device', None), self.loc))
        chain = ChainInput(ticks=times_to_ticks(self.times, self.ts), ts=float(self.ts),
                           chain_id=_chain_id(self.residue, self.cutoff))
        seed = getattr(self, 'seed', None)
        res = eng.run([chain], self.ncomp, rows, thin=1, seed=_fresh_seed() if seed is None else seed,
                      flags=_cabi.FLAG_INJECT_COEF,
                      inject={'coef_c': [np.stack([c for c, _ in coef])],
                              'coef_a': [np.stack([a for _, a in coef])]})[0]
        _raise_on_status(self.residue, res.status)
        self.indicator = np.ascontiguousarray(res.indicator[:rows])   # (niter+1)//1 rows are allocated, `rows` filled
        return self.indicator[self.burnin // self.g:]

    # ---- persistence (gibbs.py:336-381) ---------------------------------------------------
    def save(self):
        savedir = self._savedir() + '/'
        filename = f'gibbs_{self.niter}.pkl'
        if not os.path.exists(savedir):
            raise OSError(f'No such directory: {savedir}')
        if os.path.exists(savedir + filename):
            os.rename(savedir + filename, savedir + filename + '.bak')
        with open(savedir + filename, 'w+b') as f:
            pickle.dump(self, f)

    @staticmethod
    def load(file):
        keys = ['times', 'residue', 'loc', 'ncomp', 'niter', 'g', 'burnin',
                'processed_results', 'ts', 'mcweights', 'mcrates', 't',
                's', 'cutoff', 'indicator', 'whypers', 'rhypers']
        with open(file, 'r+b') as f:
            r = pickle.load(f)
        g = Gibbs()
        for attr in keys:
            try:
                setattr(g, attr, r[f'{attr}'])
            except AttributeError:
                setattr(g, attr, None)
        if isinstance(g.residue, np.ndarray):
            g.residue = g.residue[0]
        if g.t is None:
            g.t, g.s = get_s(g.times, g.ts)
        return g

    # ---- posterior processing: host side, see postprocess.py -------------------------------
    def cluster(self, method='GaussianMixture', **kwargs):
        from . import postprocess
        return postprocess.cluster(self, method=method, **kwargs)

    def process_gibbs(self, save=True, device=None):
        from . import postprocess
        return postprocess.process_gibbs(self, save=save, device=device)

    def estimate_tau(self):
        from . import postprocess
        return postprocess.estimate_tau(self)


def _raise_on_status(residue, status):
    if status & _cabi.STATUS_TIMEOUT:
        raise _cabi.BrtaError(f'residue {residue}: team rendezvous timed out on the device (status {status})')
    if status != _cabi.STATUS_OK:
        raise FloatingPointError(f'residue {residue}: sampler saw a non-finite likelihood (status {status})')


def _device_count():
    import torch
    if not torch.cuda.is_available():
        raise _cabi.BrtaError('basicrta_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.cuda.device_count()


def _pick_device(device, loc=0):
    return int(device) if device is not None else int(loc) % _device_count()


def run_batch(gibbs_list, device=None, seed=None, prepared=False, engine=None):
    """Run many residues' chains in ONE launch on one GPU and fill their output arrays.

    All members must share ``ncomp``, ``niter`` and ``g`` (``ParallelGibbs`` guarantees it,
    gibbs.py:73-75).  Hyper-parameters and ``g`` are read here, i.e. at run time.
    """
    if not gibbs_list:
        return
    first = gibbs_list[0]
    key = (first.ncomp, first.niter, first.g)
    for gb in gibbs_list:
        if (gb.ncomp, gb.niter, gb.g) != key:
            raise ValueError('a batch needs one (ncomp, niter, g)')
        if not prepared:
            gb._prepare()
    eng = engine if engine is not None else get_engine(_pick_device(device, first.loc))
    if seed is None:
        seed = getattr(first, 'seed', None)
    if seed is None:
        seed = _fresh_seed()
    res = eng.run([gb._chain_input() for gb in gibbs_list], first.ncomp, first.niter,
                  thin=first.g, seed=seed)
    for gb, r in zip(gibbs_list, res):
        _raise_on_status(gb.residue, r.status)
        gb.mcweights, gb.mcrates, gb.indicator = r.mcweights, r.mcrates, r.indicator


class ParallelGibbs(object):
    """Run a Gibbs sampler for every residue of a contact map (gibbs.py:20-88).

    ``nproc`` is the number of GPUs to shard the residues over (the reference's number of
    pool processes); whole residues go to GPUs by longest-processing-time-first on N_r --
    chains are independent, so there is no collective.
    """

    def __init__(self, contacts, nproc=1, ncomp=15, niter=110000):
        self.cutoff = float(contacts.strip('.pkl').split('/')[-1].split('_')[-1])
        self.niter = niter
        self.nproc = nproc
        self.ncomp = ncomp
        self.contacts = contacts

    @staticmethod
    def _residue_names(contacts, resids):
        """``W313``-style names from the AtomGroup in the dtype metadata (gibbs.py:62-67);
        plain ``X{resid}`` when the metadata (or MDAnalysis) is unavailable."""
        meta = getattr(contacts.dtype, 'metadata', None) or {}
        try:
            import MDAnalysis as mda
            rg = meta['ag1'].residues
            letters = {int(i): mda.lib.util.convert_aa_code(n) for i, n in zip(rg.resids, rg.resnames)}
            return [f'{letters[int(r)]}{int(r)}' for r in resids]
        except Exception:
            names = meta.get('residue_names') if isinstance(meta, dict) else None
            if names is not None:
                return [str(names[int(r)]) for r in resids]
            return [f'X{int(r)}' for r in resids]

    def _load(self, run_resids):
        with open(self.contacts, 'r+b') as f:
            contacts = pickle.load(f)
        protids = np.unique(contacts[:, 0])
        if run_resids is None or (np.ndim(run_resids) == 0 and not run_resids) or \
                (np.ndim(run_resids) > 0 and len(run_resids) == 0):
            run_resids = protids
        if not isinstance(run_resids, (list, np.ndarray)):
            run_resids = [run_resids]
        run_resids = np.asarray(run_resids)
        # one stable sort groups all residues' durations (column 3) instead of one boolean
        # mask per residue (gibbs.py:68-69 is O(R*M))
        order = np.argsort(contacts[:, 0], kind='stable')
        col0 = np.asarray(contacts[:, 0])[order]
        dur = np.asarray(contacts[:, 3], dtype=np.float64)[order]
        lo = np.searchsorted(col0, run_resids, side='left')
        hi = np.searchsorted(col0, run_resids, side='right')
        times = [dur[a:b].copy() for a, b in zip(lo, hi)]
        return run_resids, self._residue_names(contacts, run_resids), times

    def run(self, run_resids=None):
        resids, names, times = self._load(run_resids)
        n_gpu = max(1, min(int(self.nproc), _device_count()))
        gibbs = [Gibbs(t, name, i % n_gpu, ncomp=self.ncomp, niter=self.niter, cutoff=self.cutoff)
                 for i, (name, t) in enumerate(zip(names, times)) if len(t) > 0]
        dispatch(gibbs, n_gpu)
        return gibbs


def dispatch(gibbs_list, n_gpu, seed=None, save=True):
    """Shard residues over GPUs (one host thread per GPU, one launch per GPU) and write the
    reference's per-residue pickles."""
    if not gibbs_list:
        return
    shards = shard_chains([len(g.times) for g in gibbs_list], n_gpu)
    errors = []
    seed = _fresh_seed() if seed is None else seed

    def work(dev, idx):
        try:
            members = [gibbs_list[i] for i in idx]
            for gb in members:
                gb.loc = dev
            run_batch(members, device=dev, seed=seed)
            if save:
                for gb in members:
                    os.makedirs(gb._savedir(), exist_ok=True)
                    gb.save()
        except BaseException as e:                     # surfaced to the caller below
            errors.append(e)

    threads = [threading.Thread(target=work, args=(dev, idx)) for dev, idx in enumerate(shards) if len(idx)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]


if __name__ == '__main__':                             # gibbs.py:781-795
    import argparse
    parser = argparse.ArgumentParser()
    parser.add_argument('--contacts')
    parser.add_argument('--resid', type=int, default=None)
    parser.add_argument('--nproc', type=int, default=1)
    parser.add_argument('--niter', type=int, default=110000)
    parser.add_argument('--ncomp', type=int, default=15)
    args = parser.parse_args()
    ParallelGibbs(args.contacts, nproc=args.nproc, ncomp=args.ncomp, niter=args.niter).run(run_resids=args.resid)
