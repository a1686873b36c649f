"""``ProcessCluster``: the north star's name for the residue dispatcher, and the GPU form of
``ProcessProtein.reprocess``.

The reference snapshot has no class named ``ProcessCluster`` (its ``basicrta/cluster.py`` holds
``ProcessProtein``, the post-processing fan-out, cluster.py:15-175); the residue fan-out
lives in ``gibbs.ParallelGibbs`` (gibbs.py:20-88).  ``ProcessCluster`` is that dispatcher
under the name BASELINE.json uses: ``run()`` sends residues to GPUs instead of a
multiprocessing pool.

``ProcessProtein`` mirrors the two methods of the reference class that sit on the clustering path
(SURVEY.md 8 f-4): ``reprocess`` (cluster.py:54-76 -- ``Gibbs.load`` + ``process_gibbs`` of every residue
directory over a process pool, 117 scikit-learn fits each) runs the mixture fits of all residues as one
GPU batch, and ``collect_results`` (cluster.py:78-95).
"""
import os
from glob import glob

from .gibbs import Gibbs, ParallelGibbs


class ProcessCluster(ParallelGibbs):
    """``ProcessCluster(contacts, nproc, ncomp, niter).run(run_resids=None)``; ``nproc`` is
    the number of GPUs."""


class ProcessProtein(object):
    """``ProcessProtein(niter, prot, cutoff)`` (cluster.py:30-34)."""

    def __init__(self, niter, prot, cutoff):
        self.residues = {}
        self.niter = niter
        self.prot = prot
        self.cutoff = cutoff

    def __getitem__(self, item):
        return getattr(self, item)

    def _dirs(self):
        dirs = glob(f'basicrta-{self.cutoff}/?[0-9]*')                       # cluster.py:64-67
        return sorted(dirs, key=lambda d: int(os.path.basename(d)[1:]))

    def reprocess(self, nproc=1, device=0, batch=512, seed=None):
        """Re-cluster every residue that has a ``gibbs_{niter}.pkl`` (cluster.py:54-76).  ``nproc`` is kept for
        the signature; the 117 restarts of up to ``batch`` residues go to GPU ``device`` as one launch."""
        from . import postprocess
        files = [f'{d}/gibbs_{self.niter}.pkl' for d in self._dirs()]
        missing = [f for f in files if not os.path.exists(f)]
        for f in missing:
            print(f'results for {os.path.dirname(f)} do not exist')          # cluster.py:49
        files = [f for f in files if os.path.exists(f)]
        done = []
        for lo in range(0, len(files), batch):
            gibbs_list = [Gibbs.load(f) for f in files[lo:lo + batch]]
            done += [g.residue for g in postprocess.reprocess_batch(gibbs_list, device=device, seed=seed)]
        return done

    def collect_results(self):
        """cluster.py:78-95: ``residues[name]`` = path of the residue's pickle, or None."""
        for d in self._dirs():
            f = f'{d}/gibbs_{self.niter}.pkl'
            self.residues[os.path.basename(d)] = f if os.path.exists(f) else None
        return self.residues

    def get_taus(self):
        """tau of the slowest process of every collected residue and its error bars (cluster.py:97-127); residues
        without results, or not yet processed, count as ``[0, 0, 0]`` like in the reference."""
        import numpy as np

        from . import postprocess
        from .util import get_bars
        taus = []
        for res, path in self.residues.items():
            row = [0, 0, 0]
            if path is not None:
                try:
                    row = postprocess.estimate_tau(Gibbs.load(path))
                except AttributeError:                                      # no processed_results yet
                    pass
            taus.append(row)
        taus = np.array(taus, dtype=np.float64).reshape(-1, 3)
        return taus[:, 1], get_bars(taus)
