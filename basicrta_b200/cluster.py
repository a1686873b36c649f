"""``ProcessCluster``: the north star's name for the residue dispatcher.

The reference snapshot has no class of that name (its ``basicrta/cluster.py`` holds
``ProcessProtein``, the post-processing fan-out, cluster.py:15-175); the residue fan-out
lives in ``gibbs.ParallelGibbs`` (gibbs.py:20-88).  ``ProcessCluster`` is that dispatcher
under the name BASELINE.json uses: ``run()`` sends residues to GPUs instead of a
multiprocessing pool.
"""
from .gibbs import ParallelGibbs


class ProcessCluster(ParallelGibbs):
    """``ProcessCluster(contacts, nproc, ncomp, niter).run(run_resids=None)``; ``nproc`` is
    the number of GPUs."""
