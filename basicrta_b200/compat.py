"""Expose the B200 sampler under the reference's import paths.

``install()`` registers ``basicrta``, ``basicrta.gibbs``, ``basicrta.cluster`` and
``basicrta.util`` in ``sys.modules`` (only if the real package is not importable), so code
written against the reference -- ``from basicrta.gibbs import Gibbs`` (gibbs.py:114-125),
``from basicrta.util import run_residue`` (gibbs.py:50) -- runs unchanged on the GPU path.
With the real basicrta installed, apply the two-line patch of INTEGRATION.md instead.
"""
import importlib
import sys
import types


def install(force=False):
    if not force:
        try:
            importlib.import_module('basicrta.gibbs')
            return False                                  # the reference is present: do not shadow it
        except Exception:
            pass
    from . import cluster, gibbs, util
    pkg = types.ModuleType('basicrta')
    pkg.__path__ = []
    pkg.gibbs, pkg.cluster, pkg.util = gibbs, cluster, util
    sys.modules['basicrta'] = pkg
    sys.modules['basicrta.gibbs'] = gibbs
    sys.modules['basicrta.cluster'] = cluster
    sys.modules['basicrta.util'] = util
    return True
