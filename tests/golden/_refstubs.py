"""Import the UNMODIFIED reference sampler from /root/reference in the build container.

The reference imports matplotlib, seaborn and MDAnalysis at module top
(basicrta/gibbs.py:5-12, basicrta/util.py:3-17) and its ``__init__`` needs installed
package metadata (basicrta/__init__.py:9); none of that is in this image.  Registering
inert stand-ins lets ``basicrta.gibbs`` / ``basicrta.util`` import byte-for-byte
unchanged.  Only used to GENERATE golden vectors (``make_golden.py``) and by CPU-side
tests that are skipped when /root/reference is absent (it is absent on the GPU box).
"""
import os
import sys
import types

REFERENCE_ROOT = '/root/reference'


class _Anything(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        mod = _Anything(self.__name__ + '.' + name)
        setattr(self, name, mod)
        return mod

    def __call__(self, *a, **k):
        return self


class Results(dict):
    """Stand-in for MDAnalysis.analysis.base.Results: a dict with attribute access."""
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def have_reference():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'basicrta'))


def import_reference():
    """Return (basicrta.gibbs, basicrta.util) of the reference, imported with stubs."""
    if 'basicrta.gibbs' in sys.modules and getattr(sys.modules['basicrta'], '_is_ref_stub', False):
        return sys.modules['basicrta.gibbs'], sys.modules['basicrta.util']
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker', 'matplotlib.patches',
                 'matplotlib.collections', 'seaborn', 'MDAnalysis', 'MDAnalysis.analysis',
                 'MDAnalysis.analysis.base', 'MDAnalysis.lib', 'MDAnalysis.lib.util'):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    sys.modules['matplotlib'].rcParams = {}
    sys.modules['MDAnalysis.analysis.base'].Results = Results
    pkg = types.ModuleType('basicrta')
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'basicrta')]
    pkg._is_ref_stub = True
    for k in [k for k in sys.modules if k == 'basicrta' or k.startswith('basicrta.')]:
        del sys.modules[k]
    sys.modules['basicrta'] = pkg
    import importlib
    gibbs = importlib.import_module('basicrta.gibbs')
    util = importlib.import_module('basicrta.util')
    return gibbs, util
