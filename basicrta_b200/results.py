"""Attribute-dict used for ``Gibbs.processed_results`` when MDAnalysis is not installed.

The reference uses ``MDAnalysis.analysis.base.Results`` (basicrta/gibbs.py:9, 143): a ``UserDict``
whose items read and write as attributes and whose pickle state is the plain ``data`` dict.  This
stand-in has the same behaviour AND the same pickle shape (``__getstate__`` returns ``data``,
``__setstate__`` takes it back), so a pickle written here -- where :mod:`basicrta_b200.gibbs` names the
class ``MDAnalysis.analysis.base.Results`` in the stream -- loads into the real class on a machine
that has MDAnalysis, and vice versa.  With MDAnalysis available the real class is used directly.
"""
from collections import UserDict


class Results(UserDict):
    def __setattr__(self, key, value):
        if key == 'data':
            super().__setattr__(key, value)
        else:
            self[key] = value

    def __getattr__(self, key):
        if key == 'data':                                  # not yet set (during unpickling)
            raise AttributeError(key)
        try:
            return self.data[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __delattr__(self, key):
        try:
            del self.data[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __getstate__(self):
        return self.data

    def __setstate__(self, state):
        self.data = state
