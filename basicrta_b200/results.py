"""Attribute-dict used for ``Gibbs.processed_results`` when MDAnalysis is not installed.

The reference uses ``MDAnalysis.analysis.base.Results`` (basicrta/gibbs.py:9, 143); with
MDAnalysis available :mod:`basicrta_b200.gibbs` uses that very class so pickles are
interchangeable with the reference.
"""


class Results(dict):
    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as e:
            raise AttributeError(key) from e

    def __setattr__(self, key, value):
        self[key] = value

    def __delattr__(self, key):
        try:
            del self[key]
        except KeyError as e:
            raise AttributeError(key) from e
