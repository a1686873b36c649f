"""basicrta_b200 -- B200-native Gibbs sampler behind basicrta's ``Gibbs`` API.

Only the hot path of orbeckst/basicrta is here: the exponential-mixture Gibbs sweep
(basicrta/gibbs.py:176-219) and the fan-out of residues (gibbs.py:20-88), as a
hand-written sm_100a kernel behind a C ABI (``include/basicrta_b200.h``).
"""
__version__ = '0.1.0'
