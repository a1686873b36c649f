"""Developer tool: repeat tests/test_gpu_exact_parity.py::test_memoised_rows_are_bit_identical_to_recomputation[30]
(a 234-CTA team on the L2-atomics exchange path, K = 30) N times and count runs whose table / no-table results
differ.      python tools/flake_k30.py [N]"""
import sys

import numpy as np

sys.path.insert(0, '.')
from basicrta_b200 import _cabi  # noqa: E402
from basicrta_b200.engine import ChainInput, get_engine  # noqa: E402
from oracle import gibbs_oracle as O  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
eng = get_engine(0)
times = O.synth_times(30000, [0.7, 0.2, 0.1], [5, 0.3, 0.004], seed=12)
ticks = O.to_ticks(times, 0.1)
chains = [ChainInput(ticks=ticks, ts=0.1, chain_id=7), ChainInput(ticks=ticks[:1237], ts=0.1, chain_id=8)]
ref = None
bad_table = bad_plain = 0
for rep in range(n):
    a = eng.run(chains, 30, 400, thin=50, seed=21)
    b = eng.run(chains, 30, 400, thin=50, seed=21, flags=_cabi.FLAG_NO_TABLE)
    if ref is None:
        ref = b
    bad_table += any(not (np.array_equal(x.mcrates, r.mcrates) and np.array_equal(x.indicator, r.indicator)) for x, r in zip(a, ref))
    bad_plain += any(not (np.array_equal(x.mcrates, r.mcrates) and np.array_equal(x.indicator, r.indicator)) for x, r in zip(b, ref))
print(f'{n} repetitions: table runs differing from the first no-table run: {bad_table}; no-table runs differing: {bad_plain}')
