#!/bin/bash
# compute-sanitizer on a small multi-CTA batch (teams of several CTAs, mailbox exchange, saved rows):
#   racecheck (shared-memory hazards), memcheck, synccheck.  Output: gpurun_out/r2_sanitizer.log
cat > /tmp/san_run.py <<'P'
import sys
import numpy as np
sys.path.insert(0, '.')
from basicrta_b200.engine import ChainInput, get_engine
from basicrta_b200 import _cabi
rng = np.random.default_rng(3)
chains = []
for r, n in enumerate((9000, 2500, 700, 33)):
    comp = rng.choice(3, size=n, p=[0.8, 0.15, 0.05])
    x = rng.exponential(1.0 / np.array([4.0, 0.1, 0.002])[comp])
    chains.append(ChainInput(ticks=np.maximum(np.ceil(x / 0.1), 1).astype(np.int64), ts=0.1, chain_id=r))
eng = get_engine(0)
for K, flags in ((15, 0), (15, _cabi.FLAG_CTAS3), (30, 0), (40, 0)):
    res = eng.run(chains, K, 12, thin=4, seed=1, flags=flags, n_waves=None)
    print('K', K, 'flags', flags, 'status', [x.status for x in res], 'teams ok', res[0].mcweights.shape)
P
for tool in racecheck memcheck synccheck; do
  echo "=== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_run.py 2>&1 | grep -v "^$" | tail -12
done
