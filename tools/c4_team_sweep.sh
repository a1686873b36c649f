#!/bin/bash
# C4 (one chain, times sharded over the GPUs of the box): team size per GPU against the exchange cost.
# usage: tools/c4_team_sweep.sh NGPUS
N=${1:-2}
for q in 32 128 256 512 1024 2048; do
  echo "== BRTA_MIN_SLICE_QUADS=$q gpus=$N"
  if [ "$N" = "1" ]; then
    BRTA_MIN_SLICE_QUADS=$q python bench.py --config C4 --niter 11000 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step']/11000*1e3, 'us/iter', d['config']['launch'])"
  else
    BRTA_MIN_SLICE_QUADS=$q python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config C4 --niter 11000 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step']/11000*1e3, 'us/iter', d['config']['launch'])"
  fi
done
