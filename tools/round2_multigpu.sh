#!/bin/bash
# Multi-GPU evidence of round 2: bench.py under torchrun on N GPUs (C2 + its c4 / e2e_api legs) and the giant
# single residue C4 as its own configuration.  N = number of GPUs of the box (gpurun --gpus N).
N=${1:-8}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
run 29521 --steps 2 --warmup 3 > gpurun_out/r2_bench_C2_n$N.json 2> gpurun_out/r2_bench_C2_n$N.err
run 29522 --config C4 --steps 2 --warmup 3 > gpurun_out/r2_bench_C4_n$N.json 2> gpurun_out/r2_bench_C4_n$N.err
tail -3 gpurun_out/r2_bench_C2_n$N.err gpurun_out/r2_bench_C4_n$N.err
cat gpurun_out/r2_bench_C2_n$N.json gpurun_out/r2_bench_C4_n$N.json
