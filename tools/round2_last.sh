#!/bin/bash
# Last evidence run of round 2 (one B200): full GPU suite, default bench line, --set full capture of the schedule
# the engine now chooses for C2 (3 CTAs per SM, 3 waves).
set -x
python -m pytest tests -q -m gpu > gpurun_out/r2y_gputest.log 2>&1; tail -3 gpurun_out/r2y_gputest.log
python bench.py > gpurun_out/r2y_bench_n1.json 2> gpurun_out/r2y_bench_n1.err; cut -c1-300 gpurun_out/r2y_bench_n1.json
BRTA_CALIBRATE=1 python tools/perf.py 400 300 > gpurun_out/r2y_plain_perf_400_300.log 2>&1; cat gpurun_out/r2y_plain_perf_400_300.log
BRTA_CALIBRATE=1 ncu --set full --clock-control none --import-source on -k regex:gibbs_sweep -s 15 -c 1 -o gpurun_out/r2y_full_c2 -f \
    python tools/perf.py 400 300 > gpurun_out/r2y_ncu_full_c2.log 2>&1; tail -2 gpurun_out/r2y_ncu_full_c2.log
