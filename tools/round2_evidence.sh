#!/bin/bash
# Evidence runs of round 2 (one B200): every named configuration through bench.py, the launch list, the
# --set full captures and the DRAM traffic of one full launch.  Outputs land in gpurun_out/ and are copied to
# profiles/ (see profiles/README.md).
set -x
for c in C1 C4 C5; do
  python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
done
python bench.py --config C3 --steps 1 --warmup 3 > gpurun_out/r2_bench_C3.json 2> gpurun_out/r2_bench_C3.err
python bench.py --niter 2200 --steps 2 --warmup 3 --no-cpu-baseline --skip-legs > gpurun_out/r2_plain_niter2200.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_niter2200.csv \
    python bench.py --niter 2200 --steps 2 --warmup 3 --no-cpu-baseline --skip-legs > gpurun_out/r2_ncu_launches.log 2>&1
python tools/perf.py 400 300 > gpurun_out/r2_plain_perf_400_300.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gibbs_sweep -s 1 -c 1 -o gpurun_out/r2_full_c2 -f \
    python tools/perf.py 400 300 > gpurun_out/r2_ncu_full_c2.log 2>&1
BRTA_CALIBRATE=1 python tools/perf.py 50 500 shard=8:0 > gpurun_out/r2_plain_perf_50_500.log 2>&1
BRTA_CALIBRATE=1 ncu --set full --clock-control none --import-source on -k regex:gibbs_sweep -s 12 -c 1 -o gpurun_out/r2_full_50chains -f \
    python tools/perf.py 50 500 shard=8:0 > gpurun_out/r2_ncu_full_50.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gibbs_sweep -s 4 -c 1 --csv \
    --log-file gpurun_out/r2_dram_bytes_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-legs > gpurun_out/r2_ncu_dram.log 2>&1
tail -2 gpurun_out/r2_bench_C*.json | cut -c1-400
