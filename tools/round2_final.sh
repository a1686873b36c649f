#!/bin/bash
# Final evidence of round 2 on one B200: full GPU test-suite, smoke, the default bench line, launch list,
# --set full captures (sweep kernel on C2, pindicator, Gaussian mixture), DRAM bytes of one full launch.
set -x
python -m pytest tests -q -m gpu > gpurun_out/r2z_gputest.log 2>&1; tail -3 gpurun_out/r2z_gputest.log
python __graft_entry__.py smoke > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err; cut -c1-600 gpurun_out/r2z_bench_n1.json
python bench.py --niter 2200 --steps 2 --warmup 3 --no-cpu-baseline --skip-legs > gpurun_out/r2z_plain_niter2200.json 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches_bench_niter2200.csv \
    python bench.py --niter 2200 --steps 2 --warmup 3 --no-cpu-baseline --skip-legs > gpurun_out/r2z_ncu_launches.log 2>&1
python tools/perf.py 400 300 > gpurun_out/r2z_plain_perf_400_300.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gibbs_sweep -s 1 -c 1 -o gpurun_out/r2z_full_c2 -f \
    python tools/perf.py 400 300 > gpurun_out/r2z_ncu_full_c2.log 2>&1
python tools/perf_pindicator.py 1000 4000000 15 4 > gpurun_out/r2z_plain_pind.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pindicator_class -s 3 -c 1 -o gpurun_out/r2z_full_pind -f \
    python tools/perf_pindicator.py 1000 4000000 15 4 > gpurun_out/r2z_ncu_full_pind.log 2>&1
python bench.py --only-gmm > gpurun_out/r2z_plain_gmm.json 2>&1
ncu --set full --clock-control none --import-source on -k regex:gmm_fit -s 1 -c 1 -o gpurun_out/r2z_full_gmm -f \
    python bench.py --only-gmm > gpurun_out/r2z_ncu_full_gmm.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gibbs_sweep -s 4 -c 1 --csv \
    --log-file gpurun_out/r2z_dram_bytes_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-legs > gpurun_out/r2z_ncu_dram.log 2>&1
ls -la gpurun_out/r2z_*
