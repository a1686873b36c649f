#!/bin/bash
# A/B of the two pindicator kernels on one box
for args in "1000 4000000 15 4" "1000 4000001 15 4" "1000 4000002 15 4" "1001 1000003 15 3" "1000 4000000 15 8" "1000 2000001 30 6" "1000 2000000 15 12" "100 4000003 15 4"; do
  python tools/perf_pindicator.py $args
  BRTA_PINDICATOR_GENERIC=1 python tools/perf_pindicator.py $args
done
