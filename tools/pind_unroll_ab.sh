#!/bin/bash
for lib in variants/lib_pf*.so; do
  echo "=== $lib"
  for args in "1000 4000000 15 4" "1000 4000001 15 4" "1001 1000003 15 3" "1000 4000000 15 8"; do
    BRTA_LIB=$PWD/$lib python tools/perf_pindicator.py $args
  done
done
