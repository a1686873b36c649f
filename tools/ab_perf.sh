#!/bin/bash
# developer tool: tools/perf.py for every variants/lib_*.so
for lib in variants/lib_*.so; do
  echo "=== $lib"
  BRTA_LIB=$PWD/$lib timeout 120 python tools/perf.py "$@" 2>&1 | tail -1
done
