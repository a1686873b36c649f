#!/bin/bash
# developer tool: run tools/phase_timing.py for every variants/lib_*.so
for lib in variants/lib_*.so; do
  echo "=== $lib"
  BRTA_LIB=$PWD/$lib timeout 120 python tools/phase_timing.py "$@" 2>&1 | tail -10
done
