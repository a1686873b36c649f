#!/bin/bash
# developer tool: tools/perf.py for every variants/lib_*.so (mkdir -p variants first).  Build a variant with
#   python -c "from basicrta_b200 import _cabi; _cabi.build(force=True, extra_flags=['-DBRTA_SERVED_UNROLL=3'], lib_path='variants/lib_u3.so', ncomps=(15,))"
for lib in variants/lib_*.so; do
  echo "=== $lib"
  BRTA_LIB=$PWD/$lib timeout 300 python tools/perf.py "$@" 2>&1 | tail -1
done
