#!/bin/bash
# time library build variants (experiments only): bash tests/variant_time.sh [fetch|hand|both]
for rep in 1 2; do
for f in gpurun_variants/*.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f python tests/quick_time.py ${1:-both} 2>&1 | tail -2
done
done
