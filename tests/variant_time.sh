#!/bin/bash
# time library build variants (experiments only)
for f in gpurun_variants/*.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f python tests/quick_time.py 2>&1 | tail -2
done
