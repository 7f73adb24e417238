#!/bin/bash
# Round-2 GPU recipe U: alignment level 2 (no barriers inside the Newton loop) against the default level 3 on the builds it would change.
tag=${1:-r2u}
mkdir -p gpurun_out
(
for rep in 1 2; do
for f in gymnasium_robotics_b200/libb200sim.so gpurun_variants/libl2.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f timeout 200 python tests/quick_time.py fetch slide reach ant 2>&1 | tail -4
done
done
) > gpurun_out/variants_${tag}.log 2>&1
cat gpurun_out/variants_${tag}.log
