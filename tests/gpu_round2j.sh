#!/bin/bash
# Round-2 GPU recipe J: where the block-wide alignment barriers sit (Fetch build): Newton-loop barriers on / off, first barrier of the
# sub-step on / off, barrier before the collision stage on / off.
tag=${1:-r2j}
mkdir -p gpurun_out
(
for rep in 1 2; do
for f in gymnasium_robotics_b200/libb200sim.so gpurun_variants/liba_*.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f timeout 300 python tests/quick_time.py fetch 2>&1 | tail -1
done
done
) > gpurun_out/variants_${tag}.log 2>&1
cat gpurun_out/variants_${tag}.log
