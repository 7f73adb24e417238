#!/bin/bash
# Round-2 GPU recipe O: alignment groups inside the block (named barriers): 28 (default) vs 2 x 14 vs 4 x 7 warps (Fetch), 14 vs 2 x 7 (Hand).
tag=${1:-r2o}
mkdir -p gpurun_out
(
for rep in 1 2; do
for f in gymnasium_robotics_b200/libb200sim.so gpurun_variants/libg14.so gpurun_variants/libg14_l2.so gpurun_variants/libg7.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f timeout 120 python tests/quick_time.py fetch hand 2>&1 | tail -2
done
done
) > gpurun_out/variants_${tag}.log 2>&1
cat gpurun_out/variants_${tag}.log
