#!/bin/bash
# Round-2 GPU recipe E: base-like build (all revert macros) with one of this round's local-memory edits switched back on each.
tag=${1:-r2e}
mkdir -p gpurun_out
(
for rep in 1 2; do
for f in gpurun_variants/libbase.so gymnasium_robotics_b200/libb200sim.so gpurun_variants/libw_*.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f timeout 300 python tests/quick_time.py fetch 2>&1 | tail -1
done
done
) > gpurun_out/variants_${tag}.log 2>&1
cat gpurun_out/variants_${tag}.log
for v in w_all w_colocal_park; do
(B200SIM_LIB=$PWD/gpurun_variants/lib$v.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k fetch 2>&1 | tail -5) > gpurun_out/pytest_gpu_${tag}_$v.log; tail -2 gpurun_out/pytest_gpu_${tag}_$v.log
done
