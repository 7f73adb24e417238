#!/bin/bash
# Round-2 GPU recipe C: A/B timing of the narrow-phase variants (result slots in shared memory vs on the stack; box-box forms), then the GPU tests.
tag=${1:-r2c}
mkdir -p gpurun_out
(
for rep in 1 2; do
for f in gpurun_variants/libbase.so gymnasium_robotics_b200/libb200sim.so gpurun_variants/libcolocal.so gpurun_variants/libboxold.so gpurun_variants/libcolocal_boxold.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f timeout 300 python tests/quick_time.py fetch hand kitchen hammer ant 2>&1 | tail -5
done
done
) > gpurun_out/variants_${tag}.log 2>&1
tail -30 gpurun_out/variants_${tag}.log
(B200_PARITY_STATS=$PWD/gpurun_out/parity_stats_${tag}.json timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -80) > gpurun_out/pytest_gpu_${tag}.log; tail -4 gpurun_out/pytest_gpu_${tag}.log
for v in colocal_boxold; do
(B200SIM_LIB=$PWD/gpurun_variants/lib$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_kitchen_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_${tag}_$v.log; tail -3 gpurun_out/pytest_gpu_${tag}_$v.log
done
