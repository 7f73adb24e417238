#!/bin/bash
# Round-2 GPU recipe D: which of the local-memory edits since c156a1f costs the 10 %?  One revert macro per variant (tests/build_variants.py),
# plus the non-aligned barrier form (timing + synccheck).
tag=${1:-r2d}
mkdir -p gpurun_out
(
for rep in 1 2; do
for f in gpurun_variants/libbase.so gymnasium_robotics_b200/libb200sim.so gpurun_variants/libv_cbf.so gpurun_variants/libv_lim.so gpurun_variants/libv_eul.so gpurun_variants/libv_park.so gpurun_variants/libv_all.so gpurun_variants/libnonaligned.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f timeout 300 python tests/quick_time.py fetch hand hammer 2>&1 | tail -3
done
done
) > gpurun_out/variants_${tag}.log 2>&1
tail -40 gpurun_out/variants_${tag}.log
B200SIM_LIB=$PWD/gpurun_variants/libnonaligned.so timeout 600 compute-sanitizer --tool synccheck --print-limit 4 python tests/sanitize_multi.py > gpurun_out/synccheck_${tag}_nonaligned.log 2>&1; tail -3 gpurun_out/synccheck_${tag}_nonaligned.log
(B200SIM_LIB=$PWD/gpurun_variants/libv_all.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -5) > gpurun_out/pytest_gpu_${tag}_v_all.log; tail -3 gpurun_out/pytest_gpu_${tag}_v_all.log
