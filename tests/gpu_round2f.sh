#!/bin/bash
# Round-2 GPU recipe F: the split narrow phase (result record on the stack, box-box buffers in shared-memory slots): timing against the
# c156a1f build, DRAM traffic, the GPU tests, synccheck of the non-aligned barrier form.
tag=${1:-r2f}
mkdir -p gpurun_out
(
for rep in 1 2; do
for f in gpurun_variants/libbase.so gymnasium_robotics_b200/libb200sim.so gpurun_variants/libslotassume2.so gpurun_variants/libnonaligned.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f timeout 300 python tests/quick_time.py fetch hand kitchen hammer ant 2>&1 | tail -5
done
done
) > gpurun_out/variants_${tag}.log 2>&1
cat gpurun_out/variants_${tag}.log
(B200_PARITY_STATS=$PWD/gpurun_out/parity_stats_${tag}.json timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/pytest_gpu_${tag}.log; tail -3 gpurun_out/pytest_gpu_${tag}.log
for f in gpurun_variants/libbase.so gymnasium_robotics_b200/libb200sim.so; do
B200SIM_LIB=$PWD/$f ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sass__inst_executed_local_loads,sass__inst_executed_local_stores --clock-control none -k regex:fetch_kernel -s 10 -c 1 python tests/prof_step.py 4096 12 2>&1 | grep -E "dram__|gpu__time|sass__" 
done > gpurun_out/traffic_${tag}.log 2>&1; cat gpurun_out/traffic_${tag}.log
B200SIM_LIB=$PWD/gpurun_variants/libnonaligned.so timeout 600 compute-sanitizer --tool synccheck --print-limit 4 python tests/sanitize_multi.py > gpurun_out/synccheck_${tag}_nonaligned.log 2>&1; tail -3 gpurun_out/synccheck_${tag}_nonaligned.log
