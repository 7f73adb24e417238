#!/bin/bash
mkdir -p gpurun_out
(
for v in "" gpurun_variants/libnoB.so gpurun_variants/libnoassume.so; do
  echo "=== variant [$v]"
  if [ -n "$v" ]; then export B200SIM_LIB=$PWD/$v; else unset B200SIM_LIB; fi
  timeout 300 python tests/sanitize_step.py 2>&1 | tail -3
  timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python tests/sanitize_step.py 2>&1 | grep -v "Host Frame\|Saved host" | head -40
done
unset B200SIM_LIB
echo "=== racecheck default"; timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tests/sanitize_step.py 2>&1 | grep -v "Host Frame\|Saved host" | head -60
) > gpurun_out/bisect.log 2>&1
tail -5 gpurun_out/bisect.log
