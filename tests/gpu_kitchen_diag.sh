#!/bin/bash
# round-2 GPU call: the kitchen build after the eq_data fix
mkdir -p gpurun_out
P=python
(
echo "=== env n=4"; timeout 300 $P tests/kitchen_diag.py 4 env
echo "=== env n=8"; timeout 300 $P tests/kitchen_diag.py 8 env
echo "=== fixture n=8"; timeout 300 $P tests/kitchen_diag.py 8 fixture4
echo "=== groups env n=8"; B200SIM_KITCHEN_GROUPS=1 timeout 300 $P tests/kitchen_diag.py 8 env
echo "=== pytest kitchen"; timeout 900 $P -m pytest tests/test_zz_kitchen_gpu.py -q -m gpu -rA 2>&1 | tail -30
echo "=== synccheck n=4 env"; timeout 900 compute-sanitizer --tool synccheck --print-limit 6 $P tests/kitchen_diag.py 4 env 2>&1 | tail -60
echo "=== racecheck n=9 env"; timeout 900 compute-sanitizer --tool racecheck --print-limit 6 $P tests/kitchen_diag.py 9 env 2>&1 | tail -30
echo "=== memcheck n=9 env"; timeout 900 compute-sanitizer --tool memcheck --print-limit 6 $P tests/kitchen_diag.py 9 env 2>&1 | tail -30
echo "=== synccheck fetch"; timeout 900 compute-sanitizer --tool synccheck --print-limit 6 $P tests/sanitize_step.py 2>&1 | tail -40
for bp in 0 1; do
echo "=== bench kitchen groups=$bp"; B200SIM_KITCHEN_GROUPS=$bp B200SIM_EXPERIMENTAL_KITCHEN=1 timeout 600 $P bench.py --workload franka_kitchen --steps 20 --warmup 3 --no-cpu-baseline
done
) > gpurun_out/kitchen_diag2.log 2>&1
tail -5 gpurun_out/kitchen_diag2.log
