#!/bin/bash
# One GPU-box round, as run through `gpurun -- 'bash tests/gpu_round.sh <tag>'` for the numbers under profiles/:
#   1. the -m gpu parity tests (CUDA path through the C-ABI against the fp64 oracle),
#   2. the headline bench line, the CPU arm and the other BASELINE workloads,
#   3. the profile recipe (launch list, ncu --set full of the Fetch / Hand / Adroit builds, memcheck) -- see run_profile.sh,
#   4. optionally (RACECHECK=1) compute-sanitizer racecheck over every kernel build (slow: ~10 minutes).
tag=${1:-r1}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40) > gpurun_out/pytest_gpu_${tag}.log; tail -3 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; cut -c1-200 gpurun_out/bench_${tag}_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_ref.err
for w in fetch_slide hand_block_touch hand_egg antmaze_large adroit_hammer adroit_relocate adroit_pen adroit_door; do
  timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${tag}_$w.json 2> gpurun_out/bench_${tag}_$w.err
  cut -c1-120 gpurun_out/bench_${tag}_$w.json
done
bash tests/run_profile.sh ${tag} 2>&1 | tail -16
if [ -n "$RACECHECK" ]; then
  timeout 1100 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/sanitize_multi.py > gpurun_out/racecheck_${tag}.log 2>&1
  tail -3 gpurun_out/racecheck_${tag}.log
fi
