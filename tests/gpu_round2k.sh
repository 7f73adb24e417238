#!/bin/bash
# Round-2 GPU recipe K: what the driver runs at round end (GPU tests, smoke, the bench line) on the final build + racecheck per family.
tag=${1:-r2k}
mkdir -p gpurun_out
(B200_PARITY_STATS=$PWD/gpurun_out/parity_stats_${tag}.json timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/pytest_gpu_${tag}.log; tail -3 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_${tag}.log 2>&1; tail -2 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --gpus 1 --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; cut -c1-300 gpurun_out/bench_${tag}_n1.json; tail -3 gpurun_out/bench_${tag}_n1.err
for spec in "FetchPickAndPlace-v4 30 2" "FetchSlide-v4 7 2" "HandManipulateBlockRotateXYZ_ContinuousTouchSensors-v1 16 1" "AdroitHandHammer-v2 16 2" "AntMaze_UMaze-v5 9 2" "FrankaKitchen-v1 9 1"; do
  echo "== racecheck $spec"
  timeout 420 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/sanitize_one.py $spec 2>&1 | grep -v "Host Frame" | grep -E "sanitize driver done|RACECHECK SUMMARY|Race reported|hazard" | head -8
done > gpurun_out/racecheck_${tag}.log 2>&1
cat gpurun_out/racecheck_${tag}.log
