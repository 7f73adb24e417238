#!/bin/bash
# Round-2 GPU recipe G: the committed profile set of the final kernels (launch list, ncu --set full per build, memcheck), racecheck,
# synccheck per family in separate processes (a barrier report aborts the launch).
tag=${1:-r2g}
mkdir -p gpurun_out
bash tests/run_profile.sh ${tag} 2>&1 | tail -20
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/sanitize_multi.py > gpurun_out/racecheck_${tag}.log 2>&1; tail -3 gpurun_out/racecheck_${tag}.log
for spec in "FetchPickAndPlace-v4 30" "FetchSlide-v4 7" "FetchSlide-v4 10" "HandManipulateBlockRotateXYZ_ContinuousTouchSensors-v1 16" "AdroitHandHammer-v2 16" "AntMaze_UMaze-v5 9" "FrankaKitchen-v1 9"; do
  echo "== synccheck $spec"
  timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python tests/sanitize_one.py $spec 2>&1 | grep -v "Host Frame" | grep -E "sanitize driver done|Barrier error|    at |by thread|ERROR SUMMARY" | head -12
done > gpurun_out/synccheck_${tag}.log 2>&1
cat gpurun_out/synccheck_${tag}.log
