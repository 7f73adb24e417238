#!/bin/bash
# Round-2 GPU recipe M: synccheck of the two flagged builds after the single-exit rewrite of the portal-refinement loops; timing + parity.
tag=${1:-r2m}
mkdir -p gpurun_out
for spec in "FetchSlide-v4 7" "FrankaKitchen-v1 9" "HandManipulateEggRotate-v1 16"; do
  echo "== synccheck $spec"
  timeout 300 compute-sanitizer --tool synccheck --print-limit 2 python tests/sanitize_one.py $spec 2>&1 | grep -v "Host Frame" | grep -E "sanitize driver done|Barrier error|    at |by thread|ERROR SUMMARY" | head -12
done > gpurun_out/synccheck_${tag}.log 2>&1
cat gpurun_out/synccheck_${tag}.log
timeout 300 python tests/quick_time.py fetch kitchen hand 2>&1 | tail -3
(timeout 900 python -m pytest tests -m gpu -q -k "slide or kitchen or egg or pen or hammer" 2>&1 | tail -3) > gpurun_out/pytest_gpu_${tag}.log; tail -2 gpurun_out/pytest_gpu_${tag}.log
