#!/bin/bash
# Round-2 GPU-box recipe (gpurun -- 'bash tests/gpu_round2.sh <tag>'): parity tests, the driver's bench line (headline + `configs`),
# the CPU arm, the profile recipe (launch list, ncu --set full of the Fetch / Hand / Adroit / Kitchen builds, memcheck), sanitizers.
tag=${1:-r2a}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -s -x 2>&1 | tail -60) > gpurun_out/pytest_gpu_${tag}.log; tail -4 gpurun_out/pytest_gpu_${tag}.log
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; cut -c1-300 gpurun_out/bench_${tag}_n1.json; tail -3 gpurun_out/bench_${tag}_n1.err
timeout 300 python bench.py --impl reference --steps 24 --warmup 2 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_ref.err; cut -c1-200 gpurun_out/bench_${tag}_reference_arm.json
if [ -z "$NOPROFILE" ]; then bash tests/run_profile.sh ${tag} 2>&1 | tail -20; fi
if [ -n "$SANITIZE" ]; then
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/sanitize_multi.py > gpurun_out/racecheck_${tag}.log 2>&1; tail -3 gpurun_out/racecheck_${tag}.log
  timeout 600 compute-sanitizer --tool synccheck --print-limit 4 python tests/sanitize_multi.py > gpurun_out/synccheck_${tag}.log 2>&1; tail -3 gpurun_out/synccheck_${tag}.log
fi
