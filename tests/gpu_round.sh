# compute-sanitizer racecheck over every kernel build (shared-memory hazards of the hand-rolled warp code)
mkdir -p gpurun_out
timeout 1100 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/sanitize_multi.py > gpurun_out/racecheck_r1j.log 2>&1; grep -c "Race reported" gpurun_out/racecheck_r1j.log; tail -25 gpurun_out/racecheck_r1j.log | cut -c1-200
