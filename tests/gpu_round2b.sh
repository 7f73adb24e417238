#!/bin/bash
# Round-2 GPU recipe B: A/B timing of library variants (gpurun_variants/*.so), stage timing, parity tests + statistics, the bench line.
tag=${1:-r2b}
mkdir -p gpurun_out
(
for rep in 1 2; do
for f in gpurun_variants/libbase.so gymnasium_robotics_b200/libb200sim.so gpurun_variants/libslotassume.so gpurun_variants/libpig.so; do
  echo "== $f"
  B200SIM_LIB=$PWD/$f timeout 300 python tests/quick_time.py fetch hand kitchen hammer ant 2>&1 | tail -5
done
done
echo "== stage timing (pig)"; B200SIM_LIB=$PWD/gpurun_variants/libpigtiming.so timeout 300 python tests/stage_timing.py fetch hand 2>&1 | tail -50
) > gpurun_out/variants_${tag}.log 2>&1
tail -40 gpurun_out/variants_${tag}.log
(B200_PARITY_STATS=$PWD/gpurun_out/parity_stats_${tag}.json timeout 1500 python -m pytest tests -m gpu -q -s 2>&1 | tail -80) > gpurun_out/pytest_gpu_${tag}.log; tail -4 gpurun_out/pytest_gpu_${tag}.log
for v in pig slotassume; do
(B200SIM_LIB=$PWD/gpurun_variants/lib$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_kitchen_gpu.py -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_gpu_${tag}_$v.log; tail -3 gpurun_out/pytest_gpu_${tag}_$v.log
done
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; cut -c1-400 gpurun_out/bench_${tag}_n1.json; tail -3 gpurun_out/bench_${tag}_n1.err
B200SIM_LIB=$PWD/gpurun_variants/libsyncwarp.so timeout 600 compute-sanitizer --tool synccheck --print-limit 4 python tests/sanitize_multi.py > gpurun_out/synccheck_${tag}_syncwarp.log 2>&1; tail -3 gpurun_out/synccheck_${tag}_syncwarp.log
ncu --set full --clock-control none --import-source on -k regex:fetch_kernel -s 10 -c 1 -o gpurun_out/prof_${tag} python tests/prof_step.py 4096 12 > gpurun_out/ncu_${tag}.log 2>&1
python tests/summarize_profile.py ${tag} > gpurun_out/summarize_${tag}.log 2>&1; cp profiles/ncu_step_kernel_${tag}.txt profiles/roofline_fetch_pick_and_place.json gpurun_out/ 2>/dev/null
head -30 gpurun_out/ncu_step_kernel_${tag}.txt
