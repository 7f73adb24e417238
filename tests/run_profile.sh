#!/bin/bash
# GPU-box recipe for the committed profile summaries (see profiles/README.md).  Usage: bash tests/run_profile.sh <tag>
# Every .ncu-rep is summarised on the box (ncu -i needs no GPU, but the reports with --import-source are ~30 MB each and
# gpurun only brings 64 MiB back) and then removed; the text summaries are copied to gpurun_out/prof_txt/.
tag=${1:-r1}
# executed FP32 operations (fadd + fmul + 2 ffma = the flops of the compute-side roofline), not part of --set full
FLOPM=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum
mkdir -p gpurun_out/prof_txt
# 1. every launch with its device time (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${tag}.log 2>&1
# 2. the step kernel, full set, once
ncu --set full --metrics $FLOPM --clock-control none --import-source on -k regex:fetch_kernel -s 10 -c 1 -o gpurun_out/prof_${tag} \
    python tests/prof_step.py 4096 12 > gpurun_out/ncu_${tag}.log 2>&1
tail -2 gpurun_out/ncu_${tag}.log
python tests/summarize_profile.py ${tag} > gpurun_out/summarize_${tag}.log 2>&1   # (this one report is kept: read here with ncu -i)
# 3. the Shadow-Hand build of the step kernel (BASELINE config 3: 2048 envs, 92 touch sensors)
ncu --set full --metrics $FLOPM --clock-control none --import-source on -k regex:fetch_kernel -s 6 -c 1 -o gpurun_out/prof_hand_${tag} \
    python tests/prof_hand.py 2048 8 touch > gpurun_out/ncu_hand_${tag}.log 2>&1
tail -2 gpurun_out/ncu_hand_${tag}.log
python tests/summarize_profile.py hand_${tag} > gpurun_out/summarize_hand_${tag}.log 2>&1; rm -f gpurun_out/prof_hand_${tag}.ncu-rep
# 4. the wide build (AdroitHandHammer-v2, 33 dofs, 2048 envs: BASELINE config 5a)
ncu --set full --metrics $FLOPM --clock-control none --import-source on -k regex:fetch_kernel -s 6 -c 1 -o gpurun_out/prof_adroit_${tag} \
    python tests/prof_adroit.py AdroitHandHammer-v2 2048 8 > gpurun_out/ncu_adroit_${tag}.log 2>&1
tail -2 gpurun_out/ncu_adroit_${tag}.log
python tests/summarize_profile.py adroit_${tag} > gpurun_out/summarize_adroit_${tag}.log 2>&1; rm -f gpurun_out/prof_adroit_${tag}.ncu-rep
# 4b. the kitchen build (FrankaKitchen-v1, two-level broad phase, 2048 envs: BASELINE config 5b)
ncu --set full --metrics $FLOPM --clock-control none --import-source on -k regex:fetch_kernel -s 4 -c 1 -o gpurun_out/prof_kitchen_${tag} \
    python tests/prof_kitchen.py 2048 6 > gpurun_out/ncu_kitchen_${tag}.log 2>&1
tail -2 gpurun_out/ncu_kitchen_${tag}.log
python tests/summarize_profile.py kitchen_${tag} > gpurun_out/summarize_kitchen_${tag}.log 2>&1; rm -f gpurun_out/prof_kitchen_${tag}.ncu-rep
cp profiles/*${tag}* profiles/traffic*.json profiles/roofline_*.json gpurun_out/prof_txt/ 2>/dev/null
# 5. compute-sanitizer memcheck: Fetch (30 envs) and an Adroit relocate batch
compute-sanitizer --tool memcheck python tests/sanitize_step.py > gpurun_out/memcheck_${tag}.log 2>&1; tail -3 gpurun_out/memcheck_${tag}.log
compute-sanitizer --tool memcheck python tests/prof_adroit.py AdroitHandRelocate-v2 416 6 > gpurun_out/memcheck_adroit_${tag}.log 2>&1; tail -3 gpurun_out/memcheck_adroit_${tag}.log
