#!/bin/bash
# Round-2 GPU recipe Q: ncu --set full of the Ant build (BASELINE config 4 shard: AntMaze_Large-v5, 1024 envs) and of the FetchSlide build.
tag=${1:-r2q}
FLOPM=smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum
mkdir -p gpurun_out/prof_txt
ncu --set full --metrics $FLOPM --clock-control none --import-source on -k regex:fetch_kernel -s 5 -c 1 -o gpurun_out/prof_ant_${tag} python tests/prof_ant.py 1024 8 > gpurun_out/ncu_ant_${tag}.log 2>&1; tail -2 gpurun_out/ncu_ant_${tag}.log
python tests/summarize_profile.py ant_${tag} > gpurun_out/summarize_ant_${tag}.log 2>&1; tail -2 gpurun_out/summarize_ant_${tag}.log; rm -f gpurun_out/prof_ant_${tag}.ncu-rep
cp profiles/*ant_${tag}* profiles/roofline_antmaze_large.json profiles/traffic*.json gpurun_out/prof_txt/ 2>/dev/null
head -30 gpurun_out/prof_txt/ncu_step_kernel_ant_${tag}.txt
