#!/bin/bash
# GPU-box recipe for the Franka-Kitchen bring-up build, cheapest check first:
#   bash tests/gpu_round_kitchen.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
# 1. torch-free replay of tests/golden/kitchen_quick.npz through the C-ABI (seconds)
timeout 120 python tests/kitchen_gpu_quick.py > gpurun_out/kitchen_quick_${tag}.log 2>&1; tail -2 gpurun_out/kitchen_quick_${tag}.log
# 2. the env-level GPU tests (7- and 10-warp variants)
(timeout 600 python -m pytest tests/test_zz_kitchen_gpu.py -m gpu -q -s -rxX 2>&1 | tail -15) > gpurun_out/pytest_kitchen_${tag}.log; tail -3 gpurun_out/pytest_kitchen_${tag}.log
# 3. memcheck + racecheck of a small batch
timeout 600 compute-sanitizer --tool memcheck python tests/prof_kitchen.py 20 2 > gpurun_out/memcheck_kitchen_${tag}.log 2>&1; tail -2 gpurun_out/memcheck_kitchen_${tag}.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis python tests/prof_kitchen.py 10 1 > gpurun_out/racecheck_kitchen_${tag}.log 2>&1; tail -2 gpurun_out/racecheck_kitchen_${tag}.log
# 4. bench line (2 048 envs, 40 sub-steps per env-step)
timeout 300 python bench.py --workload franka_kitchen --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_franka_kitchen.json 2> gpurun_out/bench_${tag}_franka_kitchen.err; cut -c1-160 gpurun_out/bench_${tag}_franka_kitchen.json
# 4b. A/B: the two-level broad-phase build (B200SIM_KITCHEN_GROUPS=1) at 10 and 11 envs per block (11 fits since its pair list
#     is out of shared memory), and the flat build at 7
timeout 120 env B200SIM_KITCHEN_GROUPS=1 python tests/kitchen_gpu_quick.py > gpurun_out/kitchen_quick_groups_${tag}.log 2>&1; tail -1 gpurun_out/kitchen_quick_groups_${tag}.log
for v in "1 10" "1 11" "0 7"; do
  set -- $v
  B200SIM_KITCHEN_GROUPS=$1 B200SIM_WPB=$2 timeout 300 python bench.py --workload franka_kitchen --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_franka_kitchen_g$1_wpb$2.json 2>/dev/null
  echo "groups $1 wpb $2: $(cut -c1-90 gpurun_out/bench_${tag}_franka_kitchen_g$1_wpb$2.json)"
done
# 5. one full ncu capture of the kitchen step kernel
ncu --set full --clock-control none --import-source on -k regex:fetch_kernel -s 3 -c 1 -o gpurun_out/prof_kitchen_${tag} \
    python tests/prof_kitchen.py 2048 5 > gpurun_out/ncu_kitchen_${tag}.log 2>&1
tail -2 gpurun_out/ncu_kitchen_${tag}.log
python tests/summarize_profile.py kitchen_${tag} > gpurun_out/summarize_kitchen_${tag}.log 2>&1; rm -f gpurun_out/prof_kitchen_${tag}.ncu-rep
# 6. BASELINE config 5 (needs `gpurun --gpus 4`): 2 ranks AdroitHandHammer-v2 + 2 ranks FrankaKitchen-v1, 1 024 envs per GPU
if [ "$(nvidia-smi -L | wc -l)" -ge 4 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 \
      --workload mixed_hammer_kitchen --steps 30 --warmup 5 > gpurun_out/bench_${tag}_mixed_n4.json 2> gpurun_out/bench_${tag}_mixed_n4.err
  cut -c1-160 gpurun_out/bench_${tag}_mixed_n4.json
fi
