#!/bin/bash
# Round-2 GPU recipe N: kitchen parity statistics + the kitchen bench row with the representative kernel arm.
tag=${1:-r2n}
mkdir -p gpurun_out
(B200_PARITY_STATS=$PWD/gpurun_out/parity_stats_${tag}.json timeout 600 python -m pytest tests/test_zz_kitchen_gpu.py -m gpu -q -s 2>&1 | tail -8) > gpurun_out/pytest_gpu_${tag}.log; tail -5 gpurun_out/pytest_gpu_${tag}.log
timeout 600 python bench.py --workload franka_kitchen --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_${tag}_kitchen.json 2> gpurun_out/bench_${tag}_kitchen.err; cut -c1-200 gpurun_out/bench_${tag}_kitchen.json; tail -2 gpurun_out/bench_${tag}_kitchen.err
