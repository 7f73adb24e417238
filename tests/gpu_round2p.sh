#!/bin/bash
# Round-2 GPU recipe P: the driver's round-end sequence on the final commit (GPU tests, smoke, both bench arms).
tag=${1:-r2p}
mkdir -p gpurun_out
(B200_PARITY_STATS=$PWD/gpurun_out/parity_stats_${tag}.json timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > gpurun_out/pytest_gpu_${tag}.log; tail -3 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_${tag}.log 2>&1; tail -2 gpurun_out/smoke_${tag}.log
timeout 400 python bench.py --impl reference --gpus 1 --steps 24 --warmup 2 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_ref.err; cut -c1-200 gpurun_out/bench_${tag}_reference_arm.json
timeout 900 python bench.py --gpus 1 --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; cut -c1-300 gpurun_out/bench_${tag}_n1.json; tail -3 gpurun_out/bench_${tag}_n1.err
