#!/bin/bash
# Round-2 GPU recipe H: the driver's two bench invocations (own arm with `configs`, CPU arm) on the final build.
tag=${1:-r2h}
mkdir -p gpurun_out
timeout 900 python bench.py --gpus 1 --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; cut -c1-300 gpurun_out/bench_${tag}_n1.json; tail -3 gpurun_out/bench_${tag}_n1.err
timeout 400 python bench.py --impl reference --steps 24 --warmup 2 > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_ref.err; cut -c1-300 gpurun_out/bench_${tag}_reference_arm.json
