#!/bin/bash
# Round-2 GPU recipe R (gpurun --gpus 2): the driver's N=2 invocation on the final commit.
tag=${1:-r2r}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/bench_${tag}_n2.json 2> gpurun_out/bench_${tag}_n2.err; wc -l gpurun_out/bench_${tag}_n2.json; cut -c1-250 gpurun_out/bench_${tag}_n2.json; tail -3 gpurun_out/bench_${tag}_n2.err
