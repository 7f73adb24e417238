#!/bin/bash
# Round-2 GPU recipe I (gpurun --gpus 2): the driver's N=2 invocation, then the same with the side-stream gather of the packed rows.
tag=${1:-r2i}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_${tag}_n2.json 2> gpurun_out/bench_${tag}_n2.err; cut -c1-250 gpurun_out/bench_${tag}_n2.json; tail -3 gpurun_out/bench_${tag}_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 --warmup 5 --gather --no-configs > gpurun_out/bench_${tag}_n2_gather.json 2> gpurun_out/bench_${tag}_n2_gather.err; cut -c1-250 gpurun_out/bench_${tag}_n2_gather.json; tail -3 gpurun_out/bench_${tag}_n2_gather.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 8 --warmup 1 > gpurun_out/bench_${tag}_n2_reference.json 2> gpurun_out/bench_${tag}_n2_reference.err; cut -c1-200 gpurun_out/bench_${tag}_n2_reference.json
