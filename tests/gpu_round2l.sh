#!/bin/bash
# Round-2 GPU recipe L (gpurun --gpus 4): the driver's N=4 invocation of both arms; stdout must be exactly one JSON line.
tag=${1:-r2l}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 50 --warmup 5 > gpurun_out/bench_${tag}_n4.json 2> gpurun_out/bench_${tag}_n4.err; wc -l gpurun_out/bench_${tag}_n4.json; cut -c1-250 gpurun_out/bench_${tag}_n4.json; tail -3 gpurun_out/bench_${tag}_n4.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 4 --steps 8 --warmup 1 > gpurun_out/bench_${tag}_n4_reference.json 2> gpurun_out/bench_${tag}_n4_reference.err; wc -l gpurun_out/bench_${tag}_n4_reference.json; cut -c1-200 gpurun_out/bench_${tag}_n4_reference.json
