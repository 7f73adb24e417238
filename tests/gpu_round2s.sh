#!/bin/bash
# Round-2 GPU recipe S: the full GPU suite incl. the conformance tests on the product path.
tag=${1:-r2s}
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15) > gpurun_out/pytest_gpu_${tag}.log; tail -15 gpurun_out/pytest_gpu_${tag}.log
