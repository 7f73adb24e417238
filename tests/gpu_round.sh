# one GPU-box round: A/B variants, parity tests, bench lines
mkdir -p gpurun_out
bash tests/variant_time.sh both 2>&1 | tee gpurun_out/variants10.log | tail -10
(timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40) > gpurun_out/pytest_gpu10.log; tail -3 gpurun_out/pytest_gpu10.log
for w in hand_block_touch antmaze_large adroit_hammer adroit_door fetch_pick_and_place; do timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/b10_$w.json 2> gpurun_out/b10_$w.err; cut -c1-120 gpurun_out/b10_$w.json; done
