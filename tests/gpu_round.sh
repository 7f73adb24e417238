# one GPU-box round: A/B variants, parity tests, bench lines
mkdir -p gpurun_out
bash tests/variant_time.sh both 2>&1 | tee gpurun_out/variants11.log | tail -10
(timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40) > gpurun_out/pytest_gpu11.log; tail -3 gpurun_out/pytest_gpu11.log
for w in hand_block_touch adroit_hammer adroit_relocate adroit_pen hand_egg fetch_slide; do timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/b11_$w.json 2> gpurun_out/b11_$w.err; cut -c1-120 gpurun_out/b11_$w.json; done
