# one GPU-box round: A/B variants, parity tests, bench lines
mkdir -p gpurun_out
bash tests/variant_time.sh both 2>&1 | tee gpurun_out/variants9.log | tail -14
(timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40) > gpurun_out/pytest_gpu9.log; tail -3 gpurun_out/pytest_gpu9.log
for w in fetch_pick_and_place adroit_relocate adroit_hammer hand_block_touch; do timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/b9_$w.json 2> gpurun_out/b9_$w.err; cut -c1-120 gpurun_out/b9_$w.json; done
