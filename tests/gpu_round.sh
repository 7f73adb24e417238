mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40) > gpurun_out/pytest_gpu5.log
tail -6 gpurun_out/pytest_gpu5.log
for w in fetch_pick_and_place adroit_hammer adroit_relocate adroit_pen adroit_door fetch_slide hand_egg hand_block_touch antmaze_large; do
  timeout 200 python bench.py --workload $w --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/b5_$w.json 2> gpurun_out/b5_$w.err
  echo $w; cut -c1-130 gpurun_out/b5_$w.json | tail -1
done
bash tests/variant_time.sh both 2>&1 | tee gpurun_out/variants5.log | tail -12
