# one GPU-box round: parity tests, the official bench lines, the profile recipe (see profiles/README.md)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40) > gpurun_out/pytest_gpu7.log; tail -4 gpurun_out/pytest_gpu7.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1i_n1.json 2> gpurun_out/bench_r1i_n1.err; cut -c1-200 gpurun_out/bench_r1i_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1i_ref.json 2> gpurun_out/bench_r1i_ref.err; cut -c1-200 gpurun_out/bench_r1i_ref.json
for w in hand_block_touch adroit_hammer antmaze_large; do timeout 200 python bench.py --workload $w --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_r1i_$w.json 2> gpurun_out/b7_$w.err; cut -c1-120 gpurun_out/bench_r1i_$w.json; done
bash tests/run_profile.sh r1i 2>&1 | tail -20
du -sh gpurun_out
