# one GPU-box round: variant A/B timing, the official bench line, the profile recipe (see profiles/README.md)
mkdir -p gpurun_out
bash tests/variant_time.sh both 2>&1 | tee gpurun_out/variants6.log | tail -12
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1i_n1.json 2> gpurun_out/bench_r1i_n1.err; cut -c1-200 gpurun_out/bench_r1i_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1i_ref.json 2> gpurun_out/bench_r1i_ref.err; cut -c1-200 gpurun_out/bench_r1i_ref.json
bash tests/run_profile.sh r1i 2>&1 | tail -20
