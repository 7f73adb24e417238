# one GPU-box round: the official bench lines and the profile recipe of the final build (see profiles/README.md)
mkdir -p gpurun_out
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1j_n1.json 2> gpurun_out/bench_r1j_n1.err; cut -c1-200 gpurun_out/bench_r1j_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1j_ref.json 2> gpurun_out/bench_r1j_ref.err; cut -c1-200 gpurun_out/bench_r1j_ref.json
bash tests/run_profile.sh r1j 2>&1 | tail -16
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1j_n1b.json 2> gpurun_out/bench_r1j_n1b.err; cut -c1-200 gpurun_out/bench_r1j_n1b.json
du -sh gpurun_out
