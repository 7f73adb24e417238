#!/bin/bash
# Round-2 GPU recipe W: the hull build of the kitchen kernels (support-map narrow phase for the Franka meshes): parity test + statistics,
# timing against the box-proxy default, memcheck of one small batch.
tag=${1:-r2w}
mkdir -p gpurun_out
(B200_PARITY_STATS=$PWD/gpurun_out/parity_stats_${tag}.json timeout 600 python -m pytest tests/test_zz_kitchen_gpu.py -m gpu -q -s 2>&1 | tail -12) > gpurun_out/pytest_gpu_${tag}.log; tail -8 gpurun_out/pytest_gpu_${tag}.log
timeout 300 python - <<'PY' 2>&1 | tail -4
import sys, time; sys.path.insert(0, '.')
import torch
import gymnasium_robotics_b200 as grb
for mc in ("box", "hull"):
    n = 2048
    env = grb.make_vec("FrankaKitchen-v1", num_envs=n, rng_mode="torch", autoreset_mode="same_step", mesh_collision=mc)
    env.reset(seed=0)
    g = torch.Generator(device="cuda").manual_seed(1)
    tape = torch.rand((16, n, 9), generator=g, device="cuda") * 2 - 1
    for k in range(5): env.step(tape[k])
    best = 1e9
    for r in range(3):
        torch.cuda.synchronize(); t = time.time()
        for k in range(20): env.step(tape[k % 16])
        torch.cuda.synchronize(); best = min(best, (time.time() - t) / 20)
    print(f"kitchen mesh_collision={mc} N={n}: {best*1e3:.3f} ms/step, {n/best:.0f} env-steps/s, overflow count {int(env.solver_overflow_count) if hasattr(env, 'solver_overflow_count') else 'n/a'}", flush=True)
    env.close()
PY
timeout 300 compute-sanitizer --tool memcheck python - <<'PY' 2>&1 | grep -E "ERROR SUMMARY|done" | tail -3
import sys; sys.path.insert(0, '.')
import torch
import gymnasium_robotics_b200 as grb
env = grb.make_vec("FrankaKitchen-v1", num_envs=9, rng_mode="torch", mesh_collision="hull")
env.reset(seed=0)
for k in range(2): env.step(torch.full((9, 9), 0.5 if k else -0.5, device="cuda"))
torch.cuda.synchronize(); print("memcheck driver done")
PY
