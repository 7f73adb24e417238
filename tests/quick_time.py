"""A/B timing helper (experiments): min over repeats of the mean step time, Fetch 4096 envs and Hand 2048 envs."""
import sys, time; sys.path.insert(0, '.')
import torch
from gymnasium_robotics_b200.fetch import FetchVectorEnv
from gymnasium_robotics_b200.hand import HandVectorEnv

def bench(env, nact, n, K=40, reps=3):
    g = torch.Generator(device="cuda").manual_seed(1)
    tape = torch.rand((16, n, nact), generator=g, device="cuda") * 2 - 1
    for k in range(5): env.step(tape[k])
    best = 1e9
    for r in range(reps):
        torch.cuda.synchronize(); t = time.time()
        for k in range(K): env.step(tape[k % 16])
        torch.cuda.synchronize(); best = min(best, (time.time() - t) / K)
    return best

which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("both", "fetch"):
    n = 4096
    env = FetchVectorEnv("FetchPickAndPlace", num_envs=n, rng_mode="torch", autoreset_mode="same_step")
    env.reset(seed=0)
    dt = bench(env, 4, n)
    print(f"fetch N={n}: {dt*1e3:.3f} ms/step, {n/dt:.0f} env-steps/s")
    env.close()
if which in ("both", "hand"):
    n = 2048
    env = HandVectorEnv("HandManipulateBlockRotateXYZ", num_envs=n, rng_mode="torch", autoreset_mode="same_step", max_episode_steps=None)
    env.reset(seed=0)
    dt = bench(env, 20, n)
    print(f"hand N={n}: {dt*1e3:.3f} ms/step, {n/dt:.0f} env-steps/s")
    env.close()
