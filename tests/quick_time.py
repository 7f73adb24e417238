"""A/B timing helper (experiments): min over repeats of the mean step time of the bench workloads' step KERNEL path
(env.step with device-resident actions).  python tests/quick_time.py [fetch hand kitchen hammer ant ...]"""
import sys, time; sys.path.insert(0, '.')
import torch
import gymnasium_robotics_b200 as grb

SPEC = {"fetch": ("FetchPickAndPlace-v4", 4096, 4), "hand": ("HandManipulateBlockRotateXYZ_ContinuousTouchSensors-v1", 2048, 20),
        "kitchen": ("FrankaKitchen-v1", 2048, 9), "hammer": ("AdroitHandHammer-v2", 2048, 26), "ant": ("AntMaze_Large-v5", 1024, 8),
        "slide": ("FetchSlide-v4", 4096, 4), "reach": ("FetchReach-v4", 4096, 4)}


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


which = sys.argv[1:] or ["fetch", "hand"]
if which == ["both"]:
    which = ["fetch", "hand"]
for w in which:
    env_id, n, nact = SPEC[w]
    env = grb.make_vec(env_id, num_envs=n, rng_mode="torch", autoreset_mode="same_step", **({"max_episode_steps": None} if w == "hand" else {}))
    env.reset(seed=0)
    dt = bench(env, nact, n, K=20 if w == "kitchen" else 40)
    print(f"{w} N={n}: {dt*1e3:.3f} ms/step, {n/dt:.0f} env-steps/s", flush=True)
    env.close()
