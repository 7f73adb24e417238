import sys, time; sys.path.insert(0,'.')
import torch
from gymnasium_robotics_b200.hand import HandVectorEnv
for n in (2048, 4096):
    env = HandVectorEnv("HandManipulateBlockRotateXYZ", num_envs=n, rng_mode="torch", autoreset_mode="same_step", max_episode_steps=None)
    env.reset(seed=0)
    g = torch.Generator(device="cuda").manual_seed(1)
    tape = torch.rand((16, n, 20), generator=g, device="cuda") * 2 - 1
    for k in range(5): env.step(tape[k])
    torch.cuda.synchronize(); t=time.time()
    K=40
    for k in range(K): env.step(tape[k % 16])
    torch.cuda.synchronize(); dt=time.time()-t
    print(f"hand N={n}: {dt/K*1e3:.2f} ms/step, {n*K/dt:.0f} env-steps/s")
    env.close()
