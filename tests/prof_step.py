"""Tiny driver for ncu captures: 4096-env FetchPickAndPlace, a few steps (see profiles/README.md)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gymnasium_robotics_b200.fetch import FetchVectorEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
env = FetchVectorEnv("FetchPickAndPlace", num_envs=n, rng_mode="torch", autoreset_mode="same_step")
env.reset(seed=0)
g = torch.Generator(device="cuda").manual_seed(1234)
for k in range(steps):
    env.step(torch.rand((n, 4), generator=g, device="cuda") * 2 - 1)
torch.cuda.synchronize()
print("done", env.backend.launches)
