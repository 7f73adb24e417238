"""Tiny driver for ncu captures of the AntMaze workload (BASELINE config 4 shard): 1024-env AntMaze_Large-v5, a few steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gymnasium_robotics_b200 as grb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
env = grb.make_vec("AntMaze_Large-v5", num_envs=n, rng_mode="torch", autoreset_mode="same_step")
env.reset(seed=0)
g = torch.Generator(device="cuda").manual_seed(1234)
for k in range(steps):
    env.step(torch.rand((n, 8), generator=g, device="cuda") * 2 - 1)
torch.cuda.synchronize()
print("done", env.backend.launches)
