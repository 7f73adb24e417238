"""Tiny driver for ncu / compute-sanitizer runs of the FrankaKitchen-v1 bring-up build (csrc/b200sim_kitchen.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gymnasium_robotics_b200 as grb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
env = grb.make_vec("FrankaKitchen-v1", num_envs=n, rng_mode="torch", autoreset_mode="same_step")
env.reset(seed=0)
g = torch.Generator(device="cuda").manual_seed(1234)
for k in range(steps):
    env.step((torch.rand((n, 9), generator=g, device="cuda") * 2 - 1).contiguous())
torch.cuda.synchronize()
print("done FrankaKitchen-v1", n, "envs, launches", env.backend.launches)
env.close()
