"""Tiny driver for ncu captures of the Shadow-Hand workload: 2048-env HandManipulateBlockRotateXYZ, a few steps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gymnasium_robotics_b200.hand import HandVectorEnv

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 12
touch = "sensordata" if len(sys.argv) > 3 and sys.argv[3] == "touch" else None
env = HandVectorEnv("HandManipulateBlockRotateXYZ", num_envs=n, rng_mode="torch", autoreset_mode="same_step", touch_get_obs=touch)
env.reset(seed=0)
g = torch.Generator(device="cuda").manual_seed(1234)
info = torch.zeros(n, dtype=torch.int32, device="cuda")
out = env.backend.new_outputs()
for k in range(steps):
    a = (torch.rand((n, 20), generator=g, device="cuda") * 2 - 1).contiguous()
    env.backend.step(a, out, info)
torch.cuda.synchronize()
it = (info & 0xffff).float()
print("done", env.backend.launches, "newton iters/env-step mean", float(it.mean()), "max", float(it.max()), "overflow bits", int((info >> 16).max()))
