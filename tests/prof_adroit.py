"""Tiny driver for ncu / compute-sanitizer runs of the Adroit workloads (wide kernel build for hammer / relocate)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gymnasium_robotics_b200 as grb

task = sys.argv[1] if len(sys.argv) > 1 else "AdroitHandHammer-v2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 12
env = grb.make_vec(task, num_envs=n, rng_mode="torch", autoreset_mode="same_step")
env.reset(seed=0)
nact = env.single_action_space.shape[0]
g = torch.Generator(device="cuda").manual_seed(1234)
info = torch.zeros(n, dtype=torch.int32, device="cuda")
out = env.backend.new_outputs()
for k in range(steps):
    a = (torch.rand((n, nact), generator=g, device="cuda") * 2 - 1).contiguous()
    env.backend.step(a, out, info)
torch.cuda.synchronize()
it = (info & 0xffff).float()
print("done", task, env.backend.launches, "newton iters/env-step mean", float(it.mean()), "max", float(it.max()), "overflow bits", int((info >> 16).max()))
