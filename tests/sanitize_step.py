"""Small driver for compute-sanitizer runs (memcheck / racecheck) on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gymnasium_robotics_b200.fetch import FetchVectorEnv
env = FetchVectorEnv("FetchPickAndPlace", num_envs=30, rng_mode="torch")
env.reset(seed=0)
for k in range(2):
    a = torch.full((30, 4), -1.0 if k else 0.3, device="cuda")
    env.step(a)
torch.cuda.synchronize()
print("sanitize driver done")
