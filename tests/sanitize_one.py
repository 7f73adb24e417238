"""compute-sanitizer driver for ONE env family (a barrier error aborts the launch, so families are checked in separate processes):
    compute-sanitizer --tool synccheck python tests/sanitize_one.py FetchSlide-v4 7 [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gymnasium_robotics_b200 as grb

env_id, n = sys.argv[1], int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
env = grb.make_vec(env_id, num_envs=n, rng_mode="torch")
env.reset(seed=0)
nact = env.single_action_space.shape[0]
for k in range(steps):
    env.step(torch.full((n, nact), -0.7 if k % 2 else 0.4, device="cuda"))
torch.cuda.synchronize()
env.close()
print("sanitize driver done", env_id, n, flush=True)
