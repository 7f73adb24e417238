"""compute-sanitizer driver (racecheck / memcheck): one small batch of each kernel build, two env-steps each."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gymnasium_robotics_b200 as grb

for env_id, n in (("FetchPickAndPlace-v4", 30), ("FetchSlide-v4", 10), ("HandManipulateBlockRotateXYZ_ContinuousTouchSensors-v1", 16),
                  ("AdroitHandHammer-v2", 16), ("AntMaze_UMaze-v5", 9), ("FrankaKitchen-v1", 9)):
    env = grb.make_vec(env_id, num_envs=n, rng_mode="torch")
    env.reset(seed=0)
    nact = env.single_action_space.shape[0]
    for k in range(2):
        env.step(torch.full((n, nact), -0.7 if k else 0.4, device="cuda"))
    torch.cuda.synchronize()
    env.close()
    print("sanitize driver done", env_id, flush=True)

# the reset-draw kernels (b200sim_reset / _uniform / _maze) and the state scan (b200sim_check_state)
for env_id, n, kw in (("FetchPickAndPlace-v4", 33, dict(auto_recover=True)), ("AdroitHandPen-v2", 9, {}), ("AdroitHandRelocate-v2", 9, {}),
                      ("PointMaze_Medium-v3", 17, {}), ("HandManipulateBlockRotateParallel-v1", 9, {})):
    env = grb.make_vec(env_id, num_envs=n, rng_mode="device", max_episode_steps=2, **kw)
    env.reset(seed=1)
    nact = env.single_action_space.shape[0]
    for k in range(4):          # two episodes: next-step autoreset calls the draw kernel with a mask
        env.step(torch.full((n, nact), 0.2, device="cuda"))
    torch.cuda.synchronize()
    env.close()
    print("sanitize driver done (device resets)", env_id, flush=True)
