"""Trajectory dump through the CUDA path: the recorder's buffers live on the device, episodes are cut on the host afterwards."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_recorder_on_the_cuda_path(tmp_path):
    import gymnasium_robotics_b200 as pkg
    from gymnasium_robotics_b200.rollout import RolloutRecorder, load_rollout

    n, T = 64, 10
    env = pkg.make_vec("FetchPickAndPlace-v4", num_envs=n, max_episode_steps=T, rng_mode="torch")
    rec = RolloutRecorder(env, capacity_steps=40, env_id="FetchPickAndPlace-v4")
    rec.reset(seed=0)
    g = torch.Generator(device="cuda").manual_seed(5)
    for _ in range(35):
        a = torch.rand((n, 4), device="cuda", generator=g) * 2 - 1
        rec.step(a)
    assert rec._buf["obs:observation"].is_cuda
    meta = rec.save(str(tmp_path / "pnp"))
    # next_step autoreset: 35 calls = 10 + (1 + 10) + (1 + 10) + (1 + 2) per env
    assert meta["total_episodes"] == 4 * n and meta["total_steps"] == 32 * n
    _, eps = load_rollout(str(tmp_path / "pnp"))
    for e in eps[:8]:
        assert e["observations"]["observation"].shape[1] == 25 and np.isfinite(e["observations"]["observation"]).all()
        r = env.compute_reward(e["observations"]["achieved_goal"][1:], e["observations"]["desired_goal"][1:], {})
        assert np.array_equal(r, e["rewards"])
    env.close()
