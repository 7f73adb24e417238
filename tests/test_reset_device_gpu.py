"""b200sim_reset on the GPU: the in-kernel draws equal the Python restatement on the same Philox numbers (tests/test_reset_device.py),
masked resets touch only their envs, and the vector env runs episodes in rng_mode="device"."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_in_kernel_reset_matches_the_python_restatement():
    import gymnasium_robotics_b200 as pkg
    from tests.test_reset_device import py_fetch_draw

    n, seed = 192, 0x0123456789ABCDEF
    env = pkg.make_vec("FetchPickAndPlace-v4", num_envs=n, rng_mode="device")
    obs, _ = env.reset(seed=seed)
    p, rest = env._dev_reset
    st, _ = env.get_state()
    lay = env.backend.layout
    q = st[:, lay["qpos"]:lay["qpos"] + 22].cpu().numpy()
    goal = st[:, lay["goal"]:lay["goal"] + 3].cpu().numpy()
    for i in range(n):
        xy, g = py_fetch_draw(p, seed, i, 0)
        assert np.abs(q[i, 15:17] - xy).max() < 1e-6 and np.abs(goal[i] - g).max() < 1e-6, i
    assert np.array_equal(obs["desired_goal"].cpu().numpy(), goal)
    assert torch.equal(st[:, lay["qpos"]:lay["qpos"] + 15], env.initial_qpos[:15].expand(n, 15))
    assert int(env._episode.min()) == 1 and int(env._episode.max()) == 1
    # masked reset: only the masked envs get a new record (episode 1 draws), the others keep theirs bit for bit
    mask = torch.zeros(n, dtype=torch.bool, device="cuda")
    mask[::3] = True
    out = env.backend.new_outputs()
    env._reset_envs(mask, out)
    st2, _ = env.get_state()
    assert torch.equal(st2[~mask], st[~mask])
    g2 = st2[:, lay["goal"]:lay["goal"] + 3].cpu().numpy()
    for i in range(0, n, 3):
        _, g = py_fetch_draw(p, seed, i, 1)
        assert np.abs(g2[i] - g).max() < 1e-6
    assert torch.equal(env._episode, 1 + mask.to(torch.int32))
    env.close()


def test_episodes_with_in_kernel_resets():
    import gymnasium_robotics_b200 as pkg

    n = 256
    env = pkg.make_vec("FetchPush-v4", num_envs=n, rng_mode="device", max_episode_steps=20, autoreset_mode="same_step")
    o, _ = env.reset(seed=3)
    g0 = o["desired_goal"].clone()
    gen = torch.Generator(device="cuda").manual_seed(0)
    launches0 = env.backend.launches
    for k in range(45):
        o, r, te, tr, info = env.step(torch.rand((n, 4), device="cuda", generator=gen) * 2 - 1)
        assert bool(tr.all()) == (k % 20 == 19)
    assert torch.isfinite(o["observation"]).all()
    assert not torch.equal(o["desired_goal"], g0) and int(env._episode.min()) == 3
    # 45 step launches + 2 resets x (draw kernel + refresh)
    assert env.backend.launches - launches0 == 45 + 4
    env.close()
