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


def test_uniform_slot_resets_of_the_adroit_envs():
    """b200sim_reset_uniform through the vector envs: Hammer (wide build) and Relocate draw per-env model poses on the device."""
    import gymnasium_robotics_b200 as pkg
    from tests.test_reset_device import philox4x32_10, u01

    n, seed = 96, 11
    env = pkg.make_vec("AdroitHandRelocate-v2", num_envs=n, rng_mode="device")
    env.reset(seed=seed)
    s = env.get_env_state()
    obj, tgt = s["obj_pos"].cpu().numpy(), s["target_pos"].cpu().numpy()
    lo, hi = np.float32([-0.15, -0.15, -0.2, -0.2, 0.15]), np.float32([0.15, 0.3, 0.2, 0.2, 0.35])
    for i in range(n):
        r = [philox4x32_10((i, 0, b, 0x0A11), (seed, 0)) for b in range(2)]
        want = [lo[k] + (hi[k] - lo[k]) * u01(r[k // 4][k % 4]) for k in range(5)]
        got = [obj[i, 0], obj[i, 1], tgt[i, 0], tgt[i, 1], tgt[i, 2]]
        assert np.abs(np.float32(got) - np.float32(want)).max() < 1e-6, i
    assert torch.equal(s["qpos"], env.init_qpos.expand(n, -1))
    env.close()
    env = pkg.make_vec("AdroitHandHammer-v2", num_envs=n, rng_mode="device", max_episode_steps=4, autoreset_mode="same_step")
    env.reset(seed=seed)
    z0 = env.get_env_state()["board_pos"][:, 2].clone()
    assert float(z0.min()) >= 0.1 and float(z0.max()) <= 0.25 and float(z0.std()) > 0.02
    for _ in range(4):
        o, r, te, tr, info = env.step(torch.zeros((n, 26), device="cuda"))
    assert bool(tr.all()) and torch.isfinite(o).all()
    z1 = env.get_env_state()["board_pos"][:, 2]
    assert not torch.equal(z0, z1) and float(z1.min()) >= 0.1 and float(z1.max()) <= 0.25
    env.close()


def test_maze_resets_in_kernel():
    """b200sim_reset_maze on the GPU against the Python restatement on the same Philox numbers (AntMaze_Large, BASELINE config 4)."""
    import gymnasium_robotics_b200 as pkg
    from tests.test_reset_device import py_maze_draw

    n, seed = 128, 17
    env = pkg.make_vec("AntMaze_Large-v5", num_envs=n, rng_mode="device", env_offset=1024)
    o, _ = env.reset(seed=seed)
    st, _ = env.get_state()
    lay = env.backend.layout
    goal_xy, reset_xy = env._goal_loc.cpu().numpy(), env._reset_loc.cpu().numpy()
    q = st[:, lay["qpos"]:lay["qpos"] + 2].cpu().numpy()
    g = o["desired_goal"].cpu().numpy()
    for i in range(n):
        wg, wp = py_maze_draw(goal_xy, reset_xy, env.scaling, 0.25, seed, 1024 + i, 0)
        assert np.abs(g[i] - wg).max() < 5e-6 and np.abs(q[i] - wp).max() < 5e-6, i   # coordinates up to 22: one fp32 ulp (FMA contraction) is 1.9e-6
    d = np.linalg.norm(q - g, axis=1)
    assert d.min() > 0.0 and torch.isfinite(o["observation"]).all()
    for _ in range(3):
        o, r, te, tr, info = env.step(torch.zeros((n, 8), device="cuda"))
    assert torch.isfinite(o["observation"]).all()
    env.close()


def test_check_state_recovers_bad_envs_on_the_gpu():
    """b200sim_check_state: NaN / huge entries injected into two state records are detected, those envs go back to the rest record
    with their goals kept, the others are untouched (opt-in `auto_recover=True`)."""
    import gymnasium_robotics_b200 as pkg

    n = 64
    env = pkg.make_vec("FetchPickAndPlace-v4", num_envs=n, rng_mode="torch", auto_recover=True)
    obs, _ = env.reset(seed=2)
    goals = obs["desired_goal"].clone()
    env.step(torch.zeros((n, 4), device="cuda"))
    st = env.backend.state
    st[5, env._sl["qvel"].start + 2] = float("nan")
    st[40, env._sl["qpos"].start] = -4e10
    o, r, te, tr, info = env.step(torch.full((n, 4), 0.3, device="cuda"))
    bad = info["bad_state"].cpu()
    assert bad.nonzero().flatten().tolist() == [5, 40] and int(env.bad_state_count) == 2
    assert torch.isfinite(o["observation"]).all() and torch.equal(o["desired_goal"], goals)
    o, r, te, tr, info = env.step(torch.zeros((n, 4), device="cuda"))
    assert not info["bad_state"].any() and torch.isfinite(o["observation"]).all()
    env.close()


def test_hand_resets_in_kernel():
    """b200sim_reset_hand_pose / _goal through HandVectorEnv (BASELINE config 3's model): every object ends on the palm, goals are unit
    quaternions drawn per env, episodes advance the counters, same seed => same reset."""
    import gymnasium_robotics_b200 as pkg

    n = 128
    env = pkg.make_vec("HandManipulateBlockRotateXYZ-v1", num_envs=n, rng_mode="device", max_episode_steps=5)
    o, _ = env.reset(seed=4)
    st, _ = env.get_state()
    assert bool((st[:, env._obj.start + 2] > 0.04).all()) and torch.isfinite(o["observation"]).all()
    g = o["desired_goal"]
    assert float((g[:, 3:].norm(dim=1) - 1).abs().max()) < 1e-5 and torch.unique(g[:, 3:].round(decimals=5), dim=0).shape[0] == n
    assert torch.equal(g[:, :3], st[:, env._obj][:, :3])              # RotateXYZ ignores the position: goal position = settled object
    env2 = pkg.make_vec("HandManipulateBlockRotateXYZ-v1", num_envs=n, rng_mode="device", max_episode_steps=5)
    o2, _ = env2.reset(seed=4)
    assert torch.equal(o["desired_goal"], o2["desired_goal"]) and torch.equal(o["observation"], o2["observation"])
    for _ in range(6):                                                # TimeLimit 5 + the next-step reset call
        o, r, te, tr, info = env.step(torch.zeros((n, 20), device="cuda"))
    assert int(env._episode.min()) == 2 and not torch.equal(o["desired_goal"], g) and torch.isfinite(o["observation"]).all()
    env.close(); env2.close()
    reach = pkg.make_vec("HandReach-v3", num_envs=64, rng_mode="device")
    o, _ = reach.reset(seed=1)
    init = reach.initial_goal.reshape(1, 15)
    moved = (o["desired_goal"] - init).abs().amax(dim=1) > 0
    assert torch.isfinite(o["observation"]).all() and 40 <= int(moved.sum()) <= 64       # ~10 % keep the initial finger tips (reach.py:118-120)
    assert float((o["desired_goal"] - init).abs().max()) < 0.2
    reach.close()
