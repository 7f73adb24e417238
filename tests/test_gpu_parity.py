"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the C-ABI, against the fp64 CPU
oracle on identical inputs.  Tolerances (fp32 kernel vs fp64 oracle):
  * one env-step (20 sub-steps) from an identical injected state: |obs_gpu - obs_oracle| <= 2e-4 absolute
    (positions in m, scaled velocities, euler angles in rad);
  * rewards / success flags: exact, given the same achieved/desired goals;
  * same-seed runs on the same GPU: bit-identical.
The physics oracle itself is "parity unpinned" w.r.t. MuJoCo (see oracle/oracle.c header).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

OBS_TOL = 2e-4


def _mk(task, n, **kw):
    from gymnasium_robotics_b200.fetch import FetchVectorEnv

    return FetchVectorEnv(task, num_envs=n, device="cuda:0", **kw)


@pytest.mark.parametrize("task", ["FetchReach", "FetchPush", "FetchPickAndPlace"])
def test_env_setup_matches_oracle(task):
    from tests.parity_util import oracle_env_from_model

    env = _mk(task, 4, rng_mode="numpy")
    orc = oracle_env_from_model(task, env.model)
    assert np.allclose(env.initial_gripper_xpos.double().cpu().numpy(), orc.initial_gripper_xpos, atol=1e-4)
    assert np.allclose(env.initial_qpos.double().cpu().numpy(), orc.initial_qpos, atol=2e-4)
    if env.height_offset is not None:
        assert env.height_offset == pytest.approx(orc.height_offset, abs=1e-5)
    env.close()


@pytest.mark.parametrize("task", ["FetchReach", "FetchPush", "FetchPickAndPlace"])
def test_step_parity_from_identical_state(task):
    """Config 1/2 of BASELINE.json: per-step comparison from identical (qpos, qvel, warmstart, ctrl, mocap, goal)."""
    from tests.parity_util import inject_oracle_state, oracle_env_from_model

    n = 8
    env = _mk(task, n, rng_mode="numpy")
    env.reset(seed=100)
    oracles = [oracle_env_from_model(task, env.model) for _ in range(n)]
    rng = np.random.default_rng(7)
    for i, o in enumerate(oracles):
        o.reset(seed=100 + i)
    errs, free_errs = [], []
    for step in range(12):
        inject_oracle_state(env, oracles)
        a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        if step >= 6:  # drive the gripper down towards the table/object to exercise contacts
            a[:, 2] = -1.0
            a[:, 3] = -1.0 if step % 2 else 1.0
        o, r, term, trunc, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            got = o["observation"][i].double().cpu().numpy()
            err = np.abs(got - oo["observation"]).max()
            (free_errs if step < 6 else errs).append(err)
            assert np.isfinite(got).all()
            # reward/success must agree unless the distance sits within tolerance of the threshold
            d = np.linalg.norm(oo["achieved_goal"] - oo["desired_goal"])
            if abs(d - 0.05) > 5e-3:
                assert float(r[i]) == float(orr) and float(info["is_success"][i]) == float(oi["is_success"])
    free_errs, errs = np.array(free_errs), np.array(errs)
    print(f"{task}: free-motion max {free_errs.max():.2e}; contact phase median {np.median(errs):.2e} max {errs.max():.2e}")
    # free motion (object at rest, arm moving): every sample within tolerance
    assert free_errs.max() < OBS_TOL
    # contact-driving phase: impacts amplify fp32/fp64 differences within a single env-step (SURVEY.md 7 "contact-rich
    # chaos"): at least 90% of the samples within tolerance, none wildly off
    assert np.mean(errs < OBS_TOL) >= 0.9 and errs.max() < 0.2
    env.close()


def test_free_running_rollout_tracks_oracle():
    """50 free-running env-steps (no re-injection) on the contact-free FetchReach task stay within 1e-3."""
    from tests.parity_util import oracle_env_from_model

    env = _mk("FetchReach", 2, rng_mode="numpy")
    obs, _ = env.reset(seed=3)
    orc = oracle_env_from_model("FetchReach", env.model)
    oo, _ = orc.reset(seed=3)
    assert np.abs(obs["desired_goal"][0].double().cpu().numpy() - oo["desired_goal"]).max() < 1e-5  # same PCG64 stream (offset by the fp32 rest pose)
    rng = np.random.default_rng(0)
    for _ in range(49):
        a = rng.uniform(-1, 1, (2, 4)).astype(np.float32)
        o, r, *_ = env.step(torch.as_tensor(a))
        oo, *_ = orc.step(a[0].astype(np.float64))
        assert np.abs(o["observation"][0].double().cpu().numpy() - oo["observation"]).max() < 1e-3
    env.close()


def test_reward_and_success_bit_exact_given_goals():
    """fetch_env.py:74-80, 168-170 on 1e5 random pairs: exact equality with the numpy restatement in fp32 inputs."""
    env = _mk("FetchPickAndPlace", 2)
    g = torch.Generator(device="cuda").manual_seed(0)
    ag = torch.rand((100000, 3), generator=g, device="cuda") * 0.2
    dg = torch.rand((100000, 3), generator=g, device="cuda") * 0.2
    r = env.compute_reward(ag, dg, {})
    d = np.sqrt(((ag.cpu().numpy().astype(np.float32) - dg.cpu().numpy().astype(np.float32)) ** 2).sum(-1, dtype=np.float32))
    near = np.abs(d - 0.05) < 1e-6
    want = -(d > np.float32(0.05)).astype(np.float32)
    assert np.array_equal(r.cpu().numpy()[~near], want[~near])
    rn = env.compute_reward(ag.cpu().numpy()[:10], dg.cpu().numpy()[:10], {})
    assert rn.dtype == np.float32 and rn.shape == (10,)
    env.close()


def test_same_seed_determinism_bitwise():
    """tests/test_envs.py:62-117 of the reference: two instances, same seed, same actions => identical outputs."""
    outs = []
    for _ in range(2):
        env = _mk("FetchPickAndPlace", 64, rng_mode="torch")
        env.reset(seed=11)
        g = torch.Generator(device="cuda").manual_seed(5)
        acc = []
        for _ in range(20):
            a = torch.rand((64, 4), generator=g, device="cuda") * 2 - 1
            o, r, te, tr, info = env.step(a)
            acc.append(torch.cat([o["observation"], o["achieved_goal"], o["desired_goal"], r[:, None]], dim=1).clone())
        outs.append(torch.stack(acc))
        env.close()
    assert torch.equal(outs[0], outs[1])


def test_full_size_batch_properties():
    """BASELINE config 2 size (4096 envs): finite outputs, identical envs stay bit-identical (lock-step invariance),
    GoalEnv reward invariant (core.py:61-62), TimeLimit truncation at 50, autoreset restores the initial state."""
    n = 4096
    env = _mk("FetchPickAndPlace", n, rng_mode="torch", autoreset_mode="same_step")
    obs, _ = env.reset(seed=0)
    st, _ = env.get_state()
    st[:] = st[0]  # every env gets env 0's state and goal
    env.set_state(st)
    g = torch.Generator(device="cuda").manual_seed(1)
    for t in range(50):
        a = (torch.rand((1, 4), generator=g, device="cuda") * 2 - 1).expand(n, 4).contiguous()
        o, r, te, tr, info = env.step(a)
        assert torch.isfinite(o["observation"]).all()
        if t < 49:
            assert torch.equal(o["observation"], o["observation"][0:1].expand_as(o["observation"]))
            rr = env.compute_reward(o["achieved_goal"], o["desired_goal"], {})
            assert torch.equal(rr, r)
            assert not bool(tr.any())
    assert bool(tr.all()) and not bool(te.any())
    # same-step autoreset: robot joints back at the initial configuration (tests/test_envs.py:175-231 of the reference)
    st2, el = env.get_state()
    lay = env.backend.layout
    q = st2[:, lay["qpos"]:lay["qpos"] + env.model.nq]
    assert torch.allclose(q[:, :15], env.initial_qpos[:15].expand(n, 15))
    assert int(el.max()) == 0
    env.close()


def test_missing_device_is_loud():
    from gymnasium_robotics_b200 import _lib

    L = _lib.lib()
    import ctypes

    h = ctypes.c_void_p()
    rc = L.b200sim_create(None, 0, None, None, None, 1, 0, ctypes.byref(h))
    assert rc != 0 and b"bad arguments" in L.b200sim_last_error(None)


# ------------------------------------------------------------------------------------------------ AntMaze (config 4)
def _mk_ant(maze, n, **kw):
    from gymnasium_robotics_b200.maze import AntMazeVectorEnv

    return AntMazeVectorEnv(maze, num_envs=n, device="cuda:0", **kw)


def test_antmaze_step_parity_from_identical_state():
    """AntMaze_Large-v5: 5 RK4 sub-steps per env-step, sphere/capsule vs floor and maze walls, from injected states."""
    from gymnasium_robotics_b200.maze import MAPS
    from gymnasium_robotics_b200.models import load_model
    from oracle.ant_maze_env import OracleAntMazeEnv

    n = 8
    env = _mk_ant("Large", n, rng_mode="numpy")
    obs, info = env.reset(seed=20)
    model = load_model("antmaze_large")
    oracles = [OracleAntMazeEnv(MAPS["Large"], model=model) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=20 + i)
        assert np.abs(obs["desired_goal"][i].double().cpu().numpy() - oo["desired_goal"]).max() < 2e-6  # same PCG64 stream
    lay = env.backend.layout
    rng = np.random.default_rng(2)
    errs = []
    for step in range(12):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            rec[i, lay["qpos"]:lay["qpos"] + 15] = o.sim.qpos
            rec[i, lay["qvel"]:lay["qvel"] + 14] = o.sim.qvel
            rec[i, lay["warm"]:lay["warm"] + 14] = o.sim.qacc_warmstart
            rec[i, lay["goal"]:lay["goal"] + 2] = o.goal
        env.set_state(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32)
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, ote, _, oi = orc.step(a[i].astype(np.float64))
            got = o["observation"][i].double().cpu().numpy()
            assert np.isfinite(got).all()
            errs.append(np.abs(got - oo["observation"]).max())
            assert float(r[i]) == float(orr) and bool(info["success"][i]) == oi["success"]
    errs = np.array(errs)
    print(f"AntMaze: median {np.median(errs):.2e} p90 {np.quantile(errs, 0.9):.2e} max {errs.max():.2e}")
    # velocities reach tens of rad/s under +-150 N.m random torques; contact events amplify fp32/fp64 differences
    assert np.median(errs) < 2e-4 and np.mean(errs < 2e-3) >= 0.9 and errs.max() < 1.0


def test_antmaze_shard_size_batch_properties():
    """Config 4 shard (8192 envs over 8 GPUs = 1024 per GPU): finite, lock-step invariance, ant never inside a wall cell
    (bit-exact integer grid indexing), truncation at 1000 handled by the host counter."""
    n = 1024
    env = _mk_ant("Large", n, rng_mode="torch")
    env.reset(seed=0)
    st, _ = env.get_state()
    st[: n // 2] = st[0]
    env.set_state(st)
    g = torch.Generator(device="cuda").manual_seed(3)
    walls = torch.as_tensor(np.array([[c == 1 for c in row] for row in env.cells.maze_map]), device="cuda")
    for t in range(40):
        a = torch.rand((n, 8), generator=g, device="cuda") * 2 - 1
        a[: n // 2] = a[0]
        o, r, te, tr, info = env.step(a)
        assert torch.isfinite(o["observation"]).all()
        assert torch.equal(o["observation"][: n // 2], o["observation"][0:1].expand(n // 2, -1))
        xy = o["achieved_goal"]
        i = torch.floor((env.cells.y_center - xy[:, 1]) / 4.0).long()
        j = torch.floor((xy[:, 0] + env.cells.x_center) / 4.0).long()
        assert not bool(walls[i, j].any())
        assert not bool(te.any()) and not bool(tr.any())
    env.close()


# ------------------------------------------------------------------------------------------------ PointMaze
def test_pointmaze_reference_known_answers_on_gpu():
    """The reference's only numeric known answers (tests/envs/maze/test_point_maze.py:20-45) through the C-ABI on the GPU."""
    import json
    import os

    from gymnasium_robotics_b200.maze import PointMazeVectorEnv
    from gymnasium_robotics_b200.mjcf import Model

    here = os.path.dirname(os.path.abspath(__file__))
    model = Model.from_blob(open(os.path.join(here, "golden", "pointmaze_4x4.b200m"), "rb").read())
    for c in json.load(open(os.path.join(here, "golden", "maze_known_answers.json"))):
        env = PointMazeVectorEnv(c["maze_map"], num_envs=3, model=model, device="cuda:0", rng_mode="numpy")
        obs, info = env.reset(seed=[c["seed"]] * 3, options=c["options"])
        if "reset_pos" in c["expect"]:
            np.testing.assert_almost_equal(np.array(c["expect"]["reset_pos"] + [0, 0]), obs["observation"][0].double().cpu().numpy(), decimal=c["decimal"])
        if "goal" in c["expect"]:
            np.testing.assert_almost_equal(np.array(c["expect"]["goal"]), obs["desired_goal"][2].double().cpu().numpy(), decimal=c["decimal"])
        env.close()


def test_pointmaze_rollout_tracks_oracle():
    from gymnasium_robotics_b200.maze import MAPS, PointMazeVectorEnv
    from gymnasium_robotics_b200.models import load_model
    from oracle.point_maze_env import OraclePointMazeEnv

    env = PointMazeVectorEnv("Medium", num_envs=4, device="cuda:0", rng_mode="numpy")
    obs, _ = env.reset(seed=11)
    model = load_model("pointmaze_medium")
    oracles = [OraclePointMazeEnv(MAPS["Medium"], model) for _ in range(4)]
    for i, o in enumerate(oracles):
        o.reset(seed=11 + i)
    rng = np.random.default_rng(4)
    worst = 0.0
    for _ in range(200):
        a = rng.uniform(-1.5, 1.5, (4, 2)).astype(np.float32)
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, *_ = orc.step(a[i].astype(np.float64))
            worst = max(worst, np.abs(o["observation"][i].double().cpu().numpy() - oo["observation"]).max())
            d = np.linalg.norm(oo["achieved_goal"] - oo["desired_goal"])
            if abs(d - 0.45) > 1e-3:
                assert float(r[i]) == float(orr) and bool(info["success"][i]) == (d <= 0.45)
    print(f"PointMaze free-running 200 steps: worst {worst:.2e}")
    assert worst < 2e-3
    env.close()


# ------------------------------------------------------------------------------------------------ Shadow Hand (config 3)
def _mk_hand(task, n, **kw):
    from gymnasium_robotics_b200.hand import HandVectorEnv

    return HandVectorEnv(task, num_envs=n, device="cuda:0", **kw)


def test_hand_reset_and_step_parity():
    """HandManipulateBlockRotateXYZ-v1: reset (same PCG64 draws, 200 settle sub-steps) and env-steps from injected
    oracle states.  Tolerances: goal rotation 2e-6 (RNG + quaternion algebra only); joint angles / block position 2e-3
    per env-step (contact forces on the 70 g block amplify fp32 geometry round-off)."""
    from gymnasium_robotics_b200.models import load_model
    from oracle.hand_env import OracleHandBlockEnv

    n = 6
    model = load_model("hand_block")
    env = _mk_hand("HandManipulateBlockRotateXYZ", n, rng_mode="numpy")
    obs, _ = env.reset(seed=40)
    oracles = [OracleHandBlockEnv(model=model) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=40 + i)
        g = obs["desired_goal"][i].double().cpu().numpy()
        assert np.abs(g[3:] - oo["desired_goal"][3:]).max() < 2e-6
        a = obs["achieved_goal"][i].double().cpu().numpy()
        assert a[2] > 0.04 and np.abs(a[:3] - oo["achieved_goal"][:3]).max() < 5e-3
    lay, m = env.backend.layout, model
    rng = np.random.default_rng(4)
    errs = []
    for step in range(10):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            s = o.sim
            rec[i, lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
            rec[i, lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
            rec[i, lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
            rec[i, lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
            rec[i, lay["goal"]:lay["goal"] + 7] = o.goal
        env.set_state(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 20)).astype(np.float32)
        info_bits = torch.zeros(n, dtype=torch.int32, device="cuda")
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            got = o["observation"][i].double().cpu().numpy()
            assert np.isfinite(got).all()
            errs.append(max(np.abs(got[:24] - oo["observation"][:24]).max(), np.abs(got[54:57] - oo["observation"][54:57]).max()))
            assert float(r[i]) == float(orr) and float(info["is_success"][i]) == float(oi["is_success"])
    errs = np.array(errs)
    print(f"Hand: median {np.median(errs):.2e} p90 {np.quantile(errs, 0.9):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 2e-4 and np.mean(errs < 2e-3) >= 0.9 and errs.max() < 0.1
    env.close()


def test_hand_full_size_batch_properties():
    """BASELINE config 3 size (2048 envs): finite outputs, unit block quaternions, lock-step invariance of identical envs,
    GoalEnv reward invariant, no capacity overflow flags, TimeLimit at 100 + same-step autoreset puts the block back on
    the palm with the hand at its initial configuration before the settle phase."""
    n = 2048
    env = _mk_hand("HandManipulateBlockRotateXYZ", n, rng_mode="torch", autoreset_mode="same_step")
    obs, _ = env.reset(seed=0)
    assert bool((obs["achieved_goal"][:, 2] > 0.04).all())
    st, _ = env.get_state()
    st[: n // 2] = st[0]
    env.set_state(st)
    g = torch.Generator(device="cuda").manual_seed(5)
    out = env.backend.new_outputs()
    info_bits = torch.zeros(n, dtype=torch.int32, device="cuda")
    for t in range(100):
        a = torch.rand((n, 20), generator=g, device="cuda") * 2 - 1
        a[: n // 2] = a[0]
        if t == 50:  # one raw backend step with the info word: Newton iterations and overflow flags
            env.backend.step(a.contiguous(), out, info_bits)
            env._elapsed += 1
            env._elapsed_ub += 1
            assert int((info_bits >> 16).max()) == 0, "contact / row capacity overflow"
            continue
        o, r, te, tr, info = env.step(a)
        assert torch.isfinite(o["observation"]).all()
        if t < 99:
            assert torch.equal(o["observation"][: n // 2], o["observation"][0:1].expand(n // 2, -1))
            assert torch.equal(env.compute_reward(o["achieved_goal"], o["desired_goal"], {}), r)
            qn = torch.linalg.norm(o["achieved_goal"][:, 3:], dim=1)
            assert bool(((qn - 1).abs() < 1e-4).all())
            assert not bool(tr.any())
    assert bool(tr.all()) and not bool(te.any())
    assert bool((o["achieved_goal"][:, 2] > 0.04).all())   # after the same-step reset
    assert int(env._elapsed.max()) == 0
    env.close()


def test_hand_touch_sensors_parity():
    """HandManipulateBlockRotateXYZ_ContinuousTouchSensors-v1 (BASELINE config 3: 24 DoF + 92 touch sensors): the same
    sensors fire as in the oracle after reset and after env-steps from injected states; forces within 5 % of the scale."""
    import gymnasium_robotics_b200 as pkg
    from gymnasium_robotics_b200.models import load_model
    from oracle.hand_env import OracleHandBlockEnv

    n = 4
    model = load_model("hand_block_touch")
    env = pkg.make_vec("HandManipulateBlockRotateXYZ_ContinuousTouchSensors-v1", num_envs=n, device="cuda:0", rng_mode="numpy")
    assert env.single_observation_space["observation"].shape == (153,)
    obs, _ = env.reset(seed=60)
    oracles = [OracleHandBlockEnv(model=model, touch_get_obs="sensordata") for _ in range(n)]
    lay, m = env.backend.layout, model
    rng = np.random.default_rng(6)
    agree, total = 0, 0
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=60 + i)
        t, ot = obs["observation"][i, 61:].double().cpu().numpy(), oo["observation"][61:]
        assert ot.sum() > 0.3 and abs(t.sum() - ot.sum()) < 0.1 * ot.sum()   # the block's weight is carried by the hand
    for step in range(6):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            s = o.sim
            rec[i, lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
            rec[i, lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
            rec[i, lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
            rec[i, lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
            rec[i, lay["goal"]:lay["goal"] + 7] = o.goal
        env.set_state(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 20)).astype(np.float32)
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, *_ = orc.step(a[i].astype(np.float64))
            t, ot = o["observation"][i, 61:].double().cpu().numpy(), oo["observation"][61:]
            assert np.isfinite(t).all() and (t >= 0).all()
            total += 1
            agree += int(np.abs(t - ot).max() <= 0.05 * max(1.0, ot.max()))
    print(f"Hand touch: {agree}/{total} env-steps with all 92 sensors within 5 % of the force scale")
    assert agree >= 0.8 * total
    env.close()


def test_hand_reach_parity():
    """HandReach-v3 on the GPU: same goals as the oracle for the same seeds, env-steps from injected states within 2e-4 on
    the fingertip positions (achieved goal), rewards / success exact."""
    import gymnasium_robotics_b200 as pkg
    from gymnasium_robotics_b200.models import load_model
    from oracle.hand_env import OracleHandReachEnv

    n = 4
    model = load_model("hand_reach")
    env = pkg.make_vec("HandReach-v3", num_envs=n, device="cuda:0", rng_mode="numpy")
    obs, _ = env.reset(seed=70)
    oracles = [OracleHandReachEnv(model=model) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=70 + i)
        assert np.abs(obs["desired_goal"][i].double().cpu().numpy() - oo["desired_goal"]).max() < 2e-6
        assert np.abs(obs["observation"][i].double().cpu().numpy() - oo["observation"]).max() < 2e-6
    lay, m = env.backend.layout, model
    rng = np.random.default_rng(7)
    errs = []
    for step in range(8):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            s = o.sim
            rec[i, lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
            rec[i, lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
            rec[i, lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
            rec[i, lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
            rec[i, lay["goal"]:lay["goal"] + 15] = o.goal
        env.set_state(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 20)).astype(np.float32)
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            errs.append(np.abs(o["achieved_goal"][i].double().cpu().numpy() - oo["achieved_goal"]).max())
            assert float(r[i]) == float(orr) and float(info["is_success"][i]) == float(oi["is_success"])
    errs = np.array(errs)
    print(f"HandReach: median {np.median(errs):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 2e-5 and errs.max() < 2e-4
    env.close()


# ------------------------------------------------------------------------------ general convex collider (cylinder, ellipsoid)
def test_fetch_slide_parity():
    """FetchSlide-v4: the cylinder puck runs through the portal-refinement collider (kernel build NVP = 22).  A flat
    cylinder rocks on its single portal contact, so its Euler angles (obs 11:14) and angular velocity (obs 17:20) are
    chaotic between fp32 and fp64; the other 19 observation entries are compared per env-step from injected states."""
    from tests.parity_util import inject_oracle_state, oracle_env_from_model

    n = 6
    env = _mk("FetchSlide", n, rng_mode="numpy")
    orc0 = oracle_env_from_model("FetchSlide", env.model)
    assert np.allclose(env.initial_gripper_xpos.double().cpu().numpy(), orc0.initial_gripper_xpos, atol=1e-4)
    assert env.height_offset == pytest.approx(orc0.height_offset, abs=5e-4)
    obs, _ = env.reset(seed=100)
    oracles = [oracle_env_from_model("FetchSlide", env.model) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=100 + i)
        assert np.abs(obs["desired_goal"][i].double().cpu().numpy() - oo["desired_goal"]).max() < 5e-4
    keep = np.array([i for i in range(25) if not 11 <= i < 14 and not 17 <= i < 20])
    rng = np.random.default_rng(7)
    errs = []
    for step in range(10):
        inject_oracle_state(env, oracles)
        a = rng.uniform(-1, 1, (n, 4)).astype(np.float32)
        if step >= 5:   # sweep the gripper across the table towards the puck
            a[:, 2] = -0.3
        o, r, term, trunc, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            got = o["observation"][i].double().cpu().numpy()
            assert np.isfinite(got).all()
            errs.append(np.abs(got - oo["observation"])[keep].max())
            d = np.linalg.norm(oo["achieved_goal"] - oo["desired_goal"])
            if abs(d - 0.05) > 5e-3:
                assert float(r[i]) == float(orr)
    errs = np.array(errs)
    print(f"FetchSlide: median {np.median(errs):.2e} max {errs.max():.2e}")
    assert np.median(errs) < OBS_TOL and np.mean(errs < 1e-3) >= 0.9 and errs.max() < 0.05
    # batch properties at the bench size: no capacity overflow, pucks stay on the table under random actions
    env.close()
    env = _mk("FetchSlide", 4096, rng_mode="torch")
    env.reset(seed=1)
    g = torch.Generator(device="cuda").manual_seed(2)
    info_bits = torch.zeros(4096, dtype=torch.int32, device="cuda")
    for _ in range(10):
        a = torch.rand((4096, 4), generator=g, device="cuda") * 2 - 1
        out = env.backend.new_outputs()
        env.backend.step(a, out, info_bits)
        assert torch.isfinite(out["obs"]).all()
        assert int((info_bits >> 16).max()) == 0
    assert float(out["obs"][:, 5].min()) > 0.35
    env.close()


def test_hand_egg_parity():
    """HandManipulateEggRotate-v1: the ellipsoid egg against the palm/finger capsules and boxes through the convex collider."""
    from gymnasium_robotics_b200.models import load_model
    from oracle.hand_env import OracleHandBlockEnv

    n = 4
    model = load_model("hand_egg")
    env = _mk_hand("HandManipulateEggRotate", n, rng_mode="numpy")
    obs, _ = env.reset(seed=40)
    oracles = [OracleHandBlockEnv(model=model) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=40 + i)
        g = obs["desired_goal"][i].double().cpu().numpy()
        assert np.abs(g[3:] - oo["desired_goal"][3:]).max() < 2e-6
        a = obs["achieved_goal"][i].double().cpu().numpy()
        assert a[2] > 0.04 and np.abs(a[:3] - oo["achieved_goal"][:3]).max() < 5e-3
    lay, m = env.backend.layout, model
    rng = np.random.default_rng(4)
    errs = []
    for step in range(8):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            s = o.sim
            rec[i, lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
            rec[i, lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
            rec[i, lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
            rec[i, lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
            rec[i, lay["goal"]:lay["goal"] + 7] = o.goal
        env.set_state(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 20)).astype(np.float32)
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            got = o["observation"][i].double().cpu().numpy()
            errs.append(max(np.abs(got[:24] - oo["observation"][:24]).max(), np.abs(got[54:57] - oo["observation"][54:57]).max()))
            assert float(r[i]) == float(orr)
    errs = np.array(errs)
    print(f"HandEgg: median {np.median(errs):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 2e-4 and np.mean(errs < 2e-3) >= 0.9
    env.close()


# ------------------------------------------------------------------------------------------------ Adroit hammer (config 5a)
def test_adroit_hammer_parity():
    """AdroitHandHammer-v2 on the wide kernel build (33 dofs: 64-bit dof masks, bordered register Cholesky): reset from the
    same PCG64 draw, env-steps from injected oracle states (free motion, then the arm lowered onto the hammer), rewards."""
    import gymnasium_robotics_b200 as pkg
    from gymnasium_robotics_b200.models import load_model
    from oracle.adroit_env import OracleAdroitHammerEnv

    n = 4
    m = load_model("adroit_hammer")
    env = pkg.make_vec("AdroitHandHammer-v2", num_envs=n, device="cuda:0", rng_mode="numpy")
    obs, _ = env.reset(seed=30)
    oracles = [OracleAdroitHammerEnv(m) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=30 + i)
        assert np.abs(obs[i].double().cpu().numpy() - oo).max() < 2e-6
    lay = env.backend.layout
    rng = np.random.default_rng(2)
    errs = []
    for step in range(12):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            s = o.sim
            rec[i, lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
            rec[i, lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
            rec[i, lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
            rec[i, lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
            rec[i, lay["penv"]:lay["penv"] + 3] = s.body_pos[o.target_body_id]
            rec[i, lay["penv"] + 3:lay["penv"] + 7] = np.asarray(m.body_quat).reshape(-1, 4)[o.target_body_id]
        env.backend.state.copy_(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 26)).astype(np.float32)
        if step >= 5:
            a[:, :2] = [-1, -0.5]   # lower the arm onto the hammer
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            got = o[i].double().cpu().numpy()
            assert np.isfinite(got).all()
            errs.append(np.abs(got - oo).max())
            assert abs(float(r[i]) - orr) < 1e-3 and bool(info["success"][i]) == bool(oi["success"])
    errs = np.array(errs)
    print(f"AdroitHammer: median {np.median(errs):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 2e-5 and np.mean(errs < 2e-4) >= 0.9 and errs.max() < 0.05
    env.close()
    # bench-size batch: finite, no capacity overflow, board heights inside the sampled range
    env = pkg.make_vec("AdroitHandHammer-v2", num_envs=2048, device="cuda:0", rng_mode="torch")
    env.reset(seed=1)
    z = env.get_env_state()["board_pos"][:, 2]
    assert float(z.min()) >= 0.1 and float(z.max()) <= 0.25
    g = torch.Generator(device="cuda").manual_seed(2)
    info_bits = torch.zeros(2048, dtype=torch.int32, device="cuda")
    for _ in range(10):
        a = torch.rand((2048, 26), generator=g, device="cuda") * 2 - 1
        out = env.backend.new_outputs()
        env.backend.step(a, out, info_bits)
        assert torch.isfinite(out["obs"]).all()
        assert int((info_bits >> 16).max()) == 0, "contact / row capacity overflow"
    env.close()


def test_adroit_relocate_parity():
    """AdroitHandRelocate-v2 (36 dofs: four border rows in the register Cholesky)."""
    import gymnasium_robotics_b200 as pkg
    from gymnasium_robotics_b200.models import load_model
    from oracle.adroit_env import OracleAdroitRelocateEnv

    n = 4
    m = load_model("adroit_relocate")
    env = pkg.make_vec("AdroitHandRelocate-v2", num_envs=n, device="cuda:0", rng_mode="numpy")
    obs, _ = env.reset(seed=30)
    oracles = [OracleAdroitRelocateEnv(m) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=30 + i)
        assert np.abs(obs[i].double().cpu().numpy() - oo).max() < 2e-6
    lay = env.backend.layout
    rng = np.random.default_rng(2)
    errs = []
    for step in range(10):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            s = o.sim
            rec[i, lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
            rec[i, lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
            rec[i, lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
            rec[i, lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
            rec[i, lay["penv"]:lay["penv"] + 3] = s.body_pos[o.obj_body_id]
            rec[i, lay["penv"] + 3:lay["penv"] + 7] = np.asarray(m.body_quat).reshape(-1, 4)[o.obj_body_id]
            rec[i, lay["goal"]:lay["goal"] + 3] = o.target_pos
        env.backend.state.copy_(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 30)).astype(np.float32)
        if step >= 4:
            a[:, :3] = [0, 0.5, -1]
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            got = o[i].double().cpu().numpy()
            assert np.isfinite(got).all()
            errs.append(np.abs(got - oo).max())
            assert abs(float(r[i]) - orr) < 1e-3 and bool(info["success"][i]) == bool(oi["success"])
    errs = np.array(errs)
    print(f"AdroitRelocate: median {np.median(errs):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 2e-5 and np.mean(errs < 2e-4) >= 0.9 and errs.max() < 0.05
    env.close()
    # bench-size batch under random actions: the arm presses the whole hand onto the table in some envs, which exceeds the
    # 13 geom-pair groups kept per env -- flagged in the info word, never a stale contact record (regression: counted but
    # unwritten records used to be finalised from stale words)
    env = pkg.make_vec("AdroitHandRelocate-v2", num_envs=2048, device="cuda:0", rng_mode="torch")
    env.reset(seed=1)
    g = torch.Generator(device="cuda").manual_seed(1234)
    info_bits = torch.zeros(2048, dtype=torch.int32, device="cuda")
    flagged = 0
    for _ in range(60):
        a = torch.rand((2048, 30), generator=g, device="cuda") * 2 - 1
        out = env.backend.new_outputs()
        env.backend.step(a, out, info_bits)
        assert torch.isfinite(out["obs"]).all()
        flagged += int(((info_bits >> 16) != 0).sum())
    print(f"AdroitRelocate 2048 x 60: {flagged} env-steps with a capacity flag")
    assert flagged < 0.01 * 2048 * 60
    env.close()


def test_adroit_pen_parity():
    """AdroitHandPen-v2 (30 dofs, cylinder pen through the portal-refinement collider, per-env target quaternion)."""
    import gymnasium_robotics_b200 as pkg
    from gymnasium_robotics_b200.models import load_model
    from oracle.adroit_env import OracleAdroitPenEnv

    n = 4
    m = load_model("adroit_pen")
    env = pkg.make_vec("AdroitHandPen-v2", num_envs=n, device="cuda:0", rng_mode="numpy")
    obs, _ = env.reset(seed=30)
    oracles = [OracleAdroitPenEnv(m) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=30 + i)
        assert np.abs(obs[i].double().cpu().numpy() - oo).max() < 5e-6
    lay = env.backend.layout
    rng = np.random.default_rng(2)
    errs = []
    for step in range(10):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            s = o.sim
            rec[i, lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
            rec[i, lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
            rec[i, lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
            rec[i, lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
            rec[i, lay["penv"]:lay["penv"] + 3] = s.body_pos[o.target_obj_body_id]
            rec[i, lay["penv"] + 3:lay["penv"] + 7] = s.body_quat[o.target_obj_body_id]
        env.backend.state.copy_(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 24)).astype(np.float32)
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            got = o[i].double().cpu().numpy()
            assert np.isfinite(got).all()
            d = np.abs(got - oo)
            errs.append(np.delete(d, [30, 31, 32]).max())
            assert d.max() < 0.05 and abs(float(r[i]) - orr) < 2e-3
    errs = np.array(errs)
    print(f"AdroitPen: median {np.median(errs):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 2e-4 and np.mean(errs < 1e-3) >= 0.9
    env.close()


def test_adroit_door_parity():
    """AdroitHandDoor-v2 (30 dofs, 278 candidate pairs -> two-byte candidates, per-env door frame position)."""
    import gymnasium_robotics_b200 as pkg
    from gymnasium_robotics_b200.models import load_model
    from oracle.adroit_env import OracleAdroitDoorEnv

    n = 4
    m = load_model("adroit_door")
    env = pkg.make_vec("AdroitHandDoor-v2", num_envs=n, device="cuda:0", rng_mode="numpy")
    obs, _ = env.reset(seed=30)
    oracles = [OracleAdroitDoorEnv(m) for _ in range(n)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=30 + i)
        assert np.abs(obs[i].double().cpu().numpy() - oo).max() < 5e-6
    lay = env.backend.layout
    rng = np.random.default_rng(2)
    errs = []
    for step in range(10):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            s = o.sim
            rec[i, lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
            rec[i, lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
            rec[i, lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
            rec[i, lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
            rec[i, lay["penv"]:lay["penv"] + 3] = s.body_pos[o.door_body_id]
            rec[i, lay["penv"] + 3:lay["penv"] + 7] = s.body_quat[o.door_body_id]
        env.backend.state.copy_(torch.as_tensor(rec, dtype=torch.float32, device="cuda"))
        a = rng.uniform(-1, 1, (n, 28)).astype(np.float32)
        if step >= 4:
            a[:, 0] = 1.0
        o, r, te, tr, info = env.step(torch.as_tensor(a))
        for i, orc in enumerate(oracles):
            oo, orr, _, _, oi = orc.step(a[i].astype(np.float64))
            got = o[i].double().cpu().numpy()
            assert np.isfinite(got).all()
            errs.append(np.abs(got - oo).max())
            assert abs(float(r[i]) - orr) < 2e-3 and bool(info["success"][i]) == bool(oi["success"])
    errs = np.array(errs)
    print(f"AdroitDoor: median {np.median(errs):.2e} max {errs.max():.2e}")
    assert np.median(errs) < 2e-5 and np.mean(errs < 1e-3) >= 0.9 and errs.max() < 0.05
    env.close()
