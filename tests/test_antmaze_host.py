"""AntMaze (BASELINE config 4) host logic + emulated kernel arithmetic on CPU, against the oracle env."""
import numpy as np
import pytest
import torch

import gymnasium_robotics_b200 as pkg
from gymnasium_robotics_b200.maze import MAPS, AntMazeVectorEnv, MazeCells
from gymnasium_robotics_b200.models import load_model
from oracle.ant_maze_env import OracleAntMazeEnv
from oracle.maze import Maze
from tests.hostsim_backend import HostSimBackend


def mk(maze="Large", n=2, **kw):
    return AntMazeVectorEnv(maze, num_envs=n, backend_factory=HostSimBackend, **kw)


def test_registry_and_spaces():
    assert pkg.ENV_IDS["AntMaze_Large-v5"] == dict(maze="Large", reward_type="sparse", max_episode_steps=1000, include_cfrc_ext_in_observation=True)
    assert pkg.ENV_IDS["AntMaze_Large-v4"] == dict(maze="Large", reward_type="sparse", max_episode_steps=1000)
    assert pkg.ENV_IDS["AntMaze_UMazeDense-v5"]["max_episode_steps"] == 700  # reference __init__.py:839-850
    # -v5 wraps Ant-v5: (105,) = 27 + 13 x 6 clipped contact forces (ant_maze_v5.py:99); -v4 wraps Ant-v4: (27,)
    env = pkg.make_vec("AntMaze_Large-v5", num_envs=2, backend_factory=HostSimBackend, rng_mode="numpy")
    assert env.single_action_space.shape == (8,) and env.single_observation_space["observation"].shape == (105,)
    obs, info = env.reset(seed=0)
    assert obs["observation"].shape == (2, 105) and obs["achieved_goal"].shape == (2, 2) and info["success"].shape == (2,)
    assert float(obs["observation"][:, 27:].abs().max()) == 0.0      # reset: cfrc_ext is zero (mj_resetData)
    with pytest.raises(ValueError):
        env.step(np.zeros((2, 4), dtype=np.float32))
    env4 = pkg.make_vec("AntMaze_Large-v4", num_envs=2, backend_factory=HostSimBackend, rng_mode="numpy")
    assert env4.single_observation_space["observation"].shape == (27,)


def test_v5_contact_force_observation_matches_oracle():
    """The 78 trailing entries of the AntMaze-v5 observation: `cfrc_ext[1:]` of the last forward pass clipped to (-1, 1), [torque;
    force] per body about the ant's com.  Unclipped, the forces are hundreds of newtons, so most entries sit at +-1 or 0: the test
    also compares the sign pattern and, with the clip range opened in the oracle, checks that the standing ant's foot forces carry
    its weight."""
    n = 3
    env = pkg.make_vec("AntMaze_UMaze-v5", num_envs=n, backend_factory=HostSimBackend, rng_mode="numpy")
    env.reset(seed=8)
    model = load_model("antmaze_umaze")
    oracles = [OracleAntMazeEnv(MAPS["UMaze"], model=model, include_cfrc_ext_in_observation=True) for _ in range(n)]
    for i, o in enumerate(oracles):
        o.reset(seed=8 + i)
        for _ in range(12):          # the ant is dropped from z = 0.75: let it land before comparing contact forces
            o.step(np.zeros(8))
    lay = env.backend.layout
    rng = np.random.default_rng(3)
    touched = 0
    for step in range(8):
        rec = np.zeros((n, lay["stride"]))
        for i, o in enumerate(oracles):
            rec[i, lay["qpos"]:lay["qpos"] + 15] = o.sim.qpos
            rec[i, lay["qvel"]:lay["qvel"] + 14] = o.sim.qvel
            rec[i, lay["warm"]:lay["warm"] + 14] = o.sim.qacc_warmstart
            rec[i, lay["goal"]:lay["goal"] + 2] = o.goal
        env.set_state(torch.as_tensor(rec, dtype=torch.float32))
        a = rng.uniform(-1, 1, (n, 8)).astype(np.float32) * (0.3 if step < 4 else 1.0)
        o, r, te, tr, info = env.step(a)
        for i, orc in enumerate(oracles):
            oo, *_ = orc.step(a[i].astype(np.float64))
            got, want = o["observation"][i].double().numpy(), oo["observation"]
            assert got.shape == (105,) and want.shape == (105,)
            cf, wf = got[27:].reshape(13, 6), want[27:].reshape(13, 6)
            raw = orc.sim.cfrc_ext()[1:]
            big = np.abs(raw) > 1.5                      # entries the clip saturates in both: must agree exactly
            assert np.array_equal(cf[big], wf[big])
            small = np.abs(raw) < 0.5
            assert np.abs(cf[small] - wf[small]).max() < 2e-2 if small.any() else True
            touched += int((np.abs(raw).sum(axis=1) > 0).sum())
    assert touched >= 12, touched     # the comparison saw real contacts (the ants are still bouncing after the drop)
    env.close()
    # magnitude and sign of the formula: an ant that has settled on its feet is carried by its contact forces
    orc = oracles[0]
    orc.reset(seed=8)
    for _ in range(60):
        orc.step(np.zeros(8))
    raw = orc.sim.cfrc_ext()
    weight = 9.81 * float(np.sum(model.body_mass))
    assert abs(raw[:, 5].sum() - weight) < 0.05 * weight and np.abs(raw[:, 3:5].sum(axis=0)).max() < 0.05 * weight, (raw[:, 3:].sum(axis=0), weight)
    assert np.abs(raw[0]).max() == 0.0      # the world row stays empty (and is left out of the observation)


def test_cell_tables_match_oracle_maze():
    for name, mp in MAPS.items():
        a, b = MazeCells(mp), Maze(mp, 4.0)
        assert np.allclose(a.goal_locations, np.array(b.unique_goal_locations)) and np.allclose(a.reset_locations, np.array(b.unique_reset_locations))
        for i in range(a.length):
            for j in range(a.width):
                assert tuple(a.cell_xy_to_rowcol(a.cell_rowcol_to_xy((i, j)))) == (i, j)  # bit-exact grid indexing


def test_reset_follows_reference_rng_stream_and_never_starts_in_success():
    env = mk("Large", 4, rng_mode="numpy")
    model = load_model("antmaze_large")
    obs, info = env.reset(seed=10)
    for i in range(4):
        orc = OracleAntMazeEnv(MAPS["Large"], model=model)
        oo, oi = orc.reset(seed=10 + i)
        assert np.allclose(obs["desired_goal"][i].double().numpy(), oo["desired_goal"], atol=2e-6)
        assert np.allclose(obs["achieved_goal"][i].double().numpy(), oo["achieved_goal"], atol=2e-6)
        assert np.allclose(obs["observation"][i].double().numpy(), oo["observation"], atol=1e-6)
    # tests/envs/maze/test_ant_maze.py:12-21 of the reference: never reset into a success state
    for s in range(30):
        obs, info = env.reset(seed=1000 + 7 * s)
        d = torch.linalg.norm(obs["achieved_goal"] - obs["desired_goal"], dim=1)
        assert not bool(info["success"].any()) and bool((d > 0.45).all())
    obs, _ = env.reset(seed=42, options={"reset_cell": [1, 2], "goal_cell": [3, 4]})
    assert tuple(env.cells.cell_xy_to_rowcol(obs["achieved_goal"][0].double().numpy())) == (1, 2)
    assert tuple(env.cells.cell_xy_to_rowcol(obs["desired_goal"][0].double().numpy())) == (3, 4)


def test_step_tracks_oracle_from_identical_state():
    env = mk("Large", 2, rng_mode="numpy")
    env.reset(seed=5)
    model = load_model("antmaze_large")
    oracles = [OracleAntMazeEnv(MAPS["Large"], model=model) for _ in range(2)]
    for i, o in enumerate(oracles):
        o.reset(seed=5 + i)
    lay = env.backend.layout
    rng = np.random.default_rng(1)
    errs = []
    for step in range(10):
        rec = np.zeros((2, lay["stride"]))
        for i, o in enumerate(oracles):
            rec[i, lay["qpos"]:lay["qpos"] + 15] = o.sim.qpos
            rec[i, lay["qvel"]:lay["qvel"] + 14] = o.sim.qvel
            rec[i, lay["warm"]:lay["warm"] + 14] = o.sim.qacc_warmstart
            rec[i, lay["goal"]:lay["goal"] + 2] = o.goal
        env.set_state(torch.as_tensor(rec, dtype=torch.float32))
        a = rng.uniform(-1, 1, (2, 8)).astype(np.float32)
        o, r, te, tr, info = env.step(a)
        for i, orc in enumerate(oracles):
            oo, orr, ote, _, oi = orc.step(a[i].astype(np.float64))
            errs.append(np.abs(o["observation"][i].double().numpy() - oo["observation"]).max())
            assert float(r[i]) == float(orr) and bool(info["success"][i]) == oi["success"] and bool(te[i]) == ote
    errs = np.array(errs)
    # one env-step = 5 RK4 sub-steps of a 150 N.m-torque ant: velocities reach tens of rad/s, tolerance is relative to that
    assert np.median(errs) < 2e-4 and np.mean(errs < 2e-3) >= 0.9, errs


def test_dense_reward_episodic_termination_and_timelimit():
    env = mk("UMaze", 1, rng_mode="numpy", reward_type="dense", continuing_task=False, max_episode_steps=3)
    obs, _ = env.reset(seed=0)
    o, r, te, tr, info = env.step(np.zeros((1, 8), dtype=np.float32))
    d = torch.linalg.norm(o["achieved_goal"] - o["desired_goal"], dim=1)
    assert torch.allclose(r, torch.exp(-d)) and not bool(te.any())
    rn = env.compute_reward(o["achieved_goal"].numpy(), o["desired_goal"].numpy(), {})
    assert rn.dtype == np.float64 and np.allclose(rn, r.numpy())
    st, _ = env.get_state()
    lay = env.backend.layout
    st[0, lay["goal"]:lay["goal"] + 2] = st[0, lay["qpos"]:lay["qpos"] + 2]  # put the goal under the ant
    env.set_state(st)
    o, r, te, tr, info = env.step(np.zeros((1, 8), dtype=np.float32))
    assert bool(te.all()) and bool(info["success"].all())  # episodic task terminates on success (maze_v4.py:390-398)
