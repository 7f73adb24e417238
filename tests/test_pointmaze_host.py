"""PointMaze-v3: the reference's own known-answer vectors reproduced by the PRODUCT env class (host logic + emulated
kernel), plus per-step parity with the oracle env."""
import json
import os

import numpy as np
import torch

import gymnasium_robotics_b200 as pkg
from gymnasium_robotics_b200.maze import MAPS, PointMazeVectorEnv
from gymnasium_robotics_b200.mjcf import Model
from gymnasium_robotics_b200.models import load_model
from oracle.point_maze_env import OraclePointMazeEnv
from tests.hostsim_backend import HostSimBackend

HERE = os.path.dirname(os.path.abspath(__file__))


def test_reference_known_answers_through_the_product_env():
    """tests/envs/maze/test_point_maze.py:20-45 of the reference, unchanged inputs and expected values."""
    model = Model.from_blob(open(os.path.join(HERE, "golden", "pointmaze_4x4.b200m"), "rb").read())
    for c in json.load(open(os.path.join(HERE, "golden", "maze_known_answers.json"))):
        env = PointMazeVectorEnv(c["maze_map"], num_envs=1, model=model, backend_factory=HostSimBackend, rng_mode="numpy")
        obs, info = env.reset(seed=c["seed"], options=c["options"])
        if "reset_pos" in c["expect"]:
            desired_obs = np.array(c["expect"]["reset_pos"] + [0, 0])
            np.testing.assert_almost_equal(desired_obs, obs["observation"][0].double().numpy(), decimal=c["decimal"])
        if "goal" in c["expect"]:
            np.testing.assert_almost_equal(np.array(c["expect"]["goal"]), obs["desired_goal"][0].double().numpy(), decimal=c["decimal"])


def test_registry():
    assert pkg.ENV_IDS["PointMaze_UMaze-v3"] == dict(maze="UMaze", agent="point", reward_type="sparse", max_episode_steps=300)
    assert pkg.ENV_IDS["PointMaze_LargeDense-v3"]["max_episode_steps"] == 800
    env = pkg.make_vec("PointMaze_Medium-v3", num_envs=2, backend_factory=HostSimBackend, rng_mode="numpy")
    assert env.single_action_space.shape == (2,) and env.single_observation_space["observation"].shape == (4,)


def test_step_tracks_oracle_and_never_resets_into_success():
    env = PointMazeVectorEnv("UMaze", num_envs=3, backend_factory=HostSimBackend, rng_mode="numpy")
    model = load_model("pointmaze_umaze")
    for s in range(50):  # tests/envs/maze/test_point_maze.py:9-17
        obs, info = env.reset(seed=100 + 3 * s)
        assert not bool(info["success"].any())
        assert bool((torch.linalg.norm(obs["achieved_goal"] - obs["desired_goal"], dim=1) > 0.45).all())
    obs, _ = env.reset(seed=7)
    oracles = [OraclePointMazeEnv(MAPS["UMaze"], model) for _ in range(3)]
    for i, o in enumerate(oracles):
        oo, _ = o.reset(seed=7 + i)
        assert np.allclose(obs["observation"][i].double().numpy(), oo["observation"], atol=1e-6)
    rng = np.random.default_rng(0)
    worst = 0
    for _ in range(150):  # free-running: the point bounces off walls, velocity clip engages
        a = rng.uniform(-1.5, 1.5, (3, 2)).astype(np.float32)
        o, r, te, tr, info = env.step(a)
        for i, orc in enumerate(oracles):
            oo, orr, *_ = orc.step(a[i].astype(np.float64))
            worst = max(worst, np.abs(o["observation"][i].double().numpy() - oo["observation"]).max())
            assert float(r[i]) == float(orr)
    assert worst < 5e-4, worst
    assert float(o["observation"].abs()[:, 2:].max()) <= 5.0 + 1.1  # clip 5 before the step, one step of acceleration after
