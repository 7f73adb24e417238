"""Single-env numpy restatement of PointMaze (v3) on the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Follows gymnasium_robotics/envs/maze/point_maze.py (ctor :316-371, reset :373-387, step :389-402, _get_obs :404-410) and
envs/maze/point.py (PointEnv.step :52-69 clips the action and the velocity, frame_skip 1; _get_obs :71-72)."""
from __future__ import annotations

import numpy as np

from .maze import MazeResetLogic, compute_reward, compute_terminated
from .oracle_sim import OracleSim


class OraclePointMazeEnv:
    def __init__(self, maze_map, model, reward_type="sparse", continuing_task=True):
        self.model = model
        self.sim = OracleSim(model)
        self.logic = MazeResetLogic(maze_map, maze_size_scaling=1.0, position_noise_range=0.25)
        self.reward_type, self.continuing_task = reward_type, continuing_task
        self.init_qpos = np.array(model.qpos0, dtype=np.float64)
        self.goal = np.zeros(2)

    def _get_obs(self):
        o = np.concatenate([self.sim.qpos.copy(), self.sim.qvel.copy()])
        return {"observation": o.copy(), "achieved_goal": o[:2].copy(), "desired_goal": self.goal.copy()}

    def reset(self, seed=None, options=None):
        self.goal, reset_pos = self.logic.reset(seed=seed, options=options)
        self.init_qpos[:2] = reset_pos
        s = self.sim
        s.reset_data()
        s.qpos[:] = self.init_qpos
        s.qvel[:] = 0
        s.forward()
        obs = self._get_obs()
        return obs, {"success": bool(np.linalg.norm(obs["achieved_goal"] - self.goal) <= 0.45)}

    def step(self, action):
        action = np.clip(np.asarray(action, dtype=np.float64), -1.0, 1.0)
        self.sim.qvel[:] = np.clip(self.sim.qvel, -5.0, 5.0)
        self.sim.ctrl[:] = action
        self.sim.step(1)
        obs = self._get_obs()
        reward = compute_reward(obs["achieved_goal"], self.goal, self.reward_type)
        terminated = compute_terminated(obs["achieved_goal"], self.goal, self.continuing_task)
        return obs, float(reward), terminated, False, {"success": bool(np.linalg.norm(obs["achieved_goal"] - self.goal) <= 0.45)}
