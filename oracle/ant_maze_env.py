"""Single-env numpy restatement of AntMaze (v5) on the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Follows gymnasium_robotics/envs/maze/ant_maze_v5.py (ctor :221-280, reset :282-293, step :295-310, _get_obs :312-320)
and maze_v4.py (MazeEnv.reset :299-358, compute_reward :381-388, compute_terminated :390-398, update_goal :400-418).
The inner `gymnasium.envs.mujoco.ant_v5.AntEnv` is un-vendored [ext]: frame_skip 5, ctrl = action, observation =
concat(qpos, qvel) with exclude_current_positions_from_observation=False, reset_noise_scale=0, its own reward and
healthy-termination are discarded by AntMazeEnv.step (ant_maze_v5.py:296).
"""
from __future__ import annotations

import numpy as np

from gymnasium_robotics_b200.mjcf import compile_mjcf, make_maze_xml
from .maze import MazeResetLogic, compute_reward, compute_terminated
from .oracle_sim import OracleSim

LARGE_MAZE = [[1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1],
              [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1],
              [1, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 1],
              [1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1],
              [1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1],
              [1, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 1],
              [1, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 1],
              [1, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1],
              [1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1]]  # envs/maze/maps.py:93-103
U_MAZE = [[1, 1, 1, 1, 1], [1, 0, 0, 0, 1], [1, 1, 1, 0, 1], [1, 0, 0, 0, 1], [1, 1, 1, 1, 1]]  # maps.py:52-58


def compile_ant_maze(maze_map, ant_xml="/root/reference/gymnasium_robotics/envs/mujoco/assets/ant.xml"):
    root, grid = make_maze_xml(ant_xml, maze_map, 4, 0.5)
    return compile_mjcf(ant_xml, root=root, grid=grid)


class OracleAntMazeEnv:
    FRAME_SKIP = 5

    def __init__(self, maze_map=LARGE_MAZE, reward_type="sparse", continuing_task=True, reset_target=False, model=None,
                 include_cfrc_ext_in_observation=False):
        self.model = model if model is not None else compile_ant_maze(maze_map)
        self.sim = OracleSim(self.model)
        self.logic = MazeResetLogic(maze_map, maze_size_scaling=4.0, position_noise_range=0.25)
        self.reward_type, self.continuing_task, self.reset_target = reward_type, continuing_task, reset_target
        self.init_qpos = np.array(self.model.qpos0, dtype=np.float64)
        self.init_qvel = np.zeros(self.model.nv)
        self.goal = np.zeros(2)
        # Ant-v5 [ext] (gymnasium ant_v5.AntEnv): include_cfrc_ext_in_observation defaults to True, contact_force_range (-1, 1);
        # AntMaze_*-v5 builds it with the defaults (ant_maze_v5.py:249-255) => observation (105,) (ant_maze_v5.py:99); -v4 (Ant-v4,
        # use_contact_forces False) => (27,)
        self.include_cfrc = include_cfrc_ext_in_observation
        self._cfrc = np.zeros((len(self.model.mjbody_rt), 6))

    def _ant_obs(self):
        o = [self.sim.qpos.copy(), self.sim.qvel.copy()]
        if self.include_cfrc:
            o.append(np.clip(self._cfrc[1:], -1.0, 1.0).ravel())      # contact_forces[1:]: the world body is left out
        return np.concatenate(o)

    def _get_obs(self):  # ant_maze_v5.py:312-320
        o = self._ant_obs()
        return {"observation": o[2:].copy(), "achieved_goal": o[:2].copy(), "desired_goal": self.goal.copy()}

    def reset(self, seed=None, options=None):  # maze_v4.py:299-358, ant_maze_v5.py:282-293
        self.goal, reset_pos = self.logic.reset(seed=seed, options=options)
        self.init_qpos[:2] = reset_pos
        s = self.sim
        s.reset_data()
        s.qpos[:] = self.init_qpos
        s.qvel[:] = self.init_qvel
        s.forward()
        self._cfrc[:] = 0.0      # mj_resetData zeroes cfrc_ext; set_state's mj_forward does not recompute it
        obs = self._get_obs()
        return obs, {"success": bool(np.linalg.norm(obs["achieved_goal"] - self.goal) <= 0.45)}

    def step(self, action):  # ant_maze_v5.py:295-310
        self.sim.ctrl[:] = np.asarray(action, dtype=np.float64)
        self.sim.step(self.FRAME_SKIP)
        self._cfrc = self.sim.cfrc_ext()      # MujocoEnv.do_simulation [ext]: mj_step, then mj_rnePostConstraint
        obs = self._get_obs()
        reward = compute_reward(obs["achieved_goal"], self.goal, self.reward_type)
        terminated = compute_terminated(obs["achieved_goal"], self.goal, self.continuing_task)
        info = {"success": bool(np.linalg.norm(obs["achieved_goal"] - self.goal) <= 0.45)}
        return obs, float(reward), terminated, False, info
