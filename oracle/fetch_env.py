"""Single-env numpy restatement of the reference's Fetch environments on top of the CPU oracle.
TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header; parity unpinned for the physics).

Every method cites the reference lines it restates (paths relative to
/root/reference/gymnasium_robotics/).  The simulator object is `OracleSim` where the reference uses
(`mujoco.MjModel`, `mujoco.MjData`).
"""
from __future__ import annotations

import numpy as np

from gymnasium_robotics_b200.mjcf import compile_mjcf, EQ_WELD
from .oracle_sim import OracleSim

FETCH_TASKS = {
    # envs/fetch/reach.py:125-147, push.py, slide.py, pick_and_place.py:139-162
    "FetchReach": dict(xml="fetch/reach.xml", has_object=False, block_gripper=True, gripper_extra_height=0.2,
                       target_in_the_air=True, target_offset=0.0, obj_range=0.15, target_range=0.15,
                       distance_threshold=0.05,
                       initial_qpos={"robot0:slide0": 0.4049, "robot0:slide1": 0.48, "robot0:slide2": 0.0}),
    "FetchPush": dict(xml="fetch/push.xml", has_object=True, block_gripper=True, gripper_extra_height=0.0,
                      target_in_the_air=False, target_offset=0.0, obj_range=0.15, target_range=0.15,
                      distance_threshold=0.05,
                      initial_qpos={"robot0:slide0": 0.405, "robot0:slide1": 0.48, "robot0:slide2": 0.0,
                                    "object0:joint": [1.25, 0.53, 0.4, 1.0, 0.0, 0.0, 0.0]}),
    "FetchPickAndPlace": dict(xml="fetch/pick_and_place.xml", has_object=True, block_gripper=False,
                              gripper_extra_height=0.2, target_in_the_air=True, target_offset=0.0, obj_range=0.15,
                              target_range=0.15, distance_threshold=0.05,
                              initial_qpos={"robot0:slide0": 0.405, "robot0:slide1": 0.48, "robot0:slide2": 0.0,
                                            "object0:joint": [1.25, 0.53, 0.4, 1.0, 0.0, 0.0, 0.0]}),
    # envs/fetch/slide.py:160-190
    "FetchSlide": dict(xml="fetch/slide.xml", has_object=True, block_gripper=True, gripper_extra_height=-0.02,
                       target_in_the_air=False, target_offset=np.array([0.4, 0.0, 0.0]), obj_range=0.1, target_range=0.3,
                       distance_threshold=0.05,
                       initial_qpos={"robot0:slide0": 0.05, "robot0:slide1": 0.48, "robot0:slide2": 0.0,
                                     "object0:joint": [1.7, 1.1, 0.41, 1.0, 0.0, 0.0, 0.0]}),
}


def mat2euler(mat):
    """utils/rotations.py:162-184"""
    mat = np.asarray(mat, dtype=np.float64)
    eps4 = np.finfo(np.float64).eps * 4.0
    cy = np.sqrt(mat[..., 2, 2] * mat[..., 2, 2] + mat[..., 1, 2] * mat[..., 1, 2])
    condition = cy > eps4
    euler = np.empty(mat.shape[:-1], dtype=np.float64)
    euler[..., 2] = np.where(condition, -np.arctan2(mat[..., 0, 1], mat[..., 0, 0]), -np.arctan2(-mat[..., 1, 0], mat[..., 1, 1]))
    euler[..., 1] = np.where(condition, -np.arctan2(-mat[..., 0, 2], cy), -np.arctan2(-mat[..., 0, 2], cy))
    euler[..., 0] = np.where(condition, -np.arctan2(mat[..., 1, 2], mat[..., 2, 2]), 0.0)
    return euler


def goal_distance(goal_a, goal_b):
    """envs/fetch/fetch_env.py:16-18"""
    assert goal_a.shape == goal_b.shape
    return np.linalg.norm(goal_a - goal_b, axis=-1)


class OracleFetchEnv:
    """envs/fetch/fetch_env.py (MujocoFetchEnv) + envs/robot_env.py (BaseRobotEnv / MujocoRobotEnv)."""

    def __init__(self, task="FetchPickAndPlace", reward_type="sparse", assets_dir="/root/reference/gymnasium_robotics/envs/assets",
                 n_substeps=20, model=None):
        cfg = dict(FETCH_TASKS[task])
        self.__dict__.update({k: v for k, v in cfg.items() if k not in ("xml", "initial_qpos")})
        self.reward_type = reward_type
        self.n_substeps = n_substeps
        self.model = model if model is not None else compile_mjcf(f"{assets_dir}/{cfg['xml']}")
        self.sim = OracleSim(self.model)
        m = self.model
        self._grip_site = m.site_id("robot0:grip")
        self._obj_site = m.site_id("object0") if self.has_object else -1
        self._gripper_frame = m.frame_site("robot0:gripper_link")  # data.xpos/xquat of the welded body
        self._robot_joints = [j for j, n in enumerate(m.names["joint"]) if n.startswith("robot")]
        self._finger_q = [m.jnt_qposadr[m.joint_id(n)] for n in ("robot0:l_gripper_finger_joint", "robot0:r_gripper_finger_joint")]
        self.goal = np.zeros(0)
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(None)))
        self._env_setup(cfg["initial_qpos"])
        # robot_env.py:301-303
        self.initial_time = float(self.sim.time[0])
        self.initial_qpos = self.sim.qpos.copy()
        self.initial_qvel = self.sim.qvel.copy()

    # ------------------------------------------------------------------ helpers
    @property
    def dt(self):  # robot_env.py:335-338
        return self.model.opt[0] * self.n_substeps

    def _set_joint_qpos(self, name, value):  # utils/mujoco_utils.py:130-151
        m = self.model
        j = m.joint_id(name)
        a = m.jnt_qposadr[j]
        n = 7 if m.jnt_type[j] == 0 else 1
        self.sim.qpos[a:a + n] = value

    def _body_xquat(self, site):
        """xquat of the MJCF body whose frame is tracked by synthetic site `site`."""
        m = self.model
        from gymnasium_robotics_b200.mjcf import qmul
        return qmul(self.sim.xquat[m.site_body[site]], m.site_quat[site])

    # ------------------------------------------------------------------ construction
    def _env_setup(self, initial_qpos):  # fetch_env.py:404-428
        s, m = self.sim, self.model
        for name, value in initial_qpos.items():
            self._set_joint_qpos(name, value)
        # utils/mujoco_utils.py:74-80 reset_mocap_welds
        for i in range(m.neq):
            if m.eq_type[i] == EQ_WELD:
                s.eq_data[i, :7] = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
        s.forward()
        gripper_target = np.array([-0.498, 0.005, -0.431 + self.gripper_extra_height]) + s.site_xpos[self._grip_site]
        s.mocap_pos[0] = gripper_target
        s.mocap_quat[0] = [1.0, 0.0, 1.0, 0.0]
        for _ in range(10):
            s.step(self.n_substeps)
        self.initial_gripper_xpos = s.site_xpos[self._grip_site].copy()
        if self.has_object:
            self.height_offset = s.site_xpos[self._obj_site][2]

    # ------------------------------------------------------------------ GoalEnv API
    def compute_reward(self, achieved_goal, goal, info):  # fetch_env.py:74-80
        d = goal_distance(achieved_goal, goal)
        if self.reward_type == "sparse":
            return -(d > self.distance_threshold).astype(np.float32)
        return -d

    def _is_success(self, achieved_goal, desired_goal):  # fetch_env.py:168-170
        d = goal_distance(achieved_goal, desired_goal)
        return (d < self.distance_threshold).astype(np.float32)

    def _set_action(self, action):  # fetch_env.py:85-105 and 305-310
        action = action.copy()
        pos_ctrl, gripper_ctrl = action[:3], action[3]
        pos_ctrl *= 0.05
        rot_ctrl = [1.0, 0.0, 1.0, 0.0]
        gripper_ctrl = np.array([gripper_ctrl, gripper_ctrl])
        if self.block_gripper:
            gripper_ctrl = np.zeros_like(gripper_ctrl)
        action = np.concatenate([pos_ctrl, rot_ctrl, gripper_ctrl])
        self._ctrl_set_action(action)
        self._mocap_set_action(action)

    def _ctrl_set_action(self, action):  # utils/mujoco_utils.py:34-48
        m, s = self.model, self.sim
        if m.nmocap > 0:
            action = action[m.nmocap * 7:]
        if m.nu > 0:
            for i in range(action.shape[0]):
                # position actuators (biastype != 0): target relative to the current joint position
                idx = m.jnt_qposadr[m.act_trnid[i]]
                s.ctrl[i] = s.qpos[idx] + action[i]

    def _mocap_set_action(self, action):  # utils/mujoco_utils.py:51-71 and 83-107
        m, s = self.model, self.sim
        if m.nmocap > 0:
            a = action[:m.nmocap * 7].reshape(m.nmocap, 7)
            # reset_mocap2body_xpos: mocap <- pose of the welded body as of the last forward pass
            s.mocap_pos[0] = s.site_xpos[self._gripper_frame]
            s.mocap_quat[0] = self._body_xquat(self._gripper_frame)
            s.mocap_pos[:] = s.mocap_pos + a[:, :3]
            s.mocap_quat[:] = s.mocap_quat + a[:, 3:]

    def _site_xvel(self, site):  # utils/mujoco_utils.py:110-127, 228-241
        jp, jr = self.sim.jac_site(site)
        return jp @ self.sim.qvel, jr @ self.sim.qvel

    def _get_obs(self):  # fetch_env.py:107-143 and 312-360
        s, m = self.sim, self.model
        grip_pos = s.site_xpos[self._grip_site].copy()
        dt = self.dt
        grip_velp = self._site_xvel(self._grip_site)[0] * dt
        # utils/mujoco_utils.py:23-31 robot_get_obs (all joints whose name starts with "robot")
        robot_qpos = np.array([s.qpos[m.jnt_qposadr[j]] for j in self._robot_joints])
        robot_qvel = np.array([s.qvel[m.jnt_dofadr[j]] for j in self._robot_joints])
        if self.has_object:
            object_pos = s.site_xpos[self._obj_site].copy()
            object_rot = mat2euler(s.site_xmat[self._obj_site].reshape(3, 3))
            vp, vr = self._site_xvel(self._obj_site)
            object_velp, object_velr = vp * dt, vr * dt
            object_rel_pos = object_pos - grip_pos
            object_velp = object_velp - grip_velp
        else:
            object_pos = object_rot = object_velp = object_velr = object_rel_pos = np.zeros(0)
        gripper_state = robot_qpos[-2:]
        gripper_vel = robot_qvel[-2:] * dt
        achieved_goal = grip_pos.copy() if not self.has_object else np.squeeze(object_pos.copy())
        obs = np.concatenate([grip_pos, object_pos.ravel(), object_rel_pos.ravel(), gripper_state, object_rot.ravel(),
                              object_velp.ravel(), object_velr.ravel(), grip_velp, gripper_vel])
        return {"observation": obs.copy(), "achieved_goal": achieved_goal.copy(), "desired_goal": self.goal.copy()}

    def _step_callback(self):  # fetch_env.py:295-303
        if self.block_gripper:
            self._set_joint_qpos("robot0:l_gripper_finger_joint", 0.0)
            self._set_joint_qpos("robot0:r_gripper_finger_joint", 0.0)
            self.sim.forward()

    def step(self, action):  # robot_env.py:114-152
        action = np.clip(np.asarray(action, dtype=np.float64), -1.0, 1.0)
        self._set_action(action)
        self.sim.step(self.n_substeps)  # robot_env.py:340-341
        self._step_callback()
        obs = self._get_obs()
        info = {"is_success": self._is_success(obs["achieved_goal"], self.goal)}
        reward = self.compute_reward(obs["achieved_goal"], self.goal, info)
        return obs, reward, False, False, info

    def _reset_sim(self):  # fetch_env.py:375-402
        s = self.sim
        s.reset_data()
        s.time[0] = self.initial_time
        s.qpos[:] = self.initial_qpos
        s.qvel[:] = self.initial_qvel
        if self.has_object:
            object_xpos = self.initial_gripper_xpos[:2]
            while np.linalg.norm(object_xpos - self.initial_gripper_xpos[:2]) < 0.1:
                object_xpos = self.initial_gripper_xpos[:2] + self.np_random.uniform(-self.obj_range, self.obj_range, size=2)
            a = self.model.jnt_qposadr[self.model.joint_id("object0:joint")]
            s.qpos[a:a + 2] = object_xpos
        s.forward()
        return True

    def _sample_goal(self):  # fetch_env.py:153-166
        if self.has_object:
            goal = self.initial_gripper_xpos[:3] + self.np_random.uniform(-self.target_range, self.target_range, size=3)
            goal += self.target_offset
            goal[2] = self.height_offset
            if self.target_in_the_air and self.np_random.uniform() < 0.5:
                goal[2] += self.np_random.uniform(0, 0.45)
        else:
            goal = self.initial_gripper_xpos[:3] + self.np_random.uniform(-self.target_range, self.target_range, size=3)
        return goal.copy()

    def reset(self, seed=None):  # robot_env.py:154-186 (+ gymnasium.Env.reset seeding)
        if seed is not None:
            self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        did = False
        while not did:
            did = self._reset_sim()
        self.goal = self._sample_goal().copy()
        return self._get_obs(), {}
