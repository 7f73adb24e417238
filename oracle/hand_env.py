"""Single-env numpy restatement of the reference's Shadow-Hand block manipulation environment on top of the CPU oracle.
TEST INFRASTRUCTURE ONLY (see oracle/oracle.c header; parity unpinned for the physics).

Restates envs/shadow_dexterous_hand/hand_env.py (MujocoHandEnv), manipulate.py (MujocoManipulateEnv) and
manipulate_block.py (MujocoHandBlockEnv) + envs/robot_env.py; every method cites the lines it follows (paths relative to
/root/reference/gymnasium_robotics/).  The visual-only `target` free body (manipulate_block.xml:32-36: contype 0,
written only by `_render_callback`, never observed) is dropped at compile time, so qpos has 24 + 7 entries.
"""
from __future__ import annotations

import numpy as np

from gymnasium_robotics_b200 import rotations
from gymnasium_robotics_b200.mjcf import compile_mjcf
from .oracle_sim import OracleSim

HAND_BLOCK_XML = "hand/manipulate_block.xml"
TARGET_POSITION_RANGE = np.array([(-0.04, 0.04), (-0.06, 0.02), (0.0, 0.06)])  # manipulate_block.py:226


def compile_hand_model(assets_dir="/root/reference/gymnasium_robotics/envs/assets", xml=HAND_BLOCK_XML):
    return compile_mjcf(f"{assets_dir}/{xml}", overrides={"drop_bodies": ["target"], "sensor_prefix": "robot0:TS_"})


class OracleHandBlockEnv:
    def __init__(self, target_position="ignore", target_rotation="xyz", reward_type="sparse", model=None,
                 randomize_initial_position=True, randomize_initial_rotation=True, distance_threshold=0.01,
                 rotation_threshold=0.1, n_substeps=20, touch_get_obs=None, ignore_z_target_rotation=False):
        # manipulate.py:24-85 (ctor), manipulate_block.py:214-230
        self.target_position, self.target_rotation = target_position, target_rotation
        self.target_position_range = TARGET_POSITION_RANGE
        self.parallel_quats = rotations.parallel_quats()
        self.randomize_initial_position = randomize_initial_position
        self.randomize_initial_rotation = randomize_initial_rotation
        self.distance_threshold, self.rotation_threshold = distance_threshold, rotation_threshold
        self.reward_type, self.n_substeps = reward_type, n_substeps
        self.touch_get_obs = touch_get_obs  # manipulate_touch_sensors.py:10-64 (None = plain env without touch observation)
        self.ignore_z_target_rotation = ignore_z_target_rotation  # True for the pen (manipulate_pen.py:231)
        assert target_position in ("ignore", "fixed", "random")
        assert target_rotation in ("ignore", "fixed", "xyz", "z", "parallel")
        self.model = model if model is not None else compile_hand_model(
            xml=HAND_BLOCK_XML if touch_get_obs is None else "hand/manipulate_block_touch_sensors.xml")
        self.sim = OracleSim(self.model)
        m = self.model
        self._robot_joints = [j for j, n in enumerate(m.names["joint"]) if n.startswith("robot")]
        self._obj_q = int(m.jnt_qposadr[m.joint_id("object:joint")])
        self._obj_v = int(m.jnt_dofadr[m.joint_id("object:joint")])
        self._center_site = m.site_id("object:center")
        self.goal = np.zeros(0)
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(None)))
        self.sim.forward()  # manipulate.py:148-151 _env_setup with initial_qpos = {}
        self.initial_time = float(self.sim.time[0])  # robot_env.py:301-303
        self.initial_qpos = self.sim.qpos.copy()
        self.initial_qvel = self.sim.qvel.copy()

    # ------------------------------------------------------------------ GoalEnv API
    def _goal_distance(self, goal_a, goal_b):  # manipulate.py:88-115
        goal_a, goal_b = np.asarray(goal_a, dtype=np.float64), np.asarray(goal_b, dtype=np.float64)
        d_pos = np.zeros_like(goal_a[..., 0])
        d_rot = np.zeros_like(goal_b[..., 0])
        if self.target_position != "ignore":
            d_pos = np.linalg.norm(goal_a[..., :3] - goal_b[..., :3], axis=-1)
        if self.target_rotation != "ignore":
            quat_a, quat_b = goal_a[..., 3:], goal_b[..., 3:]
            if self.ignore_z_target_rotation:
                euler_a, euler_b = rotations.quat2euler(quat_a), rotations.quat2euler(quat_b)
                euler_a[..., 2] = euler_b[..., 2]   # the reference indexes [2] (1-D goals); batched goals use the last axis here
                quat_a = rotations.euler2quat(euler_a)
            quat_diff = rotations.quat_mul(quat_a, rotations.quat_conjugate(quat_b))
            d_rot = 2 * np.arccos(np.clip(quat_diff[..., 0], -1.0, 1.0))
        return d_pos, d_rot

    def compute_reward(self, achieved_goal, goal, info):  # manipulate.py:120-128
        if self.reward_type == "sparse":
            return self._is_success(achieved_goal, goal).astype(np.float32) - 1.0
        d_pos, d_rot = self._goal_distance(achieved_goal, goal)
        return -(10.0 * d_pos + d_rot)

    def _is_success(self, achieved_goal, desired_goal):  # manipulate.py:133-138
        d_pos, d_rot = self._goal_distance(achieved_goal, desired_goal)
        return ((d_pos < self.distance_threshold).astype(np.float32) * (d_rot < self.rotation_threshold).astype(np.float32))

    # ------------------------------------------------------------------ RobotEnv pieces
    def _set_action(self, action):  # hand_env.py:42-61 (absolute control; the relative branch is dead code for -v1)
        m = self.model
        ctrlrange = np.asarray(m.act_ctrlrange).reshape(-1, 2)
        half = (ctrlrange[:, 1] - ctrlrange[:, 0]) / 2.0
        center = (ctrlrange[:, 1] + ctrlrange[:, 0]) / 2.0
        self.sim.ctrl[:] = np.clip(center + action * half, ctrlrange[:, 0], ctrlrange[:, 1])

    def _get_achieved_goal(self):  # manipulate.py:143-146
        return self.sim.qpos[self._obj_q:self._obj_q + 7].copy()

    def _get_obs(self):  # manipulate.py:298-314, utils/mujoco_utils.py:23-31
        m, s = self.model, self.sim
        robot_qpos = np.array([s.qpos[m.jnt_qposadr[j]] for j in self._robot_joints])
        robot_qvel = np.array([s.qvel[m.jnt_dofadr[j]] for j in self._robot_joints])
        object_qvel = s.qvel[self._obj_v:self._obj_v + 6]
        achieved = self._get_achieved_goal()
        touch = np.zeros(0)  # manipulate_touch_sensors.py:107-138; all touch sensors of this model are "robot0:TS_*"
        if self.touch_get_obs == "sensordata":
            touch = s.sensordata.copy()
        elif self.touch_get_obs == "boolean":
            touch = (s.sensordata > 0.0).astype(np.float64)
        elif self.touch_get_obs == "log":
            touch = np.log(s.sensordata + 1.0)
        return {"observation": np.concatenate([robot_qpos, robot_qvel, object_qvel, achieved, touch]),
                "achieved_goal": achieved.copy(), "desired_goal": self.goal.copy()}

    def step(self, action):  # robot_env.py:114-152
        action = np.clip(np.asarray(action, dtype=np.float64), -1.0, 1.0)
        self._set_action(action)
        self.sim.step(self.n_substeps)
        obs = self._get_obs()
        info = {"is_success": self._is_success(obs["achieved_goal"], self.goal)}
        reward = self.compute_reward(obs["achieved_goal"], self.goal, info)
        return obs, reward, False, False, info

    def _reset_sim(self):  # manipulate.py:154-224
        s, rng = self.sim, self.np_random
        s.reset_data()
        s.time[0] = self.initial_time
        s.qpos[:] = self.initial_qpos
        s.qvel[:] = self.initial_qvel
        s.forward()
        q = s.qpos[self._obj_q:self._obj_q + 7].copy()
        initial_pos, initial_quat = q[:3], q[3:]
        if self.randomize_initial_rotation:
            if self.target_rotation == "z":
                angle = rng.uniform(-np.pi, np.pi)
                offset = rotations.quat_from_angle_and_axis(angle, np.array([0.0, 0.0, 1.0]))
                initial_quat = rotations.quat_mul(initial_quat, offset)
            elif self.target_rotation == "parallel":
                angle = rng.uniform(-np.pi, np.pi)
                z_quat = rotations.quat_from_angle_and_axis(angle, np.array([0.0, 0.0, 1.0]))
                parallel = self.parallel_quats[rng.integers(len(self.parallel_quats))]
                initial_quat = rotations.quat_mul(initial_quat, rotations.quat_mul(z_quat, parallel))
            elif self.target_rotation in ("xyz", "ignore"):
                angle = rng.uniform(-np.pi, np.pi)
                axis = rng.uniform(-1.0, 1.0, size=3)
                initial_quat = rotations.quat_mul(initial_quat, rotations.quat_from_angle_and_axis(angle, axis))
        if self.randomize_initial_position and self.target_position != "fixed":
            initial_pos = initial_pos + rng.normal(size=3, scale=0.005)
        initial_quat = initial_quat / np.linalg.norm(initial_quat)
        s.qpos[self._obj_q:self._obj_q + 7] = np.concatenate([initial_pos, initial_quat])
        for _ in range(10):  # settle
            self._set_action(np.zeros(20))
            s.step(self.n_substeps)
        s.forward()
        return bool(s.site_xpos[self._center_site][2] > 0.04)  # is_on_palm

    def _sample_goal(self):  # manipulate.py:226-279
        rng = self.np_random
        obj = self.sim.qpos[self._obj_q:self._obj_q + 7]
        if self.target_position == "random":
            offset = rng.uniform(self.target_position_range[:, 0], self.target_position_range[:, 1])
            target_pos = obj[:3] + offset
        else:
            target_pos = obj[:3].copy()
        if self.target_rotation == "z":
            target_quat = rotations.quat_from_angle_and_axis(rng.uniform(-np.pi, np.pi), np.array([0.0, 0.0, 1.0]))
        elif self.target_rotation == "parallel":
            target_quat = rotations.quat_from_angle_and_axis(rng.uniform(-np.pi, np.pi), np.array([0.0, 0.0, 1.0]))
            target_quat = rotations.quat_mul(target_quat, self.parallel_quats[rng.integers(len(self.parallel_quats))])
        elif self.target_rotation == "xyz":
            angle = rng.uniform(-np.pi, np.pi)
            target_quat = rotations.quat_from_angle_and_axis(angle, rng.uniform(-1.0, 1.0, size=3))
        else:
            # the reference reads all 7 numbers of the joint here (manipulate.py:266-267) and its own shape assert
            # would fail; the quaternion part is what the later code needs
            target_quat = obj[3:].copy()
        target_quat = target_quat / np.linalg.norm(target_quat)
        return np.concatenate([target_pos, target_quat])

    def reset(self, seed=None):  # robot_env.py:154-186 (+ gymnasium.Env.reset seeding)
        if seed is not None:
            self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        did = False
        while not did:
            did = self._reset_sim()
        self.goal = self._sample_goal().copy()
        return self._get_obs(), {}


# ---------------------------------------------------------------------------------------------------------------- HandReach
class OracleHandReachEnv:
    """envs/shadow_dexterous_hand/reach.py (MujocoHandReachEnv) + hand_env.py + robot_env.py, on the oracle simulator."""

    def __init__(self, reward_type="sparse", model=None, distance_threshold=0.01, n_substeps=20, initial_qpos=None):
        from gymnasium_robotics_b200.hand import FINGERTIP_SITE_NAMES, REACH_INITIAL_QPOS

        self.site_names = FINGERTIP_SITE_NAMES
        self.distance_threshold, self.reward_type, self.n_substeps = distance_threshold, reward_type, n_substeps
        self.model = model if model is not None else compile_mjcf(
            "/root/reference/gymnasium_robotics/envs/assets/hand/reach.xml", overrides={"sensor_prefix": "robot0:TS_"})
        self.sim = OracleSim(self.model)
        m, s = self.model, self.sim
        self._tips = [m.site_id(n) for n in self.site_names]
        self._robot_joints = [j for j, n in enumerate(m.names["joint"]) if n.startswith("robot")]
        self.goal = np.zeros(0)
        self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(None)))
        # _env_setup, reach.py:286-296
        for name, value in (initial_qpos or REACH_INITIAL_QPOS).items():
            s.qpos[m.jnt_qposadr[m.joint_id(name)]] = value
        s.forward()
        self.initial_goal = self._get_achieved_goal().copy()
        self.palm_xpos = s.xpos[m.names["body_map"]["robot0:palm"]].copy()
        self.initial_time = float(s.time[0])
        self.initial_qpos, self.initial_qvel = s.qpos.copy(), s.qvel.copy()

    def _get_achieved_goal(self):  # reach.py:278-283
        return np.array([self.sim.site_xpos[i] for i in self._tips]).flatten()

    def compute_reward(self, achieved_goal, goal, info):  # reach.py:88-93
        d = np.linalg.norm(np.asarray(achieved_goal) - np.asarray(goal), axis=-1)
        return -(d > self.distance_threshold).astype(np.float32) if self.reward_type == "sparse" else -d

    def _is_success(self, achieved_goal, desired_goal):  # reach.py:123-125
        d = np.linalg.norm(np.asarray(achieved_goal) - np.asarray(desired_goal), axis=-1)
        return (d < self.distance_threshold).astype(np.float32)

    _set_action = OracleHandBlockEnv._set_action

    def _get_obs(self):  # reach.py:298-310
        m, s = self.model, self.sim
        robot_qpos = np.array([s.qpos[m.jnt_qposadr[j]] for j in self._robot_joints])
        robot_qvel = np.array([s.qvel[m.jnt_dofadr[j]] for j in self._robot_joints])
        achieved = self._get_achieved_goal()
        return {"observation": np.concatenate([robot_qpos, robot_qvel, achieved]), "achieved_goal": achieved.copy(),
                "desired_goal": self.goal.copy()}

    def _sample_goal(self):  # reach.py:95-121
        rng = self.np_random
        finger_names = [n for n in self.site_names if n != "robot0:S_thtip"]
        finger_name = rng.choice(finger_names)
        thumb_idx, finger_idx = self.site_names.index("robot0:S_thtip"), self.site_names.index(finger_name)
        meeting_pos = self.palm_xpos + np.array([0.0, -0.09, 0.05])
        meeting_pos = meeting_pos + rng.normal(scale=0.005, size=meeting_pos.shape)
        goal = self.initial_goal.copy().reshape(-1, 3)
        for idx in (thumb_idx, finger_idx):
            offset_direction = meeting_pos - goal[idx]
            offset_direction /= np.linalg.norm(offset_direction)
            goal[idx] = meeting_pos - 0.005 * offset_direction
        if rng.uniform() < 0.1:
            goal = self.initial_goal.copy()
        return goal.flatten()

    def step(self, action):  # robot_env.py:114-152
        action = np.clip(np.asarray(action, dtype=np.float64), -1.0, 1.0)
        self._set_action(action)
        self.sim.step(self.n_substeps)
        obs = self._get_obs()
        info = {"is_success": self._is_success(obs["achieved_goal"], self.goal)}
        return obs, self.compute_reward(obs["achieved_goal"], self.goal, info), False, False, info

    def reset(self, seed=None):  # robot_env.py:154-186, 305-316
        if seed is not None:
            self.np_random = np.random.Generator(np.random.PCG64(np.random.SeedSequence(seed)))
        s = self.sim
        s.reset_data()
        s.time[0] = self.initial_time
        s.qpos[:] = self.initial_qpos
        s.qvel[:] = self.initial_qvel
        s.forward()
        self.goal = self._sample_goal().copy()
        return self._get_obs(), {}
