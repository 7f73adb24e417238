"""Helpers shared by the parity tests, smoke() and the bench's CPU-baseline leg (the only allowed oracle users)."""
from __future__ import annotations

import numpy as np
import torch

from oracle.fetch_env import OracleFetchEnv


def oracle_env_from_model(task, model, reward_type="sparse"):
    """Oracle env built from an already compiled model (so nothing reads /root/reference at run time)."""
    return OracleFetchEnv(task, reward_type=reward_type, model=model)


def oracle_state_record(env, orc):
    """One fp32 state record (layout of include/b200sim.h) from the oracle env's current fp64 state."""
    lay = env.backend.layout
    m = env.model
    rec = np.zeros(lay["stride"], dtype=np.float64)
    s = orc.sim
    rec[lay["qpos"]:lay["qpos"] + m.nq] = s.qpos
    rec[lay["qvel"]:lay["qvel"] + m.nv] = s.qvel
    rec[lay["warm"]:lay["warm"] + m.nv] = s.qacc_warmstart
    rec[lay["ctrl"]:lay["ctrl"] + m.nu] = s.ctrl
    rec[lay["mocap"]:lay["mocap"] + 3] = s.mocap_pos[0]
    rec[lay["mocap"] + 3:lay["mocap"] + 7] = s.mocap_quat[0]
    rec[lay["pose"]:lay["pose"] + 3] = s.site_xpos[orc._gripper_frame]
    rec[lay["pose"] + 3:lay["pose"] + 7] = orc._body_xquat(orc._gripper_frame)
    rec[lay["goal"]:lay["goal"] + 3] = orc.goal
    return rec


def inject_oracle_state(env, oracles):
    recs = torch.as_tensor(np.stack([oracle_state_record(env, o) for o in oracles]), dtype=torch.float32)
    env.backend.state.copy_(recs)
    obs = env.set_state(env.backend.state.clone())
    # set_state refreshes derived data, which re-anchors the stored "pose of the welded body as of the last forward
    # pass" to the injected qpos; the oracle's value is one sub-step stale: restore it
    lay = env.backend.layout
    env.backend.state[:, lay["pose"]:lay["pose"] + 7] = recs[:, lay["pose"]:lay["pose"] + 7].to(env.backend.state.device)
    return obs


def oracle_obs_vector(obs):
    return np.asarray(obs["observation"], dtype=np.float64)
