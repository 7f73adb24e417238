"""Batched Fetch environments (`gym.vector.VectorEnv`-style) on the b200sim CUDA path.

Host-side mirror of the reference's Fetch stack, batched over `num_envs`:
  * task tables              envs/fetch/{reach,push,slide,pick_and_place}.py ctor kwargs (e.g. pick_and_place.py:139-162)
  * construction/_env_setup  envs/fetch/fetch_env.py:404-428, envs/robot_env.py:292-303
  * reset/_reset_sim/_sample_goal   envs/robot_env.py:154-186, envs/fetch/fetch_env.py:375-402, 153-166
  * step                     envs/robot_env.py:114-152  (runs entirely inside one CUDA kernel, csrc/fetch_task.cuh)
  * compute_reward/_is_success      envs/fetch/fetch_env.py:74-80, 168-170
  * TimeLimit (max_episode_steps=50) and vector autoreset as gymnasium's wrappers/vector envs do.
All per-step arithmetic happens on the GPU; this file only owns reset-time sampling and bookkeeping.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _lib
from .mjcf import EQ_WELD, JNT_FREE
from .models import load_model
from .rollout import CtorPickle
from .spaces import Box, Dict as DictSpace, batch_space

FETCH_TASKS = {
    "FetchReach": dict(model="fetch_reach", has_object=False, block_gripper=True, gripper_extra_height=0.2,
                       target_in_the_air=True, target_offset=0.0, obj_range=0.15, target_range=0.15, distance_threshold=0.05,
                       initial_qpos={"robot0:slide0": 0.4049, "robot0:slide1": 0.48, "robot0:slide2": 0.0}),
    "FetchPush": dict(model="fetch_push", has_object=True, block_gripper=True, gripper_extra_height=0.0,
                      target_in_the_air=False, target_offset=0.0, obj_range=0.15, target_range=0.15, distance_threshold=0.05,
                      initial_qpos={"robot0:slide0": 0.405, "robot0:slide1": 0.48, "robot0:slide2": 0.0,
                                    "object0:joint": [1.25, 0.53, 0.4, 1.0, 0.0, 0.0, 0.0]}),
    "FetchPickAndPlace": dict(model="fetch_pick_and_place", has_object=True, block_gripper=False, gripper_extra_height=0.2,
                              target_in_the_air=True, target_offset=0.0, obj_range=0.15, target_range=0.15,
                              distance_threshold=0.05,
                              initial_qpos={"robot0:slide0": 0.405, "robot0:slide1": 0.48, "robot0:slide2": 0.0,
                                            "object0:joint": [1.25, 0.53, 0.4, 1.0, 0.0, 0.0, 0.0]}),
    # envs/fetch/slide.py:160-190 (cylinder puck; target_offset = [0.4, 0, 0])
    "FetchSlide": dict(model="fetch_slide", has_object=True, block_gripper=True, gripper_extra_height=-0.02,
                       target_in_the_air=False, target_offset=(0.4, 0.0, 0.0), obj_range=0.1, target_range=0.3, distance_threshold=0.05,
                       initial_qpos={"robot0:slide0": 0.05, "robot0:slide1": 0.48, "robot0:slide2": 0.0,
                                     "object0:joint": [1.7, 1.1, 0.41, 1.0, 0.0, 0.0, 0.0]}),
}
N_SUBSTEPS = 20
REF_POINT = (1.0, 0.75, 0.4)  # fixed world point the device spatial algebra is expressed about


def make_task_struct(model, cfg, reward_type, n_substeps=N_SUBSTEPS):
    t = _lib.FetchTaskC()
    t.has_object, t.block_gripper = int(cfg["has_object"]), int(cfg["block_gripper"])
    t.n_substeps, t.reward_dense = n_substeps, int(reward_type == "dense")
    t.grip_site = model.site_id("robot0:grip")
    t.obj_site = model.site_id("object0") if cfg["has_object"] else -1
    t.frame_site = model.frame_site("robot0:gripper_link")
    robot = [j for j, n in enumerate(model.names["joint"]) if n.startswith("robot")]
    t.nrobot = len(robot)
    for i, j in enumerate(robot):
        t.robot_qadr[i], t.robot_dadr[i] = int(model.jnt_qposadr[j]), int(model.jnt_dofadr[j])
    t.finger_qadr[0] = int(model.jnt_qposadr[model.joint_id("robot0:l_gripper_finger_joint")])
    t.finger_qadr[1] = int(model.jnt_qposadr[model.joint_id("robot0:r_gripper_finger_joint")])
    t.nobs = 25 if cfg["has_object"] else 10
    t.distance_threshold = float(cfg["distance_threshold"])
    t.dt = float(model.opt[0] * n_substeps)
    return t


def welded_eq_data(model):
    """utils/mujoco_utils.py:74-80 reset_mocap_welds, applied to the model constants before upload."""
    eq = np.array(model.eq_data, dtype=np.float64).reshape(-1, 11).copy()
    for i in range(model.neq):
        if model.eq_type[i] == EQ_WELD:
            eq[i, :7] = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0]
    return eq


class _DevArray:
    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (int(ptr), False), "version": 3, "strides": None}


class CudaBackend:
    """Thin owner of a `b200sim_t` handle; all arguments are torch CUDA tensors passed as raw device pointers.

    Outputs are PACKED (b200sim_set_packed): one [N, W] fp32 row per env, obs | achieved | desired | reward | success | terminated |
    truncated; `new_outputs()` hands out that buffer under "packed" next to column views of it under the classic names, so one
    device->host copy (or one all-gather) moves everything a step produced."""

    def __init__(self, model, eq_data, task, num_envs, device):
        if not torch.cuda.is_available():
            raise RuntimeError("b200sim needs a CUDA device: there is no CPU fallback on the product path")
        self.device = torch.device(device)
        self.L = _lib.lib()
        blob = model.to_blob()
        h = ctypes.c_void_p()
        ref = np.asarray(getattr(self, "REF", REF_POINT), dtype=np.float32)
        # eq_data: NULL (keep the blob's equality data) or exactly neq x 11 doubles (include/b200sim.h); the C side cannot see the
        # length, so an empty override must become NULL here -- a pointer to a zero-length array would be read out of bounds
        eq = np.ascontiguousarray(eq_data, dtype=np.float64)
        if eq.size not in (0, 11 * int(model.neq)):
            raise ValueError(f"eq_data must be empty or [neq = {int(model.neq)}, 11], got shape {eq.shape}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        rc = self.L.b200sim_create(blob, len(blob), eq.ctypes.data if eq.size else None, ref.ctypes.data, ctypes.byref(task), num_envs,
                                   self.device.index, ctypes.byref(h))
        if rc != 0:
            raise RuntimeError(f"b200sim_create failed ({rc}): {self.L.b200sim_last_error(None).decode()}")
        self.h = h
        self.num_envs, self.nobs = num_envs, task.nobs
        self.ngoal, self.nact = (3, 4) if task.kind == 0 else (task.ngoal, task.nact)
        lay = (ctypes.c_int * 9)()
        self.L.b200sim_layout(h, lay)
        self.layout = dict(zip(("qpos", "qvel", "warm", "ctrl", "mocap", "pose", "goal", "stride", "penv"), list(lay)))
        self.state = torch.as_tensor(_DevArray(self.L.b200sim_state(h), (num_envs, self.layout["stride"])), device=self.device)
        # per-env step counters of the in-kernel TimeLimit (b200sim_set_time_limit), capacity-overflow counter, info words
        self.elapsed = torch.as_tensor(_DevArray(self.L.b200sim_elapsed(h), (num_envs,), "<i4"), device=self.device)
        self.overflow_counter = torch.as_tensor(_DevArray(self.L.b200sim_overflow_counter(h), (1,), "<i8"), device=self.device)
        self.info = torch.zeros(num_envs, dtype=torch.int32, device=self.device)
        self.packed_w = int(self.L.b200sim_set_packed(h, 1))

    def set_time_limit(self, max_episode_steps, terminate_on_success=False):
        self._check(self.L.b200sim_set_time_limit(self.h, int(max_episode_steps or 0), int(bool(terminate_on_success))))

    def close(self):
        if getattr(self, "h", None):
            self.state = None
            self.L.b200sim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"b200sim call failed ({rc}): {self.L.b200sim_last_error(self.h).decode()}")

    def new_outputs(self):
        n, d, no, ng = self.num_envs, self.device, self.nobs, self.ngoal
        p = torch.zeros((n, self.packed_w), dtype=torch.float32, device=d)   # zeros: refresh / raw launches do not write the flags
        flags = torch.zeros((2, n), dtype=torch.uint8, device=d)
        k = no + 2 * ng
        return dict(packed=p, obs=p[:, :no], achieved=p[:, no:no + ng], desired=p[:, no + ng:k], reward=p[:, k], success=p[:, k + 1],
                    terminated=flags[0].view(torch.bool), truncated=flags[1].view(torch.bool), flags=flags)

    def _ptrs(self, out):
        return [out["packed"].data_ptr(), None, None, None, None]

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def step(self, actions, out, info=None):
        """One env-step of every env; the flags of the step land in out["terminated"] / out["truncated"], the solver's info word
        (Newton iterations | capacity-overflow bits << 16) in `info` (default: the backend's persistent `self.info`)."""
        assert actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous() and actions.shape == (self.num_envs, self.nact)
        info = self.info if info is None else info
        f = out["flags"]
        self._check(self.L.b200sim_step(self.h, actions.data_ptr(), *self._ptrs(out), f[0].data_ptr(), f[1].data_ptr(), info.data_ptr(),
                                        self._stream()))

    def refresh(self, mask, out):
        self._check(self.L.b200sim_refresh(self.h, mask.data_ptr() if mask is not None else None, *self._ptrs(out), self._stream()))

    def raw_step(self, nstep, out, mask=None):
        if mask is None:
            self._check(self.L.b200sim_raw_step(self.h, int(nstep), *self._ptrs(out), self._stream()))
        else:
            self._check(self.L.b200sim_raw_step_masked(self.h, mask.data_ptr(), int(nstep), *self._ptrs(out), self._stream()))

    def reset_draw(self, mask, rest_record, params, seed, env_offset, episode, out):
        """b200sim_reset: in-kernel draw of object start + goal for the masked envs (None = all), then mj_forward + _get_obs."""
        assert rest_record.is_cuda and rest_record.dtype == torch.float32 and rest_record.numel() == self.layout["stride"]
        assert episode.is_cuda and episode.dtype == torch.int32 and episode.numel() == self.num_envs
        self._check(self.L.b200sim_reset(self.h, mask.data_ptr() if mask is not None else None, rest_record.data_ptr(), ctypes.byref(params),
                                         int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_offset), episode.data_ptr(), *self._ptrs(out), self._stream()))

    def reset_uniform(self, mask, rest_record, params, seed, env_offset, episode, out):
        """b200sim_reset_uniform: record <- rest record + a fixed list of uniform draws, then mj_forward + _get_obs."""
        assert rest_record.is_cuda and rest_record.dtype == torch.float32 and rest_record.numel() == self.layout["stride"]
        assert episode.is_cuda and episode.dtype == torch.int32 and episode.numel() == self.num_envs
        self._check(self.L.b200sim_reset_uniform(self.h, mask.data_ptr() if mask is not None else None, rest_record.data_ptr(), ctypes.byref(params),
                                                 int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_offset), episode.data_ptr(), *self._ptrs(out), self._stream()))

    def reset_maze(self, mask, rest_record, params, goal_xy, reset_xy, seed, env_offset, episode, out):
        """b200sim_reset_maze: goal cell + noise, reset cell away from the goal + noise, then mj_forward + _get_obs."""
        assert rest_record.is_cuda and rest_record.dtype == torch.float32 and rest_record.numel() == self.layout["stride"]
        assert episode.is_cuda and episode.dtype == torch.int32 and episode.numel() == self.num_envs
        for t, n in ((goal_xy, params.n_goal), (reset_xy, params.n_reset)):
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == (n, 2)
        self._check(self.L.b200sim_reset_maze(self.h, mask.data_ptr() if mask is not None else None, rest_record.data_ptr(), ctypes.byref(params),
                                              goal_xy.data_ptr(), reset_xy.data_ptr(), int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_offset),
                                              episode.data_ptr(), *self._ptrs(out), self._stream()))

    def reset_hand_pose(self, mask, rest_record, params, parallel, seed, env_offset, episode, attempt):
        """b200sim_reset_hand_pose: records of the masked envs <- rest record + drawn object start pose (goal kept); no refresh."""
        assert rest_record.is_cuda and rest_record.numel() == self.layout["stride"] and parallel.is_cuda and tuple(parallel.shape) == (24, 4)
        self._check(self.L.b200sim_reset_hand_pose(self.h, mask.data_ptr() if mask is not None else None, rest_record.data_ptr(), ctypes.byref(params),
                                                   parallel.data_ptr(), int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_offset), episode.data_ptr(),
                                                   int(attempt), self._stream()))

    def reset_hand_goal(self, mask, params, parallel, seed, env_offset, episode, out):
        """b200sim_reset_hand_goal: goal drawn from the settled object pose, episode counters incremented, then the refresh."""
        self._check(self.L.b200sim_reset_hand_goal(self.h, mask.data_ptr() if mask is not None else None, ctypes.byref(params), parallel.data_ptr(),
                                                   int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_offset), episode.data_ptr(), *self._ptrs(out),
                                                   self._stream()))

    def reset_reach(self, mask, rest_record, params, seed, env_offset, episode, out):
        """b200sim_reset_reach: HandReach goal drawn on the device, then mj_forward + _get_obs."""
        assert rest_record.is_cuda and rest_record.numel() == self.layout["stride"] and episode.is_cuda and episode.dtype == torch.int32
        self._check(self.L.b200sim_reset_reach(self.h, mask.data_ptr() if mask is not None else None, rest_record.data_ptr(), ctypes.byref(params),
                                               int(seed) & 0xFFFFFFFFFFFFFFFF, int(env_offset), episode.data_ptr(), *self._ptrs(out), self._stream()))

    def check_state(self, bad, rest_record, keep):
        """b200sim_check_state: bad[i] = record i holds NaN / |x| > 1e10; such records are put back to `rest_record` (if given)."""
        assert bad.is_cuda and bad.dtype == torch.uint8 and bad.numel() == self.num_envs
        self._check(self.L.b200sim_check_state(self.h, bad.data_ptr(), rest_record.data_ptr() if rest_record is not None else None,
                                               ctypes.byref(keep) if keep is not None else None, self._stream()))

    def compute_reward(self, ag, dg):
        ag = ag.to(self.device, torch.float32).contiguous().reshape(-1, self.ngoal)
        dg = dg.to(self.device, torch.float32).contiguous().reshape(-1, self.ngoal)
        out = torch.empty(ag.shape[0], dtype=torch.float32, device=self.device)
        self._check(self.L.b200sim_compute_reward(self.h, ag.data_ptr(), dg.data_ptr(), ag.shape[0], out.data_ptr(), self._stream()))
        return out

    @property
    def launches(self):
        return int(self.L.b200sim_launch_count(self.h))


class FetchVectorEnv(CtorPickle):
    """`gym.make_vec("FetchPickAndPlace-v4", num_envs=N)` replacement.  Observations, rewards and flags are torch
    tensors on `device` (float32 / bool) with a leading `num_envs` axis."""

    metadata = {"render_modes": [], "render_fps": 25, "autoreset_mode": "next_step"}

    def __init__(self, task: str = "FetchPickAndPlace", num_envs: int = 1, reward_type: str = "sparse",
                 max_episode_steps: Optional[int] = 50, device="cuda:0", rng_mode: str = "auto",
                 autoreset_mode: str = "next_step", n_substeps: int = N_SUBSTEPS, backend_factory=None, **kwargs):
        if task not in FETCH_TASKS:
            raise KeyError(f"unknown Fetch task {task!r}")
        if reward_type not in ("sparse", "dense"):
            raise ValueError("reward_type must be 'sparse' or 'dense'")
        if autoreset_mode not in ("next_step", "same_step", "disabled"):
            raise ValueError("autoreset_mode must be next_step, same_step or disabled")
        if kwargs.get("render_mode") is not None:
            raise NotImplementedError("rendering is out of scope for the batched CUDA path")
        cfg = dict(FETCH_TASKS[task])
        self.task_name, self.cfg, self.reward_type = task, cfg, reward_type
        self.num_envs, self.max_episode_steps, self.autoreset_mode = int(num_envs), max_episode_steps, autoreset_mode
        self.metadata = dict(self.metadata, autoreset_mode=autoreset_mode)
        self.n_substeps = n_substeps
        self.model = load_model(cfg["model"])
        self.task = make_task_struct(self.model, cfg, reward_type, n_substeps)
        eq = welded_eq_data(self.model)
        factory = backend_factory or CudaBackend
        self.backend = factory(self.model, eq, self.task, self.num_envs, device)
        self.device = self.backend.device
        # "numpy": per-env PCG64 streams in the reference's draw order (value-equal resets); "torch": torch's device generator;
        # "device": the draws happen inside the library (b200sim_reset, csrc/reset_sample.cuh) -- no host work per reset
        if rng_mode not in ("auto", "numpy", "torch", "device"):
            raise ValueError("rng_mode must be auto, numpy, torch or device")
        self.rng_mode = rng_mode if rng_mode != "auto" else ("numpy" if self.num_envs <= 64 else "torch")
        self.env_offset = int(kwargs.get("env_offset", 0))   # global index of env 0 (sharded runs, sharding.py)
        # opt-in failure detection: after every step the state records are scanned for NaN / huge values and such envs are put
        # back to their rest state with the goal kept ([ext] mj_checkPos / mj_checkVel / mj_checkAcc + mj_resetData inside mj_step)
        self.auto_recover = bool(kwargs.get("auto_recover", False))
        self._np_rngs = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(None))) for _ in range(self.num_envs)] \
            if self.rng_mode == "numpy" else None
        self._gen = torch.Generator(device=self.device)
        self._gen.seed()
        self._dev_seed = int(self._gen.initial_seed())
        lay = self.backend.layout
        self._sl = {k: slice(lay[k], lay[k] + n) for k, n in (("qpos", self.model.nq), ("qvel", self.model.nv), ("warm", self.model.nv),
                                                              ("ctrl", self.model.nu), ("mocap", 7), ("pose", 7), ("goal", 3))}
        self.dt = float(self.model.opt[0] * n_substeps)
        nobs = self.task.nobs
        self.single_action_space = Box(-1.0, 1.0, shape=(4,), dtype=np.float32)
        self.single_observation_space = DictSpace(dict(
            desired_goal=Box(-np.inf, np.inf, shape=(3,), dtype=np.float64),
            achieved_goal=Box(-np.inf, np.inf, shape=(3,), dtype=np.float64),
            observation=Box(-np.inf, np.inf, shape=(nobs,), dtype=np.float64)))
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        # TimeLimit and the terminated / truncated flags are computed by the step kernel (b200sim_set_time_limit); the per-env
        # step counters live in the library and are visible here as a tensor
        self._elapsed = self.backend.elapsed
        self.backend.set_time_limit(max_episode_steps, False)   # robot_env.py:106-112: compute_terminated is constant False
        self._needs_reset = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self._env_setup()
        self.closed = False

    # ------------------------------------------------------------------ construction
    def _env_setup(self):
        """envs/fetch/fetch_env.py:404-428 run once for every env in lock-step (all envs are identical here)."""
        m, st, sl = self.model, self.backend.state, self._sl
        qpos = np.array(m.qpos0, dtype=np.float64)
        for name, value in self.cfg["initial_qpos"].items():
            j = m.joint_id(name)
            a = int(m.jnt_qposadr[j])
            n = 7 if m.jnt_type[j] == JNT_FREE else 1
            qpos[a:a + n] = value
        st.zero_()
        st[:, sl["qpos"]] = torch.as_tensor(qpos, dtype=torch.float32, device=self.device)
        st[:, sl["mocap"]] = torch.tensor([0, 0, 0, 1, 0, 0, 0], dtype=torch.float32, device=self.device)
        out = self.backend.new_outputs()
        self.backend.refresh(None, out)  # mj_forward
        grip = out["obs"][:, 0:3]
        target = grip + torch.tensor([-0.498, 0.005, -0.431 + self.cfg["gripper_extra_height"]], dtype=torch.float32, device=self.device)
        st[:, sl["mocap"]] = torch.cat([target, torch.tensor([1.0, 0.0, 1.0, 0.0], device=self.device).expand(self.num_envs, 4)], dim=1)
        self.backend.raw_step(10 * self.n_substeps, out)  # 10 x mj_step(nstep=n_substeps)
        # site positions as of the last forward pass (the reference reads data.site_xpos without a new mj_forward)
        self.initial_gripper_xpos = out["obs"][0, 0:3].clone()
        self.height_offset = float(out["obs"][0, 5]) if self.cfg["has_object"] else None
        self.initial_qpos = st[0, sl["qpos"]].clone()
        self.initial_qvel = st[0, sl["qvel"]].clone()
        self._mocap_rest = torch.tensor([0, 0, 0, 1, 0, 0, 0], dtype=torch.float32, device=self.device)
        self._obj_qadr = int(m.jnt_qposadr[m.joint_id("object0:joint")]) if self.cfg["has_object"] else -1
        self._last = out

    # ------------------------------------------------------------------ sampling
    def _sample_reset(self, idx):
        """Object start position (fetch_env.py:386-399) and goal (fetch_env.py:153-166) for the envs in `idx`."""
        cfg, n = self.cfg, idx.numel()
        g0 = self.initial_gripper_xpos
        off = cfg["target_offset"]
        if self.rng_mode == "numpy":
            g0n = g0.double().cpu().numpy()
            obj = np.zeros((n, 2))
            goals = np.zeros((n, 3))
            for k, i in enumerate(idx.tolist()):
                rng = self._np_rngs[i]
                if cfg["has_object"]:
                    xy = g0n[:2]
                    while np.linalg.norm(xy - g0n[:2]) < 0.1:
                        xy = g0n[:2] + rng.uniform(-cfg["obj_range"], cfg["obj_range"], size=2)
                    obj[k] = xy
                    goal = g0n[:3] + rng.uniform(-cfg["target_range"], cfg["target_range"], size=3)
                    goal += off
                    goal[2] = self.height_offset
                    if cfg["target_in_the_air"] and rng.uniform() < 0.5:
                        goal[2] += rng.uniform(0, 0.45)
                else:
                    goal = g0n[:3] + rng.uniform(-cfg["target_range"], cfg["target_range"], size=3)
                goals[k] = goal
            return (torch.as_tensor(obj, dtype=torch.float32, device=self.device) if cfg["has_object"] else None,
                    torch.as_tensor(goals, dtype=torch.float32, device=self.device))
        # device RNG (Philox): same distributions, different stream
        u = lambda *s: torch.rand(*s, generator=self._gen, device=self.device)
        obj = None
        if cfg["has_object"]:
            # rejection sampling (fetch_env.py:386-392) without a host round trip: masked redraws until a rejected sample is
            # left with probability p^(k+1) < 1e-11 per env, p = pi 0.1^2 / (2 obj_range)^2 -- 24 redraws for obj_range 0.15
            # (p = 0.35), 105 for FetchSlide's 0.1 (p = 0.79; 24 would leave 0.2 % of its resets inside the excluded disc)
            p_rej = min(np.pi * 0.01 / (2 * cfg["obj_range"]) ** 2, 0.999)
            redraws = 24 if p_rej < 0.36 else int(np.ceil(np.log(1e-11) / np.log(p_rej)))
            obj = g0[:2] + (u(n, 2) * 2 - 1) * cfg["obj_range"]
            for _ in range(redraws):
                bad = torch.linalg.norm(obj - g0[:2], dim=1) < 0.1
                obj = torch.where(bad[:, None], g0[:2] + (u(n, 2) * 2 - 1) * cfg["obj_range"], obj)
        goals = g0[:3] + (u(n, 3) * 2 - 1) * cfg["target_range"]
        if cfg["has_object"]:
            goals = goals + torch.as_tensor(off, dtype=torch.float32, device=self.device)
            goals[:, 2] = self.height_offset
            if cfg["target_in_the_air"]:
                air = u(n) < 0.5
                goals[:, 2] += torch.where(air, u(n) * 0.45, torch.zeros(n, device=self.device))
        return obj, goals

    def _mask_indices(self, mask):
        """Indices of the envs in `mask`; no device round trip when the caller knows that every env is due (`_reset_all`)."""
        if getattr(self, "_reset_all", False):
            if getattr(self, "_all_idx", None) is None or self._all_idx.numel() != self.num_envs:
                self._all_idx = torch.arange(self.num_envs, device=self.device)
            return self._all_idx
        return torch.nonzero(mask, as_tuple=False).flatten()

    def _device_reset_params(self):
        from ._lib import FetchResetC

        cfg, p = self.cfg, FetchResetC()
        p.has_object, p.target_in_the_air, p.obj_qadr = int(cfg["has_object"]), int(cfg["target_in_the_air"]), max(self._obj_qadr, 0)
        p.obj_range, p.target_range = float(cfg.get("obj_range", 0.0)), float(cfg["target_range"])
        g0 = self.initial_gripper_xpos.cpu().tolist()
        for k in range(3):
            p.target_offset[k] = float(np.broadcast_to(np.asarray(cfg["target_offset"], dtype=np.float64), (3,))[k])
            p.gripper_xpos[k] = float(g0[k])
        p.height_offset = float(self.height_offset or 0.0)
        rest = torch.zeros(self.backend.state.shape[1], dtype=torch.float32, device=self.device)   # mj_resetData
        rest[self._sl["qpos"]] = self.initial_qpos
        rest[self._sl["qvel"]] = self.initial_qvel
        rest[self._sl["mocap"]] = self._mocap_rest
        return p, rest

    def _recovery_record(self):
        """(rest record, ranges of the state record a recovered env keeps) for b200sim_check_state."""
        from ._lib import KeepC

        keep = KeepC()
        keep.n, keep.start[0], keep.len[0] = 1, self._sl["goal"].start, self._sl["goal"].stop - self._sl["goal"].start
        return self._device_reset_params()[1], keep

    def _check_and_recover(self, out, info):
        if getattr(self, "_recovery", None) is None:
            self._recovery = self._recovery_record()
            self._bad = torch.zeros(self.num_envs, dtype=torch.uint8, device=self.device)
            self.bad_state_count = torch.zeros((), dtype=torch.int64, device=self.device)
        rest, keep = self._recovery
        self.backend.check_state(self._bad, rest, keep)
        self.backend.refresh(self._bad, out)          # mj_forward + _get_obs of the recovered envs (none, almost always)
        bad = self._bad.bool()
        self.bad_state_count += bad.sum()
        info["bad_state"] = bad

    def _reset_envs(self, mask, out):
        if self.rng_mode == "device":
            if getattr(self, "_dev_reset", None) is None:
                self._dev_reset = self._device_reset_params()
                self._episode = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
            every = getattr(self, "_reset_all", False)
            p, rest = self._dev_reset
            self.backend.reset_draw(None if every else mask.to(torch.uint8), rest, p, self._dev_seed, self.env_offset, self._episode, out)
            if every:
                self._elapsed.zero_()
            else:
                self._elapsed.masked_fill_(mask, 0)
            return
        idx = self._mask_indices(mask)
        if idx.numel() == 0:
            return
        st, sl = self.backend.state, self._sl
        obj, goals = self._sample_reset(idx)
        rec = torch.zeros((idx.numel(), st.shape[1]), dtype=torch.float32, device=self.device)  # mj_resetData
        rec[:, sl["qpos"]] = self.initial_qpos
        rec[:, sl["qvel"]] = self.initial_qvel
        rec[:, sl["mocap"]] = self._mocap_rest
        if obj is not None:
            rec[:, self._sl["qpos"].start + self._obj_qadr: self._sl["qpos"].start + self._obj_qadr + 2] = obj
        rec[:, sl["goal"]] = goals
        st[idx] = rec
        self._elapsed[idx] = 0
        self.backend.refresh(mask.to(torch.uint8), out)  # mj_forward + _get_obs for the reset envs

    # ------------------------------------------------------------------ gymnasium API
    def _obs_dict(self, out):
        return self._cast_obs({"observation": out["obs"], "achieved_goal": out["achieved"], "desired_goal": out["desired"]})

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            seeds = [seed + i for i in range(self.num_envs)] if isinstance(seed, (int, np.integer)) else list(seed)
            if self.rng_mode == "numpy":
                self._np_rngs = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(s))) for s in seeds]
            self._gen.manual_seed(int(seeds[0]))
            if self.rng_mode == "device":   # one key for the batch; env index and episode counter select the stream
                self._dev_seed = int(seeds[0])
                if getattr(self, "_episode", None) is not None:
                    self._episode.zero_()
        out = self.backend.new_outputs()
        mask = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
        self._reset_all = True
        try:
            self._reset_envs(mask, out)
        finally:
            self._reset_all = False
        self._needs_reset.zero_()
        self._elapsed_ub, self._pending_reset = 0, False
        self._in_phase = True   # every env has the same step count (these envs never terminate): TimeLimit is host-known
        self._last = out
        return self._obs_dict(out), {}

    def step(self, actions):
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float32))
        if tuple(actions.shape) != (self.num_envs, self.single_action_space.shape[0]):
            raise ValueError("Action dimension mismatch")
        actions = actions.to(self.device, torch.float32, non_blocking=True).contiguous()
        out = self.backend.new_outputs()
        # clip + _set_action + n_substeps x mj_step + _get_obs + reward + TimeLimit / terminated / truncated: one kernel
        self.backend.step(actions, out)
        self._elapsed_ub = getattr(self, "_elapsed_ub", 0) + 1   # host-side upper bound of max(_elapsed): no sync on most steps
        reward, success = out["reward"], out["success"]
        terminated, truncated = out["terminated"], out["truncated"]
        # solver_info: Newton iterations (low 16 bits) | capacity-overflow flags << 16 of this step, per env (the backend's
        # persistent tensor: valid until the next step); solver_overflow_count: device counter over the env's lifetime
        info = {"is_success": success, "solver_info": self.backend.info}
        if getattr(self, "auto_recover", False):
            self._check_and_recover(out, info)
        if getattr(self, "_const_true", None) is None or self._const_true.numel() != self.num_envs:
            self._const_true = torch.ones(self.num_envs, dtype=torch.bool, device=self.device)
        in_phase = getattr(self, "_in_phase", False)
        if self.autoreset_mode == "next_step" and getattr(self, "_pending_reset", False):
            self._pending_reset = False
            if in_phase or bool(self._needs_reset.any()):
                self._reset_all = in_phase
                # envs that finished on the previous call are reset now; their action is ignored (gymnasium NEXT_STEP)
                pre = self._needs_reset.clone()
                self._reset_envs(pre, out)
                self._mask_step_outputs(out, pre)
                self._needs_reset.zero_()
                self._reset_all = False
                self._elapsed_ub = 0 if in_phase else int(self._elapsed.max())
        # TimeLimit: the flags come from the kernel; the host only looks at them (and synchronises) once the bound says an env may be due
        may_truncate = self.max_episode_steps is not None and self._elapsed_ub >= self.max_episode_steps
        if may_truncate:
            done = truncated | terminated
            if self.autoreset_mode == "next_step":
                self._needs_reset = done
                self._pending_reset = True
            elif self.autoreset_mode == "same_step":
                # in phase (all envs reset together and none terminates): the bound IS every env's step count -- no device read
                if in_phase or bool(done.any()):
                    fo = self._obs_dict(out)
                    info["final_obs"] = {k: v.clone() for k, v in fo.items()} if isinstance(fo, dict) else fo.clone()
                    info["_final_obs"] = done.clone()
                    # gymnasium's SAME_STEP convention: the info of the finished episodes next to their last observation
                    info["final_info"] = {"is_success": success.clone(), "_is_success": done.clone()}
                    info["_final_info"] = done.clone()
                    self._reset_all = in_phase
                    self._reset_envs(done, out)
                    self._reset_all = False
                self._elapsed_ub = 0 if in_phase else int(self._elapsed.max())
        info["_is_success"] = self._const_true
        self._last = out
        return self._obs_dict(out), reward, terminated, truncated, info

    def _mask_step_outputs(self, out, pre):
        """NEXT_STEP autoreset: the envs in `pre` were reset instead of stepped -- reward 0, no success, no flags (in place, so the
        packed row stays the single source of the step's results)."""
        k = self.backend.nobs + 2 * self.backend.ngoal
        out["packed"][:, k:k + 4].masked_fill_(pre[:, None], 0.0)
        out["flags"].masked_fill_(pre[None, :], 0)

    @property
    def solver_overflow_count(self):
        """Env-steps so far in which a capacity limit (broad-phase candidates, contacts, contact groups, limit rows) dropped
        something (DESIGN.md deviation 5); reads the device counter (synchronises)."""
        return int(self.backend.overflow_counter[0])

    # GoalEnv API (core.py:45-114), batched; accepts numpy or torch, any leading shape
    def compute_reward(self, achieved_goal, desired_goal, info=None):
        is_np = not torch.is_tensor(achieved_goal)
        ag = torch.as_tensor(np.asarray(achieved_goal)) if is_np else achieved_goal
        dg = torch.as_tensor(np.asarray(desired_goal)) if not torch.is_tensor(desired_goal) else desired_goal
        lead = ag.shape[:-1]
        r = self.backend.compute_reward(ag, dg).reshape(lead)
        if is_np:
            r = r.cpu().numpy()
            return r.astype(np.float32) if self.reward_type == "sparse" else r.astype(np.float64)
        return r

    def compute_terminated(self, achieved_goal, desired_goal, info=None):
        return False

    def compute_truncated(self, achieved_goal, desired_goal, info=None):
        return False

    # state access (checkpoint / parity injection), SURVEY.md section 5
    def get_state(self):
        return self.backend.state.clone(), self._elapsed.clone()

    def set_state(self, state, elapsed=None):
        self.backend.state.copy_(state)
        if elapsed is not None:
            self._elapsed.copy_(elapsed)
            self._in_phase = False
        self._elapsed_ub = int(self._elapsed.max())
        out = self.backend.new_outputs()
        self.backend.refresh(None, out)
        self._last = out
        return self._obs_dict(out)

    def close(self):
        if not getattr(self, "closed", True):
            self.backend.close()
            self.closed = True
