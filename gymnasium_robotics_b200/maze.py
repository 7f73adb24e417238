"""Batched AntMaze (v5) and PointMaze (v3) environments on the b200sim CUDA path.

Host-side mirror of the reference's maze stack, batched over `num_envs`:
  * map tables               envs/maze/maps.py:52-135 (data), ids/kwargs/max_episode_steps __init__.py:839-958
  * Maze grid math           envs/maze/maze_v4.py:135-146 (cell<->xy), :148-242 (goal/reset cell collection)
  * MazeEnv.reset / noise / update_goal     envs/maze/maze_v4.py:276-297, 299-379, 400-418
  * AntMazeEnv ctor/reset/step/_get_obs     envs/maze/ant_maze_v5.py:221-320 (inner AntEnv [ext]: frame_skip 5, RK4)
  * PointMazeEnv / PointEnv                 envs/maze/point_maze.py:316-434, envs/maze/point.py:22-77 (frame_skip 1, Euler)
  * compute_reward / compute_terminated     envs/maze/maze_v4.py:381-398
The per-step arithmetic (5 RK4 sub-steps, observation, reward, success) runs inside one CUDA kernel launch.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from .fetch import CudaBackend
from .models import load_model
from .rollout import CtorPickle
from .spaces import Box, Dict as DictSpace, batch_space

R, G, C = "r", "g", "c"
MAPS = {
    "Open": [[1, 1, 1, 1, 1, 1, 1], [1, 0, 0, 0, 0, 0, 1], [1, 0, 0, 0, 0, 0, 1], [1, 0, 0, 0, 0, 0, 1], [1, 1, 1, 1, 1, 1, 1]],
    "Open_Diverse_G": [[1, 1, 1, 1, 1, 1, 1], [1, R, G, G, G, G, 1], [1, G, G, G, G, G, 1], [1, G, G, G, G, G, 1], [1, 1, 1, 1, 1, 1, 1]],
    "Open_Diverse_GR": [[1, 1, 1, 1, 1, 1, 1], [1, C, C, C, C, C, 1], [1, C, C, C, C, C, 1], [1, C, C, C, C, C, 1], [1, 1, 1, 1, 1, 1, 1]],
    "UMaze": [[1, 1, 1, 1, 1], [1, 0, 0, 0, 1], [1, 1, 1, 0, 1], [1, 0, 0, 0, 1], [1, 1, 1, 1, 1]],
    "Medium": [[1, 1, 1, 1, 1, 1, 1, 1], [1, 0, 0, 1, 1, 0, 0, 1], [1, 0, 0, 1, 0, 0, 0, 1], [1, 1, 0, 0, 0, 1, 1, 1],
               [1, 0, 0, 1, 0, 0, 0, 1], [1, 0, 1, 0, 0, 1, 0, 1], [1, 0, 0, 0, 1, 0, 0, 1], [1, 1, 1, 1, 1, 1, 1, 1]],
    "Medium_Diverse_G": [[1, 1, 1, 1, 1, 1, 1, 1], [1, R, 0, 1, 1, 0, 0, 1], [1, 0, 0, 1, 0, 0, G, 1], [1, 1, 0, 0, 0, 1, 1, 1],
                         [1, 0, 0, 1, 0, 0, 0, 1], [1, G, 1, 0, 0, 1, 0, 1], [1, 0, 0, 0, 1, G, 0, 1], [1, 1, 1, 1, 1, 1, 1, 1]],
    "Medium_Diverse_GR": [[1, 1, 1, 1, 1, 1, 1, 1], [1, C, 0, 1, 1, 0, 0, 1], [1, 0, 0, 1, 0, 0, C, 1], [1, 1, 0, 0, 0, 1, 1, 1],
                          [1, 0, 0, 1, 0, 0, 0, 1], [1, C, 1, 0, 0, 1, 0, 1], [1, 0, 0, 0, 1, C, 0, 1], [1, 1, 1, 1, 1, 1, 1, 1]],
    "Large": [[1] * 12, [1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 1], [1, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 1], [1, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 1],
              [1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1], [1, 0, 0, 1, 0, 1, 0, 0, 0, 0, 0, 1], [1, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 1],
              [1, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1], [1] * 12],
    "Large_Diverse_G": [[1] * 12, [1, R, 0, 0, 0, 1, G, 0, 0, 0, 0, 1], [1, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 1], [1, 0, 0, 0, 0, G, 0, 1, 0, 0, G, 1],
                        [1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1], [1, 0, G, 1, 0, 1, 0, 0, 0, 0, 0, 1], [1, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 1],
                        [1, 0, 0, 1, G, 0, G, 1, 0, G, 0, 1], [1] * 12],
    "Large_Diverse_GR": [[1] * 12, [1, C, 0, 0, 0, 1, C, 0, 0, 0, 0, 1], [1, 0, 1, 1, 0, 1, 0, 1, 0, 1, 0, 1], [1, 0, 0, 0, 0, C, 0, 1, 0, 0, C, 1],
                         [1, 0, 1, 1, 1, 1, 0, 1, 1, 1, 0, 1], [1, 0, C, 1, 0, 1, 0, 0, 0, 0, 0, 1], [1, 1, 0, 1, 0, 1, 0, 1, 0, 1, 1, 1],
                         [1, 0, 0, 1, C, 0, C, 1, 0, C, 0, 1], [1] * 12],
}
# wall layout -> compiled physics model (the diverse variants only relabel free cells)
NOISE, SUCCESS_RADIUS = 0.25, 0.45
# per agent: maze_size_scaling, maze_height, frame_skip, first qpos entry in `observation`, velocity clip, episode lengths
AGENTS = {
    "ant": dict(scaling=4.0, height=0.5, frame_skip=5, obs_qpos_start=2, vel_clip=0.0, fps=50,
                steps={k: (700 if k.startswith(("Open", "UMaze")) else 1000) for k in MAPS}),     # __init__.py:839-958
    "point": dict(scaling=1.0, height=0.4, frame_skip=1, obs_qpos_start=0, vel_clip=5.0, fps=100,
                  steps={k: (300 if k.startswith(("Open", "UMaze")) else (600 if k.startswith("Medium") else 800)) for k in MAPS}),  # :960-1080
}
SCALING, HEIGHT, FRAME_SKIP = AGENTS["ant"]["scaling"], AGENTS["ant"]["height"], AGENTS["ant"]["frame_skip"]


def model_name(agent, maze):
    """wall layout -> compiled physics model (the diverse variants only relabel free cells)"""
    return f"{agent}maze_" + maze.split("_")[0].lower()


class MazeCells:
    """Cell bookkeeping of `Maze` (maze_v4.py:26-242) without the XML part."""

    def __init__(self, maze_map, scaling=SCALING):
        self.maze_map, self.scaling = maze_map, scaling
        self.length, self.width = len(maze_map), len(maze_map[0])
        self.x_center, self.y_center = self.width / 2 * scaling, self.length / 2 * scaling
        goals, resets, combined, empty = [], [], [], []
        for i in range(self.length):
            for j in range(self.width):
                cell, xy = maze_map[i][j], self.cell_rowcol_to_xy((i, j))
                if cell == R:
                    resets.append(xy)
                elif cell == G:
                    goals.append(xy)
                elif cell == C:
                    combined.append(xy)
                elif cell == 0:
                    empty.append(xy)
        if not goals and not resets and not combined:
            combined = empty
        elif not resets and not combined:
            resets = empty
        elif not goals and not combined:
            goals = empty
        self.goal_locations = np.array(goals + combined)
        self.reset_locations = np.array(resets + combined)

    def cell_rowcol_to_xy(self, rowcol):
        return np.array([(rowcol[1] + 0.5) * self.scaling - self.x_center, self.y_center - (rowcol[0] + 0.5) * self.scaling])

    def cell_xy_to_rowcol(self, xy):
        return np.array([math.floor((self.y_center - xy[1]) / self.scaling), math.floor((xy[0] + self.x_center) / self.scaling)])


def make_maze_task(model, reward_type, agent="ant", contact_forces=False):
    """contact_forces: append Ant-v5's clipped `cfrc_ext[1:]` (6 values per body) to the observation -- (105,) instead of (27,)"""
    cfg = AGENTS[agent]
    t = _lib.FetchTaskC()
    t.kind, t.nact, t.ngoal = 1, int(model.nu), 2
    t.n_substeps, t.reward_dense = cfg["frame_skip"], int(reward_type == "dense")
    t.obs_qpos_start, t.vel_clip = cfg["obs_qpos_start"], cfg["vel_clip"]
    t.touch_mode = int(bool(contact_forces))
    t.nobs = int(model.nq - cfg["obs_qpos_start"] + model.nv) + (6 * (len(model.mjbody_rt) - 1) if contact_forces else 0)
    t.success_radius = SUCCESS_RADIUS
    t.dt = float(model.opt[0] * cfg["frame_skip"])
    return t


def make_antmaze_task(model, reward_type):
    return make_maze_task(model, reward_type, "ant")


class _AntBackend(CudaBackend):
    """CudaBackend of the maze agents (goal dim 2; the action dim comes from the task)."""


class MazeVectorEnv(CtorPickle):
    """`gym.make_vec("AntMaze_Large-v5" | "PointMaze_UMaze-v3", num_envs=N)` replacement (torch CUDA tensors, leading
    `num_envs` axis).  `maze_map` may be a name from `MAPS` or (for the point agent's tests) an explicit cell list."""

    metadata = {"render_modes": [], "render_fps": 50, "autoreset_mode": "next_step"}
    AGENT = "ant"

    def __init__(self, maze="Large", num_envs: int = 1, reward_type: str = "sparse", continuing_task: bool = True,
                 reset_target: bool = False, max_episode_steps: Optional[int] = None, device="cuda:0", rng_mode: str = "auto",
                 autoreset_mode: str = "next_step", backend_factory=None, agent: Optional[str] = None, model=None,
                 include_cfrc_ext_in_observation: bool = False, **kwargs):
        self.agent = agent or self.AGENT
        cfg = AGENTS[self.agent]
        if isinstance(maze, str) and maze not in MAPS:
            raise KeyError(f"unknown maze {maze!r}")
        if reward_type not in ("sparse", "dense"):
            raise ValueError("reward_type must be 'sparse' or 'dense'")
        if kwargs.get("render_mode") is not None:
            raise NotImplementedError("rendering is out of scope for the batched CUDA path")
        if autoreset_mode not in ("next_step", "same_step", "disabled"):
            raise ValueError("autoreset_mode must be next_step, same_step or disabled")
        self.maze_name, self.reward_type = maze, reward_type
        self.continuing_task, self.reset_target = continuing_task, reset_target
        self.num_envs, self.autoreset_mode = int(num_envs), autoreset_mode
        self.metadata = dict(self.metadata, autoreset_mode=autoreset_mode)
        self.scaling, self.frame_skip = cfg["scaling"], cfg["frame_skip"]
        self.metadata["render_fps"] = cfg["fps"]
        named = isinstance(maze, str)
        self.max_episode_steps = (cfg["steps"][maze] if named else None) if max_episode_steps is None else max_episode_steps
        self.cells = MazeCells(MAPS[maze] if named else maze, cfg["scaling"])
        if model is None and not named:
            raise ValueError("an explicit maze map needs its compiled `model` (see models.compile_maze_model)")
        self.model = model if model is not None else load_model(model_name(self.agent, maze))
        # Ant-v5 keyword [ext]: AntMaze_*-v5 observes the clipped per-body contact forces (ant_maze_v5.py:99: (105,) = 27 + 13 x 6);
        # AntMaze_*-v4 (Ant-v4, use_contact_forces False) and the point agent do not.  The registry sets it per id.
        self.include_cfrc = bool(include_cfrc_ext_in_observation) and self.agent == "ant"
        self.task = make_maze_task(self.model, reward_type, self.agent, self.include_cfrc)
        factory = backend_factory or _AntBackend
        self.backend = factory(self.model, np.zeros((0, 11)), self.task, self.num_envs, device)
        self.device = self.backend.device
        # "device": goal / reset cells and their noise are drawn inside the library (b200sim_reset_maze, csrc/reset_sample.cuh)
        self.rng_mode = rng_mode if rng_mode != "auto" else ("numpy" if self.num_envs <= 64 else "torch")
        self.env_offset = int(kwargs.get("env_offset", 0))
        self._np_rngs = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(None))) for _ in range(self.num_envs)]
        self._gen = torch.Generator(device=self.device)
        self._gen.seed()
        self._dev_seed = int(self._gen.initial_seed())
        lay, m = self.backend.layout, self.model
        self._sl = {k: slice(lay[k], lay[k] + n) for k, n in (("qpos", m.nq), ("qvel", m.nv), ("warm", m.nv), ("ctrl", m.nu), ("goal", 2))}
        nobs = self.task.nobs
        self.single_action_space = Box(-1.0, 1.0, shape=(m.nu,), dtype=np.float32)
        self.single_observation_space = DictSpace(dict(
            observation=Box(-np.inf, np.inf, shape=(nobs,), dtype=np.float64),
            achieved_goal=Box(-np.inf, np.inf, shape=(2,), dtype=np.float64),
            desired_goal=Box(-np.inf, np.inf, shape=(2,), dtype=np.float64)))
        self.action_space = batch_space(self.single_action_space, self.num_envs)
        self.observation_space = batch_space(self.single_observation_space, self.num_envs)
        self.init_qpos = torch.as_tensor(np.array(m.qpos0), dtype=torch.float32, device=self.device)
        self._goal_loc = torch.as_tensor(self.cells.goal_locations, dtype=torch.float32, device=self.device)
        self._reset_loc = torch.as_tensor(self.cells.reset_locations, dtype=torch.float32, device=self.device)
        # TimeLimit and compute_terminated (maze_v4.py:390-398: success ends the episode unless continuing_task) run inside the
        # step kernel; the step counters are library memory
        self._elapsed = self.backend.elapsed
        self.backend.set_time_limit(self.max_episode_steps, not continuing_task)
        self._needs_reset = torch.zeros(self.num_envs, dtype=torch.bool, device=self.device)
        self.dt = float(m.opt[0] * self.frame_skip)
        self.closed = False

    # ------------------------------------------------------------------ sampling (MazeEnv.reset, maze_v4.py:299-358)
    def _noise_np(self, rng, xy):
        nx = rng.uniform(low=-NOISE, high=NOISE) * self.scaling
        ny = rng.uniform(low=-NOISE, high=NOISE) * self.scaling
        return np.array([xy[0] + nx, xy[1] + ny])

    def _sample_np(self, i, options):
        rng, cells = self._np_rngs[i], self.cells
        options = options or {}
        if options.get("goal_cell") is not None:
            gc = options["goal_cell"]
            assert cells.length > gc[0] and cells.width > gc[1] and cells.maze_map[gc[0]][gc[1]] != 1, f"Goal can't be placed in a wall cell, {gc}"
            goal = cells.cell_rowcol_to_xy(gc)
        else:
            goal = cells.goal_locations[rng.integers(low=0, high=len(cells.goal_locations))].copy()
        goal = self._noise_np(rng, goal)
        if options.get("reset_cell") is not None:
            rc = options["reset_cell"]
            assert cells.length > rc[0] and cells.width > rc[1] and cells.maze_map[rc[0]][rc[1]] != 1, f"Reset can't be placed in a wall cell, {rc}"
            pos = cells.cell_rowcol_to_xy(rc)
        else:
            pos = goal.copy()
            while np.linalg.norm(pos - goal) <= 0.5 * self.scaling:
                pos = cells.reset_locations[rng.integers(low=0, high=len(cells.reset_locations))].copy()
        return goal, self._noise_np(rng, pos)

    def _sample(self, idx, options=None):
        n = idx.numel()
        if self.rng_mode == "numpy" or options:  # explicit cells always go through the reference-ordered numpy path
            gs, ps = zip(*[self._sample_np(i, options) for i in idx.tolist()])
            return (torch.as_tensor(np.array(gs), dtype=torch.float32, device=self.device),
                    torch.as_tensor(np.array(ps), dtype=torch.float32, device=self.device))
        u = lambda *s: torch.rand(*s, generator=self._gen, device=self.device)
        ri = lambda hi, k: torch.randint(0, hi, (k,), generator=self._gen, device=self.device)
        goal = self._goal_loc[ri(len(self._goal_loc), n)] + (u(n, 2) * 2 - 1) * NOISE * self.scaling
        pos = self._reset_loc[ri(len(self._reset_loc), n)]
        bad = torch.linalg.norm(pos - goal, dim=1) <= 0.5 * self.scaling
        while bool(bad.any()):
            pos[bad] = self._reset_loc[ri(len(self._reset_loc), int(bad.sum()))]
            bad = torch.linalg.norm(pos - goal, dim=1) <= 0.5 * self.scaling
        return goal, pos + (u(n, 2) * 2 - 1) * NOISE * self.scaling

    def _device_reset(self, mask, out):
        if getattr(self, "_dev_reset", None) is None:
            from ._lib import MazeResetC

            p = MazeResetC()
            p.n_goal, p.n_reset, p.scaling, p.noise = len(self._goal_loc), len(self._reset_loc), float(self.scaling), float(NOISE)
            rest = torch.zeros(self.backend.state.shape[1], dtype=torch.float32, device=self.device)
            rest[self._sl["qpos"]] = self.init_qpos
            self._dev_reset = (p, rest, self._goal_loc.contiguous(), self._reset_loc.contiguous())
            self._episode = torch.zeros(self.num_envs, dtype=torch.int32, device=self.device)
        p, rest, gl, rl = self._dev_reset
        self.backend.reset_maze(mask.to(torch.uint8), rest, p, gl, rl, self._dev_seed, self.env_offset, self._episode, out)
        self._elapsed.masked_fill_(mask, 0)

    def _reset_envs(self, mask, out, options=None):
        if self.rng_mode == "device" and not options:   # explicit cells keep the reference-ordered host path
            return self._device_reset(mask, out)
        idx = torch.nonzero(mask, as_tuple=False).flatten()
        if idx.numel() == 0:
            return
        st, sl = self.backend.state, self._sl
        goal, pos = self._sample(idx, options)
        rec = torch.zeros((idx.numel(), st.shape[1]), dtype=torch.float32, device=self.device)
        rec[:, sl["qpos"]] = self.init_qpos          # ant_env.init_qpos with [:2] = reset_pos (ant_maze_v5.py:285)
        rec[:, sl["qpos"].start: sl["qpos"].start + 2] = pos
        rec[:, sl["goal"]] = goal
        st[idx] = rec
        self._elapsed[idx] = 0
        self.backend.refresh(mask.to(torch.uint8), out)

    # ------------------------------------------------------------------ gymnasium API
    def _obs_dict(self, out):
        return self._cast_obs({"observation": out["obs"], "achieved_goal": out["achieved"], "desired_goal": out["desired"]})

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            seeds = [seed + i for i in range(self.num_envs)] if isinstance(seed, (int, np.integer)) else list(seed)
            self._np_rngs = [np.random.Generator(np.random.PCG64(np.random.SeedSequence(s))) for s in seeds]
            self._gen.manual_seed(int(seeds[0]))
            self._dev_seed = int(seeds[0])
            if getattr(self, "_episode", None) is not None:
                self._episode.zero_()
        out = self.backend.new_outputs()
        self._reset_envs(torch.ones(self.num_envs, dtype=torch.bool, device=self.device), out, options)
        self._needs_reset.zero_()
        self._elapsed_ub, self._pending_reset = 0, False
        self._last = out
        return self._obs_dict(out), {"success": out["success"] > 0}

    def step(self, actions):
        if not torch.is_tensor(actions):
            actions = torch.as_tensor(np.asarray(actions, dtype=np.float32))
        if tuple(actions.shape) != (self.num_envs, self.model.nu):
            raise ValueError("Action dimension mismatch")
        actions = actions.to(self.device, torch.float32, non_blocking=True).contiguous()
        out = self.backend.new_outputs()
        self.backend.step(actions, out)   # physics + observation + reward + success + terminated / truncated flags: one kernel
        self._elapsed_ub = getattr(self, "_elapsed_ub", 0) + 1   # host-side upper bound of max(_elapsed)
        reward, success = out["reward"], out["success"] > 0
        terminated, truncated = out["terminated"], out["truncated"]
        info = {"success": success, "solver_info": self.backend.info}
        # a continuing task can only end by TimeLimit: then the host knows from its step counter when a check is due and
        # does not synchronise with the device on the other steps
        lazy = self.continuing_task
        pre = None
        if self.autoreset_mode == "next_step" and (not lazy or getattr(self, "_pending_reset", True)):
            self._pending_reset = False
            if bool(self._needs_reset.any()):
                # envs that finished on the previous call are reset now: their action was ignored, so the step they did not
                # take reports reward 0, no success and no flags (gymnasium NEXT_STEP)
                pre = self._needs_reset.clone()
                self._reset_envs(pre, out)
                k = self.backend.nobs + 2 * self.backend.ngoal
                out["packed"][:, k:k + 4].masked_fill_(pre[:, None], 0.0)
                out["flags"].masked_fill_(pre[None, :], 0)
                success = success & ~pre
                info["success"] = success
                self._needs_reset.zero_()
                self._elapsed_ub = int(self._elapsed.max())
        if self.continuing_task and self.reset_target and len(self.cells.goal_locations) > 1 and bool(success.any()):
            self._update_goal(success, out)  # maze_v4.py:400-418 (never for the envs that were just reset: `success` is masked)
        may_truncate = self.max_episode_steps is not None and (not lazy or self._elapsed_ub >= self.max_episode_steps)
        if may_truncate or not lazy:
            done = truncated | terminated
            if self.autoreset_mode == "next_step":
                self._needs_reset = done
                self._pending_reset = True
            elif self.autoreset_mode == "same_step":
                if bool(done.any()):
                    info["final_obs"] = {k: v.clone() for k, v in self._obs_dict(out).items()}
                    info["_final_obs"] = done.clone()
                    # gymnasium's SAME_STEP convention: the info of the finished episodes next to their last observation
                    info["final_info"] = {"success": success.clone(), "_success": done.clone()}
                    info["_final_info"] = done.clone()
                    self._reset_envs(done, out)
                self._elapsed_ub = int(self._elapsed.max())
        self._last = out      # the packed rows of this step (obs | achieved | desired | reward | success | flags)
        return self._obs_dict(out), reward, terminated, truncated, info

    @property
    def solver_overflow_count(self):
        return int(self.backend.overflow_counter[0])

    def _update_goal(self, success, out):
        idx = torch.nonzero(success, as_tuple=False).flatten()
        st, sl = self.backend.state, self._sl
        for i in idx.tolist():
            ag = out["achieved"][i].double().cpu().numpy()
            goal = st[i, sl["goal"]].double().cpu().numpy()
            rng = self._np_rngs[i]
            while np.linalg.norm(ag - goal) <= SUCCESS_RADIUS:
                goal = self._noise_np(rng, self.cells.goal_locations[rng.integers(low=0, high=len(self.cells.goal_locations))].copy())
            st[i, sl["goal"]] = torch.as_tensor(goal, dtype=torch.float32, device=self.device)

    def compute_reward(self, achieved_goal, desired_goal, info=None):
        is_np = not torch.is_tensor(achieved_goal)
        ag = torch.as_tensor(np.asarray(achieved_goal)) if is_np else achieved_goal
        dg = torch.as_tensor(np.asarray(desired_goal)) if not torch.is_tensor(desired_goal) else desired_goal
        r = self.backend.compute_reward(ag, dg).reshape(ag.shape[:-1])
        return r.cpu().numpy().astype(np.float64) if is_np else r

    def compute_terminated(self, achieved_goal, desired_goal, info=None):
        if not self.continuing_task:
            return bool(np.linalg.norm(np.asarray(achieved_goal) - np.asarray(desired_goal)) <= SUCCESS_RADIUS)
        return False

    def compute_truncated(self, achieved_goal, desired_goal, info=None):
        return False

    def get_state(self):
        return self.backend.state.clone(), self._elapsed.clone()

    def set_state(self, state, elapsed=None):
        self.backend.state.copy_(state)
        if elapsed is not None:
            self._elapsed.copy_(elapsed)
        self._elapsed_ub = int(self._elapsed.max())
        out = self.backend.new_outputs()
        self.backend.refresh(None, out)
        return self._obs_dict(out)

    def close(self):
        if not getattr(self, "closed", True):
            self.backend.close()
            self.closed = True


class AntMazeVectorEnv(MazeVectorEnv):
    AGENT = "ant"


class PointMazeVectorEnv(MazeVectorEnv):
    AGENT = "point"
