"""CPU test backend for FetchVectorEnv: drives the WARP_W == 1 host emulation of the kernel source, one env at a time.
Lets the `-m "not gpu"` suite cover the host logic (reset sampling, autoreset, TimeLimit, spaces) and the kernel
arithmetic without a GPU.  Test infrastructure only."""
from __future__ import annotations

import numpy as np
import torch

from tests.hostsim import FetchTaskC as HostTaskC, HostSim


class HostSimBackend:
    def __init__(self, model, eq_data, task, num_envs, device):
        from gymnasium_robotics_b200.fetch import REF_POINT

        self.device = torch.device("cpu")
        self.num_envs, self.nobs = num_envs, task.nobs
        penv = int(task.penv_body) if task.kind in (4, 5, 6, 7) else -1
        self.sim = HostSim(model, eq_data=eq_data if len(eq_data) else None, ref=getattr(self, "REF", REF_POINT), penv_body=penv,
                           ngrp_cap=getattr(self, "NGRP_CAP", 0), flavor=getattr(self, "FLAVOR", None))
        t = HostTaskC()
        for name, _ in task._fields_:
            setattr(t, name, getattr(task, name))
        o = 0
        lay = {}
        fetch = task.kind == 0
        if fetch:
            t.nact, t.ngoal = 4, 3
        self.ngoal, self.nact = int(t.ngoal), int(t.nact)
        for k, n in (("qpos", model.nq), ("qvel", model.nv), ("warm", model.nv), ("ctrl", model.nu), ("mocap", 7 * model.nmocap),
                     ("pose", 7 if fetch else 0), ("goal", self.ngoal), ("penv", 7 if penv > 0 else 0)):
            lay[k] = o
            o += n
        lay["stride"] = (o + 3) & ~3
        for k in ("qpos", "qvel", "warm", "ctrl", "mocap", "pose", "goal", "stride", "penv"):
            setattr(t, "st_" + k, lay[k])
        self.task, self.layout = t, lay
        self.state = torch.zeros((num_envs, lay["stride"]), dtype=torch.float32)
        self.launches = 0
        # the same episode bookkeeping as the step kernel (csrc/step_kernel.cuh): per-env step counters, flags, info words
        self.elapsed = torch.zeros(num_envs, dtype=torch.int32)
        self.info = torch.zeros(num_envs, dtype=torch.int32)
        self.overflow_counter = torch.zeros(1, dtype=torch.int64)
        self.max_steps, self.term_on_success = 0, False
        self.packed_w = (self.nobs + 2 * self.ngoal + 4 + 3) & ~3

    def set_time_limit(self, max_episode_steps, terminate_on_success=False):
        self.max_steps, self.term_on_success = int(max_episode_steps or 0), bool(terminate_on_success)

    def close(self):
        pass

    def new_outputs(self):
        n, no, ng = self.num_envs, self.nobs, self.ngoal
        p = torch.zeros((n, self.packed_w))
        flags = torch.zeros((2, n), dtype=torch.uint8)
        k = no + 2 * ng
        return dict(packed=p, obs=p[:, :no], achieved=p[:, no:no + ng], desired=p[:, no + ng:k], reward=p[:, k], success=p[:, k + 1],
                    terminated=flags[0].view(torch.bool), truncated=flags[1].view(torch.bool), flags=flags)

    def _run(self, mode, nraw, actions, mask, out, info=None):
        st = self.state.numpy()
        k = self.nobs + 2 * self.ngoal
        for i in range(self.num_envs):
            if mask is not None and not bool(mask[i]):
                continue
            a = actions[i].numpy() if actions is not None else np.zeros(self.nact, dtype=np.float32)
            obs, ag, dg, rew, suc, it = self.sim.env_step(self.task, mode, nraw, st[i], a, self.nobs, self.ngoal)
            self.overflow_bits = getattr(self, "overflow_bits", 0) | (it >> 16)   # capacity flags of the info word
            out["obs"][i] = torch.from_numpy(obs); out["achieved"][i] = torch.from_numpy(ag); out["desired"][i] = torch.from_numpy(dg)
            out["reward"][i] = rew; out["success"][i] = suc
            if mode == 0:   # flags are written by step launches only (refresh / raw leave them alone)
                self.elapsed[i] += 1
                trunc = self.max_steps > 0 and int(self.elapsed[i]) >= self.max_steps
                term = self.term_on_success and suc != 0
                out["flags"][0, i], out["flags"][1, i] = int(term), int(trunc)
                out["packed"][i, k + 2], out["packed"][i, k + 3] = float(term), float(trunc)
            if info is not None:
                info[i] = it
            if it >> 16:
                self.overflow_counter += 1
        self.launches += 1

    def step(self, actions, out, info=None):
        self._run(0, 0, actions, None, out, self.info if info is None else info)

    def refresh(self, mask, out):
        self._run(1, 0, None, mask, out)

    def raw_step(self, nstep, out, mask=None):
        self._run(2, nstep, None, mask, out)

    def reset_draw(self, mask, rest_record, params, seed, env_offset, episode, out):
        """b200sim_reset on the emulation: the same csrc/reset_sample.cuh code, env by env, then the refresh."""
        import ctypes

        L = self.sim._L
        L.hostsim_fetch_reset_record.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p,
                                                 ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.hostsim_fetch_reset_record.restype = None
        st, rest = self.state.numpy(), rest_record.numpy()
        for i in range(self.num_envs):
            if mask is not None and not bool(mask[i]):
                continue
            L.hostsim_fetch_reset_record(ctypes.byref(params), int(seed) & 0xFFFFFFFFFFFFFFFF, i + int(env_offset), int(episode[i]), rest.ctypes.data,
                                         self.layout["stride"], self.layout["qpos"], self.layout["goal"], st[i].ctypes.data)
            episode[i] += 1
            self.elapsed[i] = 0
        self.launches += 1
        self.refresh(mask, out)

    def reset_uniform(self, mask, rest_record, params, seed, env_offset, episode, out):
        import ctypes

        L = self.sim._L
        L.hostsim_uniform_reset_record.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int,
                                                   ctypes.c_void_p]
        L.hostsim_uniform_reset_record.restype = None
        st, rest = self.state.numpy(), rest_record.numpy()
        for i in range(self.num_envs):
            if mask is not None and not bool(mask[i]):
                continue
            L.hostsim_uniform_reset_record(ctypes.byref(params), int(seed) & 0xFFFFFFFFFFFFFFFF, i + int(env_offset), int(episode[i]), rest.ctypes.data,
                                           self.layout["stride"], st[i].ctypes.data)
            episode[i] += 1
            self.elapsed[i] = 0
        self.launches += 1
        self.refresh(mask, out)

    def reset_maze(self, mask, rest_record, params, goal_xy, reset_xy, seed, env_offset, episode, out):
        import ctypes

        L = self.sim._L
        L.hostsim_maze_reset_record.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint,
                                                ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.hostsim_maze_reset_record.restype = None
        st, rest, g, r = self.state.numpy(), rest_record.numpy(), goal_xy.numpy(), reset_xy.numpy()
        for i in range(self.num_envs):
            if mask is not None and not bool(mask[i]):
                continue
            L.hostsim_maze_reset_record(ctypes.byref(params), g.ctypes.data, r.ctypes.data, int(seed) & 0xFFFFFFFFFFFFFFFF, i + int(env_offset),
                                        int(episode[i]), rest.ctypes.data, self.layout["stride"], self.layout["qpos"], self.layout["goal"],
                                        st[i].ctypes.data)
            episode[i] += 1
            self.elapsed[i] = 0
        self.launches += 1
        self.refresh(mask, out)

    def reset_hand_pose(self, mask, rest_record, params, parallel, seed, env_offset, episode, attempt):
        import ctypes

        L = self.sim._L
        L.hostsim_hand_pose_record.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, ctypes.c_uint,
                                               ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        L.hostsim_hand_pose_record.restype = None
        st, rest, par = self.state.numpy(), rest_record.numpy(), parallel.numpy()
        for i in range(self.num_envs):
            if mask is None or bool(mask[i]):
                L.hostsim_hand_pose_record(ctypes.byref(params), par.ctypes.data, int(seed) & 0xFFFFFFFFFFFFFFFF, i + int(env_offset), int(episode[i]),
                                           int(attempt), rest.ctypes.data, self.layout["stride"], self.layout["qpos"], self.layout["goal"], self.ngoal,
                                           st[i].ctypes.data)
        self.launches += 1

    def reset_hand_goal(self, mask, params, parallel, seed, env_offset, episode, out):
        import ctypes

        L = self.sim._L
        L.hostsim_hand_goal.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p]
        L.hostsim_hand_goal.restype = None
        st, par = self.state.numpy(), parallel.numpy()
        for i in range(self.num_envs):
            if mask is None or bool(mask[i]):
                L.hostsim_hand_goal(ctypes.byref(params), par.ctypes.data, int(seed) & 0xFFFFFFFFFFFFFFFF, i + int(env_offset), int(episode[i]),
                                    self.layout["qpos"], self.layout["goal"], st[i].ctypes.data)
                episode[i] += 1
                self.elapsed[i] = 0
        self.launches += 1
        self.refresh(mask, out)

    def reset_reach(self, mask, rest_record, params, seed, env_offset, episode, out):
        import ctypes

        L = self.sim._L
        L.hostsim_reach_reset_record.argtypes = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int,
                                                 ctypes.c_int, ctypes.c_void_p]
        L.hostsim_reach_reset_record.restype = None
        st, rest = self.state.numpy(), rest_record.numpy()
        for i in range(self.num_envs):
            if mask is None or bool(mask[i]):
                L.hostsim_reach_reset_record(ctypes.byref(params), int(seed) & 0xFFFFFFFFFFFFFFFF, i + int(env_offset), int(episode[i]), rest.ctypes.data,
                                             self.layout["stride"], self.layout["goal"], st[i].ctypes.data)
                episode[i] += 1
                self.elapsed[i] = 0
        self.launches += 1
        self.refresh(mask, out)

    def check_state(self, bad, rest_record, keep):
        import ctypes

        L = self.sim._L
        L.hostsim_check_record.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        st = self.state.numpy()
        rest = rest_record.numpy() if rest_record is not None else None
        for i in range(self.num_envs):
            bad[i] = L.hostsim_check_record(st[i].ctypes.data, self.layout["stride"], rest.ctypes.data if rest is not None else None,
                                            ctypes.byref(keep) if keep is not None else None)
        self.launches += 1

    def compute_reward(self, ag, dg):
        ag = ag.to(torch.float32).reshape(-1, self.ngoal); dg = dg.to(torch.float32).reshape(-1, self.ngoal)
        if self.task.kind == 2:  # same arithmetic as the kernel's hand_goal_distance / hand_reward (fetch_task.cuh)
            t = self.task
            dp = torch.sqrt(((ag[:, :3] - dg[:, :3]) ** 2).sum(-1)) if t.goal_flags & 1 else torch.zeros(ag.shape[0])
            qa = ag[:, 3:]
            if t.goal_flags & 4:   # ignore_z_target_rotation (the pen)
                from gymnasium_robotics_b200 import rotations as R
                ea, eb = R.quat2euler(qa.double().numpy()), R.quat2euler(dg[:, 3:].double().numpy())
                ea[:, 2] = eb[:, 2]
                qa = torch.as_tensor(R.euler2quat(ea), dtype=torch.float32)
            dr = 2 * torch.acos(torch.clamp((qa * dg[:, 3:]).sum(-1), -1, 1)) if t.goal_flags & 2 else torch.zeros(ag.shape[0])
            suc = (dp < t.distance_threshold).to(torch.float32) * (dr < t.rotation_threshold).to(torch.float32)
            return -(10 * dp + dr) if t.reward_dense else suc - 1
        d = torch.sqrt(((ag - dg) ** 2).sum(-1))
        if self.task.kind == 1:
            return torch.exp(-d) if self.task.reward_dense else (d <= self.task.success_radius).to(torch.float32)
        return -d if self.task.reward_dense else -(d > self.task.distance_threshold).to(torch.float32)
