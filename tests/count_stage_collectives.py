"""Warp collectives per pipeline stage of one FetchPickAndPlace sub-step (32-lane fiber emulation with -DB200_STAGE_TIMING: the stage
clock is the scheduler-round counter).  A GPU-free view of where the synchronisation points are:
    PYTHONPATH=. python tests/count_stage_collectives.py"""
import ctypes

import numpy as np

import gymnasium_robotics_b200 as pkg
from tests import hostsim
from tests.hostsim_backend import HostSimBackend

NAMES = ["kinematics", "com + M", "collision", "constraint rows", "smooth forces", "newton begin", "newton check", "build H", "newton dir (SPD solve)",
         "newton move", "integrate", "barrier", "other", "(move: mulM)", "(move: rows)", "(move: line search)", "(check: pass F)", "(move: update)"]


class W1(HostSimBackend):
    pass


class WT(HostSimBackend):
    FLAVOR = "warp_timing"


env = pkg.make_vec("FetchPickAndPlace-v4", num_envs=1, backend_factory=W1, rng_mode="numpy")
env.reset(seed=3)
rng = np.random.default_rng(5)
for _ in range(4):
    env.step(rng.uniform(-1, 1, size=(1, 4)).astype(np.float32))
from gymnasium_robotics_b200.fetch import welded_eq_data
b = WT(env.model, welded_eq_data(env.model), env.task, 1, "cpu")
b.state.copy_(env.backend.state)
L = hostsim.lib("warp_timing")
buf = (ctypes.c_longlong * 32)()
L.hostsim_stage_rounds(buf, 1)
import torch
out = b.new_outputs()
steps = 3
for _ in range(steps):
    b.step(torch.as_tensor(rng.uniform(-1, 1, size=(1, 4)).astype(np.float32)), out)
n = L.hostsim_stage_rounds(buf, 1)
tot = sum(buf[k] for k in range(13))
print(f"warp collectives per sub-step (FetchPickAndPlace, {steps} env-steps x 20 sub-steps): {tot / steps / 20:.0f}")
for k in range(n):
    if buf[k]:
        print(f"  {NAMES[k]:28s} {buf[k] / steps / 20:8.1f}  {100 * buf[k] / tot:5.1f} %")
