"""Experiments only: wall-clock share of the pipeline stages inside the step kernel (library built with -DB200_STAGE_TIMING)."""
import sys, ctypes; sys.path.insert(0, '.')
import torch
from gymnasium_robotics_b200 import _lib
from gymnasium_robotics_b200.fetch import FetchVectorEnv
from gymnasium_robotics_b200.hand import HandVectorEnv
NAMES = ["kinematics", "com+mass_matrix", "collision", "make_constraint", "smooth_forces", "newton_begin", "newton_check", "build_H",
         "newton_direction", "newton_move", "integrate", "barrier wait", "other (load/observe/store)", "  (newton_move: M*search)", "  (newton_move: J*search + dots)", "  (newton_move: line search)",
         "  (newton_check: J^T f)", "  (newton_move: update)"]
L = _lib.lib()
for which in sys.argv[1:] or ["fetch", "hand"]:
    if which == "fetch":
        n, nact = 4096, 4
        env = FetchVectorEnv("FetchPickAndPlace", num_envs=n, rng_mode="torch", autoreset_mode="same_step")
    else:
        n, nact = 2048, 20
        env = HandVectorEnv("HandManipulateBlockRotateXYZ", num_envs=n, rng_mode="torch", autoreset_mode="same_step", max_episode_steps=None)
    env.reset(seed=0)
    g = torch.Generator(device="cuda").manual_seed(1)
    tape = torch.rand((16, n, nact), generator=g, device="cuda") * 2 - 1
    for k in range(5): env.step(tape[k])
    out = (ctypes.c_ulonglong * 32)()
    L.b200sim_debug_stage_cycles(out, 1)
    for k in range(10): env.step(tape[k])
    cnt = L.b200sim_debug_stage_cycles(out, 1)
    tot = sum(out[:13])
    print(f"== {which}: cycles per env-step per warp {tot / (10 * n):.0f}")
    for k in range(cnt):
        print(f"  {NAMES[k]:28s} {100 * out[k] / tot:5.1f}%")
    env.close()
