"""Diagnostic driver for the FrankaKitchen CUDA build: where does a non-finite value first appear?
Usage: python tests/kitchen_diag.py [n_envs] [mode]   (mode: env | substeps | fixture4)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gymnasium_robotics_b200 as grb

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
mode = sys.argv[2] if len(sys.argv) > 2 else "substeps"
print("lib:", os.environ.get("B200SIM_LIB", "default"), "n =", n, "mode =", mode, flush=True)


def report(tag, t):
    t = t.detach().float().cpu()
    bad = ~torch.isfinite(t)
    print(f"  {tag}: shape {tuple(t.shape)} nonfinite {int(bad.sum())} max|x| {float(t[~bad].abs().max()) if (~bad).any() else float('nan'):.4g}"
          + (f" first bad idx {bad.nonzero()[:6].tolist()}" if bad.any() else ""), flush=True)


os.environ["B200SIM_EXPERIMENTAL_KITCHEN"] = "1"
env = grb.make_vec("FrankaKitchen-v1", num_envs=n, rng_mode="numpy")
obs, info = env.reset(seed=21)
report("reset obs", obs["observation"])
be = env.backend
lay = be.layout
print("  launch config", be.L.b200sim_launch_count(be.h), lay, flush=True)
rng = np.random.default_rng(2)
a = torch.as_tensor(rng.uniform(-1, 1, size=(n, 9)), dtype=torch.float32, device="cuda")
if mode == "env":
    for k in range(3):
        obs, rew, term, trunc, info = env.step(a)
        report(f"step {k} obs", obs["observation"])
        report(f"step {k} state", be.state)
elif mode == "substeps":
    vel = torch.clamp(torch.clamp(a, -1.0, 1.0) * 2.0, env._vel_lo, env._vel_hi)
    ctrl = torch.clamp(env._last_robot_qpos + vel * env.dt, env._pos_lo, env._pos_hi).contiguous()
    be.state[:, lay["ctrl"]:lay["ctrl"] + 9] = ctrl
    out = be.new_outputs()
    for s in range(40):
        be.raw_step(1, out)
        torch.cuda.synchronize()
        st = be.state
        bad = ~torch.isfinite(st)
        qv = st[:, lay["qvel"]:lay["warm"]]
        wa = st[:, lay["warm"]:lay["ctrl"]]
        fm = lambda t: float(t[torch.isfinite(t)].abs().max()) if torch.isfinite(t).any() else float("nan")
        print(f"  sub {s}: nonfinite {int(bad.sum())} max|qvel| {fm(qv):.4g} max|qacc| {fm(wa):.4g}", flush=True)
        if bad.any():
            print("   bad cols per env:", [(i, bad[i].nonzero().flatten()[:8].tolist()) for i in range(n) if bad[i].any()][:4])
            break
elif mode == "fixture4":
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kitchen_quick.npz"))
    be.state.copy_(torch.as_tensor(g["state0"][:n]))
    out = be.new_outputs()
    be.refresh(None, out)
    report("fixture refresh obs", out["obs"])
    for k in range(g["ctrls"].shape[0]):
        be.step(torch.as_tensor(g["ctrls"][k][:n]).cuda().contiguous(), out)
        report(f"fixture step {k}", out["obs"])
        print("   err vs emulation", float((out["obs"].cpu() - torch.as_tensor(g["obs"][k + 1][:n])).abs().max()))
torch.cuda.synchronize()
print("diag done", flush=True)
env.close()
