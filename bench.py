#!/usr/bin/env python
"""bench.py -- env-steps/s of the b200sim CUDA path on the BASELINE.json headline workload.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference [...]                       the CPU arm: this repo's fp64 restatement of the
                                                               reference's mj_step path (the reference itself cannot
                                                               be imported: `mujoco`/`gymnasium` are absent), on all
                                                               host cores, on a bounded sample of the same workload

A "step" is one `step()` of every env of the batch (FetchPickAndPlace-v4, 4096 envs per GPU, 20 physics sub-steps per
env-step, same-step autoreset so every counted env-step contains a physics step).  One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TASK, ENV_ID, ENVS_PER_GPU = "FetchPickAndPlace", "FetchPickAndPlace-v4", 4096
# algorithmic HBM bytes per env-step: state read+write, action, obs/goals/reward/flags written
B_ALG = 2 * 4 * (22 + 2 * 21 + 2 + 7 + 3 + 1) + 4 * 4 + 4 * (25 + 2 * 3) + 10
# other workloads (BASELINE.json configs): name -> (env id, action dim, sub-steps, algorithmic bytes per env-step, default envs/GPU)
WORKLOADS = {
    "fetch_pick_and_place": ("FetchPickAndPlace-v4", 4, 20, B_ALG, 4096),
    # config 3 (plain 61-dim observation; nq = 31, nv = 30 without the visual-only target body)
    "hand_block": ("HandManipulateBlockRotateXYZ-v1", 20, 20, 2 * 4 * (31 + 60 + 20 + 0 + 7 + 1) + 80 + 4 * (61 + 14) + 10, 2048),
    # config 3 as named: 24 DoF + 92 touch sensors (153-dim observation)
    "hand_block_touch": ("HandManipulateBlockRotateXYZ_ContinuousTouchSensors-v1", 20, 20,
                         2 * 4 * (31 + 60 + 20 + 0 + 7 + 1) + 80 + 4 * (153 + 14) + 10, 2048),
    # further env families on the same kernels (general convex collider: cylinder puck, ellipsoid egg)
    "fetch_slide": ("FetchSlide-v4", 4, 20, 2 * 4 * (22 + 2 * 21 + 0 + 7 + 3 + 1) + 4 * 4 + 4 * (25 + 2 * 3) + 10, 4096),
    "hand_egg": ("HandManipulateEggRotate-v1", 20, 20, 2 * 4 * (31 + 60 + 20 + 0 + 7 + 1) + 80 + 4 * (61 + 14) + 10, 2048),
    # config 5a: AdroitHandHammer (33 dofs, wide kernel build); registered id is -v2 (SURVEY.md 8, config-name caveats)
    "adroit_hammer": ("AdroitHandHammer-v2", 26, 5, 2 * 4 * (33 + 66 + 26 + 7 + 3 + 1) + 104 + 4 * (46 + 6) + 10, 2048),
    "adroit_relocate": ("AdroitHandRelocate-v2", 30, 5, 2 * 4 * (36 + 72 + 30 + 7 + 3 + 1) + 120 + 4 * (39 + 6) + 10, 2048),
    "adroit_pen": ("AdroitHandPen-v2", 24, 5, 2 * 4 * (30 + 60 + 24 + 7 + 3 + 1) + 96 + 4 * (45 + 6) + 10, 2048),
    "adroit_door": ("AdroitHandDoor-v2", 28, 5, 2 * 4 * (30 + 60 + 28 + 7 + 3 + 1) + 112 + 4 * (39 + 6) + 10, 2048),
    # config 5b: FrankaKitchen-v1 (bring-up build, opt-in: csrc/b200sim_kitchen.cu); 40 sub-steps per env-step
    "franka_kitchen": ("FrankaKitchen-v1", 9, 40, 2 * 4 * (30 + 2 * 29 + 9) + 36 + 4 * (59 + 2 * 30 + 2) + 4, 2048),
    "antmaze_large": ("AntMaze_Large-v5", 8, 5, 2 * 4 * (15 + 28 + 0 + 0 + 2 + 1) + 32 + 124 + 10, 1024),  # config 4: 8192 envs over 8 GPUs
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: oracle on all host cores
_W = {}


def _worker_init(nenv, seed0):
    from oracle.fetch_env import OracleFetchEnv
    from gymnasium_robotics_b200.models import load_model

    model = load_model("fetch_pick_and_place")
    _W["envs"] = [OracleFetchEnv(TASK, model=model) for _ in range(nenv)]
    for i, e in enumerate(_W["envs"]):
        e.reset(seed=seed0 + i)
    _W["t"] = [0] * nenv


def _worker_step(actions):
    import numpy as np

    for i, e in enumerate(_W["envs"]):
        e.step(np.asarray(actions[i], dtype=np.float64))
        _W["t"][i] += 1
        if _W["t"][i] >= 50:  # TimeLimit + autoreset, as in the GPU arm
            e.reset()
            _W["t"][i] = 0
    return len(_W["envs"])


def run_reference(args, quiet=False):
    import multiprocessing as mp
    import numpy as np

    cores = max(1, os.cpu_count() or 1)
    per = max(1, max(args.sample_envs, 16 * cores) // cores)  # >= 16 envs per worker so IPC does not dominate
    nenv = per * cores
    ctx = mp.get_context("fork")
    pools = [ctx.Pool(1, initializer=_worker_init, initargs=(per, 1000 * w)) for w in range(cores)]
    rng = np.random.default_rng(1234)
    tape = rng.uniform(-1, 1, (64, nenv, 4)).astype(np.float32)

    def one_step(k):
        a = tape[k % 64]
        res = [p.apply_async(_worker_step, (a[w * per:(w + 1) * per],)) for w, p in enumerate(pools)]
        return sum(r.get() for r in res)

    for k in range(args.warmup):
        one_step(k)
    t0 = time.perf_counter()
    for k in range(args.steps):
        one_step(k)
    dt = time.perf_counter() - t0
    for p in pools:
        p.close()
    value = nenv * args.steps / dt
    sample = f"{nenv} envs x {args.steps} env-steps (one env per worker slot, {cores} worker processes, TimeLimit 50 + reset)"
    line = {"impl": "reference", "metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{ENV_ID}, CPU restatement of the reference mj_step path (NOT MuJoCo: dependency absent), "
                                   f"bounded sample of {nenv} envs per step", "n_substeps": 20},
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if not quiet:
        print(json.dumps(line))
    return line


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from gymnasium_robotics_b200.fetch import FetchVectorEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    mixed = None
    if args.workload == "mixed_hammer_kitchen":
        # BASELINE config 5: heterogeneous batch, whole ranks per model (sharding.mixed_batch_assignment), 1 024 envs per GPU; the
        # line reports the aggregate over both models, the roofline object is rank 0's model (the Hammer)
        from gymnasium_robotics_b200.sharding import mixed_batch_assignment

        if world < 2:
            raise SystemExit("--workload mixed_hammer_kitchen needs at least 2 ranks (torchrun --nproc-per-node 2|4|8)")
        mixed = mixed_batch_assignment(["adroit_hammer", "franka_kitchen"], world)
        args.workload = mixed[rank]
        args.envs_per_gpu = args.envs_per_gpu or 1024
    env_id, nact, nsub, b_alg, default_n = WORKLOADS[args.workload]
    n = args.envs_per_gpu or default_n
    if args.workload == "fetch_pick_and_place":
        # --rng-mode device: resets drawn inside the library (b200sim_reset) with seeds invariant to the world size
        kw = dict(env_offset=rank * n) if args.rng_mode == "device" else {}
        env = FetchVectorEnv(TASK, num_envs=n, device=dev, rng_mode=args.rng_mode, autoreset_mode="same_step", **kw)
    else:
        import gymnasium_robotics_b200 as grb

        extra = {"experimental": True} if args.workload == "franka_kitchen" else {}
        env = grb.make_vec(env_id, num_envs=n, device=dev, rng_mode="torch", autoreset_mode="same_step", **extra)
    env.reset(seed=0 if args.rng_mode == "device" else 1000 * rank)  # seeds seed0 + global env index would need numpy streams; device RNG is per rank
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    tape = torch.rand((64, n, nact), generator=g, device=dev) * 2 - 1  # pre-generated action tape (RNG outside the timed region)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > L2 (126 MB)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident arm
    for k in range(args.warmup):
        env.step(tape[k % 64])
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = env.backend.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    out = env.backend.new_outputs()
    nvtx = torch.cuda.nvtx if args.nvtx else None   # --nvtx: ranges for ncu --nvtx / nsys filtering (tracing)
    for k in range(args.steps):
        flush.fill_(float(k))  # evict L2 between timed iterations (outside the timed interval)
        ev[k][0].record()
        if nvtx:
            nvtx.range_push(f"env.step {k}")
        env.step(tape[k % 64])
        if nvtx:
            nvtx.range_pop()
        ev[k][1].record()
    barrier()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = env.backend.launches - launches0
    # dominant kernel alone (the step kernel), same stream, CUDA events around the launch only
    for k in range(args.steps):
        flush.fill_(float(k))
        kev[k][0].record()
        env.backend.step(tape[k % 64], out)
        kev[k][1].record()
    barrier()
    kms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- end to end through the public API with HOST buffers: pinned actions H2D, results D2H, every step
    host_tape = [tape[k].cpu().pin_memory() for k in range(8)]
    nobs = env.task.nobs
    ngoal = env.backend.ngoal
    host_out = {"observation": torch.empty((n, nobs), dtype=torch.float32).pin_memory(),
                "achieved_goal": torch.empty((n, ngoal), dtype=torch.float32).pin_memory(),
                "desired_goal": torch.empty((n, ngoal), dtype=torch.float32).pin_memory(),
                "reward": torch.empty(n, dtype=torch.float32).pin_memory(),
                "truncated": torch.empty(n, dtype=torch.bool).pin_memory(), "terminated": torch.empty(n, dtype=torch.bool).pin_memory(),
                "is_success": torch.empty(n, dtype=torch.float32).pin_memory()}
    if args.workload.startswith("adroit") or args.workload == "franka_kitchen":   # flat observation / per-task goal dicts
        del host_out["achieved_goal"], host_out["desired_goal"]
    h2d = n * nact * 4
    d2h = sum(v.numel() * v.element_size() for v in host_out.values())

    def e2e_step(k):
        o, r, te, tr, info = env.step(host_tape[k % 8])  # FetchVectorEnv.step copies the pinned host actions to the device
        if isinstance(o, dict):
            host_out["observation"].copy_(o["observation"], non_blocking=True)
            if "achieved_goal" in host_out:
                host_out["achieved_goal"].copy_(o["achieved_goal"], non_blocking=True)
                host_out["desired_goal"].copy_(o["desired_goal"], non_blocking=True)
        else:   # flat observation (Adroit)
            host_out["observation"].copy_(o, non_blocking=True)
        host_out["reward"].copy_(r, non_blocking=True)
        host_out["terminated"].copy_(te, non_blocking=True)
        host_out["truncated"].copy_(tr, non_blocking=True)
        suc = info["is_success"] if "is_success" in info else (info["success"] if "success" in info else te)
        host_out["is_success"].copy_(suc.to(torch.float32), non_blocking=True)
        torch.cuda.synchronize(dev)

    for k in range(args.warmup):
        e2e_step(k)
    barrier()
    e2e_s = 0.0
    for k in range(args.steps):
        flush.fill_(float(k))
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_step(k)
        e2e_s += time.perf_counter() - t0
    barrier()
    if os.environ.get("B200SIM_BENCH_DEBUG"):
        print(f"[rank {rank}] ms/step {ms / args.steps:.3f} kernel {kms:.3f} e2e {e2e_s * 1e3 / args.steps:.3f}", file=sys.stderr)
    t = torch.tensor([ms, e2e_s * 1e3, kms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kms = [float(x) for x in t.tolist()]
    total_envs = n * world
    value = total_envs * args.steps / (ms / 1e3)
    e2e_value = total_envs * args.steps / (e2e_ms / 1e3)
    if rank == 0:
        peak, how = measured_peaks()
        achieved = b_alg * n / (kms / 1e3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and args.workload == "fetch_pick_and_place":
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        cpu = None
        if world == 1 and not args.no_cpu_baseline and args.workload == "fetch_pick_and_place":
            cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "20", "--warmup", "2"]
            try:
                outp = subprocess.run(cmd, capture_output=True, text=True, timeout=600).stdout.strip().splitlines()
                cpu = json.loads(outp[-1])["cpu_baseline"]
            except Exception as e:  # noqa: BLE001
                cpu = {"value": None, "unit": "env-steps/s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
        line = {"metric": "env-steps/s", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": ("AdroitHandHammer-v2 + FrankaKitchen-v1 mixed batch, whole ranks per model; rank 0: " if mixed else "") +
                                       f"{env_id}, {n} envs/GPU, {nsub} sub-steps/env-step, random actions U(-1,1), TimeLimit, "
                                       "same-step autoreset", "envs_per_gpu": n, "l2": "flushed between timed iterations (256 MB fill)",
                           "parallelism": f"env-sharded x{world}, no data-path collective" + (f"; models per rank: {mixed}" if mixed else ""),
                           "reset_rng": "in-kernel Philox (b200sim_reset)" if args.rng_mode == "device" else "torch device generator"},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "peak_source": how, "algorithmic_bytes_per_env_step": b_alg,
                             "kernel_ms": kms,
                             "note": "path is FP32-issue/latency bound (SURVEY.md 0.4, 8d); HBM fraction is reported because the metric asks for it"},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line))
    env.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=None)
    ap.add_argument("--workload", default="fetch_pick_and_place", choices=sorted(WORKLOADS) + ["mixed_hammer_kitchen"])
    ap.add_argument("--sample-envs", type=int, default=256, help="envs per step of the CPU arm's bounded sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nvtx", action="store_true", help="NVTX range around every timed env.step (profiling runs only)")
    ap.add_argument("--rng-mode", default="torch", choices=["torch", "device"],
                    help="reset draws of the Fetch workload: torch's device generator (default) or in-kernel (b200sim_reset)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) == 0:
            run_reference(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
