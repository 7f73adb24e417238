import sys, time; sys.path.insert(0,'.')
import torch
from gymnasium_robotics_b200.fetch import FetchVectorEnv
for n in (4096,):
    env = FetchVectorEnv("FetchPickAndPlace", num_envs=n, rng_mode="torch", autoreset_mode="same_step")
    env.reset(seed=0)
    a = torch.rand((n,4), device="cuda")*2-1
    for _ in range(5): env.step(a)
    torch.cuda.synchronize(); t=time.time()
    K=50
    for _ in range(K): env.step(a)
    torch.cuda.synchronize(); dt=time.time()-t
    print(f"N={n}: {dt/K*1e3:.2f} ms/step, {n*K/dt:.0f} env-steps/s")
    import ctypes
    sm=ctypes.c_int(); epb=ctypes.c_int(); bl=ctypes.c_int()
    env.backend.L.b200sim_launch_config(env.backend.h, ctypes.byref(sm), ctypes.byref(epb), ctypes.byref(bl)); print("smem", sm.value, "envs/block", epb.value, "blocks", bl.value)
