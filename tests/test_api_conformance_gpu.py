"""The API conformance checks of tests/test_api_conformance.py on the product path (CUDA backend through the C-ABI): spaces, output
shapes / dtypes, info keys, the GoalEnv contract `reward == compute_reward(achieved_goal, desired_goal, info)` (core.py:45-62) -- bit-exact
here, dense rewards included, because `compute_reward` runs the same device routine as the step kernel -- and same-seed determinism."""
import numpy as np
import pytest
import torch

import gymnasium_robotics_b200 as pkg
from tests.test_api_conformance import CASES, _leaves, _same

pytestmark = pytest.mark.gpu

# more dense-reward ids on the product path (the GoalEnv contract is the point here)
CASES = CASES + [("AntMaze_UMazeDense-v5", True, 8, 105), ("PointMaze_MediumDense-v3", True, 2, 4), ("FetchPickAndPlaceDense-v4", True, 4, 25),
                 ("HandManipulateBlockRotateXYZDense-v1", True, 20, 61), ("HandManipulateEggRotateDense-v1", True, 20, 61)]


@pytest.mark.parametrize("env_id,goal_env,nact,nobs", CASES, ids=[c[0] for c in CASES])
def test_api_conformance_on_gpu(env_id, goal_env, nact, nobs):
    n = 8
    env = pkg.make_vec(env_id, num_envs=n, rng_mode="numpy")
    assert env.single_action_space.shape == (nact,) and env.action_space.shape == (n, nact)
    obs, info = env.reset(seed=11)
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(3):
        a = torch.rand((n, nact), generator=g, device="cuda") * 2 - 1
        o, r, te, tr, inf = env.step(a)
        ob = o["observation"] if goal_env else o
        assert ob.is_cuda and tuple(ob.shape) == (n, nobs) and ob.dtype == torch.float32
        for k, v in _leaves(o):
            assert bool(torch.isfinite(v).all()), k
        assert tuple(r.shape) == (n,) and r.dtype == torch.float32 and te.dtype == torch.bool and tr.dtype == torch.bool
        want = ("is_success",) if env_id.startswith(("Fetch", "Hand")) else (("success",) if not env_id.startswith("Franka") else
                ("tasks_to_complete", "step_task_completions", "episode_task_completions"))
        assert all(k in inf for k in want), sorted(inf)
        if goal_env and not env_id.startswith("Franka"):
            rr = torch.as_tensor(env.compute_reward(o["achieved_goal"], o["desired_goal"], inf), dtype=torch.float32).reshape(n)
            if r.device.type == "cuda":
                assert torch.equal(rr, r), (env_id, float((rr - r).abs().max()))
            else:   # (tests/dryrun_gpu_tests.py: the emulation backend's reward routine is a numpy twin)
                assert torch.allclose(rr, r, rtol=2e-6, atol=2e-6)
    with pytest.raises((ValueError, AssertionError, RuntimeError)):
        env.step(torch.zeros((n, nact + 1), device="cuda"))
    env.close()


@pytest.mark.parametrize("env_id", ["FetchPickAndPlace-v4", "HandManipulateEggFull-v1", "AdroitHandRelocate-v2", "AntMaze_Medium-v5", "FrankaKitchen-v1"])
def test_same_seed_rollouts_are_bit_identical_on_gpu(env_id):
    n = 6
    e1, e2 = (pkg.make_vec(env_id, num_envs=n, rng_mode="numpy") for _ in range(2))
    nact = e1.single_action_space.shape[0]
    o1, _ = e1.reset(seed=5)
    o2, _ = e2.reset(seed=5)
    _same(o1, o2)
    rng = np.random.default_rng(9)
    for t in range(5):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, nact)).astype(np.float32)).cuda()
        s1, s2 = e1.step(a.clone()), e2.step(a.clone())
        _same(s1[0], s2[0])
        assert torch.equal(s1[1], s2[1]) and torch.equal(s1[2], s2[2]) and torch.equal(s1[3], s2[3])
    e1.close()
    e2.close()
