"""FrankaKitchen-v1 (BASELINE config 5b), oracle-first: the model compiles (mesh-derived link inertias, joint
equalities, condim-6 pairs), the fp64 oracle steps it, and the env logic of envs/franka_kitchen/{franka_env,kitchen_env}.py is
restated on top.  The CUDA path refuses the model loudly (no kernel support for joint equalities / condim 6 / 3 708 pairs yet)."""
import numpy as np
import pytest

from gymnasium_robotics_b200.models import load_model
from oracle.kitchen_env import BONUS_THRESH, INIT_QPOS, OBS_ELEMENT_GOALS, OBS_ELEMENT_INDICES, OracleKitchenEnv


@pytest.fixture(scope="module")
def model():
    return load_model("franka_kitchen")


def test_model_facts(model):
    m = model
    assert (m.nq, m.nv, m.nu, m.neq) == (30, 29, 9, 5) and len(m.pair_geom1) > 3000
    # link masses come from `mass=` on the collision meshes (franka_assets/chain.xml:8-42): the compiler integrates the mesh
    bm = {n: m.body_mass[i] for n, i in m.names["body_map"].items()}
    assert bm["panda0_link1"] == pytest.approx(2.7063, rel=1e-6) and bm["panda0_link5"] == pytest.approx(3.00049, rel=1e-6)
    # the five joint couplings of oven_asset.xml:40-46 (knob = 174 x burner, switch = 14 x light)
    eq = np.asarray(m.eq_data).reshape(-1, 11)
    assert np.allclose(eq[:4, 1], 174) and eq[4, 1] == pytest.approx(14)
    assert set(m.pair_condim.tolist()) >= {3, 6}


def test_mesh_inertia_of_a_cube():
    from gymnasium_robotics_b200.mjcf import mesh_volume_inertia

    c = np.array([[x, y, z] for x in (0, 2) for y in (0, 1) for z in (0, 3)], dtype=float)
    quads = [(0, 1, 3, 2), (4, 6, 7, 5), (0, 4, 5, 1), (2, 3, 7, 6), (0, 2, 6, 4), (1, 5, 7, 3)]
    tris = np.array([[c[q[0]], c[q[i]], c[q[i + 1]]] for q in quads for i in (1, 2)])
    vol, com, I = mesh_volume_inertia(tris)
    assert vol == pytest.approx(6.0) and np.allclose(com, [1, 0.5, 1.5])
    assert np.allclose(I, np.diag([6 * (1 + 9) / 12, 6 * (4 + 9) / 12, 6 * (4 + 1) / 12]), atol=1e-12)


def test_env_logic_matches_the_reference_contract(model):
    env = OracleKitchenEnv(model)
    obs, info = env.reset(seed=3)
    assert obs["observation"].shape == (59,) and set(obs["achieved_goal"]) == set(OBS_ELEMENT_GOALS)
    assert sorted(info["tasks_to_complete"]) == sorted(OBS_ELEMENT_GOALS) and info["episode_task_completions"] == []
    # observation = noisy robot qpos / qvel (ratio 0.01 x amplitude 0.1) | noisy object qpos / qvel
    assert np.abs(obs["observation"][:9] - INIT_QPOS[:9]).max() <= 0.01 * 0.1 + 1e-12
    assert np.abs(obs["observation"][18:18 + 21] - INIT_QPOS[9:]).max() <= 0.0005 * 0.1 + 1e-12
    # same seed => same noise stream
    obs2, _ = OracleKitchenEnv(model).reset(seed=3)
    assert np.array_equal(obs["observation"], obs2["observation"])
    # zero action: position targets = last (noisy) robot qpos; the arm holds its pose, the kitchen stays put
    o, r, term, trunc, info = env.step(np.zeros(9))
    assert r == 0.0 and not term and not trunc and env.sim.overflow == 0
    assert np.abs(env.sim.qpos[:7] - INIT_QPOS[:7]).max() < 0.05 and np.abs(env.sim.qpos[9:23] - INIT_QPOS[9:23]).max() < 2e-2
    # joint equalities hold: knob angle = 174 x burner slide
    q = env.sim.qpos
    assert abs(q[9] - 174 * q[10]) < 5e-3 and abs(q[17] - 14 * q[18]) < 5e-3


def test_task_completion_bookkeeping(model):
    env = OracleKitchenEnv(model, tasks_to_complete=["microwave", "slide cabinet"])
    env.reset(seed=0)
    s = env.sim
    # put the microwave door at its goal: reward 1 this step, task removed, not yet terminated
    s.qpos[OBS_ELEMENT_INDICES["microwave"]] = OBS_ELEMENT_GOALS["microwave"]
    s.qvel[:] = 0
    o, r, term, _, info = env.step(np.zeros(9))
    assert r == 1.0 and info["step_task_completions"] == ["microwave"] and info["tasks_to_complete"] == ["slide cabinet"] and not term
    o, r, term, _, info = env.step(np.zeros(9))
    assert r == 0.0 and info["episode_task_completions"] == ["microwave"]
    s.qpos[OBS_ELEMENT_INDICES["slide cabinet"]] = OBS_ELEMENT_GOALS["slide cabinet"]
    o, r, term, _, info = env.step(np.zeros(9))
    assert r == 1.0 and term and sorted(info["episode_task_completions"]) == ["microwave", "slide cabinet"]
    d = np.linalg.norm(o["achieved_goal"]["slide cabinet"] - OBS_ELEMENT_GOALS["slide cabinet"])
    assert d < BONUS_THRESH
    with pytest.raises(ValueError):
        OracleKitchenEnv(model, tasks_to_complete=["dishwasher"])


def test_cuda_path_never_falls_back_for_the_kitchen():
    """FrankaKitchen-v1 is a regular id of the CUDA path (GPU-validated in round 2); without a CUDA device it fails loudly
    instead of falling back to the oracle, and the reference's registry holds no other kitchen id."""
    import gymnasium_robotics_b200 as pkg
    import torch

    assert "FrankaKitchen-v1" in pkg.ENV_IDS
    with pytest.raises(KeyError):
        pkg.make_vec("FrankaKitchen-v2", num_envs=2)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            pkg.make_vec("FrankaKitchen-v1", num_envs=2)
