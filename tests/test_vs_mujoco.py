"""Oracle vs the real MuJoCo -- auto-enables when the `mujoco` wheel AND the reference's MJCF assets are present (neither is in
this image: the oracle is "parity unpinned", see oracle/oracle.c and DESIGN.md section 2).  When it runs it pins the fp64 CPU
restatement to `mujoco.mj_step` on the reference's own models from identical (qpos, qvel, ctrl, mocap) -- which is what turns
the GPU-vs-oracle parity tests into GPU-vs-reference statements."""
import os

import numpy as np
import pytest

mujoco = pytest.importorskip("mujoco", reason="the mujoco wheel is not installed (pyproject.toml:27 of the reference)")
ASSETS = os.environ.get("B200SIM_REFERENCE_ASSETS", "/root/reference/gymnasium_robotics/envs/assets")
pytestmark = pytest.mark.skipif(not os.path.isdir(ASSETS), reason="reference MJCF assets not available")

# model, sub-steps per comparison, max |dqpos| tolerated (mesh geoms are box proxies in the oracle: contact-free phases only)
CASES = [("fetch/reach.xml", 20, 1e-6), ("hand/reach.xml", 20, 1e-6), ("adroit_hand/adroit_relocate.xml", 5, 1e-6)]


@pytest.mark.parametrize("rel,nstep,tol", CASES)
def test_oracle_matches_mj_step_in_contact_free_motion(rel, nstep, tol):
    from gymnasium_robotics_b200.mjcf import compile_mjcf
    from oracle.oracle_sim import OracleSim

    path = os.path.join(ASSETS, rel)
    mm = mujoco.MjModel.from_xml_path(path)
    md = mujoco.MjData(mm)
    orc = OracleSim(compile_mjcf(path))
    if orc.model.nq != mm.nq:
        pytest.skip("the compiled model fuses / drops bodies for this file; compare through the env-level tests instead")
    rng = np.random.default_rng(0)
    md.ctrl[:] = rng.uniform(-0.1, 0.1, mm.nu)
    orc.ctrl[:] = md.ctrl
    orc.qpos[:], orc.qvel[:] = md.qpos, md.qvel
    for _ in range(5):
        mujoco.mj_step(mm, md, nstep=nstep)
        orc.step(nstep)
        assert np.abs(orc.qpos - md.qpos).max() < tol
        assert np.abs(orc.qvel - md.qvel).max() < 100 * tol
