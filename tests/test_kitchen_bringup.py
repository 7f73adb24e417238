"""Bring-up of the Franka-Kitchen kernel features, step (i): joint equalities as two-sided dof rows.
Compiled only into the emulation build `-DB200_KITCHEN`; the product library does not carry it yet."""
import numpy as np

from gymnasium_robotics_b200.mjcf import compile_mjcf
from oracle.oracle_sim import OracleSim
from tests.hostsim import HostSim

COUPLED = """
<mujoco><option timestep="0.002"/>
<worldbody>
  <body name="a" pos="0 0 1"><joint name="knob" type="hinge" axis="0 1 0" damping="0.05" armature="0.001"/>
    <geom type="capsule" size="0.02 0.1" pos="0 0 -0.1" mass="0.3"/></body>
  <body name="b" pos="0.5 0 1"><joint name="burner" type="slide" axis="0 0 1" damping="2" armature="0.001"/>
    <geom type="sphere" size="0.03" mass="0.01"/></body>
</worldbody>
<equality><joint joint1="knob" joint2="burner" polycoef="0 20 0 0 0"/></equality>
</mujoco>"""


def test_joint_equality_rows_track_the_oracle(mjcf_file):
    model = compile_mjcf(mjcf_file(COUPLED))
    assert model.neq == 1
    orc, hs = OracleSim(model), HostSim(model, ref=(0.0, 0.0, 1.0), flavor="kitchen")
    q0 = np.array([0.3, 0.0])
    orc.qpos[:] = q0
    hs.qpos[:] = q0
    hs.qvel[:] = 0
    hs.qacc[:] = 0
    for k in range(4):
        orc.step(25)
        hs.step(25)
        assert np.abs(hs.qpos - orc.qpos).max() < 2e-5 and np.abs(hs.qvel - orc.qvel).max() < 2e-3
    # the coupling pulls the residual knob - 20 * burner towards zero (soft constraint, default solref)
    assert abs(orc.qpos[0] - 20 * orc.qpos[1]) < 0.05 and abs(hs.qpos[0] - 20 * hs.qpos[1]) < 0.05
