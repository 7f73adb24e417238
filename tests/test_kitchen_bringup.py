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


ROLLING = """
<mujoco><option timestep="0.002"/>
<worldbody>
  <geom type="plane" size="2 2 0.1" condim="CD" friction="1 0.02 0.02"/>
  <body name="ball" pos="0 0 0.05"><freejoint/><geom type="sphere" size="0.05" mass="0.5" condim="CD" friction="1 0.02 0.02"/></body>
</worldbody></mujoco>"""


def test_condim6_rolling_friction_tracks_the_oracle(mjcf_file):
    """Step (ii): two rolling base rows per contact (10 pyramid edges).  A ball rolling on a plane slows down under rolling
    friction; emulation (fp32, -DB200_KITCHEN) against the oracle (fp64), and against the condim-3 ball that keeps rolling."""
    m6 = compile_mjcf(mjcf_file(ROLLING.replace("CD", "6")))
    assert set(m6.pair_condim.tolist()) == {6}
    orc, hs = OracleSim(m6), HostSim(m6, ref=(0.0, 0.0, 0.05), flavor="kitchen")
    v0 = np.array([0.5, 0, 0, 0, 0.5 / 0.05, 0])          # rolling without slipping along +x (angular velocity in the body frame)
    hs.qpos[:] = orc.qpos                                  # the emulation scratch starts zeroed: load the model's qpos0
    for s in (orc, hs):
        s.qvel[:] = v0
    hs.qacc[:] = 0
    for k in range(5):
        orc.step(50)
        hs.step(50)
        assert np.abs(hs.qpos[:3] - orc.qpos[:3]).max() < 5e-5 and np.abs(hs.qvel - orc.qvel).max() < 5e-3
    o3 = OracleSim(compile_mjcf(mjcf_file(ROLLING.replace("CD", "3"))))
    o3.qvel[:] = v0
    o3.step(250)
    assert orc.qvel[0] < 0.8 * o3.qvel[0] and o3.qvel[0] > 0.45      # rolling friction decelerates; without it the ball keeps its speed
    assert orc.nefc == 10
