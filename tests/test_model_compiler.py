"""MJCF -> constant tables: sizes, MuJoCo compile rules, blob round trip, committed blobs in sync with the sources."""
import os

import numpy as np
import pytest

from gymnasium_robotics_b200.mjcf import Model, compile_mjcf
from gymnasium_robotics_b200.models import MODEL_DIR, MODEL_OVERRIDES, MODEL_SOURCES, REFERENCE_ASSETS, load_model


def test_committed_blobs_load_and_have_expected_sizes():
    m = load_model("fetch_pick_and_place")
    assert (m.nq, m.nv, m.nu, m.nmocap, m.neq) == (22, 21, 2, 1, 1)  # SURVEY.md section 8 sizes table
    r = load_model("fetch_reach")
    assert (r.nq, r.nv, r.nu) == (15, 15, 0)
    robot = [n for n in m.names["joint"] if n.startswith("robot")]
    assert len(robot) == 15 and robot[-2:] == ["robot0:r_gripper_finger_joint", "robot0:l_gripper_finger_joint"]


def test_blob_round_trip_is_lossless():
    m = load_model("fetch_pick_and_place")
    m2 = Model.from_blob(m.to_blob())
    for f in Model.INT_FIELDS + Model.FLT_FIELDS:
        assert np.array_equal(np.asarray(getattr(m, f)), np.asarray(getattr(m2, f))), f
    assert m.names == m2.names


@pytest.mark.needs_reference
def test_committed_blobs_match_a_fresh_compile():
    for name, rel in MODEL_SOURCES.items():
        fresh = compile_mjcf(os.path.join(REFERENCE_ASSETS, rel), overrides=MODEL_OVERRIDES.get(name)).to_blob()
        assert open(os.path.join(MODEL_DIR, name + ".b200m"), "rb").read() == fresh, name


@pytest.mark.needs_reference
def test_fused_runtime_model_keeps_the_dynamics():
    """Fusing jointless bodies (MuJoCo `fusestatic`) must not change the mass matrix; collision filters follow MuJoCo."""
    from oracle.oracle_sim import OracleSim

    m = compile_mjcf(os.path.join(REFERENCE_ASSETS, "fetch/pick_and_place.xml"))
    s = OracleSim(m)
    s.forward()
    assert np.abs(s.M - m._full_arrays["M0"]).max() < 1e-10  # oracle CRB on the fused tree vs dense sum on the MJCF tree
    assert m.nbody == 16 and len(m._full.bodies) == 33  # 33 MJCF bodies (world included) fuse into 16
    geoms = m.names["geom"]
    pairs = {(geoms[a], geoms[b]) for a, b in zip(m.pair_geom1, m.pair_geom2)}
    assert ("robot0:r_gripper_finger_link", "robot0:l_gripper_finger_link") not in pairs  # <exclude>
    assert not any("robot0:gripper_link" in p and "finger" in p[0] + p[1] for p in pairs)  # parent-child filter
    assert ("object0", "robot0:r_gripper_finger_link") in pairs or ("robot0:r_gripper_finger_link", "object0") in pairs


def test_defaults_childclass_euler_fromto(mjcf_file):
    xml = """
    <mujoco><compiler angle="degree"/>
      <default><joint damping="3"/><default class="a"><geom friction="0.7 0.1 0.1" condim="4"/><joint armature="2"/></default></default>
      <worldbody>
        <body name="b1" pos="0 0 1" euler="0 0 90" childclass="a">
          <joint name="j1" type="hinge" axis="0 1 0" range="-90 90" limited="true"/>
          <geom type="capsule" fromto="0 0 0 0 0 0.4" size="0.05"/>
          <body name="b2" pos="0 0 0.4"><joint name="j2" type="slide" axis="1 0 0" class="main"/><geom type="sphere" size="0.1" class="main"/></body>
        </body>
      </worldbody>
    </mujoco>"""
    m = compile_mjcf(mjcf_file(xml))
    assert m.nv == 2 and list(m.dof_damping) == [3, 3] and list(m.dof_armature) == [2, 0]
    assert np.allclose(m.jnt_range[0], [-np.pi / 2, np.pi / 2])
    assert np.allclose(m.body_quat[1], [np.cos(np.pi / 4), 0, 0, np.sin(np.pi / 4)])
    assert np.allclose(m.geom_size[0][:2], [0.05, 0.2]) and np.allclose(m.geom_pos[0], [0, 0, 0.2])
    assert m.pair_condim.size == 0  # both geoms hang off a parent-child pair: filtered
