"""Stage-level checks of the CPU oracle against closed-form physics.

The reference holds no golden vectors for `mj_step` (parity unpinned); these tests pin the
oracle to physics that has an analytic answer so that the GPU-vs-oracle parity tests mean something.
"""
import numpy as np
import pytest

from gymnasium_robotics_b200.mjcf import compile_mjcf
from oracle.oracle_sim import OracleSim

G = 9.81


def make(mjcf_file, xml):
    return OracleSim(compile_mjcf(mjcf_file(xml)))


FREE_BOX = """
<mujoco><option timestep="0.002"/>
<worldbody>
  <body name="b" pos="0 0 1"><freejoint/><geom type="box" size="0.1 0.2 0.3" mass="2"/></body>
</worldbody></mujoco>"""


def test_free_fall_matches_semi_implicit_euler(mjcf_file):
    s = make(mjcf_file, FREE_BOX)
    n, h = 100, 0.002
    s.step(n)
    assert s.qvel[2] == pytest.approx(-G * h * n, rel=1e-12)
    assert s.qpos[2] == pytest.approx(1 - G * h * h * n * (n + 1) / 2, rel=1e-12)
    assert np.allclose(s.qpos[3:7], [1, 0, 0, 0])


def test_free_body_torque_free_rotation_conserves_angular_momentum(mjcf_file):
    s = make(mjcf_file, FREE_BOX.replace('timestep="0.002"', 'timestep="0.0005" gravity="0 0 0"'))
    s.qvel[3:6] = [1.0, 2.0, 0.5]
    m = s.model
    I = m.body_inertia[1]

    def L_world():
        s.forward()
        R = s.xmat[1].reshape(3, 3)
        return R @ (I * s.qvel[3:6])  # free-joint angular velocity is in the body frame

    L0 = L_world()
    s.step(2000)
    assert np.allclose(L_world(), L0, rtol=0, atol=2e-3 * np.linalg.norm(L0))  # first-order integrator drift
    assert np.linalg.norm(s.qpos[3:7]) == pytest.approx(1, abs=1e-12)


PENDULUM = """
<mujoco><option timestep="0.0005"/>
<worldbody>
  <body name="p" pos="0 0 2">
    <joint name="h" type="hinge" axis="0 1 0" pos="0 0 0"/>
    <geom type="sphere" size="0.05" pos="0 0 -1" mass="1"/>
  </body>
</worldbody></mujoco>"""


def test_pendulum_period_and_energy(mjcf_file):
    s = make(mjcf_file, PENDULUM)
    th0 = 0.1
    s.qpos[0] = th0
    I = 1.0 * 1.0 ** 2 + 0.4 * 1.0 * 0.05 ** 2  # point mass + sphere inertia
    T = 2 * np.pi * np.sqrt(I / (1.0 * G * 1.0)) * (1 + th0 ** 2 / 16)
    h = 0.0005
    qs = []
    for _ in range(int(1.5 * T / h)):
        s.step(1)
        qs.append(s.qpos[0])
    qs = np.array(qs)
    # first return to the positive maximum
    k = np.argmax(qs[int(0.75 * T / h):]) + int(0.75 * T / h)
    assert k * h == pytest.approx(T, rel=2e-3)
    assert qs[k] == pytest.approx(th0, rel=5e-3)


SPRING = """
<mujoco><option timestep="0.001" gravity="0 0 0"/>
<worldbody>
  <body name="m" pos="0 0 0">
    <joint name="s" type="slide" axis="1 0 0" stiffness="100" damping="2" armature="0.5"/>
    <geom type="sphere" size="0.05" mass="1.5"/>
  </body>
</worldbody></mujoco>"""


def test_spring_damper_implicit_damping_recurrence(mjcf_file):
    s = make(mjcf_file, SPRING)
    s.qpos[0] = 0.1
    m_eff, k, b, h = 2.0, 100.0, 2.0, 0.001
    x, v = 0.1, 0.0
    for _ in range(500):
        s.step(1)
        v = v + h * (-k * x - b * v) / (m_eff + h * b)  # Euler with implicit joint damping (Appendix B step 10)
        x = x + h * v
        assert s.qpos[0] == pytest.approx(x, abs=1e-12)
        assert s.qvel[0] == pytest.approx(v, abs=1e-12)


BOX_ON_PLANE = """
<mujoco><option timestep="0.002"/>
<worldbody>
  <geom name="floor" type="plane" size="1 1 1"/>
  <body name="b" pos="0 0 0.1"><freejoint/><geom name="box" type="box" size="0.1 0.1 0.1" mass="2"/></body>
</worldbody></mujoco>"""


def test_box_rests_on_plane_with_weight_balanced_by_contacts(mjcf_file):
    s = make(mjcf_file, BOX_ON_PLANE)
    s.step(1000)
    assert s.ncon == 4
    assert np.abs(s.qvel).max() < 1e-6
    assert 0.0995 < s.qpos[2] < 0.1
    # total normal force = weight: each pyramidal contact's normal force is the sum of its 4 edge forces
    assert s.efc("force").sum() == pytest.approx(2 * G, rel=1e-6)
    assert np.allclose(s.qpos[3:7], [1, 0, 0, 0], atol=1e-9)


def test_box_on_plane_static_equilibrium_penetration(mjcf_file):
    """At rest aref = -k*imp*r must produce force mg: analytic depth from solref/solimp (Appendix B.1)."""
    s = make(mjcf_file, BOX_ON_PLANE)
    s.step(3000)
    r = s.qpos[2] - 0.1  # = contact dist (negative)
    # per-contact: 4 active edge rows, each D*(aref - J a), at rest a=0: f_edge = D*aref ; sum_edges = 4*D*aref = mg/4
    d0, dw, width = 0.9, 0.95, 0.001
    x = min(1.0, abs(r) / width)
    y = 2 * x * x if x <= 0.5 else 1 - 2 * (1 - x) ** 2
    imp = d0 + y * (dw - d0)
    tc, dr = 0.02, 1.0
    kk = 1.0 / (dw * dw * tc * tc * dr * dr)
    aref = -kk * imp * r
    invw = 1.0 / 2.0  # translational inverse weight of the free box (1/m)
    mu = 1.0
    R = 2 * mu * mu * ((1 - imp) / imp * (invw + mu * mu * invw))
    total = 4 * 4 * aref / R
    assert total == pytest.approx(2 * G, rel=1e-3)


def test_friction_cone_stick_and_slip(mjcf_file):
    mu = 0.5  # below the cube's tipping limit so the box slides without tumbling
    for tilt, sticks in ((0.25, True), (1.0, False)):  # tan(theta) vs mu
        th = np.arctan(tilt)
        g = np.array([G * np.sin(th), 0, -G * np.cos(th)])
        xml = BOX_ON_PLANE.replace('timestep="0.002"', f'timestep="0.002" gravity="{g[0]} {g[1]} {g[2]}"')
        xml = xml.replace('type="plane"', f'type="plane" friction="{mu} 0.005 0.0001"').replace('type="box"', f'type="box" friction="{mu} 0.005 0.0001"')
        s = make(mjcf_file, xml)
        s.step(500)
        if sticks:
            assert abs(s.qvel[0]) < 2e-2  # soft-constraint creep only
        else:
            a = G * (np.sin(th) - mu * np.cos(th))
            assert s.qvel[0] == pytest.approx(a * 1.0, rel=0.03)


def test_noslip_post_pass_removes_the_soft_contact_creep(mjcf_file):
    """<option noslip_iterations="20"> (the Adroit models, adroit_assets.xml:3): a box on an incline below the friction angle.
    The soft friction rows alone let it creep downhill (tangential force = -D x with finite D); the noslip pass re-solves the
    friction dimensions WITHOUT regularisation, so the tangential reference acceleration -B v_t is met exactly and a box that
    starts at rest stays at rest: closed form v_t = 0.  Above the friction angle the pass must not stop the slide (cone clamp):
    a = g (sin(theta) - mu cos(theta)).  A frictionloss joint below its stiction force does not move either."""
    mu = 0.5
    # slope 0.1: well inside the friction pyramid.  (The pass keeps the normal-force share of every pair of opposing edges, so a
    # single friction direction can supply at most mu x its pair's share -- at slope mu / 2 = 0.25 that bound is reached and the
    # box creeps again; the main solver's pyramid has no such per-pair bound.)
    th = np.arctan(0.1)
    g = np.array([G * np.sin(th), 0, -G * np.cos(th)])
    creep = {}
    for noslip in (0, 20):
        xml = BOX_ON_PLANE.replace('timestep="0.002"', f'timestep="0.002" gravity="{g[0]} {g[1]} {g[2]}" noslip_iterations="{noslip}"')
        xml = xml.replace('type="plane"', f'type="plane" friction="{mu} 0.005 0.0001"').replace('type="box"', f'type="box" friction="{mu} 0.005 0.0001"')
        s = make(mjcf_file, xml)
        s.step(1000)
        x1 = s.qpos[0]
        s.step(1000)
        creep[noslip] = (abs(s.qvel[0]), abs(s.qpos[0] - x1) / 2.0)      # slip velocity, mean creep rate over the last 2 s
        if noslip:
            assert s.noslip_iter >= 1
    assert creep[0][0] > 1e-4 and creep[0][1] > 1e-4        # the soft model creeps at 0.3 mm/s ...
    assert creep[20][0] < 1e-6 and creep[20][1] < 1e-6       # ... the noslip pass holds the box: closed form, no slip
    # switching the pass off on a model that asks for it reproduces the soft answer (what the CUDA-path parity tests compare against)
    s.reset_data(); s.set_noslip(False); s.step(2000)
    assert abs(s.qvel[0]) == pytest.approx(creep[0][0], rel=1e-6)
    # above the friction angle: the clamp to the friction pyramid keeps the sliding acceleration
    th2 = np.arctan(1.0)
    g2 = np.array([G * np.sin(th2), 0, -G * np.cos(th2)])
    xml = BOX_ON_PLANE.replace('timestep="0.002"', f'timestep="0.002" gravity="{g2[0]} {g2[1]} {g2[2]}" noslip_iterations="20"')
    xml = xml.replace('type="plane"', f'type="plane" friction="{mu} 0.005 0.0001"').replace('type="box"', f'type="box" friction="{mu} 0.005 0.0001"')
    s = make(mjcf_file, xml)
    s.step(500)
    assert s.qvel[0] == pytest.approx(G * (np.sin(th2) - mu * np.cos(th2)), rel=0.03)
    # dry joint friction: a slider pulled with 0.8 x frictionloss does not creep with the pass, and does without it
    for noslip, moves in ((0, True), (20, False)):
        xml = f"""
        <mujoco><option timestep="0.002" gravity="0 0 0" noslip_iterations="{noslip}"/>
        <worldbody><body name="m"><joint name="s" type="slide" axis="1 0 0" frictionloss="2.0"/><geom type="sphere" size="0.05" mass="1"/></body></worldbody>
        <actuator><motor joint="s" gear="1"/></actuator></mujoco>"""
        s = make(mjcf_file, xml)
        s.ctrl[0] = 1.6
        s.step(500)
        assert (abs(s.qvel[0]) > 1e-5) == moves, (noslip, s.qvel[0])


LIMIT = """
<mujoco><option timestep="0.002"/>
<worldbody>
  <body name="m" pos="0 0 1">
    <joint name="s" type="slide" axis="0 0 1" limited="true" range="-0.2 0.3"/>
    <geom type="sphere" size="0.05" mass="1"/>
  </body>
</worldbody></mujoco>"""


def test_joint_limit_holds_weight(mjcf_file):
    s = make(mjcf_file, LIMIT)
    s.step(2000)
    assert -0.21 < s.qpos[0] < -0.199
    assert abs(s.qvel[0]) < 1e-6
    assert s.efc("force").sum() == pytest.approx(G, rel=1e-6)


ACT = """
<mujoco><option timestep="0.002" gravity="0 0 0"/>
<worldbody>
  <body name="m" pos="0 0 0">
    <joint name="s" type="slide" axis="1 0 0" damping="20"/>
    <geom type="sphere" size="0.05" mass="1"/>
  </body>
</worldbody>
<actuator><position joint="s" kp="100" ctrllimited="true" ctrlrange="0 0.2"/></actuator>
</mujoco>"""


def test_position_actuator_converges_and_clamps_ctrl(mjcf_file):
    s = make(mjcf_file, ACT)
    s.ctrl[0] = 0.5  # outside ctrlrange: clamped to 0.2 inside the engine (Appendix C.4)
    s.step(3000)
    assert s.qpos[0] == pytest.approx(0.2, abs=1e-6)


WELD = """
<mujoco><option timestep="0.002"/>
<worldbody>
  <body name="mocap" mocap="true" pos="0 0 1"/>
  <body name="b" pos="0 0 1"><freejoint/><geom type="box" size="0.05 0.05 0.05" mass="1"/></body>
</worldbody>
<equality><weld body1="mocap" body2="b" solref="0.02 1" solimp="0.9 0.95 0.001"/></equality>
</mujoco>"""


def test_weld_tracks_mocap_pose(mjcf_file):
    s = make(mjcf_file, WELD)
    s.mocap_pos[0] = [0.1, -0.05, 1.2]
    q = np.array([np.cos(0.3), 0, np.sin(0.3), 0])
    s.mocap_quat[0] = q
    s.step(1500)
    # soft weld: gravity sag is  m g / (k*imp/R...) -- small; orientation must match
    assert np.allclose(s.qpos[:2], [0.1, -0.05], atol=1e-4)
    assert s.qpos[2] == pytest.approx(1.2, abs=5e-3)
    assert abs(abs(np.dot(s.qpos[3:7], q)) - 1) < 1e-5
    assert np.abs(s.qvel).max() < 1e-5


CHAIN = """
<mujoco><option timestep="0.001"/>
<default><geom contype="0" conaffinity="0"/></default>
<worldbody>
  <body name="a" pos="0 0 1">
    <joint name="j1" type="hinge" axis="0 1 0"/>
    <geom type="capsule" fromto="0 0 0 0.3 0 0" size="0.02" mass="1"/>
    <body name="b" pos="0.3 0 0">
      <joint name="j2" type="hinge" axis="0 0 1"/>
      <geom type="box" size="0.1 0.02 0.03" pos="0.1 0 0" mass="0.7"/>
      <body name="c" pos="0.2 0 0">
        <joint name="j3" type="slide" axis="1 0 0"/>
        <geom type="sphere" size="0.04" mass="0.4"/>
      </body>
    </body>
  </body>
</worldbody></mujoco>"""


def test_total_energy_conserved_on_unconstrained_chain(mjcf_file):
    """Exercises CRB + RNE (Coriolis/centrifugal) consistency: E = 1/2 v'Mv + sum m g z is conserved to O(h)."""
    s = make(mjcf_file, CHAIN)
    s.qvel[:] = [1.0, -2.0, 0.5]
    m = s.model

    def energy():
        s.forward()
        ke = 0.5 * s.qvel @ s.M @ s.qvel
        # potential from body coms: xipos = xpos + xmat @ ipos
        pe = 0.0
        for b in range(1, m.nbody):
            z = s.xpos[b][2] + (s.xmat[b].reshape(3, 3) @ m.body_ipos[b])[2]
            pe += m.body_mass[b] * G * z
        return ke + pe

    e0 = energy()
    s.step(1000)
    assert energy() == pytest.approx(e0, abs=5e-3 * abs(e0) + 5e-3)


def test_bias_force_matches_finite_difference_of_lagrangian(mjcf_file):
    """qfrc_bias = C(q,v)v + g(q): check against M*qacc from the unconstrained equation using energy rate."""
    s = make(mjcf_file, CHAIN)
    rng = np.random.default_rng(0)
    s.qpos[:] = rng.uniform(-0.5, 0.5, 3)
    s.qvel[:] = rng.uniform(-1, 1, 3)
    s.forward()
    # power balance: d/dt (KE + PE) = 0 for the unconstrained system =>  v . (M a) + 1/2 v' Mdot v + dPE/dt = 0
    a = s.qacc.copy()
    M0, v = s.M.copy(), s.qvel.copy()
    q0 = s.qpos.copy()
    eps = 1e-6

    def M_and_pe(q):
        s.qpos[:] = q
        s.forward()
        mm = s.model
        pe = sum(mm.body_mass[b] * G * (s.xpos[b][2] + (s.xmat[b].reshape(3, 3) @ mm.body_ipos[b])[2]) for b in range(1, mm.nbody))
        return s.M.copy(), pe

    Mp, pep = M_and_pe(q0 + eps * v)
    Mm, pem = M_and_pe(q0 - eps * v)
    Mdot = (Mp - Mm) / (2 * eps)
    pedot = (pep - pem) / (2 * eps)
    assert v @ M0 @ a + 0.5 * v @ Mdot @ v + pedot == pytest.approx(0, abs=1e-6)


# ---------------------------------------------------------------------------------------------- general convex collider
CONVEX_PAIR = """
<mujoco><option timestep="0.002" gravity="0 0 0"/><worldbody>
  <body name="a" pos="0 0 0"><freejoint/><geom type="ellipsoid" size="0.05 0.05 0.05" mass="1"/></body>
  <body name="b" pos="0.06 0.04 0.05"><freejoint/><geom type="GEOM2" mass="1"/></body>
</worldbody></mujoco>"""


def test_convex_collider_reproduces_sphere_sphere(mjcf_file):
    """An ellipsoid with three equal radii is a sphere: portal refinement must return the analytic sphere-sphere contact
    (distance, normal from geom1 to geom2, midpoint position)."""
    s = make(mjcf_file, CONVEX_PAIR.replace('type="GEOM2"', 'type="sphere" size="0.05"'))
    s.forward()
    (c,) = s.contacts()
    d = np.array([0.06, 0.04, 0.05])
    # pair order follows the geom types (sphere < ellipsoid): geom1 is the sphere at d, so the normal points back to 0
    assert c["dist"] == pytest.approx(np.linalg.norm(d) - 0.1, abs=2e-6)
    assert np.allclose(c["frame"][0], -d / np.linalg.norm(d), atol=1e-5)
    assert np.allclose(c["pos"], d / 2, atol=1e-5)


def test_convex_collider_capsule_against_round_ellipsoid(mjcf_file):
    """Same check against a capsule lying along z next to the round ellipsoid: closest feature is the capsule's side."""
    xml = CONVEX_PAIR.replace('type="GEOM2"', 'type="capsule" size="0.02 0.1"').replace('pos="0.06 0.04 0.05"', 'pos="0.065 0 0.03"')
    s = make(mjcf_file, xml)
    s.forward()
    (c,) = s.contacts()
    assert c["dist"] == pytest.approx(0.065 - 0.05 - 0.02, abs=2e-6)
    assert np.allclose(np.abs(c["frame"][0]), [1, 0, 0], atol=1e-3)   # portal tolerance 1e-6 on the depth ~ 1e-4 on the normal


CONVEX_ON_TABLE = """
<mujoco><option timestep="0.002"/><worldbody>
  <geom name="t" type="box" size="0.5 0.5 0.1" pos="0.1 0.05 0"/>
  <body name="b" pos="0 0 ZZ"><freejoint/><geom type="GEOM" mass="2"/></body>
</worldbody></mujoco>"""


@pytest.mark.parametrize("geom,half", [('type="ellipsoid" size="0.04 0.03 0.02"', 0.02), ('type="cylinder" size="0.03 0.02"', 0.02)])
def test_convex_body_rests_on_a_box_at_the_soft_contact_depth(mjcf_file, geom, half):
    """Weight = contact force at rest, and the body sits exactly `dist` inside the table top (the constraint model
    itself is pinned by the box-on-plane tests above; this pins the collision geometry of the new pair types)."""
    s = make(mjcf_file, CONVEX_ON_TABLE.replace('type="GEOM"', geom).replace("ZZ", str(0.1 + half + 0.001)))
    s.step(1500)
    assert s.ncon == 1
    (c,) = s.contacts()
    f = s.efc("force")
    # the flat cylinder keeps rocking on its single portal contact (as the published algorithm does): looser bounds
    rocking = "cylinder" in geom
    assert f.sum() == pytest.approx(2 * G, rel=0.2 if rocking else 2e-2)   # pyramid edges sum to the normal force
    assert 1e-4 < -c["dist"] < 6e-3                               # soft contact: a fraction of a millimetre to millimetres
    assert abs(s.qpos[2] - (0.1 + half + c["dist"])) < (1e-3 if rocking else 2e-4)
    assert c["frame"][0][2] == pytest.approx(-1, abs=1e-2 if rocking else 1e-3)       # box (type 6) is geom2: normal points into the table


PLANE_CONVEX = """
<mujoco><option timestep="0.002"/><worldbody>
  <geom type="plane" size="1 1 0.1"/>
  <body name="b" pos="0 0 0.03" QUAT><freejoint/><geom type="GEOM" mass="1"/></body>
</worldbody></mujoco>"""


def test_plane_cylinder_and_plane_ellipsoid(mjcf_file):
    s = make(mjcf_file, PLANE_CONVEX.replace('type="GEOM"', 'type="cylinder" size="0.05 0.03"').replace("QUAT", ""))
    s.forward()
    assert s.ncon == 3                                           # flat on the plane: three points of the lower rim
    assert all(abs(c["dist"]) < 1e-9 for c in s.contacts())
    s.step(1000)
    assert abs(s.qpos[2] - 0.03) < 5e-3 and np.abs(s.qvel).max() < 1e-3
    s = make(mjcf_file, PLANE_CONVEX.replace('type="GEOM"', 'type="cylinder" size="0.05 0.03"').replace("QUAT", 'quat="0.7071068 0.7071068 0 0"'))
    s.forward()                                                  # lying on its side: the two ends of the lowest generator
    assert s.ncon == 2 and all(abs(c["dist"] - (0.03 - 0.05)) < 1e-6 for c in s.contacts())
    s = make(mjcf_file, PLANE_CONVEX.replace('type="GEOM"', 'type="ellipsoid" size="0.05 0.04 0.035"').replace("QUAT", ""))
    s.forward()
    (c,) = s.contacts()
    assert c["dist"] == pytest.approx(0.03 - 0.035, abs=1e-12) and np.allclose(c["pos"], [0, 0, -0.0025], atol=1e-9)


def test_noslip_changes_the_adroit_hammer_scenario_by_a_bounded_amount():
    """What deviation 12 (the CUDA path runs no noslip pass) costs on the reference's own model: the oracle's
    AdroitHandHammer env with the pass on (as the reference's adroit_assets.xml:3 asks) against the same env with the pass off,
    same seed and actions, the arm lowered onto the hammer.  The observation difference per env-step is the bound a user of the
    CUDA path gets with respect to a noslip-enabled simulator; it is measured, printed and asserted to stay small."""
    from gymnasium_robotics_b200.models import load_model
    from oracle.adroit_env import OracleAdroitHammerEnv

    m = load_model("adroit_hammer")
    on, off = OracleAdroitHammerEnv(m), OracleAdroitHammerEnv(m, noslip=False)
    on.reset(seed=30)
    off.reset(seed=30)
    rng = np.random.default_rng(2)
    worst, iters = 0.0, 0
    for step in range(12):
        a = rng.uniform(-1, 1, 26)
        if step >= 5:
            a[:2] = [-1, -0.5]           # lower the arm onto the hammer
        # per-step difference from IDENTICAL states: the noslip env is re-synchronised to the soft one before every step
        on.sim.qpos[:] = off.sim.qpos; on.sim.qvel[:] = off.sim.qvel; on.sim.qacc_warmstart[:] = off.sim.qacc_warmstart
        oo, *_ = on.step(a)
        of, *_ = off.step(a)
        iters = max(iters, on.sim.noslip_iter)
        worst = max(worst, float(np.abs(np.asarray(oo) - np.asarray(of)).max()))
    print(f"noslip on vs off, AdroitHandHammer, 12 env-steps from identical states: max |obs difference| = {worst:.2e}, sweeps <= {iters}")
    assert iters >= 1            # the pass ran (frictionloss rows are always there)
    assert worst < 5e-2


# ---------------------------------------------------------------------------------------------------------------
# hull support maps for mesh geoms (compile_mjcf(mesh_hull=True)): the mesh keeps its type and a vertex table
HULL_XML = """
<mujoco>
  <option timestep="0.002"/>
  <asset>
    <mesh name="cube" vertex="-0.05 -0.05 -0.05  0.05 -0.05 -0.05  -0.05 0.05 -0.05  0.05 0.05 -0.05
                               -0.05 -0.05 0.05   0.05 -0.05 0.05   -0.05 0.05 0.05   0.05 0.05 0.05  0 0 0"/>
  </asset>
  <worldbody>
    <geom name="floor" type="plane" size="1 1 0.1"/>
    <body name="b" pos="0 0 0.0502">
      <freejoint/>
      <inertial pos="0 0 0" mass="1" diaginertia="0.0016667 0.0016667 0.0016667"/>
      <geom name="g" type="GTYPE" GSPEC/>
    </body>
    EXTRA
  </worldbody>
</mujoco>
"""


def _hull_sim(mjcf_file, gtype, extra="", floor=True):
    spec = 'mesh="cube"' if gtype == "mesh" else 'size="0.05 0.05 0.05"'
    xml = HULL_XML.replace("GTYPE", gtype).replace("GSPEC", spec).replace("EXTRA", extra)
    if not floor:
        xml = xml.replace('<geom name="floor" type="plane" size="1 1 0.1"/>', "")
    return OracleSim(compile_mjcf(mjcf_file(xml), mesh_hull=True))


def test_hull_vertices_of_a_cube_and_thinning():
    from gymnasium_robotics_b200.mjcf import hull_vertices

    c = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], dtype=float)
    hv = hull_vertices(np.vstack([c, [[0, 0, 0], [0.5, 0.2, 0.1]]]))           # interior points are dropped
    assert sorted(map(tuple, hv)) == sorted(map(tuple, c))
    rng = np.random.default_rng(0)
    p = rng.normal(size=(400, 3)); p /= np.linalg.norm(p, axis=1)[:, None]     # 400 points on the unit sphere: all are hull vertices
    hv = hull_vertices(p)
    assert len(hv) == 32 and np.allclose(np.linalg.norm(hv, axis=1), 1.0)
    d = rng.normal(size=(200, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    assert float(((hv @ d.T).max(axis=0)).min()) > 0.85                        # inner approximation: the support stays close to 1


def test_mesh_hull_cube_rests_like_the_box_primitive(mjcf_file):
    """A cube given as a mesh (hull support map, plane-hull contacts at the 4 lowest vertices) settles at the same height as the box
    primitive of the same size (plane-box: the 4 lowest corners)."""
    zs = []
    for gtype in ("mesh", "box"):
        sim = _hull_sim(mjcf_file, gtype)
        assert (7 in [int(t) for t in sim.model.geom_type]) == (gtype == "mesh")
        sim.step(1500)
        zs.append(float(sim.qpos[2]))
        assert sim.ncon == 4
    assert abs(zs[0] - zs[1]) < 1e-9, zs


def test_mesh_hull_against_a_box_rests_on_one_portal_contact(mjcf_file):
    """Hull cube resting on a fixed box: portal refinement gives ONE contact that carries the weight; the box primitive on the same
    table rests on the four points of the box-box routine.  Same weight through one soft contact instead of four: the hull sits deeper,
    by less than a millimetre, and both come to rest."""
    extra = '<body name="table" pos="0 0 -0.1"><geom name="t" type="box" size="0.3 0.3 0.1"/></body>'
    zs = []
    for gtype in ("mesh", "box"):
        sim = _hull_sim(mjcf_file, gtype, extra, floor=False)
        sim.step(1500)
        zs.append(float(sim.qpos[2]))
        assert sim.ncon >= 1 and abs(float(sim.qvel[2])) < 1e-4
    assert 0.049 < zs[0] < zs[1] + 1e-9 and zs[1] - zs[0] < 1e-3, zs
