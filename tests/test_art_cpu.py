"""CPU checks of the articulated-robot path: the model compiler against golden constants
made from the reference's own robot files (tools/make_golden_mjcf.py), hand-computed
compiler rules, the numpy mjx.step restatement against the closed-form Point dynamics of the
C oracle, and the CUDA kernels' arithmetic (gx_art.cuh built as plain C++, LANES = 1)
against the numpy restatement.  No GPU needed."""
import ctypes as C
import json
import math
import os
import subprocess

import numpy as np
import pytest

from guardx_b200 import art_model, mjcf, robots
from oracle import mjx_numpy as mx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/safe_rl_envs/safe_rl_envs"


def _compile(name):
    return mjcf.compile_spec(mjcf.add_world_geoms(robots.robot_spec(name), 0.5, 0.3, 8))


@pytest.fixture(scope="module")
def gold(golden_dir):
    with open(os.path.join(golden_dir, "mjcf_constants.json")) as fh:
        return json.load(fh)


@pytest.mark.parametrize("name", ['point', 'ant', 'walker', 'swimmer'])
def test_builtin_robot_equals_golden_reference_compile(gold, name):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_mjcf import constants
    got = constants(_compile(name))
    for k, v in gold[name].items():
        if isinstance(v, list) and v and not isinstance(v[0], str):
            np.testing.assert_allclose(np.asarray(got[k], float), np.asarray(v, float), rtol=1e-12, atol=1e-15, err_msg=k)
        else:
            assert got[k] == v, k


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")
@pytest.mark.parametrize("f,name", list(robots.ROBOT_FILES.items()))
def test_builtin_robot_equals_parsed_reference_xml(f, name):
    a = mjcf.compile_spec(mjcf.parse_mjcf(open(os.path.join(REF, f)).read()))
    b = mjcf.compile_spec(robots.robot_spec(f))
    for k in ('body_pos', 'body_mass', 'body_ipos', 'body_inertia_full', 'jnt_axis', 'jnt_range', 'dof_damping',
              'dof_armature', 'dof_invweight0', 'body_invweight0', 'geom_size', 'actuator_gear'):
        np.testing.assert_array_equal(getattr(a, k), getattr(b, k), err_msg=k)
    assert a.jnt_name == b.jnt_name and a.contact_pairs == b.contact_pairs


def test_compiler_rules_by_hand():
    """Numbers that can be derived on paper from the MJCF (SURVEY.md Appendix C and xmls/ant.xml)."""
    p = _compile('point')
    r = p.body_id('robot')
    m1, m2 = 4 / 3 * math.pi * 1e-3, 1e-3
    assert p.body_mass[r] == pytest.approx(m1 + m2, rel=1e-14)
    assert p.body_ipos[r][0] == pytest.approx(m2 * 0.1 / (m1 + m2), rel=1e-14)
    cx = p.body_ipos[r][0]
    izz = 0.4 * m1 * 0.01 + m1 * cx * cx + m2 * (0.05 ** 2 * 2) / 3 + m2 * (0.1 - cx) ** 2
    assert p.body_inertia_full[r][2, 2] == pytest.approx(izz, rel=1e-13)
    assert p.body_mass[p.body_id('goal')] == pytest.approx(4 / 3 * math.pi * 0.125, rel=1e-14)      # density 1 (point.xml:5)
    assert p.actuator_biasprm[0].tolist() == [0, 0, -1] and p.actuator_forcelimited.tolist() == [1, 1, 1]   # A1
    a = _compile('ant')
    assert a.nq == 29 and a.nu == 8 and a.opt['timestep'] == 0.09
    j = a.jnt_id('ankle_1')
    np.testing.assert_allclose(a.jnt_range[j], [math.radians(30), math.radians(70)])     # compiler angle default: degree
    np.testing.assert_allclose(a.jnt_axis[j], [-math.sqrt(0.5), math.sqrt(0.5), 0])      # axis normalised
    r = a.body_id('robot')
    assert a.body_mass[r] == pytest.approx(5 * 4 / 3 * math.pi * 0.06 ** 3, rel=1e-14)      # torso sphere, density 5
    # capsule: r .02, half-length |(.05,.05,0)|/2, density 5
    h = math.hypot(0.05, 0.05)
    vol = math.pi * 0.02 ** 2 * h + 4 / 3 * math.pi * 0.02 ** 3
    assert a.body_mass[a.body_id('front_left_leg')] == pytest.approx(5 * vol, rel=1e-13)
    assert a.body_mass[a.body_id('goal')] == pytest.approx(5 * 4 / 3 * math.pi * 0.125, rel=1e-14)   # inherits density 5
    assert a.contact_pairs == [(0, 5), (0, 9), (0, 13), (0, 17)]                          # floor x 4 foot spheres only
    # root x slide (armature 0): its diagonal mass-matrix entry is the total robot mass; hip: armature 1 + inertia
    M = mjcf.mass_matrix(a, mjcf.kinematics0(a))
    assert M[0, 0] == pytest.approx(a.body_mass[1:14].sum(), rel=1e-12)
    assert 1 < M[3, 3] < 1.001 and np.allclose(M[:11, 11:], 0) and np.allclose(M, M.T)
    assert a.stat_meaninertia == pytest.approx(np.trace(M) / a.nv, rel=1e-14)


def test_numpy_mjx_step_equals_closed_form_point_oracle(oracle_lib):
    """Generic CRB / RNE / actuation / implicit-damping Euler on the compiled Point model vs the
    hand-derived 3-DoF dynamics of oracle/gx_oracle.c: two derivations, one answer."""
    from oracle.oracle import OracleEngine
    m = _compile('point')
    n = 64
    o = OracleEngine({'env_num': n, '_seed': 1})
    o.reset()
    rng = np.random.default_rng(0)
    for _ in range(30):
        o.step(rng.uniform(-1, 1, (n, 2)).astype(np.float32))
    s = o.get_state()
    a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    ctrl = np.stack([a[:, 0] * s[:, 6], a[:, 0] * s[:, 7], a[:, 1]], 1)
    _, _, _, info = o.step(a, want_qacc=True)
    s2 = o.get_state()
    mm = mx.MjxModel(m, np.float64)
    qpos = np.zeros((n, m.nq)); qvel = np.zeros((n, m.nv))
    qpos[:, :3] = s[:, 0:3]; qvel[:, :3] = s[:, 3:6]; qpos[:, 3:5] = s[:, 11:13]; qpos[:, 5:21] = s[:, 13:29]
    q2, v2, d = mx.step(mm, qpos, qvel, ctrl, np.zeros((n, m.nv)))
    np.testing.assert_allclose(q2[:, :3], s2[:, 0:3], atol=2e-7)
    np.testing.assert_allclose(v2[:, :3], s2[:, 3:6], atol=3e-6)
    np.testing.assert_allclose(d['qacc'][:, :3], info['qacc'], rtol=2e-6, atol=1e-3)
    assert np.abs(v2[:, 3:]).max() == 0          # goal / hazard bodies never move
    np.testing.assert_allclose(d['xmat'][:, 1, 0, 0], s2[:, 6], atol=3e-7)   # lagged heading = this forward's R


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("emu") / "libartemu.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-x", "c++",
                    "-I" + os.path.join(ROOT, "guardx_b200", "csrc"), os.path.join(ROOT, "tests", "emu", "art_emu.cpp"),
                    "-o", out], check=True)
    lib = C.CDLL(out)
    assert lib.art_emu_model_size() == C.sizeof(art_model.GxArtModel)
    return lib


@pytest.mark.parametrize("name", ['ant', 'walker', 'swimmer'])
def test_kernel_arithmetic_on_host_equals_numpy_oracle(emu, name):
    """gx_art.cuh compiled as C++ (fp32, leg-wise arrowhead solves) vs the dense fp64 numpy
    restatement: one step from identical states along an 80-step random-action trajectory with
    joint limits and foot contacts active."""
    m = _compile(name)
    am, info = art_model.lower_art(m)
    nl, K, NB = info['shape']
    n, nvr = 24, info['nv']
    mm = mx.MjxModel(m, np.float64)
    rng = np.random.default_rng(0)
    qpos = np.zeros((n, m.nq)); qvel = np.zeros((n, m.nv)); warm = np.zeros((n, m.nv))
    qpos[:, 0] = rng.uniform(-1, 1, n)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    active_rows = 0
    for t in range(80):
        ctrl = rng.uniform(-1.2, 1.2, (n, m.nu))
        q1, v1, w1 = (x[:, :nvr].astype(np.float32).copy() for x in (qpos, qvel, warm))
        snap = np.zeros((n, 4), np.float32); it = np.zeros(n, np.int32)
        assert emu.art_emu_step(C.byref(am), nl, K, NB, n, p(q1), p(v1), p(w1), p(ctrl.astype(np.float32)), 0, p(snap), p(it)) == 0
        qpos, qvel, d = mx.step(mm, qpos, qvel, ctrl, warm)
        warm = d['qacc']
        scale = 1 + np.abs(d['qacc'][:, :nvr]).max()
        assert np.abs(q1 - qpos[:, :nvr]).max() < 3e-6, t
        assert np.abs(v1 - qvel[:, :nvr]).max() < 3e-6 * scale * m.opt['timestep'] * 20, t
        assert np.abs(w1 - d['qacc'][:, :nvr]).max() < 2e-5 * scale, t
        assert np.abs(snap[:, 0] - d['xpos'][:, 1, 0]).max() < 1e-6 and np.abs(snap[:, 2] - d['xmat'][:, 1, 0, 0]).max() < 1e-6
        assert it.max() <= 12
        active_rows = max(active_rows, int((d['efc_D'] > 0).sum(1).max()))
    assert active_rows >= (2 if name == 'swimmer' else 8)      # limits (and contacts) really were active


def test_art_model_lowering_and_rejections():
    am, info = art_model.lower_art(_compile('ant'))
    assert info['shape'] == (4, 2, 3) and am.nv == 11 and am.nu == 8 and (am.root_x, am.root_y) == (0, 2)
    assert [am.leg[l].con.on for l in range(4)] == [1, 1, 1, 1] and am.leg[0].con.body == 2
    assert am.leg[0].dof[1].act == 1 and am.leg[3].dof[0].act == 6
    am, info = art_model.lower_art(_compile('walker'))
    assert info['shape'] == (2, 5, 3) and (am.root_x, am.root_y) == (0, 2)
    am, info = art_model.lower_art(_compile('swimmer'))
    assert info['shape'] == (1, 2, 2) and (am.root_x, am.root_y) == (0, 1) and len(info['ignored_pairs']) == 3
    with pytest.raises(AssertionError):
        robots.robot_spec('xmls/car.xml')              # named by world.py:55, absent from the reference tree
    with pytest.raises(AssertionError, match='planar root'):
        art_model.lower_art(_compile('doggo'))         # Doggo goes to the general-tree kernels instead


# ---------------------------------------------------------------------------
# The mjx.step restatement against closed-form rigid-body results.  MuJoCo itself is absent
# (parity unpinned), so these anchor the oracle -- and through the oracle the kernels -- to
# physics that can be checked on paper.
# ---------------------------------------------------------------------------
def _pendulum(length=0.3, radius=0.02, density=800.0, armature=0.0, damping=0.0, axis=(0., 1., 0.)):
    """One capsule hanging from a hinge at the world origin along -z (plus the planar root the
    lowering expects is NOT needed for the oracle: it takes any hinge/slide tree)."""
    s = mjcf.Spec()
    s.opt = dict(mjcf.OPT_DEFAULT, timestep=0.001)
    s.angle = 'radian'
    s.bodies.append(dict(name='world', parent=-1, pos=(0., 0., 0.), quat=(1., 0., 0., 0.), joints=[], geoms=[]))
    s.bodies.append(dict(name='rod', parent=0, pos=(0., 0., 1.0), quat=(1., 0., 0., 0.),
                         joints=[dict(type='hinge', axis=axis, name='h', armature=armature, damping=damping)],
                         geoms=[dict(type='capsule', fromto=(0., 0., 0., 0., 0., -length), size=(radius,), density=density,
                                     contype=0, conaffinity=0)]))
    return mjcf.compile_spec(s)


def test_pendulum_acceleration_and_energy():
    m = _pendulum()
    mm = mx.MjxModel(m, np.float64)
    mass, com = m.body_mass[1], -m.body_ipos[1][2]
    I_pivot = m.body_inertia_full[1][1, 1] + mass * com * com          # parallel axis, about the hinge (y axis)
    for th in (0.0, 0.3, -1.2, 2.5):
        d = mx.forward(mm, np.array([[th]]), np.zeros((1, 1)), np.zeros((1, 0)), np.zeros((1, 1)))
        assert d['qM'][0, 0, 0] == pytest.approx(I_pivot, rel=1e-12)
        # rotation by +theta about y moves the bob (0,0,-l) to (-l sin, 0, -l cos): torque = -m g l sin(theta)
        assert d['qacc'][0, 0] == pytest.approx(-mass * 9.81 * com * np.sin(th) / I_pivot, rel=1e-10, abs=1e-12)
    # semi-implicit Euler at dt = 1 ms keeps the energy of a large swing within 0.5 % over 2 s
    q, v = np.array([[1.0]]), np.zeros((1, 1))
    energy = lambda q, v: 0.5 * I_pivot * v[0, 0] ** 2 - mass * 9.81 * com * np.cos(q[0, 0])
    e0 = energy(q, v)
    qmax = 0.0
    for _ in range(2000):
        q, v, _ = mx.step(mm, q, v, np.zeros((1, 0)), np.zeros((1, 1)))
        qmax = max(qmax, abs(q[0, 0]))
    assert abs(energy(q, v) - e0) < 5e-3 * abs(e0) and 0.99 < qmax < 1.01


def test_centrifugal_and_coriolis_terms_of_a_two_link_arm():
    """Planar double pendulum in a horizontal plane (hinge axes = z, so gravity does no work):
    the bias forces must equal the textbook  h * (-(2 w1 w2 + w2^2), w1^2)  with h = m2 l1 lc2 sin(q2)."""
    s = mjcf.Spec()
    s.opt = dict(mjcf.OPT_DEFAULT, timestep=0.001)
    s.angle = 'radian'
    s.bodies.append(dict(name='world', parent=-1, pos=(0., 0., 0.), quat=(1., 0., 0., 0.), joints=[], geoms=[]))
    cap = dict(type='capsule', size=(0.02,), density=500.0, contype=0, conaffinity=0)
    s.bodies.append(dict(name='l1', parent=0, pos=(0., 0., 0.5), quat=(1., 0., 0., 0.),
                         joints=[dict(type='hinge', axis=(0., 0., 1.), name='j1')], geoms=[dict(cap, fromto=(0., 0., 0., 0.4, 0., 0.))]))
    s.bodies.append(dict(name='l2', parent=1, pos=(0.4, 0., 0.), quat=(1., 0., 0., 0.),
                         joints=[dict(type='hinge', axis=(0., 0., 1.), name='j2')], geoms=[dict(cap, fromto=(0., 0., 0., 0.3, 0., 0.))]))
    m = mjcf.compile_spec(s)
    mm = mx.MjxModel(m, np.float64)
    m1, m2 = m.body_mass[1], m.body_mass[2]
    lc1, lc2, l1 = m.body_ipos[1][0], m.body_ipos[2][0], 0.4
    I1, I2 = m.body_inertia_full[1][2, 2], m.body_inertia_full[2][2, 2]
    q = np.array([[0.7, -1.1]]); w = np.array([[1.3, -2.2]])
    d = mx.forward(mm, q, w, np.zeros((1, 0)), np.zeros((1, 2)))
    c2, s2 = np.cos(q[0, 1]), np.sin(q[0, 1])
    M11 = I1 + I2 + m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * c2)
    M12 = I2 + m2 * (lc2 ** 2 + l1 * lc2 * c2)
    M22 = I2 + m2 * lc2 ** 2
    np.testing.assert_allclose(d['qM'][0], [[M11, M12], [M12, M22]], rtol=1e-12)
    h = m2 * l1 * lc2 * s2
    np.testing.assert_allclose(d['qfrc_bias'][0], [-h * (2 * w[0, 0] * w[0, 1] + w[0, 1] ** 2), h * w[0, 0] ** 2], rtol=1e-10, atol=1e-14)


def test_joint_limit_and_contact_are_soft_springs_with_the_documented_constants():
    """A limited hinge pushed 0.1 rad past its range and a sphere 5 mm into the floor: the solver's
    force must be the solref/solimp spring  D * (aref - J qacc)  with k = 1/(dmax^2 tc^2 zeta^2),
    b = 2/(dmax tc), imp = dmax beyond `width`, R = invweight (1-imp)/imp -- checked in closed form
    for the 1-DoF limit case."""
    m = _pendulum(armature=0.5, axis=(0., 1., 0.))
    m.jnt_limited[0] = 1
    m.jnt_range[0] = [-0.2, 0.2]
    mm = mx.MjxModel(m, np.float64)
    th, w = 0.3, 0.4
    d = mx.forward(mm, np.array([[th]]), np.array([[w]]), np.zeros((1, 0)), np.zeros((1, 1)))
    dt, tc, dmax = m.opt['timestep'], 0.02, 0.95
    k, b = 1 / (dmax ** 2 * tc ** 2), 2 / (dmax * tc)
    pos = 0.2 - th                                   # -0.1: past the upper limit, J = -1
    aref = -b * (-w) - k * dmax * pos
    R = m.dof_invweight0[0] * (1 - dmax) / dmax
    M = d['qM'][0, 0, 0]
    fs = d['qfrc_smooth'][0, 0]
    # minimise 0.5 M (a - a_s)^2 + 0.5/R (-a - aref)^2  =>  a = (fs - aref/R) / (M + 1/R)
    a = (fs - aref / R) / (M + 1 / R)
    assert (-a - aref) < 0                           # the row is active
    assert d['qacc'][0, 0] == pytest.approx(a, rel=1e-9)
    assert d['efc_force'][0, 0] == pytest.approx(-(1 / R) * (-a - aref), rel=1e-9) and d['efc_force'][0, 0] > 0


def test_mjcf_parser_rules_and_rejections():
    """Defaults resolution, the actuator-shortcut rule behind assumption A1, autolimits, fromto
    frames, and clear errors for what the compiler does not cover."""
    xml = """<mujoco><option timestep="0.005"/>
      <default><joint damping="2" armature="0.5"/><geom density="10" contype="0" conaffinity="0"/>
        <motor ctrllimited="true" ctrlrange="-1 1" forcelimited="true" forcerange="-3 3"/><velocity kv="4"/></default>
      <worldbody><geom name="floor" type="plane" size="1 1 .1" conaffinity="1"/>
        <body name="robot" pos="0 0 .2">
          <joint type="slide" axis="2 0 0" name="robot_x"/><joint type="slide" axis="0 1 0" name="robot_y" damping="7"/>
          <joint type="hinge" axis="0 0 1" name="robot_z" range="-90 90"/>
          <geom type="capsule" fromto="0 0 0 0 0.4 0" size="0.05"/>
          <body name="tip" pos="0 0.4 0"><joint name="j" axis="1 0 0" limited="false" range="-10 10"/>
            <geom type="sphere" size="0.03" contype="1"/></body></body></worldbody>
      <actuator><general joint="j" gear="5"/><motor joint="robot_z" gear="2"/></actuator></mujoco>"""
    m = mjcf.compile_spec(mjcf.parse_mjcf(xml))
    assert m.opt['timestep'] == 0.005 and m.nq == 4 and m.nu == 2
    np.testing.assert_allclose(m.jnt_axis[0], [1, 0, 0])                       # normalised
    assert m.dof_damping.tolist() == [2, 7, 2, 2] and m.dof_armature.tolist() == [0.5] * 4
    assert m.jnt_limited.tolist() == [0, 0, 1, 0]                              # autolimits; explicit false wins
    np.testing.assert_allclose(m.jnt_range[2], [-math.pi / 2, math.pi / 2])    # degrees by default
    # <general> inherits the default actuator as left by the LAST shortcut (<velocity kv=4>): gain 4, bias (0,0,-4);
    # <motor> resets it to gain 1, no bias; both keep the limits of the defaults class
    np.testing.assert_allclose(m.actuator_gainprm0, [4, 1]); np.testing.assert_allclose(m.actuator_biasprm, [[0, 0, -4], [0, 0, 0]])
    assert m.actuator_forcelimited.tolist() == [1, 1] and m.actuator_gear.tolist() == [5, 2]
    # fromto capsule: centre at the midpoint, z axis along +y, half-length 0.2
    g = 1
    np.testing.assert_allclose(m.geom_pos[g], [0, 0.2, 0]); assert m.geom_size[g][1] == pytest.approx(0.2)
    np.testing.assert_allclose(mjcf.quat2mat(m.geom_quat[g])[:, 2], [0, 1, 0], atol=1e-12)
    assert m.contact_pairs == [(0, 2)]                                        # only the tip sphere meets the floor
    for bad, exc in (('<mujoco><default><default class="a"/></default><worldbody/></mujoco>', NotImplementedError),
                     ('<mujoco><compiler coordinate="global"/><worldbody/></mujoco>', NotImplementedError)):
        with pytest.raises(exc):
            mjcf.parse_mjcf(bad)
    free = mjcf.parse_mjcf('<mujoco><worldbody><body name="robot"><freejoint name="robot"/><geom size=".1"/></body></worldbody></mujoco>')
    fm = mjcf.compile_spec(free)                          # free joints compile (groundwork for Doggo) ...
    assert fm.nq == 7 and fm.nv == 6 and fm.qpos0.tolist() == [0, 0, 0, 1, 0, 0, 0]
    with pytest.raises(AssertionError, match="planar root"):
        art_model.lower_art(fm)                           # ... but the kernels of this round refuse them
    # topology the kernels do not cover is refused by the lowering, not silently mis-simulated
    tree = mjcf.parse_mjcf(xml.replace('range="-90 90"', '').replace('<body name="tip" pos="0 0.4 0">', '<body name="side" pos="0.1 0 0"><joint name="s" axis="0 0 1"/><geom size=".02"/><body name="a"><joint name="a1"/><geom size=".02"/></body><body name="b"><joint name="b1"/><geom size=".02"/></body></body><body name="tip" pos="0 0.4 0">'))
    with pytest.raises(AssertionError, match="serial chains"):
        art_model.lower_art(mjcf.compile_spec(tree))


# ---------------------------------------------------------------------------
# Free-joint root and capsule contacts (Doggo groundwork: compiled and restated in the oracle;
# the CUDA kernels for it are the next row of SURVEY 8f-2)
# ---------------------------------------------------------------------------
def test_doggo_description_equals_golden_reference_compile(gold):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_golden_mjcf import constants
    m = mjcf.compile_spec(mjcf.add_world_geoms(robots.doggo(), 0.5, 0.3, 8))
    got = constants(m)
    for k, v in gold['doggo'].items():
        if isinstance(v, list) and v and not isinstance(v[0], str) and v[0] is not None:
            np.testing.assert_allclose(np.asarray(got[k], float), np.asarray(v, float), rtol=1e-12, atol=1e-15, err_msg=k)
    assert (m.nq, m.nv, m.nu) == (38, 37, 12) and m.qpos0[:7].tolist() == [0, 0, 0.3, 1, 0, 0, 0]
    # <freejoint> ignores the defaults class (stiffness 0.01 / limited); the goal / hazard slides inherit it
    assert m.jnt_stiffness[0] == 0 and m.jnt_limited[0] == 0
    assert m.jnt_stiffness[m.jnt_id('goal_x')] == 0.01 and m.jnt_stiffness[m.jnt_id('hip_1_y')] == 0.01
    assert m.qpos_spring[m.jnt_qposadr[m.jnt_id('ankle_3')]] == pytest.approx(math.radians(-20))
    # weld-id filters: 14 capsules x floor + 60 capsule pairs (same welded body and weld parent-child excluded)
    assert len(m.contact_pairs) == 74 and sum(1 for a, b in m.contact_pairs if a == 0) == 14


def _free_body(geom, pos=(0., 0., 1.0)):
    s = mjcf.Spec()
    s.opt = dict(mjcf.OPT_DEFAULT, timestep=0.001)
    s.angle = 'radian'
    s.bodies.append(dict(name='world', parent=-1, pos=(0., 0., 0.), quat=(1., 0., 0., 0.), joints=[],
                         geoms=[dict(name='floor', type='plane', size=(5., 5., .1))]))
    s.bodies.append(dict(name='robot', parent=0, pos=pos, quat=(1., 0., 0., 0.), joints=[dict(type='free', name='robot')], geoms=[geom]))
    return mjcf.compile_spec(s)


def test_free_body_falls_and_tumbles_like_a_rigid_body():
    m = _free_body(dict(type='box', size=(0.05, 0.1, 0.2), density=700.0))
    mm = mx.MjxModel(m, np.float64, capsules=False)
    mass, I = m.body_mass[1], np.diag(m.body_inertia_full[1])
    q = m.qpos0[None].copy(); v = np.zeros((1, 6))
    v[0, 3:] = [3.0, 0.2, -0.1]          # spin mostly about the LONGEST principal axis... of a box: unstable middle axis is y
    v[0, :3] = [0.5, 0.0, 2.0]
    d = mx.forward(mm, q, v, np.zeros((1, 0)), np.zeros((1, 6)))
    np.testing.assert_allclose(d['qM'][0], np.diag([mass] * 3 + list(I)), atol=1e-14)
    # linear: gravity only; angular (body frame): Euler's equations  I w' = -(w x I w)
    w = v[0, 3:]
    np.testing.assert_allclose(d['qacc'][0, :3], [0, 0, -9.81], atol=1e-10)
    np.testing.assert_allclose(d['qacc'][0, 3:], -np.cross(w, I * w) / I, rtol=1e-10, atol=1e-12)
    L0 = mjcf.quat2mat(q[0, 3:7]) @ (I * w)
    e0 = 0.5 * np.sum(I * w * w)
    for _ in range(1500):
        q, v, _ = mx.step(mm, q, v, np.zeros((1, 0)), np.zeros((1, 6)))
    w = v[0, 3:]
    L1 = mjcf.quat2mat(q[0, 3:7]) @ (I * w)
    assert np.linalg.norm(q[0, 3:7]) == pytest.approx(1.0, abs=1e-12)
    np.testing.assert_allclose(L1, L0, rtol=2e-3, atol=2e-5)            # angular momentum in the world frame
    assert 0.5 * np.sum(I * w * w) == pytest.approx(e0, rel=5e-3)       # rotational energy
    assert v[0, 2] == pytest.approx(2.0 - 9.81 * 1.5, rel=1e-9) and q[0, 0] == pytest.approx(0.5 * 1.5, rel=1e-9)


def test_capsule_resting_on_the_floor_carries_its_weight():
    m = _free_body(dict(type='capsule', fromto=(-0.1, 0., 0., 0.1, 0., 0.), size=(0.04,), density=900.0), pos=(0., 0., 0.041))
    mm = mx.MjxModel(m, np.float64)
    assert len(mm.capsule_plane) == 1
    q = m.qpos0[None].copy(); v = np.zeros((1, 6)); warm = np.zeros((1, 6))
    for _ in range(600):
        q, v, d = mx.step(mm, q, v, np.zeros((1, 0)), warm)
        warm = d['qacc']
    assert np.abs(v).max() < 1e-4 and np.abs(d['qacc']).max() < 1e-3          # at rest
    assert -0.002 < q[0, 2] - 0.04 < 0                                           # a fraction of a mm of soft penetration
    # two contacts (the capsule's ends) x 4 pyramid edges; every edge force pushes along n +- mu t, so the normal
    # components add up to the weight
    assert d['efc_force'].shape[1] == 8 and (d['efc_force'] > 0).all()
    assert d['efc_force'].sum() == pytest.approx(m.body_mass[1] * 9.81, rel=2e-3)
    np.testing.assert_allclose(d['qfrc_constraint'][0, 2], m.body_mass[1] * 9.81, rtol=2e-3)


def test_capsule_capsule_closest_points_against_brute_force():
    rng = np.random.default_rng(0)
    n = 200
    a0, a1, b0, b1 = (rng.uniform(-1, 1, (n, 3)) for _ in range(4))
    pa, pb = mx.closest_segment_to_segment_points(a0, a1, b0, b1)
    t = np.linspace(0, 1, 201)
    A = a0[:, None, :] + t[None, :, None] * (a1 - a0)[:, None, :]
    B = b0[:, None, :] + t[None, :, None] * (b1 - b0)[:, None, :]
    brute = np.sqrt(((A[:, :, None, :] - B[:, None, :, :]) ** 2).sum(-1)).reshape(n, -1).min(1)
    got = np.linalg.norm(pa - pb, axis=1)
    assert (got <= brute + 1e-9).all() and (got >= brute - 0.02).all()          # never worse than the sampled optimum
    # the returned points lie on their segments
    for p, s0, s1 in ((pa, a0, a1), (pb, b0, b1)):
        u = np.einsum('ni,ni->n', p - s0, s1 - s0) / np.einsum('ni,ni->n', s1 - s0, s1 - s0)
        assert (u > -1e-9).all() and (u < 1 + 1e-9).all()
        assert np.linalg.norm(np.cross(p - s0, s1 - s0), axis=1).max() < 1e-9
