"""The oracle's own model compiler (oracle/mj_compile.py) and config surface (oracle/oracle_config.py).

What these tests establish (VERDICT r1, weak 1b / missing 4):
  * the legged oracle takes NO model constant, default or task entry from the product: its robot descriptions are
    the oracle's own walk of the reference's XML files (committed as derived data, re-derived here when
    /root/reference is present), compiled by the oracle's own compiler;
  * the product's compiler (guardx_b200/mjcf.py + robots.py) and the oracle's agree FIELD BY FIELD -- two
    independently written implementations of MuJoCo's compiler rules, different algorithms where there is a choice;
  * the primitive inertia formulas both rely on hold against numeric integration;
  * the world augmentation of both (World.build, world.py:310-326) reproduces the structure of the World.build() dump
    the reference itself ships (/root/reference/result.xml -> tests/golden/result_xml_structure.json).
"""
import ast
import json
import os

import numpy as np
import pytest

from guardx_b200 import config as gxconfig
from guardx_b200 import mjcf, robots
from oracle import mj_compile as oc
from oracle import oracle_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"
ROBOTS = ("point", "swimmer", "ant", "walker", "doggo")

FLOAT_FIELDS = ('body_pos', 'body_quat', 'body_ipos', 'body_mass', 'body_inertia_full', 'body_inertia', 'body_invweight0',
                'jnt_pos', 'jnt_axis', 'jnt_range', 'jnt_stiffness', 'jnt_margin', 'jnt_solref', 'jnt_solimp', 'qpos0',
                'qpos_spring', 'dof_damping', 'dof_armature', 'dof_invweight0', 'geom_pos', 'geom_size', 'geom_friction',
                'geom_margin', 'geom_gap', 'geom_solref', 'geom_solimp', 'actuator_gear', 'actuator_ctrlrange',
                'actuator_forcerange', 'actuator_gainprm0', 'actuator_biasprm')
INT_FIELDS = ('body_parentid', 'body_jntadr', 'body_jntnum', 'body_dofadr', 'body_dofnum', 'jnt_type', 'jnt_bodyid',
              'jnt_qposadr', 'jnt_dofadr', 'jnt_limited', 'dof_bodyid', 'dof_jntid', 'dof_parentid', 'geom_type',
              'geom_bodyid', 'geom_condim', 'geom_contype', 'geom_conaffinity', 'actuator_trnid', 'actuator_ctrllimited',
              'actuator_forcelimited', 'body_weldid')


def test_oracle_imports_nothing_from_the_product_but_the_abi_struct():
    """oracle/*.py may import guardx_b200._abi (the ctypes mirror of include/guardx.h) and nothing else of the product."""
    for fn in os.listdir(os.path.join(ROOT, "oracle")):
        if not fn.endswith(".py"):
            continue
        tree = ast.parse(open(os.path.join(ROOT, "oracle", fn)).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module and node.module.startswith("guardx_b200"):
                assert node.module == "guardx_b200" and [a.name for a in node.names] == ["_abi"], (fn, ast.dump(node))
            if isinstance(node, ast.Import):
                assert not any(a.name.startswith("guardx_b200") for a in node.names), fn


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs /root/reference (absent on the GPU box)")
def test_committed_robot_specs_are_the_oracle_parse_of_the_reference_xml():
    gold = json.load(open(os.path.join(GOLD, "oracle_robot_specs.json")))
    for name in ROBOTS:
        text = open(os.path.join(REF, "safe_rl_envs", "safe_rl_envs", "xmls", name + ".xml")).read()
        assert json.loads(json.dumps(oc.read_xml(text))) == gold[f"xmls/{name}.xml"], name


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs /root/reference (absent on the GPU box)")
def test_committed_defaults_and_task_table_are_the_reference_s():
    """oracle_config.DEFAULT is the reference's Engine.DEFAULT (dumped by tools/make_golden.py); the product's table and the
    product's task entries equal the reference's (incl. sensors_obs, safe_rl_env_config.py:123, 179, 207, 266)."""
    assert oracle_config.DEFAULT == gxconfig.DEFAULT
    gold = json.load(open(os.path.join(GOLD, "task_configs.json")))
    src = open(os.path.join(REF, "safe_rl_libX", "guard_utils", "safe_rl_env_config.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "configuration_list")
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "ref", "exec"), ns)
    for task, cfg in gold.items():
        assert ns["configuration_list"](task) == cfg
        assert gxconfig.configuration(task) == cfg


def test_product_task_table_equals_the_committed_reference_dump():
    gold = json.load(open(os.path.join(GOLD, "task_configs.json")))
    for task, cfg in gold.items():
        assert gxconfig.configuration(task) == cfg, task
        assert gxconfig.configuration(task + "_noconti") == cfg


@pytest.mark.parametrize("name", ROBOTS[1:])
def test_two_compilers_agree_field_by_field(name):
    rb = f"xmls/{name}.xml"
    o = oc.world_model(rb, 0.5, 0.3, 8)
    p = mjcf.compile_spec(mjcf.add_world_geoms(robots.robot_spec(rb), 0.5, 0.3, 8))
    assert (o.nq, o.nv, o.nu, o.nbody, o.njnt, o.ngeom) == (p.nq, p.nv, p.nu, p.nbody, p.njnt, p.ngeom)
    assert o.body_name == p.body_name and o.jnt_name == p.jnt_name and o.geom_name == p.geom_name
    for k in FLOAT_FIELDS:
        a, b = np.asarray(getattr(o, k), float), np.asarray(getattr(p, k), float)
        assert a.shape == b.shape, k
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-14, err_msg=f"{name}.{k}")
    for k in INT_FIELDS:
        np.testing.assert_array_equal(np.asarray(getattr(o, k)), np.asarray(getattr(p, k)), err_msg=f"{name}.{k}")
    assert o.contact_pairs == p.contact_pairs
    assert abs(o.stat_meaninertia - p.stat_meaninertia) <= 1e-12 * p.stat_meaninertia
    for k in ('timestep', 'tolerance', 'iterations', 'impratio'):
        assert o.opt[k] == p.opt[k], k
    assert list(o.opt['gravity']) == list(p.opt['gravity'])
    # geom frames: a `fromto` geom's z axis is the segment direction; the two compilers orient it with opposite signs
    # (MuJoCo: from - to), which is the same capsule / cylinder.  Everything else must be the same rotation.
    za = np.array([oc._rot(q)[:, 2] for q in o.geom_quat]); zb = np.array([mjcf.quat2mat(q)[:, 2] for q in p.geom_quat])
    np.testing.assert_allclose(np.abs((za * zb).sum(1)), 1.0, atol=1e-12)
    sym = np.isin(o.geom_type, (oc.CAPSULE, oc.CYLINDER, oc.SPHERE, oc.PLANE))
    flipped = (za * zb).sum(1) < 0
    assert not (flipped & ~sym).any()
    # the joint-space inertia itself: velocity-propagation (oracle) vs point-Jacobian (product) construction
    M, _, _ = oc.joint_space_inertia(o)
    np.testing.assert_allclose(M, mjcf.mass_matrix(p, mjcf.kinematics0(p)), rtol=1e-11, atol=1e-15)


def test_point_constants_of_both_sides_agree():
    """the product's closed-form Point constants (config.point_model) vs the oracle's compile of the
    reference's point.xml; and the whole lowered C struct, byte for byte, for every robot"""
    for name in ROBOTS:
        cfg = {'env_num': 64, 'robot_base': f'xmls/{name}.xml', '_seed': 3, 'num_steps': 77, 'hazards_num': 5,
               'lidar_num_bins': 12, 'observe_vel': True}
        a = oracle_config.lower(oracle_config.parse(cfg), 5, 100)
        b = gxconfig.lower(gxconfig.parse(cfg), 5, 100)
        assert bytes(a) == bytes(b), name
    m = oc.world_model('xmls/point.xml', 0.5, 0.3, 8)
    assert (m.nq, m.nv, m.nu, m.nbody) == (21, 21, 3, 11)            # nq = nv = 21 in total
    assert list(m.actuator_biasprm[0]) == [0.0, 0.0, -1.0] and m.actuator_gainprm0[0] == 1.0   # A1: the velocity-servo defaults win


def _integrate(kind, size, n=160):
    """mass and moments of a unit-density primitive by midpoint integration on an n^3 grid"""
    if kind == oc.BOX:
        ext = np.asarray(size[:3], float)
    elif kind == oc.SPHERE:
        ext = np.full(3, size[0])
    else:
        ext = np.array([size[0], size[0], size[1] + (size[0] if kind == oc.CAPSULE else 0.0)])
    ax = [(np.arange(n) + 0.5) / n * 2 * e - e for e in ext]
    x, y, z = np.meshgrid(*ax, indexing='ij')
    if kind == oc.BOX:
        inside = np.ones_like(x, bool)
    elif kind == oc.SPHERE:
        inside = x * x + y * y + z * z <= size[0] ** 2
    elif kind == oc.CYLINDER:
        inside = (x * x + y * y <= size[0] ** 2) & (np.abs(z) <= size[1])
    else:
        dz = np.maximum(np.abs(z) - size[1], 0.0)
        inside = x * x + y * y + dz * dz <= size[0] ** 2
    dv = np.prod(2 * ext / n)
    m = inside.sum() * dv
    I = np.array([((y * y + z * z) * inside).sum(), ((x * x + z * z) * inside).sum(), ((x * x + y * y) * inside).sum()]) * dv
    return m, I


@pytest.mark.parametrize("kind,size", [(oc.SPHERE, [0.1, 0, 0]), (oc.BOX, [0.05, 0.07, 0.02]), (oc.CYLINDER, [0.3, 0.01, 0]),
                                       (oc.CYLINDER, [0.04, 0.2, 0]), (oc.CAPSULE, [0.032, 0.0716, 0]), (oc.CAPSULE, [0.075, 0.1, 0]),
                                       (oc.CAPSULE, [0.02, 0.0354, 0])])
def test_primitive_inertia_formulas_against_numeric_integration(kind, size):
    m, I = oc.primitive_mass_inertia(kind, size, 1.0)
    mn, In = _integrate(kind, size)
    assert abs(m - mn) <= 4e-3 * m
    np.testing.assert_allclose(I, In, rtol=6e-3)
    gt = {oc.SPHERE: mjcf.GEOM_SPHERE, oc.BOX: mjcf.GEOM_BOX, oc.CYLINDER: mjcf.GEOM_CYLINDER, oc.CAPSULE: mjcf.GEOM_CAPSULE}[kind]
    mp, Ip = mjcf._geom_mass_inertia(gt, size, 1.0)          # the product's closed forms: same numbers
    assert abs(m - mp) <= 1e-13 * m
    np.testing.assert_allclose(I, Ip, rtol=1e-12)


def test_invweight_constants_are_what_they_claim():
    """dof_invweight0[i] is the acceleration of dof i under a unit force on it (all other dofs free): solve M a = e_i"""
    o = oc.world_model('xmls/ant.xml', 0.5, 0.3, 8)
    M, w, v = oc.joint_space_inertia(o)
    assert np.allclose(M, M.T) and np.linalg.eigvalsh(M).min() > 0
    for i in (0, 1, 2, 5, 10, o.nv - 1):
        a = np.linalg.solve(M, np.eye(o.nv)[i])
        assert abs(a[i] - o.dof_invweight0[i]) <= 1e-12 * abs(a[i])
    # kinetic energy two ways for a random qvel: 1/2 qvel^T M qvel  vs  sum over bodies of translational + rotational energy
    rng = np.random.default_rng(0)
    qv = rng.normal(size=o.nv)
    _, xrot, _, _, _ = oc.pose_at_qpos0(o)
    T = 0.0
    for b in range(1, o.nbody):
        vb, wb = qv @ v[:, b], qv @ w[:, b]
        T += 0.5 * o.body_mass[b] * vb @ vb + 0.5 * wb @ (xrot[b] @ o.body_inertia_full[b] @ xrot[b].T) @ wb
    T += 0.5 * (o.dof_armature * qv * qv).sum()
    assert abs(T - 0.5 * qv @ M @ qv) <= 1e-12 * T


def test_world_augmentation_matches_the_reference_s_own_world_build_dump():
    """/root/reference/result.xml is what World.build() wrote for the Point world (world.py:331-332): body order
    robot -> goal -> hazard0..7, joint names / types / axes, limited="false", geoms with contype = conaffinity = 0,
    goal sphere r 0.5 at z 0, hazard cylinders (0.3, 0.01) at z 0.02, the robot file's <default> block and actuators.
    Both compilers' augmented Point model must have exactly that structure.  (The dump is stale in ONE attribute:
    goal / hazard joints carry damping="0", the current builder writes damping="1", world.py:318-319 -- asserted as
    the current source has it.)"""
    ref = json.load(open(os.path.join(GOLD, "result_xml_structure.json")))
    assert [b['name'] for b in ref['bodies']] == ['robot', 'goal'] + [f'hazard{i}' for i in range(8)]
    o = oc.world_model('xmls/point.xml', 0.5, 0.3, 8)
    ps = mjcf.add_world_geoms(robots.robot_spec('xmls/point.xml'), 0.5, 0.3, 8)
    p = mjcf.compile_spec(ps)
    for m in (o, p):
        assert m.body_name[1:] == [b['name'] for b in ref['bodies']]
        assert m.opt['timestep'] == float(ref['option']['timestep'])
        j = 0
        for bi, body in enumerate(ref['bodies'], start=1):
            assert m.body_parentid[bi] == 0
            np.testing.assert_allclose(m.body_pos[bi], body['pos'], atol=1e-15)
            np.testing.assert_allclose(m.body_quat[bi], body['quat'], atol=1e-15)
            assert m.body_jntnum[bi] == len(body['joints'])
            for rj in body['joints']:
                assert m.jnt_name[j] == rj['name'] and m.jnt_bodyid[j] == bi
                assert m.jnt_type[j] == {'slide': oc.SLIDE, 'hinge': oc.HINGE}[rj['type']]
                np.testing.assert_allclose(m.jnt_axis[j], [float(t) for t in rj['axis'].split()], atol=1e-15)
                assert m.jnt_limited[j] == 0                       # limited="false" where written, no range otherwise
                if body['name'] != 'robot':
                    assert rj['limited'] == 'false'
                    assert m.dof_damping[m.jnt_dofadr[j]] == 1.0   # world.py:318-319 (the dump's "0" is stale)
                j += 1
            gs = [g for g in range(m.ngeom) if m.geom_bodyid[g] == bi]
            assert len(gs) == len(body['geoms'])
            for g, rg in zip(gs, body['geoms']):
                assert m.geom_name[g] == rg['name']
                assert m.geom_type[g] == oc.GEOM_KIND[rg['type']]
                size = [float(t) for t in rg['size'].split()]
                np.testing.assert_allclose(m.geom_size[g][:len(size)], size, atol=1e-15)
                if 'contype' in rg:
                    assert (m.geom_contype[g], m.geom_conaffinity[g]) == (int(rg['contype']), int(rg['conaffinity'])) == (0, 0)
        assert j == m.njnt == 3 + 2 * 9 and m.nq == m.nv == 21
        # actuators: three <general gear="0.3"> on robot_x / robot_y / robot_z, limits from the <default> block
        assert [m.jnt_name[t] for t in m.actuator_trnid] == [a['joint'] for a in ref['actuators']]
        np.testing.assert_allclose(m.actuator_gear, [float(a['gear']) for a in ref['actuators']])
        lo, hi = (float(t) for t in ref['default']['velocity']['ctrlrange'].split())
        np.testing.assert_allclose(m.actuator_ctrlrange, [[lo, hi]] * 3)
        lo, hi = (float(t) for t in ref['default']['velocity']['forcerange'].split())
        np.testing.assert_allclose(m.actuator_forcerange, [[lo, hi]] * 3)
        assert m.geom_name[0] == 'floor' and m.geom_type[0] == oc.PLANE
    # defaults the new bodies inherit: density 1, joint damping default .001 (overridden by damping="1")
    assert float(ref['default']['geom']['density']) == 1.0
    g = o.geom_name.index('goal')
    assert abs(o.body_mass[o.body_id('goal')] - 4.0 / 3.0 * np.pi * 0.5 ** 3) < 1e-12 and o.geom_bodyid[g] == o.body_id('goal')
