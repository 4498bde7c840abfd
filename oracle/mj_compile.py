"""Oracle-side MuJoCo model compiler (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.py).

PARITY UNPINNED: ``mujoco`` is an un-vendored, un-pinned dependency of the reference
(requirements.txt:3; 3.1.x by API use) and is not installable here.  This
file restates, from MuJoCo's published compiler rules, what
``mujoco.MjModel.from_xml_string(World.build())`` (reference world.py:104-333, engine.py:228)
produces for the robot files the JAX Engine runs (xmls/{point,swimmer,ant,walker,doggo}.xml).

It exists so that the legged-robot oracle (oracle/mjx_numpy.py, oracle/art_oracle.py) does NOT
share its model constants with the product: nothing here imports guardx_b200, the XML walk, the
geom inertia formulas, the inertial-frame merge, the joint-space inertia at qpos0 and the
``*_invweight0`` / ``stat.meaninertia`` constants are written independently of
guardx_b200/mjcf.py (different algorithms where there is a choice: velocity propagation instead of
point Jacobians for M, hemisphere decomposition for capsules), and tests/test_oracle_model.py
asserts the two compilers agree field by field and checks the inertia formulas against numeric
integration.

Three stages:
  read_xml(text)            MJCF text -> {'option', 'angle', 'bodies': [...], 'actuators': [...]}
                            with every element's attributes ALREADY merged over the file's
                            <default> block and MuJoCo's built-in defaults (plain JSON-able data;
                            tests/golden/oracle_robot_specs.json is this stage's output for the
                            reference's five robot files, written by tests/tools/make_oracle_specs.py,
                            because /root/reference does not exist on the GPU box)
  with_world(spec, ...)     + the goal / hazard bodies World.build appends (world.py:310-326),
                            which carry no class and therefore inherit the file's defaults
  build(spec)               -> OModel: arrays named like mjModel, float64
"""
from __future__ import annotations

import json
import math
import os
import xml.etree.ElementTree as ET
from types import SimpleNamespace

import numpy as np

# enum values of mjtJoint / mjtGeom (mjmodel.h)
FREE, BALL, SLIDE, HINGE = 0, 1, 2, 3
PLANE, SPHERE, CAPSULE, CYLINDER, BOX = 0, 2, 3, 5, 6
JOINT_KIND = {'free': FREE, 'ball': BALL, 'slide': SLIDE, 'hinge': HINGE}
GEOM_KIND = {'plane': PLANE, 'sphere': SPHERE, 'capsule': CAPSULE, 'cylinder': CYLINDER, 'box': BOX}
MINVAL = 1e-15

# MuJoCo's built-in element defaults (XML reference, "default" column)
_JOINT0 = {'type': 'hinge', 'pos': [0, 0, 0], 'axis': [0, 0, 1], 'range': [0, 0], 'limited': 'auto', 'damping': 0.0,
           'armature': 0.0, 'stiffness': 0.0, 'springref': 0.0, 'ref': 0.0, 'margin': 0.0,
           'solreflimit': [0.02, 1.0], 'solimplimit': [0.9, 0.95, 0.001, 0.5, 2.0]}
_GEOM0 = {'type': 'sphere', 'pos': [0, 0, 0], 'quat': [1, 0, 0, 0], 'size': [0, 0, 0], 'fromto': None,
          'density': 1000.0, 'friction': [1.0, 0.005, 0.0001], 'margin': 0.0, 'gap': 0.0, 'condim': 3,
          'contype': 1, 'conaffinity': 1, 'solref': [0.02, 1.0], 'solimp': [0.9, 0.95, 0.001, 0.5, 2.0]}
_ACT0 = {'gear': 1.0, 'ctrllimited': False, 'ctrlrange': [0, 0], 'forcelimited': False, 'forcerange': [0, 0],
         'gain': 1.0, 'bias': [0.0, 0.0, 0.0]}
_OPTION0 = {'timestep': 0.002, 'gravity': [0, 0, -9.81], 'tolerance': 1e-8, 'iterations': 100, 'impratio': 1.0}

_VEC = {'pos', 'axis', 'range', 'quat', 'size', 'fromto', 'friction', 'solref', 'solimp', 'solreflimit',
        'solimplimit', 'ctrlrange', 'forcerange', 'gravity'}
_INT = {'condim', 'contype', 'conaffinity', 'iterations'}
_TEXT = {'type', 'name', 'joint', 'limited'}
_BOOL = {'ctrllimited', 'forcelimited'}


def _value(key, text):
    if key in _VEC:
        return [float(t) for t in text.split()]
    if key in _INT:
        return int(text)
    if key in _BOOL:
        return text.strip().lower() == 'true'
    if key in _TEXT:
        return text.strip()
    return float(text)


def _take(attrib, allowed):
    return {k: _value(k, v) for k, v in attrib.items() if k in allowed}


def _actuator_update(rec, tag, attrib):
    """The XML shortcuts all write the ONE general-actuator record (also inside <default>, where a
    later shortcut overrides what an earlier one implied): motor = unit fixed gain, no bias;
    velocity = gain kv, bias (0, 0, -kv); position = gain kp, bias (0, -kp, 0); general sets only
    what it spells out."""
    rec.update(_take(attrib, {'gear', 'ctrllimited', 'ctrlrange', 'forcelimited', 'forcerange', 'joint', 'name'}))
    if isinstance(rec.get('gear'), list):
        rec['gear'] = rec['gear'][0]
    if 'gear' in attrib:
        rec['gear'] = float(attrib['gear'].split()[0])
    if tag == 'motor':
        rec['gain'], rec['bias'] = 1.0, [0.0, 0.0, 0.0]
    elif tag == 'velocity':
        if 'kv' in attrib:
            rec['gain'] = float(attrib['kv'])
        rec['bias'] = [0.0, 0.0, -rec['gain']]
    elif tag == 'position':
        if 'kp' in attrib:
            rec['gain'] = float(attrib['kp'])
        rec['bias'] = [0.0, -rec['gain'], 0.0]
    elif tag == 'general':
        if 'gainprm' in attrib:
            rec['gain'] = float(attrib['gainprm'].split()[0])
        if 'biasprm' in attrib:
            b = [float(t) for t in attrib['biasprm'].split()] + [0.0, 0.0, 0.0]
            rec['bias'] = b[:3]
    else:
        raise NotImplementedError(f'<{tag}> actuator')


def read_xml(text: str) -> dict:
    """MJCF subset: one top-level <default>, local coordinates, bodies with hinge / slide / free joints
    and plane / sphere / capsule / cylinder / box geoms, motor / velocity / position / general actuators.
    Sites, cameras, sensors, assets and visual attributes do not reach mjModel's dynamics and are skipped."""
    root = ET.fromstring(text)
    comp = root.find('compiler')
    angle = 'degree'
    if comp is not None:
        angle = comp.attrib.get('angle', 'degree')
        if comp.attrib.get('coordinate', 'local') != 'local':
            raise NotImplementedError('global coordinates')
        if comp.attrib.get('inertiafromgeom', 'auto') not in ('auto', 'true'):
            raise NotImplementedError('inertiafromgeom="false"')
    option = dict(_OPTION0)
    for el in root.findall('option'):
        option.update(_take(el.attrib, set(_OPTION0)))
    jdef, gdef, adef = dict(_JOINT0), dict(_GEOM0), dict(_ACT0)
    dflt = root.find('default')
    if dflt is not None:
        for el in dflt:
            if el.tag == 'default':
                raise NotImplementedError('nested default classes')
            if el.tag == 'joint':
                jdef.update(_take(el.attrib, set(_JOINT0)))
            elif el.tag == 'geom':
                gdef.update(_take(el.attrib, set(_GEOM0)))
            elif el.tag in ('motor', 'velocity', 'position', 'general'):
                _actuator_update(adef, el.tag, el.attrib)
    bodies = []

    def geom_of(el):
        g = dict(gdef)
        g.update(_take(el.attrib, set(_GEOM0) | {'name'}))
        g.setdefault('name', None)
        return g

    def visit(el, parent):
        me = len(bodies)
        rec = {'name': el.attrib.get('name'), 'parent': parent,
               'pos': _value('pos', el.attrib.get('pos', '0 0 0')), 'quat': _value('quat', el.attrib.get('quat', '1 0 0 0')),
               'joints': [], 'geoms': []}
        bodies.append(rec)
        for ch in el:
            if ch.tag == 'inertial':
                raise NotImplementedError('explicit <inertial>')
            if ch.tag == 'freejoint':
                # a <freejoint> takes NO defaults: no damping / armature / stiffness, never limited
                rec['joints'].append(dict(_JOINT0, type='free', name=ch.attrib.get('name'), limited='false'))
            elif ch.tag == 'joint':
                j = dict(jdef)
                j.update(_take(ch.attrib, set(_JOINT0) | {'name'}))
                j.setdefault('name', None)
                if j['type'] == 'free':
                    raise NotImplementedError('<joint type="free"> (with defaults); the robot files use <freejoint>')
                rec['joints'].append(j)
            elif ch.tag == 'geom':
                rec['geoms'].append(geom_of(ch))
        for ch in el:              # children after the body's own elements: depth-first, document order
            if ch.tag == 'body':
                visit(ch, me)

    world = root.find('worldbody')
    bodies.append({'name': 'world', 'parent': -1, 'pos': [0.0, 0.0, 0.0], 'quat': [1.0, 0.0, 0.0, 0.0], 'joints': [],
                   'geoms': [geom_of(g) for g in world if g.tag == 'geom']})
    for ch in world:
        if ch.tag == 'body':
            visit(ch, 0)
    actuators = []
    sect = root.find('actuator')
    if sect is not None:
        for el in sect:
            rec = dict(adef)
            rec.setdefault('name', None)
            _actuator_update(rec, el.tag, el.attrib)
            actuators.append(rec)
    return {'option': option, 'angle': angle, 'joint_default': jdef, 'geom_default': gdef, 'bodies': bodies,
            'actuators': actuators}


def with_world(spec: dict, goal_size: float, hazards_size: float, hazards_num: int) -> dict:
    """Engine.build_world_config + World.build (engine.py:335-384, world.py:310-326): one top-level body per geom,
    `<joint type="slide" axis="1 0 0" name="{name}_x" damping="1" limited="false"/>` and the same along y, a geom
    with contype = conaffinity = 0; goal: sphere r = goal_size at z 0; hazard i: cylinder (hazards_size, 1e-2) at
    z 2e-2 (rot = random_rot() = 0, engine.py:330-333).  The elements name no class, so every attribute they do
    not spell out comes from the robot file's <default> (that is how Doggo's stiffness reaches them)."""
    out = json.loads(json.dumps(spec))
    jd, gd = out['joint_default'], out['geom_default']

    def add(name, kind, size, z):
        joints = [dict(jd, type='slide', axis=list(ax), name=f'{name}_{c}', damping=1.0, limited='false')
                  for c, ax in (('x', (1.0, 0.0, 0.0)), ('y', (0.0, 1.0, 0.0)))]
        geom = dict(gd, name=name, type=kind, size=[float(v) for v in size], contype=0, conaffinity=0,
                    pos=[0.0, 0.0, 0.0], quat=[1.0, 0.0, 0.0, 0.0], fromto=None)
        out['bodies'].append({'name': name, 'parent': 0, 'pos': [0.0, 0.0, float(z)], 'quat': [1.0, 0.0, 0.0, 0.0],
                              'joints': joints, 'geoms': [geom]})

    add('goal', 'sphere', [goal_size], 0.0)
    for h in range(int(hazards_num)):
        add(f'hazard{h}', 'cylinder', [hazards_size, 1e-2], 2e-2)
    return out


# ----------------------------------------------------------------------------- rotations
def _rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _qmul(p, q):
    pw, pv, qw, qv = p[0], np.asarray(p[1:]), q[0], np.asarray(q[1:])
    return np.concatenate([[pw * qw - pv @ qv], pw * qv + qw * pv + np.cross(pv, qv)])


def _quat_of_axis_angle(axis, angle):
    return np.concatenate([[math.cos(angle / 2)], np.asarray(axis, float) * math.sin(angle / 2)])


def _quat_z_to(v):
    """unit quaternion taking +z onto the direction of v (shortest arc; x axis for the antiparallel case)"""
    v = np.asarray(v, float) / np.linalg.norm(v)
    c = v[2]
    if c > 1 - 1e-12:
        return np.array([1.0, 0.0, 0.0, 0.0])
    if c < -1 + 1e-12:
        return np.array([0.0, 1.0, 0.0, 0.0])
    ax = np.array([-v[1], v[0], 0.0])
    ax /= np.linalg.norm(ax)
    return _quat_of_axis_angle(ax, math.acos(c))


def _quat_of_rot(R):
    """rotation matrix -> unit quaternion via the eigenvector of the symmetric 4x4 K matrix (Bar-Itzhack)"""
    K = np.array([[R[0, 0] - R[1, 1] - R[2, 2], R[1, 0] + R[0, 1], R[2, 0] + R[0, 2], R[2, 1] - R[1, 2]],
                  [R[1, 0] + R[0, 1], R[1, 1] - R[0, 0] - R[2, 2], R[2, 1] + R[1, 2], R[0, 2] - R[2, 0]],
                  [R[2, 0] + R[0, 2], R[2, 1] + R[1, 2], R[2, 2] - R[0, 0] - R[1, 1], R[1, 0] - R[0, 1]],
                  [R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1], R[0, 0] + R[1, 1] + R[2, 2]]]) / 3.0
    w, V = np.linalg.eigh(K)
    x, y, z, s = V[:, -1]
    q = np.array([s, x, y, z])
    return q if q[0] >= 0 else -q


# ----------------------------------------------------------------------------- primitives
def primitive_mass_inertia(kind, size, density):
    """(mass, principal moments about the centre, in the geom frame; symmetry axis = z)."""
    if kind == PLANE:
        return 0.0, np.zeros(3)
    if kind == SPHERE:
        r = size[0]
        m = density * (4.0 / 3.0) * math.pi * r ** 3
        return m, np.full(3, 2.0 / 5.0 * m * r * r)
    if kind == BOX:
        hx, hy, hz = size[:3]
        m = density * (2 * hx) * (2 * hy) * (2 * hz)
        return m, m / 3.0 * np.array([hy * hy + hz * hz, hx * hx + hz * hz, hx * hx + hy * hy])
    if kind == CYLINDER:
        r, L = size[0], 2.0 * size[1]
        m = density * math.pi * r * r * L
        lateral = m * (r * r / 4.0 + L * L / 12.0)
        return m, np.array([lateral, lateral, m * r * r / 2.0])
    if kind == CAPSULE:
        # a cylinder of length L plus two hemispheres: a hemisphere of mass mh has its centre of mass 3r/8 from
        # its flat face and moments (83/320) mh r^2 (lateral, about its own COM) and (2/5) mh r^2 (axial)
        r, L = size[0], 2.0 * size[1]
        m_cyl = density * math.pi * r * r * L
        m_hemi = density * (2.0 / 3.0) * math.pi * r ** 3
        off = L / 2.0 + 3.0 * r / 8.0
        lateral = m_cyl * (r * r / 4.0 + L * L / 12.0) + 2.0 * (83.0 / 320.0 * m_hemi * r * r + m_hemi * off * off)
        axial = m_cyl * r * r / 2.0 + 2.0 * (2.0 / 5.0 * m_hemi * r * r)
        return m_cyl + 2.0 * m_hemi, np.array([lateral, lateral, axial])
    raise NotImplementedError(kind)


# ----------------------------------------------------------------------------- compile
class OModel(SimpleNamespace):
    def body_id(self, name):
        return self.body_name.index(name)

    def jnt_id(self, name):
        return self.jnt_name.index(name)


def build(spec: dict) -> OModel:
    rad = math.pi / 180.0 if spec['angle'] == 'degree' else 1.0
    B = spec['bodies']
    nb = len(B)
    m = OModel()
    m.opt = dict(spec['option'])
    m.opt['gravity'] = [float(v) for v in m.opt['gravity']]
    m.nbody = nb
    m.body_name = [b['name'] for b in B]
    m.body_parentid = np.array([b['parent'] for b in B], np.int32)
    m.body_pos = np.array([b['pos'] for b in B], float)
    m.body_quat = np.array([np.asarray(b['quat'], float) / np.linalg.norm(b['quat']) for b in B])
    # ---- joints / dofs ---------------------------------------------------------------------------
    J = []            # (body, record, qposadr, dofadr, nq, nv)
    nq = nv = 0
    m.body_jntadr = np.full(nb, -1, np.int32); m.body_jntnum = np.zeros(nb, np.int32)
    m.body_dofadr = np.full(nb, -1, np.int32); m.body_dofnum = np.zeros(nb, np.int32)
    for b, body in enumerate(B):
        if body['joints']:
            m.body_jntadr[b], m.body_dofadr[b] = len(J), nv
        for j in body['joints']:
            kind = JOINT_KIND[j['type']]
            if kind == BALL:
                raise NotImplementedError('ball joint')
            if kind == FREE and (len(body['joints']) != 1 or body['parent'] != 0):
                raise NotImplementedError('free joint below the top level / beside other joints')
            w = (7, 6) if kind == FREE else (1, 1)
            J.append((b, j, nq, nv, *w))
            nq += w[0]; nv += w[1]
        m.body_jntnum[b] = len(body['joints'])
        m.body_dofnum[b] = sum(6 if JOINT_KIND[j['type']] == FREE else 1 for j in body['joints'])
    nj = len(J)
    m.nq, m.nv, m.njnt = nq, nv, nj
    m.jnt_name = [j['name'] for _, j, *_ in J]
    m.jnt_type = np.array([JOINT_KIND[j['type']] for _, j, *_ in J], np.int32)
    m.jnt_bodyid = np.array([b for b, *_ in J], np.int32)
    m.jnt_qposadr = np.array([a for _, _, a, *_ in J], np.int32)
    m.jnt_dofadr = np.array([a for _, _, _, a, *_ in J], np.int32)
    m.jnt_pos = np.zeros((nj, 3)); m.jnt_axis = np.zeros((nj, 3)); m.jnt_range = np.zeros((nj, 2))
    m.jnt_limited = np.zeros(nj, np.int32); m.jnt_stiffness = np.zeros(nj); m.jnt_margin = np.zeros(nj)
    m.jnt_solref = np.zeros((nj, 2)); m.jnt_solimp = np.zeros((nj, 5))
    m.qpos0 = np.zeros(nq); m.qpos_spring = np.zeros(nq)
    m.dof_bodyid = np.zeros(nv, np.int32); m.dof_jntid = np.zeros(nv, np.int32)
    m.dof_damping = np.zeros(nv); m.dof_armature = np.zeros(nv)
    for k, (b, j, qa, da, wq, wv) in enumerate(J):
        kind = m.jnt_type[k]
        m.dof_bodyid[da:da + wv] = b
        m.dof_jntid[da:da + wv] = k
        m.jnt_solref[k], m.jnt_solimp[k] = j['solreflimit'], j['solimplimit']
        if kind == FREE:
            m.jnt_axis[k] = (0.0, 0.0, 1.0)
            m.qpos0[qa:qa + 3] = m.body_pos[b]
            m.qpos0[qa + 3:qa + 7] = m.body_quat[b]
            m.qpos_spring[qa:qa + 7] = m.qpos0[qa:qa + 7]
            continue
        unit = rad if kind == HINGE else 1.0      # angles of hinge joints are written in the compiler's angle unit
        ax = np.asarray(j['axis'], float)
        m.jnt_axis[k] = ax / np.linalg.norm(ax)
        m.jnt_pos[k] = j['pos']
        m.jnt_range[k] = np.asarray(j['range'], float) * unit
        lim = j['limited']
        m.jnt_limited[k] = int((m.jnt_range[k] != 0).any()) if lim == 'auto' else int(lim == 'true')
        m.jnt_stiffness[k], m.jnt_margin[k] = j['stiffness'], j['margin']
        m.qpos0[qa], m.qpos_spring[qa] = j['ref'] * unit, j['springref'] * unit
        m.dof_damping[da], m.dof_armature[da] = j['damping'], j['armature']
    # parent dof: the previous dof on the same body, else the last dof of the closest moving ancestor
    m.dof_parentid = np.full(nv, -1, np.int32)
    last_dof_of = {}
    for b in range(nb):
        anc = m.body_parentid[b]
        while anc > 0 and m.body_dofnum[anc] == 0:
            anc = m.body_parentid[anc]
        prev = last_dof_of.get(anc, -1) if anc > 0 else -1
        for d in range(m.body_dofnum[b]):
            i = m.body_dofadr[b] + d
            m.dof_parentid[i] = prev
            prev = i
        if m.body_dofnum[b]:
            last_dof_of[b] = prev
    # ---- geoms -------------------------------------------------------------------------------------
    G = [(b, g) for b, body in enumerate(B) for g in body['geoms']]
    ng = len(G)
    m.ngeom = ng
    m.geom_name = [g.get('name') for _, g in G]
    m.geom_bodyid = np.array([b for b, _ in G], np.int32)
    m.geom_type = np.array([GEOM_KIND[g['type']] for _, g in G], np.int32)
    m.geom_pos = np.zeros((ng, 3)); m.geom_quat = np.zeros((ng, 4)); m.geom_size = np.zeros((ng, 3))
    m.geom_friction = np.zeros((ng, 3)); m.geom_margin = np.zeros(ng); m.geom_gap = np.zeros(ng)
    m.geom_condim = np.zeros(ng, np.int32); m.geom_contype = np.zeros(ng, np.int32); m.geom_conaffinity = np.zeros(ng, np.int32)
    m.geom_solref = np.zeros((ng, 2)); m.geom_solimp = np.zeros((ng, 5))
    gmass = np.zeros(ng); ginertia = np.zeros((ng, 3))
    for k, (b, g) in enumerate(G):
        kind = m.geom_type[k]
        size = (list(g['size']) + [0.0, 0.0, 0.0])[:3]
        pos, quat = np.asarray(g['pos'], float), np.asarray(g['quat'], float)
        if g.get('fromto') is not None:
            a, e = np.asarray(g['fromto'][:3], float), np.asarray(g['fromto'][3:], float)
            if kind not in (CAPSULE, CYLINDER, BOX):
                raise NotImplementedError('fromto on this geom type')
            pos = (a + e) / 2.0
            quat = _quat_z_to(a - e)            # user_objects.cc: the z axis points from `to` back to `from`
            size[2 if kind == BOX else 1] = np.linalg.norm(e - a) / 2.0
        m.geom_pos[k], m.geom_quat[k], m.geom_size[k] = pos, quat / np.linalg.norm(quat), size
        fr = list(g['friction']) + _GEOM0['friction'][len(g['friction']):]
        m.geom_friction[k] = fr[:3]
        m.geom_margin[k], m.geom_gap[k] = g['margin'], g['gap']
        m.geom_condim[k], m.geom_contype[k], m.geom_conaffinity[k] = g['condim'], g['contype'], g['conaffinity']
        m.geom_solref[k], m.geom_solimp[k] = g['solref'], g['solimp']
        gmass[k], ginertia[k] = primitive_mass_inertia(kind, size, g['density'])
    # ---- inertial frames from the geoms (inertiafromgeom) ------------------------------------------
    m.body_mass = np.zeros(nb); m.body_ipos = np.zeros((nb, 3)); m.body_inertia_full = np.zeros((nb, 3, 3))
    m.body_inertia = np.zeros((nb, 3)); m.body_iquat = np.tile([1.0, 0.0, 0.0, 0.0], (nb, 1))
    for b in range(1, nb):
        ks = np.flatnonzero(m.geom_bodyid == b)
        mass = gmass[ks].sum()
        if mass <= 0:
            raise ValueError(f'massless body {m.body_name[b]}')
        com = (gmass[ks, None] * m.geom_pos[ks]).sum(0) / mass
        I = np.zeros((3, 3))
        for k in ks:
            Rg = _rot(m.geom_quat[k])
            r = m.geom_pos[k] - com
            I += Rg @ np.diag(ginertia[k]) @ Rg.T                     # own inertia, rotated into the body frame
            I += gmass[k] * ((r @ r) * np.eye(3) - np.outer(r, r))    # parallel axes
        m.body_mass[b], m.body_ipos[b], m.body_inertia_full[b] = mass, com, I
        ev, evec = np.linalg.eigh(I)
        ev, evec = ev[::-1], evec[:, ::-1]                             # principal moments, largest first (mju_eig3)
        if np.linalg.det(evec) < 0:
            evec[:, 2] *= -1
        m.body_inertia[b], m.body_iquat[b] = ev, _quat_of_rot(evec)
    # ---- actuators ---------------------------------------------------------------------------------
    A = spec['actuators']
    m.nu = len(A)
    m.actuator_trnid = np.array([m.jnt_name.index(a['joint']) for a in A], np.int32)
    m.actuator_gear = np.array([a['gear'] for a in A], float)
    m.actuator_ctrllimited = np.array([int(a['ctrllimited']) for a in A], np.int32)
    m.actuator_ctrlrange = np.array([a['ctrlrange'] for a in A], float).reshape(m.nu, 2)
    m.actuator_forcelimited = np.array([int(a['forcelimited']) for a in A], np.int32)
    m.actuator_forcerange = np.array([a['forcerange'] for a in A], float).reshape(m.nu, 2)
    m.actuator_gainprm0 = np.array([a['gain'] for a in A], float)
    m.actuator_biasprm = np.array([a['bias'] for a in A], float).reshape(m.nu, 3)
    # ---- broad-phase pair filter (engine_collision_driver.c: contype/conaffinity, same weld body, parent-child) ----
    m.body_weldid = np.zeros(nb, np.int32)
    for b in range(1, nb):
        m.body_weldid[b] = b if m.body_jntnum[b] else m.body_weldid[m.body_parentid[b]]
    pairs = []
    for g1 in range(ng):
        for g2 in range(g1 + 1, ng):
            w1, w2 = m.body_weldid[m.geom_bodyid[g1]], m.body_weldid[m.geom_bodyid[g2]]
            compatible = (m.geom_contype[g1] & m.geom_conaffinity[g2]) or (m.geom_contype[g2] & m.geom_conaffinity[g1])
            if w1 == w2 or not compatible:
                continue
            up1 = m.body_weldid[m.body_parentid[w1]] if w1 else -1
            up2 = m.body_weldid[m.body_parentid[w2]] if w2 else -1
            if (w1 and w1 == up2) or (w2 and w2 == up1):     # parent-child pairs are skipped unless the parent is the world
                continue
            pairs.append((g1, g2))
    m.contact_pairs = pairs
    _constants_at_qpos0(m)
    return m


# ----------------------------------------------------------------------------- qpos0 constants
def pose_at_qpos0(m: OModel):
    """world frames of bodies, inertial frames and joints at qpos0 (all joint displacements are zero there)"""
    xpos = np.zeros((m.nbody, 3)); xrot = np.tile(np.eye(3), (m.nbody, 1, 1))
    for b in range(1, m.nbody):
        p = m.body_parentid[b]
        xpos[b] = xpos[p] + xrot[p] @ m.body_pos[b]
        xrot[b] = xrot[p] @ _rot(m.body_quat[b])
    xipos = xpos + np.einsum('bij,bj->bi', xrot, m.body_ipos)
    anchor = np.array([xpos[m.jnt_bodyid[j]] + xrot[m.jnt_bodyid[j]] @ m.jnt_pos[j] for j in range(m.njnt)]).reshape(m.njnt, 3)
    axis = np.array([xrot[m.jnt_bodyid[j]] @ m.jnt_axis[j] for j in range(m.njnt)]).reshape(m.njnt, 3)
    return xpos, xrot, xipos, anchor, axis


def unit_velocity_fields(m: OModel):
    """For every dof i: the angular velocity w[i, b] of body b and the linear velocity v[i, b] of its centre of mass
    when qvel = e_i, by propagating twists down the tree (free joint: linear dofs along world axes, angular dofs
    about the body's own axes)."""
    xpos, xrot, xipos, anchor, axis = pose_at_qpos0(m)
    w = np.zeros((m.nv, m.nbody, 3)); v = np.zeros((m.nv, m.nbody, 3))
    for b in range(1, m.nbody):
        p = m.body_parentid[b]
        w[:, b] = w[:, p]
        v[:, b] = v[:, p] + np.cross(w[:, p], xipos[b] - xipos[p])     # rigid transport of the parent's motion
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            d = m.jnt_dofadr[j]
            if m.jnt_type[j] == FREE:
                for c in range(3):
                    v[d + c, b, c] += 1.0
                    w[d + 3 + c, b] += xrot[b][:, c]
                    v[d + 3 + c, b] += np.cross(xrot[b][:, c], xipos[b] - xpos[b])
            elif m.jnt_type[j] == SLIDE:
                v[d, b] += axis[j]
            else:
                w[d, b] += axis[j]
                v[d, b] += np.cross(axis[j], xipos[b] - anchor[j])
    return w, v, xrot


def joint_space_inertia(m: OModel):
    """M_ij = sum_b  m_b v_i.v_j + w_i . I_b w_j  (twice the kinetic energy's bilinear form) + armature."""
    w, v, xrot = unit_velocity_fields(m)
    M = np.diag(m.dof_armature.astype(float))
    for b in range(1, m.nbody):
        Iw = xrot[b] @ m.body_inertia_full[b] @ xrot[b].T
        M += m.body_mass[b] * v[:, b] @ v[:, b].T + w[:, b] @ Iw @ w[:, b].T
    return M, w, v


def _constants_at_qpos0(m: OModel):
    """engine_setconst.c (set0): stat.meaninertia = mean diagonal of M; dof_invweight0 = diag(M^-1), averaged over the
    three translational and the three rotational dofs of a free joint; body_invweight0 = mean diagonal of the
    translational / rotational 3x3 blocks of J M^-1 J^T at the body's centre of mass."""
    m.body_invweight0 = np.zeros((m.nbody, 2)); m.dof_invweight0 = np.zeros(m.nv); m.stat_meaninertia = 1.0
    if m.nv == 0:
        return
    M, w, v = joint_space_inertia(m)
    Minv = np.linalg.solve(M, np.eye(m.nv))
    m.stat_meaninertia = float(np.mean(np.diag(M)))
    m.dof_invweight0 = np.diag(Minv).copy()
    for j in np.flatnonzero(m.jnt_type == FREE):
        d = m.jnt_dofadr[j]
        m.dof_invweight0[d:d + 3] = m.dof_invweight0[d:d + 3].mean()
        m.dof_invweight0[d + 3:d + 6] = m.dof_invweight0[d + 3:d + 6].mean()
    for b in range(1, m.nbody):
        Jv, Jw = v[:, b].T, w[:, b].T          # 3 x nv
        m.body_invweight0[b, 0] = max(MINVAL, np.trace(Jv @ Minv @ Jv.T) / 3.0)
        m.body_invweight0[b, 1] = max(MINVAL, np.trace(Jw @ Minv @ Jw.T) / 3.0)


# ----------------------------------------------------------------------------- fixtures
_SPECS = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden', 'oracle_robot_specs.json')
_cache = {}


def robot_spec(robot_base: str) -> dict:
    """the read_xml() output for one of the reference's robot files, from the committed fixture"""
    if 'specs' not in _cache:
        with open(_SPECS) as f:
            _cache['specs'] = json.load(f)
    return _cache['specs'][robot_base]


def world_model(robot_base: str, goal_size: float, hazards_size: float, hazards_num: int) -> OModel:
    key = (robot_base, float(goal_size), float(hazards_size), int(hazards_num))
    if key not in _cache:
        _cache[key] = build(with_world(robot_spec(robot_base), goal_size, hazards_size, hazards_num))
    return _cache[key]
