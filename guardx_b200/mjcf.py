"""Model compiler: the part of ``mujoco.MjModel.from_xml_string`` this path needs.

The reference builds its world as MJCF text (``World.build``, world.py:104-333: the
robot's base XML + one jointed body per goal/hazard geom, :310-326) and hands it to
the MuJoCo compiler (engine.py:228).  MuJoCo is a third-party dependency that is
absent from this image, so the constants it would produce are computed here, from the
published compiler rules (XML defaults, ``fromto`` frames, inertia from geoms, principal
frames, ``dof_invweight0`` / ``body_invweight0`` / ``stat.meaninertia`` at ``qpos0``).
All arithmetic is float64, as in the MuJoCo compiler; consumers cast to fp32 the way
``mjx.put_model`` does.

Two front ends produce the same intermediate ``Spec``:
  * ``parse_mjcf(text)``   -- an MJCF subset parser (one top-level <default> class,
                              hinge/slide/free joints, sphere/capsule/box/cylinder/plane
                              geoms, motor/velocity/general actuators);
  * ``guardx_b200.robots`` -- the built-in robot descriptions (no XML needed at run time).
``add_world_geoms`` appends the goal/hazard bodies the way ``World.build`` does.
``compile_spec`` turns a ``Spec`` into a ``Model`` of numpy arrays named like ``mjModel``.

Scope: kinematic trees of hinge / slide joints, optionally under one free joint (Doggo's root).
The CUDA kernels of this round cover the planar-root robots (Point, Swimmer, Ant, Walker); the
free-joint path is compiled and restated in the oracle (oracle/mjx_numpy.py) as groundwork.
"""
from __future__ import annotations

import copy
import math
import xml.etree.ElementTree as ET
from dataclasses import dataclass, field

import numpy as np

mjMINVAL = 1e-15

JNT_FREE, JNT_BALL, JNT_SLIDE, JNT_HINGE = 0, 1, 2, 3
GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE, GEOM_CYLINDER, GEOM_BOX = 0, 2, 3, 5, 6
_GEOM_TYPES = {'plane': GEOM_PLANE, 'sphere': GEOM_SPHERE, 'capsule': GEOM_CAPSULE,
               'cylinder': GEOM_CYLINDER, 'box': GEOM_BOX}
_JNT_TYPES = {'free': JNT_FREE, 'ball': JNT_BALL, 'slide': JNT_SLIDE, 'hinge': JNT_HINGE}

# MuJoCo's built-in defaults (mjmodel / user objects)
JOINT_DEFAULT = dict(type='hinge', pos=(0., 0., 0.), axis=(0., 0., 1.), range=(0., 0.), limited=None,
                     damping=0., armature=0., stiffness=0., springref=0., ref=0., margin=0.,
                     solreflimit=(0.02, 1.), solimplimit=(0.9, 0.95, 0.001, 0.5, 2.), frictionloss=0., name=None)
GEOM_DEFAULT = dict(type='sphere', pos=(0., 0., 0.), quat=(1., 0., 0., 0.), size=(0., 0., 0.), fromto=None,
                    density=1000., friction=(1., 0.005, 0.0001), margin=0., gap=0., condim=3, contype=1,
                    conaffinity=1, solref=(0.02, 1.), solimp=(0.9, 0.95, 0.001, 0.5, 2.), solmix=1.,
                    priority=0, name=None)
ACT_DEFAULT = dict(joint=None, gear=1., ctrllimited=False, ctrlrange=(0., 0.), forcelimited=False,
                   forcerange=(0., 0.), gainprm0=1., biasprm=(0., 0., 0.), name=None)
OPT_DEFAULT = dict(timestep=0.002, gravity=(0., 0., -9.81), tolerance=1e-8, iterations=100, ls_iterations=50,
                   ls_tolerance=0.01, impratio=1., solver='newton', cone='pyramidal', integrator='euler')


@dataclass
class Spec:
    """Parsed, defaults-resolved model description (bodies in depth-first document order)."""
    opt: dict = field(default_factory=lambda: dict(OPT_DEFAULT))
    angle: str = 'degree'                       # MJCF compiler default
    defaults: dict = field(default_factory=lambda: {'joint': {}, 'geom': {}, 'actuator': {}})
    bodies: list = field(default_factory=list)  # dict(name, parent, pos, quat, joints[], geoms[]); bodies[0] = world
    actuators: list = field(default_factory=list)


def _floats(s):
    return tuple(float(x) for x in s.split())


def _bool(s):
    return {'true': True, 'false': False, 'auto': None}[s.strip().lower()]


_JOINT_PARSE = dict(pos=_floats, axis=_floats, range=_floats, limited=_bool, damping=float, armature=float,
                    stiffness=float, springref=float, ref=float, margin=float, solreflimit=_floats,
                    solimplimit=_floats, frictionloss=float, type=str, name=str)
_GEOM_PARSE = dict(pos=_floats, quat=_floats, size=_floats, fromto=_floats, density=float, friction=_floats,
                   margin=float, gap=float, condim=int, contype=int, conaffinity=int, solref=_floats,
                   solimp=_floats, solmix=float, priority=int, type=str, name=str)
_ACT_PARSE = dict(joint=str, gear=lambda s: _floats(s)[0], ctrllimited=_bool, ctrlrange=_floats,
                  forcelimited=_bool, forcerange=_floats, name=str)


def _apply(dst: dict, attrib: dict, table: dict):
    for k, v in attrib.items():
        if k in table:
            dst[k] = table[k](v)


def _apply_actuator(dst: dict, tag: str, attrib: dict):
    """XML actuator shortcuts write into ONE actuator record (also in <default>, where a
    later shortcut overrides the implied fields of an earlier one)."""
    _apply(dst, attrib, _ACT_PARSE)
    if tag == 'motor':
        dst['gainprm0'] = 1.0
        dst['biasprm'] = (0., 0., 0.)
    elif tag == 'velocity':
        if 'kv' in attrib:
            dst['gainprm0'] = float(attrib['kv'])
        dst['biasprm'] = (0., 0., -dst.get('gainprm0', 1.0))
    elif tag == 'position':
        if 'kp' in attrib:
            dst['gainprm0'] = float(attrib['kp'])
        dst['biasprm'] = (0., -dst.get('gainprm0', 1.0), 0.)
    elif tag == 'general':
        if 'gainprm' in attrib:
            dst['gainprm0'] = _floats(attrib['gainprm'])[0]
        if 'biasprm' in attrib:
            b = _floats(attrib['biasprm']) + (0., 0., 0.)
            dst['biasprm'] = b[:3]
    else:
        raise NotImplementedError(f'actuator <{tag}>')


def parse_mjcf(text: str) -> Spec:
    root = ET.fromstring(text)
    spec = Spec()
    comp = root.find('compiler')
    if comp is not None:
        spec.angle = comp.attrib.get('angle', 'degree')
        if comp.attrib.get('coordinate', 'local') != 'local':
            raise NotImplementedError('compiler coordinate="global"')
    for opt in root.findall('option'):
        for k, v in opt.attrib.items():
            if k == 'gravity':
                spec.opt[k] = _floats(v)
            elif k in ('timestep', 'tolerance', 'ls_tolerance', 'impratio'):
                spec.opt[k] = float(v)
            elif k in ('iterations', 'ls_iterations'):
                spec.opt[k] = int(v)
            elif k in ('solver', 'cone', 'integrator'):
                spec.opt[k] = v.lower()
    dflt = root.find('default')
    if dflt is not None:
        for el in dflt:
            if el.tag == 'default':
                raise NotImplementedError('nested default classes')
            if el.tag == 'joint':
                _apply(spec.defaults['joint'], el.attrib, _JOINT_PARSE)
            elif el.tag == 'geom':
                _apply(spec.defaults['geom'], el.attrib, _GEOM_PARSE)
            elif el.tag in ('motor', 'velocity', 'position', 'general'):
                _apply_actuator(spec.defaults['actuator'], el.tag, el.attrib)

    def walk(el, parent):
        idx = len(spec.bodies)
        body = dict(name=el.attrib.get('name'), parent=parent,
                    pos=_floats(el.attrib.get('pos', '0 0 0')), quat=_floats(el.attrib.get('quat', '1 0 0 0')),
                    joints=[], geoms=[])
        spec.bodies.append(body)
        for ch in el:
            if ch.tag == 'joint':
                j = {}
                _apply(j, ch.attrib, _JOINT_PARSE)
                body['joints'].append(j)
            elif ch.tag == 'freejoint':
                body['joints'].append({'type': 'free', 'name': ch.attrib.get('name')})
            elif ch.tag == 'geom':
                g = {}
                _apply(g, ch.attrib, _GEOM_PARSE)
                body['geoms'].append(g)
            elif ch.tag == 'inertial':
                raise NotImplementedError('<inertial>')
        for ch in el:
            if ch.tag == 'body':
                walk(ch, idx)

    wb = root.find('worldbody')
    world = dict(name='world', parent=-1, pos=(0., 0., 0.), quat=(1., 0., 0., 0.), joints=[], geoms=[])
    spec.bodies.append(world)
    for ch in wb:
        if ch.tag == 'geom':
            g = {}
            _apply(g, ch.attrib, _GEOM_PARSE)
            world['geoms'].append(g)
    for ch in wb:
        if ch.tag == 'body':
            walk(ch, 0)
    act = root.find('actuator')
    if act is not None:
        for el in act:
            a = {'_tag': el.tag, '_attrib': dict(el.attrib)}
            spec.actuators.append(a)
    return spec


def add_world_geoms(spec: Spec, goal_size: float, hazards_size: float, hazards_num: int) -> Spec:
    """What ``World.build`` appends for ``build_world_config`` (engine.py:335-384, world.py:310-326):
    one body per geom with two slide joints (damping 1, limited false), geom contype =
    conaffinity = 0; goal = sphere at z 0, hazards = cylinders (r, 1e-2) at z 2e-2.  The new
    elements carry no class, so they inherit the robot XML's top-level defaults."""
    spec = copy.deepcopy(spec)

    def body(name, gtype, size, z):
        return dict(name=name, parent=0, pos=(0., 0., z), quat=(1., 0., 0., 0.),
                    joints=[dict(type='slide', axis=(1., 0., 0.), name=name + '_x', damping=1., limited=False),
                            dict(type='slide', axis=(0., 1., 0.), name=name + '_y', damping=1., limited=False)],
                    geoms=[dict(name=name, type=gtype, size=size, contype=0, conaffinity=0)])
    spec.bodies.append(body('goal', 'sphere', (float(goal_size),), 0.0))
    for i in range(int(hazards_num)):
        spec.bodies.append(body(f'hazard{i}', 'cylinder', (float(hazards_size), 1e-2), 2e-2))
    return spec


# ------------------------------------------------------------------ small math
def quat_mul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3],
                     a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1],
                     a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]])


def quat2mat(q):
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def z2quat(vec):
    """Quaternion that rotates the z axis onto ``vec`` (mjuu_z2quat)."""
    v = np.asarray(vec, float)
    v = v / np.linalg.norm(v)
    axis = np.cross([0., 0., 1.], v)
    s = np.linalg.norm(axis)
    if s < 1e-10:
        axis = np.array([1., 0., 0.])
    else:
        axis = axis / s
    ang = math.atan2(s, v[2])
    return np.array([math.cos(ang / 2), *(axis * math.sin(ang / 2))])


def _geom_mass_inertia(gtype, size, density):
    """Mass and principal inertia (geom frame) of a uniform-density primitive."""
    if gtype == GEOM_SPHERE:
        r = size[0]
        m = density * 4.0 / 3.0 * math.pi * r ** 3
        i = 0.4 * m * r * r
        return m, np.array([i, i, i])
    if gtype == GEOM_CAPSULE:
        r, h = size[0], 2 * size[1]
        m = density * (math.pi * r * r * h + 4.0 / 3.0 * math.pi * r ** 3)
        ms = m * 4 * r / (4 * r + 3 * h)   # the two hemispheres
        mc = m - ms
        ix = mc * (3 * r * r + h * h) / 12 + ms * (0.4 * r * r + 0.375 * r * h + 0.25 * h * h)
        iz = mc * r * r / 2 + 0.4 * ms * r * r
        return m, np.array([ix, ix, iz])
    if gtype == GEOM_CYLINDER:
        r, h = size[0], 2 * size[1]
        m = density * math.pi * r * r * h
        return m, np.array([m * (3 * r * r + h * h) / 12, m * (3 * r * r + h * h) / 12, m * r * r / 2])
    if gtype == GEOM_BOX:
        a, b, c = size[:3]
        m = density * 8 * a * b * c
        return m, np.array([m * (b * b + c * c) / 3, m * (a * a + c * c) / 3, m * (a * a + b * b) / 3])
    if gtype == GEOM_PLANE:
        return 0.0, np.zeros(3)
    raise NotImplementedError(gtype)


@dataclass
class Model:
    """Compiled constants, named like mjModel.  Arrays are float64 / int32 numpy."""
    opt: dict
    nq: int
    nv: int
    nu: int
    nbody: int
    njnt: int
    ngeom: int
    body_name: list
    body_parentid: np.ndarray
    body_pos: np.ndarray
    body_quat: np.ndarray
    body_ipos: np.ndarray       # COM in the body frame
    body_inertia_full: np.ndarray  # [nbody,3,3] inertia about the COM in BODY axes (= R_i diag(body_inertia) R_i^T)
    body_inertia: np.ndarray    # principal moments (descending)
    body_iquat: np.ndarray
    body_mass: np.ndarray
    body_jntadr: np.ndarray
    body_jntnum: np.ndarray
    body_dofadr: np.ndarray
    body_dofnum: np.ndarray
    body_invweight0: np.ndarray  # [nbody, 2]
    jnt_name: list
    jnt_type: np.ndarray
    jnt_bodyid: np.ndarray
    jnt_pos: np.ndarray
    jnt_axis: np.ndarray
    jnt_qposadr: np.ndarray
    jnt_dofadr: np.ndarray
    jnt_limited: np.ndarray
    jnt_range: np.ndarray
    jnt_stiffness: np.ndarray
    jnt_margin: np.ndarray
    jnt_solref: np.ndarray
    jnt_solimp: np.ndarray
    qpos0: np.ndarray
    qpos_spring: np.ndarray
    dof_bodyid: np.ndarray
    dof_jntid: np.ndarray
    dof_parentid: np.ndarray
    dof_damping: np.ndarray
    dof_armature: np.ndarray
    dof_invweight0: np.ndarray
    geom_name: list
    geom_type: np.ndarray
    geom_bodyid: np.ndarray
    geom_pos: np.ndarray
    geom_quat: np.ndarray
    geom_size: np.ndarray
    geom_friction: np.ndarray
    geom_margin: np.ndarray
    geom_gap: np.ndarray
    geom_condim: np.ndarray
    geom_contype: np.ndarray
    geom_conaffinity: np.ndarray
    geom_solref: np.ndarray
    geom_solimp: np.ndarray
    actuator_trnid: np.ndarray   # joint id
    actuator_gear: np.ndarray
    actuator_ctrllimited: np.ndarray
    actuator_ctrlrange: np.ndarray
    actuator_forcelimited: np.ndarray
    actuator_forcerange: np.ndarray
    actuator_gainprm0: np.ndarray
    actuator_biasprm: np.ndarray  # [nu, 3]
    stat_meaninertia: float
    body_weldid: np.ndarray = None   # set by compile_spec
    contact_pairs: list = None          # (geom1, geom2) that pass the contype/conaffinity + body filters

    def body_id(self, name):
        return self.body_name.index(name)

    def jnt_id(self, name):
        return self.jnt_name.index(name)


def _resolve(defaults: dict, builtin: dict, given: dict) -> dict:
    out = dict(builtin)
    out.update(defaults)
    out.update(given)
    return out


def compile_spec(spec: Spec) -> Model:
    deg = math.pi / 180.0 if spec.angle == 'degree' else 1.0
    nbody = len(spec.bodies)
    # ---- bodies, joints, geoms ------------------------------------------------
    body_parent, body_pos, body_quat = [], [], []
    body_jntadr, body_jntnum, body_dofadr, body_dofnum = [], [], [], []
    J = dict(name=[], type=[], body=[], pos=[], axis=[], qposadr=[], dofadr=[], limited=[], range=[],
             stiffness=[], margin=[], solref=[], solimp=[], damping=[], armature=[], springref=[], ref=[])
    G = dict(name=[], type=[], body=[], pos=[], quat=[], size=[], friction=[], margin=[], gap=[], condim=[],
             contype=[], conaffinity=[], solref=[], solimp=[], mass=[], inertia=[])
    nq = nv = 0
    for b, body in enumerate(spec.bodies):
        body_parent.append(body['parent'])
        body_pos.append(np.asarray(body['pos'], float))
        q = np.asarray(body['quat'], float)
        body_quat.append(q / np.linalg.norm(q))
        body_jntadr.append(len(J['name']) if body['joints'] else -1)
        body_jntnum.append(len(body['joints']))
        body_dofadr.append(nv if body['joints'] else -1)
        dof0 = nv
        for jg in body['joints']:
            j = _resolve(spec.defaults['joint'], JOINT_DEFAULT, jg)
            t = _JNT_TYPES[j['type']]
            if t == JNT_BALL:
                raise NotImplementedError('ball joints are not compiled by this build')
            if t == JNT_FREE:
                # <freejoint>: no attributes, the defaults class is NOT applied (no damping, armature,
                # stiffness, never limited); 7 qpos (position + unit quaternion, qpos0 = the body frame),
                # 6 dofs (linear in world axes, angular in body axes)
                if len(body['joints']) != 1 or body['parent'] != 0:
                    raise NotImplementedError('a free joint must be the only joint of a top-level body')
                J['name'].append(jg.get('name')); J['type'].append(t); J['body'].append(b)
                J['pos'].append(np.zeros(3)); J['axis'].append(np.array([0., 0., 1.]))
                J['qposadr'].append(nq); J['dofadr'].append(nv)
                J['limited'].append(0); J['range'].append(np.zeros(2)); J['stiffness'].append(0.0); J['margin'].append(0.0)
                J['solref'].append(np.asarray(JOINT_DEFAULT['solreflimit'], float)); J['solimp'].append(np.asarray(JOINT_DEFAULT['solimplimit'], float))
                J['damping'].append(0.0); J['armature'].append(0.0); J['springref'].append(0.0); J['ref'].append(0.0)
                J.setdefault('nq', []).append(7); J.setdefault('nv', []).append(6)
                nq += 7; nv += 6
                continue
            axis = np.asarray(j['axis'], float)
            axis = axis / np.linalg.norm(axis)
            rng = np.asarray(j['range'], float) * (deg if t == JNT_HINGE else 1.0)
            limited = j['limited']
            if limited is None:   # autolimits: limited iff a range is given
                limited = bool(rng[0] != 0.0 or rng[1] != 0.0)
            scale = deg if t == JNT_HINGE else 1.0
            J['name'].append(j['name']); J['type'].append(t); J['body'].append(b)
            J['pos'].append(np.asarray(j['pos'], float)); J['axis'].append(axis)
            J['qposadr'].append(nq); J['dofadr'].append(nv)
            J['limited'].append(int(bool(limited))); J['range'].append(rng)
            J['stiffness'].append(j['stiffness']); J['margin'].append(j['margin'])
            J['solref'].append(np.asarray(j['solreflimit'], float)); J['solimp'].append(np.asarray(j['solimplimit'], float))
            J['damping'].append(j['damping']); J['armature'].append(j['armature'])
            J['springref'].append(j['springref'] * scale); J['ref'].append(j['ref'] * scale)
            J.setdefault('nq', []).append(1); J.setdefault('nv', []).append(1)
            nq += 1; nv += 1
        body_dofnum.append(nv - dof0)
        for gg in body['geoms']:
            g = _resolve(spec.defaults['geom'], GEOM_DEFAULT, gg)
            t = _GEOM_TYPES[g['type']]
            size = list(g['size']) + [0.0] * (3 - len(g['size']))
            pos, quat = np.asarray(g['pos'], float), np.asarray(g['quat'], float)
            if g.get('fromto') is not None:
                ft = np.asarray(g['fromto'], float)
                p0, p1 = ft[:3], ft[3:]
                pos = 0.5 * (p0 + p1)
                quat = z2quat(p1 - p0)   # geom z axis along from -> to
                half = 0.5 * np.linalg.norm(p1 - p0)
                if t in (GEOM_CAPSULE, GEOM_CYLINDER):
                    size[1] = half
                elif t == GEOM_BOX:
                    size[2] = half
            quat = quat / np.linalg.norm(quat)
            m, inert = _geom_mass_inertia(t, size, g['density'])
            fr = list(g['friction']) + list(GEOM_DEFAULT['friction'])[len(g['friction']):]
            G['name'].append(g['name']); G['type'].append(t); G['body'].append(b); G['pos'].append(pos)
            G['quat'].append(quat); G['size'].append(np.asarray(size[:3], float)); G['friction'].append(np.asarray(fr[:3], float))
            G['margin'].append(g['margin']); G['gap'].append(g['gap']); G['condim'].append(g['condim'])
            G['contype'].append(g['contype']); G['conaffinity'].append(g['conaffinity'])
            G['solref'].append(np.asarray(g['solref'], float)); G['solimp'].append(np.asarray(g['solimp'], float))
            G['mass'].append(m); G['inertia'].append(inert)
    njnt, ngeom = len(J['name']), len(G['name'])
    # ---- body inertial frames from geoms (inertiafromgeom) ----------------------
    body_mass = np.zeros(nbody); body_ipos = np.zeros((nbody, 3)); body_I = np.zeros((nbody, 3, 3))
    body_inertia = np.zeros((nbody, 3)); body_iquat = np.tile([1., 0, 0, 0], (nbody, 1))
    for b in range(1, nbody):
        gs = [k for k in range(ngeom) if G['body'][k] == b]
        m = sum(G['mass'][k] for k in gs)
        if m <= 0:
            raise ValueError(f'body {spec.bodies[b]["name"]} has no mass')
        com = sum(G['mass'][k] * G['pos'][k] for k in gs) / m
        I = np.zeros((3, 3))
        for k in gs:
            R = quat2mat(G['quat'][k])
            d = G['pos'][k] - com
            I += R @ np.diag(G['inertia'][k]) @ R.T + G['mass'][k] * (d @ d * np.eye(3) - np.outer(d, d))
        body_mass[b], body_ipos[b], body_I[b] = m, com, I
        w, V = np.linalg.eigh(I)
        order = np.argsort(-w)
        w, V = w[order], V[:, order]
        if np.linalg.det(V) < 0:
            V[:, 2] = -V[:, 2]
        body_inertia[b] = w
        body_iquat[b] = _mat2quat(V)
    # per-qpos / per-dof expansions (a free joint owns 7 qpos and 6 dofs)
    jnt_nq, jnt_nv = J.get('nq', []), J.get('nv', [])
    jnt_qpos0 = np.zeros(nq); qpos_spring = np.zeros(nq)
    dof_jnt, dof_damping, dof_armature = [], [], []
    for jid in range(njnt):
        a = J['qposadr'][jid]
        if J['type'][jid] == JNT_FREE:
            b_ = J['body'][jid]
            jnt_qpos0[a:a + 3] = body_pos[b_]; jnt_qpos0[a + 3:a + 7] = body_quat[b_]
            qpos_spring[a:a + 7] = jnt_qpos0[a:a + 7]
        else:
            jnt_qpos0[a] = J['ref'][jid]; qpos_spring[a] = J['springref'][jid]
        dof_jnt += [jid] * jnt_nv[jid]; dof_damping += [J['damping'][jid]] * jnt_nv[jid]; dof_armature += [J['armature'][jid]] * jnt_nv[jid]
    mdl = Model(
        opt=dict(spec.opt), nq=nq, nv=nv, nu=0, nbody=nbody, njnt=njnt, ngeom=ngeom,
        body_name=[b['name'] for b in spec.bodies], body_parentid=np.asarray(body_parent, np.int32),
        body_pos=np.asarray(body_pos), body_quat=np.asarray(body_quat), body_ipos=body_ipos,
        body_inertia_full=body_I, body_inertia=body_inertia, body_iquat=body_iquat, body_mass=body_mass,
        body_jntadr=np.asarray(body_jntadr, np.int32), body_jntnum=np.asarray(body_jntnum, np.int32),
        body_dofadr=np.asarray(body_dofadr, np.int32), body_dofnum=np.asarray(body_dofnum, np.int32),
        body_invweight0=np.zeros((nbody, 2)),
        jnt_name=J['name'], jnt_type=np.asarray(J['type'], np.int32), jnt_bodyid=np.asarray(J['body'], np.int32),
        jnt_pos=np.asarray(J['pos']).reshape(njnt, 3), jnt_axis=np.asarray(J['axis']).reshape(njnt, 3),
        jnt_qposadr=np.asarray(J['qposadr'], np.int32), jnt_dofadr=np.asarray(J['dofadr'], np.int32),
        jnt_limited=np.asarray(J['limited'], np.int32), jnt_range=np.asarray(J['range']).reshape(njnt, 2),
        jnt_stiffness=np.asarray(J['stiffness'], float), jnt_margin=np.asarray(J['margin'], float),
        jnt_solref=np.asarray(J['solref']).reshape(njnt, 2), jnt_solimp=np.asarray(J['solimp']).reshape(njnt, 5),
        qpos0=jnt_qpos0.copy(), qpos_spring=qpos_spring,
        dof_bodyid=np.asarray([J['body'][jid] for jid in dof_jnt], np.int32), dof_jntid=np.asarray(dof_jnt, np.int32),
        dof_parentid=np.zeros(nv, np.int32), dof_damping=np.asarray(dof_damping, float),
        dof_armature=np.asarray(dof_armature, float), dof_invweight0=np.zeros(nv),
        geom_name=G['name'], geom_type=np.asarray(G['type'], np.int32), geom_bodyid=np.asarray(G['body'], np.int32),
        geom_pos=np.asarray(G['pos']).reshape(ngeom, 3), geom_quat=np.asarray(G['quat']).reshape(ngeom, 4),
        geom_size=np.asarray(G['size']).reshape(ngeom, 3), geom_friction=np.asarray(G['friction']).reshape(ngeom, 3),
        geom_margin=np.asarray(G['margin'], float), geom_gap=np.asarray(G['gap'], float),
        geom_condim=np.asarray(G['condim'], np.int32), geom_contype=np.asarray(G['contype'], np.int32),
        geom_conaffinity=np.asarray(G['conaffinity'], np.int32),
        geom_solref=np.asarray(G['solref']).reshape(ngeom, 2), geom_solimp=np.asarray(G['solimp']).reshape(ngeom, 5),
        actuator_trnid=np.zeros(0, np.int32), actuator_gear=np.zeros(0), actuator_ctrllimited=np.zeros(0, np.int32),
        actuator_ctrlrange=np.zeros((0, 2)), actuator_forcelimited=np.zeros(0, np.int32),
        actuator_forcerange=np.zeros((0, 2)), actuator_gainprm0=np.zeros(0), actuator_biasprm=np.zeros((0, 3)),
        stat_meaninertia=1.0, contact_pairs=[])
    mdl.contact_pairs = []
    # dof tree: parent dof = previous dof of the same body, else last dof of the nearest ancestor with dofs
    for i in range(nv):
        b = mdl.dof_bodyid[i]
        if i > 0 and mdl.dof_bodyid[i - 1] == b:
            mdl.dof_parentid[i] = i - 1
        else:
            p = mdl.body_parentid[b]
            while p > 0 and mdl.body_dofnum[p] == 0:
                p = mdl.body_parentid[p]
            mdl.dof_parentid[i] = (mdl.body_dofadr[p] + mdl.body_dofnum[p] - 1) if p > 0 else -1
    # ---- actuators --------------------------------------------------------------
    acts = []
    for a in spec.actuators:
        rec = dict(ACT_DEFAULT)
        rec.update(spec.defaults['actuator'])
        if '_tag' in a:
            _apply_actuator(rec, a['_tag'], a['_attrib'])
        else:
            rec.update(a)
        acts.append(rec)
    nu = len(acts)
    mdl.nu = nu
    mdl.actuator_trnid = np.asarray([mdl.jnt_id(a['joint']) for a in acts], np.int32)
    mdl.actuator_gear = np.asarray([a['gear'] for a in acts], float)
    mdl.actuator_ctrllimited = np.asarray([int(bool(a['ctrllimited'])) for a in acts], np.int32)
    mdl.actuator_ctrlrange = np.asarray([a['ctrlrange'] for a in acts], float).reshape(nu, 2)
    mdl.actuator_forcelimited = np.asarray([int(bool(a['forcelimited'])) for a in acts], np.int32)
    mdl.actuator_forcerange = np.asarray([a['forcerange'] for a in acts], float).reshape(nu, 2)
    mdl.actuator_gainprm0 = np.asarray([a['gainprm0'] for a in acts], float)
    mdl.actuator_biasprm = np.asarray([a['biasprm'] for a in acts], float).reshape(nu, 3)
    # ---- collision pairs (broad-phase filters, engine_collision_driver.c) ---------------------
    # bodies without joints are welded to their parent: filters work on weld ids
    weld = np.zeros(nbody, np.int32)
    for b in range(1, nbody):
        weld[b] = b if mdl.body_jntnum[b] > 0 else weld[mdl.body_parentid[b]]
    mdl.body_weldid = weld
    for g1 in range(ngeom):
        for g2 in range(g1 + 1, ngeom):
            w1, w2 = weld[mdl.geom_bodyid[g1]], weld[mdl.geom_bodyid[g2]]
            if w1 == w2:
                continue
            if not ((mdl.geom_contype[g1] & mdl.geom_conaffinity[g2]) or (mdl.geom_contype[g2] & mdl.geom_conaffinity[g1])):
                continue
            # parent-child filter on welded bodies (does not apply when the parent is the world)
            p1, p2 = weld[mdl.body_parentid[w1]] if w1 else -1, weld[mdl.body_parentid[w2]] if w2 else -1
            if (w1 != 0 and w1 == p2) or (w2 != 0 and w2 == p1):
                continue
            mdl.contact_pairs.append((g1, g2))
    _set_const(mdl)
    return mdl


def _mat2quat(R):
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = [0, 0, 0, 0]
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q = np.asarray(q)
    return q / np.linalg.norm(q)


# ------------------------------------------------------------------ qpos0 constants
def kinematics0(m: Model, qpos=None):
    """World poses at ``qpos`` (default qpos0): xpos, xmat, xipos, xanchor, xaxis (float64)."""
    qpos = m.qpos0 if qpos is None else qpos
    xpos = np.zeros((m.nbody, 3)); xquat = np.tile([1., 0, 0, 0], (m.nbody, 1)); xmat = np.tile(np.eye(3), (m.nbody, 1, 1))
    xanchor = np.zeros((m.njnt, 3)); xaxis = np.zeros((m.njnt, 3))
    for b in range(1, m.nbody):
        p = m.body_parentid[b]
        pos = xpos[p] + xmat[p] @ m.body_pos[b]
        quat = quat_mul(xquat[p], m.body_quat[b])
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            if m.jnt_type[j] == JNT_FREE:      # the pose IS the qpos
                a = m.jnt_qposadr[j]
                pos = np.array(qpos[a:a + 3], float); quat = np.array(qpos[a + 3:a + 7], float)
                quat = quat / np.linalg.norm(quat)
                xanchor[j] = pos; xaxis[j] = quat2mat(quat) @ m.jnt_axis[j]
                continue
            R = quat2mat(quat)
            xanchor[j] = pos + R @ m.jnt_pos[j]
            xaxis[j] = R @ m.jnt_axis[j]
            dq = qpos[m.jnt_qposadr[j]] - m.qpos0[m.jnt_qposadr[j]]
            if m.jnt_type[j] == JNT_SLIDE:
                pos = pos + xaxis[j] * dq
            else:
                ql = np.array([math.cos(dq / 2), *(m.jnt_axis[j] * math.sin(dq / 2))])
                quat = quat_mul(quat, ql)
                pos = xanchor[j] - quat2mat(quat) @ m.jnt_pos[j]
        quat = quat / np.linalg.norm(quat)
        xpos[b], xquat[b], xmat[b] = pos, quat, quat2mat(quat)
    xipos = xpos + np.einsum('bij,bj->bi', xmat, m.body_ipos)
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, xanchor=xanchor, xaxis=xaxis)


def mass_matrix(m: Model, kin) -> np.ndarray:
    """Joint-space inertia at the pose ``kin`` (+ armature), by summing body Jacobian terms."""
    M = np.diag(m.dof_armature.copy())
    for b in range(1, m.nbody):
        jp, jr = body_jac(m, kin, b, kin['xipos'][b])
        Iw = kin['xmat'][b] @ m.body_inertia_full[b] @ kin['xmat'][b].T
        M += m.body_mass[b] * jp.T @ jp + jr.T @ Iw @ jr
    return M


def body_jac(m: Model, kin, b: int, point):
    jp = np.zeros((3, m.nv)); jr = np.zeros((3, m.nv))
    while b > 0:
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            i = m.jnt_dofadr[j]
            if m.jnt_type[j] == JNT_FREE:      # 3 translations along world axes, 3 rotations about body axes
                jp[:, i:i + 3] = np.eye(3)
                R = kin['xmat'][b]
                for a in range(3):
                    jr[:, i + 3 + a] = R[:, a]
                    jp[:, i + 3 + a] = np.cross(R[:, a], point - kin['xpos'][b])
            elif m.jnt_type[j] == JNT_SLIDE:
                jp[:, i] = kin['xaxis'][j]
            else:
                jr[:, i] = kin['xaxis'][j]
                jp[:, i] = np.cross(kin['xaxis'][j], point - kin['xanchor'][j])
        b = m.body_parentid[b]
    return jp, jr


def _set_const(m: Model):
    """dof_invweight0, body_invweight0, stat.meaninertia at qpos0 (engine_setconst.c: set0)."""
    if m.nv == 0:
        return
    kin = kinematics0(m)
    M = mass_matrix(m, kin)
    Minv = np.linalg.inv(M)
    m.stat_meaninertia = float(np.trace(M) / m.nv)
    m.dof_invweight0 = np.diag(Minv).copy()
    for j in range(m.njnt):                # free joint: one value for the 3 translations, one for the 3 rotations
        if m.jnt_type[j] == JNT_FREE:
            i = m.jnt_dofadr[j]
            m.dof_invweight0[i:i + 3] = m.dof_invweight0[i:i + 3].mean()
            m.dof_invweight0[i + 3:i + 6] = m.dof_invweight0[i + 3:i + 6].mean()
    for b in range(1, m.nbody):
        jp, jr = body_jac(m, kin, b, kin['xipos'][b])
        A = jp @ Minv @ jp.T
        B = jr @ Minv @ jr.T
        m.body_invweight0[b, 0] = max(mjMINVAL, np.trace(A) / 3)
        m.body_invweight0[b, 1] = max(mjMINVAL, np.trace(B) / 3)
