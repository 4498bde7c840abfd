"""CPU restatement of ``mujoco.mjx.step`` / ``mjx.forward`` for hinge/slide trees.

TEST INFRASTRUCTURE ONLY (see oracle/oracle.py).  PARITY UNPINNED: ``mujoco`` and
``mujoco_mjx`` are un-vendored, un-pinned dependencies of the reference
(requirements.txt:3-4; 3.1.x by API use) and cannot be installed in
this image, and the reference ships no test or fixture for this path.  This file restates
the published MuJoCo/MJX pipeline at the reference's call sites (engine.py:655 forward,
:668 and :721 step), function by function, on the FULL compiled model -- the robot and
the 2 + 2*hazards_num static slide DoF of the goal/hazard bodies that World.build adds
(world.py:310-326) -- with a dense mass matrix, as MJX does for nv < 60:

  fwd_position   kinematics, com_pos (subtree COM frame, cinert, cdof), crb, factor_m,
                 collision (sphere-plane), make_constraint (joint limits, pyramidal contacts)
  fwd_velocity   com_vel, passive (spring + damper), rne (bias incl. gravity)
  fwd_actuation  ctrl clamp, gain/affine bias, force clamp, joint transmission
  fwd_accel.     qacc_smooth = M^-1 (passive - bias + actuator)
  solver.solve   Newton on the primal cost with warm start (cost(qacc_warmstart) vs
                 cost(qacc_smooth)), exact line search on the piecewise-quadratic 1-D cost
                 (MJX's iterative bracketing search is not reproduced: both find the
                 same minimiser; the Newton fixed point is unique),
                 termination on scaled improvement / gradient < opt.tolerance
  euler          implicit joint damping, semi-implicit integration

Everything is batched over a leading env dimension and evaluated in ``dtype``
(float32 mirrors MJX; float64 gives the rounding-free answer the fp32 CUDA kernels are
also compared with).  The CUDA kernels (guardx_b200/csrc/gx_art.cuh) use a different
formulation (robot block only, world-origin frame, fused limits/contacts), so agreement is
an independent check, not shared code.
"""
from __future__ import annotations

import numpy as np

from . import mj_compile as mjc     # the oracle's own model compiler: nothing here imports the product

mjMINVAL = 1e-15
mjMINIMP, mjMAXIMP = 0.0001, 0.9999


def _cross(a, b):
    return np.cross(a, b)


def quat_mul(a, b):
    return np.stack([a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1] - a[..., 2] * b[..., 2] - a[..., 3] * b[..., 3],
                     a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0] + a[..., 2] * b[..., 3] - a[..., 3] * b[..., 2],
                     a[..., 0] * b[..., 2] - a[..., 1] * b[..., 3] + a[..., 2] * b[..., 0] + a[..., 3] * b[..., 1],
                     a[..., 0] * b[..., 3] + a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1] + a[..., 3] * b[..., 0]], -1)


def quat2mat(q):
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    return np.stack([np.stack([w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)], -1),
                     np.stack([2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)], -1),
                     np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z], -1)], -2)


class MjxModel:
    """mjx.put_model: the compiled constants cast to the working dtype."""

    def __init__(self, m: 'mjc.OModel', dtype=np.float32, capsules=None):
        self.m = m
        self.dtype = dtype
        f = lambda a: np.asarray(a, dtype)
        self.nq, self.nv, self.nu, self.nbody, self.njnt = m.nq, m.nv, m.nu, m.nbody, m.njnt
        for k in ('body_pos', 'body_quat', 'body_ipos', 'body_mass', 'body_inertia_full', 'body_invweight0',
                  'jnt_pos', 'jnt_axis', 'jnt_range', 'jnt_stiffness', 'jnt_margin', 'jnt_solref', 'jnt_solimp',
                  'qpos0', 'qpos_spring', 'dof_damping', 'dof_armature', 'dof_invweight0', 'geom_pos', 'geom_quat',
                  'geom_size', 'geom_friction', 'geom_margin', 'geom_gap', 'geom_solref', 'geom_solimp',
                  'actuator_gear', 'actuator_ctrlrange', 'actuator_forcerange', 'actuator_gainprm0', 'actuator_biasprm'):
            setattr(self, k, f(getattr(m, k)))
        self.timestep = dtype(m.opt['timestep'])
        self.gravity = f(m.opt['gravity'])
        self.tolerance = m.opt['tolerance']
        self.iterations = m.opt['iterations']
        self.impratio = dtype(m.opt['impratio'])
        self.meaninertia = dtype(m.stat_meaninertia)
        # tree bookkeeping
        self.body_rootid = np.zeros(m.nbody, np.int32)
        for b in range(1, m.nbody):
            p = m.body_parentid[b]
            self.body_rootid[b] = b if p == 0 else self.body_rootid[p]
        # contact candidates by collision function: (plane, sphere), (plane, capsule), (capsule, capsule)
        self.sphere_plane, self.capsule_plane, self.capsule_capsule = [], [], []
        self.ignored_pairs = []
        # capsule contacts are restated for free-joint robots (Doggo); for the planar robots the
        # capsule-plane pairs of the swimmer sit exactly at dist == margin and are treated as never
        # active (assumption A2), as is the Point's box-plane pair
        with_capsules = bool((m.jnt_type == mjc.FREE).any()) if capsules is None else capsules
        for g1, g2 in m.contact_pairs:
            t1, t2 = int(m.geom_type[g1]), int(m.geom_type[g2])
            if t2 == mjc.PLANE:
                g1, g2, t1, t2 = g2, g1, t2, t1
            if t1 == mjc.PLANE and t2 == mjc.SPHERE:
                self.sphere_plane.append((g1, g2))
            elif with_capsules and t1 == mjc.PLANE and t2 == mjc.CAPSULE:
                self.capsule_plane.append((g1, g2))
            elif with_capsules and t1 == mjc.CAPSULE and t2 == mjc.CAPSULE:
                self.capsule_capsule.append((g1, g2))
            else:
                self.ignored_pairs.append((g1, g2))
        self.limited = [j for j in range(m.njnt) if m.jnt_limited[j]]


# ----------------------------------------------------------------- fwd_position
def kinematics(mm: MjxModel, qpos):
    m, dt = mm.m, mm.dtype
    N = qpos.shape[0]
    xpos = np.zeros((N, m.nbody, 3), dt)
    xquat = np.zeros((N, m.nbody, 4), dt); xquat[:, :, 0] = 1
    xmat = np.zeros((N, m.nbody, 3, 3), dt); xmat[:] = np.eye(3, dtype=dt)
    xanchor = np.zeros((N, m.njnt, 3), dt)
    xaxis = np.zeros((N, m.njnt, 3), dt)
    for b in range(1, m.nbody):
        p = m.body_parentid[b]
        pos = xpos[:, p] + np.einsum('nij,j->ni', xmat[:, p], mm.body_pos[b])
        quat = quat_mul(xquat[:, p], np.broadcast_to(mm.body_quat[b], (N, 4)))
        for k in range(m.body_jntnum[b]):
            j = m.body_jntadr[b] + k
            if m.jnt_type[j] == mjc.FREE:     # the pose is the qpos (position + normalised quaternion)
                a = m.jnt_qposadr[j]
                pos = qpos[:, a:a + 3].copy()
                quat = qpos[:, a + 3:a + 7] / np.linalg.norm(qpos[:, a + 3:a + 7], axis=1, keepdims=True)
                xanchor[:, j] = pos
                xaxis[:, j] = np.einsum('nij,j->ni', quat2mat(quat), mm.jnt_axis[j])
                continue
            R = quat2mat(quat)
            xanchor[:, j] = pos + np.einsum('nij,j->ni', R, mm.jnt_pos[j])
            xaxis[:, j] = np.einsum('nij,j->ni', R, mm.jnt_axis[j])
            dq = qpos[:, m.jnt_qposadr[j]] - mm.qpos0[m.jnt_qposadr[j]]
            if m.jnt_type[j] == mjc.SLIDE:
                pos = pos + xaxis[:, j] * dq[:, None]
            else:
                half = dq * dt(0.5)
                ql = np.concatenate([np.cos(half)[:, None], mm.jnt_axis[j][None, :] * np.sin(half)[:, None]], 1).astype(dt)
                quat = quat_mul(quat, ql)
                pos = xanchor[:, j] - np.einsum('nij,j->ni', quat2mat(quat), mm.jnt_pos[j])
        quat = quat / np.linalg.norm(quat, axis=1, keepdims=True)
        xpos[:, b], xquat[:, b], xmat[:, b] = pos, quat, quat2mat(quat)
    xipos = xpos + np.einsum('nbij,bj->nbi', xmat, mm.body_ipos)
    return dict(xpos=xpos, xquat=xquat, xmat=xmat, xipos=xipos, xanchor=xanchor, xaxis=xaxis)


def com_pos(mm: MjxModel, d):
    """subtree COM of each kinematic tree, cinert (10-vector) and cdof about it."""
    m, dt = mm.m, mm.dtype
    N = d['xpos'].shape[0]
    mpos = d['xipos'] * mm.body_mass[None, :, None]
    sub_m = mm.body_mass.copy()
    sub_mp = mpos.copy()
    for b in range(m.nbody - 1, 0, -1):
        p = m.body_parentid[b]
        sub_m[p] += sub_m[b]
        sub_mp[:, p] += sub_mp[:, b]
    subtree_com = sub_mp / np.maximum(sub_m, dt(mjMINVAL))[None, :, None]
    root_com = subtree_com[:, mm.body_rootid]          # [N, nbody, 3]
    # cinert
    Iw = np.einsum('nbij,bjk,nblk->nbil', d['xmat'], mm.body_inertia_full, d['xmat'])
    dif = d['xipos'] - root_com
    mass = mm.body_mass[None, :]
    dd = np.einsum('nbi,nbi->nb', dif, dif)
    Ic = Iw + mass[..., None, None] * (dd[..., None, None] * np.eye(3, dtype=dt) - np.einsum('nbi,nbj->nbij', dif, dif))
    cin = dict(I=Ic, mh=dif * mass[..., None], m=np.broadcast_to(mass, (N, m.nbody)).astype(dt))
    # cdof: [ang; lin] about the tree COM
    cdof = np.zeros((N, m.nv, 6), dt)
    for j in range(m.njnt):
        i = m.jnt_dofadr[j]
        b = m.jnt_bodyid[j]
        if m.jnt_type[j] == mjc.FREE:
            # three translations along the world axes, three rotations about the body axes
            for a in range(3):
                cdof[:, i + a, 3 + a] = 1
                ax = d['xmat'][:, b, :, a]
                cdof[:, i + 3 + a, :3] = ax
                cdof[:, i + 3 + a, 3:] = _cross(ax, root_com[:, b] - d['xpos'][:, b])
        elif m.jnt_type[j] == mjc.SLIDE:
            cdof[:, i, 3:] = d['xaxis'][:, j]
        else:
            cdof[:, i, :3] = d['xaxis'][:, j]
            cdof[:, i, 3:] = _cross(d['xaxis'][:, j], root_com[:, b] - d['xanchor'][:, j])
    d.update(subtree_com=subtree_com, root_com=root_com, cinert=cin, cdof=cdof)
    return d


def _inert_mul(I, mh, mass, v):
    """spatial inertia (about the reference point) times motion vector [ang; lin] -> force [ang; lin]."""
    ang = np.einsum('...ij,...j->...i', I, v[..., :3]) + _cross(mh, v[..., 3:])
    lin = mass[..., None] * v[..., 3:] - _cross(mh, v[..., :3])
    return np.concatenate([ang, lin], -1)


def crb(mm: MjxModel, d):
    m, dt = mm.m, mm.dtype
    cin = d['cinert']
    I, mh, ms = cin['I'].copy(), cin['mh'].copy(), cin['m'].copy()
    for b in range(m.nbody - 1, 0, -1):
        p = m.body_parentid[b]
        if p > 0:
            I[:, p] += I[:, b]; mh[:, p] += mh[:, b]; ms[:, p] += ms[:, b]
    N = I.shape[0]
    M = np.zeros((N, m.nv, m.nv), dt)
    for i in range(m.nv):
        b = m.dof_bodyid[i]
        f = _inert_mul(I[:, b], mh[:, b], ms[:, b], d['cdof'][:, i])
        j = i
        while j >= 0:
            M[:, i, j] = M[:, j, i] = np.einsum('nk,nk->n', d['cdof'][:, j], f)
            j = m.dof_parentid[j]
    M[:, np.arange(m.nv), np.arange(m.nv)] += mm.dof_armature
    d['qM'] = M
    return d


def _cross_motion(v, w):
    return np.concatenate([_cross(v[..., :3], w[..., :3]),
                           _cross(v[..., :3], w[..., 3:]) + _cross(v[..., 3:], w[..., :3])], -1)


def _cross_force(v, f):
    return np.concatenate([_cross(v[..., :3], f[..., :3]) + _cross(v[..., 3:], f[..., 3:]),
                           _cross(v[..., :3], f[..., 3:])], -1)


# ----------------------------------------------------------------- fwd_velocity
def com_vel(mm: MjxModel, d, qvel):
    m, dt = mm.m, mm.dtype
    N = qvel.shape[0]
    cvel = np.zeros((N, m.nbody, 6), dt)
    cdof_dot = np.zeros((N, m.nv, 6), dt)
    for b in range(1, m.nbody):
        v = cvel[:, m.body_parentid[b]].copy()
        k0 = 0
        if m.body_jntnum[b] and m.jnt_type[m.body_jntadr[b]] == mjc.FREE:
            # mj_comVel: the translations have cdof_dot = 0; the three rotations all use the velocity
            # after the translations (not sequentially updated)
            i0 = m.body_dofadr[b]
            for a in range(3):
                v = v + d['cdof'][:, i0 + a] * qvel[:, i0 + a, None]
            for a in range(3):
                cdof_dot[:, i0 + 3 + a] = _cross_motion(v, d['cdof'][:, i0 + 3 + a])
            for a in range(3):
                v = v + d['cdof'][:, i0 + 3 + a] * qvel[:, i0 + 3 + a, None]
            k0 = 6
        for k in range(k0, m.body_dofnum[b]):
            i = m.body_dofadr[b] + k
            cdof_dot[:, i] = _cross_motion(v, d['cdof'][:, i])
            v = v + d['cdof'][:, i] * qvel[:, i, None]
        cvel[:, b] = v
    d.update(cvel=cvel, cdof_dot=cdof_dot)
    return d


def passive(mm: MjxModel, qpos, qvel):
    m = mm.m
    f = -mm.dof_damping[None, :] * qvel
    for j in range(m.njnt):
        if m.jnt_stiffness[j] != 0 and m.jnt_type[j] in (mjc.SLIDE, mjc.HINGE):
            f[:, m.jnt_dofadr[j]] -= mm.jnt_stiffness[j] * (qpos[:, m.jnt_qposadr[j]] - mm.qpos_spring[m.jnt_qposadr[j]])
    return f


def rne(mm: MjxModel, d, qvel):
    m, dt = mm.m, mm.dtype
    N = qvel.shape[0]
    cin = d['cinert']
    cacc = np.zeros((N, m.nbody, 6), dt)
    cacc[:, 0, 3:] = -mm.gravity
    cfrc = np.zeros((N, m.nbody, 6), dt)
    for b in range(1, m.nbody):
        a = cacc[:, m.body_parentid[b]].copy()
        for k in range(m.body_dofnum[b]):
            i = m.body_dofadr[b] + k
            a = a + d['cdof_dot'][:, i] * qvel[:, i, None]
        cacc[:, b] = a
        Ia = _inert_mul(cin['I'][:, b], cin['mh'][:, b], cin['m'][:, b], a)
        Iv = _inert_mul(cin['I'][:, b], cin['mh'][:, b], cin['m'][:, b], d['cvel'][:, b])
        cfrc[:, b] = Ia + _cross_force(d['cvel'][:, b], Iv)
    for b in range(m.nbody - 1, 0, -1):
        p = m.body_parentid[b]
        if p > 0:
            cfrc[:, p] += cfrc[:, b]
    bias = np.zeros((N, m.nv), dt)
    for i in range(m.nv):
        bias[:, i] = np.einsum('nk,nk->n', d['cdof'][:, i], cfrc[:, m.dof_bodyid[i]])
    return bias


# ----------------------------------------------------------------- fwd_actuation
def actuation(mm: MjxModel, qvel, ctrl):
    m, dt = mm.m, mm.dtype
    N = qvel.shape[0]
    qfrc = np.zeros((N, m.nv), dt)
    for a in range(m.nu):
        i = m.jnt_dofadr[m.actuator_trnid[a]]
        u = ctrl[:, a]
        if m.actuator_ctrllimited[a]:
            u = np.clip(u, mm.actuator_ctrlrange[a, 0], mm.actuator_ctrlrange[a, 1])
        force = mm.actuator_gainprm0[a] * u
        bp = mm.actuator_biasprm[a]
        if bp[2] != 0 or bp[1] != 0 or bp[0] != 0:
            # affine bias: b0 + b1*length + b2*velocity, joint transmission: velocity = gear*qvel
            force = force + bp[0] + bp[2] * (mm.actuator_gear[a] * qvel[:, i])
            if bp[1] != 0:
                raise NotImplementedError('position-servo bias')
        if m.actuator_forcelimited[a]:
            force = np.clip(force, mm.actuator_forcerange[a, 0], mm.actuator_forcerange[a, 1])
        qfrc[:, i] += mm.actuator_gear[a] * force
    return qfrc


# ----------------------------------------------------------------- constraints
def _jac_point(mm: MjxModel, d, b, point):
    """translational Jacobian [N,3,nv] of the world point `point` moving with body b (from cdof)."""
    m = mm.m
    N = point.shape[0]
    jacp = np.zeros((N, 3, m.nv), mm.dtype)
    off = point - d['root_com'][:, b]
    i = m.body_dofadr[b] + m.body_dofnum[b] - 1 if m.body_dofnum[b] else -1
    if i < 0:
        p = m.body_parentid[b]
        while p > 0 and m.body_dofnum[p] == 0:
            p = m.body_parentid[p]
        i = m.body_dofadr[p] + m.body_dofnum[p] - 1 if p > 0 else -1
    while i >= 0:
        jacp[:, :, i] = d['cdof'][:, i, 3:] + _cross(d['cdof'][:, i, :3], off)
        i = m.dof_parentid[i]
    return jacp


def _kbi(mm: MjxModel, solref, solimp, pos):
    dt = mm.dtype
    timeconst, dampratio = dt(solref[0]), dt(solref[1])
    timeconst = max(timeconst, dt(2) * mm.timestep)        # REFSAFE
    dmin, dmax, width, mid, power = (dt(x) for x in solimp)
    dmin = min(max(dmin, dt(mjMINIMP)), dt(mjMAXIMP)); dmax = min(max(dmax, dt(mjMINIMP)), dt(mjMAXIMP))
    width = max(dt(mjMINVAL), width); mid = min(max(mid, dt(mjMINIMP)), dt(mjMAXIMP)); power = max(dt(1), power)
    k = dt(1) / (dmax * dmax * timeconst * timeconst * dampratio * dampratio)
    b = dt(2) / (dmax * timeconst)
    if solref[0] <= 0:
        k = -dt(solref[0]) / (dmax * dmax)
    if solref[1] <= 0:
        b = -dt(solref[1]) / dmax
    imp_x = np.abs(pos) / width
    imp_a = (dt(1) / np.power(mid, power - 1)) * np.power(imp_x, power)
    imp_b = dt(1) - (dt(1) / np.power(dt(1) - mid, power - 1)) * np.power(np.maximum(dt(1) - imp_x, 0), power)
    imp_y = np.where(imp_x < mid, imp_a, imp_b)
    imp = np.clip(dmin + imp_y * (dmax - dmin), dmin, dmax)
    imp = np.where(imp_x > 1.0, dmax, imp).astype(dt)
    return k, b, imp


def make_constraint(mm: MjxModel, d, qpos, qvel):
    """efc rows (J, D, aref) for joint limits and pyramidal sphere-plane contacts.  Rows that
    MJX would zero out (inactive) carry J = 0, aref = 0, D = 0 here."""
    m, dt = mm.m, mm.dtype
    N = qpos.shape[0]
    Js, Ds, arefs, meta = [], [], [], []
    for j in mm.limited:
        q = qpos[:, m.jnt_qposadr[j]]
        dist_min, dist_max = q - mm.jnt_range[j, 0], mm.jnt_range[j, 1] - q
        pos = np.minimum(dist_min, dist_max) - mm.jnt_margin[j]
        active = pos < 0
        sign = np.where(dist_min < dist_max, dt(1), dt(-1))
        J = np.zeros((N, m.nv), dt)
        J[:, m.jnt_dofadr[j]] = sign * active
        k, b, imp = _kbi(mm, m.jnt_solref[j], m.jnt_solimp[j], pos)
        R = np.maximum(mm.dof_invweight0[m.jnt_dofadr[j]] * (1 - imp) / imp, dt(mjMINVAL))
        aref = -b * np.einsum('nv,nv->n', J, qvel) - k * imp * pos
        Js.append(J); Ds.append(np.where(active, 1 / R, 0).astype(dt)); arefs.append(np.where(active, aref, 0).astype(dt))
        meta.append(('limit', j))
    contact_dist = []

    def geom_pose(g):
        bd = m.geom_bodyid[g]
        pos = d['xpos'][:, bd] + np.einsum('nij,j->ni', d['xmat'][:, bd], mm.geom_pos[g])
        mat = np.einsum('nij,jk->nik', d['xmat'][:, bd], mjc._rot(m.geom_quat[g]).astype(dt))
        return pos, mat

    def make_frame(a):
        """math.make_frame: tangent from the y (or z) axis, orthogonalised against the normal."""
        yref = np.where((np.abs(a[:, 1]) < 0.5)[:, None], np.array([0, 1, 0], dt), np.array([0, 0, 1], dt))
        t1 = yref - a * np.einsum('ni,ni->n', a, yref)[:, None]
        t1 = t1 / np.linalg.norm(t1, axis=1, keepdims=True)
        return t1, _cross(a, t1)

    def add_contact(g1, g2, dist, cpos, nrm, t1, t2):
        """pyramidal rows (condim 3: 4 per contact) of one contact between geoms g1 and g2."""
        b1, b2 = m.geom_bodyid[g1], m.geom_bodyid[g2]
        margin = max(mm.geom_margin[g1], mm.geom_margin[g2])
        gap = max(mm.geom_gap[g1], mm.geom_gap[g2])
        includemargin = margin - gap
        fri = np.maximum(mm.geom_friction[g1], mm.geom_friction[g2])
        condim = max(m.geom_condim[g1], m.geom_condim[g2])
        assert condim == 3, 'only condim 3 (pyramidal, 4 rows per contact) is restated'
        solref = 0.5 * (m.geom_solref[g1] + m.geom_solref[g2])   # solmix 1:1
        solimp = 0.5 * (m.geom_solimp[g1] + m.geom_solimp[g2])
        jac = _jac_point(mm, d, b2, cpos) - _jac_point(mm, d, b1, cpos)     # body2 - body1
        jn = np.einsum('ni,niv->nv', nrm, jac)
        jt = [np.einsum('ni,niv->nv', t1, jac), np.einsum('ni,niv->nv', t2, jac)]
        pos = dist - includemargin
        active = pos < 0
        invweight = mm.body_invweight0[b1, 0] + mm.body_invweight0[b2, 0]
        invweight = invweight + fri[0] * fri[0] * invweight
        invweight = invweight * 2 * fri[0] * fri[0] / mm.impratio
        k, b, imp = _kbi(mm, solref, solimp, pos)
        R = np.maximum(invweight * (1 - imp) / imp, dt(mjMINVAL))
        for t in range(2):
            for sgn in (1, -1):
                J = (jn + dt(sgn) * fri[0] * jt[t]) * active[:, None]   # both tangents: sliding friction
                aref = -b * np.einsum('nv,nv->n', J, qvel) - k * imp * pos
                Js.append(J.astype(dt)); Ds.append(np.where(active, 1 / R, 0).astype(dt))
                arefs.append(np.where(active, aref, 0).astype(dt))
                meta.append(('contact', g2))
        contact_dist.append(dist)

    def plane_sphere(n, ppos, spos, r):
        dist = np.einsum('ni,ni->n', spos - ppos, n) - r
        return dist, spos - n * (r + dist * dt(0.5))[:, None]

    for (gp, gs) in mm.sphere_plane:                     # collision_primitive.plane_sphere
        ppos, pmat = geom_pose(gp)
        spos, _ = geom_pose(gs)
        n = pmat[:, :, 2]
        dist, cpos = plane_sphere(n, ppos, spos, mm.geom_size[gs, 0])
        t1, t2 = make_frame(n)
        add_contact(gp, gs, dist, cpos, n, t1, t2)
    for (gp, gc) in mm.capsule_plane:                    # collision_primitive.plane_capsule: one contact per end
        ppos, pmat = geom_pose(gp)
        cpos_, cmat = geom_pose(gc)
        n, axis = pmat[:, :, 2], cmat[:, :, 2]
        bvec = axis - n * np.einsum('ni,ni->n', n, axis)[:, None]
        bn = np.linalg.norm(bvec, axis=1)
        with np.errstate(all='ignore'):
            bdir = bvec / bn[:, None]
        alt = np.where(((-0.5 < n[:, 1]) & (n[:, 1] < 0.5))[:, None], np.array([0, 1, 0], dt), np.array([0, 0, 1], dt))
        bdir = np.where((bn < 0.5)[:, None], alt, bdir).astype(dt)
        t2 = _cross(n, bdir)
        seg = axis * mm.geom_size[gc, 1]
        for off in (seg, -seg):
            dist, cp = plane_sphere(n, ppos, cpos_ + off, mm.geom_size[gc, 0])
            add_contact(gp, gc, dist, cp, n, bdir, t2)
    for (g1, g2) in mm.capsule_capsule:                  # collision_primitive.capsule_capsule
        p1, m1 = geom_pose(g1)
        p2, m2 = geom_pose(g2)
        a1, a2 = m1[:, :, 2] * mm.geom_size[g1, 1], m2[:, :, 2] * mm.geom_size[g2, 1]
        pt1, pt2 = closest_segment_to_segment_points(p1 - a1, p1 + a1, p2 - a2, p2 + a2)
        dvec = pt2 - pt1
        dn = np.linalg.norm(dvec, axis=1)
        with np.errstate(all='ignore'):
            n = np.where((dn == 0)[:, None], np.array([1, 0, 0], dt), dvec / dn[:, None]).astype(dt)
        r1, r2 = mm.geom_size[g1, 0], mm.geom_size[g2, 0]
        dist = dn - (r1 + r2)
        cp = pt1 + n * (r1 + dist * dt(0.5))[:, None]
        t1, t2 = make_frame(n)
        add_contact(g1, g2, dist, cp, n, t1, t2)
    if not Js:
        return np.zeros((N, 0, m.nv), dt), np.zeros((N, 0), dt), np.zeros((N, 0), dt), meta, contact_dist
    return np.stack(Js, 1), np.stack(Ds, 1), np.stack(arefs, 1), meta, contact_dist


def closest_segment_point(a, b, pt):
    ab = b - a
    t = np.einsum('ni,ni->n', pt - a, ab) / (np.einsum('ni,ni->n', ab, ab) + 1e-6)
    return a + np.clip(t, 0.0, 1.0)[:, None] * ab


def closest_segment_to_segment_points(a0, a1, b0, b1):
    """math.closest_segment_to_segment_points: closest points of the two infinite lines, clipped to
    the segments, then each clipped point re-projected onto the other segment."""
    da, db = a1 - a0, b1 - b0
    la, lb = np.linalg.norm(da, axis=1), np.linalg.norm(db, axis=1)
    dir_a, dir_b = da / la[:, None], db / lb[:, None]
    ha, hb = la * 0.5, lb * 0.5
    a_mid, b_mid = a0 + dir_a * ha[:, None], b0 + dir_b * hb[:, None]
    trans = a_mid - b_mid
    d_a = np.einsum('ni,ni->n', dir_a, trans)
    d_b = np.einsum('ni,ni->n', dir_b, trans)
    dd = np.einsum('ni,ni->n', dir_a, dir_b)
    denom = 1 - dd * dd
    t_a = (-d_a + d_b * dd) / (denom + 1e-6)
    t_b = (d_b + t_a * dd)
    t_a = np.clip(t_a, -ha, ha)
    t_b = np.clip(t_b, -hb, hb)
    best_a = a_mid + dir_a * t_a[:, None]
    best_b = b_mid + dir_b * t_b[:, None]
    new_a = closest_segment_point(a0, a1, best_b)
    new_b = closest_segment_point(b0, b1, best_a)
    d1 = np.sum((new_a - best_b) ** 2, 1)
    d2 = np.sum((new_b - best_a) ** 2, 1)
    best_a2 = np.where((d1 < d2)[:, None], new_a, best_a)
    best_b2 = np.where((d1 < d2)[:, None], best_b, new_b)
    return best_a2.astype(a0.dtype), best_b2.astype(a0.dtype)


# ----------------------------------------------------------------- solver
def _cost(mm, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc):
    Ma = np.einsum('nij,nj->ni', M, qacc)
    jaref = np.einsum('nev,nv->ne', J, qacc) - aref
    act = jaref < 0
    gauss = 0.5 * np.einsum('ni,ni->n', Ma - qfrc_smooth, qacc - qacc_smooth)
    cost = gauss + 0.5 * np.sum(D * jaref * jaref * act, 1)
    return cost, Ma, jaref, act


def linesearch_exact(g0, h0, jaref, jv, D, iters=40):
    """argmin_{alpha>=0} of  g0*alpha + 0.5*h0*alpha^2 + sum_i 0.5*D_i*min(0, jaref_i + alpha*jv_i)^2
    (convex, C1, piecewise quadratic): Newton on the piecewise-linear derivative with a bracket."""
    dt = jaref.dtype
    N = jaref.shape[0]
    alpha = np.zeros(N, dt)
    lo = np.zeros(N, dt)
    hi = np.full(N, np.inf, dt)
    done = np.zeros(N, bool)
    for _ in range(iters):
        x = jaref + alpha[:, None] * jv
        act = x < 0
        d1 = g0 + alpha * h0 + np.sum(D * jv * x * act, 1)
        d2 = h0 + np.sum(D * jv * jv * act, 1)
        lo = np.where(~done & (d1 < 0), alpha, lo)
        hi = np.where(~done & (d1 > 0), alpha, hi)
        with np.errstate(all='ignore'):
            nxt = alpha - d1 / d2
        bad = ~((nxt > lo) & (nxt < hi))
        mid = np.where(np.isfinite(hi), 0.5 * (lo + hi), 2 * np.maximum(lo, 1) )
        nxt = np.where(bad, mid, nxt)
        # converged when the Newton step lands on the same active set (then d1(nxt) == 0 exactly in exact arithmetic)
        xn = jaref + nxt[:, None] * jv
        same = np.all((xn < 0) == act, 1) & ~bad
        newly = ~done & (same | (d1 == 0))
        alpha = np.where(done, alpha, np.where(d1 == 0, alpha, nxt)).astype(dt)
        done |= newly
        if done.all():
            break
    return alpha


def _spd_solve(H, g):
    """H^-1 g per env; envs whose matrix is singular / not positive definite in the working
    precision (the planar roots of ant / walker degenerate at yaw = +-90 deg) get NaN, which the
    env's NaN guard turns into done = 1 (engine.py:696-699)."""
    out = np.full(g.shape, np.nan, H.dtype)
    H64, g64 = H.astype(np.float64), g.astype(np.float64)
    try:
        L = np.linalg.cholesky(H64)
        ok = np.arange(H.shape[0])
    except np.linalg.LinAlgError:
        ok = []
        for e in range(H.shape[0]):
            try:
                np.linalg.cholesky(H64[e]); ok.append(e)
            except np.linalg.LinAlgError:
                pass
        ok = np.asarray(ok, int)
        if ok.size == 0:
            return out
        L = np.linalg.cholesky(H64[ok])
    y = np.linalg.solve(L, g64[ok][..., None])
    out[ok] = np.linalg.solve(np.swapaxes(L, 1, 2), y)[..., 0].astype(H.dtype)
    return out


def solve(mm: MjxModel, M, J, D, aref, qfrc_smooth, qacc_smooth, warm, return_iters=False):
    m, dt = mm.m, mm.dtype
    N = M.shape[0]
    scale = 1.0 / (float(mm.meaninertia) * max(1, m.nv))
    c_w, *_ = _cost(mm, M, J, D, aref, qfrc_smooth, qacc_smooth, warm)
    c_s, *_ = _cost(mm, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc_smooth)
    qacc = np.where((c_w < c_s)[:, None], warm, qacc_smooth).astype(dt)
    cost, Ma, jaref, act = _cost(mm, M, J, D, aref, qfrc_smooth, qacc_smooth, qacc)
    done = np.zeros(N, bool)
    niter = np.zeros(N, np.int32)
    for it in range(mm.iterations):
        efc_force = -D * jaref * act
        grad = Ma - qfrc_smooth - np.einsum('nev,ne->nv', J, efc_force)
        gnorm = np.sqrt(np.sum(grad * grad, 1))
        done |= ~(gnorm * scale >= mm.tolerance)
        if done.all():
            break
        H = M + np.einsum('nev,ne,new->nvw', J, D * act, J)
        search = -_spd_solve(H, grad)
        mv = np.einsum('nij,nj->ni', M, search)
        jv = np.einsum('nev,nv->ne', J, search)
        g0 = np.einsum('ni,ni->n', search, Ma - qfrc_smooth)
        h0 = np.einsum('ni,ni->n', search, mv)
        alpha = linesearch_exact(g0, h0, jaref, jv, D)
        alpha = np.where(done, 0, alpha).astype(dt)
        qacc = qacc + alpha[:, None] * search
        Ma = Ma + alpha[:, None] * mv
        jaref = jaref + alpha[:, None] * jv
        act = jaref < 0
        prev = cost
        gauss = 0.5 * np.einsum('ni,ni->n', Ma - qfrc_smooth, qacc - qacc_smooth)
        cost = gauss + 0.5 * np.sum(D * jaref * jaref * act, 1)
        niter += ~done
        done |= ~((prev - cost) * scale >= mm.tolerance)
    efc_force = -D * jaref * act
    qfrc_constraint = np.einsum('nev,ne->nv', J, efc_force)
    if return_iters:
        return qacc.astype(dt), qfrc_constraint.astype(dt), efc_force, niter
    return qacc.astype(dt), qfrc_constraint.astype(dt), efc_force


# ----------------------------------------------------------------- forward / step
def _chol_solve(M, f):
    return _spd_solve(M, f)


def forward(mm: MjxModel, qpos, qvel, ctrl, warm):
    dt = mm.dtype
    qpos, qvel, ctrl, warm = (np.asarray(x, dt) for x in (qpos, qvel, ctrl, warm))
    d = kinematics(mm, qpos)
    d = com_pos(mm, d)
    d = crb(mm, d)
    d = com_vel(mm, d, qvel)
    qfrc_passive = passive(mm, qpos, qvel)
    qfrc_bias = rne(mm, d, qvel)
    qfrc_actuator = actuation(mm, qvel, ctrl)
    qfrc_smooth = qfrc_passive - qfrc_bias + qfrc_actuator
    qacc_smooth = _chol_solve(d['qM'], qfrc_smooth)
    J, D, aref, meta, cdist = make_constraint(mm, d, qpos, qvel)
    if J.shape[1] == 0:
        qacc, qfrc_c, efc_force, niter = qacc_smooth, np.zeros_like(qacc_smooth), np.zeros((qpos.shape[0], 0), dt), None
    else:
        qacc, qfrc_c, efc_force, niter = solve(mm, d['qM'], J, D, aref, qfrc_smooth, qacc_smooth, warm, True)
    d.update(qfrc_passive=qfrc_passive, qfrc_bias=qfrc_bias, qfrc_actuator=qfrc_actuator, qfrc_smooth=qfrc_smooth,
             qacc_smooth=qacc_smooth, qacc=qacc, qfrc_constraint=qfrc_c, efc_J=J, efc_D=D, efc_aref=aref,
             efc_force=efc_force, efc_meta=meta, contact_dist=cdist, solver_niter=niter, qacc_warmstart=qacc)
    return d


def step(mm: MjxModel, qpos, qvel, ctrl, warm):
    """mjx.step with the Euler integrator: returns (qpos', qvel', data of the forward pass)."""
    dt = mm.dtype
    d = forward(mm, qpos, qvel, ctrl, warm)
    qpos, qvel = np.asarray(qpos, dt), np.asarray(qvel, dt)
    Mh = d['qM'].copy()
    idx = np.arange(mm.nv)
    Mh[:, idx, idx] += mm.timestep * mm.dof_damping
    qacc = _chol_solve(Mh, d['qfrc_smooth'] + d['qfrc_constraint'])
    qvel2 = qvel + mm.timestep * qacc
    qpos2 = integrate_pos(mm, qpos, qvel2)
    d['qacc_euler'] = qacc
    return qpos2.astype(dt), qvel2.astype(dt), d


def integrate_pos(mm: MjxModel, qpos, qvel):
    """mj_integratePos: hinge / slide q += dt v; free joint: position += dt v_lin, quaternion rotated by
    the body-frame angular velocity over dt (math.quat_integrate) and renormalised."""
    m, dt = mm.m, mm.dtype
    if m.nq == m.nv:
        return qpos + mm.timestep * qvel
    out = qpos.copy()
    for j in range(m.njnt):
        a, i = m.jnt_qposadr[j], m.jnt_dofadr[j]
        if m.jnt_type[j] == mjc.FREE:
            out[:, a:a + 3] = qpos[:, a:a + 3] + mm.timestep * qvel[:, i:i + 3]
            w = qvel[:, i + 3:i + 6]
            wn = np.linalg.norm(w, axis=1)
            ang = wn * mm.timestep
            with np.errstate(all='ignore'):
                axis = np.where((wn > 0)[:, None], w / wn[:, None], np.array([1, 0, 0], dt))
            ql = np.concatenate([np.cos(ang * 0.5)[:, None], axis * np.sin(ang * 0.5)[:, None]], 1)
            q = quat_mul(qpos[:, a + 3:a + 7], ql)
            out[:, a + 3:a + 7] = q / np.linalg.norm(q, axis=1, keepdims=True)
        else:
            out[:, a] = qpos[:, a] + mm.timestep * qvel[:, i]
    return out
