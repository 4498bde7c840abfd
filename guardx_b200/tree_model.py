"""Lowering of a compiled model (mjcf.Model) to the device struct of the general-tree kernels
(csrc/gx_tree.cuh: GxTreeModel, include/guardx_tree.h) -- the Doggo path: free-joint (or jointed)
root, hinge / slide joints, capsule / sphere geoms against the floor and against each other.
The goal / hazard bodies are lowered to ``nworld`` independent damped-spring slide dofs."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import mjcf
from .art_model import _kb

MAXB, MAXJ, MAXV, MAXQ, MAXU, MAXG, MAXPAIR, MAXCON = 16, 16, 20, 21, 16, 20, 96, 24
F, I, U = C.c_float, C.c_int32, C.c_uint32
TP_PLANE_SPHERE, TP_PLANE_CAPSULE, TP_CAPSULE_CAPSULE = 0, 1, 2


class GxTreeBody(C.Structure):
    _fields_ = [("parent", I), ("jntadr", I), ("jntnum", I), ("dofadr", I), ("dofnum", I), ("dofmask", U),
                ("pos", F * 3), ("quat", F * 4), ("ipos", F * 3), ("inert", F * 6), ("mass", F)]


class GxTreeJoint(C.Structure):
    _fields_ = [("type", I), ("body", I), ("qposadr", I), ("dofadr", I), ("limited", I), ("pad", I),
                ("axis", F * 3), ("pos", F * 3), ("qpos0", F), ("stiffness", F), ("springref", F), ("range", F * 2),
                ("margin", F), ("invw", F), ("k", F), ("b", F), ("imp", F * 5)]


class GxTreeDof(C.Structure):
    _fields_ = [("body", I), ("parent", I), ("act", I), ("ctrllimited", I), ("forcelimited", I), ("pad", I),
                ("damping", F), ("armature", F), ("gear", F), ("gain", F), ("biasv", F), ("ctrl_lo", F), ("ctrl_hi", F),
                ("force_lo", F), ("force_hi", F)]


class GxTreeGeom(C.Structure):
    _fields_ = [("type", I), ("body", I), ("pos", F * 3), ("mat", F * 9), ("radius", F), ("halflen", F)]


class GxTreePair(C.Structure):
    _fields_ = [("type", I), ("g1", I), ("g2", I), ("pad", I), ("margin", F), ("mu", F), ("invw", F), ("k", F), ("b", F),
                ("imp", F * 5)]


class GxTreeModel(C.Structure):
    _fields_ = [("nb", I), ("njnt", I), ("nq", I), ("nv", I), ("nu", I), ("ngeom", I), ("npair", I), ("max_iter", I),
                ("root_free", I), ("nworld", I), ("root_xadr", I), ("root_yadr", I), ("dt", F), ("grav_z", F), ("tol_scale", F),
                ("world_mass", F * 2), ("world_damping", F), ("world_stiffness", F),
                ("body", GxTreeBody * MAXB), ("jnt", GxTreeJoint * MAXJ), ("dof", GxTreeDof * MAXV),
                ("geom", GxTreeGeom * MAXG), ("pair", GxTreePair * MAXPAIR)]


def lower_tree(m: mjcf.Model) -> tuple[GxTreeModel, dict]:
    if m.opt['solver'] != 'newton' or m.opt['cone'] != 'pyramidal' or m.opt['integrator'] != 'euler':
        raise AssertionError('only the MuJoCo defaults (Newton, pyramidal, Euler) are built')
    r = m.body_id('robot')
    assert m.body_parentid[r] == 0 and r == 1, 'robot must be the first body under the world'
    # robot bodies: the subtree of `robot`, in model order (parents before children)
    robot = [r]
    for b in range(r + 1, m.nbody):
        if m.body_parentid[b] in robot:
            robot.append(b)
    bmap = {b: k for k, b in enumerate(robot)}
    jnts = [j for j in range(m.njnt) if m.jnt_bodyid[j] in bmap]
    nq = sum(7 if m.jnt_type[j] == mjcf.JNT_FREE else 1 for j in jnts)
    nv = sum(6 if m.jnt_type[j] == mjcf.JNT_FREE else 1 for j in jnts)
    assert jnts == list(range(len(jnts))) and m.jnt_qposadr[jnts[0]] == 0, 'robot joints must come first'
    assert len(robot) <= MAXB and len(jnts) <= MAXJ and nv <= MAXV and nq <= MAXQ and m.nu <= MAXU
    out = GxTreeModel()
    dt = float(m.opt['timestep'])
    out.nb, out.njnt, out.nq, out.nv, out.nu = len(robot), len(jnts), nq, nv, m.nu
    out.max_iter = int(m.opt['iterations'])
    out.dt, out.grav_z = dt, m.opt['gravity'][2]
    out.tol_scale = float(m.opt['tolerance']) * m.stat_meaninertia * max(1, m.nv)
    out.root_free = int(m.jnt_type[jnts[0]] == mjcf.JNT_FREE)
    if out.root_free:      # joint literally named 'robot': x, y, z_height, quaternion
        out.root_xadr, out.root_yadr = int(m.jnt_qposadr[jnts[0]]), int(m.jnt_qposadr[jnts[0]]) + 1
    else:
        out.root_xadr, out.root_yadr = int(m.jnt_qposadr[m.jnt_id('robot_x')]), int(m.jnt_qposadr[m.jnt_id('robot_y')])
    act_of_dof = {int(m.jnt_dofadr[m.actuator_trnid[a]]): a for a in range(m.nu)}
    for k, b in enumerate(robot):
        B = out.body[k]
        B.parent = bmap[m.body_parentid[b]] if m.body_parentid[b] in bmap else -1
        B.jntadr, B.jntnum = int(m.body_jntadr[b]) if m.body_jntnum[b] else 0, int(m.body_jntnum[b])
        B.dofadr, B.dofnum = int(m.body_dofadr[b]) if m.body_dofnum[b] else 0, int(m.body_dofnum[b])
        mask = 0
        a = b
        while a in bmap:
            for i in range(int(m.body_dofnum[a])):
                mask |= 1 << (int(m.body_dofadr[a]) + i)
            a = int(m.body_parentid[a])
        B.dofmask = mask
        B.pos[:] = list(m.body_pos[b]); B.quat[:] = list(m.body_quat[b]); B.ipos[:] = list(m.body_ipos[b])
        Jm = m.body_inertia_full[b]
        B.inert[:] = [Jm[0, 0], Jm[1, 1], Jm[2, 2], Jm[0, 1], Jm[0, 2], Jm[1, 2]]
        B.mass = m.body_mass[b]
    for j in jnts:
        Jn = out.jnt[j]
        t = int(m.jnt_type[j])
        Jn.type, Jn.body = t, bmap[int(m.jnt_bodyid[j])]
        Jn.qposadr, Jn.dofadr, Jn.limited = int(m.jnt_qposadr[j]), int(m.jnt_dofadr[j]), int(m.jnt_limited[j])
        Jn.axis[:] = list(m.jnt_axis[j]); Jn.pos[:] = list(m.jnt_pos[j])
        a = int(m.jnt_qposadr[j])
        Jn.qpos0 = 0.0 if t == mjcf.JNT_FREE else m.qpos0[a]
        Jn.stiffness = m.jnt_stiffness[j]
        Jn.springref = 0.0 if t == mjcf.JNT_FREE else m.qpos_spring[a]
        Jn.range[:] = list(m.jnt_range[j]); Jn.margin = m.jnt_margin[j]
        Jn.invw = m.dof_invweight0[int(m.jnt_dofadr[j])]
        k_, b_, imp = _kb(m.jnt_solref[j], m.jnt_solimp[j], dt)
        Jn.k, Jn.b = k_, b_
        Jn.imp[:] = list(imp)
    for i in range(nv):
        D = out.dof[i]
        D.body, D.parent = bmap[int(m.dof_bodyid[i])], int(m.dof_parentid[i])
        D.damping, D.armature = m.dof_damping[i], m.dof_armature[i]
        a = act_of_dof.get(i, -1)
        D.act = a
        if a >= 0:
            bp = m.actuator_biasprm[a]
            assert bp[0] == 0 and bp[1] == 0, 'only motor / velocity-servo actuators'
            D.gear, D.gain, D.biasv = m.actuator_gear[a], m.actuator_gainprm0[a], bp[2]
            D.ctrllimited, D.forcelimited = int(m.actuator_ctrllimited[a]), int(m.actuator_forcelimited[a])
            D.ctrl_lo, D.ctrl_hi = m.actuator_ctrlrange[a]
            D.force_lo, D.force_hi = m.actuator_forcerange[a]
    # ---- colliding geoms and candidate pairs
    gmap = {}

    def geom(g):
        if g in gmap:
            return gmap[g]
        k = len(gmap)
        assert k < MAXG
        G = out.geom[k]
        t = int(m.geom_type[g])
        assert t in (mjcf.GEOM_SPHERE, mjcf.GEOM_CAPSULE), 'sphere / capsule geoms only'
        G.type, G.body = t, bmap[int(m.geom_bodyid[g])]
        G.pos[:] = list(m.geom_pos[g]); G.mat[:] = list(mjcf.quat2mat(m.geom_quat[g]).reshape(-1))
        G.radius, G.halflen = m.geom_size[g, 0], (m.geom_size[g, 1] if t == mjcf.GEOM_CAPSULE else 0.0)
        gmap[g] = k
        return k
    npair = 0
    for g1, g2 in m.contact_pairs:
        t1, t2 = int(m.geom_type[g1]), int(m.geom_type[g2])
        if t2 == mjcf.GEOM_PLANE:
            g1, g2, t1, t2 = g2, g1, t2, t1
        P = out.pair[npair]
        if t1 == mjcf.GEOM_PLANE:
            assert m.geom_bodyid[g1] == 0 and np.allclose(m.geom_pos[g1], 0) and np.allclose(m.geom_quat[g1], [1, 0, 0, 0])
            P.type = TP_PLANE_SPHERE if t2 == mjcf.GEOM_SPHERE else TP_PLANE_CAPSULE
            P.g1, P.g2 = -1, geom(g2)
        else:
            assert t1 == mjcf.GEOM_CAPSULE and t2 == mjcf.GEOM_CAPSULE, 'capsule-capsule is the only robot-robot pair built'
            P.type, P.g1, P.g2 = TP_CAPSULE_CAPSULE, geom(g1), geom(g2)
        assert max(m.geom_condim[g1], m.geom_condim[g2]) == 3
        fri = np.maximum(m.geom_friction[g1], m.geom_friction[g2])
        P.margin = max(m.geom_margin[g1], m.geom_margin[g2]) - max(m.geom_gap[g1], m.geom_gap[g2])
        solref = 0.5 * (m.geom_solref[g1] + m.geom_solref[g2])
        solimp = 0.5 * (m.geom_solimp[g1] + m.geom_solimp[g2])
        invw = m.body_invweight0[m.geom_bodyid[g1], 0] + m.body_invweight0[m.geom_bodyid[g2], 0]
        invw = invw + fri[0] * fri[0] * invw
        invw = invw * 2 * fri[0] * fri[0] / float(m.opt['impratio'])
        k_, b_, imp = _kb(solref, solimp, dt)
        P.mu, P.invw, P.k, P.b = fri[0], invw, k_, b_
        P.imp[:] = list(imp)
        npair += 1
        assert npair <= MAXPAIR
    out.ngeom, out.npair = len(gmap), npair
    # ---- goal / hazard slide dofs (world.py:316-323): mass + armature, damping, stiffness
    gx_, hz = m.jnt_id('goal_x'), [j for j in range(m.njnt) if (m.jnt_name[j] or '').startswith('hazard')]
    out.nworld = 2 + len(hz)
    out.world_mass[0] = m.body_mass[m.body_id('goal')] + m.dof_armature[int(m.jnt_dofadr[gx_])]
    out.world_mass[1] = (m.body_mass[int(m.jnt_bodyid[hz[0]])] + m.dof_armature[int(m.jnt_dofadr[hz[0]])]) if hz else 1.0
    out.world_damping = m.dof_damping[int(m.jnt_dofadr[gx_])]
    out.world_stiffness = m.jnt_stiffness[gx_]
    info = dict(nq=nq, nv=nv, nu=m.nu, nb=len(robot), npair=npair, ngeom=len(gmap), z_height=float(m.body_pos[r][2]))
    return out, info
