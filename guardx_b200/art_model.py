"""Lowering of a compiled model (mjcf.Model) to the device struct of the articulated
kernels (csrc/gx_art.cuh: GxArtModel) -- what ``mjx.put_model`` (engine.py:231) is to the
reference.  Checks the STAR topology the kernels are written for: one root body with
three planar root joints, and serial chains ("legs") hanging off it, each with at most
one sphere that can touch the floor.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import mjcf

MAXLEG, MAXK, MAXB = 4, 5, 3
mjMINVAL = 1e-15
mjMINIMP, mjMAXIMP = 0.0001, 0.9999
F, I = C.c_float, C.c_int32


class GxArtBody(C.Structure):
    _fields_ = [("pos", F * 3), ("quat", F * 4), ("ipos", F * 3), ("inert", F * 6), ("mass", F), ("ndof", I)]


class GxArtDof(C.Structure):
    _fields_ = [("type", I), ("limited", I), ("act", I), ("ctrllimited", I), ("forcelimited", I), ("pad0", I),
                ("axis", F * 3), ("jpos", F * 3),
                ("damping", F), ("armature", F), ("stiffness", F), ("springref", F), ("qpos0", F),
                ("range", F * 2), ("margin", F), ("invw", F), ("k", F), ("b", F), ("imp", F * 5),
                ("gear", F), ("gain", F), ("biasv", F), ("ctrl_lo", F), ("ctrl_hi", F), ("force_lo", F), ("force_hi", F)]


class GxArtContact(C.Structure):
    _fields_ = [("on", I), ("body", I), ("pos", F * 3), ("radius", F), ("margin", F), ("mu", F), ("invw", F),
                ("k", F), ("b", F), ("imp", F * 5)]


class GxArtSeg(C.Structure):
    _fields_ = [("nb", I), ("nk", I), ("dofbody", I * MAXK), ("body", GxArtBody * MAXB), ("dof", GxArtDof * MAXK),
                ("con", GxArtContact)]


class GxArtModel(C.Structure):
    _fields_ = [("nleg", I), ("nv", I), ("nu", I), ("max_iter", I), ("root_qadr", I), ("leg_qadr", I * MAXLEG), ("root_x", I), ("root_y", I),
                ("dt", F), ("grav_z", F), ("tol_scale", F), ("root", GxArtSeg), ("leg", GxArtSeg * MAXLEG)]


def _kb(solref, solimp, timestep):
    """solref -> (k, b) with REFSAFE, solimp clipped (mjx constraint._kbi)."""
    timeconst, dampratio = float(solref[0]), float(solref[1])
    dmin, dmax, width, mid, power = (float(x) for x in solimp)
    dmin = min(max(dmin, mjMINIMP), mjMAXIMP); dmax = min(max(dmax, mjMINIMP), mjMAXIMP)
    width = max(mjMINVAL, width); mid = min(max(mid, mjMINIMP), mjMAXIMP); power = max(1.0, power)
    tc = max(timeconst, 2 * timestep)
    k = 1.0 / (dmax * dmax * tc * tc * dampratio * dampratio)
    b = 2.0 / (dmax * tc)
    if solref[0] <= 0:
        k = -solref[0] / (dmax * dmax)
    if solref[1] <= 0:
        b = -solref[1] / dmax
    return k, b, (dmin, dmax, width, mid, power)


def lower_art(m: mjcf.Model) -> tuple[GxArtModel, dict]:
    """Returns the device struct and a small description (shape = (nleg, K, NB), dof order)."""
    if m.opt['solver'] != 'newton' or m.opt['cone'] != 'pyramidal' or m.opt['integrator'] != 'euler':
        raise AssertionError('only the MuJoCo defaults (Newton, pyramidal, Euler) are built')
    r = m.body_id('robot')
    assert m.body_parentid[r] == 0, 'robot body must hang off the world'
    children = {b: [c for c in range(m.nbody) if m.body_parentid[c] == b] for b in range(m.nbody)}
    dt = float(m.opt['timestep'])
    out = GxArtModel()
    act_of_dof = {int(m.jnt_dofadr[m.actuator_trnid[a]]): a for a in range(m.nu)}

    def fill_body(dst: GxArtBody, b: int):
        dst.pos[:] = list(m.body_pos[b]); dst.quat[:] = list(m.body_quat[b]); dst.ipos[:] = list(m.body_ipos[b])
        J = m.body_inertia_full[b]
        dst.inert[:] = [J[0, 0], J[1, 1], J[2, 2], J[0, 1], J[0, 2], J[1, 2]]
        dst.mass = m.body_mass[b]
        dst.ndof = int(m.body_dofnum[b])

    def fill_dof(dst: GxArtDof, j: int):
        t = int(m.jnt_type[j])
        assert t in (mjcf.JNT_SLIDE, mjcf.JNT_HINGE)
        i = int(m.jnt_dofadr[j])
        dst.type = 0 if t == mjcf.JNT_SLIDE else 1
        dst.limited = int(m.jnt_limited[j])
        dst.axis[:] = list(m.jnt_axis[j]); dst.jpos[:] = list(m.jnt_pos[j])
        dst.damping = m.dof_damping[i]; dst.armature = m.dof_armature[i]; dst.stiffness = m.jnt_stiffness[j]
        dst.springref = m.qpos_spring[j]; dst.qpos0 = m.qpos0[int(m.jnt_qposadr[j])]
        dst.range[:] = list(m.jnt_range[j]); dst.margin = m.jnt_margin[j]; dst.invw = m.dof_invweight0[i]
        k, b, imp = _kb(m.jnt_solref[j], m.jnt_solimp[j], dt)
        dst.k, dst.b = k, b
        dst.imp[:] = list(imp)
        a = act_of_dof.get(i, -1)
        dst.act = a
        if a >= 0:
            bp = m.actuator_biasprm[a]
            assert bp[0] == 0 and bp[1] == 0, 'only motor / velocity-servo actuators'
            dst.gear = m.actuator_gear[a]; dst.gain = m.actuator_gainprm0[a]; dst.biasv = bp[2]
            dst.ctrllimited = int(m.actuator_ctrllimited[a]); dst.forcelimited = int(m.actuator_forcelimited[a])
            dst.ctrl_lo, dst.ctrl_hi = m.actuator_ctrlrange[a]
            dst.force_lo, dst.force_hi = m.actuator_forcerange[a]

    # ---- root segment
    assert m.body_dofnum[r] == 3, 'planar root (3 joints) expected'
    out.root.nb, out.root.nk = 1, 3
    fill_body(out.root.body[0], r)
    for k in range(3):
        j = int(m.body_jntadr[r]) + k
        fill_dof(out.root.dof[k], j)
        out.root.dofbody[k] = 0
        assert not m.jnt_limited[j], 'limited root joints are not built'
    out.root.con.on = 0
    out.root.con.body = -1
    out.root_qadr = int(m.body_dofadr[r])
    out.root_x = int(m.jnt_dofadr[m.jnt_id('robot_x')]) - out.root_qadr
    out.root_y = int(m.jnt_dofadr[m.jnt_id('robot_y')]) - out.root_qadr
    # ---- legs
    legs = children[r]
    assert len(legs) <= MAXLEG, f'at most {MAXLEG} chains'
    body_leg = {}
    kmax = nbmax = 0
    for l, b0 in enumerate(legs):
        seg = out.leg[l]
        chain, b = [], b0
        while True:
            chain.append(b)
            ch = children[b]
            assert len(ch) <= 1, 'legs must be serial chains (star topology)'
            if not ch:
                break
            b = ch[0]
        assert len(chain) <= MAXB, f'at most {MAXB} bodies per chain'
        seg.nb = len(chain)
        nk = 0
        first = None
        for kb_, b in enumerate(chain):
            fill_body(seg.body[kb_], b)
            body_leg[b] = (l, kb_)
            for q in range(int(m.body_jntnum[b])):
                j = int(m.body_jntadr[b]) + q
                assert nk < MAXK, f'at most {MAXK} dofs per chain'
                fill_dof(seg.dof[nk], j)
                seg.dofbody[nk] = kb_
                if first is None:
                    first = int(m.jnt_dofadr[j])
                assert int(m.jnt_dofadr[j]) == first + nk, 'chain dofs must be contiguous'
                nk += 1
        seg.nk = nk
        seg.con.on = 0
        seg.con.body = -1
        out.leg_qadr[l] = first if first is not None else 0
        kmax, nbmax = max(kmax, nk), max(nbmax, len(chain))
    out.nleg = len(legs)
    nv_robot = 3 + sum(out.leg[l].nk for l in range(out.nleg))
    out.nv, out.nu = nv_robot, m.nu
    # ---- contacts: spheres against the floor plane
    ignored = []
    for g1, g2 in m.contact_pairs:
        t1, t2 = int(m.geom_type[g1]), int(m.geom_type[g2])
        if {t1, t2} != {mjcf.GEOM_PLANE, mjcf.GEOM_SPHERE}:
            ignored.append((m.geom_name[g1], m.geom_name[g2]))   # assumption A2: never active
            continue
        gp, gs = (g1, g2) if t1 == mjcf.GEOM_PLANE else (g2, g1)
        assert m.geom_bodyid[gp] == 0 and np.allclose(m.geom_pos[gp], 0) and np.allclose(m.geom_quat[gp], [1, 0, 0, 0]), \
            'floor must be the world plane z = 0'
        bs = int(m.geom_bodyid[gs])
        if bs == r:
            ignored.append((m.geom_name[gp], m.geom_name[gs]))   # root sphere resting on the floor (Point): A2
            continue
        l, kb_ = body_leg[bs]
        con = out.leg[l].con
        assert not con.on, 'one floor contact per chain'
        assert max(m.geom_condim[gp], m.geom_condim[gs]) == 3
        fri = np.maximum(m.geom_friction[gp], m.geom_friction[gs])
        margin = max(m.geom_margin[gp], m.geom_margin[gs]) - max(m.geom_gap[gp], m.geom_gap[gs])
        solref = 0.5 * (m.geom_solref[gp] + m.geom_solref[gs])
        solimp = 0.5 * (m.geom_solimp[gp] + m.geom_solimp[gs])
        invw = m.body_invweight0[0, 0] + m.body_invweight0[bs, 0]
        invw = invw + fri[0] * fri[0] * invw
        invw = invw * 2 * fri[0] * fri[0] / float(m.opt['impratio'])
        k, b, imp = _kb(solref, solimp, dt)
        con.on, con.body = 1, kb_
        con.pos[:] = list(m.geom_pos[gs]); con.radius = m.geom_size[gs, 0]; con.margin = margin
        con.mu, con.invw, con.k, con.b = fri[0], invw, k, b
        con.imp[:] = list(imp)
    out.dt = dt
    out.grav_z = m.opt['gravity'][2]
    assert m.opt['gravity'][0] == 0 and m.opt['gravity'][1] == 0
    out.max_iter = int(m.opt['iterations'])
    out.tol_scale = float(m.opt['tolerance']) * m.stat_meaninertia * max(1, m.nv)
    info = dict(shape=(out.nleg, kmax, nbmax), nv=nv_robot, nu=m.nu, ignored_pairs=ignored,
                z_height=float(m.body_pos[r][2]))
    return out, info
