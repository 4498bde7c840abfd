"""CPU oracle of the config-3 extension (Button / Push / Goal x Point / Car among hazards, vases, pillars, gremlins).

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED BY CONSTRUCTION: the reference has no implementation of these tasks, objects
or of the Car, so there is nothing to pin against.  This file is an INDEPENDENT restatement
-- numpy, float64 arithmetic, numpy's own arctan2 / exp / sin / cos -- of the specification in DESIGN.md section 8, which
the CUDA kernels (guardx_b200/csrc/gx_task.cuh) implement in fp32 with their own elementary functions; the two are
compared under a stated tolerance (floats 2e-5, integer outputs exact away from thresholds).  It imports nothing from
the product; its parameters arrive as a plain dict built by the test from the C ABI's config struct.
State rows / layout rows have the order include/guardx_task.h and csrc/gx_task.cuh document.
"""
from __future__ import annotations

import numpy as np

GOAL, BOX, BUTTON, HAZARD, VASE, PILLAR, GREMLIN, ROBOT = range(8)
SLOT_OBJ = 16


class TaskOracle:
    def __init__(self, P: dict):
        self.P = P
        self.count = list(P['count'])
        self.first = np.concatenate([[0], np.cumsum(self.count[:ROBOT])]).astype(int)
        self.nobj = int(sum(self.count[:ROBOT]))
        self.ng = self.count[GREMLIN]
        self.nb = P['lidar_num_bins']
        self.n = P['env_num']
        self.state_dim = SLOT_OBJ + 2 * self.nobj + self.ng
        self.S = np.zeros((self.n, self.state_dim))
        # observation layout: sorted keys
        nu = 2 if P['robot'] == 1 else 3
        comps = [('box_compass', 2, self.count[BOX] > 0), ('box_lidar', self.nb, self.count[BOX] > 0),
                 ('buttons_lidar', self.nb, self.count[BUTTON] > 0), ('ctrl', nu, True), ('goal_compass', 2, True),
                 ('goal_lidar', self.nb, True), ('gremlins_lidar', self.nb, self.count[GREMLIN] > 0),
                 ('hazards_lidar', self.nb, self.count[HAZARD] > 0), ('pillars_lidar', self.nb, self.count[PILLAR] > 0),
                 ('qpos', 3, True), ('qvel', 3, True), ('vases_lidar', self.nb, self.count[VASE] > 0)]
        self.off, o = {}, 0
        for bit, (name, w, exists) in enumerate(comps):
            if exists and (P['obs_mask'] >> bit) & 1:
                self.off[name] = (o, o + w)
                o += w
        self.obs_dim = o

    # ---- state access ---------------------------------------------------------------------------------------------
    def set_state(self, rows):
        self.S = np.asarray(rows, np.float64).copy()

    def get_state(self):
        return self.S.astype(np.float32)

    def obj(self, cls, k):
        o = self.first[cls] + k
        return self.S[:, SLOT_OBJ + 2 * o:SLOT_OBJ + 2 * o + 2]

    def gremlin_xy(self, k, time):
        ph = self.S[:, SLOT_OBJ + 2 * self.nobj + k] + time
        return self.obj(GREMLIN, k) + self.P['gremlins_travel'] * np.stack([np.cos(ph), np.sin(ph)], 1)

    def target(self):
        if self.P['task'] == 1:
            b = np.clip(self.S[:, 14].astype(int), 0, self.count[BUTTON] - 1)
            return np.stack([self.obj(BUTTON, k)[i] for i, k in enumerate(b)])
        return self.obj(GOAL, 0).copy()

    # ---- lidar (engine.py:846-900) --------------------------------------------------------------------------------
    def lidar(self, pos_list, P0, c, s):
        nb = self.nb
        out = np.zeros((self.n, nb))
        bins_first = None
        for pos in pos_list:
            d = pos - P0
            ex, ey = d[:, 0] * c + d[:, 1] * s, -d[:, 0] * s + d[:, 1] * c
            dist = np.hypot(ex, ey)
            ang = np.arctan2(ey, ex) % (2 * np.pi)
            bsz = 2 * np.pi / nb
            b = (ang / bsz).astype(int)
            if self.P['lidar_max_dist'] > 0:
                sensor = np.maximum(0, self.P['lidar_max_dist'] - dist) / self.P['lidar_max_dist']
            else:
                sensor = np.exp(-self.P['lidar_exp_gain'] * dist)
            if bins_first is None:
                bins_first = b.copy()
            idx = np.arange(self.n)
            ok = b < nb
            out[idx[ok], b[ok]] = np.maximum(out[idx[ok], b[ok]], sensor[ok])
            if self.P['lidar_alias']:
                alias = (ang - b * bsz) / bsz
                bp, bm = (b + 1) % nb, (b - 1) % nb
                out[idx, bp] = np.maximum(out[idx, bp], alias * sensor)
                out[idx, bm] = np.maximum(out[idx, bm], (1 - alias) * sensor)
        return out, bins_first

    # ---- robots ---------------------------------------------------------------------------------------------------
    def point_step(self, a):
        P, S = self.P, self.S
        q, v = S[:, 0:3].copy(), S[:, 3:6].copy()
        ctrl = np.stack([S[:, 6] * a[:, 0], S[:, 7] * a[:, 0], a[:, 1]], 1)       # lagged heading (engine.py:676-680)
        th = q[:, 2]
        c, s = np.cos(th), np.sin(th)
        m, mc, I = P['pt_mass'], P['pt_mass'] * P['pt_com_x'], P['pt_inertia_zz']
        damp, gear = np.array(P['pt_damping']), np.array(P['pt_gear'])
        u = np.clip(ctrl, P['pt_ctrl'][0], P['pt_ctrl'][1])
        force = np.clip(P['pt_gain'] * u + P['pt_bias2'] * (gear * v), P['pt_force'][0], P['pt_force'][1]) * gear
        bias = np.stack([-mc * c * v[:, 2] ** 2, -mc * s * v[:, 2] ** 2, np.zeros(self.n)], 1)
        f = -damp * v - bias + force
        vn = np.zeros_like(v)
        dt = P['pt_timestep']
        for i in range(self.n):
            M = np.array([[m, 0, -mc * s[i]], [0, m, mc * c[i]], [-mc * s[i], mc * c[i], I]])
            vn[i] = v[i] + dt * np.linalg.solve(M + dt * np.diag(damp), f[i])
        qn = q + dt * vn
        return ctrl, q[:, :2].copy(), c, s, qn, vn

    def car_step(self, a):
        P, S = self.P, self.S
        q, v = S[:, 0:3].copy(), S[:, 3:6].copy()
        ctrl = np.clip(a, -1, 1)
        c, s = np.cos(q[:, 2]), np.sin(q[:, 2])
        tl, tr = P['car_gear'] * ctrl[:, 0], P['car_gear'] * ctrl[:, 1]
        F, N = (tl + tr) / P['car_wheel_r'], P['car_half_track'] * (tr - tl) / P['car_wheel_r']
        dt = P['car_timestep']
        u = v[:, 0] * c + v[:, 1] * s
        un = (P['car_mass'] * u + dt * F) / (P['car_mass'] + dt * P['car_damp_lin'])
        wn = (P['car_inertia'] * v[:, 2] + dt * N) / (P['car_inertia'] + dt * P['car_damp_ang'])
        vn = np.stack([un * c, un * s, wn], 1)
        return ctrl, q[:, :2].copy(), c, s, q + dt * vn, vn

    # ---- observation + cost at a pose snapshot --------------------------------------------------------------------
    def observe(self, P0, c, s, time, ctrl, q, v):
        P = self.P
        obs = np.zeros((self.n, self.obs_dim))
        tgt = self.target()
        rr = P['size'][ROBOT]

        def put(name, val):
            if name in self.off:
                a, b = self.off[name]
                obs[:, a:b] = val

        put('ctrl', ctrl)
        d = tgt - P0
        put('goal_compass', np.stack([d[:, 0] * c + d[:, 1] * s, -d[:, 0] * s + d[:, 1] * c], 1))
        gl, goal_bin = self.lidar([tgt], P0, c, s)
        put('goal_lidar', gl)
        if self.count[BOX]:
            bx = self.obj(BOX, 0)
            d = bx - P0
            put('box_compass', np.stack([d[:, 0] * c + d[:, 1] * s, -d[:, 0] * s + d[:, 1] * c], 1))
            put('box_lidar', self.lidar([bx], P0, c, s)[0])
        for name, cls in (('buttons_lidar', BUTTON), ('hazards_lidar', HAZARD), ('pillars_lidar', PILLAR), ('vases_lidar', VASE)):
            if name in self.off:
                put(name, self.lidar([self.obj(cls, k) for k in range(self.count[cls])], P0, c, s)[0])
        if 'gremlins_lidar' in self.off:
            put('gremlins_lidar', self.lidar([self.gremlin_xy(k, time) for k in range(self.ng)], P0, c, s)[0])
        put('qpos', q)
        put('qvel', v)
        dist_task = np.hypot(*(tgt - P0).T)
        dist_box = np.zeros(self.n)
        if P['task'] == 2:
            dist_box = np.hypot(*(self.obj(BOX, 0) - P0).T)
            dist_task = np.hypot(*(tgt - self.obj(BOX, 0)).T)
        cost = np.zeros(self.n); viol = np.zeros(self.n, int); contacts = np.zeros(self.n, int)
        near = {}        # distance - threshold per indicator test, for the "away from thresholds" filter of the tests
        margins = []
        gb = self.S[:, 14].astype(int)
        for k in range(self.count[BUTTON]):
            dd = np.hypot(*(self.obj(BUTTON, k) - P0).T) - (rr + P['size'][BUTTON])
            hit = (dd < 0) & ~((P['task'] == 1) & (gb == k))
            cost += P['buttons_cost'] * hit; contacts += hit; margins.append(np.where((P['task'] == 1) & (gb == k), 1.0, dd))
        for k in range(self.count[HAZARD]):
            dist = np.hypot(*(self.obj(HAZARD, k) - P0).T)
            term = P['size'][HAZARD] - np.minimum(dist, P['size'][HAZARD])
            cost += P['hazards_cost_scale'] * term; viol += term > 0; margins.append(dist - P['size'][HAZARD])
        for cls, key in ((VASE, 'vases_contact_cost'), (PILLAR, 'pillars_cost')):
            for k in range(self.count[cls]):
                dd = np.hypot(*(self.obj(cls, k) - P0).T) - (rr + P['size'][cls])
                cost += P[key] * (dd < 0); contacts += dd < 0; margins.append(dd)
        for k in range(self.ng):
            dist = np.hypot(*(self.gremlin_xy(k, time) - P0).T)
            dd = dist - (rr + P['size'][GREMLIN])
            cost += P['gremlins_contact_cost'] * (dd < 0); contacts += dd < 0; margins.append(dd)
            gap = dist - P['size'][GREMLIN]
            cost += np.where(gap < P['gremlins_dist_threshold'], P['gremlins_dist_cost'] * (P['gremlins_dist_threshold'] - gap), 0.0)
            margins.append(gap - P['gremlins_dist_threshold'])
        margin = np.min(np.abs(np.stack(margins, 1)), 1) if margins else np.full(self.n, 1.0)
        return obs, dist_task, dist_box, cost, viol, contacts, goal_bin, margin

    # ---- Engine.step (no auto-reset) ------------------------------------------------------------------------------
    def step(self, action):
        P, S = self.P, self.S
        a = np.asarray(action, np.float64)
        last_dist, last_box, last_done, steps = S[:, 8].copy(), S[:, 9].copy(), S[:, 10].copy(), S[:, 11].copy()
        dt = P['car_timestep'] if P['robot'] == 1 else P['pt_timestep']
        ctrl, P0, c, s, qn, vn = (self.car_step if P['robot'] == 1 else self.point_step)(a)
        time = steps * dt
        # contacts on the post-step state: pillars stop the robot ...
        rr = P['size'][ROBOT]
        q2, v2 = qn.copy(), vn.copy()
        for k in range(self.count[PILLAR]):
            o = self.obj(PILLAR, k)
            d = q2[:, :2] - o
            dist = np.hypot(*d.T)
            lim = rr + P['size'][PILLAR]
            hit = (dist < lim) & (dist > 0)
            nrm = d / np.where(dist > 0, dist, 1)[:, None]
            q2[hit, :2] = (o + nrm * lim)[hit]
            vnorm = (v2[:, :2] * nrm).sum(1)
            fix = hit & (vnorm < 0)
            v2[fix, :2] = (v2[:, :2] - vnorm[:, None] * nrm)[fix]
        obs, dist_task, dist_box, cost, viol, contacts, goal_bin, margin = self.observe(P0, c, s, time, ctrl, q2, v2)
        # ... and the robot pushes the box and the vases out of its way (after the observation saw them where they were)
        for cls in (BOX, VASE):
            for k in range(self.count[cls]):
                o = self.obj(cls, k)
                d = o - q2[:, :2]
                dist = np.hypot(*d.T)
                lim = rr + P['size'][cls]
                hit = (dist < lim) & (dist > 0)
                nrm = d / np.where(dist > 0, dist, 1)[:, None]
                o[hit] = (q2[:, :2] + nrm * lim)[hit]
        last = np.where(last_done > 0, dist_task, last_dist)
        d_dist = last - dist_task
        if P['task'] == 2:
            lastb = np.where(last_done > 0, dist_box, last_box)
            reward = (lastb - dist_box) * P['reward_box_dist'] + d_dist * P['reward_box_goal']
            done = dist_task < P['size'][GOAL]
            thr = P['size'][GOAL]
        elif P['task'] == 1:
            reward = d_dist * P['reward_distance']
            thr = rr + P['size'][BUTTON]
            done = dist_task < thr
        else:
            reward = d_dist * P['reward_distance']
            thr = P['size'][GOAL]
            done = dist_task < thr
        flags = np.where(done, 4, 0)
        jump = np.abs(d_dist) > 1
        done = done | jump; reward = np.where(jump, 0, reward); flags |= np.where(jump, 8, 0)
        bad = ~np.isfinite(obs).all(1)
        done = done | bad; reward = np.where(bad, 0, reward); flags |= np.where(bad, 1, 0)
        to = steps > P['num_steps']
        done = done | to; flags |= np.where(to, 2, 0)
        S[:, 0:3], S[:, 3:6] = q2, v2
        S[:, 6], S[:, 7] = c, s
        S[:, 8], S[:, 9] = dist_task, dist_box
        S[:, 10] = done
        S[:, 11] = np.where(done, 0, steps + 1)
        info = dict(cost=cost, violations=viol, contacts=contacts, goal_bin=goal_bin, flags=flags, margin=margin,
                    dist=dist_task, d_dist=d_dist, success_threshold=thr)
        return obs, reward, done.astype(np.float64), info

    def reset_obs(self):
        """observation of freshly placed envs (pose = state, theta as stored, qvel = ctrl = 0 in the row, time 0)"""
        S = self.S
        c, s = np.cos(S[:, 2]), np.sin(S[:, 2])
        nu = 2 if self.P['robot'] == 1 else 3
        return self.observe(S[:, 0:2].copy(), c, s, np.zeros(self.n), np.zeros((self.n, nu)), S[:, 0:3], S[:, 3:6])
