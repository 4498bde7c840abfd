"""CPU oracle of the batched env for articulated robots (Ant / Walker / Swimmer).

TEST INFRASTRUCTURE ONLY.  Restates ``Engine.reset / step / reset_done`` and
``mjx_reset / mjx_step / mjx_reset_done`` (engine.py:454-505, 623-731) around
  * ``oracle/mjx_numpy.py`` for ``mjx.forward`` / ``mjx.step`` (PARITY UNPINNED, see there), and
  * the single-env pieces of ``oracle/gx_oracle.c`` for ``obs_lidar`` / ``obs_compass`` /
    ``cost`` (engine.py:804-900), which ARE pinned against the reference's own source
    (tests/golden/refsrc_*.npz).
State is the full ``mjx.Data`` the reference carries: qpos / qvel over robot + goal + hazard
joints, ctrl, qacc_warmstart, plus Engine._done / _steps and the previous forward's distance.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import oracle_config as gxconfig     # the oracle's own defaults / lowering / model compile
from . import mjx_numpy as mx
from . import oracle as _o


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class ArtOracleEngine:
    def __init__(self, config: dict, dtype=np.float32):
        self.L = _o.lib()
        self.config = gxconfig.parse(config)
        self.model = gxconfig.world_model(self.config)
        self.gx_config = gxconfig.lower(self.config, 0, 0)
        self.mm = mx.MjxModel(self.model, dtype)
        self.dtype = dtype
        m = self.model
        self.n = int(self.config['env_num'])
        self.nh = int(self.config['hazards_num'])
        self.nb = int(self.config['lidar_num_bins'])
        self.substeps = int(self.config['physics_steps_per_control_step'])
        self.robot_body = m.body_id('robot')
        self.nq_robot = m.nq - 2 - 2 * self.nh   # robot.nq / robot.nv (world.py:436-437): the robot's own joints
        self.nv_robot = m.nv - 2 - 2 * self.nh
        self.free_root = bool((m.jnt_type == 0).any())          # Doggo: free joint named 'robot'
        self.nu = m.nu
        self.obs_dim = self.nu + 2 + 2 * self.nb + self.nq_robot + self.nv_robot
        assert self.config['observe_goal_lidar'] and self.config['observe_goal_comp'] and self.config['observe_hazards']
        assert self.config['observe_qpos'] and self.config['observe_qvel'] and self.config['observe_ctrl']
        self.with_vel, self.with_acc = bool(self.config['observe_vel']), bool(self.config['observe_acc'])
        self.obs_dim += 2 * (self.with_vel + self.with_acc)
        # joint_name2qpos_id (engine.py:310-316)
        self.qadr = {m.jnt_name[j]: int(m.jnt_qposadr[j]) for j in range(m.njnt) if m.jnt_name[j]}
        n = self.n
        self.qpos = np.zeros((n, m.nq), dtype); self.qvel = np.zeros((n, m.nv), dtype)
        self.ctrl = np.zeros((n, m.nu), dtype); self.warm = np.zeros((n, m.nv), dtype)
        self.done = np.zeros(n, np.float32); self.steps = np.zeros(n, np.float32)
        self.prev_dist = np.zeros(n, np.float32)
        self._obs = None
        self.last_data = None
        # history read by ego_vel_acc (engine.py:902-929): xpos of Engine._data / _last_data, _last_done, and how
        # many of _last_done / _last_last_done are not None yet (update_data, engine.py:426-431)
        self.hx_data = np.zeros((n, 3), np.float32); self.hx_last = np.zeros((n, 3), np.float32)
        self.h_last_done = np.zeros(n, np.float32); self.h_depth = 0

    # layout rows: goal xy, robot xy, hazard0 xy, ... (same as guardx.h)
    def layout2qpos(self, layout):
        q = np.zeros((layout.shape[0], self.model.nq), self.dtype)
        q[:, self.qadr['goal_x']] = layout[:, 0]; q[:, self.qadr['goal_y']] = layout[:, 1]
        if 'robot' in self.qadr:                 # free joint: x, y, z_height, unit quaternion (engine.py:626-631)
            a = self.qadr['robot']
            q[:, a] = layout[:, 2]; q[:, a + 1] = layout[:, 3]
            q[:, a + 2] = self.model.body_pos[self.robot_body][2]; q[:, a + 3] = 1.0
        else:
            q[:, self.qadr['robot_x']] = layout[:, 2]; q[:, self.qadr['robot_y']] = layout[:, 3]
        for h in range(self.nh):
            q[:, self.qadr[f'hazard{h}_x']] = layout[:, 4 + 2 * h]; q[:, self.qadr[f'hazard{h}_y']] = layout[:, 5 + 2 * h]
        return q

    def _geometry(self, d, qpos):
        """robot xpos/xmat of the forward pass + goal / hazard positions (their slide qpos)."""
        xpos = d['xpos'][:, self.robot_body].astype(np.float32)
        xmat = d['xmat'][:, self.robot_body].astype(np.float32).reshape(-1, 9)
        goal = np.stack([qpos[:, self.qadr['goal_x']], qpos[:, self.qadr['goal_y']]], 1).astype(np.float32)
        hz = np.stack([np.stack([qpos[:, self.qadr[f'hazard{h}_x']], qpos[:, self.qadr[f'hazard{h}_y']]], 1)
                       for h in range(self.nh)], 1).astype(np.float32)
        return xpos, xmat, goal, hz

    def _obs_rows(self, d, qpos_geom, qpos, qvel, ctrl, vel=None, acc=None):
        """Engine.obs (engine.py:738-778), sorted key order:
        [acc] | ctrl | goal_compass | goal_lidar | hazards_lidar | qpos | qvel | [vel]."""
        n = qpos.shape[0]
        xpos, xmat, goal, hz = self._geometry(d, qpos_geom)
        obs = np.zeros((n, self.obs_dim), np.float32)
        cost = np.zeros(n, np.float32)
        gb = np.zeros((n, 1), np.int32); hb = np.zeros((n, self.nh), np.int32)
        o = 0
        if self.with_acc:
            if acc is not None:
                obs[:, 0:2] = acc
            o = 2
        if self.with_vel and vel is not None:
            obs[:, -2:] = vel
        obs[:, o:o + self.nu] = ctrl; o += self.nu
        comp = np.zeros(2, np.float32); lid = np.zeros(self.nb, np.float32); c1 = np.zeros(1, np.float32)
        cfg = C.byref(self.gx_config)
        for i in range(n):
            xp, xm = np.ascontiguousarray(xpos[i]), np.ascontiguousarray(xmat[i])
            g, h = np.ascontiguousarray(goal[i]), np.ascontiguousarray(hz[i])
            compass, lidar = (self.L.gxo_piece_compass3, self.L.gxo_piece_lidar3) if self.free_root else (self.L.gxo_piece_compass, self.L.gxo_piece_lidar)
            compass(_p(xp), _p(xm), _p(g), _p(comp))
            obs[i, o:o + 2] = comp
            b = np.zeros(1, np.int32)
            lidar(cfg, _p(xp), _p(xm), _p(g), 1, _p(lid), _p(b))
            obs[i, o + 2:o + 2 + self.nb] = lid; gb[i] = b
            b = np.zeros(self.nh, np.int32)
            lidar(cfg, _p(xp), _p(xm), _p(h), self.nh, _p(lid), _p(b))
            obs[i, o + 2 + self.nb:o + 2 + 2 * self.nb] = lid; hb[i] = b
            rxy = np.ascontiguousarray(xp[:2])
            self.L.gxo_piece_cost(cfg, _p(rxy), _p(h), _p(c1))
            cost[i] = c1[0]
        o += 2 + 2 * self.nb
        obs[:, o:o + self.nq_robot] = qpos[:, :self.nq_robot]; o += self.nq_robot
        obs[:, o:o + self.nv_robot] = qvel[:, :self.nv_robot]
        dist = np.sqrt(((goal - xpos[:, :2]) ** 2).sum(1, dtype=np.float32)).astype(np.float32)
        return obs, cost, dist, dict(goal_bin=gb, hazard_bins=hb, xpos=xpos, xmat=xmat)

    # ------------------------------------------------------------------ API
    def reset(self, layout):
        """Engine.reset + mjx_reset (engine.py:454-467, 644-657) with the layouts injected."""
        layout = np.asarray(layout, np.float32)
        self.qpos = self.layout2qpos(layout)
        self.qvel[:] = 0; self.ctrl[:] = 0
        d = mx.forward(self.mm, self.qpos, self.qvel, self.ctrl, np.zeros_like(self.warm))   # pristine mjx_data
        self.warm = d['qacc'].copy()
        obs, _, dist, _ = self._obs_rows(d, self.qpos, self.qpos, self.qvel, self.ctrl)
        self.prev_dist = dist
        self.steps[:] = 0
        self._obs = obs
        self.last_data = d
        self.hx_data = d['xpos'][:, self.robot_body, :3].astype(np.float32)   # Engine._data <- reset data; _last_data stays
        return obs

    def step(self, action):
        """Engine.step + mjx_step (engine.py:469-495, 659-700)."""
        self.ctrl = np.asarray(action, self.dtype).copy()          # non-Point robots: ctrl = action (engine.py:673)
        qpos_pre = self.qpos
        for _ in range(self.substeps):
            qpos_pre = self.qpos
            self.qpos, self.qvel, d = mx.step(self.mm, self.qpos, self.qvel, self.ctrl, self.warm)
            self.warm = d['qacc'].copy()
        vel = acc = None
        xp = d['xpos'][:, self.robot_body, :3].astype(np.float32)
        if (self.with_vel or self.with_acc) and self.free_root:
            vel, acc = self._vel_acc3(d['xmat'][:, self.robot_body].astype(np.float32).reshape(-1, 3, 3), xp)
        elif self.with_vel or self.with_acc:
            # update_data: _last_last_data <- _last_data <- _data, same for done (engine.py:426-430)
            vel = np.zeros((self.n, 2), np.float32); acc = np.zeros((self.n, 2), np.float32)
            xm = d['xmat'][:, self.robot_body].astype(np.float32).reshape(-1, 9)
            cfg = C.byref(self.gx_config)
            for i in range(self.n):
                cur, la, ll = (np.ascontiguousarray(a[:2]) for a in (xp[i], self.hx_data[i], self.hx_last[i]))
                v2, a2 = np.zeros(2, np.float32), np.zeros(2, np.float32)
                self.L.gxo_piece_vel_acc(cfg, _p(np.ascontiguousarray(xm[i])), _p(cur), _p(la), _p(ll), C.c_int(self.h_depth),
                                         C.c_float(float(self.done[i])), C.c_float(float(self.h_last_done[i])), _p(v2), _p(a2))
                vel[i], acc[i] = v2, a2
        self.hx_last, self.hx_data = self.hx_data, xp.copy()
        self.h_last_done = self.done.copy()
        self.h_depth = min(self.h_depth + 1, 2)
        obs, cost, dist, extra = self._obs_rows(d, qpos_pre, self.qpos, self.qvel, self.ctrl, vel, acc)
        last = np.where(self.done > 0, dist, self.prev_dist).astype(np.float32)
        d_dist = (last - dist).astype(np.float32)
        reward = (d_dist * np.float32(self.config['reward_distance'])).astype(np.float32)
        done = (dist < np.float32(self.config['goal_size'])).astype(np.float32)
        jump = np.abs(d_dist) > 1
        done[jump] = 1; reward[jump] = 0
        bad = ~np.isfinite(obs).all(1)
        reward[bad] = 0; done[bad] = 1
        done[self.steps > np.float32(self.config['num_steps'])] = 1
        self.steps = np.where(done > 0, 0, self.steps + 1).astype(np.float32)
        self.done = done
        self.prev_dist = dist
        self._obs = obs
        self.last_data = d
        info = dict(cost=cost, qacc=d['qacc'][:, :self.nv_robot].astype(np.float32), dist=dist, d_dist=d_dist, **extra)
        return obs, reward, done, info

    def _vel_acc3(self, xmat, cur):
        """ego_vel_acc (engine.py:902-929) for a free-joint root: all three components of xpos, full xmat; fp32."""
        f = np.float32
        dt = f(float(self.gx_config.opt_timestep) * self.substeps)
        pl, pll = cur, cur
        if self.h_depth >= 1:                                   # last_done is not None
            pl = np.where((self.done > 0)[:, None], cur, self.hx_data)
            if self.h_depth >= 2:                               # last_last_done is not None
                pll = np.where((self.h_last_done + self.done > 0)[:, None], pl, self.hx_last)
        vw = ((cur - pl) / dt).astype(f)
        lv = ((pl - pll) / dt).astype(f)
        aw = ((vw - lv) / dt).astype(f)
        vel = np.einsum('nk,nkj->nj', vw, xmat).astype(f)[:, :2]
        acc = np.einsum('nk,nkj->nj', aw, xmat).astype(f)[:, :2]
        return vel, acc

    def reset_done(self, layout):
        """Engine.reset_done + mjx_reset_done (engine.py:497-505, 702-731)."""
        layout = np.asarray(layout, np.float32)
        sel = self.done > 0
        if not sel.any():
            return self._obs.copy()
        q = self.layout2qpos(layout)
        self.qpos = np.where(sel[:, None], q, self.qpos).astype(self.dtype)
        self.qvel = np.where(sel[:, None], 0, self.qvel).astype(self.dtype)
        self.ctrl = np.where(sel[:, None], 0, self.ctrl).astype(self.dtype)
        # throw-away step(s) on the re-placed envs only (the other rows are masked out anyway)
        qp, qv, w = self.qpos[sel], self.qvel[sel], self.warm[sel]
        qpre = qp
        for _ in range(self.substeps):
            qpre = qp
            qp, qv, d = mx.step(self.mm, qp, qv, self.ctrl[sel], w)
            w = d['qacc']
        obs_reset, _, _, _ = self._obs_rows(d, qpre, qp, qv, self.ctrl[sel])
        obs = self._obs.copy()
        obs[sel] = obs_reset
        self._obs = obs
        return obs

    # state rows as gx_get_state documents for GX_ROBOT_ART
    def get_state(self):
        nr = self.nq_robot
        goal = np.stack([self.qpos[:, self.qadr['goal_x']], self.qpos[:, self.qadr['goal_y']]], 1)
        hz = np.concatenate([np.stack([self.qpos[:, self.qadr[f'hazard{h}_x']], self.qpos[:, self.qadr[f'hazard{h}_y']]], 1)
                             for h in range(self.nh)], 1)
        nvr = self.nv_robot
        parts = [self.qpos[:, :nr], self.qvel[:, :nvr], self.warm[:, :nvr], self.prev_dist[:, None],
                 self.done[:, None], self.steps[:, None], goal, hz]
        if self.free_root:       # general-tree rows also carry the goal / hazard slide velocities
            parts.append(self.qvel[:, nvr:])
        return np.concatenate(parts, 1).astype(np.float32)

    def set_state(self, rows):
        """inverse of get_state(): load env states read from the CUDA env (gx_get_state rows), so that the oracle can
        take ONE step from wherever a large CUDA batch has got to (tests at BASELINE.json configs[3] size)."""
        rows = np.asarray(rows, np.float32)
        assert rows.shape[0] == self.n
        nr, nvr = self.nq_robot, self.nv_robot
        o = 0
        self.qpos = np.zeros((self.n, self.model.nq), self.dtype); self.qvel = np.zeros((self.n, self.model.nv), self.dtype)
        self.warm = np.zeros((self.n, self.model.nv), self.dtype)
        self.qpos[:, :nr] = rows[:, o:o + nr]; o += nr
        self.qvel[:, :nvr] = rows[:, o:o + nvr]; o += nvr
        self.warm[:, :nvr] = rows[:, o:o + nvr]; o += nvr
        self.prev_dist = rows[:, o].copy(); self.done = rows[:, o + 1].copy(); self.steps = rows[:, o + 2].copy(); o += 3
        self.qpos[:, self.qadr['goal_x']] = rows[:, o]; self.qpos[:, self.qadr['goal_y']] = rows[:, o + 1]; o += 2
        for h in range(self.nh):
            self.qpos[:, self.qadr[f'hazard{h}_x']] = rows[:, o]; self.qpos[:, self.qadr[f'hazard{h}_y']] = rows[:, o + 1]; o += 2
        if self.free_root:
            self.qvel[:, nvr:] = rows[:, o:o + self.model.nv - nvr]; o += self.model.nv - nvr
        assert o == rows.shape[1], (o, rows.shape)

    def round_state_to_fp32(self):
        """Make the float64 state exactly what get_state() hands to the CUDA env (one-step parity
        from bit-identical inputs)."""
        f = lambda a: a.astype(np.float32).astype(self.dtype)
        self.qpos, self.qvel, self.warm = f(self.qpos), f(self.qvel), f(self.warm)
