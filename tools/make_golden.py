#!/usr/bin/env python
"""Generate tests/golden/refsrc_*.npz by executing the REFERENCE'S OWN SOURCE.

The reference (``/root/reference/safe_rl_envs/safe_rl_envs/envs/engine.py``) cannot
be run as a whole here: jax, mujoco and gym are not installable in this image.
But everything in it except ``mjx.step`` / ``mjx.forward`` is plain array code
(``Engine.obs``, ``obs_lidar``, ``obs_compass``, ``ego_xy``, ``reward_done``,
``cost``, ``sample_layout``, ``draw_placement``).  This script imports the
unmodified file with ``tools/jnp_shim.py`` standing in for jax.numpy and stubs
for the other imports, builds an ``Engine`` WITHOUT running its MuJoCo-dependent
``__init__`` (only ``parse`` / ``build_placements_dict`` /
``build_observation_space`` and the body-id table that ``__init__`` would fill),
feeds it hand-made ``mjx.Data``-like objects, and records what the reference's
methods return.

Run here (needs /root/reference; the GPU box does not have it):
    python tools/make_golden.py
The vectors are committed; tests/test_oracle_golden.py checks the C oracle
against them, and tests/test_parity_gpu.py checks the CUDA kernels against the
same vectors through state injection.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
REF = '/root/reference/safe_rl_envs'


def load_reference_engine():
    import jnp_shim
    jnp_shim.install()
    sys.path.insert(0, REF)
    from safe_rl_envs.envs.engine import Engine  # the unmodified reference class
    return Engine, jnp_shim


def make_engine(Engine, config):
    eng = Engine.__new__(Engine)
    eng.parse(config)                                   # engine.py:322-328
    nh = eng.hazards_num
    eng.robot = SimpleNamespace(nq=3, nv=3, nu=3, z_height=0.1)  # world.py:430-438 for point.xml
    eng.body_name2xpos_id = {'robot': 1, 'goal': 2, 'hazards': list(range(3, 3 + nh))}  # engine.py:303-308
    eng.dt = 0.02 * eng.physics_steps_per_control_step   # engine.py:235
    eng.build_observation_space()                        # engine.py:386-418
    eng.build_placements_dict()                          # engine.py:533-544
    return eng


def fake_data(shim, robot_xy, theta_cs, goal_xy, hazards_xy, qpos, qvel, ctrl, qacc):
    """An object with the mjx.Data fields Engine reads (engine.py:754-766, 781-784, 806-807)."""
    nh = len(hazards_xy)
    nbody = 3 + nh
    xpos = np.zeros((nbody, 3), np.float32)
    xpos[1] = [robot_xy[0], robot_xy[1], 0.1]
    xpos[2] = [goal_xy[0], goal_xy[1], 0.0]
    for h in range(nh):
        xpos[3 + h] = [hazards_xy[h][0], hazards_xy[h][1], 0.02]
    xmat = np.tile(np.eye(3, dtype=np.float32), (nbody, 1, 1))
    c, s = theta_cs
    xmat[1] = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)
    nq = 3 + 2 + 2 * nh
    full = lambda v: np.concatenate([np.asarray(v, np.float32), np.zeros(nq - 3, np.float32)])
    W = lambda a: np.asarray(a, np.float32).view(shim.JArr)
    return SimpleNamespace(xpos=W(xpos), xmat=W(xmat), qpos=W(full(qpos)), qvel=W(full(qvel)),
                           qacc=W(full(qacc)), ctrl=W(np.asarray(ctrl, np.float32)))


def gen_obs_cases(Engine, shim, config, n_random, rng, tag):
    eng = make_engine(Engine, config)
    nh = eng.hazards_num
    cases = []
    for k in range(n_random):
        robot = rng.uniform(-3, 3, 2)
        th = rng.uniform(-20, 20)
        cs = (np.float32(np.cos(th)), np.float32(np.sin(th)))
        goal = rng.uniform(-2, 2, 2)
        haz = rng.uniform(-2, 2, (nh, 2))
        if k % 5 == 0:   # robot inside a hazard / near the goal
            haz[rng.integers(nh)] = robot + rng.uniform(-0.2, 0.2, 2)
        if k % 7 == 0:
            goal = robot + rng.uniform(-0.45, 0.45, 2)
        cases.append((robot, cs, goal, haz, th))
    # hand-made edges
    one = (np.float32(1.0), np.float32(0.0))
    hz0 = np.tile(np.array([[5.0, 5.0]]), (nh, 1))
    edge = lambda gx, gy: ((0.0, 0.0), one, (gx, gy), hz0.copy(), 0.0)
    cases += [
        edge(1.0, -1e-8),        # tiny negative angle: angle % 2pi == 2pi in fp32 -> bin == nbins
        edge(1.0, -1e-7),
        edge(-1.0, 0.0),         # angle == pi exactly
        edge(-1.0, -0.0),
        edge(0.0, 0.0),          # object on top of the robot: atan2(0, 0)
        edge(0.0, 1.0), edge(0.0, -1.0), edge(1.0, 0.0),
        edge(0.5, 0.0),          # dist == goal_size exactly (strict <)
        edge(np.cos(np.pi / 8), np.sin(np.pi / 8)),    # on a bin boundary
        edge(np.cos(-np.pi / 8), np.sin(-np.pi / 8)),
    ]
    hz_edge = hz0.copy(); hz_edge[0] = [0.3, 0.0]   # dist == hazards_size exactly
    if nh > 1:
        hz_edge[1] = [0.0, 0.1]
    cases.append(((0.0, 0.0), one, (3.0, 0.0), hz_edge, 0.0))
    out = {k: [] for k in ('robot', 'cs', 'goal', 'hazards', 'qpos', 'qvel', 'ctrl', 'qacc', 'flat', 'cost',
                           'last_robot', 'last_goal', 'last_done', 'reward', 'done', 'reward_nolast',
                           'done_nolast', 'goal_bin_lidar', 'compass')}
    for robot, cs, goal, haz, th in cases:
        qpos = np.r_[robot, th]   # self-consistent pose: cs = (cos, sin)(qpos[2])
        qvel = rng.uniform(-3, 3, 3)
        ctrl = rng.normal(size=3)
        qacc = rng.normal(size=3)
        data = fake_data(shim, robot, cs, goal, haz, qpos, qvel, ctrl, qacc)
        flat, d = eng.obs(data, None, None, None, None)           # engine.py:738-778
        cost = eng.cost(data)                                     # engine.py:804-811
        # reward_done with a previous frame (engine.py:787-802)
        jump = rng.random() < 0.15
        lrobot = np.asarray(robot) + (rng.uniform(1.0, 2.0, 2) if jump else rng.uniform(-0.07, 0.07, 2))
        lgoal = goal
        ldone = float(rng.random() < 0.3)
        ldata = fake_data(shim, lrobot, cs, lgoal, haz, qpos, qvel, ctrl, qacc)
        r1, d1 = eng.reward_done(data, ldata, np.float32(ldone).view(shim.JArr) if False else shim._wrap(np.float32(ldone)))
        r0, d0 = eng.reward_done(data, ldata, None)
        out['robot'].append(robot); out['cs'].append(cs); out['goal'].append(goal); out['hazards'].append(haz)
        out['qpos'].append(qpos); out['qvel'].append(qvel); out['ctrl'].append(ctrl); out['qacc'].append(qacc)
        out['flat'].append(np.asarray(flat)); out['cost'].append(float(cost))
        out['last_robot'].append(lrobot); out['last_goal'].append(lgoal); out['last_done'].append(ldone)
        out['reward'].append(float(r1)); out['done'].append(float(d1))
        out['reward_nolast'].append(float(r0)); out['done_nolast'].append(float(d0))
        out['goal_bin_lidar'].append(np.asarray(d['goal_lidar']))
        out['compass'].append(np.asarray(d['goal_compass']))
    res = {k: np.asarray(v, np.float32) for k, v in out.items()}
    res['obs_keys'] = np.array(sorted(eng.obs_space_dict.keys()))
    path = os.path.join(ROOT, 'tests', 'golden', f'refsrc_obs_{tag}.npz')
    np.savez_compressed(path, config=np.array(repr(config)), **res)
    print(path, res['flat'].shape)


def gen_velacc_cases(Engine, shim, n, rng):
    """ego_vel_acc (engine.py:902-929) on random pose histories, for the three None-ness
    states of (last_done, last_last_done) the Engine goes through on its first steps."""
    eng = make_engine(Engine, {'observe_vel': True, 'observe_acc': True})
    nh = eng.hazards_num
    out = {k: [] for k in ('cs', 'cur', 'last', 'last_last', 'have', 'last_done', 'last_last_done', 'vel', 'acc', 'flat')}
    z3 = np.zeros(3)
    for k in range(n):
        th = rng.uniform(-10, 10)
        cs = (np.float32(np.cos(th)), np.float32(np.sin(th)))
        cur = rng.uniform(-2, 2, 2)
        last = cur - rng.uniform(-0.06, 0.06, 2)
        ll = last - rng.uniform(-0.06, 0.06, 2)
        have = k % 3
        ld, lld = float(rng.random() < 0.25), float(rng.random() < 0.25)
        goal, haz = rng.uniform(-2, 2, 2), rng.uniform(-2, 2, (nh, 2))
        mk = lambda xy: fake_data(shim, xy, cs, goal, haz, np.r_[xy, th], z3, z3, z3)
        data, ldata, lldata = mk(cur), mk(last), mk(ll)
        W = lambda v: shim._wrap(np.float32(v))
        a_ld = None if have == 0 else W(ld)
        a_lld = None if have <= 1 else W(lld)
        vel, acc = eng.ego_vel_acc(data, ldata, lldata if have == 2 else ldata, a_ld, a_lld)
        flat, _ = eng.obs(data, ldata, lldata if have == 2 else None, a_ld, a_lld)
        out['cs'].append(cs); out['cur'].append(cur); out['last'].append(last); out['last_last'].append(ll)
        out['have'].append(have); out['last_done'].append(ld); out['last_last_done'].append(lld)
        out['vel'].append(np.asarray(vel)); out['acc'].append(np.asarray(acc)); out['flat'].append(np.asarray(flat))
    res = {k: np.asarray(v, np.float32) for k, v in out.items()}
    res['obs_keys'] = np.array(sorted(eng.obs_space_dict.keys()))
    path = os.path.join(ROOT, 'tests', 'golden', 'refsrc_velacc.npz')
    np.savez_compressed(path, **res)
    print(path, res['flat'].shape)


def gen_layout_cases(Engine, shim, config, n, tag):
    eng = make_engine(Engine, config)
    import jax
    names = list(eng.placements.keys())           # placement order (engine.py:538-542)
    draws, layouts, succ = [], [], []
    for seed in range(n):
        shim.DRAW_LOG.clear()
        layout, success = eng.sample_layout(jax.random.PRNGKey(seed))   # engine.py:546-572
        d = np.asarray(shim.DRAW_LOG, np.float32).reshape(len(names), 10, 2)
        lay = np.stack([np.asarray(layout[k], np.float32) for k in names])
        draws.append(d); layouts.append(lay); succ.append(float(success))
    path = os.path.join(ROOT, 'tests', 'golden', f'refsrc_layout_{tag}.npz')
    np.savez_compressed(path, config=np.array(repr(config)), names=np.array(names),
                        draws=np.asarray(draws), layouts=np.asarray(layouts), success=np.asarray(succ, np.float32))
    print(path, 'success rate', np.mean(succ))


def main():
    Engine, shim = load_reference_engine()
    rng = np.random.default_rng(20261018)
    default = {}  # Engine.DEFAULT == Goal_Point_8Hazards for every key that matters
    variant = {'lidar_num_bins': 10, 'lidar_alias': False, 'lidar_exp_gain': 0.5, 'hazards_num': 3,
               'goal_size': 0.3, 'hazards_size': 0.2, 'reward_distance': 2.0}
    gen_obs_cases(Engine, shim, default, 400, rng, 'default')
    gen_obs_cases(Engine, shim, variant, 200, rng, 'variant')
    gen_velacc_cases(Engine, shim, 300, rng)
    gen_layout_cases(Engine, shim, default, 600, 'default')
    gen_layout_cases(Engine, shim, {'hazards_num': 5, 'hazards_keepout': 0.25, 'goal_keepout': 0.35,
                                    'robot_keepout': 0.3, 'placements_margin': 0.05,
                                    'placements_extents': [-2.5, -2, 2.5, 2]}, 400, 'variant')
    # the key set of Engine.DEFAULT, for tests/test_config.py
    import json
    with open(os.path.join(ROOT, 'tests', 'golden', 'engine_default.json'), 'w') as f:
        json.dump({k: (v if not isinstance(v, np.ndarray) else v.tolist()) for k, v in Engine.DEFAULT.items()},
                  f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()
