#!/usr/bin/env python
"""Dump a TRUE parity fixture from the real reference (needs jax + mujoco + mujoco-mjx +
gym + xmltodict, none of which are installable in the build image -- this script has NOT
been run there; it exists so that the assumption tags A1-A4 of DESIGN.md can be
discharged on any machine that has the reference's dependencies).

    # on a machine with the reference installed (JAX_PLATFORMS=cpu is fine)
    python tests/tools/dump_reference_fixture.py dump  --ref /path/to/guardX --out fixture_true.npz
    # anywhere (CPU is enough): replay the same layouts/actions through the oracle
    python tests/tools/dump_reference_fixture.py check --fixture fixture_true.npz

`dump` drives the reference exactly as BASELINE.json configs[0] says (Goal_Point_8Hazards,
64 envs, 1000 random-action steps), but with INJECTED layouts and actions so that RNG
streams do not matter: layouts come from `Engine._reset(layout)` (engine.py:269, 460) with a
layout dict built from --layouts (default: the first 64 accepted layouts of the reference's own
pool), actions are U(-1,1) from numpy's PCG64(seed 0).  It records, per step, obs / reward /
done / cost and the robot qpos / qvel / xpos / xmat, plus the compiled model constants that
`gx_point_model` hard-codes (body_mass, body_ipos, body_inertia, dof_damping, actuator_*), so A1
can be checked directly.

`check` replays the fixture through oracle/gx_oracle.c and prints, per quantity, the first
step at which |diff| exceeds 1e-5 and the divergence curve (max |diff| every 50 steps).

`--robot ant|walker|swimmer` does the same for the articulated robots: `dump` additionally
records the compiled model constants the model compiler (guardx_b200/mjcf.py) reproduces
(body_mass, body_ipos, body_inertia + body_iquat, dof_armature / damping / invweight0,
body_invweight0, jnt_range, stat.meaninertia, geom margins / friction), `check` compares them
with ours one by one and replays the steps through oracle/art_oracle.py (numpy mjx.step).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

N, T = 64, 1000
ROBOT_BASE = {'point': 'xmls/point.xml', 'ant': 'xmls/ant.xml', 'walker': 'xmls/walker.xml', 'swimmer': 'xmls/swimmer.xml'}


def dump(args):
    sys.path.insert(0, os.path.join(args.ref, 'safe_rl_envs'))
    import jax
    import jax.numpy as jp
    import torch
    from safe_rl_envs.envs.engine import Engine
    env = Engine({'env_num': N, 'num_steps': 1000, '_seed': 0, 'robot_base': ROBOT_BASE[args.robot]})
    env.reset()
    nq_r, nu = env.robot.nq, env.action_space.shape[0]
    names = list(env.placements.keys())                      # goal, hazard0.., robot
    if args.layouts:
        lay = np.load(args.layouts)['layout'].astype(np.float32)   # [N, 20] goal, robot, hazards
        order = {'goal': lay[:, 0:2], 'robot': lay[:, 2:4]}
        for h in range(8):
            order[f'hazard{h}'] = lay[:, 4 + 2 * h:6 + 2 * h]
        layout = {k: jp.asarray(order[k]) for k in names}
    else:
        layout = {k: env.layout[k][:N] for k in names}
        lay = np.concatenate([np.asarray(layout['goal']), np.asarray(layout['robot'])] +
                             [np.asarray(layout[f'hazard{h}']) for h in range(8)], axis=1)
    obs, data = env._reset(layout)
    env._steps = jp.zeros(N); env._obs = obs; env._data = data
    m = env.mj_model
    rec = {k: [] for k in ('obs', 'reward', 'done', 'cost', 'qpos', 'qvel', 'xpos', 'xmat')}
    acts = np.random.default_rng(0).uniform(-1, 1, (T, N, nu)).astype(np.float32)
    rid = env.body_name2xpos_id['robot']
    for t in range(T):
        o, r, d, info = env.step(torch.as_tensor(acts[t]))
        rec['obs'].append(o.cpu().numpy()); rec['reward'].append(r.cpu().numpy())
        rec['done'].append(d.cpu().numpy()); rec['cost'].append(info['cost'].cpu().numpy())
        rec['qpos'].append(np.asarray(env._data.qpos[:, :nq_r])); rec['qvel'].append(np.asarray(env._data.qvel[:, :nq_r]))
        rec['xpos'].append(np.asarray(env._data.xpos[:, rid])); rec['xmat'].append(np.asarray(env._data.xmat[:, rid]))
    model = dict(body_mass=m.body_mass[rid], body_ipos=m.body_ipos[rid], body_inertia=m.body_inertia[rid],
                 dof_damping=m.dof_damping[:3], actuator_gear=m.actuator_gear[:, 0],
                 actuator_gainprm=m.actuator_gainprm[:, :3], actuator_biasprm=m.actuator_biasprm[:, :3],
                 actuator_ctrlrange=m.actuator_ctrlrange, actuator_forcerange=m.actuator_forcerange,
                 actuator_ctrllimited=m.actuator_ctrllimited, actuator_forcelimited=m.actuator_forcelimited,
                 timestep=m.opt.timestep,
                 all_body_mass=m.body_mass, all_body_ipos=m.body_ipos, all_body_inertia=m.body_inertia, all_body_iquat=m.body_iquat,
                 all_body_pos=m.body_pos, all_body_quat=m.body_quat, all_body_invweight0=m.body_invweight0,
                 all_dof_armature=m.dof_armature, all_dof_damping=m.dof_damping, all_dof_invweight0=m.dof_invweight0,
                 all_jnt_range=m.jnt_range, all_jnt_axis=m.jnt_axis, all_jnt_stiffness=m.jnt_stiffness, meaninertia=m.stat.meaninertia,
                 all_geom_margin=m.geom_margin, all_geom_friction=m.geom_friction, all_geom_size=m.geom_size,
                 all_actuator_gear=m.actuator_gear[:, 0], nq=m.nq, nv=m.nv)
    np.savez_compressed(args.out, layout=lay, actions=acts, reset_obs=np.asarray(obs),
                        **{k: np.asarray(v) for k, v in rec.items()},
                        **{'model_' + k: np.asarray(v) for k, v in model.items()})
    print('wrote', args.out)


def check_art(args):
    """Articulated robots: constants of the model compiler vs the real MjModel, then the steps
    through the numpy mjx.step oracle (free-running; the first divergence step is what matters)."""
    from oracle.art_oracle import ArtOracleEngine
    from guardx_b200 import mjcf
    f = np.load(args.fixture)
    o = ArtOracleEngine({'env_num': N, 'num_steps': 1000, '_seed': 0, 'robot_base': ROBOT_BASE[args.robot]}, dtype=np.float32)
    m = o.model
    R = lambda q: np.stack([mjcf.quat2mat(x) for x in q])
    I_ref = np.einsum('bij,bj,bkj->bik', R(f['model_all_body_iquat']), f['model_all_body_inertia'], R(f['model_all_body_iquat']))
    pairs = dict(body_mass=(m.body_mass, f['model_all_body_mass']), body_ipos=(m.body_ipos, f['model_all_body_ipos']),
                 body_inertia_full=(m.body_inertia_full, I_ref), body_pos=(m.body_pos, f['model_all_body_pos']),
                 body_invweight0=(m.body_invweight0, f['model_all_body_invweight0']), dof_armature=(m.dof_armature, f['model_all_dof_armature']),
                 dof_damping=(m.dof_damping, f['model_all_dof_damping']), dof_invweight0=(m.dof_invweight0, f['model_all_dof_invweight0']),
                 jnt_range=(m.jnt_range, f['model_all_jnt_range']), jnt_axis=(m.jnt_axis, f['model_all_jnt_axis']),
                 geom_margin=(m.geom_margin, f['model_all_geom_margin']), geom_friction=(m.geom_friction, f['model_all_geom_friction']),
                 actuator_gear=(m.actuator_gear, f['model_all_actuator_gear']),
                 meaninertia=(np.array(m.stat_meaninertia), f['model_meaninertia']))
    for k, (ours, ref) in pairs.items():
        ours, ref = np.asarray(ours, float), np.asarray(ref, float)
        ok = ours.shape == ref.shape and np.allclose(ours, ref, rtol=1e-6, atol=1e-9)
        print(f'{k:20s} {"ok" if ok else "DIFFERS"}', '' if ok else (ours.shape, ref.shape, float(np.abs(ours - ref).max()) if ours.shape == ref.shape else ''))
    obs = o.reset(f['layout'])
    print('reset obs max|diff|', np.abs(obs - f['reset_obs']).max())
    first = {}
    for t in range(T):
        ob, r, d, info = o.step(f['actions'][t])
        cur = {'obs': np.abs(ob - f['obs'][t]).max(), 'reward': np.abs(r - f['reward'][t]).max(),
               'cost': np.abs(info['cost'] - f['cost'][t]).max(), 'done': np.abs(d - f['done'][t]).max()}
        for k, v in cur.items():
            if v > 1e-4 and k not in first:
                first[k] = t
        if t % 50 == 0:
            print(t, {k: float(v) for k, v in cur.items()})
        if d.any():   # the reference loop calls reset_done with pool layouts; a fixture run should avoid dones (num_steps 1000)
            print('done at step', t, '- stopping the free-running comparison here'); break
    print('first step with |diff| > 1e-4:', first or 'never')


def check(args):
    if args.robot != 'point':
        return check_art(args)
    from oracle.oracle import OracleEngine
    f = np.load(args.fixture)
    o = OracleEngine({'env_num': N, 'num_steps': 1000, '_seed': 0}, threads=os.cpu_count() or 1)
    for k in ('model_body_mass', 'model_dof_damping', 'model_actuator_gainprm', 'model_actuator_biasprm',
              'model_actuator_ctrlrange', 'model_actuator_forcerange'):
        print(k, f[k])
    pm = o.gx_config.point
    print('ours: mass', pm.mass, 'com_x', pm.com_x, 'I_zz(joint)', pm.inertia_zz, 'gain', pm.gainprm0, 'bias2', pm.biasprm2)
    obs = o.reset(f['layout'])
    print('reset obs max|diff|', np.abs(obs - f['reset_obs']).max())
    first = {}
    for t in range(T):
        ob, r, d, info = o.step(f['actions'][t])
        cur = {'obs': np.abs(ob - f['obs'][t]).max(), 'reward': np.abs(r - f['reward'][t]).max(),
               'cost': np.abs(info['cost'] - f['cost'][t]).max(), 'done': np.abs(d - f['done'][t]).max()}
        for k, v in cur.items():
            if v > 1e-5 and k not in first:
                first[k] = t
        if t % 50 == 0:
            print(t, {k: float(v) for k, v in cur.items()})
    print('first step with |diff| > 1e-5:', first or 'never')


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest='cmd', required=True)
    d = sub.add_parser('dump'); d.add_argument('--ref', required=True); d.add_argument('--out', default='fixture_true.npz')
    d.add_argument('--layouts', default=None); d.add_argument('--robot', default='point', choices=sorted(ROBOT_BASE))
    c = sub.add_parser('check'); c.add_argument('--fixture', required=True); c.add_argument('--robot', default='point', choices=sorted(ROBOT_BASE))
    a = ap.parse_args()
    (dump if a.cmd == 'dump' else check)(a)
