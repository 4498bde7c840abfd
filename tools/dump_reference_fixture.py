#!/usr/bin/env python
"""Dump a TRUE parity fixture from the real reference (needs jax + mujoco + mujoco-mjx +
gym + xmltodict, none of which are installable in the build image -- this script has NOT
been run there; it exists so that the assumption tags A1-A4 of DESIGN.md can be
discharged on any machine that has the reference's dependencies).

    # on a machine with the reference installed (JAX_PLATFORMS=cpu is fine)
    python tools/dump_reference_fixture.py dump  --ref /path/to/guardX --out fixture_true.npz
    # anywhere (CPU is enough): replay the same layouts/actions through the oracle
    python tools/dump_reference_fixture.py check --fixture fixture_true.npz

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
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

N, T = 64, 1000


def dump(args):
    sys.path.insert(0, os.path.join(args.ref, 'safe_rl_envs'))
    import jax
    import jax.numpy as jp
    import torch
    from safe_rl_envs.envs.engine import Engine
    env = Engine({'env_num': N, 'num_steps': 1000, '_seed': 0})
    env.reset()
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
    acts = np.random.default_rng(0).uniform(-1, 1, (T, N, 2)).astype(np.float32)
    rid = env.body_name2xpos_id['robot']
    for t in range(T):
        o, r, d, info = env.step(torch.as_tensor(acts[t]))
        rec['obs'].append(o.cpu().numpy()); rec['reward'].append(r.cpu().numpy())
        rec['done'].append(d.cpu().numpy()); rec['cost'].append(info['cost'].cpu().numpy())
        rec['qpos'].append(np.asarray(env._data.qpos[:, :3])); rec['qvel'].append(np.asarray(env._data.qvel[:, :3]))
        rec['xpos'].append(np.asarray(env._data.xpos[:, rid])); rec['xmat'].append(np.asarray(env._data.xmat[:, rid]))
    model = dict(body_mass=m.body_mass[rid], body_ipos=m.body_ipos[rid], body_inertia=m.body_inertia[rid],
                 dof_damping=m.dof_damping[:3], actuator_gear=m.actuator_gear[:, 0],
                 actuator_gainprm=m.actuator_gainprm[:, :3], actuator_biasprm=m.actuator_biasprm[:, :3],
                 actuator_ctrlrange=m.actuator_ctrlrange, actuator_forcerange=m.actuator_forcerange,
                 actuator_ctrllimited=m.actuator_ctrllimited, actuator_forcelimited=m.actuator_forcelimited,
                 timestep=m.opt.timestep)
    np.savez_compressed(args.out, layout=lay, actions=acts, reset_obs=np.asarray(obs),
                        **{k: np.asarray(v) for k, v in rec.items()},
                        **{'model_' + k: np.asarray(v) for k, v in model.items()})
    print('wrote', args.out)


def check(args):
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
    d.add_argument('--layouts', default=None)
    c = sub.add_parser('check'); c.add_argument('--fixture', required=True)
    a = ap.parse_args()
    (dump if a.cmd == 'dump' else check)(a)
