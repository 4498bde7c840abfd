"""Oracle-side config surface (TEST INFRASTRUCTURE ONLY -- see oracle/oracle.py).

The oracle must not take its defaults, its task table or its model constants from the product
(guardx_b200/config.py), or a wrong default would pass every comparison on both sides.  Here:
  * DEFAULT is the JSON dump of the reference's own ``Engine.DEFAULT`` (engine.py:98-204) that
    tools/make_golden.py wrote from the unmodified reference source
    (tests/golden/engine_default.json);
  * ``parse`` restates Engine.parse (engine.py:322-328);
  * ``lower`` fills the C ABI's config struct -- the ctypes mirror of include/guardx.h in
    guardx_b200/_abi.py is the ONE thing the oracle shares with the product, because the struct IS
    the boundary -- with the Point constants taken from the oracle's own model compile
    (oracle/mj_compile.py) of the reference's xmls/point.xml.
"""
from __future__ import annotations

import copy
import json
import os

from guardx_b200 import _abi          # ctypes mirror of include/guardx.h only
from . import mj_compile

_HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(_HERE, '..', 'tests', 'golden', 'engine_default.json')) as _f:
    DEFAULT = json.load(_f)

ROBOT_KIND = {'xmls/point.xml': _abi.GX_ROBOT_POINT, 'xmls/swimmer.xml': _abi.GX_ROBOT_ART, 'xmls/ant.xml': _abi.GX_ROBOT_ART,
              'xmls/walker.xml': _abi.GX_ROBOT_ART, 'xmls/doggo.xml': _abi.GX_ROBOT_TREE}


def parse(config: dict) -> dict:
    merged = copy.deepcopy(DEFAULT)
    merged.update(copy.deepcopy(config))
    for key in merged:
        assert key in DEFAULT, f'Bad key {key}'
    return merged


def world_model(cfg: dict) -> mj_compile.OModel:
    """what World.build + MjModel.from_xml_string give for this config (engine.py:220-232)"""
    return mj_compile.world_model(cfg['robot_base'], cfg['goal_size'], cfg['hazards_size'], cfg['hazards_num'])


def point_constants(m: mj_compile.OModel) -> _abi.GxPointModel:
    """the closed-form constants of the Point robot read off the compiled model of xmls/point.xml"""
    r = m.body_id('robot')
    pm = _abi.GxPointModel()
    pm.timestep, pm.body_z = m.opt['timestep'], m.body_pos[r][2]
    # inertia_zz of the C struct is the yaw inertia about the JOINT axis (body origin): centre-of-mass value + m c^2
    pm.mass, pm.com_x = m.body_mass[r], m.body_ipos[r][0]
    pm.inertia_zz = m.body_inertia_full[r][2, 2] + m.body_mass[r] * float(m.body_ipos[r] @ m.body_ipos[r])
    d0 = m.body_dofadr[r]
    pm.damping[:] = [float(v) for v in m.dof_damping[d0:d0 + 3]]
    pm.gear[:] = [float(v) for v in m.actuator_gear[:3]]
    pm.gainprm0, pm.biasprm2 = m.actuator_gainprm0[0], m.actuator_biasprm[0][2]
    pm.ctrllimited, pm.forcelimited = int(m.actuator_ctrllimited[0]), int(m.actuator_forcelimited[0])
    pm.ctrlrange[:] = [float(v) for v in m.actuator_ctrlrange[0]]
    pm.forcerange[:] = [float(v) for v in m.actuator_forcerange[0]]
    return pm


def lower(cfg: dict, env_id_offset: int = 0, pool_size: int = 0) -> _abi.GxConfig:
    assert cfg['task'] == 'goal' and cfg['observation_flatten'] and cfg['hazards_num']
    c = _abi.GxConfig()
    c.abi_version = _abi.GX_ABI_VERSION
    c.env_num, c.num_steps = int(cfg['env_num']), int(cfg['num_steps'])
    c.robot = ROBOT_KIND[cfg['robot_base']]
    c.seed = int(cfg['_seed']) & 0xFFFFFFFFFFFFFFFF
    c.env_id_offset = int(env_id_offset)
    c.placements_extents[:] = [float(v) for v in cfg['placements_extents']]
    c.placements_margin = float(cfg['placements_margin'])
    c.robot_keepout, c.goal_keepout, c.hazards_keepout = (float(cfg[k]) for k in ('robot_keepout', 'goal_keepout', 'hazards_keepout'))
    c.goal_size, c.hazards_size = float(cfg['goal_size']), float(cfg['hazards_size'])
    c.reward_distance = float(cfg['reward_distance'])
    c.lidar_max_dist = 0.0 if cfg['lidar_max_dist'] is None else float(cfg['lidar_max_dist'])
    c.lidar_exp_gain = float(cfg['lidar_exp_gain'])
    c.goal_min_robot_dist = 3.0          # engine.py:571
    c.layout_tries = 10                  # engine.py:562
    m = world_model(cfg)
    nworld = 2 + 2 * int(cfg['hazards_num'])
    c.opt_timestep = float(m.opt['timestep'])
    if c.robot != _abi.GX_ROBOT_POINT:
        c.art_nu, c.art_nv, c.art_nq = int(m.nu), int(m.nv) - nworld, int(m.nq) - nworld
    c.hazards_num, c.lidar_num_bins = int(cfg['hazards_num']), int(cfg['lidar_num_bins'])
    c.lidar_alias = 1 if cfg['lidar_alias'] else 0
    c.physics_steps_per_control_step = int(cfg['physics_steps_per_control_step'])
    c.pool_size = int(pool_size)
    mask = 0
    for key, bit in (('observe_acc', _abi.GX_OBS_ACC), ('observe_ctrl', _abi.GX_OBS_CTRL), ('observe_goal_comp', _abi.GX_OBS_GOAL_COMPASS),
                     ('observe_goal_lidar', _abi.GX_OBS_GOAL_LIDAR), ('observe_hazards', _abi.GX_OBS_HAZARDS_LIDAR),
                     ('observe_qpos', _abi.GX_OBS_QPOS), ('observe_qvel', _abi.GX_OBS_QVEL), ('observe_vel', _abi.GX_OBS_VEL)):
        if cfg[key]:
            mask |= bit
    c.obs_mask = mask
    c.point = point_constants(mj_compile.world_model('xmls/point.xml', cfg['goal_size'], cfg['hazards_size'], cfg['hazards_num']))
    return c
