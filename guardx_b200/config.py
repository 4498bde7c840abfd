"""Config surface of the reference ``Engine`` and its lowering to ``gx_config``.

``DEFAULT`` reproduces the key set and default values of ``Engine.DEFAULT``
(safe_rl_envs/safe_rl_envs/envs/engine.py:98-204); ``parse`` reproduces
``Engine.parse`` (engine.py:322-328) including the ``Bad key`` assertion.
``configuration(task)`` holds the entries of
safe_rl_libX/guard_utils/safe_rl_env_config.py that the JAX Engine can actually
run.
"""
from __future__ import annotations

import math
from copy import deepcopy

from . import _abi

DEFAULT = {
    'num_steps': 1000,
    'device_id': 0,
    'env_num': 1,
    'placements_extents': [-2, -2, 2, 2],
    'placements_margin': 0.0,
    'floor_display_mode': False,
    'robot_placements': None,
    'robot_locations': [],
    'robot_keepout': 0.4,
    'robot_base': 'xmls/point.xml',
    'robot_rot': None,
    'observation_flatten': True,
    'observe_goal_lidar': True,
    'observe_goal_comp': True,
    'observe_hazards': True,
    'observe_qpos': True,
    'observe_qvel': True,
    'observe_qacc': True,
    'observe_vel': False,
    'observe_acc': False,
    'observe_ctrl': True,
    'observe_vision': False,
    'render_labels': False,
    'render_lidar_markers': True,
    'render_lidar_radius': 0.15,
    'render_lidar_size': 0.025,
    'render_lidar_offset_init': 0.5,
    'render_lidar_offset_delta': 0.06,
    'sensors_obs': ['accelerometer', 'velocimeter', 'gyro', 'magnetometer'],
    'sensors_hinge_joints': True,
    'sensors_ball_joints': True,
    'sensors_angle_components': True,
    'lidar_num_bins': 16,
    'lidar_num_bins3D': 1,
    'lidar_max_dist': None,
    'lidar_exp_gain': 1.0,
    'lidar_type': 'pseudo',
    'lidar_alias': True,
    'lidar_body': ['robot'],
    'task': 'goal',
    'push_object': 'box',
    'goal_mode': 'random',
    'goal_travel': 3.0,
    'goal_velocity': 0.5,
    'goal_placements': None,
    'goal_locations': [],
    'goal_keepout': 0.5,
    'goal_size': 0.5,
    'goal_3D': False,
    'goal_z_range': [1.0, 1.0],
    'reward_distance': 1.0,
    'reward_goal': 1.0,
    'reward_box_dist': 1.0,
    'reward_box_goal': 1.0,
    'reward_orientation': False,
    'reward_orientation_scale': 0.002,
    'reward_orientation_body': 'robot',
    'reward_exception': -10.0,
    'reward_x': 1.0,
    'reward_z': 1.0,
    'reward_circle': 1e-1,
    'reward_clip': 10,
    'reward_defense': 1.0,
    'reward_chase': 1.0,
    'constrain_hazards': False,
    'constrain_indicator': True,
    'hazards_num': 8,
    'hazards_placements': None,
    'hazards_locations': [],
    'hazards_keepout': 0.4,
    'hazards_size': 0.3,
    'hazards_cost': 1.0,
    'physics_steps_per_control_step': 1,
    '_seed': 0,
}

# robots this build has dynamics for -> GX_ROBOT_*
_ROBOTS = {'xmls/point.xml': _abi.GX_ROBOT_POINT, 'xmls/ant.xml': _abi.GX_ROBOT_ART,
           'xmls/walker.xml': _abi.GX_ROBOT_ART, 'xmls/swimmer.xml': _abi.GX_ROBOT_ART,
           'xmls/doggo.xml': _abi.GX_ROBOT_TREE}
_TASK_ROBOT = {'Point': 'xmls/point.xml', 'Ant': 'xmls/ant.xml', 'Walker': 'xmls/walker.xml', 'Swimmer': 'xmls/swimmer.xml',
               'Doggo': 'xmls/doggo.xml'}


def parse(config: dict) -> dict:
    """Engine.parse (engine.py:322-328): defaults <- config, unknown key asserts."""
    merged = deepcopy(DEFAULT)
    merged.update(deepcopy(config))
    for key in merged:
        assert key in DEFAULT, f'Bad key {key}'
    return merged


# touch sensors each robot file defines; the JAX Engine never reads `sensors_obs`, but callers see it in
# `env.config` and it is part of what a pickled env carries (safe_rl_env_config.py:123, 179, 207, 266)
_BASE_SENSORS = ['accelerometer', 'velocimeter', 'gyro', 'magnetometer']
_ANKLE_TOUCH = [f'touch_ankle_{i}{s}' for s in 'ab' for i in (1, 2, 3, 4)]
_TASK_SENSORS = {'Swimmer': _BASE_SENSORS + [f'touch_point{i}' for i in (1, 2, 3, 4)],
                 'Ant': _BASE_SENSORS + _ANKLE_TOUCH, 'Doggo': _BASE_SENSORS + _ANKLE_TOUCH,
                 'Walker': _BASE_SENSORS + ['touch_right_foot', 'touch_left_foot']}


def configuration(task: str) -> dict:
    """The runnable subset of safe_rl_env_config.configuration_list: equal, key by key, to the reference's
    dicts (tests/golden/task_configs.json, tests/test_oracle_golden.py)."""
    # safe_rl_env_config.py:59-82 (Point), :111-136 (Swimmer), :167-193 (Doggo), :195-221 (Ant), :254-280 (Walker)
    name = task[:-len("_noconti")] if task.endswith("_noconti") else task
    parts = name.split("_")
    if len(parts) == 3 and parts[0] == "Goal" and parts[2] == "8Hazards" and parts[1] in _TASK_ROBOT:
        cfg = {
            'robot_base': _TASK_ROBOT[parts[1]],
            'task': 'goal',
            'goal_size': 0.5,
            'observe_goal_comp': True,
            'observe_hazards': True,
            'constrain_hazards': True,
            'constrain_indicator': False,
            'lidar_num_bins': 16,
            'hazards_num': 8,
            'hazards_size': 0.3,
        }
        if parts[1] in _TASK_SENSORS:
            cfg['sensors_obs'] = list(_TASK_SENSORS[parts[1]])
        return cfg
    raise NotImplementedError(
        f"task {task!r}: built are Goal_{{Point,Swimmer,Ant,Walker,Doggo}}_8Hazards "
        "(the Push/Chase/Defense/Ghost entries are rejected by the reference Engine itself, engine.py:326-328)")


def compiled_model(cfg: dict):
    """The MuJoCo model World.build() would emit for this config (world.py:104-333), compiled
    by guardx_b200.mjcf -- None for the Point, whose constants are closed-form (point_model)."""
    if _ROBOTS.get(cfg['robot_base']) not in (_abi.GX_ROBOT_ART, _abi.GX_ROBOT_TREE):
        return None
    from . import mjcf, robots
    spec = mjcf.add_world_geoms(robots.robot_spec(cfg['robot_base']), cfg['goal_size'], cfg['hazards_size'],
                                cfg['hazards_num'])
    return mjcf.compile_spec(spec)


def obs_mask_of(cfg: dict) -> int:
    m = 0
    if cfg['observe_acc']:
        m |= _abi.GX_OBS_ACC
    if cfg['observe_ctrl']:
        m |= _abi.GX_OBS_CTRL
    if cfg['observe_goal_comp']:
        m |= _abi.GX_OBS_GOAL_COMPASS
    if cfg['observe_goal_lidar']:
        m |= _abi.GX_OBS_GOAL_LIDAR
    if cfg['observe_hazards']:
        m |= _abi.GX_OBS_HAZARDS_LIDAR
    if cfg['observe_qpos']:
        m |= _abi.GX_OBS_QPOS
    if cfg['observe_qvel']:
        m |= _abi.GX_OBS_QVEL
    if cfg['observe_vel']:
        m |= _abi.GX_OBS_VEL
    return m


def point_model() -> _abi.GxPointModel:
    """SURVEY.md Appendix C, evaluated in double from xmls/point.xml:5,13,16-20,37-39
    (same arithmetic as gx_fill_default_config in csrc/gx_params.h)."""
    m1 = 4.0 / 3.0 * math.pi * 0.1 * 0.1 * 0.1
    m2 = 8.0 * 0.05 * 0.05 * 0.05
    m = m1 + m2
    cx = m2 * 0.1 / m
    izz = 0.4 * m1 * 0.01 + m2 * (0.05 * 0.05 + 0.05 * 0.05) / 3.0 + m2 * 0.01
    pm = _abi.GxPointModel()
    pm.timestep, pm.body_z = 0.02, 0.1
    pm.mass, pm.com_x, pm.inertia_zz = m, cx, izz
    pm.damping[:] = [0.01, 0.01, 0.005]
    pm.gear[:] = [0.3, 0.3, 0.3]
    pm.gainprm0, pm.biasprm2 = 1.0, -1.0
    pm.ctrllimited = 1
    pm.ctrlrange[:] = [-1.0, 1.0]
    pm.forcelimited = 1
    pm.forcerange[:] = [-0.05, 0.05]
    return pm


def lower(cfg: dict, env_id_offset: int = 0, pool_size: int = 0, model=None) -> _abi.GxConfig:
    """Parsed Engine config -> gx_config.  Raises for what the reference itself
    cannot run (SURVEY.md Appendix A quirks marked "don't")."""
    if cfg['robot_base'] not in _ROBOTS:
        raise NotImplementedError(f"robot_base {cfg['robot_base']!r} has no CUDA dynamics yet")
    assert cfg['task'] == 'goal', "the JAX Engine only places a goal for task='goal' (engine.py:538-539)"
    assert cfg['observation_flatten'], "observation_flatten=False crashes the reference (engine.py:773-778)"
    assert cfg['hazards_num'], "hazards_num == 0 crashes the reference (engine.py:370-384)"
    for k in ('robot_placements', 'goal_placements', 'hazards_placements'):
        assert cfg[k] is None, f"{k}: only the full-extents placement is supported (engine.py:616 uses an undefined self.rs otherwise)"
    for k in ('robot_locations', 'goal_locations', 'hazards_locations'):
        assert len(cfg[k]) == 0, f"{k}: fixed locations are not supported"
    assert not cfg['observe_vision'], "observe_vision is not implemented by the reference Engine"
    c = _abi.GxConfig()
    c.abi_version = _abi.GX_ABI_VERSION
    c.env_num = int(cfg['env_num'])
    c.num_steps = int(cfg['num_steps'])
    c.robot = _ROBOTS[cfg['robot_base']]
    c.seed = int(cfg['_seed']) & 0xFFFFFFFFFFFFFFFF
    c.env_id_offset = int(env_id_offset)
    c.placements_extents[:] = [float(v) for v in cfg['placements_extents']]
    c.placements_margin = float(cfg['placements_margin'])
    c.robot_keepout = float(cfg['robot_keepout'])
    c.goal_keepout = float(cfg['goal_keepout'])
    c.hazards_keepout = float(cfg['hazards_keepout'])
    c.goal_size = float(cfg['goal_size'])
    c.hazards_size = float(cfg['hazards_size'])
    c.reward_distance = float(cfg['reward_distance'])
    c.lidar_max_dist = 0.0 if cfg['lidar_max_dist'] is None else float(cfg['lidar_max_dist'])
    c.lidar_exp_gain = float(cfg['lidar_exp_gain'])
    c.goal_min_robot_dist = 3.0  # engine.py:571
    c.opt_timestep = 0.02  # point.xml:3 (mj_model.opt.timestep is a double)
    if c.robot in (_abi.GX_ROBOT_ART, _abi.GX_ROBOT_TREE):
        if model is None:
            model = compiled_model(cfg)
        nworld = 2 + 2 * int(cfg['hazards_num'])
        c.opt_timestep = float(model.opt['timestep'])
        c.art_nu, c.art_nv, c.art_nq = int(model.nu), int(model.nv) - nworld, int(model.nq) - nworld
    c.hazards_num = int(cfg['hazards_num'])
    c.lidar_num_bins = int(cfg['lidar_num_bins'])
    c.lidar_alias = 1 if cfg['lidar_alias'] else 0
    c.physics_steps_per_control_step = int(cfg['physics_steps_per_control_step'])
    c.layout_tries = 10  # engine.py:562
    c.pool_size = int(pool_size)
    c.obs_mask = obs_mask_of(cfg)
    c.point = point_model()
    return c
