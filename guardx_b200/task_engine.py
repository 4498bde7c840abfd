"""``TaskEngine`` -- Button / Push / Goal tasks with the Car or Point robot among hazards, vases, pillars and gremlins
(BASELINE.json configs[2]), on the kernels behind ``include/guardx_task.h``.

NO REFERENCE, PARITY UNPINNED BY CONSTRUCTION: nothing of this exists in the reference's JAX Engine.
The semantics are the specification in DESIGN.md section 8 (upstream Safety-Gym's object set under the JAX Engine's
conventions); the checker is ``oracle/task_oracle.py``.  The class keeps the reference Engine's calling convention
(flat config dict merged over ``DEFAULT`` with the ``Bad key`` assertion, ``reset() / step(a) / reset_done()`` returning
torch tensors, ``observation_space`` / ``action_space``, ``info['cost']``) so that the safe_rl_libX algorithms can drive it.
"""
from __future__ import annotations

import ctypes as C
from copy import deepcopy

import numpy as np
import torch

from . import _abi
from .engine import _make_box

GXT_ABI_VERSION = 1
GXT_NCLASS = 8
GOAL, BOX, BUTTON, HAZARD, VASE, PILLAR, GREMLIN, ROBOT = range(8)
TASKS = {'goal': 0, 'button': 1, 'push': 2}
ROBOTS = {'xmls/point.xml': 0, 'xmls/car.xml': 1}
# flat observation = sorted key order (engine.py:773-777)
OBS_COMPONENTS = (("box_compass", 1 << 0), ("box_lidar", 1 << 1), ("buttons_lidar", 1 << 2), ("ctrl", 1 << 3),
                  ("goal_compass", 1 << 4), ("goal_lidar", 1 << 5), ("gremlins_lidar", 1 << 6), ("hazards_lidar", 1 << 7),
                  ("pillars_lidar", 1 << 8), ("qpos", 1 << 9), ("qvel", 1 << 10), ("vases_lidar", 1 << 11))


class GxtConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("env_num", C.c_int32), ("num_steps", C.c_int32), ("task", C.c_int32), ("robot", C.c_int32),
        ("seed", C.c_uint64), ("env_id_offset", C.c_uint64),
        ("count", C.c_int32 * GXT_NCLASS), ("size", C.c_double * GXT_NCLASS), ("keepout", C.c_double * GXT_NCLASS),
        ("placements_extents", C.c_double * 4), ("placements_margin", C.c_double), ("goal_min_robot_dist", C.c_double),
        ("layout_tries", C.c_int32), ("pool_size", C.c_int32),
        ("reward_distance", C.c_double), ("reward_box_dist", C.c_double), ("reward_box_goal", C.c_double),
        ("hazards_cost_scale", C.c_double), ("vases_contact_cost", C.c_double), ("pillars_cost", C.c_double), ("buttons_cost", C.c_double),
        ("gremlins_contact_cost", C.c_double), ("gremlins_dist_threshold", C.c_double), ("gremlins_dist_cost", C.c_double),
        ("gremlins_travel", C.c_double),
        ("lidar_num_bins", C.c_int32), ("lidar_alias", C.c_int32), ("lidar_max_dist", C.c_double), ("lidar_exp_gain", C.c_double),
        ("obs_mask", C.c_uint32),
        ("pt_timestep", C.c_float), ("pt_mass", C.c_float), ("pt_com_x", C.c_float), ("pt_inertia_zz", C.c_float),
        ("pt_damping", C.c_float * 3), ("pt_gear", C.c_float * 3), ("pt_gain", C.c_float), ("pt_bias2", C.c_float),
        ("pt_ctrl", C.c_float * 2), ("pt_force", C.c_float * 2),
        ("car_mass", C.c_float), ("car_inertia", C.c_float), ("car_wheel_r", C.c_float), ("car_half_track", C.c_float),
        ("car_gear", C.c_float), ("car_damp_lin", C.c_float), ("car_damp_ang", C.c_float), ("car_timestep", C.c_float),
    ]


_P, _I = C.c_void_p, C.c_int
PROTOTYPES = {
    "gxt_config_size": (_I, []),
    "gxt_default_config": (None, [C.POINTER(GxtConfig)]),
    "gxt_create": (_I, [C.POINTER(GxtConfig), _I, C.POINTER(_P)]),
    "gxt_destroy": (_I, [_P]),
    "gxt_obs_dim": (_I, [_P]), "gxt_act_dim": (_I, [_P]), "gxt_state_dim": (_I, [_P]), "gxt_layout_dim": (_I, [_P]),
    "gxt_pool_size": (_I, [_P]), "gxt_obs_offset": (_I, [_P, C.c_uint32]),
    "gxt_reset": (_I, [_P, _P, _P, _P]),
    "gxt_step": (_I, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "gxt_reset_done": (_I, [_P, _P, _P, _P, _P]),
    "gxt_get_state": (_I, [_P, _P, _P]), "gxt_set_state": (_I, [_P, _P, _P]),
    "gxt_pool_read": (_I, [_P, _P, _P, _P]),
}
_bound = False


def load_library():
    global _bound
    lib = _abi.load_library()
    if not _bound:
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        if lib.gxt_config_size() != C.sizeof(GxtConfig):
            raise _abi.GuardXError("gxt_config layout mismatch between guardx_task.h and task_engine.py")
        _bound = True
    return lib


# Config surface: the reference Engine's keys that apply (same names, same defaults, engine.py:98-204) plus the object
# keys of upstream Safety-Gym (names and defaults as recalled from safety_gym/envs/engine.py).
DEFAULT = {
    'num_steps': 1000, 'env_num': 1, 'device_id': 0, '_seed': 0,
    'placements_extents': [-2, -2, 2, 2], 'placements_margin': 0.0,
    'robot_base': 'xmls/car.xml', 'robot_keepout': 0.4,
    'task': 'button',
    'observe_goal_lidar': True, 'observe_goal_comp': True, 'observe_box_lidar': True, 'observe_box_comp': True,
    'observe_buttons': True, 'observe_hazards': True, 'observe_vases': True, 'observe_pillars': True, 'observe_gremlins': True,
    'observe_qpos': True, 'observe_qvel': True, 'observe_ctrl': True,
    'lidar_num_bins': 16, 'lidar_max_dist': None, 'lidar_exp_gain': 1.0, 'lidar_alias': True,
    'goal_keepout': 0.305, 'goal_size': 0.3,
    'box_keepout': 0.2, 'box_size': 0.2,
    'reward_distance': 1.0, 'reward_box_dist': 1.0, 'reward_box_goal': 1.0,
    'buttons_num': 4, 'buttons_keepout': 0.2, 'buttons_size': 0.1, 'buttons_cost': 1.0,
    'hazards_num': 4, 'hazards_keepout': 0.18, 'hazards_size': 0.2, 'hazards_cost': 1.0,
    'vases_num': 4, 'vases_keepout': 0.15, 'vases_size': 0.1, 'vases_contact_cost': 1.0,
    'pillars_num': 2, 'pillars_keepout': 0.3, 'pillars_size': 0.2, 'pillars_cost': 1.0,
    'gremlins_num': 4, 'gremlins_keepout': 0.5, 'gremlins_size': 0.1, 'gremlins_travel': 0.3,
    'gremlins_contact_cost': 1.0, 'gremlins_dist_threshold': 0.2, 'gremlins_dist_cost': 1.0,
    'robot_size': None,               # contact radius; None: 0.1 (Point's sphere) / 0.2 (Car's half length)
    'goal_min_robot_dist': None,      # None: 3.0 for the goal task (engine.py:571), 0 otherwise
}


def configuration(task: str) -> dict:
    """``<Task>_<Robot>_Level2``-style names in the spirit of safe_rl_env_config.configuration_list:
    Goal / Button / Push x Point / Car, with every object class present."""
    parts = task.split('_')
    assert len(parts) >= 2 and parts[0].lower() in TASKS and parts[1] in ('Point', 'Car'), \
        f"task {task!r}: expected <Goal|Button|Push>_<Point|Car>[_...]"
    return {'task': parts[0].lower(), 'robot_base': f'xmls/{parts[1].lower()}.xml'}


def parse(config: dict) -> dict:
    merged = deepcopy(DEFAULT)
    merged.update(deepcopy(config))
    for key in merged:
        assert key in DEFAULT, f'Bad key {key}'
    return merged


def lower(cfg: dict, env_id_offset: int = 0, pool_size: int = 0) -> GxtConfig:
    lib = load_library()
    c = GxtConfig()
    lib.gxt_default_config(C.byref(c))           # Point / Car model constants
    c.env_num, c.num_steps = int(cfg['env_num']), int(cfg['num_steps'])
    c.task, c.robot = TASKS[cfg['task']], ROBOTS[cfg['robot_base']]
    c.seed, c.env_id_offset = int(cfg['_seed']) & 0xFFFFFFFFFFFFFFFF, int(env_id_offset)
    task = cfg['task']
    counts = {GOAL: 0 if task == 'button' else 1, BOX: 1 if task == 'push' else 0, BUTTON: int(cfg['buttons_num']) if task == 'button' else 0,
              HAZARD: int(cfg['hazards_num']), VASE: int(cfg['vases_num']), PILLAR: int(cfg['pillars_num']),
              GREMLIN: int(cfg['gremlins_num']), ROBOT: 1}
    rs = cfg['robot_size'] if cfg['robot_size'] is not None else (0.2 if c.robot == 1 else 0.1)
    sizes = {GOAL: cfg['goal_size'], BOX: cfg['box_size'], BUTTON: cfg['buttons_size'], HAZARD: cfg['hazards_size'],
             VASE: cfg['vases_size'], PILLAR: cfg['pillars_size'], GREMLIN: cfg['gremlins_size'], ROBOT: rs}
    keeps = {GOAL: cfg['goal_keepout'], BOX: cfg['box_keepout'], BUTTON: cfg['buttons_keepout'], HAZARD: cfg['hazards_keepout'],
             VASE: cfg['vases_keepout'], PILLAR: cfg['pillars_keepout'], GREMLIN: cfg['gremlins_keepout'], ROBOT: cfg['robot_keepout']}
    for k in range(GXT_NCLASS):
        c.count[k], c.size[k], c.keepout[k] = counts[k], float(sizes[k]), float(keeps[k])
    c.placements_extents[:] = [float(v) for v in cfg['placements_extents']]
    c.placements_margin = float(cfg['placements_margin'])
    gm = cfg['goal_min_robot_dist']
    c.goal_min_robot_dist = float(gm) if gm is not None else (3.0 if task == 'goal' else 0.0)
    c.layout_tries, c.pool_size = 10, int(pool_size)
    c.reward_distance, c.reward_box_dist, c.reward_box_goal = (float(cfg[k]) for k in ('reward_distance', 'reward_box_dist', 'reward_box_goal'))
    c.hazards_cost_scale = float(cfg['hazards_cost'])
    c.vases_contact_cost, c.pillars_cost, c.buttons_cost = float(cfg['vases_contact_cost']), float(cfg['pillars_cost']), float(cfg['buttons_cost'])
    c.gremlins_contact_cost, c.gremlins_dist_threshold = float(cfg['gremlins_contact_cost']), float(cfg['gremlins_dist_threshold'])
    c.gremlins_dist_cost, c.gremlins_travel = float(cfg['gremlins_dist_cost']), float(cfg['gremlins_travel'])
    c.lidar_num_bins, c.lidar_alias = int(cfg['lidar_num_bins']), 1 if cfg['lidar_alias'] else 0
    c.lidar_max_dist = 0.0 if cfg['lidar_max_dist'] is None else float(cfg['lidar_max_dist'])
    c.lidar_exp_gain = float(cfg['lidar_exp_gain'])
    flags = {'box_compass': 'observe_box_comp', 'box_lidar': 'observe_box_lidar', 'buttons_lidar': 'observe_buttons', 'ctrl': 'observe_ctrl',
             'goal_compass': 'observe_goal_comp', 'goal_lidar': 'observe_goal_lidar', 'gremlins_lidar': 'observe_gremlins',
             'hazards_lidar': 'observe_hazards', 'pillars_lidar': 'observe_pillars', 'qpos': 'observe_qpos', 'qvel': 'observe_qvel',
             'vases_lidar': 'observe_vases'}
    c.obs_mask = sum(bit for name, bit in OBS_COMPONENTS if cfg[flags[name]])
    return c


class TaskEngine:
    DEFAULT = DEFAULT

    def __init__(self, config={}, *, device=None, env_id_offset: int = 0, pool_size: int = 0):
        self.config = parse(config)
        for k, v in self.config.items():
            setattr(self, k, v)
        self._ctor = (dict(config), dict(device=device, env_id_offset=env_id_offset, pool_size=pool_size))
        if not torch.cuda.is_available():
            raise _abi.GuardXError("guardx_b200.TaskEngine needs a CUDA device (sm_100a); there is no CPU path")
        self._lib = load_library()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else torch.device(device).index or 0)
        self._gxt_config = lower(self.config, env_id_offset, pool_size)
        h = C.c_void_p()
        _abi.check(self._lib, self._lib.gxt_create(C.byref(self._gxt_config), self.device.index, C.byref(h)), "gxt_create")
        self._h = h
        L = self._lib
        self.obs_flat_size, self.act_dim = L.gxt_obs_dim(h), L.gxt_act_dim(h)
        self.state_dim, self.layout_dim, self.layout_size = L.gxt_state_dim(h), L.gxt_layout_dim(h), L.gxt_pool_size(h)
        self.action_space = _make_box(-np.ones(self.act_dim, np.float32), np.ones(self.act_dim, np.float32))
        self.observation_space = _make_box(-np.inf, np.inf, (self.obs_flat_size,))
        nb, nu = int(self.lidar_num_bins), (2 if self._gxt_config.robot == 1 else 3)
        width = {'box_compass': 2, 'ctrl': nu, 'goal_compass': 2, 'qpos': 3, 'qvel': 3}
        self._obs_slices = {}
        for name, bit in OBS_COMPONENTS:
            off = L.gxt_obs_offset(h, bit)
            if off >= 0:
                self._obs_slices[name] = (off, off + width.get(name, nb))
        self._obs = self._done = None
        self._pool_checked = False

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _empty(self, *shape, dtype=torch.float32):
        return torch.empty(shape, dtype=dtype, device=self.device)

    @staticmethod
    def _ptr(t):
        return None if t is None else C.c_void_p(t.data_ptr())

    def _as_input(self, t, shape):
        if not isinstance(t, torch.Tensor):
            t = torch.as_tensor(np.asarray(t), dtype=torch.float32)
        t = t.to(device=self.device, dtype=torch.float32).contiguous()
        assert tuple(t.shape) == tuple(shape), f'expected shape {tuple(shape)}, got {tuple(t.shape)}'
        return t

    def reset(self, layout=None):
        obs = self._empty(self.env_num, self.obs_flat_size)
        lay = None if layout is None else self._as_input(layout, (self.env_num, self.layout_dim))
        _abi.check(self._lib, self._lib.gxt_reset(self._h, self._ptr(lay), self._ptr(obs), self._stream()), "gxt_reset")
        if layout is None and not self._pool_checked:
            # engine.py:443-444 asserts that enough valid layouts exist; here: every pool slot found one
            _, tries = self.layout_pool()
            if int(tries.min().item()) <= 0:
                from .engine import ResamplingError
                raise ResamplingError('no valid layout found: too many objects / keepouts for placements_extents')
            self._pool_checked = True
        self._obs = obs
        return obs

    def _step(self, action, autoreset, want_aux):
        action = self._as_input(action, (self.env_num, self.act_dim))
        n = self.env_num
        obs, rdc = self._empty(n, self.obs_flat_size), self._empty(3, n)
        aux = self._empty(n, 4, dtype=torch.int32) if want_aux else None
        _abi.check(self._lib, self._lib.gxt_step(self._h, self._ptr(action), self._ptr(obs), self._ptr(rdc[0]), self._ptr(rdc[1]),
                                                 self._ptr(rdc[2]), self._ptr(aux), 1 if autoreset else 0, self._stream()), "gxt_step")
        info = {'cost': rdc[2]}
        if want_aux:
            info['aux'] = aux
        self._obs, self._done = obs, rdc[1]
        return obs, rdc[0], rdc[1], info

    def step(self, action, want_aux: bool = False):
        return self._step(action, False, want_aux)

    def step_autoreset(self, action, want_aux: bool = False):
        return self._step(action, True, want_aux)

    def reset_done(self, layout=None):
        assert self._obs is not None, 'reset_done() before reset()'
        out = self._empty(self.env_num, self.obs_flat_size)
        lay = None if layout is None else self._as_input(layout, (self.env_num, self.layout_dim))
        _abi.check(self._lib, self._lib.gxt_reset_done(self._h, self._ptr(self._obs), self._ptr(out), self._ptr(lay), self._stream()),
                   "gxt_reset_done")
        self._obs = out
        return out

    def get_state(self):
        s = self._empty(self.env_num, self.state_dim)
        _abi.check(self._lib, self._lib.gxt_get_state(self._h, self._ptr(s), self._stream()), "gxt_get_state")
        return s

    def set_state(self, state):
        s = self._as_input(state, (self.env_num, self.state_dim))
        _abi.check(self._lib, self._lib.gxt_set_state(self._h, self._ptr(s), self._stream()), "gxt_set_state")

    def layout_pool(self):
        pool = self._empty(self.layout_size, self.layout_dim)
        tries = self._empty(self.layout_size, dtype=torch.int32)
        _abi.check(self._lib, self._lib.gxt_pool_read(self._h, self._ptr(pool), self._ptr(tries), self._stream()), "gxt_pool_read")
        return pool, tries

    def render(self, *a, **k):
        raise NotImplementedError('render() is out of scope')

    def close(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._lib.gxt_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __reduce__(self):
        return (_rebuild, self._ctor)


def _rebuild(config, kwargs):
    return TaskEngine(config, **kwargs)


def make_task(task: str = "Button_Car_Level2", env_num: int = 1, seed: int = 0, max_ep_len: int = 1000, **overrides):
    cfg = configuration(task)
    cfg.update(env_num=env_num, _seed=seed, num_steps=max_ep_len)
    engine_kwargs = {k: overrides.pop(k) for k in ('device', 'env_id_offset', 'pool_size') if k in overrides}
    cfg.update(overrides)
    return TaskEngine(cfg, **engine_kwargs)
