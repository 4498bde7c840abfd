"""``Engine`` -- the reference env API on top of libguardx.so.

Mirrors ``safe_rl_envs.envs.engine.Engine`` (reference engine.py:84-505): same
constructor (a flat config dict merged over ``DEFAULT``, unknown key ->
``AssertionError('Bad key ...')``), same ``reset() / step(a) / reset_done()``
protocol returning torch tensors on the simulation device, same
``observation_space`` / ``action_space`` attributes, same flat observation
layout (sorted key order, engine.py:773-777).

Host code is Python; all arithmetic happens in the CUDA kernels behind the C ABI
of ``include/guardx.h``.  There is no CPU fallback.
"""
from __future__ import annotations

import collections.abc
import ctypes as C

import numpy as np
import torch

from . import _abi
from . import config as gxconfig

try:  # the reference uses gym.spaces.Box (engine.py:296, 393-413)
    from gym.spaces import Box as _GymBox  # type: ignore
except Exception:  # gym is not installed in this image
    _GymBox = None


class Box:
    """Minimal stand-in for ``gym.spaces.Box`` (gym is absent here).  Callers of
    the reference only read ``.shape`` (ppo_train.py:263-267) and test
    ``isinstance(space, Box)`` (ppo_core.py:131)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            low = np.asarray(low, dtype=dtype)
            high = np.asarray(high, dtype=dtype)
            shape = low.shape
        else:
            low = np.full(shape, low, dtype=dtype)
            high = np.full(shape, high, dtype=dtype)
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def _make_box(low, high, shape=None):
    if _GymBox is not None:
        if shape is None:
            return _GymBox(np.asarray(low, np.float32), np.asarray(high, np.float32), dtype=np.float32)
        return _GymBox(low, high, shape, dtype=np.float32)
    return Box(low, high, shape)


class _ObsDict(collections.abc.Mapping):
    """``info['obs']`` (engine.py:693-695): per-key views into the flat observation
    plus ``qacc``.  Views are created on access so that the per-step host cost is
    not paid by callers that never look (none of the shipped algorithms do)."""

    def __init__(self, obs, slices, qacc):
        self._obs, self._slices, self._qacc = obs, slices, qacc

    def __getitem__(self, k):
        if k == 'qacc' and self._qacc is not None:
            return self._qacc
        a, b = self._slices[k]
        return self._obs[:, a:b]

    def __iter__(self):
        yield from self._slices
        if self._qacc is not None:
            yield 'qacc'

    def __len__(self):
        return len(self._slices) + (self._qacc is not None)


class _Robot:
    """reference world.py:422-445, the attributes Engine reads."""

    def __init__(self, nq, nv, nu, z_height):
        self.nq, self.nv, self.nu, self.z_height = nq, nv, nu, z_height


class ResamplingError(AssertionError):
    ''' Raised when we fail to sample a valid distribution of objects or goals '''


class Engine:
    DEFAULT = gxconfig.DEFAULT

    def __init__(self, config={}, *, device=None, env_id_offset: int = 0, pool_size: int = 0):
        self.parse(config)
        self._ctor_config = dict(config)
        self._ctor_kwargs = dict(device=device, env_id_offset=env_id_offset, pool_size=pool_size)
        if not torch.cuda.is_available():
            raise _abi.GuardXError("guardx_b200.Engine needs a CUDA device (sm_100a); there is no CPU path")
        self._lib = _abi.load_library()
        # the reference ignores device_id (engine.py:100) and runs on JAX's default
        # device; here: explicit `device`, else torch's current CUDA device
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None
                                   else torch.device(device).index or 0)
        # World.build + mujoco.MjModel.from_xml_string + mjx.put_model (engine.py:220-232):
        # the Point's constants are closed-form; the other robots go through the model compiler
        self.mj_model = gxconfig.compiled_model(self.config)
        self._gx_config = gxconfig.lower(self.config, env_id_offset, pool_size, self.mj_model)
        h = C.c_void_p()
        if self.mj_model is None:
            _abi.check(self._lib, self._lib.gx_create(C.byref(self._gx_config), self.device.index, C.byref(h)),
                       "gx_create")
            self._art_info = None
        elif self._gx_config.robot == _abi.GX_ROBOT_TREE:
            from . import tree_model
            self._art_model, self._art_info = tree_model.lower_tree(self.mj_model)
            _abi.check(self._lib, self._lib.gx_create_tree(C.byref(self._gx_config), C.byref(self._art_model),
                                                           self.device.index, C.byref(h)), "gx_create_tree")
        else:
            from . import art_model
            self._art_model, self._art_info = art_model.lower_art(self.mj_model)
            _abi.check(self._lib, self._lib.gx_create_art(C.byref(self._gx_config), C.byref(self._art_model),
                                                          self.device.index, C.byref(h)), "gx_create_art")
        self._h = h
        L = self._lib
        self.obs_flat_size = L.gx_obs_dim(h)
        self.act_dim = L.gx_act_dim(h)
        self.aux_dim = L.gx_aux_dim(h)
        self.state_dim = L.gx_state_dim(h)
        self.layout_dim = 2 * (2 + int(self.hazards_num))
        self.layout_size = L.gx_pool_size(h)  # reference: number of valid layouts (engine.py:442)
        if self._art_info is None:
            self.robot = _Robot(3, 3, 3, 0.1)  # point.xml: nq = nv = nu = 3, body z 0.1
        else:
            self.robot = _Robot(self._art_info.get('nq', self._art_info['nv']), self._art_info['nv'], self._art_info['nu'],
                                self._art_info['z_height'])
        self.dt = float(self._gx_config.opt_timestep) * self.physics_steps_per_control_step
        lo = (C.c_float * self.act_dim)()
        hi = (C.c_float * self.act_dim)()
        _abi.check(L, L.gx_action_bounds(h, lo, hi), "gx_action_bounds")
        self.action_space = _make_box(np.array(lo[:], np.float32), np.array(hi[:], np.float32))
        self.build_observation_space()
        self._obs = None
        self._reward = None
        self._done = None
        self._info = None
        self._pool_checked = False
        #: emit info['obs'] (incl. qacc) like the reference does; costs 12 B/env-step
        self.emit_info_obs = True
        #: reset_done() writes done rows into the tensor step() returned instead of a fresh one
        self.inplace_reset_done = False

    # ---------------------------------------------------------------- config
    def parse(self, config):
        ''' Parse a config dict - see self.DEFAULT for description (engine.py:322-328) '''
        self.config = gxconfig.parse(config)
        for key, value in self.config.items():
            setattr(self, key, value)

    def build_observation_space(self):
        ''' engine.py:386-418 '''
        self.obs_space_dict = collections.OrderedDict()
        self._obs_slices = {}
        nb = int(self.lidar_num_bins)
        sizes = {'acc': 2, 'ctrl': self.robot.nu, 'goal_compass': 2, 'goal_lidar': nb,
                 'hazards_lidar': nb, 'qpos': self.robot.nq, 'qvel': self.robot.nv, 'vel': 2}
        for name, bit in _abi.OBS_COMPONENTS:
            if not (self._gx_config.obs_mask & bit):
                continue
            off = self._lib.gx_obs_offset(self._h, bit)
            assert off >= 0
            self._obs_slices[name] = (off, off + sizes[name])
            if 'lidar' in name:
                self.obs_space_dict[name] = _make_box(0.0, 1.0, (sizes[name],))
            else:
                self.obs_space_dict[name] = _make_box(-np.inf, np.inf, (sizes[name],))
        assert sum(b - a for a, b in self._obs_slices.values()) == self.obs_flat_size
        self.observation_space = _make_box(-np.inf, np.inf, (self.obs_flat_size,))

    # ------------------------------------------------------------- plumbing
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

    def _check_pool(self):
        """engine.py:443-444 asserts that enough valid layouts exist; here: every
        pool slot found one (checked once, the only host sync of reset())."""
        if self._pool_checked:
            return
        tries = self._empty(self.layout_size, dtype=torch.int32)
        _abi.check(self._lib, self._lib.gx_pool_read(self._h, None, self._ptr(tries), self._stream()), "gx_pool_read")
        if int(tries.min().item()) <= 0:
            raise ResamplingError('no valid layout found: placements_extents / keepouts cannot satisfy '
                                  'the robot-goal distance of engine.py:571')
        self._pool_checked = True

    # ------------------------------------------------------------------ API
    def reset(self, layout=None):
        ''' Reset the physics simulation and return observation (engine.py:454-467).
        `layout` (optional, [env_num, 2*(2+hazards_num)]: goal, robot, hazards xy)
        replaces the pool draw -- the reference's `self._reset(layout)`. '''
        obs = self._empty(self.env_num, self.obs_flat_size)
        L = self._lib
        if layout is None:
            _abi.check(L, L.gx_reset(self._h, self._ptr(obs), self._stream()), "gx_reset")
            self._check_pool()
        else:
            lay = self._as_input(layout, (self.env_num, self.layout_dim))
            _abi.check(L, L.gx_reset_with_layout(self._h, self._ptr(lay), self._ptr(obs), self._stream()),
                       "gx_reset_with_layout")
        self._obs = obs
        return obs

    def step(self, action, want_aux: bool = False):
        ''' Take a step and return observation, reward, done, and info (engine.py:469-495) '''
        action = self._as_input(action, (self.env_num, self.act_dim))
        n = self.env_num
        obs = self._empty(n, self.obs_flat_size)
        rdc = self._empty(3, n)
        reward, done, cost = rdc[0], rdc[1], rdc[2]
        qacc = self._empty(n, self.robot.nv) if self.emit_info_obs else None
        aux = self._empty(n, self.aux_dim, dtype=torch.int32) if want_aux else None
        L = self._lib
        _abi.check(L, L.gx_step(self._h, self._ptr(action), self._ptr(obs), self._ptr(reward), self._ptr(done),
                                self._ptr(cost), self._ptr(qacc), self._ptr(aux), self._stream()), "gx_step")
        info = {'cost': cost}
        if self.emit_info_obs:
            info['obs'] = _ObsDict(obs, self._obs_slices, qacc)
        if want_aux:
            info['aux'] = aux
        self._obs, self._reward, self._done, self._info = obs, reward, done, info
        return obs, reward, done, info

    def step_autoreset(self, action):
        ''' step() and reset_done() fused into one kernel launch: returns
        (obs after re-placing finished envs, reward, done, info) -- the same
        tensors `step(a)` followed by `reset_done()` would hand the caller. '''
        action = self._as_input(action, (self.env_num, self.act_dim))
        n = self.env_num
        obs = self._empty(n, self.obs_flat_size)
        rdc = self._empty(3, n)
        L = self._lib
        _abi.check(L, L.gx_step_autoreset(self._h, self._ptr(action), self._ptr(obs), self._ptr(rdc[0]),
                                          self._ptr(rdc[1]), self._ptr(rdc[2]), self._stream()),
                   "gx_step_autoreset")
        info = {'cost': rdc[2]}
        self._obs, self._reward, self._done, self._info = obs, rdc[0], rdc[1], info
        return obs, rdc[0], rdc[1], info

    def reset_done(self, layout=None):
        ''' Re-place the envs whose last `done` was 1 and return the full
        observation batch with their rows replaced (engine.py:497-505). '''
        assert self._obs is not None, 'reset_done() before reset()'
        out = self._obs if self.inplace_reset_done else self._empty(self.env_num, self.obs_flat_size)
        lay = None if layout is None else self._as_input(layout, (self.env_num, self.layout_dim))
        L = self._lib
        _abi.check(L, L.gx_reset_done(self._h, self._ptr(self._obs), self._ptr(out), self._ptr(lay), self._stream()),
                   "gx_reset_done")
        self._obs = out
        return out

    def done_indices(self):
        """(idx, count) of the envs whose last `done` was 1, on the device (rollout_ops.done_indices):
        what SCPO's ``env.reset_done()[np.where(done == 1)]`` (scpo/scpo.py:722) needs, without the
        device->host copy of `done`."""
        assert self._done is not None, 'done_indices() before the first step()'
        from . import rollout_ops
        return rollout_ops.done_indices(self._done)

    def rollout(self, T: int, want=("obs", "act", "reward", "cost", "done")):
        ''' T steps of a ~ U(-1,1) (as trpo/myTest.py:31 drives the reference) with
        reset_done after every step, in one kernel launch. '''
        n = self.env_num
        out = {}
        if "obs" in want:
            out["obs"] = self._empty(T, n, self.obs_flat_size)
        if "act" in want:
            out["act"] = self._empty(T, n, self.act_dim)
        for k in ("reward", "cost", "done"):
            if k in want:
                out[k] = self._empty(T, n)
        L = self._lib
        _abi.check(L, L.gx_rollout(self._h, T, self._ptr(out.get("obs")), self._ptr(out.get("act")),
                                   self._ptr(out.get("reward")), self._ptr(out.get("cost")),
                                   self._ptr(out.get("done")), self._stream()), "gx_rollout")
        # Engine._obs / _done / _reward describe the LAST step: keep them current when the rollout
        # produced them, otherwise invalidate them so that a following reset_done() / done_indices()
        # asserts instead of working on pre-rollout tensors
        self._obs = out["obs"][-1] if "obs" in out else None
        self._done = out["done"][-1] if "done" in out else None
        self._reward = out["reward"][-1] if "reward" in out else None
        self._info = None
        return out

    # ------------------------------------------------- state / diagnostics
    def get_state(self):
        s = self._empty(self.env_num, self.state_dim)
        _abi.check(self._lib, self._lib.gx_get_state(self._h, self._ptr(s), self._stream()), "gx_get_state")
        return s

    def set_state(self, state):
        s = self._as_input(state, (self.env_num, self.state_dim))
        _abi.check(self._lib, self._lib.gx_set_state(self._h, self._ptr(s), self._stream()), "gx_set_state")

    def counters(self):
        """(t_global, epoch): env steps taken and layout pools generated -- the two
        device-resident counters that key every random draw.  Synchronises."""
        t, e = C.c_uint64(), C.c_uint64()
        _abi.check(self._lib, self._lib.gx_get_counters(self._h, C.byref(t), C.byref(e), self._stream()),
                   "gx_get_counters")
        return int(t.value), int(e.value)

    def set_counters(self, t_global: int, epoch: int):
        _abi.check(self._lib, self._lib.gx_set_counters(self._h, t_global, epoch, self._stream()),
                   "gx_set_counters")

    def layout_pool(self):
        pool = self._empty(self.layout_size, self.layout_dim)
        tries = self._empty(self.layout_size, dtype=torch.int32)
        _abi.check(self._lib, self._lib.gx_pool_read(self._h, self._ptr(pool), self._ptr(tries), self._stream()),
                   "gx_pool_read")
        return pool, tries

    def stats_reset(self):
        _abi.check(self._lib, self._lib.gx_stats_reset(self._h, self._stream()), "gx_stats_reset")

    def stats(self):
        """device tensor of 8 doubles: n_episodes, sum_ret, sum_cost, sum_len, sum_ret^2,
        sum_cost^2, total_cost, total_env_steps (since stats_reset)."""
        s = torch.empty(8, dtype=torch.float64, device=self.device)
        _abi.check(self._lib, self._lib.gx_stats_read(self._h, self._ptr(s), self._stream()), "gx_stats_read")
        return s

    def step_host(self, h_action, h_out, auto_reset_done: bool = False):
        """End-to-end step with HOST buffers: `h_action` pinned [N, act_dim];
        `h_out` pinned [N*(obs_dim+3)] receiving obs | reward | done | cost.
        Synchronises the stream."""
        n, od = self.env_num, self.obs_flat_size
        base = h_out.data_ptr()
        p = lambda off: C.c_void_p(base + 4 * off)
        L = self._lib
        _abi.check(L, L.gx_step_host(self._h, C.c_void_p(h_action.data_ptr()), p(0), p(n * od), p(n * od + n),
                                     p(n * od + 2 * n), 1 if auto_reset_done else 0, self._stream()), "gx_step_host")

    def rollout_host(self, h_action, h_obs, h_reward, h_done, h_cost):
        """T pipelined end-to-end steps with HOST buffers (gx_rollout_host): `h_action`
        pinned [T, N, act_dim]; outputs pinned [T, N, obs_dim] and [T, N].  Same results
        as T calls of step_host(auto_reset_done=True); returns when they are in host memory."""
        T = int(h_action.shape[0])
        n, od = self.env_num, self.obs_flat_size
        assert tuple(h_action.shape) == (T, n, self.action_space.shape[0]) and h_action.is_contiguous()
        assert tuple(h_obs.shape) == (T, n, od) and h_obs.is_contiguous()
        for h in (h_reward, h_done, h_cost):
            assert tuple(h.shape) == (T, n) and h.is_contiguous()
        p = lambda t: C.c_void_p(t.data_ptr())
        L = self._lib
        _abi.check(L, L.gx_rollout_host(self._h, T, p(h_action), p(h_obs), p(h_reward), p(h_done), p(h_cost),
                                        self._stream()), "gx_rollout_host")

    # ------------------------------------------------------------ lifetime
    def render(self, *a, **k):
        raise NotImplementedError('render() needs a MuJoCo GL viewer (engine.py:1036-1070); out of scope')

    def close(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            self._lib.gx_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __reduce__(self):
        # EzPickle(config=config) (engine.py:214): pickles as constructor args, not state
        return (_rebuild, (self._ctor_config, self._ctor_kwargs))


def _rebuild(config, kwargs):
    return Engine(config, **kwargs)
