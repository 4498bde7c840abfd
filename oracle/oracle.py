"""numpy-facing wrapper of the CPU oracle (oracle/gx_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by guardx_b200.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from guardx_b200 import _abi      # ctypes mirror of include/guardx.h only (the struct IS the boundary)
from . import oracle_config as gxconfig

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgxoracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gx_oracle.c")
    deps = [src, os.path.join(_HERE, "..", "include", "guardx.h")] + [
        os.path.join(_HERE, "..", "guardx_b200", "csrc", h)
        for h in ("gx_math.h", "gx_philox.h", "gx_params.h")]
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.gxo_last_error.restype = C.c_char_p
        for name in ("gxo_create", "gxo_destroy", "gxo_reset", "gxo_reset_with_layout", "gxo_step",
                     "gxo_reset_done", "gxo_get_state", "gxo_set_state", "gxo_pool_read",
                     "gxo_rollout", "gxo_stats_read", "gxo_stats_reset", "gxo_obs_dim",
                     "gxo_aux_dim", "gxo_state_dim", "gxo_pool_size", "gxo_max_threads",
                     "gxo_config_size"):
            getattr(_lib, name).restype = C.c_int
        _lib.gxo_set_counters.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        _lib.gxo_set_counters.restype = None
        assert _lib.gxo_config_size() == C.sizeof(_abi.GxConfig)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class OracleEngine:
    """Same call protocol as guardx_b200.Engine, numpy arrays on the host."""

    def __init__(self, config: dict | None = None, env_id_offset: int = 0, pool_size: int = 0,
                 threads: int = 1, gx_config: _abi.GxConfig | None = None):
        self.L = lib()
        if gx_config is None:
            self.config = gxconfig.parse(config or {})
            gx_config = gxconfig.lower(self.config, env_id_offset, pool_size)
        self.gx_config = gx_config
        h = C.c_void_p()
        rc = self.L.gxo_create(C.byref(gx_config), C.byref(h))
        if rc != 0:
            raise AssertionError(self.L.gxo_last_error().decode())
        self.h = h
        self.n = gx_config.env_num
        self.obs_dim = self.L.gxo_obs_dim(h)
        self.aux_dim = self.L.gxo_aux_dim(h)
        self.state_dim = self.L.gxo_state_dim(h)
        self.pool_size = self.L.gxo_pool_size(h)
        self.lay_dim = 2 * (2 + gx_config.hazards_num)
        self.set_threads(threads)
        self._obs = None

    def set_threads(self, t: int):
        self.L.gxo_set_threads(self.h, C.c_int(int(t)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.gxo_destroy(self.h)
            self.h = None

    def _check(self, rc):
        if rc != 0:
            raise AssertionError(self.L.gxo_last_error().decode())

    def reset(self, layout=None):
        obs = np.empty((self.n, self.obs_dim), np.float32)
        if layout is None:
            self._check(self.L.gxo_reset(self.h, _p(obs)))
        else:
            layout = _f32(layout)
            assert layout.shape == (self.n, self.lay_dim)
            self._check(self.L.gxo_reset_with_layout(self.h, _p(layout), _p(obs)))
        self._obs = obs
        return obs

    def step(self, action, want_qacc=False, want_aux=False):
        action = _f32(action)
        assert action.shape == (self.n, 2)
        obs = np.empty((self.n, self.obs_dim), np.float32)
        reward = np.empty(self.n, np.float32)
        done = np.empty(self.n, np.float32)
        cost = np.empty(self.n, np.float32)
        qacc = np.empty((self.n, 3), np.float32) if want_qacc else None
        aux = np.empty((self.n, self.aux_dim), np.int32) if want_aux else None
        self._check(self.L.gxo_step(self.h, _p(action), _p(obs), _p(reward), _p(done), _p(cost),
                                    _p(qacc), _p(aux)))
        self._obs = obs
        info = {"cost": cost}
        if want_qacc:
            info["qacc"] = qacc
        if want_aux:
            info["aux"] = aux
        return obs, reward, done, info

    def reset_done(self, layout=None):
        out = np.empty((self.n, self.obs_dim), np.float32)
        lay = None if layout is None else _f32(layout)
        self._check(self.L.gxo_reset_done(self.h, _p(self._obs), _p(out), _p(lay)))
        self._obs = out
        return out

    def get_state(self):
        s = np.empty((self.n, self.state_dim), np.float32)
        self._check(self.L.gxo_get_state(self.h, _p(s)))
        return s

    def set_state(self, s):
        s = _f32(s)
        assert s.shape == (self.n, self.state_dim)
        self._check(self.L.gxo_set_state(self.h, _p(s)))

    def set_counters(self, t_global: int, epoch: int):
        self.L.gxo_set_counters(self.h, t_global, epoch)

    def pool(self):
        pool = np.empty((self.pool_size, self.lay_dim), np.float32)
        tries = np.empty(self.pool_size, np.int32)
        self._check(self.L.gxo_pool_read(self.h, _p(pool), _p(tries)))
        return pool, tries

    def rollout(self, T, want=("obs", "act", "reward", "cost", "done")):
        n = self.n
        out = {}
        if "obs" in want:
            out["obs"] = np.empty((T, n, self.obs_dim), np.float32)
        if "act" in want:
            out["act"] = np.empty((T, n, 2), np.float32)
        for k in ("reward", "cost", "done"):
            if k in want:
                out[k] = np.empty((T, n), np.float32)
        self._check(self.L.gxo_rollout(self.h, C.c_int32(T), _p(out.get("obs")), _p(out.get("act")),
                                       _p(out.get("reward")), _p(out.get("cost")), _p(out.get("done"))))
        if "obs" in out:
            self._obs = out["obs"][-1].copy()
        return out

    def stats(self):
        s = np.zeros(8, np.float64)
        self._check(self.L.gxo_stats_read(self.h, _p(s)))
        return s

    def stats_reset(self):
        self._check(self.L.gxo_stats_reset(self.h))
