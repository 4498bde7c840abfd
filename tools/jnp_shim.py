"""A numpy stand-in for the parts of ``jax`` / ``jax.numpy`` that the reference's
observation, reward, cost and layout code uses (engine.py:546-621, 738-929).

Purpose: execute the reference's OWN, UNMODIFIED source for those functions in a
container that has no jax / mujoco, to generate golden vectors that pin the CPU
oracle (tools/make_golden.py).  It reproduces the JAX semantics that matter for
results:
  * default dtype float32 / int32; Python scalars are weakly typed (a Python
    float never widens a float32 array; int array op Python float -> float32);
  * ``x[i]`` with an out-of-range integer index clamps (gather mode "clip"),
    negative indices wrap;
  * ``x.at[i].set(v)`` returns a copy, out-of-range scatter is dropped;
  * ``%`` is floored remainder (np.remainder has the same definition).
It does NOT reproduce XLA's elementary-function rounding (numpy's exp/arctan2 are
used), so float comparisons against these vectors carry a ~1e-6 tolerance.
The physics (mjx.step / mjx.forward) is not emulated at all.
"""
from __future__ import annotations

import sys
import types

import numpy as np


def _is_float(dt):
    return np.issubdtype(dt, np.floating)


def _coerce(x):
    """Inputs of an op -> plain ndarrays / scalars with JAX-style typing."""
    if isinstance(x, JArr):
        return np.asarray(x)
    if isinstance(x, np.ndarray):
        if x.dtype == np.float64:
            return x.astype(np.float32)
        if x.dtype == np.int64:
            return x.astype(np.int32)
        return x
    if isinstance(x, np.generic):
        return _coerce(np.asarray(x))
    return x  # python scalar: weak


def _promote(args):
    arrs = [a for a in args if isinstance(a, np.ndarray)]
    any_float = any(_is_float(a.dtype) for a in arrs) or any(isinstance(a, float) for a in args)
    out = []
    for a in args:
        if isinstance(a, np.ndarray):
            if any_float and not _is_float(a.dtype) and a.dtype != np.bool_:
                a = a.astype(np.float32)
            out.append(a)
        elif isinstance(a, bool):
            out.append(a)
        elif isinstance(a, float):
            out.append(np.float32(a))
        elif isinstance(a, int):
            out.append(np.float32(a) if any_float else np.int32(a))
        else:
            out.append(a)
    return out


def _wrap(r):
    if isinstance(r, tuple):
        return tuple(_wrap(v) for v in r)
    r = np.asarray(r)
    if r.dtype == np.float64:
        r = r.astype(np.float32)
    elif r.dtype == np.int64:
        r = r.astype(np.int32)
    return r.view(JArr)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _AtIdx(self.arr, idx)


class _AtIdx:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, v):
        out = np.array(self.arr, copy=True)
        idx = self.idx
        if isinstance(idx, (JArr, np.ndarray, np.generic)):
            idx = int(np.asarray(idx))
        n = out.shape[0]
        if idx < 0:
            idx += n
        if 0 <= idx < n:  # out-of-bounds scatter is dropped
            out[idx] = np.asarray(v, dtype=out.dtype)
        return out.view(JArr)


class JArr(np.ndarray):
    __array_priority__ = 100

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        args = _promote([_coerce(x) for x in inputs])
        kwargs.pop('out', None)
        return _wrap(getattr(ufunc, method)(*args, **kwargs))

    def __getitem__(self, idx):
        base = np.asarray(self)
        if isinstance(idx, (JArr, np.generic)) or (isinstance(idx, np.ndarray) and idx.ndim == 0):
            idx = int(np.asarray(idx))
        if isinstance(idx, (int, np.integer)) and base.ndim >= 1:
            n = base.shape[0]
            i = int(idx)
            if i < 0:
                i += n
            i = min(max(i, 0), n - 1)  # gather clamps
            return _wrap(base[i])
        if isinstance(idx, list):
            idx = np.asarray(idx)
        return _wrap(base[idx])

    def __iter__(self):
        base = np.asarray(self)
        for i in range(base.shape[0]):
            yield _wrap(base[i])

    @property
    def at(self):
        return _At(self)

    def astype(self, dt):
        if dt is int:
            dt = np.int32
        elif dt is float or dt == 'float32':
            dt = np.float32
        return _wrap(np.asarray(self).astype(dt))

    def reshape(self, *shape):
        return _wrap(np.asarray(self).reshape(*shape))

    def flatten(self):
        return _wrap(np.asarray(self).flatten())

    def transpose(self, *a):
        return _wrap(np.asarray(self).transpose(*a))

    def any(self, *a, **k):
        return _wrap(np.asarray(self).any(*a, **k))

    def __bool__(self):
        return bool(np.asarray(self))

    def __float__(self):
        return float(np.asarray(self))

    def __int__(self):
        return int(np.asarray(self))


def _fn(npf):
    def f(*args, **kwargs):
        a = _promote([_coerce(x) for x in args])
        return _wrap(npf(*a, **kwargs))
    return f


def _array(x, dtype=None):
    if isinstance(x, (list, tuple)):
        x = [np.asarray(_coerce(v)) for v in x]
        anyf = any(_is_float(np.asarray(v).dtype) for v in x)
        x = np.asarray(x, dtype=np.float32 if anyf else None)
    r = np.asarray(_coerce(x) if not isinstance(x, np.ndarray) else x)
    if dtype is not None:
        r = r.astype(dtype)
    return _wrap(r)


def _concatenate(seq, axis=0):
    arrs = [np.atleast_1d(np.asarray(_coerce(a))) for a in seq]
    anyf = any(_is_float(a.dtype) for a in arrs)
    if anyf:
        arrs = [a.astype(np.float32) for a in arrs]
    return _wrap(np.concatenate(arrs, axis=axis))


def _matmul(a, b):
    a, b = np.asarray(_coerce(a), np.float32), np.asarray(_coerce(b), np.float32)
    # explicit fp32 dot products in index order (what a scalar XLA:CPU loop does)
    if a.ndim == 1 and b.ndim == 2:
        out = np.zeros(b.shape[1], np.float32)
        for j in range(b.shape[1]):
            acc = np.float32(0)
            for k in range(a.shape[0]):
                acc = np.float32(acc + np.float32(a[k] * b[k, j]))
            out[j] = acc
        return _wrap(out)
    if a.ndim == 2 and b.ndim == 1:
        out = np.zeros(a.shape[0], np.float32)
        for i in range(a.shape[0]):
            acc = np.float32(0)
            for k in range(a.shape[1]):
                acc = np.float32(acc + np.float32(a[i, k] * b[k]))
            out[i] = acc
        return _wrap(out)
    return _wrap(a @ b)


def _norm(x, axis=None):
    x = np.asarray(_coerce(x), np.float32)
    return _wrap(np.sqrt(np.sum(np.square(x), axis=axis, dtype=np.float32), dtype=np.float32))


def _where(c, a, b):
    c = np.asarray(_coerce(c))
    a, b = _promote([_coerce(a), _coerce(b)])
    return _wrap(np.where(c, a, b))


class _Key:
    """Stand-in for a threefry key: a path of split indices.  Draw values come
    from a hash of the path, and every uniform draw is RECORDED so the oracle
    can be fed exactly the draws the reference code consumed."""

    def __init__(self, path=()):
        self.path = tuple(path)


DRAW_LOG: list = []


def _split(key, n=2):
    return [_Key(key.path + (i,)) for i in range(n)]


def _uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    seed = abs(hash(key.path)) % (2 ** 63)
    u = np.float32(np.random.default_rng(seed).integers(0, 1 << 24) * 2.0 ** -24)
    lo, hi = np.float32(minval), np.float32(maxval)
    v = np.float32(np.float32(u * np.float32(hi - lo)) + lo)  # jax.random.uniform: u*(max-min)+min
    v = max(lo, v)
    DRAW_LOG.append(float(v))
    return _wrap(np.float32(v))


def install():
    """Put stub modules for jax / gym / mujoco / xmltodict into sys.modules."""
    jnp = types.ModuleType('jax.numpy')
    jnp.pi = np.pi
    jnp.inf = np.inf
    jnp.ndarray = JArr
    jnp.array = _array
    jnp.asarray = _array
    jnp.zeros = lambda shape, dtype=np.float32: _wrap(np.zeros(shape, dtype))
    jnp.ones = lambda shape, dtype=np.float32: _wrap(np.ones(shape, dtype))
    jnp.concatenate = _concatenate
    jnp.matmul = _matmul
    jnp.where = _where
    for name in ('sqrt', 'square', 'exp', 'arctan2', 'maximum', 'minimum', 'isnan', 'isinf', 'abs'):
        setattr(jnp, name, _fn(getattr(np, name)))
    jnp.sum = lambda x, axis=None: _wrap(np.sum(np.asarray(_coerce(x)), axis=axis, dtype=np.asarray(_coerce(x)).dtype))
    linalg = types.ModuleType('jax.numpy.linalg')
    linalg.norm = _norm
    jnp.linalg = linalg

    jax = types.ModuleType('jax')
    jax.numpy = jnp
    jax.Array = JArr
    jax.jit = lambda f, *a, **k: f
    jax.vmap = lambda f, *a, **k: f
    jax.tree_map = lambda f, t: f(t)
    rnd = types.ModuleType('jax.random')
    rnd.PRNGKey = lambda seed: _Key((int(seed),))
    rnd.split = _split
    rnd.uniform = _uniform
    jax.random = rnd
    jax.lax = types.SimpleNamespace()
    dl = types.ModuleType('jax.dlpack')
    jax.dlpack = dl

    gym = types.ModuleType('gym')

    class Env:
        pass

    class EzPickle:
        def __init__(self, *a, **k):
            pass

    class BoxStub:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.shape = tuple(shape) if shape is not None else np.asarray(low).shape
            self.low, self.high, self.dtype = low, high, dtype

    gym.Env = Env
    gym.utils = types.ModuleType('gym.utils')
    gym.utils.EzPickle = EzPickle
    gym.spaces = types.ModuleType('gym.spaces')
    gym.spaces.Box = BoxStub
    gym.spaces.Dict = dict
    gym.vector = types.ModuleType('gym.vector')
    gym.vector.utils = types.ModuleType('gym.vector.utils')

    mujoco = types.ModuleType('mujoco')
    mujoco.viewer = types.ModuleType('mujoco.viewer')
    mjx = types.ModuleType('mujoco.mjx')
    mjx.Data = object
    mjx.Model = object
    mujoco.mjx = mjx

    mods = {
        'jax': jax, 'jax.numpy': jnp, 'jax.numpy.linalg': linalg, 'jax.random': rnd, 'jax.dlpack': dl,
        'gym': gym, 'gym.utils': gym.utils, 'gym.spaces': gym.spaces, 'gym.vector': gym.vector,
        'gym.vector.utils': gym.vector.utils,
        'mujoco': mujoco, 'mujoco.viewer': mujoco.viewer, 'mujoco.mjx': mjx,
        'xmltodict': types.ModuleType('xmltodict'),
    }
    dev = types.ModuleType('safe_rl_envs.envs.mjx_device')
    dev.device_put = lambda *a, **k: None
    mods['safe_rl_envs.envs.mjx_device'] = dev
    sys.modules.update(mods)
    return jnp
