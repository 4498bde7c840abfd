"""Rollout-side device ops: what the reference does on the host
between env steps and the policy update, as CUDA kernels behind the C ABI."""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi


def gae(reward: torch.Tensor, value: torch.Tensor, done: torch.Tensor, last_value: torch.Tensor | None = None,
        gamma: float = 0.99, lam: float = 0.97):
    """GAE-lambda advantages and rewards-to-go of a time-major rollout.

    Replaces ``PPOBufferX.finish_path`` (safe_rl_libX/ppo/ppo_train.py:61-113).
    ``reward``, ``value``, ``done``: CUDA fp32 ``[T, N]``; ``done[t, i] = 1`` closes env i's
    path after step t (bootstrapped with 0, as the reference does for finished and
    timed-out paths alike); the open tail is bootstrapped with ``last_value`` (``[N]``) or 0.
    Returns ``(adv, ret)``, both ``[T, N]``.
    """
    assert reward.is_cuda and reward.dtype == torch.float32 and reward.dim() == 2
    T, n = reward.shape
    reward, value, done = reward.contiguous(), value.contiguous(), done.contiguous()
    assert value.shape == (T, n) and done.shape == (T, n)
    if last_value is not None:
        last_value = last_value.to(torch.float32).contiguous()
        assert last_value.shape == (n,)
    adv, ret = torch.empty_like(reward), torch.empty_like(reward)
    lib = _abi.load_library()
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(reward.device):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _abi.check(lib, lib.gx_gae(p(reward), p(value), p(done), p(last_value), T, n, float(gamma), float(lam),
                                   p(adv), p(ret), stream), "gx_gae")
    return adv, ret


def done_indices(done: torch.Tensor):
    """Ascending indices of the envs whose ``done`` flag is set, without a host round trip.

    Replaces ``np.where(done == 1)`` on a device->host copy (ppo_train.py:403-411,
    scpo/scpo.py:722: ``env.reset_done()[np.where(done == 1)]``).  Returns ``(idx, count)``:
    ``idx`` int32 ``[N]`` whose first ``count`` entries are valid, ``count`` int32 ``[1]`` -- both
    on the device; ``idx[:int(count)]`` is the only point that needs the host.
    """
    assert done.is_cuda and done.dtype == torch.float32 and done.dim() == 1
    done = done.contiguous()
    n = done.numel()
    idx = torch.empty(n, dtype=torch.int32, device=done.device)
    count = torch.empty(1, dtype=torch.int32, device=done.device)
    lib = _abi.load_library()
    with torch.cuda.device(done.device):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        _abi.check(lib, lib.gx_compact_done(C.c_void_p(done.data_ptr()), n, C.c_void_p(idx.data_ptr()),
                                            C.c_void_p(count.data_ptr()), stream), "gx_compact_done")
    return idx, count
