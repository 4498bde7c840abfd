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
