"""Factory functions of the reference (safe_rl_libX/guard_utils/safe_rl_env_config.py:2597-2614)."""
from __future__ import annotations

from types import SimpleNamespace

from .config import configuration


def create_env(args, **engine_kwargs):
    """Same contract as the reference's ``create_env(args)``: ``args`` carries
    ``task``, ``env_num``, ``seed`` and ``max_ep_len`` (an argparse namespace)."""
    from .engine import Engine
    config = configuration(args.task)
    config['env_num'] = args.env_num
    config['_seed'] = args.seed
    config['num_steps'] = args.max_ep_len
    config['device_id'] = 0
    return Engine(config, **engine_kwargs)


def make(task: str = "Goal_Point_8Hazards", env_num: int = 1, seed: int = 0, max_ep_len: int = 1000,
         **engine_kwargs):
    """``make`` in BASELINE.json's sense: a thin alias of ``create_env`` (the
    reference registers nothing with gym)."""
    return create_env(SimpleNamespace(task=task, env_num=env_num, seed=seed, max_ep_len=max_ep_len),
                      **engine_kwargs)
