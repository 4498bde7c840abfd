"""guardx_b200 -- B200-native batched ``safe_rl_envs`` environment step.

Drop-in for the reference's env API:
``create_env(args)`` / ``Engine(config)`` -> ``reset() / step(a) / reset_done()``.
"""
from .config import DEFAULT, configuration, parse  # noqa: F401


def __getattr__(name):  # Engine pulls in torch + the CUDA library: import lazily
    if name in ("Engine", "Box", "ResamplingError"):
        from . import engine
        return getattr(engine, name)
    if name in ("create_env", "make"):
        from . import env_config
        return getattr(env_config, name)
    raise AttributeError(name)
