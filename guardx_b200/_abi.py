"""ctypes mirror of ``include/guardx.h`` and the loader of ``libguardx.so``.

There is no CPU fallback: if the CUDA library is missing or fails to load,
importing the env raises.  The oracle (``oracle/``) is test infrastructure and is
never imported from this package.
"""
from __future__ import annotations

import ctypes as C
import os

GX_ABI_VERSION = 3
GX_MAX_HAZARDS = 32
GX_MAX_LIDAR_BINS = 64

GX_OK, GX_EINVAL, GX_ECUDA, GX_ENOMEM, GX_ESTATE = 0, -1, -2, -3, -4
GX_ROBOT_POINT, GX_ROBOT_ART, GX_ROBOT_TREE = 0, 1, 2

GX_OBS_ACC = 1 << 0
GX_OBS_CTRL = 1 << 1
GX_OBS_GOAL_COMPASS = 1 << 2
GX_OBS_GOAL_LIDAR = 1 << 3
GX_OBS_HAZARDS_LIDAR = 1 << 4
GX_OBS_QPOS = 1 << 5
GX_OBS_QVEL = 1 << 6
GX_OBS_VEL = 1 << 7

GX_AUX_GUARD, GX_AUX_TIMEOUT, GX_AUX_GOAL, GX_AUX_JUMP = 1, 2, 4, 8

# flat observation order = sorted key order (engine.py:773-777)
OBS_COMPONENTS = (
    ("acc", GX_OBS_ACC),
    ("ctrl", GX_OBS_CTRL),
    ("goal_compass", GX_OBS_GOAL_COMPASS),
    ("goal_lidar", GX_OBS_GOAL_LIDAR),
    ("hazards_lidar", GX_OBS_HAZARDS_LIDAR),
    ("qpos", GX_OBS_QPOS),
    ("qvel", GX_OBS_QVEL),
    ("vel", GX_OBS_VEL),
)


class GxPointModel(C.Structure):
    _fields_ = [
        ("timestep", C.c_float),
        ("body_z", C.c_float),
        ("mass", C.c_float),
        ("com_x", C.c_float),
        ("inertia_zz", C.c_float),
        ("damping", C.c_float * 3),
        ("gear", C.c_float * 3),
        ("gainprm0", C.c_float),
        ("biasprm2", C.c_float),
        ("ctrllimited", C.c_int32),
        ("ctrlrange", C.c_float * 2),
        ("forcelimited", C.c_int32),
        ("forcerange", C.c_float * 2),
    ]


class GxConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("env_num", C.c_int32),
        ("num_steps", C.c_int32),
        ("robot", C.c_int32),
        ("seed", C.c_uint64),
        ("env_id_offset", C.c_uint64),
        ("placements_extents", C.c_double * 4),
        ("placements_margin", C.c_double),
        ("robot_keepout", C.c_double),
        ("goal_keepout", C.c_double),
        ("hazards_keepout", C.c_double),
        ("goal_size", C.c_double),
        ("hazards_size", C.c_double),
        ("reward_distance", C.c_double),
        ("lidar_max_dist", C.c_double),
        ("lidar_exp_gain", C.c_double),
        ("goal_min_robot_dist", C.c_double),
        ("opt_timestep", C.c_double),
        ("hazards_num", C.c_int32),
        ("lidar_num_bins", C.c_int32),
        ("lidar_alias", C.c_int32),
        ("physics_steps_per_control_step", C.c_int32),
        ("layout_tries", C.c_int32),
        ("pool_size", C.c_int32),
        ("obs_mask", C.c_uint32),
        ("art_nu", C.c_int32),
        ("art_nv", C.c_int32),
        ("art_nq", C.c_int32),
        ("point", GxPointModel),
    ]


_P = C.c_void_p
_I = C.c_int
_CFGP = C.POINTER(GxConfig)

# name -> (restype, argtypes); every symbol include/guardx.h declares
PROTOTYPES = {
    "gx_abi_version": (_I, []),
    "gx_config_size": (_I, []),
    "gx_last_error": (C.c_char_p, []),
    "gx_default_config": (None, [_CFGP]),
    "gx_create": (_I, [_CFGP, _I, C.POINTER(_P)]),
    "gx_create_art": (_I, [_CFGP, _P, _I, C.POINTER(_P)]),
    "gx_create_tree": (_I, [_CFGP, _P, _I, C.POINTER(_P)]),
    "gx_destroy": (_I, [_P]),
    "gx_obs_dim": (_I, [_P]),
    "gx_act_dim": (_I, [_P]),
    "gx_env_num": (_I, [_P]),
    "gx_pool_size": (_I, [_P]),
    "gx_obs_offset": (_I, [_P, C.c_uint32]),
    "gx_action_bounds": (_I, [_P, _P, _P]),
    "gx_reset": (_I, [_P, _P, _P]),
    "gx_reset_with_layout": (_I, [_P, _P, _P, _P]),
    "gx_step": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "gx_step_autoreset": (_I, [_P, _P, _P, _P, _P, _P, _P]),
    "gx_aux_dim": (_I, [_P]),
    "gx_reset_done": (_I, [_P, _P, _P, _P, _P]),
    "gx_state_dim": (_I, [_P]),
    "gx_get_state": (_I, [_P, _P, _P]),
    "gx_set_state": (_I, [_P, _P, _P]),
    "gx_set_counters": (_I, [_P, C.c_uint64, C.c_uint64, _P]),
    "gx_get_counters": (_I, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), _P]),
    "gx_pool_read": (_I, [_P, _P, _P, _P]),
    "gx_rollout": (_I, [_P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "gx_stats_read": (_I, [_P, _P, _P]),
    "gx_stats_reset": (_I, [_P, _P]),
    "gx_step_host": (_I, [_P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "gx_rollout_host": (_I, [_P, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "gx_compact_done": (_I, [_P, C.c_int32, _P, _P, _P]),
    "gx_gae": (_I, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_double, C.c_double, _P, _P, _P]),
    "gx_debug_math": (_I, [_I, _P, _P, _P, _P, _I, _P]),
    "gx_debug_sm_clock": (_I, [_P, _I, _P]),
    "gx_rollout_host_route": (_I, [_P, _P]),
    "gx_policy_weight_floats": (C.c_int64, [_P]),
    "gx_policy_step": (_I, [_P, _P, _P, C.c_uint64, _P, _P, _P, _P, _P, _P]),
    "gx_policy_rollout_supported": (_I, [_P]),
    "gx_policy_rollout": (_I, [_P, _P, C.c_int32, C.c_uint64, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
}

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GX_LIB") or os.path.join(_HERE, "libguardx.so")  # GX_LIB: kernel-tuning builds
_lib = None


class GuardXError(RuntimeError):
    pass


def load_library(path: str | None = None) -> C.CDLL:
    """dlopen libguardx.so and bind every prototype.  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("GUARDX_LIB") or LIB_PATH   # GUARDX_LIB: developer override (kernel build variants)
    if not os.path.exists(p):
        raise GuardXError(
            f"{p} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C guardx_b200/csrc`).  guardx_b200 has no CPU fallback."
        )
    lib = C.CDLL(p)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.gx_abi_version() != GX_ABI_VERSION:
        raise GuardXError("libguardx.so ABI version mismatch")
    if lib.gx_config_size() != C.sizeof(GxConfig):
        raise GuardXError("gx_config layout mismatch between guardx.h and _abi.py")
    if path is None:
        _lib = lib
    return lib


def check(lib: C.CDLL, rc: int, what: str) -> None:
    if rc == GX_OK:
        return
    msg = lib.gx_last_error()
    msg = msg.decode() if msg else ""
    if rc in (GX_EINVAL, GX_ESTATE):
        raise AssertionError(f"{what}: {msg}")
    raise GuardXError(f"{what}: {msg} (code {rc})")
