"""CPU-side checks of the drop-in boundary: libguardx.so loads, exports every
symbol include/guardx.h declares, agrees with the ctypes mirror on the config
layout and defaults, and fails LOUDLY (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import pytest

from guardx_b200 import _abi, config as gxconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "guardx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    lib = _abi.load_library()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in guardx.h but not exported"
        assert name in _abi.PROTOTYPES, f"{name} has no ctypes prototype"
    assert sorted(_abi.PROTOTYPES) == declared


def test_task_extension_symbols_exported_and_bound():
    """include/guardx_task.h (config-3 extension): every declared symbol is exported and has a ctypes prototype; the
    config struct has the same layout on both sides; the default config is sane."""
    from guardx_b200 import task_engine
    lib = task_engine.load_library()
    src = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "guardx_task.h")).read(), flags=re.S)
    declared = sorted(set(re.findall(r"\b(gxt_[a-z0-9_]+)\s*\(", src)))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in guardx_task.h but not exported"
    assert sorted(task_engine.PROTOTYPES) == declared
    assert lib.gxt_config_size() == C.sizeof(task_engine.GxtConfig)
    c = task_engine.lower(task_engine.parse({'env_num': 16384}))
    assert list(c.count) == [0, 0, 4, 4, 4, 2, 4, 1] and c.task == 1 and c.robot == 1 and c.lidar_num_bins == 16
    with pytest.raises(AssertionError, match="Bad key"):
        task_engine.parse({'ghosts_num': 8})
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        assert lib.gxt_create(C.byref(c), 0, C.byref(h)) == _abi.GX_ECUDA and not h.value     # no CPU fallback


def test_config_layout_and_defaults_agree():
    lib = _abi.load_library()
    assert lib.gx_config_size() == C.sizeof(_abi.GxConfig)
    c = _abi.GxConfig()
    lib.gx_default_config(C.byref(c))
    py = gxconfig.lower(gxconfig.parse(gxconfig.configuration("Goal_Point_8Hazards")))
    assert bytes(c) == bytes(py)
    assert c.hazards_num == 8 and c.lidar_num_bins == 16 and c.num_steps == 1000
    assert abs(c.point.mass - 5.18879e-3) < 1e-8 and abs(c.point.inertia_zz - 2.842183e-5) < 1e-10


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _abi.load_library()
    c = _abi.GxConfig()
    lib.gx_default_config(C.byref(c))
    h = C.c_void_p()
    rc = lib.gx_create(C.byref(c), 0, C.byref(h))
    assert rc == _abi.GX_ECUDA and not h.value
    assert b"cuda" in lib.gx_last_error().lower()
    from guardx_b200.engine import Engine
    with pytest.raises(_abi.GuardXError):
        Engine({'env_num': 4})


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under guardx_b200/ may import,
    include, link or dlopen it."""
    bad = re.compile(r"^\s*(from|import)\s+oracle\b|#\s*include\s*[\"<][^\">]*oracle|gxoracle|gxo_", re.M)
    for sub in ("guardx_b200", "tools", "include"):     # tools/ are developer scripts of the product side too
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")) or f == "Makefile":
                    text = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(text), f"{sub}/{f} references the oracle"


def test_bad_config_is_rejected_like_the_reference():
    with pytest.raises(AssertionError, match="Bad key"):
        gxconfig.parse({'ghosts_num': 8})
    with pytest.raises(NotImplementedError):
        gxconfig.lower(gxconfig.parse({'robot_base': 'xmls/barkour_v0/assets/barkour.xml'}))   # not a robot of the JAX Engine's configs
    with pytest.raises(AssertionError):
        gxconfig.lower(gxconfig.parse({'hazards_num': 0}))
