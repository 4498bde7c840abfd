"""Accuracy of the shared deterministic fp32 math (guardx_b200/csrc/gx_math.h),
measured against float64 numpy.  Bit-equality with the device build of the same
header is checked on the GPU in test_parity_gpu.py."""
import ctypes as C

import numpy as np


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ulp(got, ref64):
    ref32 = ref64.astype(np.float32)
    return np.abs(got.astype(np.float64) - ref64) / np.spacing(np.abs(ref32)).astype(np.float64)


def test_sincos_accuracy(oracle_lib):
    rng = np.random.default_rng(0)
    n = 1_000_000
    for scale in (1.0, 100.0, 1e4):
        x = rng.uniform(-scale, scale, n).astype(np.float32)
        s = np.empty_like(x)
        c = np.empty_like(x)
        oracle_lib.gxo_v_sincos(_p(x), _p(s), _p(c), n)
        x64 = x.astype(np.float64)
        assert np.abs(s - np.sin(x64)).max() < 1.5e-7
        assert np.abs(c - np.cos(x64)).max() < 1.5e-7
        assert _ulp(s, np.sin(x64)).max() < 2.0 and _ulp(c, np.cos(x64)).max() < 2.0


def test_sincos_exact_points(oracle_lib):
    x = np.array([0.0, -0.0], np.float32)
    s = np.empty_like(x)
    c = np.empty_like(x)
    oracle_lib.gxo_v_sincos(_p(x), _p(s), _p(c), 2)
    assert s[0] == 0.0 and c[0] == 1.0 and c[1] == 1.0  # theta = 0 must give R = I exactly
    bad = np.array([np.inf, -np.inf, np.nan], np.float32)
    s = np.empty_like(bad)
    c = np.empty_like(bad)
    oracle_lib.gxo_v_sincos(_p(bad), _p(s), _p(c), 3)
    assert np.isnan(s).all() and np.isnan(c).all()


def test_atan2_accuracy_and_quadrants(oracle_lib):
    rng = np.random.default_rng(1)
    n = 1_000_000
    y = (rng.normal(size=n) * rng.choice([1e-3, 1, 10, 1e3], n)).astype(np.float32)
    x = (rng.normal(size=n) * rng.choice([1e-3, 1, 10], n)).astype(np.float32)
    o = np.empty_like(x)
    oracle_lib.gxo_v_atan2(_p(y), _p(x), _p(o), n)
    assert _ulp(o, np.arctan2(y.astype(np.float64), x.astype(np.float64))).max() < 3.0
    yy = np.array([0, 0, -0.0, 1, -1, -1e-8, 0.0], np.float32)
    xx = np.array([0, -1, -1, 0, 0, 1, 1], np.float32)
    o = np.empty_like(xx)
    oracle_lib.gxo_v_atan2(_p(yy), _p(xx), _p(o), len(xx))
    np.testing.assert_array_equal(o, np.arctan2(yy, xx).astype(np.float32))
    nn = np.array([np.nan, 1.0], np.float32)
    o = np.empty(2, np.float32)
    oracle_lib.gxo_v_atan2(_p(nn), _p(nn[::-1].copy()), _p(o), 2)
    assert np.isnan(o).all()


def test_exp_accuracy(oracle_lib):
    rng = np.random.default_rng(2)
    n = 1_000_000
    x = rng.uniform(-87, 5, n).astype(np.float32)
    o = np.empty_like(x)
    oracle_lib.gxo_v_exp(_p(x), _p(o), n)
    assert _ulp(o, np.exp(x.astype(np.float64))).max() < 2.0
    sp = np.array([0.0, -0.0, -200.0, 100.0, np.nan, -np.inf], np.float32)
    o = np.empty_like(sp)
    oracle_lib.gxo_v_exp(_p(sp), _p(o), len(sp))
    assert o[0] == 1.0 and o[1] == 1.0 and o[2] == 0.0 and np.isinf(o[3]) and np.isnan(o[4]) and o[5] == 0.0
