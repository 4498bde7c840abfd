"""Pin the CPU oracle against the reference's OWN source.

tests/golden/refsrc_*.npz were produced by tools/make_golden.py, which imports the
unmodified reference engine.py with a numpy stand-in for jax.numpy and records
what Engine.obs / obs_lidar / obs_compass / reward_done / cost / sample_layout
return.  Floats: |diff| <= 2e-6 (numpy's exp/arctan2 vs gx_math.h, <= 2.5 ulp of
values <= ~7); integers (lidar bins, done flags, layout acceptance) exact.
The MJX physics is not covered by these vectors (parity unpinned there).
"""
import ast
import ctypes as C
import json
import os

import numpy as np
import pytest

from guardx_b200 import _abi, config as gxconfig

TOL = 2e-6


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cfg(gold):
    return gxconfig.lower(gxconfig.parse(ast.literal_eval(str(gold['config']))))


def _xmat9(cs):
    c, s = cs
    return np.array([c, -s, 0, s, c, 0, 0, 0, 1], np.float32)


@pytest.mark.parametrize("tag", ["default", "variant"])
def test_obs_reward_cost_against_reference_source(oracle_lib, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"refsrc_obs_{tag}.npz"))
    cfg = _cfg(g)
    nh, nb = cfg.hazards_num, cfg.lidar_num_bins
    keys = [str(k) for k in g['obs_keys']]
    assert keys == ['ctrl', 'goal_compass', 'goal_lidar', 'hazards_lidar', 'qpos', 'qvel']
    off_gl, off_hl = 5, 5 + nb
    n = g['flat'].shape[0]
    assert g['flat'].shape[1] == 3 + 2 + 2 * nb + 6
    n_bin_edge = 0
    for i in range(n):
        xpos = np.array([g['robot'][i, 0], g['robot'][i, 1], 0.1], np.float32)
        xmat = _xmat9(g['cs'][i])
        flat = g['flat'][i]
        # obs_lidar (engine.py:846-900), goal then hazards
        lid = np.zeros(nb, np.float32)
        bins = np.zeros(nh, np.int32)
        goal = np.ascontiguousarray(g['goal'][i], np.float32)
        assert oracle_lib.gxo_piece_lidar(C.byref(cfg), _p(xpos), _p(xmat), _p(goal), 1, _p(lid), _p(bins)) == 0
        np.testing.assert_allclose(lid, flat[off_gl:off_gl + nb], rtol=0, atol=TOL)
        n_bin_edge += int(bins[0] == nb)
        # the golden goal lidar has one object: the reference's bin holds the maximum
        # (except at the wrap edge, where the reference drops the main scatter)
        if bins[0] < nb:
            assert flat[off_gl + int(bins[0])] == flat[off_gl:off_gl + nb].max()
        haz = np.ascontiguousarray(g['hazards'][i], np.float32)
        lid = np.zeros(nb, np.float32)
        assert oracle_lib.gxo_piece_lidar(C.byref(cfg), _p(xpos), _p(xmat), _p(haz), nh, _p(lid), _p(bins)) == 0
        np.testing.assert_allclose(lid, flat[off_hl:off_hl + nb], rtol=0, atol=TOL)
        # obs_compass (engine.py:834-844)
        comp = np.zeros(2, np.float32)
        oracle_lib.gxo_piece_compass(_p(xpos), _p(xmat), _p(goal), _p(comp))
        np.testing.assert_allclose(comp, flat[3:5], rtol=0, atol=TOL)
        np.testing.assert_allclose(comp, g['compass'][i], rtol=0, atol=TOL)
        # pass-through parts of Engine.obs (engine.py:759-766): ctrl | ... | qpos | qvel
        np.testing.assert_array_equal(flat[:3], g['ctrl'][i])
        np.testing.assert_array_equal(flat[-6:-3], g['qpos'][i])
        np.testing.assert_array_equal(flat[-3:], g['qvel'][i])
        # cost (engine.py:804-811)
        cost = C.c_float()
        rxy = np.ascontiguousarray(g['robot'][i], np.float32)
        oracle_lib.gxo_piece_cost(C.byref(cfg), _p(rxy), _p(haz), C.byref(cost))
        assert abs(cost.value - g['cost'][i]) <= TOL
        assert (cost.value > 0) == (g['cost'][i] > 0)
        # reward_done (engine.py:787-802), with and without a previous frame
        lr = np.ascontiguousarray(g['last_robot'][i], np.float32)
        lg = np.ascontiguousarray(g['last_goal'][i], np.float32)
        r, d = C.c_float(), C.c_float()
        oracle_lib.gxo_piece_reward_done(C.byref(cfg), _p(rxy), _p(goal), _p(lr), _p(lg),
                                         C.c_float(float(g['last_done'][i])), C.byref(r), C.byref(d))
        assert d.value == g['done'][i]
        assert abs(r.value - g['reward'][i]) <= TOL
        oracle_lib.gxo_piece_reward_done(C.byref(cfg), _p(rxy), _p(goal), _p(rxy), _p(goal),
                                         C.c_float(1.0), C.byref(r), C.byref(d))
        assert d.value == g['done_nolast'][i] and r.value == g['reward_nolast'][i] == 0.0
    assert n_bin_edge >= 1, "the fp32 wrap edge (bin == nbins) must be exercised"


@pytest.mark.parametrize("tag", ["default", "variant"])
def test_sample_layout_against_reference_source(oracle_lib, golden_dir, tag):
    g = np.load(os.path.join(golden_dir, f"refsrc_layout_{tag}.npz"))
    cfg = _cfg(g)
    nh = cfg.hazards_num
    names = [str(s) for s in g['names']]
    assert names == ['goal'] + [f'hazard{i}' for i in range(nh)] + ['robot']  # engine.py:538-542
    n_ok = 0
    for i in range(g['draws'].shape[0]):
        draws = np.ascontiguousarray(g['draws'][i], np.float32)
        out = np.zeros(2 * (2 + nh), np.float32)
        ok = C.c_int()
        assert oracle_lib.gxo_piece_layout_from_draws(C.byref(cfg), _p(draws), _p(out), C.byref(ok)) == 0
        assert ok.value == int(g['success'][i])
        ref = g['layouts'][i]  # placement order: goal, hazards..., robot
        want = np.concatenate([ref[0], ref[-1], ref[1:-1].reshape(-1)])
        placed = np.isfinite(want)
        # a conflicted object keeps the reference's -inf placeholder; ours is -1e30
        np.testing.assert_array_equal(out[placed], want[placed])
        assert (out[~placed] == np.float32(-1e30)).all()
        n_ok += ok.value
    assert n_ok >= 5


def test_ego_vel_acc_against_reference_source(oracle_lib, golden_dir):
    """engine.py:902-929 incl. the None-ness of last_done / last_last_done on the first
    two steps, and the sorted-key position of acc (first) and vel (last) in the flat obs."""
    g = np.load(os.path.join(golden_dir, "refsrc_velacc.npz"))
    assert [str(k) for k in g['obs_keys']] == ['acc', 'ctrl', 'goal_compass', 'goal_lidar', 'hazards_lidar',
                                                'qpos', 'qvel', 'vel']
    cfg = gxconfig.lower(gxconfig.parse({'observe_vel': True, 'observe_acc': True}))
    n = g['flat'].shape[0]
    assert g['flat'].shape[1] == 47
    seen = set()
    for i in range(n):
        vel, acc = np.zeros(2, np.float32), np.zeros(2, np.float32)
        f = lambda k: np.ascontiguousarray(g[k][i], np.float32)
        have = int(g['have'][i])
        ll = f('last_last') if have == 2 else f('last')
        assert oracle_lib.gxo_piece_vel_acc(C.byref(cfg), _p(_xmat9(g['cs'][i])), _p(f('cur')), _p(f('last')), _p(ll),
                                            have, C.c_float(float(g['last_done'][i])),
                                            C.c_float(float(g['last_last_done'][i])), _p(vel), _p(acc)) == 0
        # finite differences divide by dt twice: |acc| ~ 1e2, so the tolerance is relative
        np.testing.assert_allclose(vel, g['vel'][i], rtol=2e-5, atol=2e-5)
        np.testing.assert_allclose(acc, g['acc'][i], rtol=2e-5, atol=2e-3)
        np.testing.assert_array_equal(g['flat'][i][:2], g['acc'][i])
        np.testing.assert_array_equal(g['flat'][i][-2:], g['vel'][i])
        if have == 0:
            assert (vel == 0).all() and (acc == 0).all()
        seen.add(have)
    assert seen == {0, 1, 2}


def test_default_keys_match_reference(golden_dir):
    with open(os.path.join(golden_dir, "engine_default.json")) as f:
        ref = json.load(f)
    assert set(ref) == set(gxconfig.DEFAULT)
    for k, v in ref.items():
        assert gxconfig.DEFAULT[k] == v, k
    with pytest.raises(AssertionError, match="Bad key"):
        gxconfig.parse({'observe_ghosts': True})
