"""Parity of the CUDA path (through the C ABI, via guardx_b200.Engine) against
the CPU oracle on identical seeds, layouts and action sequences.

Bar (BASELINE.json north_star): integer outputs -- done flags, reset masks,
lidar bin indices, violation counts, guard/timeout flags -- bit-exact; floats
within a stated fp32 tolerance.  Because kernels and oracle share the same
deterministic fp32 math and the same evaluation order, the tolerance these
tests actually assert for floats is ZERO (bit-equality, NaN == NaN).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def bits_equal(a, b, what=""):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    if a.dtype.kind == "f":
        nan = np.isnan(a) & np.isnan(b)
        same = (a.view(np.uint32) == b.view(np.uint32)) | nan
    else:
        same = a == b
    if not same.all():
        idx = np.argwhere(~same)[:5]
        raise AssertionError(f"{what}: {int((~same).sum())} of {same.size} differ; first at {idx.tolist()} "
                             f"gpu={a[tuple(idx[0])]!r} cpu={b[tuple(idx[0])]!r}")


def cpu(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module", params=["lanes16", "thread_per_env"])
def mods(request):
    """Every test runs twice: with the 16-lanes-per-env kernels (the default up to
    16384 envs) and with the one-thread-per-env kernels forced for all sizes."""
    _need_gpu()
    from guardx_b200.engine import Engine
    from oracle.oracle import OracleEngine
    old = os.environ.get("GX_LANES_MAX_ENVS")
    os.environ["GX_LANES_MAX_ENVS"] = "1048576" if request.param == "lanes16" else "0"
    yield Engine, OracleEngine
    if old is None:
        os.environ.pop("GX_LANES_MAX_ENVS", None)
    else:
        os.environ["GX_LANES_MAX_ENVS"] = old


def make_pair(mods, cfg, threads=8):
    Engine, OracleEngine = mods
    return Engine(dict(cfg)), OracleEngine(dict(cfg), threads=threads)


def layouts_from(oracle, n, seed):
    pool, _ = oracle.pool()
    rng = np.random.default_rng(seed)
    return pool[rng.integers(0, pool.shape[0], n)].copy()


# ------------------------------------------------------------------ math header
def test_shared_math_is_bit_identical_on_device(mods, oracle_lib):
    from guardx_b200 import _abi
    lib = _abi.load_library()
    rng = np.random.default_rng(3)
    n = 1 << 18
    x = np.concatenate([rng.uniform(-400, 400, n - 8), [0.0, -0.0, 1e9, -1e9, np.inf, -np.inf, np.nan, 1e-30]]).astype(np.float32)
    y = rng.normal(size=n).astype(np.float32) * rng.choice([1e-6, 1e-3, 1.0, 50.0], n).astype(np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    dx, dy = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    o1, o2 = torch.empty_like(dx), torch.empty_like(dx)
    dp = lambda t: C.c_void_p(t.data_ptr())
    s = np.empty_like(x)
    c = np.empty_like(x)
    oracle_lib.gxo_v_sincos(p(x), p(s), p(c), n)
    assert lib.gx_debug_math(0, dp(dx), None, dp(o1), dp(o2), n, None) == 0
    bits_equal(cpu(o1), s, "sin")
    bits_equal(cpu(o2), c, "cos")
    o = np.empty_like(x)
    oracle_lib.gxo_v_atan2(p(y), p(x), p(o), n)
    assert lib.gx_debug_math(1, dp(dy), dp(dx), dp(o1), None, n, None) == 0
    bits_equal(cpu(o1), o, "atan2")
    xe = np.clip(x, -120, 90)
    xe[np.isnan(x)] = np.nan
    oracle_lib.gxo_v_exp(p(xe), p(o), n)
    assert lib.gx_debug_math(2, dp(torch.from_numpy(xe).cuda()), None, dp(o1), None, n, None) == 0
    bits_equal(cpu(o1), o, "exp")
    # max / min: NaN propagation and signed zeros
    special = np.array([0.0, -0.0, np.nan, np.inf, -np.inf, 1.0, -1.0, 1e-45, -1e-45], np.float32)
    aa, bb = [v.ravel().copy() for v in np.meshgrid(special, special)]
    da, db, do_ = torch.from_numpy(aa).cuda(), torch.from_numpy(bb).cuda(), torch.empty(len(aa), device='cuda')
    ref = np.empty_like(aa)
    for fn, name in ((3, "gxo_v_max"), (4, "gxo_v_min")):
        getattr(oracle_lib, name)(p(aa), p(bb), p(ref), len(aa))
        assert lib.gx_debug_math(fn, dp(da), dp(db), dp(do_), None, len(aa), None) == 0
        bits_equal(cpu(do_), ref, name)
    oracle_lib.gxo_v_max(p(x), p(y), p(o), n)
    assert lib.gx_debug_math(3, dp(dx), dp(dy), dp(o1), None, n, None) == 0
    bits_equal(cpu(o1), o, "max")


# ---------------------------------------------------------------- layout pool
def test_layout_pool_and_reset_bit_exact(mods):
    g, o = make_pair(mods, {'env_num': 100, '_seed': 7})
    og = g.reset()
    oo = o.reset()
    pool_g, tries_g = g.layout_pool()
    pool_o, tries_o = o.pool()
    bits_equal(cpu(tries_g), tries_o, "pool tries")
    bits_equal(cpu(pool_g), pool_o, "pool layouts")
    assert 0.015 < 1.0 / tries_o.mean() < 0.025  # ~1.96 % acceptance
    bits_equal(cpu(og), oo, "reset obs")
    bits_equal(cpu(g.get_state()), o.get_state(), "reset state")
    # second reset(): new pool (epoch 1), new picks
    og2, oo2 = g.reset(), o.reset()
    bits_equal(cpu(og2), oo2, "second reset obs")
    assert not np.array_equal(cpu(og2), cpu(og))
    # placement constraints of engine.py:546-572 hold for every pool entry
    pool = pool_o.reshape(-1, 10, 2)
    goal, robot, haz = pool[:, 0], pool[:, 1], pool[:, 2:]
    assert (np.linalg.norm(robot - goal, axis=1) >= 3.0).all()
    assert (np.abs(goal) <= 1.5 + 1e-6).all() and (np.abs(robot) <= 1.6 + 1e-6).all() and (np.abs(haz) <= 1.6 + 1e-6).all()
    d_hg = np.linalg.norm(haz - goal[:, None], axis=2)
    assert (d_hg >= 0.9 - 1e-6).all()
    d_hr = np.linalg.norm(haz - robot[:, None], axis=2)
    assert (d_hr >= 0.8 - 1e-6).all()
    for a in range(8):
        for b in range(a + 1, 8):
            assert (np.linalg.norm(haz[:, a] - haz[:, b], axis=1) >= 0.8 - 1e-6).all()


# -------------------------------------------- config 1: 64 envs x 1000 steps
def test_free_running_64x1000_random_actions(mods):
    """BASELINE.json configs[0]: Goal_Point_8Hazards, 64 envs, random actions,
    1000 steps.  Every output of every step is compared, plus the final state."""
    n, T = 64, 1000
    g, o = make_pair(mods, {'env_num': n, 'num_steps': 1000, '_seed': 0})
    lay = layouts_from(OracleSeed(mods, n), n, 11)
    bits_equal(cpu(g.reset(torch.from_numpy(lay).cuda())), o.reset(lay), "reset obs")
    rng = np.random.default_rng(5)
    acts = rng.uniform(-1, 1, (T, n, 2)).astype(np.float32)
    acts_d = torch.from_numpy(acts).cuda()
    viol = 0
    for t in range(T):
        og, rg, dg, ig = g.step(acts_d[t], want_aux=True)
        oo, ro, do, io = o.step(acts[t], want_qacc=True, want_aux=True)
        bits_equal(cpu(ig['aux']), io['aux'], f"aux t={t}")          # flags, violations, lidar bins
        bits_equal(cpu(dg), do, f"done t={t}")
        bits_equal(cpu(og), oo, f"obs t={t}")
        bits_equal(cpu(rg), ro, f"reward t={t}")
        bits_equal(cpu(ig['cost']), io['cost'], f"cost t={t}")
        bits_equal(cpu(ig['obs']['qacc']), io['qacc'], f"qacc t={t}")
        viol += int(io['aux'][:, 1].sum())
    bits_equal(cpu(g.get_state()), o.get_state(), "final state")
    assert viol > 0, "the fixture should visit hazards"


def OracleSeed(mods, n):
    """an oracle with a generated pool, used as a source of valid layouts"""
    _, OracleEngine = mods
    e = OracleEngine({'env_num': n, '_seed': 99}, threads=8)
    e.reset()
    return e


def seek_policy(obs):
    """drive towards the goal from the compass (obs[3:5]); pure numpy so that both
    implementations are fed identical actions"""
    cx, cy = obs[:, 3], obs[:, 4]
    ang = np.arctan2(cy, cx)
    fwd = np.where(np.abs(ang) < 0.6, 1.0, 0.05)
    turn = np.clip(2.0 * ang, -1, 1)
    return np.stack([fwd, turn], 1).astype(np.float32)


@pytest.mark.parametrize("n", [96, 257])
def test_goal_seeking_with_reset_done_and_timeouts(mods, n):
    """Episodes that actually END: goal reached, timeouts (num_steps small),
    reset_done from the pool after every step, stale heading after reset_done."""
    g, o = make_pair(mods, {'env_num': n, 'num_steps': 110, '_seed': 3})
    og, oo = g.reset(), o.reset()
    bits_equal(cpu(og), oo, "reset")
    n_done = n_goal = n_timeout = 0
    for t in range(700):
        a = seek_policy(oo)
        og, rg, dg, ig = g.step(torch.from_numpy(a).cuda(), want_aux=True)
        oo, ro, do, io = o.step(a, want_aux=True)
        bits_equal(cpu(ig['aux']), io['aux'], f"aux t={t}")
        bits_equal(cpu(dg), do, f"done t={t}")
        bits_equal(cpu(og), oo, f"obs t={t}")
        bits_equal(cpu(rg), ro, f"reward t={t}")
        bits_equal(cpu(ig['cost']), io['cost'], f"cost t={t}")
        n_done += int(do.sum())
        n_goal += int(((io['aux'][:, 0] & 4) != 0).sum())
        n_timeout += int(((io['aux'][:, 0] & 2) != 0).sum())
        if do.any():  # the PPO loop's protocol (ppo_train.py:387-417)
            og2, oo2 = g.reset_done(), o.reset_done()
            bits_equal(cpu(og2), oo2, f"reset_done obs t={t}")
            mask = do > 0
            assert np.array_equal(oo2[~mask], oo[~mask])            # other rows untouched
            assert (oo2[mask][:, :3] == 0).all() and (oo2[mask][:, -3:] == 0).all()  # ctrl, qvel = 0
            oo = oo2
        if t % 100 == 0:
            bits_equal(cpu(g.get_state()), o.get_state(), f"state t={t}")
    bits_equal(cpu(g.get_state()), o.get_state(), "final state")
    assert n_goal > 10 and n_timeout > 10 and n_done >= n_goal


def test_reward_is_zero_after_reset_and_lags_one_step(mods):
    """reward_t = dist(q_{t-1}) - dist(q_t), zero
    on the first step after reset() and after reset_done()."""
    g, o = make_pair(mods, {'env_num': 32, 'num_steps': 1000, '_seed': 1})
    g.reset(); o.reset()
    a = np.tile(np.array([[1.0, 0.3]], np.float32), (32, 1))
    _, r1, _, _ = g.step(torch.from_numpy(a).cuda())
    assert (cpu(r1) == 0).all()
    s_before = cpu(g.get_state())
    _, r2, _, _ = g.step(torch.from_numpy(a).cuda())
    # the second step pays for the motion of the first one
    assert (np.abs(cpu(r2)) > 0).any()
    o.step(a); _, r2o, _, _ = o.step(a)
    bits_equal(cpu(r2), r2o, "lagged reward")
    assert s_before.shape == (32, 19 + 16)


# ----------------------------------------------------------- fused rollout
@pytest.mark.parametrize("n,T,num_steps", [(333, 300, 40), (4096, 120, 25)])
def test_rollout_kernel_matches_oracle_rollout(mods, n, T, num_steps):
    g, o = make_pair(mods, {'env_num': n, 'num_steps': num_steps, '_seed': 5})
    bits_equal(cpu(g.reset()), o.reset(), "reset")
    rg = g.rollout(T)
    ro = o.rollout(T)
    for k in ("act", "done", "reward", "cost", "obs"):
        bits_equal(cpu(rg[k]), ro[k], f"rollout {k}")
    bits_equal(cpu(g.get_state()), o.get_state(), "state after rollout")
    assert ro["done"].sum() > n  # timeouts fired and were reset in-kernel
    # a second rollout continues the same streams
    rg2, ro2 = g.rollout(7), o.rollout(7)
    bits_equal(cpu(rg2["obs"]), ro2["obs"], "second rollout obs")


def test_rollout_equals_step_plus_reset_done_on_gpu(mods):
    """Self-consistency at BASELINE configs[1] size: the fused multi-step kernel
    and the single-step API produce the same bits."""
    Engine, _ = mods
    cfg = {'env_num': 4096, 'num_steps': 30, '_seed': 9}
    a, b = Engine(dict(cfg)), Engine(dict(cfg))
    oa, ob = a.reset(), b.reset()
    assert torch.equal(oa, ob)
    T = 64
    ra = a.rollout(T)
    for t in range(T):
        o_t, r_t, d_t, info = b.step(ra["act"][t])
        o_t = b.reset_done()
        assert torch.equal(d_t, ra["done"][t]), t
        assert torch.equal(r_t.view(torch.int32), ra["reward"][t].view(torch.int32)), t
        assert torch.equal(info['cost'].view(torch.int32), ra["cost"][t].view(torch.int32)), t
        assert torch.equal(o_t.view(torch.int32), ra["obs"][t].view(torch.int32)), t
    assert torch.equal(a.get_state().view(torch.int32), b.get_state().view(torch.int32))
    assert ra["done"].sum().item() > 4096


# ------------------------------------------------- other configs and edges
VARIANTS = [
    {'hazards_num': 3, 'lidar_num_bins': 10, 'lidar_alias': False, 'lidar_exp_gain': 0.5,
     'goal_size': 0.3, 'hazards_size': 0.2, 'reward_distance': 2.0},
    {'hazards_num': 5, 'hazards_keepout': 0.25, 'goal_keepout': 0.35, 'robot_keepout': 0.3,
     'placements_margin': 0.05, 'placements_extents': [-2.5, -2, 2.5, 2], 'lidar_num_bins': 7},
    {'physics_steps_per_control_step': 3, 'hazards_num': 8},
    {'lidar_max_dist': 3.0, 'hazards_num': 12, 'hazards_keepout': 0.2, 'lidar_num_bins': 32},
    {'observe_ctrl': False, 'observe_qvel': False, 'observe_goal_lidar': False},
    {'observe_hazards': False, 'observe_goal_comp': False, 'hazards_num': 1, 'lidar_num_bins': 1},
    {'hazards_num': 32, 'hazards_keepout': 0.12, 'hazards_size': 0.1, 'lidar_num_bins': 64},
    {'observe_vel': True, 'observe_acc': True},                                  # ego_vel_acc, engine.py:902-929
    {'observe_vel': True, 'observe_qpos': False, 'hazards_num': 4, 'physics_steps_per_control_step': 2},
]


@pytest.mark.parametrize("variant", range(len(VARIANTS)))
def test_non_default_configs(mods, variant):
    cfg = dict(VARIANTS[variant], env_num=77, num_steps=35, _seed=variant)
    g, o = make_pair(mods, cfg)
    assert g.obs_flat_size == o.obs_dim and g.observation_space.shape == (o.obs_dim,)
    og, oo = g.reset(), o.reset()
    bits_equal(cpu(og), oo, "reset")
    bits_equal(cpu(g.layout_pool()[0]), o.pool()[0], "pool")
    rng = np.random.default_rng(variant)
    for t in range(150):
        a = rng.uniform(-1.5, 1.5, (77, 2)).astype(np.float32)
        og, rg, dg, ig = g.step(torch.from_numpy(a).cuda(), want_aux=True)
        oo, ro, do, io = o.step(a, want_aux=True)
        bits_equal(cpu(ig['aux']), io['aux'], f"aux t={t}")
        bits_equal(cpu(og), oo, f"obs t={t}")
        bits_equal(cpu(rg), ro, f"reward t={t}")
        bits_equal(cpu(dg), do, f"done t={t}")
        bits_equal(cpu(ig['cost']), io['cost'], f"cost t={t}")
        og, oo = g.reset_done(), o.reset_done()
        bits_equal(cpu(og), oo, f"reset_done t={t}")
    rg, ro = g.rollout(40), o.rollout(40)
    for k in ("obs", "reward", "cost", "done", "act"):
        bits_equal(cpu(rg[k]), ro[k], f"rollout {k}")
    bits_equal(cpu(g.get_state()), o.get_state(), "state incl. history columns")
    # reset() leaves the history stale, exactly like the reference (engine.py:454-467)
    bits_equal(cpu(g.reset()), o.reset(), "second reset")
    a = rng.uniform(-1, 1, (77, 2)).astype(np.float32)
    bits_equal(cpu(g.step(torch.from_numpy(a).cuda())[0]), o.step(a)[0], "step after second reset")


def test_nan_inf_actions_trip_the_guard(mods):
    """engine.py:696-699: NaN/Inf anywhere in obs => reward 0, done 1."""
    n = 40
    g, o = make_pair(mods, {'env_num': n, '_seed': 2})
    g.reset(); o.reset()
    a = np.random.default_rng(0).uniform(-1, 1, (n, 2)).astype(np.float32)
    a[3, 0] = np.nan; a[5, 1] = np.inf; a[9, 0] = -np.inf; a[11] = [np.nan, np.nan]; a[13, 0] = 1e30
    for t in range(4):
        og, rg, dg, ig = g.step(torch.from_numpy(a).cuda(), want_aux=True)
        oo, ro, do, io = o.step(a, want_aux=True)
        bits_equal(cpu(ig['aux']), io['aux'], "aux")
        bits_equal(cpu(og), oo, "obs"); bits_equal(cpu(rg), ro, "reward"); bits_equal(cpu(dg), do, "done")
        bits_equal(cpu(ig['cost']), io['cost'], "cost")
        for k in ((3, 5, 9, 11) if t == 0 else (3,) if t == 1 else ()):
            assert do[k] == 1.0 and ro[k] == 0.0 and (io['aux'][k, 0] & 1)
        if t >= 2:
            assert not (io['aux'][:, 0] & 1).any()  # re-placed envs recover
        og, oo = g.reset_done(), o.reset_done()
        bits_equal(cpu(og), oo, "reset_done")
        a = np.random.default_rng(t + 1).uniform(-1, 1, (n, 2)).astype(np.float32)
        if t == 0:
            a[3, 0] = np.nan
    bits_equal(cpu(g.get_state()), o.get_state(), "state")


@pytest.mark.parametrize("n,lanes", [(96, True), (20000, False)])
def test_exotic_states_through_the_fused_autoreset_paths(mods, n, lanes):
    """The fused step + auto-reset kernels decide `done` before the lidar only for states whose pose and objects are finite and
    below 1e15 (gx_lanes.cuh / guardx.cu: `sane`); anything else takes the reference order in an out-of-line function.  Inject
    such states (NaN / Inf / astronomically large pose, goal and hazards, NaN velocities, huge step counters) and compare
    step_autoreset with the oracle's step + reset_done, bit for bit, on both kernel families (16 lanes per env; thread per env)."""
    g, o = make_pair(mods, {'env_num': n, '_seed': 6, 'num_steps': 30})
    g.reset(); o.reset()
    rng = np.random.default_rng(3)
    for rep in range(3):
        s = o.get_state()
        k = np.arange(0, n, 7)
        vals = np.array([np.nan, np.inf, -np.inf, 1e20, -3e16, 2e15, 1e14, -1e30], np.float32)
        col = rng.choice([0, 1, 2, 3, 5, 11, 12, 13, 14, 20, 27], len(k))       # qx qy qth vx vw goal_x goal_y hazard coords
        s[k, col] = vals[rng.integers(0, len(vals), len(k))]
        s[k[::3], 10] = 40.0                                                     # and some step counters beyond num_steps
        g.set_state(torch.from_numpy(s).cuda()); o.set_state(s)
        for t in range(3):
            a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
            og, rg, dg, ig = g.step_autoreset(torch.from_numpy(a).cuda())
            oo, ro, do, io = o.step(a)
            oo = o.reset_done()
            bits_equal(cpu(og), oo, f"obs {rep}/{t}"); bits_equal(cpu(rg), ro, "reward"); bits_equal(cpu(dg), do, "done")
            bits_equal(cpu(ig['cost']), io['cost'], "cost")
        bits_equal(cpu(g.get_state()), o.get_state(), "state")
        assert do.sum() >= 0


def test_state_injection_roundtrip_and_one_step_parity(mods):
    n = 500
    g, o = make_pair(mods, {'env_num': n, '_seed': 4})
    g.reset(); o.reset()
    rng = np.random.default_rng(8)
    s = o.get_state()
    s[:, 0:2] = rng.uniform(-2.5, 2.5, (n, 2)); s[:, 2] = rng.uniform(-30, 30, n)
    s[:, 3:5] = rng.uniform(-3, 3, (n, 2)); s[:, 5] = rng.uniform(-12, 12, n)
    th = rng.uniform(-3, 3, n)
    s[:, 6], s[:, 7] = np.cos(th), np.sin(th)
    s[:, 8] = rng.uniform(0, 4, n); s[:, 9] = rng.integers(0, 2, n); s[:, 10] = np.where(s[:, 9] > 0, 0, rng.integers(0, 900, n))
    s = s.astype(np.float32)
    g.set_state(torch.from_numpy(s).cuda()); o.set_state(s)
    bits_equal(cpu(g.get_state()), s, "get(set(s))")
    a = rng.uniform(-1, 1, (n, 2)).astype(np.float32)
    og, rg, dg, ig = g.step(torch.from_numpy(a).cuda(), want_aux=True)
    oo, ro, do, io = o.step(a, want_aux=True)
    bits_equal(cpu(og), oo, "obs"); bits_equal(cpu(rg), ro, "reward"); bits_equal(cpu(dg), do, "done")
    bits_equal(cpu(ig['aux']), io['aux'], "aux")
    bits_equal(cpu(g.get_state()), o.get_state(), "state")


def test_golden_reference_source_vectors_through_cuda(mods, golden_dir):
    """The vectors produced by the reference's own obs/cost code
    (tools/make_golden.py) replayed through the CUDA step by state injection: the
    geometry of a step's observation comes from the PRE-step pose, so injecting
    (robot pose, goal, hazards) and stepping once reproduces the golden lidar /
    compass / cost.  Tolerance 2e-5: the golden rotation is float32(cos/sin theta),
    the kernel's comes from the MJX quaternion path."""
    Engine, _ = mods
    gold = np.load(os.path.join(golden_dir, "refsrc_obs_default.npz"))
    n = gold['flat'].shape[0]
    g = Engine({'env_num': n})
    g.reset()
    s = cpu(g.get_state())
    s[:, 0:2] = gold['robot']; s[:, 2] = gold['qpos'][:, 2]; s[:, 3:6] = 0
    s[:, 6] = 1; s[:, 7] = 0; s[:, 8] = 0; s[:, 9] = 1; s[:, 10] = 0
    s[:, 11:13] = gold['goal']; s[:, 13:29] = gold['hazards'].reshape(n, -1)
    g.set_state(torch.from_numpy(s).cuda())
    obs, r, d, info = g.step(torch.zeros(n, 2, device='cuda'), want_aux=True)
    obs = cpu(obs)
    flat = gold['flat']
    np.testing.assert_allclose(obs[:, 3:5], flat[:, 3:5], rtol=0, atol=2e-5)      # compass
    np.testing.assert_allclose(obs[:, 5:37], flat[:, 5:37], rtol=0, atol=2e-5)    # both lidars
    np.testing.assert_allclose(cpu(info['cost']), gold['cost'], rtol=0, atol=2e-6)
    dist = np.linalg.norm(gold['goal'] - gold['robot'], axis=1)
    clear = np.abs(dist - 0.5) > 1e-5
    assert ((cpu(d) > 0) == (dist < 0.5))[clear].all()
    assert (cpu(r) == 0).all()  # prev_done injected as 1 => reward 0 (engine.py:795)


@pytest.mark.parametrize("pinned", [True, False])
def test_step_host_matches_step(mods, pinned):
    """pinned + mapped buffers take the zero-copy route (the kernel reads / writes host memory),
    pageable ones the staged-copy route; both equal step() + reset_done()."""
    Engine, _ = mods
    cfg = {'env_num': 1000, 'num_steps': 20, '_seed': 6}
    a, b = Engine(dict(cfg)), Engine(dict(cfg))
    a.reset(); b.reset()
    n, od = 1000, a.obs_flat_size
    h_act = torch.empty(n, 2)
    h_out = torch.empty(n * (od + 3))
    if pinned:
        h_act, h_out = h_act.pin_memory(), h_out.pin_memory()
    rng = np.random.default_rng(1)
    for t in range(50):
        act = torch.from_numpy(rng.uniform(-1, 1, (n, 2)).astype(np.float32))
        h_act.copy_(act)
        a.step_host(h_act, h_out, auto_reset_done=True)
        o, r, d, info = b.step(act.cuda())
        o = b.reset_done()
        assert torch.equal(h_out[:n * od].view(n, od), o.cpu())
        assert torch.equal(h_out[n * od:n * od + n], r.cpu())
        assert torch.equal(h_out[n * od + n:n * od + 2 * n], d.cpu())
        assert torch.equal(h_out[n * od + 2 * n:], info['cost'].cpu())


@pytest.mark.parametrize("n,T", [(1000, 37), (4096, 70), (70000, 5)])
def test_rollout_host_matches_step_host(mods, n, T):
    """The pipelined host rollout (chunked, two streams) == T synchronous host steps,
    including a ragged last chunk and a second call reusing the ring."""
    Engine, _ = mods
    cfg = {'env_num': n, 'num_steps': 20, '_seed': 9}
    a, b = Engine(dict(cfg)), Engine(dict(cfg))
    a.reset(); b.reset()
    od = a.obs_flat_size
    h_out = torch.empty(n * (od + 3)).pin_memory()
    for rep in range(6):      # calls 0..3 alternate the DMA ring and the zero-copy route (probing), 4.. use the faster one
        h_act = torch.empty(T, n, 2).uniform_(-1, 1).pin_memory()
        h_obs = torch.empty(T, n, od).pin_memory()
        h_r, h_d, h_c = (torch.empty(T, n).pin_memory() for _ in range(3))
        a.rollout_host(h_act, h_obs, h_r, h_d, h_c)
        assert a._lib.gx_rollout_host_route(a._h, None) == (-1 if rep < 3 else a._lib.gx_rollout_host_route(a._h, None))
        for t in range(T):
            b.step_host(h_act[t], h_out, auto_reset_done=True)
            assert torch.equal(h_out[:n * od].view(n, od), h_obs[t]), (rep, t)
            assert torch.equal(h_out[n * od:n * od + n], h_r[t])
            assert torch.equal(h_out[n * od + n:n * od + 2 * n], h_d[t])
            assert torch.equal(h_out[n * od + 2 * n:], h_c[t])
    assert a._lib.gx_rollout_host_route(a._h, None) in (0, 1)
    # pageable outputs cannot be mapped: the DMA ring is the only route, results unchanged
    h_act = torch.empty(T, n, 2).uniform_(-1, 1).pin_memory()
    p_obs, p_r, p_d, p_c = torch.empty(T, n, od), torch.empty(T, n), torch.empty(T, n), torch.empty(T, n)
    a.rollout_host(h_act, p_obs, p_r, p_d, p_c)
    for t in range(T):
        b.step_host(h_act[t], h_out, auto_reset_done=True)
        assert torch.equal(h_out[:n * od].view(n, od), p_obs[t]) and torch.equal(h_out[n * od + 2 * n:], p_c[t])


def test_episode_statistics_match_oracle(mods):
    n = 300
    g, o = make_pair(mods, {'env_num': n, 'num_steps': 30, '_seed': 12})
    g.reset(); o.reset()
    g.stats_reset(); o.stats_reset()
    g.rollout(200); o.rollout(200)
    sg, so = cpu(g.stats()), o.stats()
    assert sg[0] == so[0] and sg[3] == so[3] and sg[7] == so[7] and so[0] > n   # counts: exact
    # sums: per-env fp32 partials reduced in fp64 on the GPU vs serial fp64 in the oracle
    np.testing.assert_allclose(sg, so, rtol=2e-6)
    a = np.zeros((n, 2), np.float32)
    g.step(torch.from_numpy(a).cuda()); o.step(a)
    np.testing.assert_allclose(cpu(g.stats()), o.stats(), rtol=2e-6)
    assert torch.equal(g.stats(), g.stats())   # deterministic reduction


# ------------------------------------------------------- BASELINE full sizes
@pytest.mark.parametrize("n,T", [(4096, 200), (16384, 60), (65536, 24)])
def test_full_size_against_oracle(mods, n, T):
    """BASELINE.json configs[1], [2]-size and [3]-size batches (Point robot),
    every output bit-compared with the OpenMP oracle."""
    g, o = make_pair(mods, {'env_num': n, 'num_steps': 15, '_seed': n}, threads=os.cpu_count() or 8)
    bits_equal(cpu(g.reset()), o.reset(), "reset")
    rg = g.rollout(T, want=("obs", "reward", "cost", "done"))
    ro = o.rollout(T, want=("obs", "reward", "cost", "done"))
    for k in ("done", "reward", "cost", "obs"):
        bits_equal(cpu(rg[k]), ro[k], f"{k}")
    bits_equal(cpu(g.get_state()), o.get_state(), "state")
    # size-independent properties
    obs = ro["obs"]
    assert ((obs[..., 5:37] >= 0) & (obs[..., 5:37] <= 1)).all()       # lidar range
    assert np.isin(ro["done"], (0.0, 1.0)).all() and (ro["cost"] >= 0).all()
    assert ro["done"].sum() >= n  # every env timed out at least once


def test_sharding_is_invariant_to_the_number_of_ranks(mods):
    """global_env_id = offset + local id keys every draw, so 2 shards of
    512 reproduce one batch of 1024."""
    Engine, _ = mods
    cfg = {'env_num': 1024, 'num_steps': 20, '_seed': 21}
    whole = Engine(dict(cfg))
    half = dict(cfg, env_num=512)
    lo, hi = Engine(dict(half), env_id_offset=0, pool_size=whole.layout_size), Engine(dict(half), env_id_offset=512, pool_size=whole.layout_size)
    ow = whole.reset()
    assert torch.equal(torch.cat([lo.reset(), hi.reset()]), ow)
    rw = whole.rollout(50)
    rl, rh = lo.rollout(50), hi.rollout(50)
    for k in ("obs", "reward", "cost", "done", "act"):
        assert torch.equal(torch.cat([rl[k], rh[k]], dim=1), rw[k]), k


def test_api_surface(mods):
    Engine, _ = mods
    import pickle
    from guardx_b200 import create_env, make
    from types import SimpleNamespace
    env = create_env(SimpleNamespace(task='Goal_Point_8Hazards', env_num=16, seed=0, max_ep_len=200))
    assert env.observation_space.shape == (43,) and env.action_space.shape == (2,)
    assert (env.action_space.low == -1).all() and (env.action_space.high == 1).all()
    assert env.num_steps == 200 and env.env_num == 16
    o = env.reset()
    assert o.shape == (16, 43) and o.dtype == torch.float32 and o.is_cuda
    o2, r, d, info = env.step(torch.zeros(16, 2, device='cuda'))
    assert 'cost' in info.keys() and info['cost'].shape == (16,) and r.shape == (16,) and d.shape == (16,)
    assert sorted(info['obs'].keys()) == ['ctrl', 'goal_compass', 'goal_lidar', 'hazards_lidar', 'qacc', 'qpos', 'qvel']
    assert torch.equal(info['obs']['goal_lidar'], o2[:, 5:21]) and info['obs']['qacc'].shape == (16, 3)
    # returned tensors are fresh and stay valid across the next step (SURVEY 8b ownership)
    keep = o2.clone()
    o2.add_(0)  # writable
    env.step(torch.ones(16, 2, device='cuda'))
    assert torch.equal(o2, keep)
    o3 = env.reset_done()
    assert o3.data_ptr() != o2.data_ptr()
    with pytest.raises(AssertionError, match='Bad key'):
        Engine({'observe_ghosts': True})
    with pytest.raises(NotImplementedError):
        env.render()
    env2 = pickle.loads(pickle.dumps(env))
    assert env2.env_num == 16 and env2.num_steps == 200
    assert make(env_num=8).reset().shape == (8, 43)
    env.close()


def test_step_autoreset_equals_step_then_reset_done(mods):
    """The fused launch returns exactly what step() followed by reset_done() returns."""
    Engine, _ = mods
    cfg = {'env_num': 2500, 'num_steps': 12, '_seed': 14}
    a, b = Engine(dict(cfg)), Engine(dict(cfg))
    a.reset(); b.reset()
    g = torch.Generator(device='cuda').manual_seed(0)
    n_done = 0
    for t in range(60):
        act = torch.rand(2500, 2, device='cuda', generator=g) * 2 - 1
        oa, ra, da, ia = a.step_autoreset(act)
        ob, rb, db, ib = b.step(act)
        ob = b.reset_done()
        assert torch.equal(oa.view(torch.int32), ob.view(torch.int32)), t
        assert torch.equal(ra.view(torch.int32), rb.view(torch.int32)) and torch.equal(da, db)
        assert torch.equal(ia['cost'].view(torch.int32), ib['cost'].view(torch.int32))
        n_done += int(da.sum().item())
    assert n_done > 2500
    assert torch.equal(a.get_state().view(torch.int32), b.get_state().view(torch.int32))


def test_cuda_graph_capture_and_replay(mods):
    """Counters live in device memory, so a captured graph of fused steps replays
    correctly: 5 replays of a 6-step graph == 30 direct steps on a twin env."""
    Engine, _ = mods
    cfg = {'env_num': 2048, 'num_steps': 9, '_seed': 17}
    a, b = Engine(dict(cfg)), Engine(dict(cfg))
    a.reset(); b.reset()
    acts = torch.rand(6, 2048, 2, device='cuda') * 2 - 1
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        a.step_autoreset(acts[0]); b.step_autoreset(acts[0])   # warm-up outside capture
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        outs = [a.step_autoreset(acts[k]) for k in range(6)]
    assert a.counters()[0] == 1          # capture does not execute
    n_done = 0
    for rep in range(5):
        g.replay()
        torch.cuda.synchronize()
        for k in range(6):
            o, r, d, info = b.step_autoreset(acts[k])
            go, gr, gd, gi = outs[k]
            assert torch.equal(go.view(torch.int32), o.view(torch.int32)), (rep, k)
            assert torch.equal(gr.view(torch.int32), r.view(torch.int32)) and torch.equal(gd, d)
            assert torch.equal(gi['cost'].view(torch.int32), info['cost'].view(torch.int32))
            n_done += int(d.sum().item())
    assert n_done > 2048
    assert a.counters() == b.counters() == (31, 1)
    assert torch.equal(a.get_state().view(torch.int32), b.get_state().view(torch.int32))


def test_counters_roundtrip(mods):
    g, o = make_pair(mods, {'env_num': 64, 'num_steps': 5, '_seed': 33})
    g.set_counters(1 << 33, 7); o.set_counters(1 << 33, 7)   # > 32-bit step counts key correctly
    bits_equal(cpu(g.reset()), o.reset(), "reset at epoch 7")
    assert g.counters() == (1 << 33, 8)
    rg, ro = g.rollout(30), o.rollout(30)
    for k in ("obs", "act", "done"):
        bits_equal(cpu(rg[k]), ro[k], k)
    assert g.counters() == ((1 << 33) + 30, 8)
