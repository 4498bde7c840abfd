"""Doggo on the general-tree CUDA kernels (gx_tree.cuh: free-joint root, capsule contacts with the
floor and between limbs, dense Newton) vs the CPU oracle (oracle/art_oracle.py around
oracle/mjx_numpy.py).  Independent formulations => stated tolerances (see tests/test_art_gpu.py);
PARITY UNPINNED against real MJX like every mjx.step restatement here."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
NU, NQ, NV = 12, 20, 19


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from guardx_b200.engine import Engine
    from oracle.art_oracle import ArtOracleEngine
    from oracle.oracle import OracleEngine
    base = OracleEngine({'env_num': 64, '_seed': 5})
    base.reset()
    pool, _ = base.pool()
    return Engine, ArtOracleEngine, pool


def cpu(t):
    return t.detach().cpu().numpy()


def close(a, b, atol, rtol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f'{what}: max excess {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}'


def test_doggo_spaces_reset_and_one_step_parity(mods):
    Engine, Oracle, pool = mods
    n = 32
    cfg = {'env_num': n, 'robot_base': 'xmls/doggo.xml', 'num_steps': 25, '_seed': 2}
    g, o = Engine(dict(cfg)), Oracle(dict(cfg), dtype=np.float64)
    assert g.observation_space.shape == (NU + 34 + NQ + NV,) and g.action_space.shape == (NU,)
    assert g.robot.nq == NQ and g.robot.nv == NV
    close(cpu(g.reset(pool[:n])), o.reset(pool[:n]), 2e-6, 2e-6, 'reset obs')
    rng = np.random.default_rng(3)
    k = NU + 34 + NQ
    n_done = n_rows = 0
    for t in range(40):
        a = rng.uniform(-1.2, 1.2, (n, NU)).astype(np.float32)
        o.round_state_to_fp32()
        g.set_state(o.get_state())
        og, rg, dg, ig = g.step(torch.from_numpy(a).cuda(), want_aux=True)
        oo, ro, do, io = o.step(a)
        close(cpu(og)[:, :k], oo[:, :k], 5e-5, 5e-5, f'step {t} obs (ctrl..qpos)')
        close(cpu(og)[:, k:], oo[:, k:], 2e-3, 5e-4, f'step {t} qvel')
        close(cpu(rg), ro, 5e-5, 0, f'step {t} reward')
        close(cpu(ig['cost']), io['cost'], 5e-5, 0, f'step {t} cost')
        qs = 1 + np.abs(io['qacc']).max(1, keepdims=True)
        close(cpu(ig['obs']['qacc']) / qs, io['qacc'] / qs, 5e-4, 0, f'step {t} qacc relative to the env max')
        clear = (np.abs(io['dist'] - 0.5) > 1e-4) & (np.abs(np.abs(io['d_dist']) - 1) > 1e-4)
        assert (cpu(dg) == do)[clear].all()
        aux = cpu(ig['aux'])
        bins_o = np.concatenate([io['goal_bin'], io['hazard_bins']], 1)
        assert (aux[:, 2:11] != bins_o).sum() <= 1
        n_done += int(do.sum())
        n_rows = max(n_rows, int((o.last_data['efc_D'] > 0).sum(1).max()))
        if do.any():
            lay = pool[n * (1 + t % 3):n * (2 + t % 3)]
            g.set_state(o.get_state())
            close(cpu(g.reset_done(lay))[:, :k], o.reset_done(lay)[:, :k], 5e-5, 5e-5, f'step {t} reset_done obs')
    assert n_done >= n and n_rows >= 12        # timeouts fired; limits and capsule contacts really were active


def test_doggo_world_bodies_creep_like_in_the_reference_model(mods):
    """Doggo's defaults class gives the goal / hazard slide joints a stiffness (world.py:316-319 writes only
    limited / damping): they drift towards the origin, in the kernel as in the full-model oracle."""
    Engine, Oracle, pool = mods
    n = 8
    cfg = {'env_num': n, 'robot_base': 'xmls/doggo.xml', 'num_steps': 1000, '_seed': 4}
    g, o = Engine(dict(cfg)), Oracle(dict(cfg), dtype=np.float64)
    g.reset(pool[:n]); o.reset(pool[:n])
    z = np.zeros((n, NU), np.float32)
    for t in range(30):
        g.step(torch.from_numpy(z).cuda()); o.step(z)
    sg, so = cpu(g.get_state()), o.get_state()
    goal0 = pool[:n, :2]
    goal_g, goal_o = sg[:, NQ + 2 * NV + 3:NQ + 2 * NV + 5], so[:, NQ + 2 * NV + 3:NQ + 2 * NV + 5]
    assert (np.linalg.norm(goal_g, axis=1) < np.linalg.norm(goal0, axis=1)).all()      # moved towards the origin
    close(goal_g, goal_o, 1e-6, 1e-6, 'goal position after 30 steps')
    close(sg[:, -18:], so[:, -18:], 1e-7, 1e-5, 'goal / hazard slide velocities')


def test_doggo_autoreset_rollout_stats(mods):
    Engine, _, _ = mods
    n = 600
    cfg = {'env_num': n, 'robot_base': 'xmls/doggo.xml', 'num_steps': 6, '_seed': 6}
    a_env, b_env = Engine(dict(cfg)), Engine(dict(cfg))
    a_env.reset(); b_env.reset()
    a_env.stats_reset()
    ro = a_env.rollout(16)
    assert ro['obs'].shape == (16, n, 85) and torch.isfinite(ro['obs']).all()
    for t in range(16):
        o, r, d, i = b_env.step_autoreset(ro['act'][t])
        assert torch.equal(o, ro['obs'][t]) and torch.equal(d, ro['done'][t]) and torch.equal(r, ro['reward'][t])
    s = a_env.stats().cpu().numpy()
    assert s[7] == n * 16 and s[0] == float(ro['done'].sum().item()) and s[0] >= n
    # step + reset_done == step_autoreset
    c_env, d_env = Engine(dict(cfg)), Engine(dict(cfg))
    c_env.reset(); d_env.reset()
    for t in range(12):
        act = ro['act'][t]
        oa, ra, da, ia = c_env.step_autoreset(act)
        ob, rb, db, ib = d_env.step(act)
        ob = d_env.reset_done()
        assert torch.equal(oa, ob) and torch.equal(da, db)


def test_ant_through_the_general_tree_kernels_equals_the_leg_per_lane_kernels(mods):
    """The general-tree kernels are robot-agnostic: the ant lowered with tree_model runs through them and
    must agree with the specialised leg-per-lane kernels (same model, two CUDA formulations)."""
    import ctypes as C
    Engine, _, pool = mods
    from guardx_b200 import _abi, config as gxconfig, tree_model
    n = 64
    cfg = gxconfig.parse({'env_num': n, 'robot_base': 'xmls/ant.xml', 'num_steps': 1000, '_seed': 1})
    model = gxconfig.compiled_model(cfg)
    art = Engine(dict(env_num=n, robot_base='xmls/ant.xml', num_steps=1000, _seed=1))
    gxc = gxconfig.lower(cfg, 0, 0, model)
    gxc.robot = _abi.GX_ROBOT_TREE
    tm, info = tree_model.lower_tree(model)
    lib = _abi.load_library()
    h = C.c_void_p()
    _abi.check(lib, lib.gx_create_tree(C.byref(gxc), C.byref(tm), 0, C.byref(h)), "gx_create_tree")
    p = lambda t_: C.c_void_p(t_.data_ptr())
    lay = torch.from_numpy(np.ascontiguousarray(pool[:n])).cuda()
    obs_t = torch.empty(n, 64, device='cuda'); rdc = torch.empty(3, n, device='cuda')
    _abi.check(lib, lib.gx_reset_with_layout(h, p(lay), p(obs_t), None), "reset")
    obs_a = art.reset(lay)
    close(cpu(obs_t), cpu(obs_a), 2e-6, 2e-6, 'reset obs')
    gen = torch.Generator(device='cuda').manual_seed(0)
    for t in range(15):
        act = (torch.rand(n, 8, device='cuda', generator=gen) * 2 - 1).contiguous()
        _abi.check(lib, lib.gx_step(h, p(act), p(obs_t), p(rdc[0]), p(rdc[1]), p(rdc[2]), None, None, None), "step")
        oa, ra, da, ia = art.step(act)
        close(cpu(obs_t)[:, :53], cpu(oa)[:, :53], 2e-4, 1e-4, f'step {t} obs')
        close(cpu(rdc[2]), cpu(ia['cost']), 2e-4, 0, f'step {t} cost')
        assert torch.equal(rdc[1], da)
    lib.gx_destroy(h)


def test_observe_vel_acc_for_doggo_uses_the_full_3d_pose(mods):
    """ego_vel_acc (engine.py:902-929) for a free-joint root: finite differences of all three xpos
    components, rotated by the full xmat (the falling, tilting torso mixes z in); done masking and the
    None-ness of the first two steps as for the other robots; obs = acc | default | vel."""
    Engine, Oracle, pool = mods
    n = 16
    cfg = {'env_num': n, 'robot_base': 'xmls/doggo.xml', 'num_steps': 6, '_seed': 2, 'observe_vel': True, 'observe_acc': True}
    g, o = Engine(dict(cfg)), Oracle(dict(cfg), dtype=np.float64)
    assert g.observation_space.shape == (4 + NU + 34 + NQ + NV,)
    lay = pool[:n]
    close(cpu(g.reset(lay)), o.reset(lay), 2e-6, 2e-6, 'reset obs (vel = acc = 0)')
    rng = np.random.default_rng(5)
    seen_done, zmix = 0, 0.0
    k0, k1 = 2, 2 + NU + 34 + NQ
    for t in range(16):
        a = rng.uniform(-1.2, 1.2, (n, NU)).astype(np.float32)
        o.round_state_to_fp32()
        g.set_state(o.get_state())          # joint state injected; the xpos history evolves on each side
        og, rg, dg, ig = g.step(torch.from_numpy(a).cuda())
        oo, ro, do, io = o.step(a)
        og = cpu(og)
        close(og[:, k0:k1], oo[:, k0:k1], 5e-5, 5e-5, f'step {t} default part')
        close(og[:, -2:], oo[:, -2:], 5e-4, 1e-4, f'step {t} vel')          # d(xpos)/dt, dt = 0.01... of fp32 positions
        close(og[:, :2], oo[:, :2], 5e-2, 1e-3, f'step {t} acc')            # second difference / dt^2
        assert (cpu(dg) == do).all()
        seen_done += int(do.sum())
        zmix = max(zmix, float(np.abs(io['xmat'].reshape(-1, 3, 3)[:, 2, :2]).max()))
        if do.any():
            g.set_state(o.get_state())
            l2 = pool[n * (1 + t % 3):n * (2 + t % 3)]
            close(cpu(g.reset_done(l2))[:, k0:k1], o.reset_done(l2)[:, k0:k1], 5e-5, 5e-5, 'reset_done obs')
    assert seen_done >= n and np.abs(oo[:, -2:]).max() > 1e-3 and zmix > 1e-3   # the z row of xmat really took part


def test_doggo_host_buffer_paths_and_batch_independence(mods):
    """Engine.rollout_host (pipelined H2D / kernels / D2H) == T x Engine.step_host for Doggo; and an env's
    trajectory does not depend on the batch it sits in (first rows of a 2 000-env batch == a 96-env batch,
    bit for bit: one thread per env, no cross-env arithmetic, results independent of block / SM placement)."""
    Engine, _, pool = mods
    n, T = 96, 6
    cfg = {'env_num': n, 'robot_base': 'xmls/doggo.xml', 'num_steps': 4, '_seed': 8}
    a_env, b_env = Engine(dict(cfg)), Engine(dict(cfg))
    a_env.reset(); b_env.reset()
    od = a_env.obs_flat_size
    h_act = torch.empty(T, n, NU).uniform_(-1, 1).pin_memory()
    h_obs = torch.empty(T, n, od).pin_memory()
    h_r, h_d, h_c = (torch.empty(T, n).pin_memory() for _ in range(3))
    a_env.rollout_host(h_act, h_obs, h_r, h_d, h_c)
    h_out = torch.empty(n * (od + 3)).pin_memory()
    for t in range(T):
        b_env.step_host(h_act[t], h_out, auto_reset_done=True)
        assert torch.equal(h_out[:n * od].view(n, od), h_obs[t])
        assert torch.equal(h_out[n * od + n:n * od + 2 * n], h_d[t])
    assert torch.isfinite(h_obs).all() and h_d.sum() >= n
    big = 2000
    s_env = Engine({'env_num': n, 'robot_base': 'xmls/doggo.xml', 'num_steps': 1000, '_seed': 1})
    l_env = Engine({'env_num': big, 'robot_base': 'xmls/doggo.xml', 'num_steps': 1000, '_seed': 1})
    lay_s = pool[:n]
    lay_l = np.concatenate([lay_s] + [pool[:big - n]], 0)
    os_, ol = s_env.reset(lay_s), l_env.reset(lay_l)
    assert torch.equal(os_, ol[:n])
    gen = torch.Generator(device='cuda').manual_seed(3)
    for t in range(10):
        act = torch.empty(big, NU, device='cuda').uniform_(-1, 1, generator=gen)
        o1, r1, d1, i1 = s_env.step(act[:n].contiguous())
        o2, r2, d2, i2 = l_env.step(act)
        assert torch.equal(o1, o2[:n]) and torch.equal(r1, r2[:n]) and torch.equal(i1['cost'], i2['cost'][:n])


def test_doggo_step_replays_from_a_cuda_graph(mods):
    """The Doggo kernels take the compiled model (11 KB) as a by-value kernel parameter; a captured step must
    replay with it: graph replays == eager gx_step_autoreset calls, bit for bit."""
    import ctypes as C
    Engine, _, _ = mods
    n = 160
    cfg = {'env_num': n, 'robot_base': 'xmls/doggo.xml', 'num_steps': 5, '_seed': 4}
    a_env, b_env = Engine(dict(cfg)), Engine(dict(cfg))
    a_env.reset(); b_env.reset()
    act = torch.empty(n, NU, device='cuda').uniform_(-1, 1)
    obs = torch.empty(n, 85, device='cuda'); rdc = torch.empty(3, n, device='cuda')
    p = lambda t_: C.c_void_p(t_.data_ptr())
    lib = a_env._lib
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        sp = C.c_void_p(side.cuda_stream)
        with torch.cuda.graph(g, stream=side):
            lib.gx_step_autoreset(a_env._h, p(act), p(obs), p(rdc[0]), p(rdc[1]), p(rdc[2]), sp)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    for t in range(8):                       # capture does not execute: both envs start from the same state
        act.uniform_(-1, 1)
        g.replay()
        torch.cuda.synchronize()
        o, r, d, i = b_env.step_autoreset(act)
        assert torch.equal(o, obs) and torch.equal(r, rdc[0]) and torch.equal(d, rdc[1]) and torch.equal(i['cost'], rdc[2])
    assert torch.isfinite(obs).all()


def test_doggo_config4_size_one_step_vs_oracle(mods):
    """65 536 Doggo envs per GPU use the per-block chunks of the GLOBAL scratch (gx_tree.cuh) in a grid of 2048 blocks;
    the small tests above run one or two blocks.  Roll the big batch 40 steps on its own, hand 512 random envs to the
    float64 oracle, one step from bit-identical states with the same actions; tolerances of the small test."""
    Engine, Oracle, pool = mods
    n, k = 65536, 512
    g = Engine({'env_num': n, 'robot_base': 'xmls/doggo.xml', 'num_steps': 1000, '_seed': 4})
    g.reset()
    for _ in range(5):
        out = g.rollout(8)
        del out
    st = cpu(g.get_state())
    rng = np.random.default_rng(22)
    finite = np.isfinite(st).all(1)
    assert finite.mean() > 0.999
    ids = rng.choice(np.flatnonzero(finite), k, replace=False)
    o = Oracle({'env_num': k, 'robot_base': 'xmls/doggo.xml', 'num_steps': 1000, '_seed': 4}, dtype=np.float64)
    o.reset(pool[:k])
    o.set_state(st[ids])
    a = rng.uniform(-1.2, 1.2, (n, NU)).astype(np.float32)
    og, rg, dg, ig = g.step(torch.from_numpy(a).cuda(), want_aux=True)
    oo, ro, do, io = o.step(a[ids])
    kq = NU + 34 + NQ
    close(cpu(og)[ids][:, :kq], oo[:, :kq], 5e-5, 5e-5, 'doggo 65536-env grid: obs (ctrl..qpos)')
    close(cpu(og)[ids][:, kq:], oo[:, kq:], 2e-3, 5e-4, 'qvel')
    close(cpu(rg)[ids], ro, 5e-5, 0, 'reward')
    close(cpu(ig['cost'])[ids], io['cost'], 5e-5, 0, 'cost')
    qs = 1 + np.abs(io['qacc']).max(1, keepdims=True)
    close(cpu(ig['obs']['qacc'])[ids] / qs, io['qacc'] / qs, 5e-4, 0, 'qacc relative to the env max')
    clear = (np.abs(io['dist'] - 0.5) > 1e-4) & (np.abs(np.abs(io['d_dist']) - 1) > 1e-4)
    assert (cpu(dg)[ids] == do)[clear].all()
    aux = cpu(ig['aux'])[ids]
    bins_o = np.concatenate([io['goal_bin'], io['hazard_bins']], 1)
    assert (aux[:, 2:11] != bins_o).sum() <= 2
    assert int((o.last_data['efc_D'] > 0).sum(1).max()) >= 12      # the robots have landed: capsule contacts active
