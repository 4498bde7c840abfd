"""Config-3 extension (Button / Push / Goal x Car / Point among hazards, vases, pillars, gremlins; include/guardx_task.h)
on the CUDA kernels vs oracle/task_oracle.py.

NO REFERENCE, PARITY UNPINNED BY CONSTRUCTION: both sides implement the specification of DESIGN.md
section 8 -- the kernel in fp32 with the repo's deterministic elementary functions, the oracle in float64 numpy, written
independently.  Tolerances: one step from bit-identical injected states, floats |d| <= 2e-5 + 2e-5 |x|; done flags,
violation / contact counts exact wherever the deciding distance is further than 1e-5 from its threshold; lidar bins may
differ for rays within float rounding of a bin edge (counted, bounded).
"""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from guardx_b200.task_engine import TaskEngine, make_task
    from oracle.task_oracle import TaskOracle
    return TaskEngine, make_task, TaskOracle


def params_of(env):
    """plain-number view of the C ABI config struct the env was created with (what the oracle is parameterised by)"""
    c = env._gxt_config
    P = {}
    for name, _ in c._fields_:
        v = getattr(c, name)
        P[name] = list(v) if hasattr(v, '__len__') else v
    return P


def cpu(t):
    return t.detach().cpu().numpy()


def close(a, b, tol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (tol + tol * np.abs(b))
    assert (err <= 0).all(), f'{what}: max excess {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}'


@pytest.mark.parametrize("task", ["Button_Car", "Push_Car", "Goal_Car", "Button_Point", "Push_Point", "Goal_Point"])
def test_one_step_from_injected_state_vs_oracle(mods, task):
    TaskEngine, make_task, TaskOracle = mods
    n = 96
    g = make_task(task, env_num=n, seed=3, max_ep_len=70, hazards_num=8, vases_num=8, pillars_num=4, gremlins_num=6,
                  placements_extents=[-3.5, -3.5, 3.5, 3.5])
    o = TaskOracle(params_of(g))
    assert g.obs_flat_size == o.obs_dim and g.state_dim == o.state_dim
    obs0 = g.reset()
    o.set_state(cpu(g.get_state()))
    close(cpu(obs0), o.reset_obs()[0], 2e-5, 'reset obs')
    # half of the robots start right behind vase 0 (or the box), moving towards it: pushes happen from the first steps on
    st = cpu(g.get_state())
    counts = list(g._gxt_config.count)
    tgt = 16 + 2 * (sum(counts[:1]) if 'Push' in task else sum(counts[:4]))         # box, else vase 0
    gap = float(g._gxt_config.size[7]) + 0.25
    st[::2, 0], st[::2, 1], st[::2, 2] = st[::2, tgt] - gap, st[::2, tgt + 1], 0.0
    st[::2, 3], st[::2, 4], st[::2, 5] = 0.8, 0.0, 0.0
    st[::2, 6], st[::2, 7] = 1.0, 0.0
    g.set_state(torch.from_numpy(st).cuda())
    rng = np.random.default_rng(5)
    n_done = n_contact = n_viol = bin_miss = moved = 0
    a_hold = rng.uniform(-1, 1, (n, 2))
    for t in range(160):
        if t % 20 == 0:                                      # held, mostly-forward actions: the robots travel and hit things
            a_hold = rng.uniform(-0.2, 1.2, (n, 2)) if 'Car' in task else np.stack([rng.uniform(0.2, 1.2, n), rng.uniform(-0.4, 0.4, n)], 1)
        a = a_hold.astype(np.float32)
        st0 = cpu(g.get_state())
        o.set_state(st0)
        og, rg, dg, ig = g.step(torch.from_numpy(a).cuda(), want_aux=True)
        oo, ro, do, io = o.step(a)
        close(cpu(og), oo, 2e-5, f'{task} step {t} obs')
        close(cpu(rg), ro, 2e-5, f'step {t} reward')
        aux = cpu(ig['aux'])
        sure = io['margin'] > 1e-5
        close(cpu(ig['cost'])[sure], io['cost'][sure], 2e-5, f'step {t} cost')
        assert (aux[:, 1] == io['violations'])[sure].all() and (aux[:, 2] == io['contacts'])[sure].all()
        clear = (np.abs(io['dist'] - io['success_threshold']) > 1e-5) & (np.abs(np.abs(io['d_dist']) - 1) > 1e-5)
        assert (cpu(dg) == do)[clear].all() and (aux[:, 0] == io['flags'])[clear].all()
        bin_miss += int((aux[:, 3] != io['goal_bin']).sum())
        st1 = cpu(g.get_state())
        keep = np.ones(st1.shape[1], bool); keep[[12, 13]] = False        # age: bit patterns of a counter
        close(st1[:, keep][clear], o.get_state()[:, keep][clear], 2e-5, f'step {t} state')
        moved += int((np.abs(st1[:, 16:] - st0[:, 16:]) > 0).any(1).sum())
        n_done += int(do.sum()); n_contact += int(io['contacts'].sum()); n_viol += int(io['violations'].sum())
        if do.any():
            g.reset_done()
    assert n_done >= n and n_viol > 0                       # timeouts fired for every env; hazards were entered
    assert n_contact > 0                                    # vases / pillars / gremlins / wrong buttons were touched
    assert moved > 0                                        # vases (and the box) really were pushed
    assert bin_miss <= 2


def test_layout_pool_respects_keepouts_and_extents(mods):
    TaskEngine, make_task, _ = mods
    g = make_task("Button_Car", env_num=64, seed=1)
    g.reset()
    pool, tries = (cpu(x) for x in g.layout_pool())
    assert (tries > 0).all()
    c = g._gxt_config
    counts = list(c.count)
    cls = np.concatenate([[k] * counts[k] for k in range(7)] + [[7]]).astype(int)
    keep = np.array(list(c.keepout))
    nobj = len(cls) - 1
    xy = pool[:, :2 * nobj + 2].reshape(len(pool), nobj + 1, 2)
    ext = np.array(list(c.placements_extents))
    for k in range(nobj + 1):
        assert (xy[:, k, 0] >= ext[0] + keep[cls[k]] - 1e-6).all() and (xy[:, k, 0] <= ext[2] - keep[cls[k]] + 1e-6).all()
        assert (xy[:, k, 1] >= ext[1] + keep[cls[k]] - 1e-6).all() and (xy[:, k, 1] <= ext[3] - keep[cls[k]] + 1e-6).all()
        for q in range(k):
            d = np.linalg.norm(xy[:, k] - xy[:, q], axis=1)
            assert (d >= keep[cls[k]] + keep[cls[q]] - 1e-6).all(), (k, q)
    ph = pool[:, 2 * nobj + 2:]
    assert ph.shape[1] == counts[6] and (ph >= 0).all() and (ph < 2 * np.pi + 1e-6).all()
    assert len(np.unique(pool[:, 0])) > len(pool) // 2      # slots differ
    # goal task: the JAX Engine's robot-goal distance rule (engine.py:571)
    g2 = make_task("Goal_Point", env_num=32, seed=1, vases_num=2, pillars_num=1, gremlins_num=1, hazards_num=4)
    g2.reset()
    p2, t2 = (cpu(x) for x in g2.layout_pool())
    nobj2 = g2.layout_dim // 2 - 1 if False else (g2.layout_dim - 2 - 1) // 2
    assert (t2 > 0).all() and (np.linalg.norm(p2[:, 2 * nobj2:2 * nobj2 + 2] - p2[:, 0:2], axis=1) >= 3.0 - 1e-6).all()


def test_autoreset_equals_step_plus_reset_done_and_graph_replay(mods):
    TaskEngine, make_task, _ = mods
    n = 256
    a_env, b_env = (make_task("Push_Car", env_num=n, seed=7, max_ep_len=15) for _ in range(2))
    oa, ob = a_env.reset(), b_env.reset()
    assert torch.equal(oa, ob)
    gen = torch.Generator(device='cuda').manual_seed(0)
    dones = 0
    for t in range(60):
        act = torch.rand(n, 2, device='cuda', generator=gen) * 2 - 1
        o1, r1, d1, i1 = a_env.step_autoreset(act)
        o2, r2, d2, i2 = b_env.step(act)
        o2 = b_env.reset_done()
        for x, y in ((o1, o2), (r1, r2), (d1, d2), (i1['cost'], i2['cost'])):
            assert torch.equal(x.view(torch.int32), y.view(torch.int32)), t
        dones += int(d1.sum())
    assert dones >= 3 * n
    assert torch.equal(a_env.get_state().view(torch.int32), b_env.get_state().view(torch.int32))
    # the fused step is graph-capturable (every counter lives on the device)
    act = torch.rand(n, 2, device='cuda', generator=gen) * 2 - 1
    obs, rdc = torch.empty(n, a_env.obs_flat_size, device='cuda'), torch.empty(3, n, device='cuda')
    p = lambda x: C.c_void_p(x.data_ptr())
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(graph, stream=s):
            rc = a_env._lib.gxt_step(a_env._h, p(act), p(obs), p(rdc[0]), p(rdc[1]), p(rdc[2]), None, 1, C.c_void_p(s.cuda_stream))
            assert rc == 0
    torch.cuda.current_stream().wait_stream(s)
    for _ in range(5):
        graph.replay()
        o2, r2, d2, i2 = b_env.step_autoreset(act)
        torch.cuda.synchronize()
        assert torch.equal(obs.view(torch.int32), o2.view(torch.int32)) and torch.equal(rdc[0], r2)


def test_car_drives_pushes_the_box_and_is_stopped_by_a_pillar(mods):
    TaskEngine, make_task, _ = mods
    n = 4
    g = make_task("Push_Car", env_num=n, seed=0, hazards_num=0, vases_num=0, gremlins_num=0, pillars_num=1, max_ep_len=1000)
    assert g.layout_dim == 2 * 3 + 2
    # goal far away, box 0.6 m ahead of the car (+x), pillar 1.5 m ahead
    lay = np.tile(np.array([[-1.5, 1.5, 0.6, 0.0, 1.5, 0.0, 0.0, 0.0]], np.float32), (n, 1))
    g.reset(lay)
    fwd = torch.ones(n, 2, device='cuda')
    xs, box = [], []
    for t in range(400):
        g.step(fwd)
        st = cpu(g.get_state())
        xs.append(st[0, 0]); box.append(st[0, 16 + 2])
    xs, box = np.array(xs), np.array(box)
    assert xs[50] > 0.2 and abs(cpu(g.get_state())[0, 1]) < 1e-5          # equal wheel commands: straight along +x
    assert box.max() > 0.7                                                 # the box was pushed ahead of the car
    assert (box - xs)[box > 0.61].min() >= 0.4 - 1e-5                      # ... at contact distance (robot 0.2 + box 0.2)
    assert xs.max() <= 1.5 - 0.4 + 1e-5                                    # the pillar (0.2) stops the car (0.2)
    # opposite wheel commands: the car turns on the spot
    g.reset(lay)
    spin = torch.tensor([[-1.0, 1.0]], device='cuda').repeat(n, 1)
    for t in range(100):
        g.step(spin)
    st = cpu(g.get_state())
    assert abs(st[0, 0]) < 1e-4 and abs(st[0, 1]) < 1e-4 and st[0, 2] > 0.5


def test_config_surface(mods):
    TaskEngine, make_task, _ = mods
    with pytest.raises(AssertionError, match="Bad key"):
        TaskEngine({'ghosts_num': 3})
    from guardx_b200.engine import ResamplingError
    with pytest.raises(ResamplingError):          # 30 objects with these keep-outs do not fit into 4 x 4 m
        make_task("Button_Car", env_num=4, hazards_num=8, vases_num=8, pillars_num=4, gremlins_num=6).reset()
    g = make_task("Button_Car", env_num=8)
    assert g.observation_space.shape == (g.obs_flat_size,) and g.action_space.shape == (2,)
    names = [k for k, _ in sorted(g._obs_slices.items(), key=lambda kv: kv[1][0])]
    assert names == sorted(names)            # flat observation in sorted key order (engine.py:773-777)
    assert 'box_lidar' not in names and 'buttons_lidar' in names and 'gremlins_lidar' in names
    import pickle
    g2 = pickle.loads(pickle.dumps(g))
    assert g2.obs_flat_size == g.obs_flat_size
