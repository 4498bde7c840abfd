"""PPO-Lagrangian harness (guardx_b200/ppo_lag.py): the update math on synthetic buffers
(CPU), gradient averaging / parameter sync over two gloo ranks (CPU), and a short training
run on the CUDA env (GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from guardx_b200 import ppo_lag


def _data(n=512, obs_dim=7, act_dim=2, seed=0):
    g = torch.Generator().manual_seed(seed)
    obs = torch.randn(n, obs_dim, generator=g)
    act = torch.randn(n, act_dim, generator=g)
    return obs, act


def test_update_improves_surrogate_and_fits_critics():
    torch.manual_seed(0)
    obs, act = _data()
    ac = ppo_lag.ActorCritic(7, 2)
    with torch.no_grad():
        logp = ac.dist(obs).log_prob(act).sum(-1)
    adv = torch.sign(act[:, 0])                       # "actions with a0 > 0 were good"
    adv_c = torch.sign(act[:, 1])                     # "... and a1 > 0 was costly"
    ret, ret_c = obs[:, 0] * 2.0, obs[:, 1] - 1.0
    data = dict(obs=obs, act=act, logp=logp, adv=adv, adv_c=adv_c, ret=ret, ret_c=ret_c)
    cfg = ppo_lag.PPOLagConfig(train_pi_iters=40, train_v_iters=200, target_kl=0.05)
    opts = [torch.optim.Adam(ac.pi_params(), lr=3e-3), torch.optim.Adam(ac.v.parameters(), lr=1e-2),
            torch.optim.Adam(ac.vc.parameters(), lr=1e-2)]
    mu0 = ac.mu_net(obs).detach().mean(0)
    v_err0 = float(((ac.v(obs).squeeze(-1) - ret) ** 2).mean())
    info = ppo_lag.ppo_update(ac, *opts, data, cfg, lmd=1.0)
    mu1 = ac.mu_net(obs).detach().mean(0)
    assert mu1[0] > mu0[0] and mu1[1] < mu0[1]        # reward pushes a0 up, the multiplier pushes a1 down
    assert float(((ac.v(obs).squeeze(-1) - ret) ** 2).mean()) < 0.2 * v_err0
    assert float(((ac.vc(obs).squeeze(-1) - ret_c) ** 2).mean()) < 0.5
    assert 0 < info['StopIter'] <= 40 and info['KL'] <= 0.05 * 1.5 + 1e-3
    # lmd = 0 ignores the cost advantage altogether
    ac2 = ppo_lag.ActorCritic(7, 2)
    with torch.no_grad():
        data['logp'] = ac2.dist(obs).log_prob(act).sum(-1)
    m0 = ac2.mu_net(obs).detach().mean(0)
    ppo_lag.ppo_update(ac2, torch.optim.Adam(ac2.pi_params(), lr=3e-3), torch.optim.Adam(ac2.v.parameters()),
                       torch.optim.Adam(ac2.vc.parameters()), data, ppo_lag.PPOLagConfig(train_pi_iters=20, train_v_iters=1, target_kl=0.05), lmd=0.0)
    m1 = ac2.mu_net(obs).detach().mean(0)
    assert m1[0] > m0[0] and abs(m1[1] - m0[1]) < 0.5 * abs(m1[0] - m0[0])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from guardx_b200 import distributed as gxd
    torch.manual_seed(100 + rank)                      # different initial weights and data per rank
    ac = ppo_lag.ActorCritic(7, 2)
    gxd.sync_params(ac)
    obs, act = _data(seed=rank)
    with torch.no_grad():
        logp = ac.dist(obs).log_prob(act).sum(-1)
    data = dict(obs=obs, act=act, logp=logp, adv=torch.sign(act[:, 0]), adv_c=torch.zeros(len(obs)), ret=obs[:, 0], ret_c=obs[:, 1])
    cfg = ppo_lag.PPOLagConfig(train_pi_iters=5, train_v_iters=5, target_kl=1.0)
    ppo_lag.ppo_update(ac, torch.optim.Adam(ac.pi_params(), lr=1e-2), torch.optim.Adam(ac.v.parameters(), lr=1e-2),
                       torch.optim.Adam(ac.vc.parameters(), lr=1e-2), data, cfg, lmd=0.5)
    m, s = ppo_lag._global_mean_std(torch.arange(4.0) + 4 * rank)
    flat = torch.cat([p.detach().reshape(-1) for p in ac.parameters()])
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), flat=flat.numpy(), m=float(m), s=float(s))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_stay_in_lockstep(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = (np.load(os.path.join(tmp_path, f"r{r}.npz")) for r in range(2))
    np.testing.assert_array_equal(a['flat'], b['flat'])          # synced start + averaged grads => identical weights
    assert a['m'] == pytest.approx(3.5) and a['s'] == pytest.approx(np.std(np.arange(8.0)))


@pytest.mark.gpu
@pytest.mark.parametrize("task", ["Goal_Point_8Hazards", "Goal_Ant_8Hazards"])
def test_short_training_run_on_the_cuda_env(task):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from guardx_b200 import make
    env = make(task, env_num=256, seed=0, max_ep_len=60)
    cfg = ppo_lag.PPOLagConfig(epochs=3, max_ep_len=60, train_pi_iters=10, train_v_iters=10, target_cost=0.0, lam_lr=0.5)
    ac, log = ppo_lag.ppo_lag(env, cfg)
    assert len(log) == 3
    for row in log:
        assert np.isfinite([row['MeanStepReward'], row['MeanStepCost'], row['Lambda'], row['KL']]).all()
        assert row['TotalEnvInteracts'] == 256 * 60
    assert log[-1]['Lambda'] >= 0.0
    assert all(torch.isfinite(p).all() for p in ac.parameters())


@pytest.mark.gpu
def test_graphed_rollout_agrees_with_eager_rollout():
    """The CUDA-graph rollout (policy + env step + buffer writes replayed T times) trains like the
    step-by-step loop."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from guardx_b200 import make
    logs = []
    for use_graph in (False, True):
        env = make("Goal_Point_8Hazards", env_num=300, seed=3, max_ep_len=25)
        cfg = ppo_lag.PPOLagConfig(epochs=2, max_ep_len=25, train_pi_iters=3, train_v_iters=3, use_graph=use_graph, use_fused=False, seed=5)
        _, log = ppo_lag.ppo_lag(env, cfg)
        logs.append(log)
    for a, b in zip(*logs):
        assert a['TotalEnvInteracts'] == b['TotalEnvInteracts'] == 300 * 25
        assert np.isfinite([b['MeanStepReward'], b['MeanStepCost']]).all()
        # the graph warm-up consumes extra policy samples, so trajectories differ; the statistics must be sane and close
        assert abs(a['MeanStepCost'] - b['MeanStepCost']) < 0.05


def test_pack_policy_layout():
    """pack_policy writes the layout csrc/gx_policy.cuh documents (checked on the CPU with a numpy forward pass)."""
    torch.manual_seed(1)
    od, ad = 43, 2
    ac = ppo_lag.ActorCritic(od, ad)
    with torch.no_grad():
        ac.log_std.copy_(torch.tensor([-0.3, -0.7]))
    w = ppo_lag.pack_policy(ac, od, ad).numpy()
    k1 = 44
    o = 0
    W1 = w[o:o + k1 * 192].reshape(3, k1 // 4, 64, 4).transpose(1, 3, 0, 2).reshape(k1, 192); o += k1 * 192    # -> [k][net*64+u]
    W2 = w[o:o + 3 * 64 * 64].reshape(3, 16, 64, 4).transpose(0, 1, 3, 2).reshape(3, 64, 64); o += 3 * 64 * 64  # -> [net][k][u]
    B1 = w[o:o + 192]; o += 192
    B2 = w[o:o + 192]; o += 192
    W3 = w[o:o + (ad + 2) * 64].reshape(ad + 2, 64); o += (ad + 2) * 64
    B3 = w[o:o + ad + 2]; o += ad + 2
    LS = w[o:o + ad]; o += ad
    assert o == w.size and (W1[od:] == 0).all()
    x = torch.randn(5, od)
    xp = np.concatenate([x.numpy(), np.zeros((5, 1), np.float32)], 1)
    h1 = np.tanh(xp @ W1 + B1)
    h2 = np.tanh(np.concatenate([h1[:, n * 64:(n + 1) * 64] @ W2[n] for n in range(3)], 1) + B2)
    mu = h2[:, :64] @ W3[:ad].T + B3[:ad]
    v = h2[:, 64:128] @ W3[ad] + B3[ad]
    vc = h2[:, 128:] @ W3[ad + 1] + B3[ad + 1]
    with torch.no_grad():
        np.testing.assert_allclose(mu, ac.mu_net(x).numpy(), atol=2e-6)
        np.testing.assert_allclose(v, ac.v(x).squeeze(-1).numpy(), atol=2e-6)
        np.testing.assert_allclose(vc, ac.vc(x).squeeze(-1).numpy(), atol=2e-6)
    np.testing.assert_array_equal(LS, ac.log_std.detach().numpy())


@pytest.mark.gpu
@pytest.mark.parametrize("task,n", [("Goal_Point_8Hazards", 1000), ("Goal_Ant_8Hazards", 100), ("Goal_Doggo_8Hazards", 70)])
def test_policy_kernel_matches_the_torch_actor_critic(task, n):
    """gx_policy_step vs ActorCritic: mean / value / cost value within 1e-5 of the torch forward pass (fp32, different
    summation order), logp == Normal.log_prob of the stored action, noise ~ N(0, 1), replay-safe counters."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes as C
    from guardx_b200 import make
    env = make(task, env_num=n, seed=2, max_ep_len=50)
    od, ad = env.observation_space.shape[0], env.action_space.shape[0]
    torch.manual_seed(3)
    ac = ppo_lag.ActorCritic(od, ad).cuda()
    with torch.no_grad():
        for p_ in ac.parameters():
            p_.mul_(3.0)                                   # saturate some tanh units
        ac.log_std.copy_(torch.linspace(-1.0, 0.2, ad))
    w = ppo_lag.pack_policy(ac, od, ad)
    assert w.numel() == env._lib.gx_policy_weight_floats(env._h)
    obs = env.reset()
    f = lambda *s: torch.empty(*s, device='cuda')
    act, val, cval, logp, mu = f(n, ad), f(n), f(n), f(n), f(n, ad)
    p = lambda t: C.c_void_p(t.data_ptr())
    zs = []
    for t in range(6):
        rc = env._lib.gx_policy_step(env._h, p(w), p(obs), C.c_uint64(77), p(act), p(val), p(cval), p(logp), p(mu), env._stream())
        assert rc == 0, env._lib.gx_last_error()
        with torch.no_grad():
            mu_t, v_t, vc_t = ac.mu_net(obs), ac.v(obs).squeeze(-1), ac.vc(obs).squeeze(-1)
            std = torch.exp(ac.log_std)
            lp_t = torch.distributions.Normal(mu, std).log_prob(act).sum(-1)
        assert float((mu - mu_t).abs().max()) <= 1e-5 * (1 + float(mu_t.abs().max()))
        assert float((val - v_t).abs().max()) <= 1e-5 * (1 + float(v_t.abs().max()))
        assert float((cval - vc_t).abs().max()) <= 1e-5 * (1 + float(vc_t.abs().max()))
        assert float((logp - lp_t).abs().max()) <= 2e-5 * (1 + float(lp_t.abs().max()))
        zs.append(((act - mu) / std).flatten().cpu())
        act2 = act.clone()
        rc = env._lib.gx_policy_step(env._h, p(w), p(obs), C.c_uint64(77), p(act), p(val), p(cval), p(logp), None, env._stream())
        assert torch.equal(act, act2)                      # same counters -> same draw (deterministic, graph-replayable)
        obs, r, d, info = env.step_autoreset(act)          # the env's step count advances -> fresh noise next time
    z = torch.cat(zs)
    assert torch.isfinite(z).all()
    assert abs(float(z.mean())) < 5.0 / z.numel() ** 0.5 and abs(float(z.var()) - 1.0) < 8.0 / z.numel() ** 0.5
    assert float(z.abs().max()) > 2.5                      # tails are there
    # successive steps of one env and neighbouring envs are uncorrelated
    zz = torch.stack([x.view(n, ad)[:, 0] for x in zs])
    assert abs(float((zz[:-1] * zz[1:]).mean())) < 5.0 / (zz[:-1].numel()) ** 0.5
    assert abs(float((zz[:, :-1] * zz[:, 1:]).mean())) < 5.0 / (zz[:, 1:].numel()) ** 0.5


@pytest.mark.gpu
def test_fused_rollout_fills_the_buffer_like_the_step_by_step_loop():
    """FusedRollout -- both routes: the ONE persistent launch (gx_policy_rollout) and the 2T launches in one graph, outputs
    written in place -- against the same calls issued one by one with separate tensors: bit-identical buffers and episode
    statistics (same device functions, same counters)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes as C
    from guardx_b200 import make
    n, T = 500, 30
    bufs = []
    st_f = []
    for mode in ("persistent", "graph", "loop"):
        fused = mode != "loop"
        env = make("Goal_Point_8Hazards", env_num=n, seed=4, max_ep_len=12)      # timeouts inside the window
        torch.manual_seed(9)
        ac = ppo_lag.ActorCritic(43, 2).cuda()
        if fused:
            env.reset(); env.stats_reset()
            fr = ppo_lag.FusedRollout(env, ac, T, noise_seed=5, persistent=(mode == "persistent"))
            assert fr.persistent == (mode == "persistent")
            env.set_counters(0, 0)       # (capture ran no kernels; the reset above consumed pool epoch 0)
            fr.reset_into_buffer(); env.stats_reset()
            fr.run()
            torch.cuda.synchronize()
            bufs.append({k: v.clone() for k, v in fr.buf.items()})
            st_f.append(env.stats().clone())
        else:
            env.reset(); env.stats_reset()
            env.set_counters(0, 0)
            w = ppo_lag.pack_policy(ac, 43, 2)
            o = env.reset(); env.stats_reset()
            f = lambda *s: torch.empty(*s, device='cuda')
            b = dict(obs=f(T + 1, n, 43), act=f(T, n, 2), rew=f(T, n), cost=f(T, n), done=f(T, n), val=f(T, n), cval=f(T, n), logp=f(T, n))
            p = lambda t: C.c_void_p(t.data_ptr())
            for t in range(T):
                b['obs'][t] = o
                a, v, vc, lp = f(n, 2), f(n), f(n), f(n)
                assert env._lib.gx_policy_step(env._h, p(w), p(o), C.c_uint64(5), p(a), p(v), p(vc), p(lp), None, env._stream()) == 0
                o, r, d, info = env.step_autoreset(a)
                b['act'][t], b['val'][t], b['cval'][t], b['logp'][t], b['rew'][t], b['done'][t], b['cost'][t] = a, v, vc, lp, r, d, info['cost']
            b['obs'][T] = o
            bufs.append(b)
            st_e = env.stats()
    for which in (0, 1):
        for k in bufs[2]:
            assert torch.equal(bufs[which][k].view(torch.int32), bufs[2][k].view(torch.int32)), (("persistent", "graph")[which], k)
        assert torch.equal(st_f[which], st_e)
    assert float(bufs[2]['done'].sum()) >= n            # episodes really ended (and envs were re-placed) inside the launch


@pytest.mark.gpu
def test_policy_rollout_in_two_calls_equals_one_call():
    """gx_policy_rollout keeps the env state on the device between calls and takes row 0 of the observation buffer as its
    input: T = 24 steps in one launch == 10 + 14 steps in two launches, bit for bit (buffers and episode statistics)."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes as C
    from guardx_b200 import make
    n, T, T1 = 333, 24, 10
    res = []
    for split in (False, True):
        env = make("Goal_Point_8Hazards", env_num=n, seed=11, max_ep_len=9)
        torch.manual_seed(3)
        ac = ppo_lag.ActorCritic(43, 2).cuda()
        L, h = env._lib, env._h
        assert L.gx_policy_rollout_supported(h) == 1
        w = ppo_lag.pack_policy(ac, 43, 2)
        f = lambda *s: torch.empty(*s, device='cuda')
        b = dict(obs=f(T + 1, n, 43), act=f(T, n, 2), rew=f(T, n), done=f(T, n), cost=f(T, n), val=f(T, n), cval=f(T, n), logp=f(T, n))
        b['obs'][0] = env.reset(); env.stats_reset()
        p = lambda t: C.c_void_p(t.data_ptr())
        for t0, steps in (((0, T1), (T1, T - T1)) if split else ((0, T),)):
            rc = L.gx_policy_rollout(h, p(w), steps, C.c_uint64(7), p(b['obs'][t0:]), p(b['act'][t0:]), p(b['rew'][t0:]), p(b['done'][t0:]),
                                     p(b['cost'][t0:]), p(b['val'][t0:]), p(b['cval'][t0:]), p(b['logp'][t0:]), env._stream())
            assert rc == 0, L.gx_last_error()
        torch.cuda.synchronize()
        res.append((b, env.stats().clone()))
    for k in res[0][0]:
        assert torch.equal(res[0][0][k].view(torch.int32), res[1][0][k].view(torch.int32)), k
    assert torch.equal(res[0][1], res[1][1])
    assert float(res[0][0]['done'].sum()) >= n


@pytest.mark.gpu
def test_policy_rollout_is_refused_for_other_shapes_and_bad_arguments():
    """The persistent rollout kernel is built for the default Point shape only; everything else gets an error code and a
    message (and FusedRollout falls back to the two-launch graph), never a wrong answer."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import ctypes as C
    from guardx_b200 import make
    env = make("Goal_Ant_8Hazards", env_num=64, seed=0, max_ep_len=20)
    env.reset()
    L, h = env._lib, env._h
    assert L.gx_policy_rollout_supported(h) == 0
    x = torch.zeros(1 << 16, device='cuda')
    p = C.c_void_p(x.data_ptr())
    assert L.gx_policy_rollout(h, p, 4, C.c_uint64(0), p, p, p, p, p, p, p, p, env._stream()) != 0
    assert b'default Point shape' in L.gx_last_error()
    ac = ppo_lag.ActorCritic(env.observation_space.shape[0], env.action_space.shape[0]).cuda()
    fr = ppo_lag.FusedRollout(env, ac, 3)
    assert not fr.persistent and fr.graph is not None
    pt = make("Goal_Point_8Hazards", env_num=64, seed=0, max_ep_len=20)
    L, h = pt._lib, pt._h
    assert L.gx_policy_rollout(h, p, 4, C.c_uint64(0), p, p, p, p, p, p, p, p, pt._stream()) != 0       # before reset
    pt.reset()
    assert L.gx_policy_rollout(h, None, 4, C.c_uint64(0), p, p, p, p, p, p, p, p, pt._stream()) != 0    # NULL weights
    assert L.gx_policy_rollout(h, C.c_void_p(x.data_ptr() + 4), 4, C.c_uint64(0), p, p, p, p, p, p, p, p, pt._stream()) != 0   # misaligned weights
    assert L.gx_policy_rollout(h, p, -1, C.c_uint64(0), p, p, p, p, p, p, p, p, pt._stream()) != 0
    assert L.gx_policy_rollout(h, p, 0, C.c_uint64(0), p, p, p, p, p, p, p, p, pt._stream()) == 0       # T = 0: nothing to do
