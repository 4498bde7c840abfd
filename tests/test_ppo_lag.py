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
        cfg = ppo_lag.PPOLagConfig(epochs=2, max_ep_len=25, train_pi_iters=3, train_v_iters=3, use_graph=use_graph, seed=5)
        _, log = ppo_lag.ppo_lag(env, cfg)
        logs.append(log)
    for a, b in zip(*logs):
        assert a['TotalEnvInteracts'] == b['TotalEnvInteracts'] == 300 * 25
        assert np.isfinite([b['MeanStepReward'], b['MeanStepCost']]).all()
        # the graph warm-up consumes extra policy samples, so trajectories differ; the statistics must be sane and close
        assert abs(a['MeanStepCost'] - b['MeanStepCost']) < 0.05
