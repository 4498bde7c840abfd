"""PPO-Lagrangian on the CUDA env -- the caller side of the hot path (BASELINE.json
configs[4]).

The reference has no PPO-Lag: it has PPO-clip (safe_rl_libX/ppo/ppo_train.py:142-452) and
the Lagrangian multiplier update of TRPO-Lag (trpolag/trpolag.py:466-471,
``lmd <- max(0, lmd + lam_lr * (EpCost - target_cost))``) with a cost critic.  This module
is PPO-clip with that multiplier and cost critic, keeping ``ppo_train.py``'s structure
(same actor-critic: ppo_core.py:91-160, Gaussian policy with state-independent log-std
-0.5, tanh MLPs 64-64; same losses :283-305, early stop on KL :327-331, Adam, advantage
normalisation :121-130) while removing what made the reference's loop host-bound:

  reference rollout step (ppo_train.py:365-417)          here
  ------------------------------------------------       -----------------------------------
  env.step + env.reset_done: two jitted calls + a         env.step_autoreset: ONE kernel launch
    full extra physics step
  five .cpu().numpy() syncs per step for the              none: rewards/costs/dones stay on the
    episode bookkeeping                                    device, episode statistics come from the
                                                           env's in-kernel accumulators (gx_stats_*)
  finish_path: per-env Python loop over                   rollout_ops.gae: one streaming kernel per
    scipy.signal.lfilter with device<->host copies         epoch for rewards and one for costs
  mpi_avg_grads via numpy on the host                     torch.distributed all-reduce (NCCL) of the
                                                           flattened gradients

The reference bootstraps every closed path and the epoch end with 0 (ppo_train.py:388-409:
``v[done == 1] = 0`` and ``done = ones`` at the timeout); so does this.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field

import torch
import torch.nn as nn
from torch.distributions.normal import Normal

from . import distributed as gxd
from . import rollout_ops


def mlp(sizes, activation=nn.Tanh, output_activation=nn.Identity):
    layers = []
    for j in range(len(sizes) - 1):
        layers += [nn.Linear(sizes[j], sizes[j + 1]), (activation if j < len(sizes) - 2 else output_activation)()]
    return nn.Sequential(*layers)


class ActorCritic(nn.Module):
    """ppo_core.MLPActorCritic (ppo_core.py:116-160) + a cost critic (TRPO-Lag's ``vc``)."""

    def __init__(self, obs_dim, act_dim, hidden_sizes=(64, 64)):
        super().__init__()
        self.log_std = nn.Parameter(torch.full((act_dim,), -0.5))
        self.mu_net = mlp([obs_dim, *hidden_sizes, act_dim])
        self.v = mlp([obs_dim, *hidden_sizes, 1])
        self.vc = mlp([obs_dim, *hidden_sizes, 1])

    def dist(self, obs):
        return Normal(self.mu_net(obs), torch.exp(self.log_std), validate_args=False)   # validation syncs the host

    @torch.no_grad()
    def step(self, obs):
        # pi.sample() / torch.normal(mean, std) validate std >= 0 with a host sync, which stalls the
        # eager loop and is illegal under CUDA-graph capture: sample and score by hand
        mu, std = self.mu_net(obs), torch.exp(self.log_std)
        a = mu + std * torch.randn_like(mu)
        logp = (-((a - mu) ** 2) / (2 * std * std) - self.log_std - 0.9189385332046727).sum(-1)
        return a, self.v(obs).squeeze(-1), self.vc(obs).squeeze(-1), logp

    def pi_params(self):
        return [self.log_std, *self.mu_net.parameters()]


@dataclass
class PPOLagConfig:
    epochs: int = 10
    max_ep_len: int = 200          # steps per epoch per env (ppo_train.py:463-464)
    gamma: float = 0.99
    lam: float = 0.97
    clip_ratio: float = 0.2
    pi_lr: float = 3e-4
    vf_lr: float = 1e-3
    train_pi_iters: int = 80
    train_v_iters: int = 80
    target_kl: float = 0.01
    target_cost: float = 0.0       # trpolag.py: target_cost
    lam_lr: float = 0.05           # trpolag.py: lam_lr (multiplier step)
    lmd_init: float = 0.0
    hidden_sizes: tuple = (64, 64)
    seed: int = 0
    use_fused: bool = True         # rollout step = gx_policy_step + gx_step_autoreset (two launches), the epoch one CUDA graph
    use_graph: bool = True         # (when not fused) capture the torch policy + env step + buffer writes as one CUDA graph
    log: list = field(default_factory=list)


def _avg_grads(params):
    """mpi_avg_grads (mpi_pytorch.py:20-27) on the device: one all-reduce of the flat gradient."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat)
    flat /= dist.get_world_size()
    o = 0
    for p in params:
        n = p.numel()
        p.grad.copy_(flat[o:o + n].view_as(p.grad))
        o += n


def _global_mean_std(x):
    """mpi_statistics_scalar (mpi_tools.py:70-92) without leaving the device."""
    import torch.distributed as dist
    s = torch.stack([x.sum(dtype=torch.float64), (x.double() ** 2).sum(), torch.tensor(float(x.numel()), device=x.device, dtype=torch.float64)])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(s)
    mean = s[0] / s[2]
    std = torch.sqrt(torch.clamp(s[1] / s[2] - mean * mean, min=0.0))
    return mean.float(), std.float()


def ppo_update(ac: ActorCritic, pi_opt, v_opt, vc_opt, data: dict, cfg: PPOLagConfig, lmd: float) -> dict:
    """compute_loss_pi / compute_loss_v / update() of ppo_train.py:283-348 with the Lagrangian
    advantage  (adv - lmd * adv_c) / (1 + lmd).  ``data`` tensors are flat [T*N, ...]."""
    obs, act, logp_old = data['obs'], data['act'], data['logp']
    adv = (data['adv'] - lmd * data['adv_c']) / (1.0 + lmd)

    def loss_pi():
        pi = ac.dist(obs)
        logp = pi.log_prob(act).sum(-1)
        ratio = torch.exp(logp - logp_old)
        clip_adv = torch.clamp(ratio, 1 - cfg.clip_ratio, 1 + cfg.clip_ratio) * adv
        return -(torch.min(ratio * adv, clip_adv)).mean(), (logp_old - logp).mean()

    pi_params = ac.pi_params()
    stop_iter = cfg.train_pi_iters
    kl = torch.zeros((), device=obs.device)
    for i in range(cfg.train_pi_iters):
        pi_opt.zero_grad(set_to_none=False)
        lp, kl = loss_pi()
        if float(gxd.global_mean(kl)) > cfg.target_kl:     # the one host sync per policy iteration (as mpi_avg(kl))
            stop_iter = i
            break
        lp.backward()
        _avg_grads(pi_params)
        pi_opt.step()
    for opt, net, key in ((v_opt, ac.v, 'ret'), (vc_opt, ac.vc, 'ret_c')):
        params = list(net.parameters())
        for _ in range(cfg.train_v_iters):
            opt.zero_grad(set_to_none=False)
            loss = ((net(obs).squeeze(-1) - data[key]) ** 2).mean()
            loss.backward()
            _avg_grads(params)
            opt.step()
    return dict(StopIter=stop_iter, KL=float(kl.detach()))


class GraphedRolloutStep:
    """One rollout step -- policy forward + sample, fused env step + auto-reset, the eight
    buffer writes -- captured once as a CUDA graph.  Legal because the env keeps every counter
    in device memory and the write index ``t`` is a device tensor advanced
    inside the graph.  A 4096-env step is ~30 small kernels: replaying them as a graph removes
    the ~0.5 ms of Python / launch overhead per step that otherwise bounds the rollout."""

    def __init__(self, env, ac, buf):
        import ctypes as C
        dev = env.device
        self.o = torch.empty(env.env_num, env.observation_space.shape[0], device=dev)
        self.t = torch.zeros(1, dtype=torch.long, device=dev)
        o_next = torch.empty_like(self.o)
        rdc = torch.empty(3, env.env_num, device=dev)
        p = lambda x: C.c_void_p(x.data_ptr())

        def body(stream_ptr):
            a, v, vc, logp = ac.step(self.o)
            for k, x in (('obs', self.o), ('act', a), ('val', v), ('cval', vc), ('logp', logp)):
                buf[k].index_copy_(0, self.t, x.unsqueeze(0))
            rc = env._lib.gx_step_autoreset(env._h, p(a), p(o_next), p(rdc[0]), p(rdc[1]), p(rdc[2]), stream_ptr)
            assert rc == 0, env._lib.gx_last_error()
            for k, j in (('rew', 0), ('done', 1), ('cost', 2)):
                buf[k].index_copy_(0, self.t, rdc[j].unsqueeze(0))
            self.o.copy_(o_next)
            self.t += 1

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            # warm-up on a copy of the env state would change the trajectory; the warm-up step is simply part
            # of the epoch's first reset..step sequence being overwritten: callers reset() before every epoch
            sp = C.c_void_p(side.cuda_stream)
            body(sp)
            torch.cuda.synchronize(dev)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                body(sp)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

    def run(self, o0, T):
        self.o.copy_(o0)
        self.t.zero_()
        for _ in range(T):
            self.graph.replay()


def pack_policy(ac: ActorCritic, obs_dim: int, act_dim: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """The actor-critic's parameters in the layout gx_policy_step reads (csrc/gx_policy.cuh):
    W1 [3][K1/4][64][4] | W2 [3][16][64][4] | B1 [3][64] | B2 [3][64] | W3 [A+2][64] | B3 [A+2] | log_std [A], nets ordered
    mu_net, v, vc; K1 = obs_dim rounded up to 4 (padding inputs are zero); W*[net][k/4][u][k&3] is the weight of hidden
    unit u for input k (the kernel copies a net's block into shared memory as it is).  Written in place into ``out``
    when given, so that a captured CUDA graph keeps reading the same buffer after every update."""
    nets = (ac.mu_net, ac.v, ac.vc)
    lin = [[m for m in net if isinstance(m, nn.Linear)] for net in nets]
    assert all(len(l) == 3 and l[0].out_features == 64 and l[1].out_features == 64 for l in lin), \
        'gx_policy_step is built for hidden_sizes == (64, 64) (the reference default, ppo_train.py:461)'
    k1 = (obs_dim + 3) // 4 * 4
    with torch.no_grad():
        def tiles(w, k):          # Linear weight [64, in] -> [k/4][64][4]
            wp = torch.zeros(64, k, device=w.device)
            wp[:, :w.shape[1]] = w
            return wp.reshape(64, k // 4, 4).permute(1, 0, 2).reshape(-1)
        parts = [torch.cat([tiles(l[0].weight, k1) for l in lin]), torch.cat([tiles(l[1].weight, 64) for l in lin]),
                 torch.cat([l[0].bias for l in lin]), torch.cat([l[1].bias for l in lin]),
                 torch.cat([l[2].weight for l in lin]).reshape(-1),             # rows: A actor heads, v, vc
                 torch.cat([l[2].bias for l in lin]), ac.log_std]
        flat = torch.cat([x.float().reshape(-1) for x in parts])
        if out is None:
            return flat.contiguous()
        out.copy_(flat)
        return out


class FusedRollout:
    """One epoch's rollout (ppo_train.py:364-417) without torch ops, copy kernels or host work between the T steps;
    every output is written in place into the rollout buffer (``buf['obs']`` has T + 1 rows).  Two routes:

    * ``persistent`` (default where gx_policy_rollout_supported: the default Point shape): ONE launch of
      gx_policy_rollout -- each block owns 32 envs, keeps their state in registers and the three networks in shared
      memory, and alternates policy and env phases T times;
    * otherwise 2 T launches in ONE CUDA graph: per control step the policy kernel (gx_policy_step: three MLPs +
      Gaussian sample + log-prob, reading the observation row the env wrote) and the env kernel (gx_step_autoreset,
      reading that action, writing the next observation into the buffer's next row).

    Both produce the same bits (tests/test_ppo_lag.py)."""

    HOW_GRAPH = ("gx_policy_step (3 fp32 MLPs + Philox Gaussian sample + log-prob, one launch) + gx_step_autoreset per step, "
                 "outputs written in place into the rollout buffer; 2T launches replayed as one CUDA graph")
    HOW_PERSISTENT = ("gx_policy_rollout: ONE persistent launch for the T steps -- a block owns 32 envs (state in registers, the "
                      "3 fp32 MLPs' weights in shared memory) and alternates policy phases (register-tiled FFMA2, Philox Gaussian "
                      "sample + log-prob) and env phases (gx_step_autoreset's device function), four groups of warps out of phase; "
                      "outputs in place in the rollout buffer")
    how = HOW_GRAPH

    def __init__(self, env, ac, T, noise_seed=0, persistent=None):
        import ctypes as C
        N, dev = env.env_num, env.device
        od, ad = env.observation_space.shape[0], env.action_space.shape[0]
        f = lambda *s: torch.empty(*s, device=dev)
        self.env, self.ac, self.T, self.od, self.ad = env, ac, T, od, ad
        self.buf = dict(obs=f(T + 1, N, od), act=f(T, N, ad), rew=f(T, N), cost=f(T, N), done=f(T, N), val=f(T, N),
                        cval=f(T, N), logp=f(T, N))
        L, h = env._lib, env._h
        assert int(L.gx_policy_weight_floats(h)) > 0
        self.weights = torch.empty(int(L.gx_policy_weight_floats(h)), device=dev)
        self.refresh_weights()
        b = self.buf
        p = lambda x: C.c_void_p(x.data_ptr())
        supported = bool(L.gx_policy_rollout_supported(h))
        self.persistent = supported if persistent is None else bool(persistent)
        assert supported or not self.persistent, 'gx_policy_rollout handles the default Point shape only'
        self.how = self.HOW_PERSISTENT if self.persistent else self.HOW_GRAPH
        if self.persistent:
            self.graph = None
            self._args = (h, p(self.weights), C.c_int32(T), C.c_uint64(noise_seed), p(b['obs']), p(b['act']), p(b['rew']), p(b['done']),
                          p(b['cost']), p(b['val']), p(b['cval']), p(b['logp']))
            self.reset_into_buffer()
            return

        def issue(sp):
            for t in range(T):
                rc = L.gx_policy_step(h, p(self.weights), p(b['obs'][t]), C.c_uint64(noise_seed), p(b['act'][t]), p(b['val'][t]),
                                      p(b['cval'][t]), p(b['logp'][t]), None, sp)
                assert rc == 0, L.gx_last_error()
                rc = L.gx_step_autoreset(h, p(b['act'][t]), p(b['obs'][t + 1]), p(b['rew'][t]), p(b['done'][t]), p(b['cost'][t]), sp)
                assert rc == 0, L.gx_last_error()

        self.reset_into_buffer()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            sp = C.c_void_p(side.cuda_stream)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                issue(sp)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)

    def refresh_weights(self):
        pack_policy(self.ac, self.od, self.ad, out=self.weights)

    def reset_into_buffer(self):
        """env.reset() (ppo_train.py:401) with the observation written into row 0 of the rollout buffer"""
        import ctypes as C
        from . import _abi
        env = self.env
        _abi.check(env._lib, env._lib.gx_reset(env._h, C.c_void_p(self.buf['obs'][0].data_ptr()), env._stream()), "gx_reset")
        env._obs = self.buf['obs'][0]

    def run(self):
        if self.persistent:
            from . import _abi
            L = self.env._lib
            _abi.check(L, L.gx_policy_rollout(*self._args, self.env._stream()), "gx_policy_rollout")
        else:
            self.graph.replay()


class RolloutTimer:
    """Times ONE epoch's rollout (T steps: policy forward + sample, env step + auto-reset, the
    buffer writes) with CUDA events on the current stream; used by bench.py."""

    def __init__(self, env, ac, T, fused=True, persistent=None):
        N, dev = env.env_num, env.device
        self.T, self.env, self.fused = T, env, fused
        env.reset()
        env.stats_reset()
        if fused:
            self.roll = FusedRollout(env, ac, T, persistent=persistent)
            self.how = self.roll.how
        else:
            f = lambda *s: torch.empty(*s, device=dev)
            self.buf = dict(obs=f(T, N, env.observation_space.shape[0]), act=f(T, N, env.action_space.shape[0]), rew=f(T, N),
                            cost=f(T, N), done=f(T, N), val=f(T, N), cval=f(T, N), logp=f(T, N))
            self.step = GraphedRolloutStep(env, ac, self.buf)
            self.how = "torch policy forward + sample + gx_step_autoreset + 8 buffer writes per step (~30 kernels), one CUDA graph replayed T times"
        self.run()      # warm-up
        torch.cuda.synchronize(dev)

    def run(self) -> float:
        if self.fused:
            self.roll.reset_into_buffer()
        else:
            o = self.env.reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(self.env.device)
        e0.record()
        if self.fused:
            self.roll.run()
        else:
            self.step.run(o, self.T)
        e1.record()
        torch.cuda.synchronize(self.env.device)
        return e0.elapsed_time(e1) * 1e-3


def ppo_lag(env, cfg: PPOLagConfig | None = None, ac: ActorCritic | None = None):
    """Train on ``env`` (a guardx_b200.Engine).  Returns (actor_critic, per-epoch log)."""
    cfg = cfg or PPOLagConfig()
    dev = env.device
    torch.manual_seed(cfg.seed)
    obs_dim, act_dim = env.observation_space.shape[0], env.action_space.shape[0]
    ac = ac or ActorCritic(obs_dim, act_dim, cfg.hidden_sizes).to(dev)
    gxd.sync_params(ac)
    # ppo_train.py:257: seed += 10000 * proc_id() -- AFTER the parameters are shared, so that every rank
    # explores with its own noise stream on identical weights
    torch.manual_seed(cfg.seed + 10000 * gxd.rank())
    pi_opt = torch.optim.Adam(ac.pi_params(), lr=cfg.pi_lr)
    v_opt = torch.optim.Adam(ac.v.parameters(), lr=cfg.vf_lr)
    vc_opt = torch.optim.Adam(ac.vc.parameters(), lr=cfg.vf_lr)
    N, T = env.env_num, cfg.max_ep_len
    f = lambda *s: torch.empty(*s, device=dev)
    lmd = cfg.lmd_init
    graphed = fused = None
    use_fused = cfg.use_fused and tuple(cfg.hidden_sizes) == (64, 64)
    if not use_fused:
        buf = dict(obs=f(T, N, obs_dim), act=f(T, N, act_dim), rew=f(T, N), cost=f(T, N), done=f(T, N), val=f(T, N),
                   cval=f(T, N), logp=f(T, N))
    if use_fused:
        env.reset()
        env.stats_reset()     # statistics on BEFORE capture: the captured launch is the accumulating variant
        fused = FusedRollout(env, ac, T, noise_seed=cfg.seed + 10000 * gxd.rank())
        buf = dict(fused.buf, obs=fused.buf['obs'][:T])
    elif cfg.use_graph:
        env.reset()
        env.stats_reset()
        graphed = GraphedRolloutStep(env, ac, buf)
    t_start = time.time()
    for epoch in range(cfg.epochs):
        if fused is not None:
            fused.refresh_weights()                       # the parameters Adam just updated, in the kernel's layout
            fused.reset_into_buffer()                     # ppo_train.py:401 resets at every epoch end
            env.stats_reset()
            fused.run()
        else:
            o = env.reset()
            env.stats_reset()
        if graphed is not None:
            graphed.run(o, T)
        for t in range(0 if (graphed is None and fused is None) else T, T):
            a, v, vc, logp = ac.step(o)
            buf['obs'][t], buf['act'][t], buf['val'][t], buf['cval'][t], buf['logp'][t] = o, a, v, vc, logp
            o, r, d, info = env.step_autoreset(a)         # step + reset_done in one launch, nothing leaves the device
            buf['rew'][t], buf['cost'][t], buf['done'][t] = r, info['cost'], d
        # finish_path for every env at once; every closed path and the epoch end bootstrap with 0
        adv, ret = rollout_ops.gae(buf['rew'], buf['val'], buf['done'], None, cfg.gamma, cfg.lam)
        adv_c, ret_c = rollout_ops.gae(buf['cost'], buf['cval'], buf['done'], None, cfg.gamma, cfg.lam)
        m, s = _global_mean_std(adv)
        adv = (adv - m) / s
        m_c, _ = _global_mean_std(adv_c)
        adv_c = adv_c - m_c                                # cost advantage is centred, not rescaled (as in TRPO-Lag)
        flat = lambda x: x.reshape(T * N, *x.shape[2:])
        data = dict(obs=flat(buf['obs']), act=flat(buf['act']), logp=flat(buf['logp']), adv=flat(adv), adv_c=flat(adv_c),
                    ret=flat(ret), ret_c=flat(ret_c))
        # episode statistics: the env's device accumulators, all-reduced over ranks (replaces the
        # host bookkeeping of ppo_train.py:370-376 and mpi_statistics_scalar)
        stats = env.stats()
        gxd.allreduce_stats(stats)
        summ = gxd.episode_summary(stats.tolist())
        # episodes cut by the epoch end count with what they accumulated (ppo_train.py:392-396 logs them too)
        # (no finished episode anywhere: the mean accumulated return / cost per env over ALL ranks, so that every rank
        # applies the same multiplier update)
        if summ['Episodes'] == 0:
            # ppo_train.py:392-409 logs the paths the epoch end cuts off as episodes: per-env sums over the epoch
            summ['EpRet'] = float(gxd.global_mean(buf['rew'].sum(0).mean()))
            summ['EpCost'] = float(gxd.global_mean(buf['cost'].sum(0).mean()))
            summ['EpLen'] = float(T)
        ep_cost = summ['EpCost']
        info = ppo_update(ac, pi_opt, v_opt, vc_opt, data, cfg, lmd)
        lmd = max(0.0, lmd + cfg.lam_lr * (ep_cost - cfg.target_cost))     # trpolag.py:466-471
        row = dict(Epoch=epoch, Lambda=lmd, Time=time.time() - t_start, EpCostUsed=ep_cost,
                   MeanStepReward=float(buf['rew'].mean()), MeanStepCost=float(buf['cost'].mean()), **summ, **info)
        cfg.log.append(row)
    return ac, cfg.log
