"""BASELINE.json configs[4]: PPO-Lagrangian rollout + update on Goal_Point_8Hazards with the env
on the CUDA kernels.  One process per GPU (torchrun for N > 1); prints one JSON line on rank 0:
env-steps/s of rollout-only and of rollout+update, per epoch after a warm-up epoch.
    python tools/bench_ppo.py [--task T] [--envs N] [--len L] [--epochs E]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from guardx_b200 import make, ppo_lag  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--task", default="Goal_Point_8Hazards")
ap.add_argument("--envs", type=int, default=4096)
ap.add_argument("--len", type=int, default=200)
ap.add_argument("--epochs", type=int, default=4)
args = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
env = make(args.task, env_num=args.envs, seed=rank, max_ep_len=args.len)
cfg = ppo_lag.PPOLagConfig(epochs=args.epochs, max_ep_len=args.len, target_cost=0.0)
t0 = time.time()
ac, log = ppo_lag.ppo_lag(env, cfg)
torch.cuda.synchronize()
# rollout-only rates: the eager loop, and the CUDA-graph step replayed T times
o = env.reset()
torch.cuda.synchronize(); t1 = time.time()
for t in range(args.len):
    a, v, vc, lp = ac.step(o)
    o, r, d, info = env.step_autoreset(a)
torch.cuda.synchronize(); t_eager = time.time() - t1
T, N = args.len, args.envs
f = lambda *s: torch.empty(*s, device=env.device)
buf = dict(obs=f(T, N, env.observation_space.shape[0]), act=f(T, N, env.action_space.shape[0]), rew=f(T, N), cost=f(T, N),
           done=f(T, N), val=f(T, N), cval=f(T, N), logp=f(T, N))
g = ppo_lag.GraphedRolloutStep(env, ac, buf)
o = env.reset(); g.run(o, T)
torch.cuda.synchronize(); t1 = time.time()
g.run(o, T)
torch.cuda.synchronize(); t_roll = time.time() - t1
if rank == 0:
    per_epoch = (log[-1]['Time'] - log[0]['Time']) / max(1, args.epochs - 1)
    steps = args.envs * args.len * world
    print(json.dumps({"task": args.task, "n_gpus": world, "envs_per_gpu": args.envs, "steps_per_epoch_per_env": args.len,
                      "rollout_env_steps_per_s": steps / t_roll, "rollout_eager_env_steps_per_s": steps / t_eager, "train_env_steps_per_s": steps / per_epoch,
                      "epoch_s": per_epoch, "last": {k: log[-1][k] for k in ("EpRet", "EpCost", "Lambda", "KL", "StopIter", "MeanStepCost")}}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
