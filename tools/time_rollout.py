"""time the fused PPO rollout (policy kernel + env kernel per step, one CUDA graph): python tools/time_rollout.py [envs] [T]"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from statistics import median
from guardx_b200 import make, ppo_lag
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 200
env = make("Goal_Point_8Hazards", env_num=n, seed=0, max_ep_len=T)
ac = ppo_lag.ActorCritic(43, 2).cuda()
env.reset(); env.stats_reset()
for persistent in (True, False):
    if persistent and not env._lib.gx_policy_rollout_supported(env._h):
        continue
    roll = ppo_lag.RolloutTimer(env, ac, T, persistent=persistent)
    ts = [roll.run() for _ in range(9)]
    dt = median(ts)
    print(json.dumps({"route": "persistent" if persistent else "graph", "envs": n, "T": T, "us_per_step": 1e6 * dt / T,
                      "env_steps_per_s": n * T / dt, "min_us": 1e6 * min(ts) / T}))
