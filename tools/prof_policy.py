"""profiling target: the fused PPO rollout, T steps: the ONE persistent launch (gx_policy_rollout) or, with a third argument
"graph", the 2T launches (gx_policy_step + gx_step_autoreset).  python tools/prof_policy.py [envs] [T] [graph]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200 import make, ppo_lag
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = make("Goal_Point_8Hazards", env_num=n, seed=0, max_ep_len=200)
ac = ppo_lag.ActorCritic(43, 2).cuda()
env.reset(); env.stats_reset()
fr = ppo_lag.FusedRollout(env, ac, T, persistent=not (len(sys.argv) > 3 and sys.argv[3] == "graph"))
fr.reset_into_buffer()
fr.run(); fr.run()
torch.cuda.synchronize()
print("ok")
