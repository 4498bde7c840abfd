"""profiling target: the fused PPO rollout (gx_policy_step + gx_step_autoreset), T steps.  python tools/prof_policy.py [envs] [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200 import make, ppo_lag
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T = int(sys.argv[2]) if len(sys.argv) > 2 else 20
env = make("Goal_Point_8Hazards", env_num=n, seed=0, max_ep_len=200)
ac = ppo_lag.ActorCritic(43, 2).cuda()
env.reset(); env.stats_reset()
fr = ppo_lag.FusedRollout(env, ac, T)
fr.reset_into_buffer()
fr.run(); fr.run()
torch.cuda.synchronize()
print("ok")
