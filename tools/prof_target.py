"""Small profiling target: N envs, a few fused steps and one short rollout.
usage: python tools/prof_target.py [envs] [steps] [rollout_T]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32
env = Engine({'env_num': n}); env.reset(); env.stats_reset()
act = torch.rand(n, 2, device='cuda') * 2 - 1
for _ in range(steps):
    env.step_autoreset(act)
if T:
    env.rollout(T)
torch.cuda.synchronize()
print("ok", n, steps, T)
