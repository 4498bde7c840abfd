"""Small profiling target: N envs (random episode phases + burn-in, as bench.py), a few fused step + auto-reset launches.
usage: python tools/prof_target.py [envs] [steps] [rollout_T] [robot xml]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
T = int(sys.argv[3]) if len(sys.argv) > 3 else 0
xml = sys.argv[4] if len(sys.argv) > 4 else 'xmls/point.xml'
env = Engine({'env_num': n, 'robot_base': xml}); env.reset(); env.stats_reset()
if xml == 'xmls/point.xml':
    st = env.get_state()
    st[:, 10] = torch.randint(0, 1000, (n,), device='cuda').float()
    env.set_state(st)
    rc = env._lib.gx_rollout(env._h, 1200, None, None, None, None, None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
else:
    env.rollout(30)
act = torch.rand(n, env.act_dim, device='cuda') * 2 - 1
for _ in range(steps):
    env.step_autoreset(act)
if T:
    env.rollout(T)
torch.cuda.synchronize()
print("ok", n, steps, T)
