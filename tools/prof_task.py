"""profiling target: config-3 extension, fused step + auto-reset.  python tools/prof_task.py [task] [envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.task_engine import make_task
task = sys.argv[1] if len(sys.argv) > 1 else "Button_Car"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
env = make_task(task, env_num=n, seed=0)
env.reset()
st = env.get_state(); st[:, 11] = torch.randint(0, 1000, (n,), device='cuda').float(); env.set_state(st)
act = torch.rand(n, 2, device='cuda') * 2 - 1
for _ in range(steps):
    env.step_autoreset(act)
torch.cuda.synchronize()
print("ok")
