"""ncu target for the articulated step kernel: ant, 16384 envs, a few fused step + auto-reset launches.
    python tools/prof_art.py [robot] [envs] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from guardx_b200.engine import Engine  # noqa: E402

robot = sys.argv[1] if len(sys.argv) > 1 else 'ant'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 16
env = Engine({'env_num': n, 'robot_base': f'xmls/{robot}.xml', 'num_steps': 1000, '_seed': 0})
env.reset()
out = env.rollout(steps, want=("obs", "reward", "cost", "done"))
torch.cuda.synchronize()
print("ok", robot, n, steps, bool(torch.isfinite(out['obs']).all().item()))
