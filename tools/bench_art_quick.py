"""One-line timing of the articulated step: python tools/bench_art_quick.py robot envs [T]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.engine import Engine
robot, n = sys.argv[1], int(sys.argv[2]); T = int(sys.argv[3]) if len(sys.argv) > 3 else 40
env = Engine({'env_num': n, 'robot_base': f'xmls/{robot}.xml', 'num_steps': 1000, '_seed': 0}); env.reset()
out = env.rollout(T); torch.cuda.synchronize()
p = lambda t_: C.c_void_p(t_.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
env._lib.gx_rollout(env._h, T, p(out['obs']), p(out['act']), p(out['reward']), p(out['cost']), p(out['done']), env._stream())
e1.record(); torch.cuda.synchronize()
dt = e0.elapsed_time(e1) * 1e-3
print(os.environ.get('GUARDX_LIB', 'default'), robot, n, 'us/step %.1f' % (1e6 * dt / T), 'Menv-steps/s %.1f' % (n * T / dt / 1e6), bool(torch.isfinite(out['obs']).all()))
