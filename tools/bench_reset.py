import sys, torch
sys.path.insert(0, '/root/repo')
from guardx_b200.engine import Engine
for n in (4096, 65536):
    env = Engine({'env_num': n, '_seed': 0}); env.reset(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): env.reset()
    e1.record(); torch.cuda.synchronize()
    print('reset() envs', n, 'us per call %.1f' % (e0.elapsed_time(e1) * 100))
