"""Small workload for compute-sanitizer: every kernel family once, tiny sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.engine import Engine
for cfg in ({'env_num': 70, 'num_steps': 6},                                   # lanes kernels, odd tail
            {'env_num': 45, 'num_steps': 6, 'hazards_num': 3, 'lidar_num_bins': 10, 'observe_vel': True, 'observe_acc': True},
            {'env_num': 33, 'num_steps': 6, 'pool_size': 256}):
    if cfg.get('pool_size'):
        os.environ['GX_LANES_MAX_ENVS'] = '0'                                    # one-thread kernels, default shape
    ps = cfg.pop('pool_size', 512)
    env = Engine(cfg, pool_size=ps)
    env.reset(); env.stats_reset()
    n = env.env_num
    for t in range(12):
        a = torch.rand(n, 2, device='cuda') * 2 - 1
        env.step(a, want_aux=(t % 2 == 0)); env.reset_done()
        env.step_autoreset(a)
    env.rollout(10)
    env.set_state(env.get_state()); env.stats(); env.layout_pool(); env.reset()
    h_a = torch.zeros(n, 2).pin_memory(); h_o = torch.empty(n * (env.obs_flat_size + 3)).pin_memory()
    env.step_host(h_a, h_o, auto_reset_done=True)
    torch.cuda.synchronize()
    env.close()
print("sanitize target ok")
