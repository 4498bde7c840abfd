"""Throughput of the articulated-robot step (BASELINE.json configs[3]: legged robot Goal task,
65536 envs): fused step + auto-reset launches with in-kernel U(-1,1) actions, CUDA-event timed.
    python tools/bench_art.py [robot ...]        # default: ant walker swimmer
Prints one JSON line per (robot, env_num)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from guardx_b200.engine import Engine  # noqa: E402

ROB = {'ant': 'xmls/ant.xml', 'walker': 'xmls/walker.xml', 'swimmer': 'xmls/swimmer.xml', 'doggo': 'xmls/doggo.xml'}
robots = sys.argv[1:] or ['ant', 'walker', 'swimmer']
for r in robots:
    for n in (4096, 16384, 65536, 262144):
        env = Engine({'env_num': n, 'robot_base': ROB[r], 'num_steps': 1000, '_seed': 0})
        env.reset()
        env.stats_reset()
        T = 40
        env.rollout(5)
        torch.cuda.synchronize()
        out = env.rollout(T)     # allocate once; the timed call reuses the same buffers
        torch.cuda.synchronize()
        import ctypes as C
        p = lambda t_: C.c_void_p(t_.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = env._lib.gx_rollout(env._h, T, p(out['obs']), p(out['act']), p(out['reward']), p(out['cost']), p(out['done']), env._stream())
        assert rc == 0
        e1.record()
        torch.cuda.synchronize()
        dt = e0.elapsed_time(e1) * 1e-3
        s = env.stats().cpu().numpy()
        od = env.obs_flat_size
        nv, nu = env.robot.nv, env.robot.nu
        alg = 4 * (2 * 3 * nv + 8 + 8 + 2 + 16 + od + 3)   # q,v,warm r+w; d2 r+w; goal; hazards; obs; r/c/d (actions in-kernel)
        print(json.dumps({"robot": r, "envs": n, "us_per_step": 1e6 * dt / T, "env_steps_per_s": n * T / dt,
                          "alg_bytes_per_env_step": alg, "achieved_gbs": n * T * alg / dt / 1e9,
                          "episodes": float(s[0]), "mean_cost_per_step": float(s[6] / max(s[7], 1)),
                          "finite": bool(torch.isfinite(out['obs']).all().item())}), flush=True)
        del env, out
        torch.cuda.empty_cache()
