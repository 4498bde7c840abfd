"""Kernel-level timing of the fused step at a given batch size (graph replay, cold L2)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.engine import Engine
from guardx_b200 import _abi
lib = _abi.load_library()
sizes = [int(a) for a in sys.argv[1:]] or [1 << 20]
for n in sizes:
    R = max(2, int(252e6 / (n * 364)) + 1)
    envs = []
    for k in range(R):
        e = Engine({'env_num': n, '_seed': k}, pool_size=16384); e.reset(); envs.append(e)
    od = 43
    acts = [torch.rand(n, 2, device='cuda') * 2 - 1 for _ in range(R)]
    outs = [torch.empty(n * (od + 3), device='cuda') for _ in range(R)]
    p = lambda t, off=0: C.c_void_p(t.data_ptr() + 4 * off)
    def issue(sp, k):
        for t in range(k):
            j = t % R
            rc = lib.gx_step_autoreset(envs[j]._h, p(acts[j]), p(outs[j]), p(outs[j], n * od), p(outs[j], n * od + n), p(outs[j], n * od + 2 * n), sp)
            assert rc == 0
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        issue(C.c_void_p(side.cuda_stream), 2 * R); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        K = 20 * R
        with torch.cuda.graph(g, stream=side):
            issue(C.c_void_p(side.cuda_stream), K)
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); g.replay(); g.replay(); ev1.record(); torch.cuda.synchronize()
    us = ev0.elapsed_time(ev1) * 1e3 / (2 * K)
    print(f"envs {n:8d} batches {R:3d}  {us:8.2f} us/step  {n/us/1e3:8.3f} G env-steps/s  {n*348/us/1e3:8.1f} GB/s (348 B)  frac {n*348/us/1e3/6554.2:.3f}")
    del envs, acts, outs, g
    torch.cuda.empty_cache()
