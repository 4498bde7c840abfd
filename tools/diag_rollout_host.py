"""Diagnostic: per-call time of Engine.rollout_host (cold and warm calls, several T)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.engine import Engine
n, od = 4096, 43
envs = [Engine({'env_num': n, '_seed': k}) for k in range(6)]
for e in envs: e.reset()
for T in (256, 512):
    ha = torch.empty(T, n, 2).uniform_(-1, 1).pin_memory()
    ho = torch.empty(T, n, od).pin_memory()
    hr, hd, hc = (torch.empty(T, n).pin_memory() for _ in range(3))
    ho.zero_(); torch.cuda.synchronize()
    for k, e in enumerate(envs[:3]):
        for rep in range(3):
            t0 = time.perf_counter()
            e.rollout_host(ha, ho, hr, hd, hc)
            dt = time.perf_counter() - t0
            print(f"T={T} env{k} call{rep}: {dt*1e3:8.2f} ms  {n*T/dt/1e6:8.1f} M env-steps/s", flush=True)
