"""Diagnostic: where does the time of Engine.step_host go?"""
import time, sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.engine import Engine
from guardx_b200 import _abi
lib = _abi.load_library()
n = 4096
env = Engine({'env_num': n}); env.reset()
od = env.obs_flat_size
h_act = torch.empty(8, n, 2).uniform_(-1, 1).pin_memory()
h_out = torch.empty(n * (od + 3)).pin_memory()
def loop(k):
    for t in range(k):
        env.step_host(h_act[t & 7], h_out, auto_reset_done=True)
loop(20); torch.cuda.synchronize()
t0 = time.perf_counter(); loop(500); dt = time.perf_counter() - t0
print("step_host single env us/step", dt / 500 * 1e6)
envs = [Engine({'env_num': n, '_seed': k}) for k in range(20)]
for e in envs: e.reset()
def loop2(k):
    for t in range(k):
        envs[t % 20].step_host(h_act[t & 7], h_out, auto_reset_done=True)
loop2(40); torch.cuda.synchronize()
t0 = time.perf_counter(); loop2(500); dt = time.perf_counter() - t0
print("step_host 20 envs round robin us/step", dt / 500 * 1e6)
# raw launches
act = torch.rand(n, 2, device='cuda'); obs = torch.empty(n, od, device='cuda'); rdc = torch.empty(3, n, device='cuda')
p = lambda t: C.c_void_p(t.data_ptr())
args = (env._h, p(act), p(obs), p(rdc[0]), p(rdc[1]), p(rdc[2]), None)
for _ in range(50): lib.gx_step_autoreset(*args)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): lib.gx_step_autoreset(*args)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("raw ctypes launch: host us/launch", (t1 - t0) / 2000 * 1e6, " total us/launch", (t2 - t0) / 2000 * 1e6)
# CUDA graph of 100 steps
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    sp = C.c_void_p(s.cuda_stream)
    a2 = (env._h, p(act), p(obs), p(rdc[0]), p(rdc[1]), p(rdc[2]), sp)
    for _ in range(3): lib.gx_step_autoreset(*a2)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(100): lib.gx_step_autoreset(*a2)
torch.cuda.synchronize()
for _ in range(3): g.replay()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): g.replay()
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print("graph of 100 steps: us/step", dt / 2000 * 1e6)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(); lib.gx_step_autoreset(*args); ev1.record(); torch.cuda.synchronize()
print("single launch event time us", ev0.elapsed_time(ev1) * 1e3)
# floor: graph of 100 trivial 1-block kernels
xa = torch.zeros(32, device='cuda'); xo = torch.zeros(32, device='cuda')
g2 = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    lib.gx_debug_math(2, p(xa), None, p(xo), None, 32, sp); torch.cuda.synchronize()
    with torch.cuda.graph(g2, stream=s):
        for _ in range(100): lib.gx_debug_math(2, p(xa), None, p(xo), None, 32, sp)
torch.cuda.synchronize()
for _ in range(3): g2.replay()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): g2.replay()
torch.cuda.synchronize(); print("graph floor (trivial kernel) us/node", (time.perf_counter() - t0) / 2000 * 1e6)
# rollout kernel per-step time
ro = env.rollout(8, want=("obs",)); torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
obs_buf = torch.empty(256, n, od, device='cuda')
ev0.record(); lib.gx_rollout(env._h, 256, p(obs_buf), None, None, None, None, None); ev1.record(); torch.cuda.synchronize()
print("rollout kernel us/step (obs only)", ev0.elapsed_time(ev1) * 1e3 / 256)
