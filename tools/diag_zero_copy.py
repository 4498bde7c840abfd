"""Diagnostic: e2e step with the kernel reading/writing PINNED HOST memory directly (zero-copy)."""
import time, sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from guardx_b200.engine import Engine
from guardx_b200 import _abi
lib = _abi.load_library()
for n in (4096, 16384):
    env = Engine({'env_num': n}); env.reset()
    od = env.obs_flat_size
    h_act = torch.empty(8, n, 2).uniform_(-1, 1).pin_memory()
    h_out = torch.empty(n * (od + 3)).pin_memory()
    base = h_out.data_ptr()
    p = lambda off: C.c_void_p(base + 4 * off)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    def zc(k):
        for t in range(k):
            rc = lib.gx_step_autoreset(env._h, C.c_void_p(h_act[t & 7].data_ptr()), p(0), p(n * od), p(n * od + n), p(n * od + 2 * n), sp)
            assert rc == 0, lib.gx_last_error()
            torch.cuda.synchronize()
    zc(20)
    t0 = time.perf_counter(); zc(500); dt = time.perf_counter() - t0
    print(n, "zero-copy step us/step", dt / 500 * 1e6, "env-steps/s", n * 500 / dt)
    def cp(k):
        for t in range(k):
            env.step_host(h_act[t & 7], h_out, auto_reset_done=True)
    cp(20)
    t0 = time.perf_counter(); cp(500); dt = time.perf_counter() - t0
    print(n, "copy-based step_host us/step", dt / 500 * 1e6, "env-steps/s", n * 500 / dt)
    # device-output + single D2H, action zero-copy
    d_out = torch.empty(n * (od + 3), device='cuda'); db = d_out.data_ptr(); q = lambda off: C.c_void_p(db + 4 * off)
    def mix(k):
        for t in range(k):
            lib.gx_step_autoreset(env._h, C.c_void_p(h_act[t & 7].data_ptr()), q(0), q(n * od), q(n * od + n), q(n * od + 2 * n), sp)
            h_out.copy_(d_out, non_blocking=True)
            torch.cuda.synchronize()
    mix(20)
    t0 = time.perf_counter(); mix(500); dt = time.perf_counter() - t0
    print(n, "zero-copy action + D2H copy us/step", dt / 500 * 1e6, "env-steps/s", n * 500 / dt)
