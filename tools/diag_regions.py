"""diagnostic: per-region GPU time and host enqueue time of the bench's graph-replayed regions, for several region
sizes / chunk sizes / flush settings.   python tools/diag_regions.py"""
import ctypes as C, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from guardx_b200 import _abi
from guardx_b200.engine import Engine

lib = _abi.load_library()
n = int(os.environ.get('N', '4096'))
R = int(os.environ.get("R", "170"))
if os.environ.get("FRESH"):
    bench.BURN_IN = 1
    _ri = torch.randint
    torch.randint = lambda lo, hi, *a, **k: _ri(0, 1, *a, **k)    # every env at phase 0: no env finishes for ~100 steps
batches = [bench.Batch(Engine, torch, n, 0, seed=k, n_act=8, pool_size=16384 if n <= 16384 else 0) for k in range(R)]
for b in batches:
    b.env.stats_reset()
flush_buf = torch.empty(bench.FLUSH_BYTES // 4, device='cuda')
stream = torch.cuda.current_stream()
issue = lambda sp, t0, steps: bench.issue_steps(lib, batches, steps, sp, t0=t0)
clk = torch.zeros(64, device='cuda')
sp0 = C.c_void_p(stream.cuda_stream)
def issue_noreset(sp, t0, steps):
    for t in range(t0, t0 + steps):
        b = batches[t % R]
        rc = lib.gx_step(b.h, b.p_act[(t // R) % len(b.p_act)], b.p_obs, b.p_r, b.p_d, b.p_c, None, None, sp)
        assert rc == 0
mode = os.environ.get("MODE", "autoreset")
if mode == "zero":
    for b in batches:
        b.act.zero_()
if mode == "noreset":
    issue = issue_noreset
cases = [(int(x), 2048, not os.environ.get("NOFLUSH"), 5) for x in os.environ.get("SLIST", "2720").split(",")]
for S, chunk, flush, regions in cases:
    reg = bench.Regions(torch, issue, S, flush_buf, chunk=chunk)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(regions)]
    torch.cuda.synchronize()
    host = []
    t0 = time.perf_counter(); reg.replay(); host.append(time.perf_counter() - t0)
    for r, (e0, e1) in enumerate(evs):
        if flush:
            flush_buf.fill_(float(r))
        e0.record(stream)
        t0 = time.perf_counter(); reg.replay(); host.append(time.perf_counter() - t0)
        e1.record(stream)
        lib.gx_debug_sm_clock(C.c_void_p(clk.data_ptr() + 4 * r), 100000, sp0)
    t_enq = time.perf_counter()
    torch.cuda.synchronize()
    t_done = time.perf_counter()
    gpu = [e0.elapsed_time(e1) for e0, e1 in evs]
    dones = [float(b.out[n * 44: n * 45].sum()) for b in batches]
    gpu = [sorted(gpu)[len(gpu) // 2]]
    print(json.dumps({"mode": mode, "done_per_batch_last_step": sum(dones) / len(dones), "S": S, "chunk": chunk, "flush": flush, "us_per_step_gpu": [round(1e3 * g / S, 3) for g in gpu],
                       "host_waited_ms": round(1e3 * (t_done - t_enq), 1)}))
    del reg
    if os.environ.get("SLEEP"):
        time.sleep(float(os.environ["SLEEP"]))
