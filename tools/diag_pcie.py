"""Diagnostic: host<->device copy latency/bandwidth on the GPU box (pinned vs pageable)."""
import time, torch
torch.cuda.init()
s = torch.cuda.current_stream()
for nbytes in (32768, 753664, 8 << 20, 256 << 20):
    n = nbytes // 4
    d = torch.empty(n, device='cuda')
    hp = torch.empty(n).pin_memory()
    hg = torch.empty(n)
    for name, h in (("pinned", hp), ("pageable", hg)):
        for direction in ("d2h", "h2d"):
            for _ in range(3):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True)); torch.cuda.synchronize()
            reps = 50 if nbytes < (64 << 20) else 5
            t0 = time.perf_counter()
            for _ in range(reps):
                (h.copy_(d, non_blocking=True) if direction == "d2h" else d.copy_(h, non_blocking=True))
                torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / reps
            print(f"{nbytes:>10} B {name:8s} {direction}: {dt*1e6:9.1f} us  {nbytes/dt/1e9:7.2f} GB/s")
# launch latency of an empty-ish kernel through torch
x = torch.zeros(1, device='cuda')
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(2000): x.add_(1)
torch.cuda.synchronize(); print("torch add_ launch+exec per call us", (time.perf_counter()-t0)/2000*1e6)
t0 = time.perf_counter()
for _ in range(200): x.add_(1); torch.cuda.synchronize()
print("launch+sync per call us", (time.perf_counter()-t0)/200*1e6)
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max,pcie.link.width.max --format=csv; lscpu | grep -E 'Model name|^CPU\\(s\\)|NUMA'; ulimit -l", shell=True, capture_output=True, text=True).stdout)
