"""config-3 section of bench.py alone (Button_Car / Push_Car, 16384 envs, cold L2): python tools/time_task.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
torch.cuda.set_device(0)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    out = bench.task3_section(torch, s)
print(json.dumps({k: ({kk: out[k][kk] for kk in ("us_per_step", "value", "finite")} if isinstance(out[k], dict) else out[k]) for k in out if k != "workload"}))
