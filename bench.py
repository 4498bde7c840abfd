#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched Goal_Point_8Hazards step on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: one fused
step + auto-reset launch (gx_step_autoreset) for all envs of the batch with
random actions already resident in HBM.  Workload (BASELINE.json configs[1]):
Goal_Point_8Hazards, 4096 envs per GPU, random actions U(-1,1).

One JSON line on stdout (rank 0).  Keys beyond the driver's contract:
  roofline      dominant kernel (gx_step_kernel) vs the measured HBM peak,
                348 algorithmic bytes per env-step
  cpu_baseline  the CPU oracle (C restatement of the reference, OpenMP) timed on
                this box's host cores on a bounded sample of the same workload
  e2e           the same metric with HOST buffers: every step's action from pinned
                host memory, every step's obs/reward/done/cost to pinned host memory,
                through gx_rollout_host (chunks pipelined over two streams);
                e2e.per_step_sync is the blocking one-call-per-step variant
  rollout       the fused multi-step kernel (gx_rollout): T steps per launch,
                state in registers, vs its own 192-byte I/O-only roofline
  legged        Goal_Ant_8Hazards (and legged.doggo: Goal_Doggo_8Hazards), 65536 envs sharded over the ranks (BASELINE configs[3]),
                the articulated step kernel (leg-per-lane mjx.step with a Newton contact solver)
  sweep         (N=1 only) step-kernel GB/s at larger batches, to show where the
                kernel sits against the roofline once launch latency is amortised

Timing (the same whatever --steps K is): a WINDOW is exactly K fused step launches; a timed
REGION is M back-to-back windows, M chosen so that a region lasts >= 20 ms; REGIONS (7) regions
are measured and the MEDIAN is reported (min / max beside it), so ms_per_step x steps = one
window.  Every region is a chain of at most TWO CUDA graphs; all regions and their CUDA events
are enqueued while the GPU is still busy with SETTLE_SECONDS of untimed replays, so neither
graph-launch latency nor the host is inside a timed region.  (Why two: with four 2048-node
graphs launched in turn, every step of a region measured 6 % slower than with three -- 5.97 vs
5.61 us, reproducibly and at a constant SM clock -- which is graph re-upload, not kernel time;
with <= 2 graphs --steps 20, 4000 and 20000 agree to 0.006 %.)

L2 policy: every timed step touches a batch whose state, actions, layout pool and output
buffers have not been touched for R-1 steps, where R batches x ~1.5 MB (+ 1.3 MB of layout
pool each) exceed twice the 126 MB L2 (config.l2 says so); the L2-resident number is reported
beside it.  (A replay starts at batch 0 again while the previous one ended at batch
(S-1) mod R, so at most R/2 of the S launches of a region can see lines of the previous
replay: < 2 %.  GX_BENCH_FLUSH=1 adds a 512 MB write before every region instead.)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ENVS_PER_GPU = 4096
ALG_BYTES_STEP = 348      # single-step API contract
ALG_BYTES_ROLLOUT = 192   # multi-step kernel, I/O only
L2_BYTES = 126e6
FLUSH_BYTES = 512 << 20       # written before every timed region: evicts the L2
REGION_SECONDS = float(os.environ.get("GX_BENCH_REGION_S", "0.025"))   # a timed region lasts at least this long, whatever --steps is
MAX_REGION_STEPS = 1 << 16    # graph nodes per region (bounds capture time when --steps is huge)
BURN_IN = 2000                # random-action steps every batch takes before anything is timed (num_steps = 1000)
SETTLE_SECONDS = float(os.environ.get("GX_BENCH_SETTLE_S", "0.3"))   # untimed replays before the timed regions (the GPU is busy while the timed work is enqueued)
FLUSH = bool(int(os.environ.get("GX_BENCH_FLUSH", "0")))   # 512 MB write before every region (see the docstring)
REGIONS = int(os.environ.get("GX_BENCH_REGIONS", "7"))   # regions measured; the median is reported
TASK = "Goal_Point_8Hazards"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled with NVML while the timed regions run."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.sm_max = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.25)   # nvidia-smi -lms 200 cadence: NVML queries take a driver lock

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._loop, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_threads():
    """all host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the
    oracle's OpenMP regions take an explicit thread count, so that does not bind)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def workload_config(n, world):
    """the `config` object: identical for the CUDA arm and the --impl reference arm (what differs -- how each arm runs the
    workload, the measured L2 working set, the region sizes -- is in the top-level keys `how` and `timing`)"""
    return {"workload": f"{TASK}, {n} envs per GPU, random-action rollout (a ~ U(-1,1), pre-generated), "
                        f"step + reset_done every step",
            "envs_per_gpu": n, "total_envs": n * world,
            "parallelism": f"dp{world} (independent env shards, no data-path collective)",
            "l2": "CUDA arm: cold -- every timed launch touches a batch (state, actions, outputs, layout pool) not touched for "
                  "R - 1 launches, R batches > 2 x 126 MB L2 (sizes in `timing`), episode phases stationary; CPU arm: host memory"}


# ------------------------------------------------------------ reference arm
def oracle_steps_per_sec(n_envs, steps, warmup, threads):
    """Times gxo_step + gxo_reset_done on preallocated host buffers."""
    import numpy as np
    from oracle.oracle import OracleEngine
    o = OracleEngine({'env_num': n_envs, 'num_steps': 1000, '_seed': 0}, threads=threads)
    o.reset()
    L, h = o.L, o.h
    rng = np.random.default_rng(0)
    # the same stationary workload as the CUDA arm: random episode phases, then a burn-in
    st = o.get_state()
    st[:, 10] = rng.integers(0, 1000, n_envs).astype(np.float32)
    o.set_state(st)
    o.rollout(500, want=())
    acts = rng.uniform(-1, 1, (64, n_envs, 2)).astype(np.float32)
    obs = np.empty((n_envs, o.obs_dim), np.float32)
    rdc = np.empty((3, n_envs), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    po, pr, pd, pc = p(obs), p(rdc[0]), p(rdc[1]), p(rdc[2])
    pa = [p(acts[i]) for i in range(64)]

    def run(k):
        for t in range(k):
            L.gxo_step(h, pa[t & 63], po, pr, pd, pc, None, None)
            L.gxo_reset_done(h, po, po, None)

    run(warmup)
    t0 = time.perf_counter()
    run(steps)
    dt = time.perf_counter() - t0
    return n_envs * steps / dt, dt


def reference_arm(args, rank, world):
    if rank != 0:
        return
    threads = host_threads()
    n = ENVS_PER_GPU * world
    # a window is args.steps steps; M windows back to back so that the timed region is ~2 s whatever --steps is
    r0, d0 = oracle_steps_per_sec(n, max(args.steps, 50), max(args.warmup, 3), threads)
    M = max(1, int(2.0 * r0 / (n * args.steps)))
    rate, dt = oracle_steps_per_sec(n, args.steps * M, max(args.warmup, 3), threads)
    dt /= M
    sample = (f"{M} windows x {args.steps} steps x {n} envs ({dt * M:.2f} s), gxo_step + gxo_reset_done, stationary workload, "
              "actions pre-generated in host memory")
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": rate, "unit": "env-steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(ENVS_PER_GPU, world),
        "how": "CPU oracle: gxo_step + gxo_reset_done per step, all host threads",
        "cpu_baseline": {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
                         "sample": sample,
                         "note": "CPU oracle (C restatement of the reference step, OpenMP); the reference itself "
                                 "is Python on jax+mujoco, neither installable here"},
        "e2e": {"value": rate, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ------------------------------------------------------------------ our arm
class Batch:
    """One 4096-env batch with its own actions and output buffers (all in HBM)."""

    def __init__(self, Engine, torch, n, offset, seed, n_act, pool_size):
        self.env = Engine({'env_num': n, 'num_steps': 1000, '_seed': seed}, env_id_offset=offset,
                          pool_size=pool_size)
        self.env.reset()
        # STATIONARY workload: a freshly reset batch has every env at step 0 of its episode, and with random
        # actions most episodes end by the num_steps timeout -- all in the same launch.  A 4096-env launch is
        # as slow as its slowest warp, so what a launch costs depends on whether ANY env finishes in it.  The
        # timed steps therefore start from mixed episode phases: each env's step counter is set to a random
        # phase (through the public state injection), then BURN_IN random-action steps (fused rollout kernel,
        # no outputs) let positions and phases decorrelate.  ~4 envs of 4096 finish in every launch.
        g = torch.Generator(device='cuda').manual_seed(4321 + offset + 7919 * seed)
        st = self.env.get_state()
        st[:, 10] = torch.randint(0, 1000, (n,), device='cuda', generator=g).float()   # Engine._steps (gx_get_state row)
        self.env.set_state(st)
        sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = self.env._lib.gx_rollout(self.env._h, BURN_IN, None, None, None, None, None, sp)
        assert rc == 0, self.env._lib.gx_last_error()
        g = torch.Generator(device='cuda').manual_seed(1234 + offset)
        self.act = torch.rand(n_act, n, 2, device='cuda', generator=g) * 2 - 1
        od = self.env.obs_flat_size
        self.out = torch.empty(n * (od + 3), device='cuda')
        base = self.out.data_ptr()
        self.p_obs = C.c_void_p(base)
        self.p_r = C.c_void_p(base + 4 * n * od)
        self.p_d = C.c_void_p(base + 4 * n * (od + 1))
        self.p_c = C.c_void_p(base + 4 * n * (od + 2))
        self.p_act = [C.c_void_p(self.act[i].data_ptr()) for i in range(n_act)]
        self.h = self.env._h


def issue_steps(lib, batches, steps, sp, t0=0):
    """enqueue `steps` fused step launches round-robin over the batches"""
    R = len(batches)
    fn = lib.gx_step_autoreset
    for t in range(t0, t0 + steps):
        b = batches[t % R]
        rc = fn(b.h, b.p_act[(t // R) % len(b.p_act)], b.p_obs, b.p_r, b.p_d, b.p_c, sp)
        if rc:
            raise RuntimeError(lib.gx_last_error().decode())


def time_steps(torch, lib, batches, steps, stream, barrier=None):
    """K fused step launches issued one by one from Python (stream launches);
    CUDA events on the launching stream, synchronise on both sides."""
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sp = C.c_void_p(stream.cuda_stream)
    if barrier:
        barrier()
    torch.cuda.synchronize()
    ev0.record(stream)
    issue_steps(lib, batches, steps, sp)
    ev1.record(stream)
    torch.cuda.synchronize()
    if barrier:
        barrier()
    return ev0.elapsed_time(ev1) * 1e-3


class Regions:
    """A timed REGION = `S` fused step launches (S = M windows x K steps) as a chain of CUDA
    graphs (at most two per region; every chunk is its own graph, so the round-robin over
    the batches runs on unbroken from the first step of a region to the last).  Legal because
    every counter the kernels use lives in device memory.

    measure(): one untimed replay keeps the GPU busy while ALL regions are enqueued --
    flush (512 MB write, evicts the 126 MB L2), event, graphs, event, flush, ... -- so that
    neither the host nor the graph-launch latency is inside a region."""


    def __init__(self, torch, issue, S, flush_buf, chunk=None):
        self.torch, self.S, self.flush_buf = torch, S, flush_buf
        self.graphs = []
        # at most TWO graphs per region: with four 2048-node graphs launched in turn every step of the region measured
        # 6 % slower than with three (5.97 vs 5.61 us, reproducibly, whatever the settle time: re-upload of the graphs'
        # device-side images, not kernel time; tools/diag_regions.py)
        chunk = chunk or max(2048, (S + 1) // 2)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        sp = C.c_void_p(side.cuda_stream)
        with torch.cuda.stream(side):
            for c0 in range(0, S, chunk):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    issue(sp, c0, min(chunk, S - c0))
                self.graphs.append(g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()

    def replay(self):
        for g in self.graphs:
            g.replay()

    def measure(self, stream, regions, barrier=None, flush=FLUSH, settle=1):
        torch = self.torch
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(regions)]
        if barrier:
            barrier()
        torch.cuda.synchronize()
        for _ in range(max(1, settle)):    # untimed: the GPU is busy (and settled) while everything below is enqueued
            self.replay()
        for r, (e0, e1) in enumerate(evs):
            if flush and self.flush_buf is not None:
                self.flush_buf.fill_(float(r))
            e0.record(stream)
            self.replay()
            e1.record(stream)
        torch.cuda.synchronize()
        if barrier:
            barrier()
        return [e0.elapsed_time(e1) * 1e-3 for e0, e1 in evs]


def max_over_ranks(torch, dist, times):
    """device-timed durations: MAX over the ranks, element by element"""
    if dist is None:
        return list(times)
    t = torch.tensor(times, device='cuda', dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def median(xs):
    s = sorted(xs)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


def load_traffic(kernel_prefix, envs):
    """dram bytes per launch of the dominant kernel from the committed ncu --set full summary of
    THIS build (profiles/traffic.json, written by tools/ncu_summary.py); None if there is none."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        for row in json.load(f):
            if row.get("kernel", "").startswith(kernel_prefix) and row.get("envs") == envs:
                return row.get("dram_bytes_per_launch"), row.get("source")
    return None, None


_REAL_STDOUT = None


def _emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="guardx", choices=["guardx", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU, help="envs per GPU (default: BASELINE configs[1])")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-legged", action="store_true", help="skip the ant / 65536-env measurement")
    ap.add_argument("--no-ppo", action="store_true", help="skip the PPO-Lagrangian section")
    ap.add_argument("--value-only", action="store_true", help="diagnostic: print value / timing after the graph-replayed regions and stop")
    args = ap.parse_args()

    # stdout carries the ONE JSON line and nothing else: native libraries (NCCL prints its version banner with
    # printf on some boxes) get fd 2 while the run lasts; the line itself is written to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    from guardx_b200 import _abi
    from guardx_b200 import distributed as gxd
    from guardx_b200.engine import Engine

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback); use --impl reference for the CPU arm"
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    barrier = (lambda: dist.barrier()) if dist else None
    lib = _abi.load_library()
    hbm_peak, peak_src = peaks()
    n = args.envs
    warmup = max(args.warmup, 3)
    K = args.steps
    stream = torch.cuda.current_stream()

    # ---- batches: R independent n-env batches, working set > 2 x L2 -----------
    od = 43
    per_batch = n * (116 + 4 * (od + 3)) + 8 * n * 8   # state + outputs + 8 action slices (the 1.3 MB layout pool is extra)
    R = int(L2_BYTES * 2.0 / per_batch) + 1
    R = max(2, min(R, 512))
    pool_size = 16384
    batches = [Batch(Engine, torch, n, rank * n, seed=k, n_act=8, pool_size=pool_size) for k in range(R)]
    for b in batches:
        b.env.stats_reset()   # in-kernel episode statistics on (what ppo_train.py:370-376 does on the host)
    working_set = R * (per_batch + pool_size * 80)
    flush_buf = torch.empty(FLUSH_BYTES // 4, device='cuda')

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- value: regions of M windows x K steps, CUDA graphs, cold L2 -------
    time_steps(torch, lib, batches, max(warmup, R), stream)            # warm-up (>= 3 steps, every batch once)
    est = max_over_ranks(torch, dist, [time_steps(torch, lib, batches, 256, stream) / 256])[0]   # stream launches: >= graph time
    m_cap = max(1, MAX_REGION_STEPS // K)
    M = max(1, min(int(REGION_SECONDS / (K * est * 0.6)) + 1, m_cap))
    issue = lambda sp, t0, steps: issue_steps(lib, batches, steps, sp, t0=t0)
    for attempt in range(4):
        reg = Regions(torch, issue, M * K, flush_buf)
        settle = int(SETTLE_SECONDS / (M * K * est * 0.7)) + 1
        region_s = max_over_ranks(torch, dist, reg.measure(stream, REGIONS, barrier, settle=settle))
        dt = median(region_s)                                         # one region = M windows = S steps
        del reg
        if dt >= 0.8 * REGION_SECONDS or M >= m_cap:
            break
        M = max(M + 1, min(int(M * REGION_SECONDS / dt) + 1, m_cap))  # identical on every rank (dt is the max over ranks)
    S = M * K
    # per-epoch statistics all-reduce (replaces mpi_statistics_scalar), inside the job but not the hot loop
    stats = torch.stack([b.env.stats() for b in batches]).sum(0)
    gxd.allreduce_stats(stats)   # NCCL all-reduce of the 8-double episode statistics vector
    value = world * n * S / dt
    if args.value_only:
        if rank == 0:
            _emit({"value": value, "ms_per_step": 1e3 * dt / S, "steps": K, "windows_per_region": M,
                   "region_ms": [1e3 * x for x in region_s], "clocks": sampler.stop()})
        if dist:
            dist.barrier()
            dist.destroy_process_group()
        return
    # ---- the same steps launched one by one from Python (no graph) ----------
    Ss = max(K, 2000)
    dt_stream = max_over_ranks(torch, dist, [time_steps(torch, lib, batches, Ss, stream, barrier)])[0]
    # ---- same graph chain over one batch only (L2-resident state) ----------
    hot = Regions(torch, lambda sp, t0, steps: issue_steps(lib, batches[:1], steps, sp, t0=t0), S, None)
    dt_hot = median(max_over_ranks(torch, dist, hot.measure(stream, 5, barrier, flush=False, settle=4)))
    del hot

    # ---- e2e: host buffers through Engine.step_host -----------------------
    h_act = torch.empty(8, n, 2).uniform_(-1, 1).pin_memory()
    h_out = torch.empty(n * (od + 3)).pin_memory()
    Ke = max(200, min(K, 1000))

    def e2e_loop(k):
        for t in range(k):
            batches[t % R].env.step_host(h_act[t & 7], h_out, auto_reset_done=True)
        return float(h_out[-1])  # the step's result is read on the host

    e2e_loop(R + 3)   # warm-up: every batch's staging buffer exists before the timed region
    if barrier:
        barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_loop(Ke)
    dt_e2e = max_over_ranks(torch, dist, [time.perf_counter() - t0])[0]
    e2e_sync_value = world * n * Ke / dt_e2e

    # ---- e2e, pipelined: the same host-buffer steps through gx_rollout_host --
    # (random actions are known up front, so step t+1 does not wait for step t's
    # outputs to reach the host: H2D, kernels and D2H of neighbouring chunks overlap)
    Tp = 512
    hp_act = torch.empty(Tp, n, 2).uniform_(-1, 1).pin_memory()
    hp_obs = torch.empty(Tp, n, od).pin_memory()
    hp_r, hp_d, hp_c = (torch.empty(Tp, n).pin_memory() for _ in range(3))
    reps_p = 8

    Rp = 2           # PCIe, not L2, bounds this path; two batches alternate

    def e2e_pipe(k):
        acc = 0.0
        for j in range(k):
            batches[j % Rp].env.rollout_host(hp_act, hp_obs, hp_r, hp_d, hp_c)
            acc += float(hp_r[-1, -1])   # the rollout's result is read on the host
        return acc

    e2e_pipe(5 * Rp)   # warm-up: chunk rings / streams exist, and gx_rollout_host has timed both of its routes (DMA ring,
                       # zero-copy) on every handle -- with all ranks running -- and settled on the faster one
    if barrier:
        barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e2e_pipe(reps_p)
    dt_pipe = max_over_ranks(torch, dist, [time.perf_counter() - t0])[0]
    e2e_value = world * n * Tp * reps_p / dt_pipe
    _rates = (C.c_double * 2)()
    e2e_route = lib.gx_rollout_host_route(batches[0].h, _rates)
    e2e_route = {"route": {0: "dma ring", 1: "zero-copy tile stores"}.get(e2e_route, "undecided"),
                 "probe_dma": _rates[0], "probe_zero_copy": _rates[1]}
    del hp_obs

    # ---- fused rollout kernel ----------------------------------------------
    T = 256
    ro_obs = torch.empty(T, n, od, device='cuda')
    ro_rcd = torch.empty(3, T, n, device='cuda')
    ro_act = torch.empty(T, n, 2, device='cuda')
    env0 = batches[0].env
    pr = lambda t_: C.c_void_p(t_.data_ptr())
    sp = C.c_void_p(stream.cuda_stream)

    def rollout_once():
        rc = lib.gx_rollout(env0._h, T, pr(ro_obs), pr(ro_act), pr(ro_rcd[0]), pr(ro_rcd[1]), pr(ro_rcd[2]), sp)
        assert rc == 0, lib.gx_last_error()

    for _ in range(3):
        rollout_once()
    reps = 40
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record(stream)
    for _ in range(reps):
        rollout_once()
    ev1.record(stream)
    torch.cuda.synchronize()
    dt_ro = max_over_ranks(torch, dist, [ev0.elapsed_time(ev1) * 1e-3])[0]
    ro_value = world * n * T * reps / dt_ro
    del ro_obs, ro_rcd, ro_act
    # ---- BASELINE configs[3]: legged robots, 65536 envs sharded over the ranks ----------
    legged = None
    if not args.no_legged:
        def legged_run(xml, n_leg, T_, settle):
            env = Engine({'env_num': n_leg, 'robot_base': xml, 'num_steps': 1000, '_seed': 0}, env_id_offset=rank * n_leg)
            env.reset()
            env.stats_reset()
            lo = env.rollout(T_)      # warm-up + buffers
            args_ = (env._h, T_, pr(lo['obs']), pr(lo['act']), pr(lo['reward']), pr(lo['cost']), pr(lo['done']), sp)
            for _ in range(settle):   # contacts established
                lib.gx_rollout(*args_)
            torch.cuda.synchronize()
            if barrier:
                barrier()
            ts = []
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                rc = lib.gx_rollout(*args_)
                assert rc == 0, lib.gx_last_error()
                e1.record(stream)
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
            d_ = median(max_over_ranks(torch, dist, ts))
            fin = bool(torch.isfinite(lo['obs']).all().item())
            del env, lo
            torch.cuda.empty_cache()
            return d_, fin

        n_leg, Tl = 65536 // world, 64
        dt_leg, fin = legged_run('xmls/ant.xml', n_leg, Tl, 0)
        leg_bytes = 4 * (2 * 3 * 11 + 8 + 2 + 16 + 64 + 3)   # q,v,warm r+w | d2 r+w | goal | hazards | obs | r,c,d
        legged = {"workload": "Goal_Ant_8Hazards, 65536 envs in total, random actions drawn in-kernel, fused step + "
                              "auto-reset, one launch per step (BASELINE.json configs[3])",
                  "envs_per_gpu": n_leg, "steps": Tl, "value": world * n_leg * Tl / dt_leg, "unit": "env-steps/s",
                  "us_per_step": 1e6 * dt_leg / Tl, "scaling": "strong", "alg_bytes_per_env_step": leg_bytes,
                  "achieved_gbs": n_leg * Tl * leg_bytes / dt_leg / 1e9,
                  "bound": "issue/latency (Newton contact solver per env), not HBM", "finite": fin}
        Td = 10
        dt_dog, fin = legged_run('xmls/doggo.xml', n_leg, Td, 3)
        legged["doggo"] = {"workload": "Goal_Doggo_8Hazards, 65536 envs in total, same protocol (free-joint root, capsule "
                                       "contacts, dense Newton per env)",
                           "envs_per_gpu": n_leg, "steps": Td, "value": world * n_leg * Td / dt_dog, "unit": "env-steps/s",
                           "us_per_step": 1e6 * dt_dog / Td, "bound": "dependent-issue latency (DESIGN.md 5b), not HBM",
                           "finite": fin}
    # ---- BASELINE configs[2]: Button task, Car robot, vases / pillars / gremlins / hazards, 16384 envs on 1 GPU ----
    task3 = None
    if not args.no_legged and rank == 0:
        task3 = task3_section(torch, stream)
    # ---- BASELINE configs[4]: PPO-Lagrangian rollout + update, env on the CUDA kernels ----
    ppo = None
    if not args.no_ppo:
        ppo = ppo_section(torch, dist, rank, world, local_rank)
    clocks = sampler.stop()

    # the dominant kernel's average duration over the timed region (one launch per step)
    step_us = 1e6 * dt / S
    achieved = n * ALG_BYTES_STEP / (step_us * 1e-6) / 1e9

    sweep = None
    if rank == 0 and world == 1 and not args.no_sweep:
        sweep = []
        for big in (65536, 1 << 20):
            per = big * (116 + 4 * (od + 3)) + 8 * big * 2
            Rb = max(2, int(L2_BYTES * 2.0 / per) + 1)
            bl = [Batch(Engine, torch, big, 0, seed=100 + k, n_act=2, pool_size=pool_size) for k in range(Rb)]
            e_ = time_steps(torch, lib, bl, 4 * Rb, stream) / (4 * Rb)
            kk = max(2 * Rb, min(4096, int(REGION_SECONDS / e_) + 1))
            gs = Regions(torch, lambda sp_, t0, steps: issue_steps(lib, bl, steps, sp_, t0=t0), kk, flush_buf)
            ds = gs.measure(stream, 5, settle=6)
            d = median(ds)
            del gs
            gbs = big * ALG_BYTES_STEP * kk / d / 1e9
            sweep.append({"envs": big, "batches": Rb, "steps_per_region": kk, "us_per_step": 1e6 * d / kk,
                          "env_steps_per_s": big * kk / d, "achieved_gbs": gbs, "frac": gbs / hbm_peak,
                          "region_ms_min_max": [1e3 * min(ds), 1e3 * max(ds)]})
            del bl
            torch.cuda.empty_cache()

    cpu_base = None
    if rank == 0 and world == 1:
        threads = host_threads()
        rate, _ = oracle_steps_per_sec(n, 400, 20, threads)          # calibrate
        ks = int(max(400, min(12.0 * rate / n, 200000)))             # ~12 s of CPU work
        rate, d_cpu = oracle_steps_per_sec(n, ks, 20, threads)
        rate1, d1 = oracle_steps_per_sec(n, max(100, ks // (4 * threads)), 5, 1)
        cpu_base = {"value": rate, "unit": "env-steps/s", "cores": threads, "kind": "port",
                    "sample": f"{ks} steps x {n} envs ({d_cpu:.1f} s), gxo_step + gxo_reset_done, host-resident actions",
                    "single_thread": rate1}

    if rank == 0:
        kname = "gx_step_lanes_kernel" if n <= 9216 else "gx_step_kernel"
        traffic, traffic_src = load_traffic(kname, n)
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": warmup, "ms_per_step": 1e3 * dt / S, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(n, world),
            "how": "CUDA: fused step + auto-reset (gx_step_autoreset), one kernel launch per step, replayed as CUDA graphs, "
                   "launches chained by programmatic dependent launch (" + ("off: GX_PDL=0" if os.environ.get("GX_PDL", "1")[:1] == "0" else "on") + "), "
                   f"in-kernel episode statistics; stationary workload (random episode phases + {BURN_IN} burn-in steps per batch)",
            "timing": {"l2_batches": R, "l2_working_set_mb": working_set / 1e6, "l2_flush_between_regions": FLUSH,
                       "window_steps": K, "settle_seconds": SETTLE_SECONDS,
                       "windows_per_region": M, "launches_per_region": S, "regions": REGIONS,
                       "region_ms_median": 1e3 * dt, "region_ms_min": 1e3 * min(region_s), "region_ms_max": 1e3 * max(region_s),
                       "window_ms": 1e3 * dt / M},
            "value_l2_resident": world * n * S / dt_hot,
            "value_stream_launch": world * n * Ss / dt_stream,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": n * 2 * 4,
                    "d2h_bytes_per_step": n * (od + 3) * 4, "steps": Tp * reps_p, "rank0_route": e2e_route,
                    "per_step_sync": {"value": e2e_sync_value, "steps": Ke,
                                      "path": "Engine.step_host: one blocking call per step (policy in the loop)"},
                    "path": "Engine.rollout_host -> gx_rollout_host: every step's action from pinned host memory, every "
                            "step's obs|reward|done|cost into pinned host memory, chunks pipelined over two streams; "
                            "host clock around the calls, results read on the host"},
            "gpu_launches": S,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "kernel": kname + "<true,true>" if n <= 9216 else kname,
                         "alg_bytes_per_env_step": ALG_BYTES_STEP, "envs_per_launch": n,
                         "us_per_launch": step_us, "peak_source": peak_src,
                         "note": "latency-bound at 4096 envs (1.4 MB per launch, one dependent chain per env); "
                                 "see sweep for larger batches"},
            "rollout": {"value": ro_value, "unit": "env-steps/s", "steps_per_launch": T, "launches": reps,
                        "alg_bytes_per_env_step": ALG_BYTES_ROLLOUT,
                        "achieved_gbs": ro_value / world * ALG_BYTES_ROLLOUT / 1e9,
                        "frac": ro_value / world * ALG_BYTES_ROLLOUT / 1e9 / hbm_peak},
            "episode_stats": gxd.episode_summary(stats.tolist()),
        }
        if ppo is not None:
            line["ppo_lag"] = ppo
        if task3 is not None:
            line["config3"] = task3
        if legged is not None:
            line["legged"] = legged
        if sweep is not None:
            line["sweep"] = sweep
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        _emit(line)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def task3_section(torch, stream):
    """BASELINE.json configs[2]: 'Button/Push Car robot with vases, pillars and gremlins, 16384 envs, 1 B200' on the
    config-3 extension (include/guardx_task.h; NO reference implementation exists).  Fused step +
    auto-reset, random actions resident in HBM, cold L2 (rotating batches), same region protocol as the headline."""
    from guardx_b200 import task_engine
    n = 16384
    lib = task_engine.load_library()
    out = {}
    for name in ("Button_Car", "Push_Car"):
        envs = [task_engine.make_task(name, env_num=n, seed=k, max_ep_len=1000) for k in range(2)]
        od = envs[0].obs_flat_size
        per = n * 4 * (envs[0].state_dim + od + 3 + 2 * 4)
        Rb = int(L2_BYTES * 2.0 / per) + 1
        envs += [task_engine.make_task(name, env_num=n, seed=k, max_ep_len=1000) for k in range(2, Rb)]
        bufs = []
        g = torch.Generator(device='cuda').manual_seed(99)
        for e in envs:
            e.reset()
            st = e.get_state()
            st[:, 11] = torch.randint(0, 1000, (n,), device='cuda', generator=g).float()      # random episode phases
            e.set_state(st)
            act = torch.rand(4, n, 2, device='cuda', generator=g) * 2 - 1
            bufs.append((e, act, torch.empty(n, od, device='cuda'), torch.empty(3, n, device='cuda')))
        pr = lambda t_: C.c_void_p(t_.data_ptr())

        def issue(sp, t0, steps):
            for t in range(t0, t0 + steps):
                e, act, obs, rdc = bufs[t % Rb]
                rc = lib.gxt_step(e._h, pr(act[(t // Rb) % 4]), pr(obs), pr(rdc[0]), pr(rdc[1]), pr(rdc[2]), None, 1, sp)
                assert rc == 0, lib.gx_last_error()

        sp = C.c_void_p(stream.cuda_stream)
        issue(sp, 0, 40 * Rb)                      # warm-up / burn-in
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream); issue(sp, 0, 4 * Rb); ev1.record(stream); torch.cuda.synchronize()
        est = ev0.elapsed_time(ev1) * 1e-3 / (4 * Rb)
        S = max(2 * Rb, min(4096, int(REGION_SECONDS / est) + 1))
        reg = Regions(torch, issue, S, None)
        ds = reg.measure(stream, 5, settle=max(1, int(SETTLE_SECONDS / (S * est)) + 1))
        d = median(ds)
        del reg
        alg = 4 * (2 * 14 + 2 * (envs[0].state_dim - 16) // 1 // 2 + od + 3 + 2)   # dynamic state r+w, objects r, obs + r/d/c w, action r
        out[name] = {"envs": n, "obs_dim": od, "objects": int(envs[0].layout_dim - 2 - 4) // 2, "steps_per_region": S,
                     "us_per_step": 1e6 * d / S, "value": n * S / d, "unit": "env-steps/s", "alg_bytes_per_env_step": alg,
                     "achieved_gbs": n * S * alg / d / 1e9, "finite": bool(torch.isfinite(bufs[0][2]).all().item())}
        del bufs, envs
        torch.cuda.empty_cache()
    out["workload"] = ("Button / Push task, Car robot, 4 hazards + 4 vases + 2 pillars + 4 gremlins (+ 4 buttons | box + goal), 16384 envs, "
                       "fused step + auto-reset, cold L2 (BASELINE.json configs[2]); no reference implementation exists: parity unpinned")
    return out


def ppo_section(torch, dist, rank, world, local_rank):
    """BASELINE.json configs[4]: PPO-Lagrangian on Goal_Point_8Hazards, 4096 envs per GPU, max_ep_len 200
    (ppo_train.py:354-417 rollout + :314-348 update).  Returns env-steps/s of the rollout alone (policy forward +
    sample + env step + buffer writes, replayed as a CUDA graph) and of whole epochs (rollout + GAE + update)."""
    from guardx_b200 import make, ppo_lag
    n, T, epochs = 4096, 200, 3
    env = make(TASK, env_num=n, seed=rank, max_ep_len=T)
    cfg = ppo_lag.PPOLagConfig(epochs=epochs, max_ep_len=T, target_cost=0.0)
    ac, log = ppo_lag.ppo_lag(env, cfg)
    torch.cuda.synchronize()
    per_epoch = (log[-1]['Time'] - log[0]['Time']) / (epochs - 1)     # epoch 0 carries the warm-up
    roll = ppo_lag.RolloutTimer(env, ac, T)
    ts = []
    for _ in range(5):
        ts.append(roll.run())
    dt_roll = median(max_over_ranks(torch, dist, ts))
    per_epoch = max_over_ranks(torch, dist, [per_epoch])[0]
    out = {"workload": f"PPO-Lagrangian, {TASK}, {n} envs per GPU, {T} steps per epoch (BASELINE.json configs[4])",
           "rollout": {"value": world * n * T / dt_roll, "unit": "env-steps/s", "us_per_step": 1e6 * dt_roll / T,
                       "how": roll.how},
           "rollout_plus_update": {"value": world * n * T / per_epoch, "unit": "env-steps/s", "epoch_s": per_epoch,
                                   "how": "rollout + GAE kernels + 80/80/80 full-batch Adam iterations (torch), gradient "
                                          "all-reduce over NCCL"},
           "last_epoch": {k: log[-1][k] for k in ("EpRet", "EpCost", "Lambda", "KL", "StopIter") if k in log[-1]}}
    del env
    torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
