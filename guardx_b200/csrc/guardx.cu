// guardx.cu -- kernels and C ABI of libguardx.so (see include/guardx.h).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//        -Xcompiler -ffp-contract=off,-fPIC -shared guardx.cu -o ../libguardx.so
// -fmad=false / -ffp-contract=off are REQUIRED (bit-parity contract, gx_math.h).
#include <chrono>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/guardx.h"
#include "gx_device.cuh"
#include "gx_lanes.cuh"
#include "gx_policy.cuh"
#include "gx_art.cuh"
#include "gx_art_launch.h"
#include "gx_tree_launch.h"

#define GX_API extern "C" __attribute__((visibility("default")))

// Resident blocks per SM the one-thread-per-env step kernel is compiled for
// (register cap = 65536 / (128 * this)).  ncu at 2^20 envs: 85 registers -> 5
// blocks -> 20 warps/SM, issue slots 57 % busy, i.e. occupancy-limited.
#ifndef GX_STEP_MIN_BLOCKS
#define GX_STEP_MIN_BLOCKS 8
#endif

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int gx_fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

// same, for the other translation units of the library (gx_task.cu)
int gx_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define GX_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return gx_fail(GX_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                     __FILE__, __LINE__);                                               \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) {
      cudaSetDevice(dev);
      switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// ---------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------
struct gx_env {
  gx_config cfg;
  gx_params p;
  GxState s;
  int device;
  int sm_count;
  float* pool;          // [pool_size, lay_dim]
  int32_t* pool_tries;  // [pool_size]
  double* stats;        // [8] device
  // Device-resident counters, so that every entry point can be captured in a
  // CUDA graph and replayed: the step count lives per env in the state (d2.zw,
  // "age"); ctr[1] = epoch (layout pools generated so far), ctr[3] = the
  // finished-blocks ticket of the pool kernel.
  unsigned long long* ctr;
  bool have_reset;
  bool stats_on;
  unsigned long long stats_age0;  // env age at the last gx_stats_reset
  int lanes_max_envs;   // use the 16-lanes-per-env kernels up to this batch size (0 = never)
  // staging for gx_step_host
  float* stage;  // [n*act_dim | n*obs_dim | n | n | n]
  // gx_step_host zero-copy route: cached mapping of the caller's pinned buffers
  const float* zc_dact; float *zc_dobs, *zc_dr, *zc_dd, *zc_dc; bool zc_ok;
  // articulated robots (GX_ROBOT_ART): host + device copy of the model, joint-state SoA
  GxArtModel* art_host;
  GxArtDev art;
  int art_shape;  // 1 ant-like (4 legs x 2 dofs), 2 walker-like (2 x 5), 3 swimmer-like (1 x 2)
  // general-tree robots (GX_ROBOT_TREE): host + device copy of the model, state, scratch
  GxTreeModel* tree_host;
  GxTreeDev tree;
  // gx_rollout_host pipeline: two chunk buffers, a compute and a copy stream
  float* ring[2];
  int ring_steps;
  cudaStream_t rs_compute, rs_copy;
  cudaEvent_t ev_in, ev_computed[2], ev_copied[2];
  // route of gx_rollout_host: 0 = DMA ring (copy engine), 1 = zero-copy (the kernels' tile stores go to the mapped
  // host buffers).  Which one is faster depends on what else shares the host side (alone on the box the copy engine
  // wins; with 8 ranks the DMA route collapsed to 1/4 in round 1), so the first calls time both and the faster stays.
  int rh_calls, rh_route;
  double rh_rate[2];
};

static inline int div_up(int a, int b) { return (a + b - 1) / b; }

// Threads per block for one-thread-per-env kernels: small batches are spread
// over as many SMs as possible (latency-bound regime), large batches use 128.
static inline int pick_block(const gx_env* e) {
  const int n = e->p.n;
  int block = 128;
  if (n <= e->sm_count * 32 * 4) block = 32;
  else if (n <= e->sm_count * 64 * 8) block = 64;
  while (block > 32 && (size_t)block * e->p.obs_dim * sizeof(float) > 48 * 1024) block >>= 1;
  return block;
}

// ---------------------------------------------------------------------------
// The pool epoch lives in device memory and is advanced by the LAST block of the
// pool kernel to finish (a ticket per block), i.e. after every block has taken
// its read of the old value.  (The step count is per-env state, gx_device.cuh.)
// No host-side count is baked into a launch, so a captured CUDA graph replays
// correctly.
// ---------------------------------------------------------------------------
enum { GX_CTR_T = 0, GX_CTR_EPOCH = 1, GX_CTR_TICKET = 3 };

__device__ __forceinline__ void gx_tick(unsigned long long* ctr, int slot, unsigned long long inc) {
  __syncthreads();  // all reads of ctr[slot] by this block are done
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned last = gridDim.x - 1;
    if (atomicInc(reinterpret_cast<unsigned*>(ctr + GX_CTR_TICKET), last) == last) ctr[slot] += inc;
  }
}

// ---------------------------------------------------------------------------
// layout pool (reset_layout, engine.py:433-444; sample_layout :546-572)
//
// One WARP per pool slot.  Candidate `attempt` of slot j is a pure function of
// (key, j, attempt), so the 32 lanes test 32 consecutive attempts at once and
// the lowest successful one wins -- the same layout the sequential oracle finds,
// without the geometric-tail divergence of one thread per slot.
// ---------------------------------------------------------------------------
#define GX_MAX_POOL_ATTEMPTS (1u << 16)

__device__ __forceinline__ int gx_class_of(int obj, int nh) {
  return obj == 0 ? GX_CLS_GOAL : (obj <= nh ? GX_CLS_HAZARD : GX_CLS_ROBOT);
}

__device__ bool gx_sample_layout(const gx_params& p, uint32_t k0, uint32_t k1, uint32_t slot,
                                 uint32_t attempt, float (*placed)[2]) {
  const int nh = p.nh, nobj = nh + 2;
  for (int o = 0; o < nobj; ++o) {
    const int cls = gx_class_of(o, nh);
    bool conflicted = true;
    float x = -1e30f, y = -1e30f;
    for (int t = 0; t < p.tries; ++t) {
      gx_u32x4 r = gx_philox4x32_10(slot, attempt, ((uint32_t)o << 8) | (uint32_t)t, GX_STREAM_LAYOUT, k0, k1);
      float cx = gx_uniform(r.v[0], p.lo[cls][0], p.hi[cls][0]);
      float cy = gx_uniform(r.v[1], p.lo[cls][1], p.hi[cls][1]);
      bool valid = true;
      for (int q = 0; q < o; ++q) {
        float dx = cx - placed[q][0], dy = cy - placed[q][1];
        float dist = gx_sqrt(dx * dx + dy * dy);
        if (dist < p.keepsum[gx_class_of(q, nh)][cls]) valid = false;
      }
      if (valid) { x = cx; y = cy; conflicted = false; }
    }
    if (conflicted) return false;  // success can no longer be 1 (engine.py:569)
    placed[o][0] = x;
    placed[o][1] = y;
  }
  float dx = placed[nobj - 1][0] - placed[0][0], dy = placed[nobj - 1][1] - placed[0][1];
  return !(gx_sqrt(dx * dx + dy * dy) < p.min_robot_goal);
}

__device__ void gx_pool_slot(const gx_params& p, uint32_t k0, uint32_t k1, int slot, int lane,
                             float* __restrict__ pool, int32_t* __restrict__ tries) {
  float placed[2 + GX_MAX_HAZARDS][2];
  for (uint32_t base = 0; base < GX_MAX_POOL_ATTEMPTS; base += 32) {
    const uint32_t attempt = base + lane;
    bool ok = gx_sample_layout(p, k0, k1, (uint32_t)slot, attempt, placed);
    unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m) {
      const int win = __ffs(m) - 1;
      if (lane == win) {
        float* out = pool + (size_t)slot * p.lay_dim;
        const int nobj = p.nh + 2;
        out[0] = placed[0][0]; out[1] = placed[0][1];
        out[2] = placed[nobj - 1][0]; out[3] = placed[nobj - 1][1];
        for (int h = 0; h < p.nh; ++h) { out[4 + 2 * h] = placed[1 + h][0]; out[5 + 2 * h] = placed[1 + h][1]; }
        tries[slot] = (int32_t)(base + win + 1);
      }
      return;
    }
  }
  // acceptance rate ~0: the configuration cannot be laid out (the reference
  // would fail `assert self.layout_size > self.env_num`, engine.py:444)
  if (lane == 0) tries[slot] = -1;
}

__global__ void __launch_bounds__(128) gx_pool_kernel(const __grid_constant__ gx_params p,
                                                      unsigned long long* __restrict__ ctr,
                                                      float* __restrict__ pool, int32_t* __restrict__ tries) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  // per-epoch key: one Philox call on (epoch, STREAM_LAYOUT) under the seed
  const unsigned long long epoch = ctr[GX_CTR_EPOCH];
  const gx_u32x4 key = gx_philox4x32_10((uint32_t)epoch, (uint32_t)(epoch >> 32), 0u, GX_STREAM_LAYOUT,
                                        p.seed_lo, p.seed_hi);
  if (slot < p.pool_size) gx_pool_slot(p, key.v[0], key.v[1], slot, lane, pool, tries);
  gx_tick(ctr, GX_CTR_EPOCH, 1ull);
}

// ---------------------------------------------------------------------------
// placement of one env from a layout row (layout2qpos, engine.py:623-639)
// ---------------------------------------------------------------------------
template <int NH_CT>
__device__ __forceinline__ void gx_place(const gx_params& p, const GxState& s, int i, const float* lay,
                                         float& gx_, float& gy_, float4* hz_regs, GxDyn& d) {
  const int nh = NH_CT ? NH_CT : p.nh;
  if (NH_CT && ((reinterpret_cast<uintptr_t>(lay) & 15u) == 0) && (NH_CT % 2 == 0)) {
    const float4* l4 = reinterpret_cast<const float4*>(lay);
    float4 a = l4[0];
    gx_ = a.x; gy_ = a.y; d.qx = a.z; d.qy = a.w;
#pragma unroll
    for (int k = 0; k < NH_CT / 2; ++k) {
      hz_regs[k] = l4[1 + k];
      s.hz[(size_t)k * s.cap + i] = hz_regs[k];
    }
  } else {
    gx_ = lay[0]; gy_ = lay[1]; d.qx = lay[2]; d.qy = lay[3];
    for (int k = 0; k < s.nh2; ++k) {
      float4 h;
      h.x = lay[4 + 4 * k]; h.y = lay[5 + 4 * k];
      h.z = (2 * k + 1 < nh) ? lay[6 + 4 * k] : 0.0f;
      h.w = (2 * k + 1 < nh) ? lay[7 + 4 * k] : 0.0f;
      if (NH_CT) hz_regs[k] = h;
      s.hz[(size_t)k * s.cap + i] = h;
    }
  }
  s.goal[i] = make_float2(gx_, gy_);
  d.qth = 0.0f; d.vx = 0.0f; d.vy = 0.0f; d.vw = 0.0f;
}

template <int NH_CT>
__device__ __forceinline__ void gx_load_static(const GxState& s, int i, float& gx_, float& gy_, float4* hz_regs) {
  float2 g = s.goal[i];
  gx_ = g.x; gy_ = g.y;
  if (NH_CT) {
#pragma unroll
    for (int k = 0; k < (NH_CT + 1) / 2; ++k) hz_regs[k] = s.hz[(size_t)k * s.cap + i];
  }
}


// ---- Engine.reset (engine.py:454-467) -------------------------------------
template <int NH_CT>
__global__ void gx_reset_kernel(const __grid_constant__ gx_params p, const GxState s,
                                const float* __restrict__ pool, const float* __restrict__ layout,
                                float* __restrict__ obs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int wbase = i - lane;
  if (wbase >= p.n) return;
  float* tile = gx_smem + (size_t)warp * 32 * p.obs_dim;
  float* row = tile + lane * p.obs_dim;
  if (i < p.n) {
    GxDyn d;
    gx_load_dyn(s, i, d);  // reset() leaves Engine._done (engine.py:463) and the key stale
    const float* lay = gx_pick_layout(p, pool, layout, i, d.age, GX_STREAM_PICK ^ 1u);
    float4 hz_regs[NH_CT ? (NH_CT + 1) / 2 : 1];
    float gx_, gy_;
    gx_place<NH_CT>(p, s, i, lay, gx_, gy_, hz_regs, d);
    d.steps = 0.0f;
    d.prev_dist = gx_reset_row<NH_CT>(p, s, i, d, gx_, gy_, hz_regs, row);
    gx_kinematics(d.qth, d.hc, d.hs);  // mjx.forward at the reset pose (engine.py:655): R = I
    if ((p.off_vel >= 0) | (p.off_acc >= 0)) {  // Engine._data.xpos <- reset pose; _last_data stays
      float4 h = s.hx[i];
      h.x = d.qx; h.y = d.qy;
      s.hx[i] = h;
    }
    gx_store_dyn(s, i, d);
    if (s.ep) s.ep[i] = make_float4(0.f, 0.f, 0.f, s.ep[i].w);
  }
  const int rows = min(32, p.n - wbase);
  gx_store_tile(tile, obs + (size_t)wbase * p.obs_dim, rows, p.obs_dim, lane);
}

// ---- Engine.reset_done (engine.py:497-505, 702-731) -------------------------
// Re-place ONE env whose done flag is set.  Lag heading, prev_dist, steps and the
// done flag itself are deliberately left as they are (the reference returns the
// un-stepped data and never clears _done).
template <int NH_CT>
__device__ __forceinline__ void gx_reset_done_env(const gx_params& p, const GxState& s, int i,
                                                  const float* lay, GxDyn& d, float& gx_, float& gy_,
                                                  float4* hz_regs, float* row) {
  gx_place<NH_CT>(p, s, i, lay, gx_, gy_, hz_regs, d);
  gx_reset_row<NH_CT>(p, s, i, d, gx_, gy_, hz_regs, row);
}

// ---- Engine.step + Engine.reset_done with ONE observation pass (thread per env) ----------------
// Same idea as gx_step_autoreset_lanes (gx_lanes.cuh): whether the env finishes is known before the
// lidar (goal distance, jump, timeout; and the NaN/Inf guard reduces to ctrl / qpos / qvel when pose
// and objects are finite and of sane magnitude), so a finishing env is re-placed FIRST and the single
// observation pass produces its reset observation; reward, cost, violations, flags are the step's.
// A finishing thread pays a Philox draw, five 128-bit pool loads and eight square roots instead of a
// second full observation pass, which is what its 31 warp-mates would otherwise wait for.
// Exotic states -- reachable through gx_set_state / caller layouts only -- take the reference order
// (step observation, then re-place, then reset observation), out of line.
// `fused_ok` (uniform): observe_vel / observe_acc off and a lidar that cannot overflow.
// (Tried: hazards left in the state arrays and a NON-unrolled hazard loop, to shrink the ~3 k-instruction
// hot path after ncu showed stall_no_instruction 2.0 per issue at 2^20 envs: no gain there, and 13.3 -> 19.9 us
// per step at 65 536 envs, where the in-loop loads are exposed -- dropped.)
template <int NH_CT>
__device__ __noinline__ GxStepOut gx_step_then_reset_env(const gx_params& p, const GxState& s, const float* pool, int i, GxDyn& d,
                                                         float& gx_, float& gy_, float4* hz_regs, float a0, float a1, float* row,
                                                         float* qacc) {
  GxStepOut o = gx_step_env<NH_CT>(p, s, i, d, gx_, gy_, hz_regs, a0, a1, row, qacc, nullptr);
  if (o.done > 0.0f) {
    const float* lay = gx_pick_layout(p, pool, nullptr, i, d.age, GX_STREAM_PICK);
    gx_reset_done_env<NH_CT>(p, s, i, lay, d, gx_, gy_, hz_regs, row);
  }
  return o;
}

template <int NH_CT>
__device__ __forceinline__ GxStepOut gx_step_env_autoreset(const gx_params& p, const GxState& s, const float* pool, int i,
                                                           GxDyn& d, float& gx_, float& gy_, float4* hz_regs, float a0,
                                                           float a1, float* row, float* qacc, bool fused_ok) {
  bool sane = fused_ok & (gx_abs(d.qx) < GX_SANE_BOUND) & (gx_abs(d.qy) < GX_SANE_BOUND) & !gx_notfinite(d.qth) &
              (gx_abs(gx_) < GX_SANE_BOUND) & (gx_abs(gy_) < GX_SANE_BOUND);
#pragma unroll
  for (int k = 0; k < (NH_CT + 1) / 2; ++k)
    sane &= (gx_abs(hz_regs[k].x) < GX_SANE_BOUND) & (gx_abs(hz_regs[k].y) < GX_SANE_BOUND) &
            (gx_abs(hz_regs[k].z) < GX_SANE_BOUND) & (gx_abs(hz_regs[k].w) < GX_SANE_BOUND);
  if (!sane) return gx_step_then_reset_env<NH_CT>(p, s, pool, i, d, gx_, gy_, hz_regs, a0, a1, row, qacc);
  float c0 = d.hc * a0, c1 = d.hs * a0, c2 = a1;
  const float last_dist = d.prev_dist, last_done = d.done;
  float px, py, r00, r10;
  for (int k = 0; k < p.substeps; ++k) gx_mjx_step_point(p, d, c0, c1, c2, px, py, r00, r10, qacc);
  float dist;
  {
    float dx = gx_ - px, dy = gy_ - py;
    dist = gx_sqrt(dx * dx + dy * dy);
  }
  const int bad = gx_notfinite(c0) | gx_notfinite(c1) | gx_notfinite(c2) | gx_notfinite(d.qx) | gx_notfinite(d.qy) |
                  gx_notfinite(d.qth) | gx_notfinite(d.vx) | gx_notfinite(d.vy) | gx_notfinite(d.vw);
  const float last = last_done > 0.0f ? dist : last_dist;
  const float d_dist = last - dist;
  GxStepOut out;
  out.reward = d_dist * p.reward_distance;
  out.done = dist < p.goal_size ? 1.0f : 0.0f;
  out.flags = out.done > 0.0f ? GX_AUX_GOAL : 0;
  if (gx_abs(d_dist) > 1.0f) { out.done = 1.0f; out.reward = 0.0f; out.flags |= GX_AUX_JUMP; }
  if (bad) { out.reward = 0.0f; out.done = 1.0f; out.flags |= GX_AUX_GUARD; }
  if (d.steps > p.num_steps) { out.done = 1.0f; out.flags |= GX_AUX_TIMEOUT; }
  d.steps = out.done > 0.0f ? 0.0f : d.steps + 1.0f;
  d.done = out.done;
  d.prev_dist = dist;
  d.hc = r00;
  d.hs = r10;
  d.age += 1;
  float cost_step = 0.0f;
  int viol_step = 0;
  if (out.done > 0.0f) {
    // the step's cost at the step pose (engine.py:804-811), then the re-placement
#pragma unroll
    for (int h = 0; h < NH_CT; ++h) {
      const float4 v = hz_regs[h >> 1];
      const float hx = (h & 1) ? v.z : v.x, hy = (h & 1) ? v.w : v.y;
      float dx = hx - px, dy = hy - py;
      float dh = gx_sqrt(dx * dx + dy * dy);
      float term = p.hazards_size - gx_min(dh, p.hazards_size);
      if (term > 0.0f) ++viol_step;
      cost_step = cost_step + term;
    }
    const float* lay = gx_pick_layout(p, pool, nullptr, i, d.age, GX_STREAM_PICK);
    gx_place<NH_CT>(p, s, i, lay, gx_, gy_, hz_regs, d);
    px = d.qx; py = d.qy; r00 = 1.0f; r10 = 0.0f;   // gx_kinematics(0) returns exactly (1, 0)
    c0 = 0.0f; c1 = 0.0f; c2 = 0.0f;
  }
  GxObsOut o = gx_obs_row<NH_CT>(p, s, i, gx_, gy_, hz_regs, d, c0, c1, c2, px, py, r00, r10, row, nullptr);
  out.cost = out.done > 0.0f ? cost_step : o.cost;
  out.violations = out.done > 0.0f ? viol_step : o.violations;
  return out;
}

// ---- episode statistics (ppo_train.py:370-376, 405-411 on the host) --------
// Per-env accumulators only -- no atomics and no grid-wide reduction in the step
// (4096 same-address atomics per launch cost more than the step itself).  The
// sums over envs are taken, deterministically, when the statistics are read.
// out[8] = n_episodes, sum_ret, sum_cost, sum_len, sum_ret^2, sum_cost^2,
// total_cost, total_env_steps.  One block, fixed tree => deterministic.
__global__ void __launch_bounds__(1024) gx_stats_reduce_kernel(const GxState s, int n, unsigned long long age0,
                                                                double* __restrict__ out) {
  __shared__ double sh[8][32];
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float4 f0 = s.fin[i], f1 = s.fin[(size_t)s.cap + i], e = s.ep[i], c = s.d2[i];
    const unsigned long long age = (unsigned long long)__float_as_uint(c.z) | ((unsigned long long)__float_as_uint(c.w) << 32);
    acc[0] += f0.x; acc[1] += f0.y; acc[2] += f0.z; acc[3] += f0.w;
    acc[4] += f1.x; acc[5] += f1.y; acc[6] += e.w; acc[7] += (double)(age - age0);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double v = acc[k];
    for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    if (lane == 0) sh[k][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double v = lane < (blockDim.x >> 5) ? sh[k][lane] : 0.0;
      for (int off = 16; off; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
      if (lane == 0) out[k] = v;
    }
  }
}

// ---- Engine.step (engine.py:469-495) ---------------------------------------
// qacc (info['obs']['qacc'], engine.py:763-764) is emitted when the pointer is
// non-NULL (uniform branch); AUX selects the variant that also writes the integer
// aux row.
// AUTORESET fuses Engine.reset_done into the same launch: reward/done/cost are
// the step's, the observation row of a done env is its reset observation (what
// step() followed by reset_done() returns).
template <int NH_CT, bool AUX, bool STATS, bool AUTORESET>
__global__ void __launch_bounds__(128, GX_STEP_MIN_BLOCKS) gx_step_kernel(const __grid_constant__ gx_params p, const __grid_constant__ GxState s,
                               const float* __restrict__ pool,
                               const float2* __restrict__ action, float* __restrict__ obs,
                               float* __restrict__ reward, float* __restrict__ done,
                               float* __restrict__ cost, float* __restrict__ qacc,
                               int32_t* __restrict__ aux, double* __restrict__ stats, int bulk) {
  gx_pdl_sync();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int wbase = i - lane;
  if (wbase >= p.n) return;
  float* tile = gx_smem + (size_t)warp * 32 * p.obs_dim;
  float* row = tile + lane * p.obs_dim;
  if (i < p.n) {
    GxDyn d;
    gx_load_dyn(s, i, d);
    constexpr bool FUSED = AUTORESET && NH_CT != 0 && !AUX;
    float4 hz_regs[NH_CT ? (NH_CT + 1) / 2 : 1];
    float gx_, gy_;
    gx_load_static<NH_CT>(s, i, gx_, gy_, hz_regs);
    const float2 a = action[i];
    float4 ep_acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (STATS) ep_acc = gx_stats_load(s, i);
    float qa[3];
    int32_t bins[AUX ? 1 + GX_MAX_HAZARDS : 1];
    GxStepOut o;
    if constexpr (FUSED) {
      const bool fused_ok = (p.off_vel < 0) & (p.off_acc < 0) & (p.use_exp ? p.neg_exp_gain <= 0.0f : p.lidar_max_dist > 0.0f);
      o = gx_step_env_autoreset<NH_CT>(p, s, pool, i, d, gx_, gy_, hz_regs, a.x, a.y, row, qacc ? qa : nullptr, fused_ok);
    } else {
      o = gx_step_env<NH_CT>(p, s, i, d, gx_, gy_, hz_regs, a.x, a.y, row, qacc ? qa : nullptr, AUX ? bins : nullptr);
      if (AUTORESET && o.done > 0.0f) {
        const float* lay = gx_pick_layout(p, pool, nullptr, i, d.age, GX_STREAM_PICK);
        gx_reset_done_env<NH_CT>(p, s, i, lay, d, gx_, gy_, hz_regs, row);
      }
    }
    gx_store_dyn(s, i, d);
    reward[i] = o.reward;
    done[i] = o.done;
    cost[i] = o.cost;
    if (qacc) { qacc[3 * (size_t)i] = qa[0]; qacc[3 * (size_t)i + 1] = qa[1]; qacc[3 * (size_t)i + 2] = qa[2]; }
    if (AUX) {
      int32_t* ax = aux + (size_t)i * p.aux_dim;
      ax[0] = o.flags; ax[1] = o.violations;
      for (int k = 0; k < 1 + p.nh; ++k) ax[2 + k] = bins[k];
    }
    if (STATS) gx_stats_commit(s, i, ep_acc, o.reward, o.cost, o.done);
  }
  const int rows = min(32, p.n - wbase);
  gx_store_tile_any(tile, obs + (size_t)wbase * p.obs_dim, rows, p.obs_dim, lane, bulk != 0);
}

template <int NH_CT>
__global__ void gx_reset_done_kernel(const __grid_constant__ gx_params p, const GxState s,
                                     const float* __restrict__ pool, const float* __restrict__ layout,
                                     const float* obs_in, float* obs_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int wbase = i - lane;
  if (wbase >= p.n) return;
  const int rows = min(32, p.n - wbase);
  float* tile = gx_smem + (size_t)warp * 32 * p.obs_dim;
  float* row = tile + lane * p.obs_dim;
  const bool inplace = (obs_in == obs_out);
  bool is_done = false;
  if (i < p.n) is_done = (__float_as_uint(s.d2[i].y) >> 31) != 0;
  const unsigned any = __ballot_sync(0xffffffffu, is_done);
  if (!inplace) {
    // stage the previous observation rows of this warp (Engine._obs)
    const int total = rows * p.obs_dim;
    const float* src = obs_in + (size_t)wbase * p.obs_dim;
    if (rows == 32 && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0)) {
      const float4* s4 = reinterpret_cast<const float4*>(src);
      float4* t4 = reinterpret_cast<float4*>(tile);
      for (int k = lane; k < (total >> 2); k += 32) t4[k] = s4[k];
    } else {
      for (int k = lane; k < total; k += 32) tile[k] = src[k];
    }
    __syncwarp();
  } else if (!any) {
    return;
  }
  if (is_done) {
    GxDyn d;
    gx_load_dyn(s, i, d);
    float4 hz_regs[NH_CT ? (NH_CT + 1) / 2 : 1];
    float gx_, gy_;
    const float* lay = gx_pick_layout(p, pool, layout, i, d.age, GX_STREAM_PICK);
    gx_reset_done_env<NH_CT>(p, s, i, lay, d, gx_, gy_, hz_regs, row);
    gx_store_dyn(s, i, d);
  }
  if (!inplace) {
    gx_store_tile(tile, obs_out + (size_t)wbase * p.obs_dim, rows, p.obs_dim, lane);
  } else if (is_done) {
    float* dst = obs_out + (size_t)i * p.obs_dim;
    for (int k = 0; k < p.obs_dim; ++k) dst[k] = row[k];
  }
}

// ---- synthetic random-action rollout (trpo/myTest.py:28-49 + reset_done) ----
// T control steps in ONE launch; dynamic state, goal and hazards stay in
// registers between steps; HBM sees only the per-step outputs.
template <int NH_CT, bool STATS>
__global__ void gx_rollout_kernel(const __grid_constant__ gx_params p, const GxState s,
                                  const float* __restrict__ pool, int T,
                                  float* __restrict__ obs, float2* __restrict__ act,
                                  float* __restrict__ reward, float* __restrict__ cost,
                                  float* __restrict__ done, double* __restrict__ stats) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int wbase = i - lane;
  if (wbase >= p.n) return;
  const int rows = min(32, p.n - wbase);
  const bool live = i < p.n;
  float* tile = gx_smem + (size_t)warp * 32 * p.obs_dim;
  float* row = tile + lane * p.obs_dim;
  GxDyn d;
  float4 hz_regs[NH_CT ? (NH_CT + 1) / 2 : 1];
  float gx_ = 0.f, gy_ = 0.f;
  if (live) {
    gx_load_dyn(s, i, d);
    gx_load_static<NH_CT>(s, i, gx_, gy_, hz_regs);
  }
  const uint64_t g = p.env_id_offset + (uint64_t)i;
  const size_t n = (size_t)p.n;
  for (int t = 0; t < T; ++t) {
    if (live) {
      const uint64_t tg = d.age;
      gx_u32x4 u = gx_philox4x32_10((uint32_t)g, (uint32_t)tg, (uint32_t)(tg >> 32), GX_STREAM_ACTION, p.seed_lo, p.seed_hi);
      const float a0 = gx_uniform(u.v[0], -1.0f, 1.0f), a1 = gx_uniform(u.v[1], -1.0f, 1.0f);
      GxStepOut o = gx_step_env<NH_CT>(p, s, i, d, gx_, gy_, hz_regs, a0, a1, row, nullptr, nullptr);
      if (STATS) gx_stats_step(s, i, o.reward, o.cost, o.done);
      if (o.done > 0.0f) {
        const float* lay = gx_pick_layout(p, pool, nullptr, i, d.age, GX_STREAM_PICK);
        gx_reset_done_env<NH_CT>(p, s, i, lay, d, gx_, gy_, hz_regs, row);
      }
      const size_t k = (size_t)t * n + (size_t)i;
      if (act) act[k] = make_float2(a0, a1);
      if (reward) reward[k] = o.reward;
      if (cost) cost[k] = o.cost;
      if (done) done[k] = o.done;
    }
    if (obs) gx_store_tile(tile, obs + ((size_t)t * n + (size_t)wbase) * p.obs_dim, rows, p.obs_dim, lane);
  }
  if (live) gx_store_dyn(s, i, d);
}

// ---- 16-lanes-per-env variants (small batches; gx_lanes.cuh) ------------------
template <bool STATS, bool AUTORESET>
__global__ void __launch_bounds__(128) gx_step_lanes_kernel(
    const __grid_constant__ gx_params p, const __grid_constant__ GxState s, const float* __restrict__ pool,
    const float2* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
    float* __restrict__ done, float* __restrict__ cost, float* __restrict__ qacc, double* __restrict__ stats) {
  gx_pdl_sync();
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, sub = threadIdx.x & 15;
  const int env = gtid >> 4;
  if (((gtid - lane) >> 4) >= p.n) return;  // whole warp past the end
  const bool live = env < p.n;
  const int i = live ? env : p.n - 1;
  const unsigned gmask = gx_grp_mask();
  GxDyn d;
  gx_load_dyn(s, i, d);
  const float2 g = s.goal[i];
  float gx_ = g.x, gy_ = g.y, ox, oy;
  gx_load_object(s, i, sub, gx_, gy_, ox, oy);
  const float2 a = action[i];
  float4 ep_acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (STATS && sub == 9) ep_acc = gx_stats_load(s, i);   // needed at the very end: load it now
  float qa[3];
  GxLaneStep r;
  if (AUTORESET) {
    // the layout this env is re-placed with if the step ends its episode (keyed by the env's age after the step):
    // drawn and loaded now, so that it is in registers when the step knows
    const float* lay = gx_pick_layout(p, pool, nullptr, i, d.age + 1, GX_STREAM_PICK);
    const GxSpecLayout sp = gx_spec_layout(lay, sub);
    r = gx_step_autoreset_lanes(p, s, i, sub, gmask, live, lay, sp, d, gx_, gy_, ox, oy, a.x, a.y, qacc ? qa : nullptr);
  } else {
    r = gx_step_lanes(p, sub, 0xffffffffu, gmask, d, gx_, gy_, ox, oy, a.x, a.y, qacc ? qa : nullptr);
  }
  if (live) {
    gx_store_row_lanes(obs + (size_t)i * 43, sub, r.obs, r.c0, r.c1, r.c2, d);
    gx_store_dyn_lanes(s, i, sub, d);
    if (sub == 3) reward[i] = r.out.reward;
    if (sub == 4) done[i] = r.out.done;
    if (sub == 5) cost[i] = r.out.cost;
    if (qacc && sub >= 6 && sub <= 8) qacc[3 * (size_t)i + (sub - 6)] = qa[sub - 6];
    if (STATS && sub == 9) gx_stats_commit(s, i, ep_acc, r.out.reward, r.out.cost, r.out.done);
  }
}

template <bool STATS>
__global__ void __launch_bounds__(128) gx_rollout_lanes_kernel(
    const __grid_constant__ gx_params p, const GxState s, const float* __restrict__ pool, int T,
    float* __restrict__ obs, float2* __restrict__ act, float* __restrict__ reward, float* __restrict__ cost,
    float* __restrict__ done, double* __restrict__ stats) {
  const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31, sub = threadIdx.x & 15;
  const int env = gtid >> 4;
  if (((gtid - lane) >> 4) >= p.n) return;
  const bool live = env < p.n;
  const int i = live ? env : p.n - 1;
  const unsigned gmask = gx_grp_mask();
  GxDyn d;
  gx_load_dyn(s, i, d);
  const float2 g = s.goal[i];
  float gx_ = g.x, gy_ = g.y, ox, oy;
  gx_load_object(s, i, sub, gx_, gy_, ox, oy);
  const uint64_t gid = p.env_id_offset + (uint64_t)i;
  const size_t n = (size_t)p.n;
  for (int t = 0; t < T; ++t) {
    const uint64_t tg = d.age;
    gx_u32x4 u = gx_philox4x32_10((uint32_t)gid, (uint32_t)tg, (uint32_t)(tg >> 32), GX_STREAM_ACTION, p.seed_lo, p.seed_hi);
    const float a0 = gx_uniform(u.v[0], -1.0f, 1.0f), a1 = gx_uniform(u.v[1], -1.0f, 1.0f);
    GxLaneStep r = gx_step_lanes(p, sub, 0xffffffffu, gmask, d, gx_, gy_, ox, oy, a0, a1, nullptr);
    if (STATS && live && sub == 9) gx_stats_step(s, i, r.out.reward, r.out.cost, r.out.done);
    if (r.out.done > 0.0f) {
      const float* lay = gx_pick_layout(p, pool, nullptr, i, d.age, GX_STREAM_PICK);
      gx_reset_done_lanes(p, s, i, sub, gmask, live, lay, d, gx_, gy_, ox, oy, r.obs);
      r.c0 = 0.0f; r.c1 = 0.0f; r.c2 = 0.0f;
    }
    if (live) {
      const size_t k = (size_t)t * n + (size_t)i;
      if (obs) gx_store_row_lanes(obs + k * 43, sub, r.obs, r.c0, r.c1, r.c2, d);
      if (act && sub == 2) act[k] = make_float2(a0, a1);
      if (reward && sub == 3) reward[k] = r.out.reward;
      if (done && sub == 4) done[k] = r.out.done;
      if (cost && sub == 5) cost[k] = r.out.cost;
    }
  }
  if (live) gx_store_dyn_lanes(s, i, sub, d);
}

// ---- state injection / checkpoint -------------------------------------------
__global__ void gx_get_state_kernel(const __grid_constant__ gx_params p, const GxState s, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  GxDyn d;
  gx_load_dyn(s, i, d);
  float* o = out + (size_t)i * p.state_dim;
  o[0] = d.qx; o[1] = d.qy; o[2] = d.qth; o[3] = d.vx; o[4] = d.vy; o[5] = d.vw;
  o[6] = d.hc; o[7] = d.hs; o[8] = d.prev_dist; o[9] = d.done; o[10] = d.steps;
  float2 g = s.goal[i];
  o[11] = g.x; o[12] = g.y;
  for (int h = 0; h < p.nh; ++h) {
    float4 v = s.hz[(size_t)(h >> 1) * s.cap + i];
    o[13 + 2 * h] = (h & 1) ? v.z : v.x;
    o[14 + 2 * h] = (h & 1) ? v.w : v.y;
  }
  float* x = o + 13 + 2 * p.nh;
  const float4 hx = s.hx[i];
  const uint32_t hf = s.hf[i];
  x[0] = hx.x; x[1] = hx.y; x[2] = hx.z; x[3] = hx.w;
  x[4] = (hf & 1u) ? 1.0f : 0.0f;
  x[5] = (float)((hf >> 1) & 3u);
}

__global__ void gx_set_state_kernel(const __grid_constant__ gx_params p, const GxState s, const float* in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const float* o = in + (size_t)i * p.state_dim;
  GxDyn d;
  gx_load_dyn(s, i, d);  // keeps the env's age
  d.qx = o[0]; d.qy = o[1]; d.qth = o[2]; d.vx = o[3]; d.vy = o[4]; d.vw = o[5];
  d.hc = o[6]; d.hs = o[7]; d.prev_dist = o[8]; d.done = o[9]; d.steps = o[10];
  gx_store_dyn(s, i, d);
  s.goal[i] = make_float2(o[11], o[12]);
  for (int k = 0; k < s.nh2; ++k) {
    float4 v;
    v.x = o[13 + 4 * k]; v.y = o[14 + 4 * k];
    v.z = (2 * k + 1 < p.nh) ? o[15 + 4 * k] : 0.0f;
    v.w = (2 * k + 1 < p.nh) ? o[16 + 4 * k] : 0.0f;
    s.hz[(size_t)k * s.cap + i] = v;
  }
  const float* x = o + 13 + 2 * p.nh;
  s.hx[i] = make_float4(x[0], x[1], x[2], x[3]);
  int hist = (int)x[5];
  hist = hist < 0 ? 0 : (hist > 2 ? 2 : hist);
  s.hf[i] = (x[4] > 0.0f ? 1u : 0u) | ((uint32_t)hist << 1);
}

// ---- state injection / checkpoint: one row per env =
//      qpos[nv] qvel[nv] qacc_warmstart[nv] prev_dist prev_done steps goal[2] hazards[2*nh]
__global__ void gx_art_get_state_kernel(const __grid_constant__ gx_params p, const GxState s, const GxArtDev a, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  float* o = out + (size_t)i * p.state_dim;
  const size_t cap = (size_t)s.cap;
  for (int k = 0; k < 3 * a.nv; ++k) o[k] = a.aq[(size_t)k * cap + i];
  o += 3 * a.nv;
  const float4 c = s.d2[i];
  o[0] = c.x; o[1] = (__float_as_uint(c.y) >> 31) ? 1.0f : 0.0f; o[2] = fabsf(c.y);
  const float2 g = s.goal[i];
  o[3] = g.x; o[4] = g.y;
  for (int h = 0; h < p.nh; ++h) gx_art_hazard(s, nullptr, i, h, o[5 + 2 * h], o[6 + 2 * h]);
}

__global__ void gx_art_set_state_kernel(const __grid_constant__ gx_params p, const GxState s, const GxArtDev a, const float* in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const float* o = in + (size_t)i * p.state_dim;
  const size_t cap = (size_t)s.cap;
  for (int k = 0; k < 3 * a.nv; ++k) a.aq[(size_t)k * cap + i] = o[k];
  o += 3 * a.nv;
  const float4 c = s.d2[i];
  s.d2[i] = make_float4(o[0], o[1] > 0.0f ? -o[2] : o[2], c.z, c.w);
  s.goal[i] = make_float2(o[3], o[4]);
  for (int k = 0; k < s.nh2; ++k) {
    float4 h;
    h.x = o[5 + 4 * k]; h.y = o[6 + 4 * k];
    h.z = (2 * k + 1 < p.nh) ? o[7 + 4 * k] : 0.0f;
    h.w = (2 * k + 1 < p.nh) ? o[8 + 4 * k] : 0.0f;
    s.hz[(size_t)k * cap + i] = h;
  }
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
GX_API int gx_abi_version(void) { return GX_ABI_VERSION; }
GX_API int gx_config_size(void) { return (int)sizeof(gx_config); }
GX_API const char* gx_last_error(void) { return g_err; }
GX_API void gx_default_config(gx_config* cfg) {
  if (cfg) gx_fill_default_config(cfg);
}

// 16 lanes per env pay off while the batch cannot fill the machine with one
// thread per env; beyond that the one-thread kernels do less total work.
static bool lanes_path(const gx_env* e) {
  const gx_params& p = e->p;
  return e->lanes_max_envs > 0 && p.n <= e->lanes_max_envs && p.nh == 8 && p.nb == 16 && p.obs_dim == 43 &&
         p.alias && p.use_exp && p.ctrllimited && p.forcelimited && p.substeps == 1 &&
         p.off_ctrl == 0 && p.off_compass == 3 && p.off_glidar == 5 && p.off_hlidar == 21 &&
         p.off_qpos == 37 && p.off_qvel == 40;
}

static size_t tile_bytes(const gx_env* e, int block) { return (size_t)block * e->p.obs_dim * sizeof(float); }
static bool fast_path(const gx_env* e) { return e->p.nh == 8; }

static int gx_art_shape_of(const GxArtModel* m) {
  int kmax = 0, nbmax = 0;
  for (int l = 0; l < m->nleg; ++l) { kmax = m->leg[l].nk > kmax ? m->leg[l].nk : kmax; nbmax = m->leg[l].nb > nbmax ? m->leg[l].nb : nbmax; }
  if (m->nleg == 4 && kmax <= 2 && nbmax <= 3) return 1;
  if (m->nleg == 2 && kmax <= 5 && nbmax <= 3) return 2;
  if (m->nleg == 1 && kmax <= 2 && nbmax <= 2) return 3;
  return 0;
}

static int gx_create_impl(const gx_config* h_cfg, const GxArtModel* h_model, int device, gx_env** out,
                          const GxTreeModel* h_tree = nullptr) {
  if (!out) return gx_fail(GX_EINVAL, "gx_create: out is NULL");
  *out = nullptr;
  if (h_cfg && (h_cfg->robot == GX_ROBOT_ART) != (h_model != nullptr))
    return gx_fail(GX_EINVAL, "gx_create: GX_ROBOT_ART needs gx_create_art with a model (and only it)");
  if (h_cfg && (h_cfg->robot == GX_ROBOT_TREE) != (h_tree != nullptr))
    return gx_fail(GX_EINVAL, "gx_create: GX_ROBOT_TREE needs gx_create_tree with a model (and only it)");
  if (h_tree && (h_cfg->art_nv != h_tree->nv || h_cfg->art_nu != h_tree->nu || h_cfg->art_nq != h_tree->nq ||
                 h_tree->nworld != 2 + 2 * h_cfg->hazards_num || h_tree->nv > GX_TREE_MAXV || h_tree->npair > GX_TREE_MAXPAIR))
    return gx_fail(GX_EINVAL, "gx_create_tree: model and config disagree (nq / nv / nu / hazards) or exceed the build limits");
  if (h_model) {
    int nvr = 3;
    if (h_model->nleg < 1 || h_model->nleg > GX_ART_MAXLEG) return gx_fail(GX_EINVAL, "gx_create_art: nleg out of range");
    for (int l = 0; l < h_model->nleg; ++l) nvr += h_model->leg[l].nk;
    if (nvr != h_model->nv || h_cfg->art_nv != h_model->nv || h_cfg->art_nu != h_model->nu)
      return gx_fail(GX_EINVAL, "gx_create_art: nv / nu of the model and the config disagree");
    if (!gx_art_shape_of(h_model))
      return gx_fail(GX_EINVAL, "gx_create_art: no kernel instance for this robot shape (legs x dofs x bodies)");
  }
  gx_params p;
  memset(&p, 0, sizeof p);
  if (const char* why = gx_lower_config(h_cfg, &p)) return gx_fail(GX_EINVAL, "gx_create: %s", why);
  int ndev = 0;
  GX_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return gx_fail(GX_EINVAL, "gx_create: device %d of %d", device, ndev);
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  GX_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return gx_fail(GX_ECUDA, "gx_create: device %d is sm_%d%d; libguardx.so is built for sm_100a only", device, prop.major, prop.minor);
  gx_env* e = new (std::nothrow) gx_env();
  if (!e) return gx_fail(GX_ENOMEM, "gx_create: host allocation failed");
  e->cfg = *h_cfg;
  e->p = p;
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  e->ctr = nullptr; e->have_reset = false; e->stats_on = false; e->stats_age0 = 0;
  e->stage = nullptr;
  e->tree_host = nullptr; e->tree.m = nullptr; e->tree.tq = nullptr; e->tree.tw = nullptr; e->tree.scratch = nullptr;
  e->art_host = nullptr; e->art.m = nullptr; e->art.aq = nullptr; e->art.nv = 0; e->art.nu = 0; e->art_shape = 0;
  e->lanes_max_envs = 9216;   // measured crossover with the one-thread kernel (us per step, lanes vs thread-per-env, after the latter's hazard
                              // loop was rolled): 8192 -> 6.5 vs 7.2, 10 240 -> 8.4 vs 7.4, 12 288 -> 9.0 vs 7.5, 16 384 -> 10.3 vs 7.7
  if (const char* v = getenv("GX_LANES_MAX_ENVS")) e->lanes_max_envs = atoi(v);
  const int cap = div_up(p.n, 32) * 32;
  GxState& s = e->s;
  memset(&s, 0, sizeof s);
  s.cap = cap;
  s.nh2 = (p.nh + 1) / 2;
  e->pool = nullptr; e->pool_tries = nullptr; e->stats = nullptr;
#define GX_ALLOC(ptr, bytes)                                                         \
  do {                                                                               \
    cudaError_t e_ = cudaMalloc((void**)&(ptr), (bytes));                            \
    if (e_ != cudaSuccess) {                                                         \
      gx_destroy(e);                                                                 \
      return gx_fail(e_ == cudaErrorMemoryAllocation ? GX_ENOMEM : GX_ECUDA,         \
                     "gx_create: cudaMalloc(%zu) failed: %s", (size_t)(bytes), cudaGetErrorString(e_)); \
    }                                                                                \
    cudaMemset((ptr), 0, (bytes));                                                   \
  } while (0)
  GX_ALLOC(s.d0, sizeof(float4) * cap);
  GX_ALLOC(s.d1, sizeof(float4) * cap);
  GX_ALLOC(s.d2, sizeof(float4) * cap);
  GX_ALLOC(s.goal, sizeof(float2) * cap);
  GX_ALLOC(s.hz, sizeof(float4) * (size_t)cap * s.nh2);
  GX_ALLOC(s.ep, sizeof(float4) * cap);
  GX_ALLOC(s.fin, sizeof(float4) * 2 * (size_t)cap);
  GX_ALLOC(s.hx, sizeof(float4) * cap);
  GX_ALLOC(s.hf, sizeof(uint32_t) * cap);
  GX_ALLOC(e->pool, sizeof(float) * (size_t)p.pool_size * p.lay_dim);
  GX_ALLOC(e->pool_tries, sizeof(int32_t) * (size_t)p.pool_size);
  GX_ALLOC(e->stats, sizeof(double) * 8);
  GX_ALLOC(e->ctr, sizeof(unsigned long long) * 4);
  if (h_model) {
    e->art_host = new (std::nothrow) GxArtModel(*h_model);
    e->art_shape = gx_art_shape_of(h_model);
    e->art.nv = h_model->nv; e->art.nu = h_model->nu;
    GxArtModel* dm = nullptr;
    GX_ALLOC(dm, sizeof(GxArtModel));
    cudaMemcpy(dm, h_model, sizeof(GxArtModel), cudaMemcpyHostToDevice);
    e->art.m = dm;
    GX_ALLOC(e->art.aq, sizeof(float) * 3 * (size_t)h_model->nv * cap);
  }
  if (h_tree) {
    e->tree_host = new (std::nothrow) GxTreeModel(*h_tree);
    e->tree.nq = h_tree->nq; e->tree.nv = h_tree->nv; e->tree.nu = h_tree->nu; e->tree.scratch_floats = gx_tree_scratch_floats();
    GxTreeModel* dm = nullptr;
    GX_ALLOC(dm, sizeof(GxTreeModel));
    cudaMemcpy(dm, h_tree, sizeof(GxTreeModel), cudaMemcpyHostToDevice);
    e->tree.m = dm;
    GX_ALLOC(e->tree.tq, sizeof(float) * 2 * (size_t)(h_tree->nq + 2 * h_tree->nv) * cap);
    GX_ALLOC(e->tree.tw, sizeof(float) * (size_t)(h_tree->nworld + 2) * cap);   // + z history of ego_vel_acc
    GX_ALLOC(e->tree.scratch, sizeof(float) * (size_t)gx_tree_scratch_floats() * cap);
  }
#undef GX_ALLOC
  // never-forwarded mjx.Data carries the identity xmat that put_data copied in
  {
    float4* h = (float4*)malloc(sizeof(float4) * cap);
    for (int i = 0; i < cap; ++i) h[i] = make_float4(0.f, 0.f, 1.f, 0.f);
    cudaMemcpy(s.d1, h, sizeof(float4) * cap, cudaMemcpyHostToDevice);
    free(h);
  }
  GX_CUDA(cudaDeviceSynchronize());
  *out = e;
  return GX_OK;
}

GX_API int gx_create(const gx_config* h_cfg, int device, gx_env** out) { return gx_create_impl(h_cfg, nullptr, device, out); }
GX_API int gx_create_art(const gx_config* h_cfg, const GxArtModel* h_model, int device, gx_env** out) {
  if (!h_model) return gx_fail(GX_EINVAL, "gx_create_art: model is NULL");
  return gx_create_impl(h_cfg, h_model, device, out);
}

GX_API int gx_create_tree(const gx_config* h_cfg, const GxTreeModel* h_model, int device, gx_env** out) {
  if (!h_model) return gx_fail(GX_EINVAL, "gx_create_tree: model is NULL");
  return gx_create_impl(h_cfg, nullptr, device, out, h_model);
}

GX_API int gx_destroy(gx_env* e) {
  if (!e) return GX_OK;
  DeviceGuard guard(e->device);
  cudaFree(e->s.d0); cudaFree(e->s.d1); cudaFree(e->s.d2); cudaFree(e->s.goal);
  cudaFree(e->s.hz); cudaFree(e->s.ep); cudaFree(e->s.fin); cudaFree(e->s.hx); cudaFree(e->s.hf);
  cudaFree(e->pool); cudaFree(e->pool_tries); cudaFree(e->stats); cudaFree(e->stage); cudaFree(e->ctr);
  cudaFree((void*)e->art.m); cudaFree(e->art.aq); delete e->art_host;
  cudaFree((void*)e->tree.m); cudaFree(e->tree.tq); cudaFree(e->tree.tw); cudaFree(e->tree.scratch); delete e->tree_host;
  if (e->ring_steps) {
    cudaFree(e->ring[0]); cudaFree(e->ring[1]);
    cudaStreamDestroy(e->rs_compute); cudaStreamDestroy(e->rs_copy);
    cudaEventDestroy(e->ev_in);
    for (int k = 0; k < 2; ++k) { cudaEventDestroy(e->ev_computed[k]); cudaEventDestroy(e->ev_copied[k]); }
  }
  delete e;
  return GX_OK;
}

GX_API int gx_obs_dim(const gx_env* e) { return e ? e->p.obs_dim : GX_EINVAL; }
GX_API int gx_act_dim(const gx_env* e) { return e ? (e->art_host ? e->art.nu : (e->tree_host ? e->tree.nu : 2)) : GX_EINVAL; }
GX_API int gx_env_num(const gx_env* e) { return e ? e->p.n : GX_EINVAL; }
GX_API int gx_pool_size(const gx_env* e) { return e ? e->p.pool_size : GX_EINVAL; }
GX_API int gx_aux_dim(const gx_env* e) { return e ? e->p.aux_dim : GX_EINVAL; }
GX_API int gx_state_dim(const gx_env* e) { return e ? e->p.state_dim : GX_EINVAL; }

GX_API int gx_obs_offset(const gx_env* e, uint32_t component) {
  if (!e) return GX_EINVAL;
  switch (component) {
    case GX_OBS_CTRL: return e->p.off_ctrl;
    case GX_OBS_GOAL_COMPASS: return e->p.off_compass;
    case GX_OBS_GOAL_LIDAR: return e->p.off_glidar;
    case GX_OBS_HAZARDS_LIDAR: return e->p.off_hlidar;
    case GX_OBS_QPOS: return e->p.off_qpos;
    case GX_OBS_QVEL: return e->p.off_qvel;
    case GX_OBS_ACC: return e->p.off_acc;
    case GX_OBS_VEL: return e->p.off_vel;
    default: return -1;
  }
}

// engine.py:291-297: ctrlrange of the first two actuators, +-inf if not limited
GX_API int gx_action_bounds(const gx_env* e, float* h_low, float* h_high) {
  if (!e || !h_low || !h_high) return gx_fail(GX_EINVAL, "gx_action_bounds: NULL argument");
  if (e->tree_host) {
    for (int i = 0; i < e->tree_host->nv; ++i) {
      const GxTreeDof& d = e->tree_host->dof[i];
      if (d.act < 0) continue;
      h_low[d.act] = d.ctrllimited ? d.ctrl_lo : -INFINITY;
      h_high[d.act] = d.ctrllimited ? d.ctrl_hi : INFINITY;
    }
    return GX_OK;
  }
  if (e->art_host) {
    const GxArtModel* m = e->art_host;
    auto put = [&](const GxArtDof& d) {
      if (d.act < 0) return;
      h_low[d.act] = d.ctrllimited ? d.ctrl_lo : -INFINITY;
      h_high[d.act] = d.ctrllimited ? d.ctrl_hi : INFINITY;
    };
    for (int r = 0; r < 3; ++r) put(m->root.dof[r]);
    for (int l = 0; l < m->nleg; ++l) for (int k = 0; k < m->leg[l].nk; ++k) put(m->leg[l].dof[k]);
    return GX_OK;
  }
  for (int i = 0; i < 2; ++i) {
    h_low[i] = e->p.ctrllimited ? e->p.ctrl_lo : -INFINITY;
    h_high[i] = e->p.ctrllimited ? e->p.ctrl_hi : INFINITY;
  }
  return GX_OK;
}

#define GX_LAUNCH_CHECK(name)                                                   \
  do {                                                                          \
    cudaError_t e_ = cudaGetLastError();                                        \
    if (e_ != cudaSuccess)                                                      \
      return gx_fail(GX_ECUDA, "%s launch failed: %s", name, cudaGetErrorString(e_)); \
  } while (0)

static int gx_generate_pool(gx_env* e, cudaStream_t st) {
  const int block = 128;
  const int grid = div_up(e->p.pool_size * 32, block);
  gx_pool_kernel<<<grid, block, 0, st>>>(e->p, e->ctr, e->pool, e->pool_tries);
  GX_LAUNCH_CHECK("gx_pool_kernel");
  return GX_OK;
}

// ---- articulated robots: the kernels live in one translation unit per robot shape
//      (gx_art_ant.cu, gx_art_walker.cu, gx_art_swimmer.cu) so that the build parallelises
template <int MODE, bool AR>
static int gx_art_launch(gx_env* e, cudaStream_t st, const float* layout, const float* action, const float* obs_in,
                         float* obs, float* act_out, float* reward, float* done, float* cost, float* qacc, int32_t* aux = nullptr) {
  if (e->tree_host) {
    GxTreeLaunch t;
    t.p = &e->p; t.s = &e->s; t.t = &e->tree; t.host_model = e->tree_host; t.pool = e->pool; t.layout = layout; t.action = action; t.obs_in = obs_in;
    t.obs = obs; t.act_out = act_out; t.reward = reward; t.done = done; t.cost = cost; t.qacc = qacc; t.aux = aux;
    t.stats_on = e->stats_on ? 1 : 0; t.mode = MODE; t.autoreset = AR ? 1 : 0; t.stream = st;
    const cudaError_t err = gx_tree_launch(t);
    if (err != cudaSuccess) return gx_fail(GX_ECUDA, "gx_tree_kernel launch failed: %s", cudaGetErrorString(err));
    return GX_OK;
  }
  GxArtLaunch a;
  a.p = &e->p; a.s = &e->s; a.art = &e->art; a.pool = e->pool; a.layout = layout; a.action = action; a.obs_in = obs_in;
  a.obs = obs; a.act_out = act_out; a.reward = reward; a.done = done; a.cost = cost; a.qacc = qacc; a.aux = aux;
  a.stats_on = e->stats_on ? 1 : 0; a.mode = MODE; a.autoreset = AR ? 1 : 0; a.sm_count = e->sm_count; a.stream = st;
  cudaError_t err;
  switch (e->art_shape) {
    case 1: err = gx_art_launch_ant(a); break;
    case 2: err = gx_art_launch_walker(a); break;
    case 3: err = gx_art_launch_swimmer(a); break;
    default: return gx_fail(GX_EINVAL, "gx_art_launch: no kernel for this robot shape");
  }
  if (err != cudaSuccess) return gx_fail(GX_ECUDA, "gx_art_kernel launch failed: %s", cudaGetErrorString(err));
  return GX_OK;
}

static int gx_art_step(gx_env* e, cudaStream_t st, const float* action, float* obs, float* act_out, float* reward,
                       float* done, float* cost, float* qacc, bool autoreset, int32_t* aux = nullptr) {
  return autoreset ? gx_art_launch<GX_ART_STEP, true>(e, st, nullptr, action, nullptr, obs, act_out, reward, done, cost, qacc, aux)
                   : gx_art_launch<GX_ART_STEP, false>(e, st, nullptr, action, nullptr, obs, act_out, reward, done, cost, qacc, aux);
}

static int gx_reset_impl(gx_env* e, const float* layout, float* obs, cudaStream_t st) {
  if (e->art_host || e->tree_host) {
    if (int rc = gx_art_launch<GX_ART_RESET, false>(e, st, layout, nullptr, nullptr, obs, nullptr, nullptr, nullptr, nullptr, nullptr)) return rc;
    e->have_reset = true;
    return GX_OK;
  }
  const int block = pick_block(e), grid = div_up(e->p.n, block);
  const size_t sh = tile_bytes(e, block);
  if (fast_path(e))
    gx_reset_kernel<8><<<grid, block, sh, st>>>(e->p, e->s, e->pool, layout, obs);
  else
    gx_reset_kernel<0><<<grid, block, sh, st>>>(e->p, e->s, e->pool, layout, obs);
  GX_LAUNCH_CHECK("gx_reset_kernel");
  e->have_reset = true;
  return GX_OK;
}

GX_API int gx_reset(gx_env* e, float* obs, void* stream) {
  if (!e || !obs) return gx_fail(GX_EINVAL, "gx_reset: NULL argument");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (int rc = gx_generate_pool(e, st)) return rc;
  return gx_reset_impl(e, nullptr, obs, st);
}

GX_API int gx_reset_with_layout(gx_env* e, const float* layout, float* obs, void* stream) {
  if (!e || !obs || !layout) return gx_fail(GX_EINVAL, "gx_reset_with_layout: NULL argument");
  DeviceGuard guard(e->device);
  return gx_reset_impl(e, layout, obs, (cudaStream_t)stream);
}

// Programmatic dependent launch (sm_90+): the step kernels call griddepcontrol.wait before they touch global memory and
// are launched with programmatic stream serialization, so a launch is set up and its blocks become resident while the
// kernel before it in the stream is still running -- a 4096-env step is 0.43 waves and ~4 us of dependent chain, the
// ~1.5 us between two dependent launches were a quarter of the step.  Nothing changes for what the caller can observe:
// the dependent grid's wait returns only when every grid before it has completed and flushed, and a kernel that is
// launched WITHOUT the attribute after ours (torch's, anybody's) is serialized as ever.  GX_PDL=0 in the environment
// switches the attribute off (A/B measurements).
static bool gx_pdl_enabled() {
  static const bool on = [] { const char* v = getenv("GX_PDL"); return !(v && v[0] == '0'); }();
  return on;
}
template <typename... KArgs, typename... Args>
static cudaError_t gx_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t sh, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = sh; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = gx_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int NH_CT>
static void launch_step(gx_env* e, int grid, int block, size_t sh, cudaStream_t st, const float* action,
                        float* obs, float* reward, float* done, float* cost, float* qacc, int32_t* aux,
                        bool autoreset, int bulk) {
  const float2* a2 = reinterpret_cast<const float2*>(action);
#define GX_ST(AX, ST, AR) gx_launch(gx_step_kernel<NH_CT, AX, ST, AR>, grid, block, sh, st, e->p, e->s, e->pool, a2, obs, reward, done, cost, qacc, aux, e->stats, bulk)
#define GX_ST2(AX, ST) do { if (autoreset) GX_ST(AX, ST, true); else GX_ST(AX, ST, false); } while (0)
  if (aux) { if (e->stats_on) GX_ST2(true, true); else GX_ST2(true, false); }
  else { if (e->stats_on) GX_ST2(false, true); else GX_ST2(false, false); }
#undef GX_ST2
#undef GX_ST
}

static int gx_step_impl(gx_env* e, const float* action, float* obs, float* reward, float* done, float* cost,
                        float* qacc, int32_t* aux, void* stream, bool autoreset, const char* who,
                        bool tile_stores = false) {
  if (!e || !action || !obs || !reward || !done || !cost) return gx_fail(GX_EINVAL, "%s: NULL argument", who);
  if (!e->have_reset) return gx_fail(GX_ESTATE, "%s: call gx_reset first", who);
  if (reinterpret_cast<uintptr_t>(action) & 7u) return gx_fail(GX_EINVAL, "%s: action must be 8-byte aligned", who);
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (e->art_host || e->tree_host) {
    return gx_art_step(e, st, action, obs, nullptr, reward, done, cost, qacc, autoreset, aux);
  }
  // tile_stores: the caller wants the observation rows to leave as coalesced 128-bit stores
  // (outputs in mapped host memory: scattered 4-byte stores waste PCIe transactions)
  if (!aux && !tile_stores && lanes_path(e)) {
    const int block = 128, grid = div_up(e->p.n * GX_LPE, block);
    const float2* a2 = reinterpret_cast<const float2*>(action);
#define GX_SL(ST, AR) gx_launch(gx_step_lanes_kernel<ST, AR>, grid, block, 0, st, e->p, e->s, e->pool, a2, obs, reward, done, cost, qacc, e->stats)
    if (e->stats_on) { if (autoreset) GX_SL(true, true); else GX_SL(true, false); }
    else { if (autoreset) GX_SL(false, true); else GX_SL(false, false); }
#undef GX_SL
    GX_LAUNCH_CHECK("gx_step_lanes_kernel");
    return GX_OK;
  }
  const int block = pick_block(e), grid = div_up(e->p.n, block);
  const size_t sh = tile_bytes(e, block);
  // observation tiles leave as asynchronous bulk copies, except towards mapped HOST memory (tile_stores: the
  // zero-copy route of gx_step_host keeps the plain coalesced 128-bit stores)
  const int bulk = tile_stores ? 0 : 1;
  if (fast_path(e)) launch_step<8>(e, grid, block, sh, st, action, obs, reward, done, cost, qacc, aux, autoreset, bulk);
  else launch_step<0>(e, grid, block, sh, st, action, obs, reward, done, cost, qacc, aux, autoreset, bulk);
  GX_LAUNCH_CHECK("gx_step_kernel");
  return GX_OK;
}

GX_API int gx_step(gx_env* e, const float* action, float* obs, float* reward, float* done, float* cost,
                   float* qacc, int32_t* aux, void* stream) {
  return gx_step_impl(e, action, obs, reward, done, cost, qacc, aux, stream, false, "gx_step");
}

GX_API int gx_step_autoreset(gx_env* e, const float* action, float* obs, float* reward, float* done,
                             float* cost, void* stream) {
  return gx_step_impl(e, action, obs, reward, done, cost, nullptr, nullptr, stream, true, "gx_step_autoreset");
}

GX_API int gx_reset_done(gx_env* e, const float* obs_in, float* obs_out, const float* layout, void* stream) {
  if (!e || !obs_in || !obs_out) return gx_fail(GX_EINVAL, "gx_reset_done: NULL argument");
  if (!e->have_reset) return gx_fail(GX_ESTATE, "gx_reset_done: call gx_reset first");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (e->art_host || e->tree_host)
    return gx_art_launch<GX_ART_RESET_DONE, false>(e, st, layout, nullptr, obs_in, obs_out, nullptr, nullptr, nullptr, nullptr, nullptr);
  const int block = pick_block(e), grid = div_up(e->p.n, block);
  const size_t sh = tile_bytes(e, block);
  if (fast_path(e))
    gx_reset_done_kernel<8><<<grid, block, sh, st>>>(e->p, e->s, e->pool, layout, obs_in, obs_out);
  else
    gx_reset_done_kernel<0><<<grid, block, sh, st>>>(e->p, e->s, e->pool, layout, obs_in, obs_out);
  GX_LAUNCH_CHECK("gx_reset_done_kernel");
  return GX_OK;
}

GX_API int gx_get_state(gx_env* e, float* state, void* stream) {
  if (!e || !state) return gx_fail(GX_EINVAL, "gx_get_state: NULL argument");
  DeviceGuard guard(e->device);
  if (e->tree_host) gx_tree_get_state(e->p, e->s, e->tree, state, (cudaStream_t)stream);
  else if (e->art_host) gx_art_get_state_kernel<<<div_up(e->p.n, 128), 128, 0, (cudaStream_t)stream>>>(e->p, e->s, e->art, state);
  else gx_get_state_kernel<<<div_up(e->p.n, 128), 128, 0, (cudaStream_t)stream>>>(e->p, e->s, state);
  GX_LAUNCH_CHECK("gx_get_state_kernel");
  return GX_OK;
}

GX_API int gx_set_state(gx_env* e, const float* state, void* stream) {
  if (!e || !state) return gx_fail(GX_EINVAL, "gx_set_state: NULL argument");
  DeviceGuard guard(e->device);
  if (e->tree_host) gx_tree_set_state(e->p, e->s, e->tree, state, (cudaStream_t)stream);
  else if (e->art_host) gx_art_set_state_kernel<<<div_up(e->p.n, 128), 128, 0, (cudaStream_t)stream>>>(e->p, e->s, e->art, state);
  else gx_set_state_kernel<<<div_up(e->p.n, 128), 128, 0, (cudaStream_t)stream>>>(e->p, e->s, state);
  GX_LAUNCH_CHECK("gx_set_state_kernel");
  e->have_reset = true;
  return GX_OK;
}

__global__ void gx_set_age_kernel(const GxState s, int n, unsigned long long age) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 c = s.d2[i];
  c.z = __uint_as_float((uint32_t)age);
  c.w = __uint_as_float((uint32_t)(age >> 32));
  s.d2[i] = c;
}

GX_API int gx_set_counters(gx_env* e, uint64_t t_global, uint64_t epoch, void* stream) {
  if (!e) return gx_fail(GX_EINVAL, "gx_set_counters: NULL env");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned long long h[4] = {0ull, epoch, 0ull, 0ull};
  GX_CUDA(cudaMemcpyAsync(e->ctr, h, sizeof h, cudaMemcpyHostToDevice, st));
  gx_set_age_kernel<<<div_up(e->p.n, 128), 128, 0, st>>>(e->s, e->p.n, t_global);
  GX_LAUNCH_CHECK("gx_set_age_kernel");
  GX_CUDA(cudaStreamSynchronize(st));  // h is on this stack frame
  return GX_OK;
}

GX_API int gx_get_counters(gx_env* e, uint64_t* h_t_global, uint64_t* h_epoch, void* stream) {
  if (!e) return gx_fail(GX_EINVAL, "gx_get_counters: NULL env");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  unsigned long long h[4];
  float4 d2;
  GX_CUDA(cudaMemcpyAsync(h, e->ctr, sizeof h, cudaMemcpyDeviceToHost, st));
  GX_CUDA(cudaMemcpyAsync(&d2, e->s.d2, sizeof d2, cudaMemcpyDeviceToHost, st));  // env 0
  GX_CUDA(cudaStreamSynchronize(st));
  uint32_t lo, hi;
  memcpy(&lo, &d2.z, 4);
  memcpy(&hi, &d2.w, 4);
  if (h_t_global) *h_t_global = (uint64_t)lo | ((uint64_t)hi << 32);
  if (h_epoch) *h_epoch = h[GX_CTR_EPOCH];
  return GX_OK;
}

GX_API int gx_pool_read(gx_env* e, float* pool, int32_t* tries, void* stream) {
  if (!e) return gx_fail(GX_EINVAL, "gx_pool_read: NULL env");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (pool) GX_CUDA(cudaMemcpyAsync(pool, e->pool, sizeof(float) * (size_t)e->p.pool_size * e->p.lay_dim, cudaMemcpyDeviceToDevice, st));
  if (tries) GX_CUDA(cudaMemcpyAsync(tries, e->pool_tries, sizeof(int32_t) * (size_t)e->p.pool_size, cudaMemcpyDeviceToDevice, st));
  return GX_OK;
}

GX_API int gx_rollout(gx_env* e, int32_t T, float* obs, float* act, float* reward, float* cost, float* done,
                      void* stream) {
  if (!e || T < 0) return gx_fail(GX_EINVAL, "gx_rollout: bad argument");
  if (!e->have_reset) return gx_fail(GX_ESTATE, "gx_rollout: call gx_reset first");
  if (T == 0) return GX_OK;
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (e->art_host || e->tree_host) {
    // one fused step + auto-reset launch per control step, actions drawn in-kernel (the
    // per-step work of a legged robot dwarfs the launch, unlike the Point's)
    if (!obs || !reward || !cost || !done) return gx_fail(GX_EINVAL, "gx_rollout: articulated robots need obs, reward, cost and done buffers");
    const size_t n = (size_t)e->p.n;
    for (int t = 0; t < T; ++t)
      if (int rc = gx_art_step(e, st, nullptr, obs + (size_t)t * n * e->p.obs_dim, act ? act + (size_t)t * n * gx_act_dim(e) : nullptr,
                               reward + t * n, done + t * n, cost + t * n, nullptr, true))
        return rc;
    return GX_OK;
  }
  const int block = pick_block(e), grid = div_up(e->p.n, block);
  const size_t sh = tile_bytes(e, block);
  float2* a2 = reinterpret_cast<float2*>(act);
  if (lanes_path(e)) {
    const int lb = 128, lg = div_up(e->p.n * GX_LPE, lb);
    if (e->stats_on) gx_rollout_lanes_kernel<true><<<lg, lb, 0, st>>>(e->p, e->s, e->pool, T, obs, a2, reward, cost, done, e->stats);
    else gx_rollout_lanes_kernel<false><<<lg, lb, 0, st>>>(e->p, e->s, e->pool, T, obs, a2, reward, cost, done, e->stats);
    GX_LAUNCH_CHECK("gx_rollout_lanes_kernel");
    return GX_OK;
  }
#define GX_RO(NH, ST) gx_rollout_kernel<NH, ST><<<grid, block, sh, st>>>(e->p, e->s, e->pool, T, obs, a2, reward, cost, done, e->stats)
  if (fast_path(e)) { if (e->stats_on) GX_RO(8, true); else GX_RO(8, false); }
  else { if (e->stats_on) GX_RO(0, true); else GX_RO(0, false); }
#undef GX_RO
  GX_LAUNCH_CHECK("gx_rollout_kernel");
  return GX_OK;
}

GX_API int gx_stats_read(gx_env* e, double* out8, void* stream) {
  if (!e || !out8) return gx_fail(GX_EINVAL, "gx_stats_read: NULL argument");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  gx_stats_reduce_kernel<<<1, 1024, 0, st>>>(e->s, e->p.n, e->stats_age0, e->stats);
  GX_LAUNCH_CHECK("gx_stats_reduce_kernel");
  GX_CUDA(cudaMemcpyAsync(out8, e->stats, sizeof(double) * 8, cudaMemcpyDefault, st));
  return GX_OK;
}

GX_API int gx_stats_reset(gx_env* e, void* stream) {
  if (!e) return gx_fail(GX_EINVAL, "gx_stats_reset: NULL env");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  e->stats_on = true;  // the first reset switches accumulation on
  GX_CUDA(cudaMemsetAsync(e->s.ep, 0, sizeof(float4) * e->s.cap, st));
  GX_CUDA(cudaMemsetAsync(e->s.fin, 0, sizeof(float4) * 2 * (size_t)e->s.cap, st));
  // total_env_steps counts from the env age at this moment (host sync: rare call)
  uint64_t t = 0;
  if (int rc = gx_get_counters(e, &t, nullptr, stream)) return rc;
  e->stats_age0 = t;
  return GX_OK;
}

// true if the kernels can address this host pointer directly (pinned + mapped: cudaHostAlloc /
// cudaHostRegister / torch pin_memory under unified addressing)
static bool gx_host_mapped(const void* p, const void** dev) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return false;
  *dev = at.devicePointer;
  return true;
}

GX_API int gx_step_host(gx_env* e, const float* h_action, float* h_obs, float* h_reward, float* h_done,
                        float* h_cost, int32_t auto_reset_done, void* stream) {
  if (!e || !h_action || !h_obs || !h_reward || !h_done || !h_cost) return gx_fail(GX_EINVAL, "gx_step_host: NULL argument");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)e->p.n, od = (size_t)e->p.obs_dim;
  // Zero-copy route (pinned, mapped buffers): ONE launch, the kernel reads the actions from and
  // writes obs | reward | done | cost to host memory itself -- the PCIe writes overlap the
  // computation and the two copy-engine hand-offs disappear (4096 envs: 49.9 -> 40.6 us per step).
  // All five pointers are looked up on EVERY call (a few 100 ns each): a cached mapping would go
  // stale when the caller frees a pinned buffer and a pageable one lands at the same address.
  {
    const void *da, *dob, *dr, *dd, *dc;
    e->zc_ok = gx_host_mapped(h_action, &da) && gx_host_mapped(h_obs, &dob) && gx_host_mapped(h_reward, &dr) &&
               gx_host_mapped(h_done, &dd) && gx_host_mapped(h_cost, &dc) &&
               !(reinterpret_cast<uintptr_t>(da) & 7u);
    if (e->zc_ok) { e->zc_dact = (const float*)da; e->zc_dobs = (float*)dob; e->zc_dr = (float*)dr; e->zc_dd = (float*)dd; e->zc_dc = (float*)dc; }
  }
  if (e->zc_ok && !getenv("GX_NO_ZERO_COPY")) {
    if (int rc = gx_step_impl(e, e->zc_dact, e->zc_dobs, e->zc_dr, e->zc_dd, e->zc_dc, nullptr, nullptr, stream,
                              auto_reset_done != 0, "gx_step_host", true))
      return rc;
    GX_CUDA(cudaStreamSynchronize(st));
    return GX_OK;
  }
  // copy route (pageable or unmapped host memory)
  // stage = [obs n*od | reward n | done n | cost n | pad | action ad*n]
  const size_t ad = (size_t)gx_act_dim(e);
  const size_t act_off = ((od + 3) * n + 3) & ~(size_t)3;
  if (!e->stage) GX_CUDA(cudaMalloc((void**)&e->stage, sizeof(float) * (act_off + ad * n)));
  float* d_obs = e->stage;
  float* d_reward = d_obs + od * n;
  float* d_done = d_reward + n;
  float* d_cost = d_done + n;
  float* d_act = e->stage + act_off;
  GX_CUDA(cudaMemcpyAsync(d_act, h_action, sizeof(float) * ad * n, cudaMemcpyHostToDevice, st));
  if (int rc = gx_step_impl(e, d_act, d_obs, d_reward, d_done, d_cost, nullptr, nullptr, stream,
                            auto_reset_done != 0, "gx_step_host"))
    return rc;
  if (h_reward == h_obs + od * n && h_done == h_reward + n && h_cost == h_done + n) {
    GX_CUDA(cudaMemcpyAsync(h_obs, d_obs, sizeof(float) * (od + 3) * n, cudaMemcpyDeviceToHost, st));
  } else {
    GX_CUDA(cudaMemcpyAsync(h_obs, d_obs, sizeof(float) * od * n, cudaMemcpyDeviceToHost, st));
    GX_CUDA(cudaMemcpyAsync(h_reward, d_reward, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    GX_CUDA(cudaMemcpyAsync(h_done, d_done, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
    GX_CUDA(cudaMemcpyAsync(h_cost, d_cost, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  }
  GX_CUDA(cudaStreamSynchronize(st));
  return GX_OK;
}

// ---- host-buffer rollout, pipelined ------------------------------------------
// T steps with host-resident actions and host-resident outputs.  The steps are
// cut into chunks of S; chunk c's actions go up and its S fused step+auto-reset
// launches run on a compute stream while chunk c-1's outputs drain to the host on
// a copy stream (two device chunk buffers).  PCIe carries 8 B up and
// (obs_dim+3)*4 B down per env-step; at 4096 envs the kernels (~5 us/step) hide
// entirely behind the D2H stream (~14 us/step at 55 GB/s).
GX_API int gx_rollout_host(gx_env* e, int32_t T, const float* h_action, float* h_obs, float* h_reward,
                           float* h_done, float* h_cost, void* stream) {
  if (!e || T < 0 || !h_action || !h_obs || !h_reward || !h_done || !h_cost)
    return gx_fail(GX_EINVAL, "gx_rollout_host: bad argument");
  if (!e->have_reset) return gx_fail(GX_ESTATE, "gx_rollout_host: call gx_reset first");
  if (T == 0) return GX_OK;
  DeviceGuard guard(e->device);
  cudaStream_t user = (cudaStream_t)stream;
  const size_t n = (size_t)e->p.n, od = (size_t)e->p.obs_dim;
  const size_t ad = (size_t)gx_act_dim(e);
  const size_t per_step = (od + 3 + ad) * n;  // obs | reward | done | cost | action
  // ---- route ---------------------------------------------------------------------------------------------------
  const void *zo = nullptr, *zr = nullptr, *zd = nullptr, *zc = nullptr;
  const bool can_zc = !e->art_host && !e->tree_host && gx_host_mapped(h_obs, &zo) && gx_host_mapped(h_reward, &zr) &&
                      gx_host_mapped(h_done, &zd) && gx_host_mapped(h_cost, &zc);
  int route = 0;
  if (can_zc) {
    const char* force = getenv("GX_ROLLOUT_HOST_ROUTE");   // "dma" | "zc" | unset = measure
    if (force && force[0] == 'z') route = 1;
    else if (force && force[0] == 'd') route = 0;
    else if (e->rh_calls < 4) route = e->rh_calls & 1;      // probe: dma, zc, dma, zc
    else route = e->rh_route;
  }
  const auto t_begin = std::chrono::steady_clock::now();
  auto finish = [&](int rc_) {
    if (rc_ == GX_OK && can_zc) {
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count();
      const double rate = (double)n * T / (dt > 1e-9 ? dt : 1e-9);
      if (e->rh_calls >= 2 && e->rh_calls < 4) e->rh_rate[route] = rate;   // the second probe of each route (first ones allocate)
      if (e->rh_calls == 3) e->rh_route = e->rh_rate[1] > e->rh_rate[0] ? 1 : 0;
      e->rh_calls++;
    }
    return rc_;
  };
  if (!e->ring_steps) {
    // ~8 MB of output per chunk keeps the copy engine near its large-transfer rate
    int S = (int)((8u << 20) / (sizeof(float) * (od + 3) * n));
    S = S < 1 ? 1 : (S > 64 ? 64 : S);
    // build everything in locals and commit to the handle only when all of it exists
    float* ring[2] = {nullptr, nullptr};
    cudaStream_t sc = nullptr, sy = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaError_t err = cudaSuccess;
    for (int k = 0; k < 2 && err == cudaSuccess; ++k) err = cudaMalloc((void**)&ring[k], sizeof(float) * per_step * S);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&sy, cudaStreamNonBlocking);
    for (int k = 0; k < 5 && err == cudaSuccess; ++k) err = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
    if (err != cudaSuccess) {
      for (int k = 0; k < 5; ++k) if (ev[k]) cudaEventDestroy(ev[k]);
      if (sc) cudaStreamDestroy(sc);
      if (sy) cudaStreamDestroy(sy);
      for (int k = 0; k < 2; ++k) if (ring[k]) cudaFree(ring[k]);
      cudaGetLastError();
      return gx_fail(GX_ECUDA, "gx_rollout_host: pipeline setup failed: %s", cudaGetErrorString(err));
    }
    e->ring[0] = ring[0]; e->ring[1] = ring[1];
    e->rs_compute = sc; e->rs_copy = sy;
    e->ev_in = ev[0]; e->ev_computed[0] = ev[1]; e->ev_computed[1] = ev[2]; e->ev_copied[0] = ev[3]; e->ev_copied[1] = ev[4];
    e->ring_steps = S;
  }
  const int S = e->ring_steps;
  // any failure below drains both pipeline streams before returning, so that no D2H copy into the
  // caller's buffers is still in flight when the caller sees the error
  int rc = GX_OK;
  auto cu = [&](cudaError_t err, const char* what) {
    if (rc == GX_OK && err != cudaSuccess) rc = gx_fail(GX_ECUDA, "gx_rollout_host: %s failed: %s", what, cudaGetErrorString(err));
    return rc == GX_OK;
  };
  if (route == 1) {
    // zero-copy: the actions of a chunk go up by DMA (8 B per env-step), the step kernels write observation tiles,
    // reward, done and cost straight into the caller's mapped host buffers (coalesced 128-bit stores over PCIe);
    // no device-to-host copy, no second stream
    float* d_obs = (float*)zo; float* d_rew = (float*)zr; float* d_done = (float*)zd; float* d_cost = (float*)zc;
    if (cu(cudaEventRecord(e->ev_in, user), "cudaEventRecord")) cu(cudaStreamWaitEvent(e->rs_compute, e->ev_in, 0), "cudaStreamWaitEvent");
    int chunk = 0;
    for (int t0 = 0; t0 < T && rc == GX_OK; t0 += S, ++chunk) {
      const int kk = chunk & 1, s_now = (T - t0 < S) ? (T - t0) : S;
      float* d_act = e->ring[kk];          // the ring buffers only stage actions on this route
      // (the H2D copy into ring[kk] may overwrite actions that chunk - 2 has consumed: same stream, in order)
      if (!cu(cudaMemcpyAsync(d_act, h_action + ad * n * t0, sizeof(float) * ad * n * s_now, cudaMemcpyHostToDevice, e->rs_compute), "H2D copy")) break;
      for (int t = 0; t < s_now && rc == GX_OK; ++t)
        rc = gx_step_impl(e, d_act + ad * n * t, d_obs + od * n * (t0 + t), d_rew + n * (t0 + t), d_done + n * (t0 + t), d_cost + n * (t0 + t),
                          nullptr, nullptr, e->rs_compute, true, "gx_rollout_host", true);
    }
    if (rc == GX_OK) cu(cudaEventRecord(e->ev_computed[0], e->rs_compute), "cudaEventRecord");
    if (rc == GX_OK) cu(cudaStreamWaitEvent(user, e->ev_computed[0], 0), "cudaStreamWaitEvent");
    cudaError_t serr = cudaStreamSynchronize(e->rs_compute);
    if (rc == GX_OK && serr != cudaSuccess) rc = gx_fail(GX_ECUDA, "gx_rollout_host: %s", cudaGetErrorString(serr));
    return finish(rc);
  }
  int chunk = 0;
  if (cu(cudaEventRecord(e->ev_in, user), "cudaEventRecord")) cu(cudaStreamWaitEvent(e->rs_compute, e->ev_in, 0), "cudaStreamWaitEvent");
  for (int t0 = 0; t0 < T && rc == GX_OK; t0 += S, ++chunk) {
    const int k = chunk & 1, s_now = (T - t0 < S) ? (T - t0) : S;
    float* d_obs = e->ring[k];
    float* d_rew = d_obs + od * n * S;
    float* d_done = d_rew + n * S;
    float* d_cost = d_done + n * S;
    float* d_act = d_cost + n * S;
    if (chunk >= 2 && !cu(cudaStreamWaitEvent(e->rs_compute, e->ev_copied[k], 0), "cudaStreamWaitEvent")) break;
    if (!cu(cudaMemcpyAsync(d_act, h_action + ad * n * t0, sizeof(float) * ad * n * s_now, cudaMemcpyHostToDevice, e->rs_compute), "H2D copy")) break;
    for (int t = 0; t < s_now && rc == GX_OK; ++t)
      rc = gx_step_impl(e, d_act + ad * n * t, d_obs + od * n * t, d_rew + n * t, d_done + n * t, d_cost + n * t,
                        nullptr, nullptr, e->rs_compute, true, "gx_rollout_host");
    if (rc != GX_OK) break;
    if (!cu(cudaEventRecord(e->ev_computed[k], e->rs_compute), "cudaEventRecord")) break;
    if (!cu(cudaStreamWaitEvent(e->rs_copy, e->ev_computed[k], 0), "cudaStreamWaitEvent")) break;
    cu(cudaMemcpyAsync(h_obs + od * n * t0, d_obs, sizeof(float) * od * n * s_now, cudaMemcpyDeviceToHost, e->rs_copy), "D2H copy");
    cu(cudaMemcpyAsync(h_reward + n * t0, d_rew, sizeof(float) * n * s_now, cudaMemcpyDeviceToHost, e->rs_copy), "D2H copy");
    cu(cudaMemcpyAsync(h_done + n * t0, d_done, sizeof(float) * n * s_now, cudaMemcpyDeviceToHost, e->rs_copy), "D2H copy");
    cu(cudaMemcpyAsync(h_cost + n * t0, d_cost, sizeof(float) * n * s_now, cudaMemcpyDeviceToHost, e->rs_copy), "D2H copy");
    if (!cu(cudaEventRecord(e->ev_copied[k], e->rs_copy), "cudaEventRecord")) break;
  }
  if (rc != GX_OK) {
    cudaStreamSynchronize(e->rs_compute);
    cudaStreamSynchronize(e->rs_copy);
    cudaGetLastError();
    return rc;
  }
  // the caller's stream (and the host) see the finished rollout
  GX_CUDA(cudaStreamWaitEvent(user, e->ev_copied[(chunk - 1) & 1], 0));
  GX_CUDA(cudaStreamSynchronize(e->rs_copy));
  return finish(GX_OK);
}

GX_API int gx_rollout_host_route(const gx_env* e, double* rates2) {
  if (!e) return GX_EINVAL;
  if (rates2) { rates2[0] = e->rh_rate[0]; rates2[1] = e->rh_rate[1]; }
  return e->rh_calls < 4 ? -1 : e->rh_route;
}

// ---- policy half of a PPO rollout step (gx_policy.cuh) ---------------------------------
template <int K1>
static int gx_policy_launch(gx_env* e, cudaStream_t st, const float* W, const float* obs, uint64_t noise_seed, int A, float* act,
                            float* val, float* cval, float* logp, float* mu_out) {
  const size_t sh = gx_policy_smem_floats(K1) * sizeof(float);
  // (function attributes are per device: set on every call, a process may hold handles on several devices)
  GX_CUDA(cudaFuncSetAttribute(gx_policy_kernel<K1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  const int grid = div_up(e->p.n, GX_POL_ENVS);
  // No programmatic dependent launch on either side of this kernel (measured, 4096 envs: with it 15.8 us per rollout step, with
  // only the env kernel released early 16.1, without 15.1 -- blocks of the other kernel waiting on the SMs cost more
  // than the launch gap they hide): a plain launch, and the kernel never triggers its dependents early.
  gx_policy_kernel<K1><<<grid, GX_POL_THREADS, sh, st>>>(W, obs, e->p.n, e->p.obs_dim, A, e->s.d2, e->p.env_id_offset,
                                                         e->p.seed_lo ^ (uint32_t)noise_seed, e->p.seed_hi ^ (uint32_t)(noise_seed >> 32),
                                                         act, val, cval, logp, mu_out);
  GX_LAUNCH_CHECK("gx_policy_kernel");
  return GX_OK;
}

GX_API int64_t gx_policy_weight_floats(const gx_env* e) {
  return e ? (int64_t)gx_policy_weight_count(e->p.obs_dim, gx_act_dim(e)) : (int64_t)GX_EINVAL;
}

GX_API int gx_policy_step(gx_env* e, const float* weights, const float* obs, uint64_t noise_seed, float* act, float* val,
                          float* cval, float* logp, float* mu_out, void* stream) {
  if (!e || !weights || !obs || !act || !val || !cval || !logp) return gx_fail(GX_EINVAL, "gx_policy_step: NULL argument");
  if (!e->have_reset) return gx_fail(GX_ESTATE, "gx_policy_step: call gx_reset first");
  const int A = gx_act_dim(e);
  if (A > GX_POL_MAX_ACT) return gx_fail(GX_EINVAL, "gx_policy_step: act_dim %d > %d", A, GX_POL_MAX_ACT);
  if (reinterpret_cast<uintptr_t>(weights) & 15u) return gx_fail(GX_EINVAL, "gx_policy_step: weights must be 16-byte aligned (bulk-async copies)");
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  switch ((e->p.obs_dim + 3) & ~3) {   // Point 43, Swimmer 46, Ant 64, Walker 70, Doggo 85 (+ 2 / 4 with observe_vel / observe_acc)
#define GX_POL_CASE(K) case K: return gx_policy_launch<K>(e, st, weights, obs, noise_seed, A, act, val, cval, logp, mu_out)
    GX_POL_CASE(44); GX_POL_CASE(48); GX_POL_CASE(52); GX_POL_CASE(64); GX_POL_CASE(68); GX_POL_CASE(72); GX_POL_CASE(76);
    GX_POL_CASE(88); GX_POL_CASE(92);
#undef GX_POL_CASE
    default: return gx_fail(GX_EINVAL, "gx_policy_step: no kernel instantiated for obs_dim %d", e->p.obs_dim);
  }
}

// ---- a whole PPO rollout as one persistent launch (gx_policy.cuh: gx_policy_rollout_kernel) ----------------------------
static bool lanes_shape(const gx_env* e) {
  const gx_params& p = e->p;
  return !e->art_host && !e->tree_host && p.nh == 8 && p.nb == 16 && p.obs_dim == 43 && p.alias && p.use_exp && p.ctrllimited &&
         p.forcelimited && p.substeps == 1 && p.off_ctrl == 0 && p.off_compass == 3 && p.off_glidar == 5 && p.off_hlidar == 21 &&
         p.off_qpos == 37 && p.off_qvel == 40;
}

GX_API int gx_policy_rollout_supported(const gx_env* e) { return e && lanes_shape(e) && gx_act_dim(e) == 2 ? 1 : 0; }

GX_API int gx_policy_rollout(gx_env* e, const float* weights, int32_t T, uint64_t noise_seed, float* obs, float* act, float* reward,
                             float* done, float* cost, float* val, float* cval, float* logp, void* stream) {
  if (!e || !weights || !obs || !act || !reward || !done || !cost || !val || !cval || !logp)
    return gx_fail(GX_EINVAL, "gx_policy_rollout: NULL argument");
  if (!e->have_reset) return gx_fail(GX_ESTATE, "gx_policy_rollout: call gx_reset first");
  if (T < 0) return gx_fail(GX_EINVAL, "gx_policy_rollout: T = %d", T);
  if (!gx_policy_rollout_supported(e))
    return gx_fail(GX_EINVAL, "gx_policy_rollout: built for the default Point shape (obs 43, act 2); use gx_policy_step + gx_step_autoreset");
  if (reinterpret_cast<uintptr_t>(weights) & 15u) return gx_fail(GX_EINVAL, "gx_policy_rollout: weights must be 16-byte aligned (bulk-async copies)");
  if (reinterpret_cast<uintptr_t>(act) & 7u) return gx_fail(GX_EINVAL, "gx_policy_rollout: act must be 8-byte aligned");
  if (T == 0) return GX_OK;
  DeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sh = gx_policy_rollout_smem_floats() * sizeof(float);
  if (e->stats_on) GX_CUDA(cudaFuncSetAttribute(gx_policy_rollout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  else GX_CUDA(cudaFuncSetAttribute(gx_policy_rollout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  const int grid = div_up(e->p.n, GX_PR_ENVS);
  float2* a2 = reinterpret_cast<float2*>(act);
  const unsigned k0 = (uint32_t)noise_seed, k1 = (uint32_t)(noise_seed >> 32);
  if (e->stats_on)
    gx_policy_rollout_kernel<true><<<grid, GX_PR_THREADS, sh, st>>>(e->p, e->s, e->pool, weights, T, k0, k1, obs, a2, reward, done, cost, val, cval, logp);
  else
    gx_policy_rollout_kernel<false><<<grid, GX_PR_THREADS, sh, st>>>(e->p, e->s, e->pool, weights, T, k0, k1, obs, a2, reward, done, cost, val, cval, logp);
  GX_LAUNCH_CHECK("gx_policy_rollout_kernel");
  return GX_OK;
}

// ---- diagnostics -------------------------------------------------------------
__global__ void gx_debug_math_kernel(int fn, const float* a, const float* b, float* out, float* out2, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (fn == 0) {
    float s, c;
    gx_sincos(a[i], &s, &c);
    out[i] = s;
    if (out2) out2[i] = c;
  } else if (fn == 1) {
    out[i] = gx_atan2(a[i], b[i]);
  } else if (fn == 2) {
    out[i] = gx_exp(a[i]);
  } else if (fn == 3) {
    out[i] = gx_max(a[i], b[i]);
  } else {
    out[i] = gx_min(a[i], b[i]);
  }
}

GX_API int gx_debug_math(int fn, const float* a, const float* b, float* out, float* out2, int n, void* stream) {
  if (!a || !out || n < 0 || fn < 0 || fn > 4 || (fn != 0 && fn != 2 && !b)) return gx_fail(GX_EINVAL, "gx_debug_math: bad argument");
  if (n == 0) return GX_OK;
  gx_debug_math_kernel<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(fn, a, b, out, out2, n);
  GX_LAUNCH_CHECK("gx_debug_math_kernel");
  return GX_OK;
}

// SM clock as the kernels see it: one thread spins for `spin_cycles` SM cycles and reports
// cycles / globaltimer-nanoseconds (MHz).  NVML samples at a few Hz; this probe can sit
// between two graph launches on the stream (bench.py uses it to report the clock the timed
// steps actually ran at).
__global__ void gx_debug_sm_clock_kernel(long long spin_cycles, float* out_mhz) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  const long long c0 = clock64();
  long long c1 = c0;
  while (c1 - c0 < spin_cycles) c1 = clock64();
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  *out_mhz = t1 > t0 ? (float)((double)(c1 - c0) * 1000.0 / (double)(t1 - t0)) : 0.0f;
}

GX_API int gx_debug_sm_clock(float* out_mhz, int32_t spin_cycles, void* stream) {
  if (!out_mhz || spin_cycles <= 0) return gx_fail(GX_EINVAL, "gx_debug_sm_clock: bad argument");
  gx_debug_sm_clock_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long)spin_cycles, out_mhz);
  GX_LAUNCH_CHECK("gx_debug_sm_clock_kernel");
  return GX_OK;
}

// ---- GAE-lambda / rewards-to-go (ppo_train.py:61-113) --------------------------
// One thread per env walks the T steps backwards; consecutive threads read
// consecutive envs of the time-major buffers, so every access is coalesced.
// 12 B read + 8 B written per (t, env): a pure HBM stream.
__global__ void gx_gae_kernel(const float* __restrict__ reward, const float* __restrict__ value,
                              const float* __restrict__ done, const float* __restrict__ last_value, int T,
                              int n, float gamma_f, double gamma, double gl, float* __restrict__ adv,
                              float* __restrict__ ret) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v_next = last_value ? last_value[i] : 0.0f;
  double a = 0.0, r = (double)v_next;  // discount_cumsum(rews + [last_val])[:-1]
  for (int t = T - 1; t >= 0; --t) {
    const size_t k = (size_t)t * n + i;
    const float rew = reward[k], v = value[k];
    const bool closed = done[k] > 0.0f;
    if (closed) { v_next = 0.0f; a = 0.0; r = 0.0; }
    const float delta = __fsub_rn(__fadd_rn(rew, __fmul_rn(gamma_f, v_next)), v);
    a = __dadd_rn((double)delta, __dmul_rn(gl, a));
    r = __dadd_rn((double)rew, __dmul_rn(gamma, r));
    adv[k] = (float)a;
    ret[k] = (float)r;
    v_next = v;
  }
}

GX_API int gx_gae(const float* reward, const float* value, const float* done, const float* last_value,
                  int32_t T, int32_t n, double gamma, double lam, float* adv, float* ret, void* stream) {
  if (!reward || !value || !done || !adv || !ret || T < 0 || n < 0) return gx_fail(GX_EINVAL, "gx_gae: bad argument");
  if (T == 0 || n == 0) return GX_OK;
  gx_gae_kernel<<<div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(reward, value, done, last_value, T, n,
                                                                  (float)gamma, gamma, gamma * lam, adv, ret);
  GX_LAUNCH_CHECK("gx_gae_kernel");
  return GX_OK;
}

// ---- done-compaction: indices of the envs whose done flag is set, ascending,
// plus their count -- what callers of the reference obtain on the host with
// np.where(done == 1) after a device->host copy (ppo_train.py:403-411, scpo/scpo.py:722).
// One block: ballot + popc per warp, a 32-entry scan per 1024-env chunk; deterministic order.
__global__ void __launch_bounds__(1024) gx_compact_done_kernel(const float* __restrict__ done, int n,
                                                               int32_t* __restrict__ idx, int32_t* __restrict__ count) {
  __shared__ int warp_excl[32];
  __shared__ int chunk_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = 0;  // envs found in earlier chunks (same value in every thread)
  for (int start = 0; start < n; start += 1024) {
    const int i = start + threadIdx.x;
    const bool on = i < n && done[i] > 0.0f;
    const unsigned m = __ballot_sync(0xffffffffu, on);
    if (lane == 0) warp_excl[warp] = __popc(m);
    __syncthreads();
    if (warp == 0) {
      const int v = warp_excl[lane];
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      warp_excl[lane] = incl - v;
      if (lane == 31) chunk_total = incl;
    }
    __syncthreads();
    if (on) idx[base + warp_excl[warp] + __popc(m & ((1u << lane) - 1u))] = i;
    base += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *count = base;
}

GX_API int gx_compact_done(const float* done, int32_t n, int32_t* idx, int32_t* count, void* stream) {
  if (!done || !idx || !count || n < 0) return gx_fail(GX_EINVAL, "gx_compact_done: bad argument");
  gx_compact_done_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(done, n, idx, count);
  GX_LAUNCH_CHECK("gx_compact_done_kernel");
  return GX_OK;
}
