// gx_task.cu -- kernels + C ABI of the config-3 extension (include/guardx_task.h, gx_task.cuh).
// NO REFERENCE, PARITY UNPINNED: see the header of include/guardx_task.h.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "gx_task.cuh"

extern "C" const char* gx_last_error(void);
int gx_set_error(int code, const char* fmt, ...);   // guardx.cu

#define GXT_API extern "C" __attribute__((visibility("default")))
#define GXT_CUDA(call)                                                                                          \
  do {                                                                                                          \
    cudaError_t e_ = (call);                                                                                    \
    if (e_ != cudaSuccess) return gx_set_error(GX_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

struct gxt_env {
  gxt_config cfg;
  gxt_params p;
  int device, sm_count;
  float* S;            // [state_dim][cap]
  float* pool;         // [pool_size][lay_dim]
  int32_t* pool_tries;
  unsigned long long* epoch;   // device: layout pools generated so far
  bool have_reset;
};

struct GxtDeviceGuard {
  int prev = -1;
  bool sw = false;
  explicit GxtDeviceGuard(int dev) { if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) { cudaSetDevice(dev); sw = true; } }
  ~GxtDeviceGuard() { if (sw) cudaSetDevice(prev); }
};

// ---------------------------------------------------------------------------------------------------------------
// layout pool: one WARP per slot, the 32 lanes test 32 consecutive candidates of the slot's counter-based stream, the
// lowest successful one wins (same scheme as gx_pool_kernel of the Point path).  A candidate places the objects in class
// order, each with `tries` uniform draws inside the extents shrunk by its keep-out, the LAST draw that keeps the
// keep-out distance to everything placed so far wins (engine.py:546-572); the robot is placed last and must be at least
// goal_min_robot_dist from the goal (goal task; engine.py:571).  A gremlin also draws its phase.
// ---------------------------------------------------------------------------------------------------------------
#define GXT_MAX_POOL_ATTEMPTS (1u << 16)

__device__ bool gxt_sample_layout(const gxt_params& p, uint32_t k0, uint32_t k1, uint32_t slot, uint32_t attempt, float* out /* lay_dim */) {
  const int ntot = p.nobj + 1;
#pragma unroll 1
  for (int o = 0; o < ntot; ++o) {
    const int cls = o < p.nobj ? gxt_class_of_obj(p, o) : GXT_ROBOT;
    bool placed = false;
    float x = 0.0f, y = 0.0f;
#pragma unroll 1
    for (int t = 0; t < p.tries; ++t) {
      const gx_u32x4 r = gx_philox4x32_10(slot, attempt, ((uint32_t)o << 8) | (uint32_t)t, GX_STREAM_LAYOUT, k0, k1);
      const float cx = gx_uniform(r.v[0], p.lo[cls][0], p.hi[cls][0]);
      const float cy = gx_uniform(r.v[1], p.lo[cls][1], p.hi[cls][1]);
      bool valid = true;
      for (int q = 0; q < o; ++q) {
        const int qc = gxt_class_of_obj(p, q);
        const float dx = cx - out[2 * q], dy = cy - out[2 * q + 1];
        if (gx_sqrt(dx * dx + dy * dy) < p.keep[qc] + p.margin + p.keep[cls]) valid = false;
      }
      if (valid) { x = cx; y = cy; placed = true; }
    }
    if (!placed) return false;
    out[2 * o] = x; out[2 * o + 1] = y;
  }
  if (p.count[GXT_GOAL] > 0) {
    const float dx = out[2 * p.nobj] - out[2 * p.first[GXT_GOAL]], dy = out[2 * p.nobj + 1] - out[2 * p.first[GXT_GOAL] + 1];
    if (gx_sqrt(dx * dx + dy * dy) < p.min_robot_goal) return false;
  }
  for (int g = 0; g < p.ng; ++g) {
    const gx_u32x4 r = gx_philox4x32_10(slot, attempt, 0xff00u | (uint32_t)g, GX_STREAM_LAYOUT, k0, k1);
    out[2 * p.nobj + 2 + g] = gx_uniform(r.v[0], 0.0f, GX_TWO_PI_F);
  }
  return true;
}

__global__ void __launch_bounds__(128) gxt_pool_kernel(const __grid_constant__ gxt_params p, const unsigned long long* __restrict__ epoch,
                                                       float* __restrict__ pool, int32_t* __restrict__ tries) {
  const int lane = threadIdx.x & 31;
  const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= p.pool_size) return;
  const unsigned long long ep = *epoch;
  const gx_u32x4 key = gx_philox4x32_10((uint32_t)ep, (uint32_t)(ep >> 32), 0x7461736bu, GX_STREAM_LAYOUT, p.pt.seed_lo, p.pt.seed_hi);
  float lay[2 * GXT_MAX_OBJECTS + 2 + GXT_MAX_PER_CLASS];
  for (uint32_t base = 0; base < GXT_MAX_POOL_ATTEMPTS; base += 32) {
    const bool ok = gxt_sample_layout(p, key.v[0], key.v[1], (uint32_t)slot, base + lane, lay);
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (m) {
      const int win = __ffs(m) - 1;
      if (lane == win) {
        for (int k = 0; k < p.lay_dim; ++k) pool[(size_t)slot * p.lay_dim + k] = lay[k];
        tries[slot] = (int32_t)(base + win + 1);
      }
      return;
    }
  }
  if (lane == 0) tries[slot] = -1;
}

__global__ void gxt_bump_epoch(unsigned long long* epoch) { *epoch += 1ull; }

// ---------------------------------------------------------------------------------------------------------------
extern __shared__ __align__(16) float gxt_smem[];

__device__ __forceinline__ void gxt_load_dyn(const gxt_params& p, const float* __restrict__ S, int i, GxDyn& d, float& prev_box, float& goal_button) {
  const size_t c = p.cap;
  d.qx = S[i]; d.qy = S[c + i]; d.qth = S[2 * c + i]; d.vx = S[3 * c + i]; d.vy = S[4 * c + i]; d.vw = S[5 * c + i];
  d.hc = S[6 * c + i]; d.hs = S[7 * c + i]; d.prev_dist = S[8 * c + i]; prev_box = S[9 * c + i];
  d.done = S[10 * c + i]; d.steps = S[11 * c + i];
  d.age = (uint64_t)__float_as_uint(S[12 * c + i]) | ((uint64_t)__float_as_uint(S[13 * c + i]) << 32);
  goal_button = S[14 * c + i];
}

__device__ __forceinline__ void gxt_store_dyn(const gxt_params& p, float* __restrict__ S, int i, const GxDyn& d, float prev_box) {
  const size_t c = p.cap;
  S[i] = d.qx; S[c + i] = d.qy; S[2 * c + i] = d.qth; S[3 * c + i] = d.vx; S[4 * c + i] = d.vy; S[5 * c + i] = d.vw;
  S[6 * c + i] = d.hc; S[7 * c + i] = d.hs; S[8 * c + i] = d.prev_dist; S[9 * c + i] = prev_box;
  S[10 * c + i] = d.done; S[11 * c + i] = d.steps;
  S[12 * c + i] = __uint_as_float((uint32_t)d.age); S[13 * c + i] = __uint_as_float((uint32_t)(d.age >> 32));
}

// observation of a freshly placed env: pose = layout, theta = 0, qvel = ctrl = 0, episode time 0
__device__ __forceinline__ GxtObs gxt_reset_row(const gxt_params& p, const float* S, int i, const GxDyn& d, float goal_button, float* row) {
  GxtGeom g;
  g.px = d.qx; g.py = d.qy;
  gx_kinematics(d.qth, g.r00, g.r10);
  g.time = 0.0f;
  return gxt_obs_row(p, S, i, g, d, goal_button, 0.0f, 0.0f, 0.0f, row);
}

__global__ void gxt_reset_kernel(const __grid_constant__ gxt_params p, float* __restrict__ S, const float* __restrict__ pool,
                                 const float* __restrict__ layout, float* __restrict__ obs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int wbase = i - lane;
  if (wbase >= p.n) return;
  float* tile = gxt_smem + (size_t)warp * 32 * p.obs_dim;
  float* row = tile + lane * p.obs_dim;
  if (i < p.n) {
    GxDyn d;
    float prev_box, goal_button;
    gxt_load_dyn(p, S, i, d, prev_box, goal_button);   // reset() leaves Engine._done and the key alone (engine.py:463)
    const float* lay = gxt_pick_layout(p, pool, layout, i, d.age, GX_STREAM_PICK ^ 1u);
    gxt_place(p, S, i, lay, d, goal_button, d.age);
    d.steps = 0.0f;
    const GxtObs o = gxt_reset_row(p, S, i, d, goal_button, row);
    d.prev_dist = o.dist_task;
    prev_box = o.dist_box;
    gx_kinematics(d.qth, d.hc, d.hs);
    gxt_store_dyn(p, S, i, d, prev_box);
  }
  gx_store_tile(tile, obs + (size_t)wbase * p.obs_dim, min(32, p.n - wbase), p.obs_dim, lane);
}

template <bool AUTORESET>
__global__ void __launch_bounds__(64) gxt_step_kernel(const __grid_constant__ gxt_params p, float* __restrict__ S, const float* __restrict__ pool,
                                                      const float* __restrict__ action, float* __restrict__ obs, float* __restrict__ reward,
                                                      float* __restrict__ done, float* __restrict__ cost, int32_t* __restrict__ aux) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int wbase = i - lane;
  if (wbase >= p.n) return;
  float* tile = gxt_smem + (size_t)warp * 32 * p.obs_dim;
  float* row = tile + lane * p.obs_dim;
  if (i < p.n) {
    GxDyn d;
    float prev_box, goal_button;
    gxt_load_dyn(p, S, i, d, prev_box, goal_button);
    const float a0 = action[(size_t)i * 2], a1 = action[(size_t)i * 2 + 1];
    const float last_dist = d.prev_dist, last_box = prev_box, last_done = d.done;
    GxtGeom g;
    g.time = d.steps * p.dt;
    float c0, c1, c2 = 0.0f;
    if (p.robot == GXT_ROBOT_CAR) {
      gxt_car_step(p, d, a0, a1, c0, c1, g.px, g.py, g.r00, g.r10);
    } else {
      c0 = d.hc * a0; c1 = d.hs * a0; c2 = a1;    // convert_action, lagged heading (engine.py:672-685)
      gx_mjx_step_point(p.pt, d, c0, c1, c2, g.px, g.py, g.r00, g.r10, nullptr);
    }
    // pillars stop the robot, the robot pushes vases and the box: applied to the POST-step state; the observation below
    // still sees every object where it was during the forward pass (its geometry is the pre-step snapshot)
    GxDyn dn = d;
    // (object positions are written by gxt_resolve_contacts; read them for the observation first)
    const GxtObs o = gxt_obs_row(p, S, i, g, d, goal_button, c0, c1, c2, row);
    gxt_resolve_contacts(p, S, i, dn);
    if (p.off_qpos >= 0) { row[p.off_qpos] = dn.qx; row[p.off_qpos + 1] = dn.qy; }
    if (p.off_qvel >= 0) { row[p.off_qvel] = dn.vx; row[p.off_qvel + 1] = dn.vy; }
    int bad = o.bad | gx_notfinite(dn.qx) | gx_notfinite(dn.qy) | gx_notfinite(dn.vx) | gx_notfinite(dn.vy);
    d = dn;
    // reward / done (engine.py:787-802 pattern; success = target reached)
    const float dist = o.dist_task;
    const float last = last_done > 0.0f ? dist : last_dist;
    const float d_dist = last - dist;
    float rew, dn_flag;
    int flags = 0;
    if (p.task == GXT_TASK_PUSH) {
      const float lastb = last_done > 0.0f ? o.dist_box : last_box;
      rew = (lastb - o.dist_box) * p.reward_box_dist + d_dist * p.reward_box_goal;
      dn_flag = dist < p.goal_size ? 1.0f : 0.0f;
    } else if (p.task == GXT_TASK_BUTTON) {
      rew = d_dist * p.reward_distance;
      dn_flag = dist < p.size[GXT_ROBOT] + p.size[GXT_BUTTON] ? 1.0f : 0.0f;
    } else {
      rew = d_dist * p.reward_distance;
      dn_flag = dist < p.goal_size ? 1.0f : 0.0f;
    }
    if (dn_flag > 0.0f) flags |= GX_AUX_GOAL;
    if (gx_abs(d_dist) > 1.0f) { dn_flag = 1.0f; rew = 0.0f; flags |= GX_AUX_JUMP; }
    if (bad) { rew = 0.0f; dn_flag = 1.0f; flags |= GX_AUX_GUARD; }
    if (d.steps > p.num_steps) { dn_flag = 1.0f; flags |= GX_AUX_TIMEOUT; }
    d.steps = dn_flag > 0.0f ? 0.0f : d.steps + 1.0f;
    d.done = dn_flag;
    d.prev_dist = dist;
    prev_box = o.dist_box;
    d.hc = g.r00; d.hs = g.r10;
    d.age += 1;
    if (AUTORESET && dn_flag > 0.0f) {
      const float* lay = gxt_pick_layout(p, pool, nullptr, i, d.age, GX_STREAM_PICK);
      gxt_place(p, S, i, lay, d, goal_button, d.age);
      gxt_reset_row(p, S, i, d, goal_button, row);
    }
    gxt_store_dyn(p, S, i, d, prev_box);
    reward[i] = rew; done[i] = dn_flag; cost[i] = o.cost;
    if (aux) { aux[4 * (size_t)i] = flags; aux[4 * (size_t)i + 1] = o.violations; aux[4 * (size_t)i + 2] = o.contacts; aux[4 * (size_t)i + 3] = o.goal_bin; }
  }
  gx_store_tile(tile, obs + (size_t)wbase * p.obs_dim, min(32, p.n - wbase), p.obs_dim, lane);
}

// the step-pose observation pass of an env in an exotic state (non-finite / astronomically large pose or objects), out of line
__device__ __noinline__ GxtObs gxt_obs_lanes_cold(const gxt_params& p, float* __restrict__ S, const float* ob, int i, int l, unsigned mask,
                                                  const GxtGeom& g, float goal_button, float* row, const GxDyn* push) {
  return gxt_obs_lanes(p, S, ob, i, l, mask, g, goal_button, row, push);
}

// ---- Engine.step (+ fused reset_done) with 8 lanes per env: 4 envs per warp, 16 per 128-thread block ----------------
// Seven resident blocks per SM (<= 72 registers): the 16 384-env launch of BASELINE.json configs[2] is 1024 blocks, one wave on
// 148 SMs only if seven fit -- at 80 registers it is 1.15 waves and 25 % slower (measured twice by accident).
template <bool AUTORESET>
__global__ void __launch_bounds__(128, 7) gxt_step_lanes_kernel(const __grid_constant__ gxt_params p, float* __restrict__ S,
                                                             const float* __restrict__ pool, const float* __restrict__ action,
                                                             float* __restrict__ obs, float* __restrict__ reward, float* __restrict__ done,
                                                             float* __restrict__ cost, int32_t* __restrict__ aux) {
  gx_pdl_sync();      // programmatic dependent launch (gx_device.cuh): before the first access to global memory
  const int l = threadIdx.x & (GXT_LPE - 1), lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int env = (blockIdx.x * blockDim.x + threadIdx.x) / GXT_LPE;
  const int wenv0 = (blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / GXT_LPE;     // first env of this warp
  if (wenv0 >= p.n) return;
  const bool live = env < p.n;
  const int i = live ? env : p.n - 1;
  const unsigned mask = gxt_grp_mask();
  float* tile = gxt_smem + (size_t)warp * (32 / GXT_LPE) * p.obs_dim;
  float* row = tile + (size_t)(lane / GXT_LPE) * p.obs_dim;
  const int ob_stride = 2 * p.nobj + p.ng;
  float* ob = gxt_smem + (size_t)(blockDim.x / GXT_LPE) * p.obs_dim + (size_t)(threadIdx.x / GXT_LPE) * ob_stride;
  const int staged_ok = gxt_stage_objects(p, S, i, l, ob);
  // zero the warp's whole observation tile with 128-bit stores (everything that is not a lidar range is overwritten below;
  // zeroing the seven lidars one by one was 9 % of the kernel's instructions); issued after the state loads
  {
    const int cnt = (32 / GXT_LPE) * p.obs_dim;      // the tile is 16-byte aligned: a warp's tile starts at a multiple of 4 rows
    float4* t4 = reinterpret_cast<float4*>(tile);
    for (int k = lane; k < (cnt >> 2); k += 32) t4[k] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    for (int k = (cnt & ~3) + lane; k < cnt; k += 32) tile[k] = 0.0f;
    __syncwarp();      // a row is zeroed by lanes of every env of the warp
  }
  GxDyn d;
  float prev_box, goal_button;
  gxt_load_dyn(p, S, i, d, prev_box, goal_button);
  const float a0 = action[(size_t)i * 2], a1 = action[(size_t)i * 2 + 1];
  const float last_dist = d.prev_dist, last_box = prev_box, last_done = d.done;
  GxtGeom g;
  g.time = d.steps * p.dt;
  float c0, c1, c2 = 0.0f;
  if (p.robot == GXT_ROBOT_CAR) {
    gxt_car_step(p, d, a0, a1, c0, c1, g.px, g.py, g.r00, g.r10);
  } else {
    c0 = d.hc * a0; c1 = d.hs * a0; c2 = a1;
    gx_mjx_step_point(p.pt, d, c0, c1, c2, g.px, g.py, g.r00, g.r10, nullptr);
  }
  // Fused step + reset_done with ONE observation pass (as the Point kernels do it, gx_lanes.cuh): whether the env finishes is
  // decided BEFORE the lidar -- target distance, jump and timeout need no lidar, and with sane objects and a sane pose
  // snapshot every entry the lidar / compass pass writes is finite by construction, so the NaN / Inf guard reduces to ctrl /
  // qpos / qvel.  A finishing env then takes its cost from a cost-only pass at the step pose, is re-placed, and the single
  // observation pass produces its reset observation.  A launch lasts as long as its slowest warp and nearly every launch has
  // a finishing env: this takes the second lidar pass out of that warp.  Anything exotic keeps the reference order.
  const bool fast = AUTORESET && gxt_sum8i(staged_ok, mask) == GXT_LPE && (gx_abs(g.px) < GX_SANE_BOUND) && (gx_abs(g.py) < GX_SANE_BOUND) &&
                    !gx_notfinite(g.r00) && !gx_notfinite(g.r10) && !gx_notfinite(g.time) &&
                    (p.pt.use_exp ? p.pt.neg_exp_gain <= 0.0f : p.pt.lidar_max_dist > 0.0f);
  __syncwarp(mask);
  gxt_resolve_pillars(p, ob, d);
  GxtObs o;
  o.bad = 0;
  if (fast) gxt_task_dists(p, ob, gxt_target_index(p, goal_button), g, o);
  else if (AUTORESET) o = gxt_obs_lanes_cold(p, S, ob, i, l, mask, g, goal_button, row, live ? &d : nullptr);   // exotic state: reference order, out of line
  else o = gxt_obs_lanes(p, S, ob, i, l, mask, g, goal_button, row, live ? &d : nullptr);
  int bad = o.bad | gx_notfinite(c0) | gx_notfinite(c1) | gx_notfinite(c2) | gx_notfinite(d.qx) | gx_notfinite(d.qy) | gx_notfinite(d.qth) |
            gx_notfinite(d.vx) | gx_notfinite(d.vy) | gx_notfinite(d.vw);
  const float dist = o.dist_task;
  const float last = last_done > 0.0f ? dist : last_dist;
  const float d_dist = last - dist;
  float rew, dn_flag;
  int flags = 0;
  if (p.task == GXT_TASK_PUSH) {
    const float lastb = last_done > 0.0f ? o.dist_box : last_box;
    rew = (lastb - o.dist_box) * p.reward_box_dist + d_dist * p.reward_box_goal;
    dn_flag = dist < p.goal_size ? 1.0f : 0.0f;
  } else if (p.task == GXT_TASK_BUTTON) {
    rew = d_dist * p.reward_distance;
    dn_flag = dist < p.size[GXT_ROBOT] + p.size[GXT_BUTTON] ? 1.0f : 0.0f;
  } else {
    rew = d_dist * p.reward_distance;
    dn_flag = dist < p.goal_size ? 1.0f : 0.0f;
  }
  if (dn_flag > 0.0f) flags |= GX_AUX_GOAL;
  if (gx_abs(d_dist) > 1.0f) { dn_flag = 1.0f; rew = 0.0f; flags |= GX_AUX_JUMP; }
  if (bad) { rew = 0.0f; dn_flag = 1.0f; flags |= GX_AUX_GUARD; }
  if (d.steps > p.num_steps) { dn_flag = 1.0f; flags |= GX_AUX_TIMEOUT; }
  if (fast && dn_flag > 0.0f) gxt_cost_lanes(p, ob, l, mask, g, goal_button, o);
  // the ONE inlined observation pass of the fused kernel: step pose with pushes (fast path, env goes on) or reset pose
  // (env finished; after the re-placement below)
  GxtGeom gg = g;
  const GxDyn* push = live ? &d : nullptr;
  const bool second = AUTORESET && dn_flag > 0.0f;
  d.steps = dn_flag > 0.0f ? 0.0f : d.steps + 1.0f;
  d.done = dn_flag;
  d.prev_dist = dist;
  prev_box = o.dist_box;
  d.hc = g.r00; d.hs = g.r10;
  d.age += 1;
  if (AUTORESET && dn_flag > 0.0f) {
    // re-place (the 8 lanes copy the layout row), then the reset observation (the only pass over the objects' lidar on the
    // fast path; the second one otherwise)
    __syncwarp(mask);
    const float* lay = gxt_pick_layout(p, pool, nullptr, i, d.age, GX_STREAM_PICK);
    for (int k = l; k < 2 * p.nobj; k += GXT_LPE) {
      const float v = lay[k];
      ob[k] = v;
      if (live) S[(size_t)(GXT_SLOT_OBJ + k) * p.cap + i] = v;
    }
    for (int gq = l; gq < p.ng; gq += GXT_LPE) {
      const float v = lay[2 * p.nobj + 2 + gq];
      ob[2 * p.nobj + gq] = v;
      if (live) S[(size_t)(GXT_SLOT_OBJ + 2 * p.nobj + gq) * p.cap + i] = v;
    }
    d.qx = lay[2 * p.nobj]; d.qy = lay[2 * p.nobj + 1];
    d.qth = 0.0f; d.vx = 0.0f; d.vy = 0.0f; d.vw = 0.0f;
    goal_button = 0.0f;
    if (p.task == GXT_TASK_BUTTON) {
      const uint64_t gid = p.pt.env_id_offset + (uint64_t)i;
      const gx_u32x4 r = gx_philox4x32_10((uint32_t)gid, (uint32_t)d.age, (uint32_t)(d.age >> 32), GX_STREAM_BUTTON, p.pt.seed_lo, p.pt.seed_hi);
      goal_button = (float)gx_randint(r.v[0], (uint32_t)p.count[GXT_BUTTON]);
    }
    if (live && l == 0) S[(size_t)14 * p.cap + i] = goal_button;
    c0 = 0.0f; c1 = 0.0f; c2 = 0.0f;
    GxtGeom g2;
    g2.px = d.qx; g2.py = d.qy; g2.r00 = 1.0f; g2.r10 = 0.0f; g2.time = 0.0f;
    if (!fast) gxt_zero_lidars(p, row, l);      // (fast path: nothing has been scattered into the row yet)
    __threadfence_block();
    __syncwarp(mask);
    gg = g2;
    push = nullptr;
  }
  if (fast || second) {
    const GxtObs f = gxt_obs_lanes(p, S, ob, i, l, mask, gg, goal_button, row, push);
    if (fast && !second) { o.cost = f.cost; o.violations = f.violations; o.contacts = f.contacts; o.goal_bin = f.goal_bin; }
  }
  if (l == 0) {
    if (p.off_ctrl >= 0) { row[p.off_ctrl] = c0; row[p.off_ctrl + 1] = c1; if (p.nu == 3) row[p.off_ctrl + 2] = c2; }
    if (p.off_qpos >= 0) { row[p.off_qpos] = d.qx; row[p.off_qpos + 1] = d.qy; row[p.off_qpos + 2] = d.qth; }
    if (p.off_qvel >= 0) { row[p.off_qvel] = d.vx; row[p.off_qvel + 1] = d.vy; row[p.off_qvel + 2] = d.vw; }
    if (live) {
      gxt_store_dyn(p, S, i, d, prev_box);
      reward[i] = rew; done[i] = dn_flag; cost[i] = o.cost;
      if (aux) { aux[4 * (size_t)i] = flags; aux[4 * (size_t)i + 1] = o.violations; aux[4 * (size_t)i + 2] = o.contacts; aux[4 * (size_t)i + 3] = o.goal_bin; }
    }
  }
  __syncwarp();
  // the warp's 4 rows are contiguous in shared memory and in the output
  const int rows = min(32 / GXT_LPE, p.n - wenv0);
  float* dst = obs + (size_t)wenv0 * p.obs_dim;
  const int total = rows * p.obs_dim;
  if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(tile)) & 15u) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(tile);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int k = lane; k < (total >> 2); k += 32) d4[k] = s4[k];     // a full warp's four rows are 16-byte aligned at both ends
  } else {
    for (int k = lane; k < total; k += 32) dst[k] = tile[k];
  }
}

__global__ void gxt_reset_done_kernel(const __grid_constant__ gxt_params p, float* __restrict__ S, const float* __restrict__ pool,
                                      const float* __restrict__ layout, const float* obs_in, float* obs_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  float* row = gxt_smem + (size_t)threadIdx.x * p.obs_dim;
  const bool is_done = S[(size_t)10 * p.cap + i] > 0.0f;
  if (!is_done) {
    if (obs_in != obs_out) for (int k = 0; k < p.obs_dim; ++k) obs_out[(size_t)i * p.obs_dim + k] = obs_in[(size_t)i * p.obs_dim + k];
    return;
  }
  GxDyn d;
  float prev_box, goal_button;
  gxt_load_dyn(p, S, i, d, prev_box, goal_button);
  const float* lay = gxt_pick_layout(p, pool, layout, i, d.age, GX_STREAM_PICK);
  gxt_place(p, S, i, lay, d, goal_button, d.age);
  gxt_reset_row(p, S, i, d, goal_button, row);
  gxt_store_dyn(p, S, i, d, prev_box);
  for (int k = 0; k < p.obs_dim; ++k) obs_out[(size_t)i * p.obs_dim + k] = row[k];
}

__global__ void gxt_get_state_kernel(const __grid_constant__ gxt_params p, const float* __restrict__ S, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  for (int k = 0; k < p.state_dim; ++k) out[(size_t)i * p.state_dim + k] = S[(size_t)k * p.cap + i];
}
__global__ void gxt_set_state_kernel(const __grid_constant__ gxt_params p, float* __restrict__ S, const float* __restrict__ in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  for (int k = 0; k < p.state_dim; ++k)
    if (k != 12 && k != 13) S[(size_t)k * p.cap + i] = in[(size_t)i * p.state_dim + k];   // the env keeps its age
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
GXT_API void gxt_default_config(gxt_config* c) {
  if (!c) return;
  memset(c, 0, sizeof *c);
  c->abi_version = GXT_ABI_VERSION;
  c->env_num = 1; c->num_steps = 1000; c->task = GXT_TASK_BUTTON; c->robot = GXT_ROBOT_CAR;
  const int cnt[GXT_NCLASS] = {0, 0, 4, 4, 4, 2, 4, 1};
  const double size[GXT_NCLASS] = {0.3, 0.2, 0.1, 0.2, 0.1, 0.2, 0.1, 0.2};
  const double keep[GXT_NCLASS] = {0.305, 0.2, 0.2, 0.18, 0.15, 0.3, 0.5, 0.4};
  for (int k = 0; k < GXT_NCLASS; ++k) { c->count[k] = cnt[k]; c->size[k] = size[k]; c->keepout[k] = keep[k]; }
  c->placements_extents[0] = -2; c->placements_extents[1] = -2; c->placements_extents[2] = 2; c->placements_extents[3] = 2;
  c->placements_margin = 0.0; c->goal_min_robot_dist = 0.0; c->layout_tries = 10; c->pool_size = 0;
  c->reward_distance = 1.0; c->reward_box_dist = 1.0; c->reward_box_goal = 1.0; c->hazards_cost_scale = 1.0;
  c->vases_contact_cost = 1.0; c->pillars_cost = 1.0; c->buttons_cost = 1.0;
  c->gremlins_contact_cost = 1.0; c->gremlins_dist_threshold = 0.2; c->gremlins_dist_cost = 1.0; c->gremlins_travel = 0.3;
  c->lidar_num_bins = 16; c->lidar_alias = 1; c->lidar_max_dist = 0.0; c->lidar_exp_gain = 1.0;
  c->obs_mask = GXT_OBS_BUTTONS_LIDAR | GXT_OBS_CTRL | GXT_OBS_GOAL_COMPASS | GXT_OBS_GOAL_LIDAR | GXT_OBS_GREMLINS_LIDAR |
                GXT_OBS_HAZARDS_LIDAR | GXT_OBS_PILLARS_LIDAR | GXT_OBS_QPOS | GXT_OBS_QVEL | GXT_OBS_VASES_LIDAR;
  gx_config pc;
  gx_default_config(&pc);
  c->pt_timestep = pc.point.timestep; c->pt_mass = pc.point.mass; c->pt_com_x = pc.point.com_x; c->pt_inertia_zz = pc.point.inertia_zz;
  for (int k = 0; k < 3; ++k) { c->pt_damping[k] = pc.point.damping[k]; c->pt_gear[k] = pc.point.gear[k]; }
  c->pt_gain = pc.point.gainprm0; c->pt_bias2 = pc.point.biasprm2;
  c->pt_ctrl[0] = pc.point.ctrlrange[0]; c->pt_ctrl[1] = pc.point.ctrlrange[1];
  c->pt_force[0] = pc.point.forcerange[0]; c->pt_force[1] = pc.point.forcerange[1];
  // Car: chassis box 0.4 x 0.2 x 0.1 m, density 1 as in point.xml's defaults; two driven wheels r = 0.05, half track 0.1
  c->car_mass = 0.008f; c->car_inertia = 1.33e-4f; c->car_wheel_r = 0.05f; c->car_half_track = 0.1f; c->car_gear = 6.0e-4f;
  c->car_damp_lin = 0.01f; c->car_damp_ang = 2.0e-4f; c->car_timestep = 0.02f;
}

static const char* gxt_lower(const gxt_config* c, gxt_params* p) {
  if (!c) return "null config";
  if (c->abi_version != GXT_ABI_VERSION) return "abi_version mismatch";
  if (c->env_num < 1) return "env_num < 1";
  if (c->task < 0 || c->task > 2) return "unknown task";
  if (c->robot != GXT_ROBOT_POINT && c->robot != GXT_ROBOT_CAR) return "unknown robot";
  memset(p, 0, sizeof *p);
  // lidar / Point constants through the Point path's own lowering (same roundings)
  gx_config pc;
  gx_default_config(&pc);
  pc.env_num = c->env_num; pc.seed = c->seed; pc.env_id_offset = c->env_id_offset;
  pc.lidar_num_bins = c->lidar_num_bins; pc.lidar_alias = c->lidar_alias; pc.lidar_max_dist = c->lidar_max_dist; pc.lidar_exp_gain = c->lidar_exp_gain;
  pc.point.timestep = c->pt_timestep; pc.point.mass = c->pt_mass; pc.point.com_x = c->pt_com_x; pc.point.inertia_zz = c->pt_inertia_zz;
  for (int k = 0; k < 3; ++k) { pc.point.damping[k] = c->pt_damping[k]; pc.point.gear[k] = c->pt_gear[k]; }
  pc.point.gainprm0 = c->pt_gain; pc.point.biasprm2 = c->pt_bias2;
  pc.point.ctrlrange[0] = c->pt_ctrl[0]; pc.point.ctrlrange[1] = c->pt_ctrl[1];
  pc.point.forcerange[0] = c->pt_force[0]; pc.point.forcerange[1] = c->pt_force[1];
  pc.opt_timestep = c->pt_timestep;
  if (const char* why = gx_lower_config(&pc, &p->pt)) return why;
  p->n = c->env_num; p->cap = (c->env_num + 31) & ~31; p->task = c->task; p->robot = c->robot;
  int nobj = 0;
  for (int k = 0; k < GXT_ROBOT; ++k) {
    if (c->count[k] < 0 || c->count[k] > GXT_MAX_PER_CLASS) return "object count out of range";
    p->first[k] = nobj; p->count[k] = c->count[k];
    nobj += c->count[k];
  }
  p->first[GXT_ROBOT] = nobj; p->count[GXT_ROBOT] = 1;
  if (nobj > GXT_MAX_OBJECTS) return "too many objects";
  if (c->task == GXT_TASK_BUTTON && (c->count[GXT_BUTTON] < 1 || c->count[GXT_GOAL] != 0)) return "button task: buttons_num >= 1 and no goal object";
  if (c->task == GXT_TASK_GOAL && c->count[GXT_GOAL] != 1) return "goal task: exactly one goal";
  if (c->task == GXT_TASK_PUSH && (c->count[GXT_GOAL] != 1 || c->count[GXT_BOX] != 1)) return "push task: one goal and one box";
  if (c->task != GXT_TASK_PUSH && c->count[GXT_BOX] != 0) return "a box only exists in the push task";
  p->nobj = nobj; p->ng = c->count[GXT_GREMLIN];
  static_assert(GXT_MAX_OBJECTS == 64 && GXT_NCLASS <= 16, "cls_bits packs 64 objects x 4 bits");
  for (int w = 0; w < GXT_MAX_OBJECTS / 16; ++w) p->cls_bits[w] = 0;
  for (int o = 0; o < nobj; ++o) p->cls_bits[o >> 4] |= (uint64_t)gxt_class_scan(*p, o) << (4 * (o & 15));
  p->nu = c->robot == GXT_ROBOT_CAR ? 2 : 3;
  for (int k = 0; k < GXT_NCLASS; ++k) {
    p->size[k] = (float)c->size[k]; p->keep[k] = (float)c->keepout[k];
    p->lo[k][0] = (float)(c->placements_extents[0] + c->keepout[k]); p->lo[k][1] = (float)(c->placements_extents[1] + c->keepout[k]);
    p->hi[k][0] = (float)(c->placements_extents[2] - c->keepout[k]); p->hi[k][1] = (float)(c->placements_extents[3] - c->keepout[k]);
    if (!(p->lo[k][0] <= p->hi[k][0]) || !(p->lo[k][1] <= p->hi[k][1])) return "placements_extents smaller than a keepout";
  }
  p->margin = (float)c->placements_margin; p->min_robot_goal = (float)c->goal_min_robot_dist;
  if (c->layout_tries < 1 || c->layout_tries > 255) return "layout_tries out of range";
  p->tries = c->layout_tries;
  p->pool_size = c->pool_size > 0 ? c->pool_size : gx_auto_pool_size(c->env_num);
  p->num_steps = (float)c->num_steps; p->reward_distance = (float)c->reward_distance;
  p->reward_box_dist = (float)c->reward_box_dist; p->reward_box_goal = (float)c->reward_box_goal;
  p->hazards_cost_scale = (float)c->hazards_cost_scale; p->vases_contact_cost = (float)c->vases_contact_cost;
  p->pillars_cost = (float)c->pillars_cost; p->buttons_cost = (float)c->buttons_cost;
  p->gremlins_contact_cost = (float)c->gremlins_contact_cost; p->gremlins_dist_threshold = (float)c->gremlins_dist_threshold;
  p->gremlins_dist_cost = (float)c->gremlins_dist_cost; p->gremlins_travel = (float)c->gremlins_travel;
  p->goal_size = (float)c->size[GXT_GOAL];
  p->dt = c->robot == GXT_ROBOT_CAR ? c->car_timestep : c->pt_timestep;
  p->car_mass = c->car_mass; p->car_inertia = c->car_inertia; p->car_wheel_r = c->car_wheel_r; p->car_half_track = c->car_half_track;
  p->car_gear = c->car_gear;
  { float t; t = c->car_timestep * c->car_damp_lin; p->car_mh_lin = c->car_mass + t; t = c->car_timestep * c->car_damp_ang; p->car_mh_ang = c->car_inertia + t; }
  const int nb = c->lidar_num_bins;
  int off = 0;
  auto take = [&](uint32_t bit, int width, bool exists) { int r = -1; if ((c->obs_mask & bit) && exists) { r = off; off += width; } return r; };
  p->off_box_compass = take(GXT_OBS_BOX_COMPASS, 2, c->count[GXT_BOX] > 0);
  p->off_box_lidar = take(GXT_OBS_BOX_LIDAR, nb, c->count[GXT_BOX] > 0);
  p->off_buttons_lidar = take(GXT_OBS_BUTTONS_LIDAR, nb, c->count[GXT_BUTTON] > 0);
  p->off_ctrl = take(GXT_OBS_CTRL, p->nu, true);
  p->off_goal_compass = take(GXT_OBS_GOAL_COMPASS, 2, true);
  p->off_goal_lidar = take(GXT_OBS_GOAL_LIDAR, nb, true);
  p->off_gremlins_lidar = take(GXT_OBS_GREMLINS_LIDAR, nb, c->count[GXT_GREMLIN] > 0);
  p->off_hazards_lidar = take(GXT_OBS_HAZARDS_LIDAR, nb, c->count[GXT_HAZARD] > 0);
  p->off_pillars_lidar = take(GXT_OBS_PILLARS_LIDAR, nb, c->count[GXT_PILLAR] > 0);
  p->off_qpos = take(GXT_OBS_QPOS, 3, true);
  p->off_qvel = take(GXT_OBS_QVEL, 3, true);
  p->off_vases_lidar = take(GXT_OBS_VASES_LIDAR, nb, c->count[GXT_VASE] > 0);
  if (off == 0) return "empty observation";
  p->obs_dim = off;
  p->state_dim = GXT_SLOT_OBJ + 2 * nobj + p->ng;
  p->lay_dim = 2 * nobj + 2 + p->ng;
  return nullptr;
}

GXT_API int gxt_create(const gxt_config* c, int device, gxt_env** out) {
  if (!out) return gx_set_error(GX_EINVAL, "gxt_create: out is NULL");
  *out = nullptr;
  gxt_params p;
  if (const char* why = gxt_lower(c, &p)) return gx_set_error(GX_EINVAL, "gxt_create: %s", why);
  int ndev = 0;
  GXT_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return gx_set_error(GX_EINVAL, "gxt_create: device %d of %d", device, ndev);
  GxtDeviceGuard guard(device);
  cudaDeviceProp prop;
  GXT_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return gx_set_error(GX_ECUDA, "gxt_create: device %d is sm_%d%d; libguardx.so is built for sm_100a only", device, prop.major, prop.minor);
  gxt_env* e = new (std::nothrow) gxt_env();
  if (!e) return gx_set_error(GX_ENOMEM, "gxt_create: host allocation failed");
  e->cfg = *c; e->p = p; e->device = device; e->sm_count = prop.multiProcessorCount;
  cudaError_t err = cudaMalloc((void**)&e->S, sizeof(float) * (size_t)p.state_dim * p.cap);
  if (err == cudaSuccess) err = cudaMemset(e->S, 0, sizeof(float) * (size_t)p.state_dim * p.cap);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->pool, sizeof(float) * (size_t)p.pool_size * p.lay_dim);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->pool_tries, sizeof(int32_t) * (size_t)p.pool_size);
  if (err == cudaSuccess) err = cudaMalloc((void**)&e->epoch, sizeof(unsigned long long));
  if (err == cudaSuccess) err = cudaMemset(e->epoch, 0, sizeof(unsigned long long));
  if (err != cudaSuccess) {
    gxt_destroy(e);
    return gx_set_error(err == cudaErrorMemoryAllocation ? GX_ENOMEM : GX_ECUDA, "gxt_create: %s", cudaGetErrorString(err));
  }
  *out = e;
  return GX_OK;
}

GXT_API int gxt_destroy(gxt_env* e) {
  if (!e) return GX_OK;
  GxtDeviceGuard guard(e->device);
  cudaFree(e->S); cudaFree(e->pool); cudaFree(e->pool_tries); cudaFree(e->epoch);
  delete e;
  return GX_OK;
}

GXT_API int gxt_config_size(void) { return (int)sizeof(gxt_config); }
GXT_API int gxt_obs_dim(const gxt_env* e) { return e ? e->p.obs_dim : GX_EINVAL; }
GXT_API int gxt_act_dim(const gxt_env* e) { return e ? 2 : GX_EINVAL; }
GXT_API int gxt_state_dim(const gxt_env* e) { return e ? e->p.state_dim : GX_EINVAL; }
GXT_API int gxt_layout_dim(const gxt_env* e) { return e ? e->p.lay_dim : GX_EINVAL; }
GXT_API int gxt_pool_size(const gxt_env* e) { return e ? e->p.pool_size : GX_EINVAL; }
GXT_API int gxt_obs_offset(const gxt_env* e, uint32_t comp) {
  if (!e) return GX_EINVAL;
  const gxt_params& p = e->p;
  switch (comp) {
    case GXT_OBS_BOX_COMPASS: return p.off_box_compass;
    case GXT_OBS_BOX_LIDAR: return p.off_box_lidar;
    case GXT_OBS_BUTTONS_LIDAR: return p.off_buttons_lidar;
    case GXT_OBS_CTRL: return p.off_ctrl;
    case GXT_OBS_GOAL_COMPASS: return p.off_goal_compass;
    case GXT_OBS_GOAL_LIDAR: return p.off_goal_lidar;
    case GXT_OBS_GREMLINS_LIDAR: return p.off_gremlins_lidar;
    case GXT_OBS_HAZARDS_LIDAR: return p.off_hazards_lidar;
    case GXT_OBS_PILLARS_LIDAR: return p.off_pillars_lidar;
    case GXT_OBS_QPOS: return p.off_qpos;
    case GXT_OBS_QVEL: return p.off_qvel;
    case GXT_OBS_VASES_LIDAR: return p.off_vases_lidar;
    default: return GX_EINVAL;
  }
}

#define GXT_LAUNCH_CHECK(name)                                                                            \
  do {                                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                                  \
    if (e_ != cudaSuccess) return gx_set_error(GX_ECUDA, "%s launch failed: %s", name, cudaGetErrorString(e_)); \
  } while (0)

static inline int gxt_div_up(int a, int b) { return (a + b - 1) / b; }

GXT_API int gxt_reset(gxt_env* e, const float* layout, float* obs, void* stream) {
  if (!e || !obs) return gx_set_error(GX_EINVAL, "gxt_reset: NULL argument");
  GxtDeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gxt_params& p = e->p;
  if (!layout) {
    gxt_pool_kernel<<<gxt_div_up(p.pool_size * 32, 128), 128, 0, st>>>(p, e->epoch, e->pool, e->pool_tries);
    GXT_LAUNCH_CHECK("gxt_pool_kernel");
    gxt_bump_epoch<<<1, 1, 0, st>>>(e->epoch);
  }
  const int block = 64;
  gxt_reset_kernel<<<gxt_div_up(p.n, block), block, sizeof(float) * block * p.obs_dim, st>>>(p, e->S, e->pool, layout, obs);
  GXT_LAUNCH_CHECK("gxt_reset_kernel");
  e->have_reset = true;
  return GX_OK;
}

GXT_API int gxt_step(gxt_env* e, const float* action, float* obs, float* reward, float* done, float* cost, int32_t* aux,
                     int32_t autoreset, void* stream) {
  if (!e || !action || !obs || !reward || !done || !cost) return gx_set_error(GX_EINVAL, "gxt_step: NULL argument");
  if (!e->have_reset) return gx_set_error(GX_ESTATE, "gxt_step: call gxt_reset first");
  GxtDeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  const gxt_params& p = e->p;
  // 8 lanes per env while the batch cannot fill the machine with one thread per env (GXT_LANES=0|1 in the environment forces a path)
  const char* force = getenv("GXT_LANES");
  const bool lanes = force ? force[0] == '1' : p.n <= e->sm_count * 1024;
  if (lanes) {
    const int block = 128, epb = block / GXT_LPE;
    const size_t sh = sizeof(float) * epb * (p.obs_dim + 2 * p.nobj + p.ng);     // observation rows + staged object coordinates
    // chained by programmatic dependent launch, as the Point step kernels are (guardx.cu: gx_launch; GX_PDL=0 switches it off)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gxt_div_up(p.n, epb)); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = sh; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    static const bool pdl = [] { const char* v = getenv("GX_PDL"); return !(v && v[0] == '0'); }();
    cfg.numAttrs = pdl ? 1 : 0;
    float* S_ = e->S; const float* pool_ = e->pool;
    if (autoreset) cudaLaunchKernelEx(&cfg, gxt_step_lanes_kernel<true>, p, S_, pool_, action, obs, reward, done, cost, aux);
    else cudaLaunchKernelEx(&cfg, gxt_step_lanes_kernel<false>, p, S_, pool_, action, obs, reward, done, cost, aux);
    GXT_LAUNCH_CHECK("gxt_step_lanes_kernel");
    return GX_OK;
  }
  const int block = 64;
  const size_t sh = sizeof(float) * block * p.obs_dim;
  if (autoreset) gxt_step_kernel<true><<<gxt_div_up(p.n, block), block, sh, st>>>(p, e->S, e->pool, action, obs, reward, done, cost, aux);
  else gxt_step_kernel<false><<<gxt_div_up(p.n, block), block, sh, st>>>(p, e->S, e->pool, action, obs, reward, done, cost, aux);
  GXT_LAUNCH_CHECK("gxt_step_kernel");
  return GX_OK;
}

GXT_API int gxt_reset_done(gxt_env* e, const float* obs_in, float* obs_out, const float* layout, void* stream) {
  if (!e || !obs_in || !obs_out) return gx_set_error(GX_EINVAL, "gxt_reset_done: NULL argument");
  if (!e->have_reset) return gx_set_error(GX_ESTATE, "gxt_reset_done: call gxt_reset first");
  GxtDeviceGuard guard(e->device);
  const gxt_params& p = e->p;
  const int block = 64;
  gxt_reset_done_kernel<<<gxt_div_up(p.n, block), block, sizeof(float) * block * p.obs_dim, (cudaStream_t)stream>>>(p, e->S, e->pool, layout, obs_in, obs_out);
  GXT_LAUNCH_CHECK("gxt_reset_done_kernel");
  return GX_OK;
}

GXT_API int gxt_get_state(gxt_env* e, float* state, void* stream) {
  if (!e || !state) return gx_set_error(GX_EINVAL, "gxt_get_state: NULL argument");
  GxtDeviceGuard guard(e->device);
  gxt_get_state_kernel<<<gxt_div_up(e->p.n, 128), 128, 0, (cudaStream_t)stream>>>(e->p, e->S, state);
  GXT_LAUNCH_CHECK("gxt_get_state_kernel");
  return GX_OK;
}

GXT_API int gxt_set_state(gxt_env* e, const float* state, void* stream) {
  if (!e || !state) return gx_set_error(GX_EINVAL, "gxt_set_state: NULL argument");
  GxtDeviceGuard guard(e->device);
  gxt_set_state_kernel<<<gxt_div_up(e->p.n, 128), 128, 0, (cudaStream_t)stream>>>(e->p, e->S, state);
  GXT_LAUNCH_CHECK("gxt_set_state_kernel");
  e->have_reset = true;
  return GX_OK;
}

GXT_API int gxt_pool_read(gxt_env* e, float* pool, int32_t* tries, void* stream) {
  if (!e) return gx_set_error(GX_EINVAL, "gxt_pool_read: NULL env");
  GxtDeviceGuard guard(e->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (pool) GXT_CUDA(cudaMemcpyAsync(pool, e->pool, sizeof(float) * (size_t)e->p.pool_size * e->p.lay_dim, cudaMemcpyDeviceToDevice, st));
  if (tries) GXT_CUDA(cudaMemcpyAsync(tries, e->pool_tries, sizeof(int32_t) * (size_t)e->p.pool_size, cudaMemcpyDeviceToDevice, st));
  return GX_OK;
}
