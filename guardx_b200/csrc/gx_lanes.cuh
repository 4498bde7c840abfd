// gx_lanes.cuh -- the env step with 16 LANES PER ENV (two envs per warp).
//
// At BASELINE.json's headline size (4096 envs) one thread per env leaves < 1 warp
// per SM sub-partition and the step is a ~2700-instruction dependent chain per
// thread (ncu: 7.5 cycles per issued instruction, profiles/).  Here the nine
// lidar objects of an env (goal + 8 hazards) are evaluated by nine different
// lanes, the 2 x 16 lidar bins are assembled one bin per lane with sub-warp
// shuffles, the cost terms are combined by shuffles, the NaN/Inf guard is a
// ballot, and the 43-float observation row is written by the env's 16 lanes as
// contiguous stores (no shared-memory staging needed).  The dependent chain
// shrinks to dynamics + one lidar return + the gather.
//
// Only the default shape is handled here (hazards_num = 8, lidar_num_bins = 16,
// all six observation components => obs_dim 43); other configurations use the
// one-thread-per-env kernels.
//
// Bit-parity: every fp32 value is produced by the same expression as in
// gx_device.cuh / the oracle.  The lidar bins are folded in the oracle's update
// order (object order; main bin, bin+1, bin-1), so even signed zeros agree.
#ifndef GX_LANES_CUH_
#define GX_LANES_CUH_

#include "gx_device.cuh"

#define GX_LPE 16  // lanes per env

// mask of the 16 lanes that share this thread's env.  All sub-warp collectives
// below name only these lanes, so the two envs of a warp may diverge (one being
// re-placed while the other is not).
__device__ __forceinline__ unsigned gx_grp_mask() { return 0xffffu << (threadIdx.x & 16); }

// value of lane `src` of this env's 16-lane group
template <typename T>
__device__ __forceinline__ T gx_grp(unsigned gmask, T v, int src) {
  return __shfl_sync(gmask, v, src, GX_LPE);
}

// Fold one ray into the running value of output bin j, in the oracle's order.
// (lidar_alias is on in every configuration that reaches the lanes kernels.)
__device__ __forceinline__ float gx_fold_ray(float v, int j, int bin, float sensor, float up, float down) {
  const int bp = (bin + 1) & 15;  // bin in 0..16: (bin+1) % 16
  const int bm = (bin - 1) & 15;  // (bin-1) mod 16, also for bin = 0 and bin = 16
  // A contribution that targets another bin is replaced by +0, the neutral
  // element here: v is always >= +0 or NaN, and max(v, +0) == v bit for bit.
  // bin == 16 never equals j: the out-of-bounds scatter is dropped (JAX).
  v = gx_max(v, (bin == j) ? sensor : 0.0f);
  v = gx_max(v, (bp == j) ? up : 0.0f);
  v = gx_max(v, (bm == j) ? down : 0.0f);
  return v;
}

struct GxLaneObs {
  float glidar, hlidar;  // this lane's bin of the goal / hazards lidar
  float cx, cy;          // compass
  float dist_goal, cost;
  int violations, bad;
};

// Observation geometry for one env from pose (px,py,R): lane `sub` owns object
// `sub` (0 = goal, 1..8 = hazard sub-1) at (ox,oy) and output bin `sub`.
// `gmask` names the lanes that execute this call together: the whole warp on the
// main path (compile-time 0xffffffff, no membership check), the env's own 16
// lanes inside the divergent re-placement branch.  `gbits` are the env's lanes.
__device__ __forceinline__ GxLaneObs gx_obs_lanes(const gx_params& p, int sub, unsigned gmask, unsigned gbits, float ox, float oy,
                                                  float gx_, float gy_, float px, float py, float r00,
                                                  float r10) {
  GxLaneObs o;
  float ex, ey;
  gx_ego(ox, oy, px, py, r00, r10, ex, ey);
  GxRay r = gx_lidar_ray<true>(p, ex, ey);  // lanes 9..15 compute a dummy ray (same code path, result unused)
  // cost term of this lane's hazard (engine.py:808-810)
  float dx = ox - px, dy = oy - py;
  float dist = gx_sqrt(dx * dx + dy * dy);
  float term = p.hazards_size - gx_min(dist, p.hazards_size);
  const bool is_h = (sub >= 1 && sub <= 8);
  if (!is_h) term = 0.0f;
  // goal: compass and distance (lane 0's own numbers, broadcast)
  o.cx = gx_grp(gmask, ex, 0);
  o.cy = gx_grp(gmask, ey, 0);
  {
    float gdx = gx_ - px, gdy = gy_ - py;
    o.dist_goal = gx_sqrt(gdx * gdx + gdy * gdy);
  }
  // goal lidar, bin `sub`
  {
    int b = gx_grp(gmask, r.bin, 0);
    float s0 = gx_grp(gmask, r.sensor, 0), u0 = gx_grp(gmask, r.up, 0), d0 = gx_grp(gmask, r.down, 0);
    o.glidar = gx_fold_ray(0.0f, sub, b, s0, u0, d0);
  }
  // hazards lidar, bin `sub`; cost summed in hazard order
  float v = 0.0f, cost = 0.0f;
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    int b = gx_grp(gmask, r.bin, 1 + h);
    float s1 = gx_grp(gmask, r.sensor, 1 + h), u1 = gx_grp(gmask, r.up, 1 + h), d1 = gx_grp(gmask, r.down, 1 + h);
    v = gx_fold_ray(v, sub, b, s1, u1, d1);
    cost = cost + gx_grp(gmask, term, 1 + h);
  }
  o.hlidar = v;
  o.cost = cost;
  const unsigned viol = __ballot_sync(gmask, term > 0.0f) & gbits;
  o.violations = __popc(viol);
  const unsigned badm = __ballot_sync(gmask, gx_notfinite(o.glidar) | gx_notfinite(o.hlidar)) & gbits;
  o.bad = (badm != 0) | gx_notfinite(o.cx) | gx_notfinite(o.cy);
  return o;
}

// Lane `sub` stores its share of the 43-float row:
//   [0..2] ctrl  [3..4] compass  [5..20] goal lidar  [21..36] hazards lidar  [37..39] qpos  [40..42] qvel
__device__ __forceinline__ void gx_store_row_lanes(float* row, int sub, const GxLaneObs& o, float c0,
                                                   float c1, float c2, const GxDyn& d) {
  row[5 + sub] = o.glidar;
  row[21 + sub] = o.hlidar;
  float v;
  switch (sub) {
    case 0: v = c0; break;
    case 1: v = c1; break;
    case 2: v = c2; break;
    case 3: v = o.cx; break;
    case 4: v = o.cy; break;
    case 5: v = d.qx; break;
    case 6: v = d.qy; break;
    case 7: v = d.qth; break;
    case 8: v = d.vx; break;
    case 9: v = d.vy; break;
    default: v = d.vw; break;
  }
  if (sub < 11) row[sub < 5 ? sub : 32 + sub] = v;
}

// this lane's object from the SoA state: goal for lane 0, hazard sub-1 for 1..8
__device__ __forceinline__ void gx_load_object(const GxState& s, int i, int sub, float gx_, float gy_,
                                               float& ox, float& oy) {
  ox = gx_; oy = gy_;
  if (sub >= 1 && sub <= 8) {
    const int h = sub - 1;
    const float2* pair = reinterpret_cast<const float2*>(s.hz + (size_t)(h >> 1) * s.cap + i);
    float2 v = pair[h & 1];
    ox = v.x; oy = v.y;
  }
}

// Re-place a finished env from layout row `lay` (16 lanes cooperate): static
// state goes back to HBM, the lanes pick up their new objects, the observation
// is recomputed at the reset pose (R = I, qvel = ctrl = 0).
__device__ __forceinline__ void gx_reset_done_lanes(const gx_params& p, const GxState& s, int i, int sub,
                                                    unsigned gmask, bool live, const float* lay, GxDyn& d, float& gx_, float& gy_,
                                                    float& ox, float& oy, GxLaneObs& o) {
  const float2* l2 = reinterpret_cast<const float2*>(lay);  // goal, robot, hazard0..7
  float2 g = l2[0], rb = l2[1];
  gx_ = g.x; gy_ = g.y;
  d.qx = rb.x; d.qy = rb.y; d.qth = 0.0f; d.vx = 0.0f; d.vy = 0.0f; d.vw = 0.0f;
  ox = gx_; oy = gy_;
  if (sub == 0 && live) s.goal[i] = g;
  if (sub >= 1 && sub <= 8) {
    float2 hv = l2[1 + sub];
    ox = hv.x; oy = hv.y;
    float2* pair = reinterpret_cast<float2*>(s.hz + (size_t)((sub - 1) >> 1) * s.cap + i);
    if (live) pair[(sub - 1) & 1] = hv;
  }
  float r00, r10;
  gx_kinematics(d.qth, r00, r10);
  o = gx_obs_lanes(p, sub, gmask, gmask, ox, oy, gx_, gy_, d.qx, d.qy, r00, r10);
}

struct GxLaneStep {
  GxStepOut out;
  GxLaneObs obs;
  float c0, c1, c2;
};

// One transition, all 16 lanes of the env execute it (dynamics redundantly).
__device__ __forceinline__ GxLaneStep gx_step_lanes(const gx_params& p, int sub, unsigned gbits, GxDyn& d, float gx_,
                                                    float gy_, float ox, float oy, float a0, float a1,
                                                    float* qacc) {
  GxLaneStep r;
  r.c0 = d.hc * a0; r.c1 = d.hs * a0; r.c2 = a1;
  const float last_dist = d.prev_dist, last_done = d.done;
  float px, py, r00, r10;
  gx_mjx_step_point<true>(p, d, r.c0, r.c1, r.c2, px, py, r00, r10, qacc);  // substeps == 1 here
  r.obs = gx_obs_lanes(p, sub, 0xffffffffu, gbits, ox, oy, gx_, gy_, px, py, r00, r10);
  int bad = r.obs.bad | gx_notfinite(r.c0) | gx_notfinite(r.c1) | gx_notfinite(r.c2) |
            gx_notfinite(d.qx) | gx_notfinite(d.qy) | gx_notfinite(d.qth) |
            gx_notfinite(d.vx) | gx_notfinite(d.vy) | gx_notfinite(d.vw);
  const float dist = r.obs.dist_goal;
  const float last = last_done > 0.0f ? dist : last_dist;
  const float d_dist = last - dist;
  GxStepOut& out = r.out;
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
  out.cost = r.obs.cost;
  out.violations = r.obs.violations;
  return r;
}

// state write-back by three different lanes of the env (one 16-B or 8-B store each)
__device__ __forceinline__ void gx_store_dyn_lanes(const GxState& s, int i, int sub, const GxDyn& d) {
  if (sub == 0) s.d0[i] = make_float4(d.qx, d.qy, d.qth, d.vx);
  if (sub == 1) s.d1[i] = make_float4(d.vy, d.vw, d.hc, d.hs);
  if (sub == 2) s.d2[i] = gx_pack_d2(d);
}

#endif  // GX_LANES_CUH_
