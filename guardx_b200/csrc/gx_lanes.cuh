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

// Cost of one env at pose (px,py): lane `sub` in 1..8 owns hazard sub-1 at (ox,oy); the terms
// are summed in hazard order (engine.py:808-810), violations = number of positive terms.
__device__ __forceinline__ void gx_cost_lanes(const gx_params& p, int sub, unsigned gmask, unsigned gbits, float ox, float oy,
                                              float px, float py, float& cost_out, int& violations) {
  float dx = ox - px, dy = oy - py;
  float dist = gx_sqrt(dx * dx + dy * dy);
  float term = p.hazards_size - gx_min(dist, p.hazards_size);
  const bool is_h = (sub >= 1 && sub <= 8);
  if (!is_h) term = 0.0f;
  float cost = 0.0f;
#pragma unroll
  for (int h = 0; h < 8; ++h) cost = cost + gx_grp(gmask, term, 1 + h);
  cost_out = cost;
  violations = __popc(__ballot_sync(gmask, term > 0.0f) & gbits);
}

// Lidar + compass of one env from pose (px,py,R): lane `sub` owns object `sub` (0 = goal,
// 1..8 = hazard sub-1) at (ox,oy) and output bin `sub`.
// `gmask` names the lanes that execute this call together: the whole warp on the
// main path (compile-time 0xffffffff, no membership check), the env's own 16
// lanes inside a divergent branch.  `gbits` are the env's lanes.
__device__ __forceinline__ void gx_lidar_lanes(const gx_params& p, int sub, unsigned gmask, unsigned gbits, float ox, float oy,
                                               float px, float py, float r00, float r10, GxLaneObs& o) {
  float ex, ey;
  gx_ego(ox, oy, px, py, r00, r10, ex, ey);
  GxRay r = gx_lidar_ray<true>(p, ex, ey);  // lanes 9..15 compute a dummy ray (same code path, result unused)
  // goal: compass (lane 0's own numbers, broadcast)
  o.cx = gx_grp(gmask, ex, 0);
  o.cy = gx_grp(gmask, ey, 0);
  // goal lidar, bin `sub`
  {
    int b = gx_grp(gmask, r.bin, 0);
    float s0 = gx_grp(gmask, r.sensor, 0), u0 = gx_grp(gmask, r.up, 0), d0 = gx_grp(gmask, r.down, 0);
    o.glidar = gx_fold_ray(0.0f, sub, b, s0, u0, d0);
  }
  // hazards lidar, bin `sub`
  float v = 0.0f;
#pragma unroll
  for (int h = 0; h < 8; ++h) {
    int b = gx_grp(gmask, r.bin, 1 + h);
    float s1 = gx_grp(gmask, r.sensor, 1 + h), u1 = gx_grp(gmask, r.up, 1 + h), d1 = gx_grp(gmask, r.down, 1 + h);
    v = gx_fold_ray(v, sub, b, s1, u1, d1);
  }
  o.hlidar = v;
  const unsigned badm = __ballot_sync(gmask, gx_notfinite(o.glidar) | gx_notfinite(o.hlidar)) & gbits;
  o.bad = (badm != 0) | gx_notfinite(o.cx) | gx_notfinite(o.cy);
}

// Observation geometry for one env from pose (px,py,R) = lidar + compass + goal distance + cost.
__device__ __forceinline__ GxLaneObs gx_obs_lanes(const gx_params& p, int sub, unsigned gmask, unsigned gbits, float ox, float oy,
                                                  float gx_, float gy_, float px, float py, float r00,
                                                  float r10) {
  GxLaneObs o;
  gx_cost_lanes(p, sub, gmask, gbits, ox, oy, px, py, o.cost, o.violations);
  {
    float gdx = gx_ - px, gdy = gy_ - py;
    o.dist_goal = gx_sqrt(gdx * gdx + gdy * gdy);
  }
  gx_lidar_lanes(p, sub, gmask, gbits, ox, oy, px, py, r00, r10, o);
  return o;
}

// Lane `sub` stores its share of the 43-float row:
//   [0..2] ctrl  [3..4] compass  [5..20] goal lidar  [21..36] hazards lidar  [37..39] qpos  [40..42] qvel
__device__ __forceinline__ void gx_store_row_lanes(float* row, int sub, const GxLaneObs& o, float c0,
                                                   float c1, float c2, const GxDyn& d) {
  row[5 + sub] = o.glidar;
  row[21 + sub] = o.hlidar;
#ifdef GX_ROW_SWITCH
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
#else
  // a chain of selects, not a switch: the 16 lanes of an env all take different cases, and a divergent 11-way switch is
  // executed case by case
  float v = d.vw;
  v = sub == 9 ? d.vy : v;
  v = sub == 8 ? d.vx : v;
  v = sub == 7 ? d.qth : v;
  v = sub == 6 ? d.qy : v;
  v = sub == 5 ? d.qx : v;
  v = sub == 4 ? o.cy : v;
  v = sub == 3 ? o.cx : v;
  v = sub == 2 ? c2 : v;
  v = sub == 1 ? c1 : v;
  v = sub == 0 ? c0 : v;
#endif
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

// One transition, all 16 lanes of the env execute it (dynamics redundantly).  `gmask`: the lanes that
// execute the call together (0xffffffff on the convergent main path), `gbits`: the env's lanes.
__device__ __forceinline__ GxLaneStep gx_step_lanes(const gx_params& p, int sub, unsigned gmask, unsigned gbits, GxDyn& d, float gx_,
                                                    float gy_, float ox, float oy, float a0, float a1,
                                                    float* qacc) {
  GxLaneStep r;
  r.c0 = d.hc * a0; r.c1 = d.hs * a0; r.c2 = a1;
  const float last_dist = d.prev_dist, last_done = d.done;
  float px, py, r00, r10;
  gx_mjx_step_point<true>(p, d, r.c0, r.c1, r.c2, px, py, r00, r10, qacc);  // substeps == 1 here
  r.obs = gx_obs_lanes(p, sub, gmask, gbits, ox, oy, gx_, gy_, px, py, r00, r10);
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

// ---------------------------------------------------------------------------
// Engine.step + Engine.reset_done fused, with ONE lidar pass per env.
//
// A 4096-env launch is one wave: it lasts as long as its slowest warp, and nearly every launch
// of a stationary rollout has some env that finishes.  Running the step's observation and then
// the re-placed env's observation one after the other (gx_step_lanes + gx_reset_done_lanes)
// makes every such launch pay two lidar passes plus a dependent Philox draw and pool-row load.
// Here:
//   * the layout the env WOULD be re-placed with is drawn and loaded at the top of the kernel
//     (it depends only on the env id and its age), so it arrives during the dynamics;
//   * everything that decides `done` except the NaN/Inf guard needs no lidar (goal distance, jump,
//     timeout), and the guard itself reduces to the finiteness of ctrl / qpos / qvel whenever the
//     pose and the objects are finite and of sane magnitude (`sane`: then every ego vector, distance,
//     angle, sensor reading and alias weight is finite by construction);
//   * so a finishing env switches to its new pose and objects BEFORE the single lidar pass, which
//     then produces the reset observation directly; reward, cost, violations and flags are the
//     step's, computed at the step pose.
// When `sane` fails (non-finite or astronomically large state: never in a rollout, but reachable
// through gx_set_state) the env takes the reference order in gx_step_then_reset_lanes (out of line).
// Results are bit-identical either way (tests/test_parity_gpu.py).
// ---------------------------------------------------------------------------

struct GxSpecLayout {   // speculative re-placement, loaded at the top of the kernel
  float2 goal, robot, obj;
};

__device__ __forceinline__ GxSpecLayout gx_spec_layout(const float* lay, int sub) {
  const float2* l2 = reinterpret_cast<const float2*>(lay);  // goal, robot, hazard0..7
  GxSpecLayout sp;
  sp.goal = l2[0];
  sp.robot = l2[1];
  sp.obj = (sub >= 1 && sub <= 8) ? l2[1 + sub] : sp.goal;
  return sp;
}

__device__ __noinline__ void gx_step_then_reset_lanes(const gx_params& p, const GxState& s, int i, int sub, unsigned gmask,
                                                      bool live, const float* lay, GxDyn& d, float& gx_, float& gy_,
                                                      float& ox, float& oy, float a0, float a1, float* qacc, GxLaneStep& r) {
  r = gx_step_lanes(p, sub, gmask, gmask, d, gx_, gy_, ox, oy, a0, a1, qacc);
  if (r.out.done > 0.0f) {
    gx_reset_done_lanes(p, s, i, sub, gmask, live, lay, d, gx_, gy_, ox, oy, r.obs);
    r.c0 = 0.0f; r.c1 = 0.0f; r.c2 = 0.0f;
  }
}

__device__ __forceinline__ GxLaneStep gx_step_autoreset_lanes(const gx_params& p, const GxState& s, int i, int sub, unsigned gmask,
                                                              bool live, const float* lay, const GxSpecLayout& sp, GxDyn& d,
                                                              float& gx_, float& gy_, float& ox, float& oy, float a0, float a1,
                                                              float* qacc) {
  // (goal and object by reference: a finishing env leaves with its NEW goal / object, which the multi-step callers keep in
  // registers; the single-step kernel ignores them)
  GxLaneStep r;
  // exotic state: the whole env (16 lanes agree: pose is replicated, objects are checked by ballot) leaves the fast path
  const bool obj_ok = (gx_abs(ox) < GX_SANE_BOUND) & (gx_abs(oy) < GX_SANE_BOUND);
  const bool objs_ok = (__ballot_sync(0xffffffffu, obj_ok) & gmask) == gmask;
  const bool sane = objs_ok & (gx_abs(d.qx) < GX_SANE_BOUND) & (gx_abs(d.qy) < GX_SANE_BOUND) & !gx_notfinite(d.qth) &
                    (gx_abs(gx_) < GX_SANE_BOUND) & (gx_abs(gy_) < GX_SANE_BOUND) & (p.neg_exp_gain <= 0.0f);
  if (!sane) {
    GxDyn dd = d;
    float g0 = gx_, g1 = gy_, o0 = ox, o1 = oy;
    gx_step_then_reset_lanes(p, s, i, sub, gmask, live, lay, dd, g0, g1, o0, o1, a0, a1, qacc, r);
    d = dd;
    gx_ = g0; gy_ = g1; ox = o0; oy = o1;
    return r;
  }
  r.c0 = d.hc * a0; r.c1 = d.hs * a0; r.c2 = a1;
  const float last_dist = d.prev_dist, last_done = d.done;
  float px, py, r00, r10;
  gx_mjx_step_point<true>(p, d, r.c0, r.c1, r.c2, px, py, r00, r10, qacc);  // substeps == 1 here
  GxLaneObs& o = r.obs;
  // the step's goal distance, cost and violations, at the step pose
  {
    float gdx = gx_ - px, gdy = gy_ - py;
    o.dist_goal = gx_sqrt(gdx * gdx + gdy * gdy);
  }
  gx_cost_lanes(p, sub, 0xffffffffu, gmask, ox, oy, px, py, o.cost, o.violations);
  // NaN/Inf guard (engine.py:696-699): with a sane pose and sane objects the compass and both lidars are
  // finite, so the guard is decided by the remaining entries of the row
  const int bad = gx_notfinite(r.c0) | gx_notfinite(r.c1) | gx_notfinite(r.c2) |
                  gx_notfinite(d.qx) | gx_notfinite(d.qy) | gx_notfinite(d.qth) |
                  gx_notfinite(d.vx) | gx_notfinite(d.vy) | gx_notfinite(d.vw);
  const float dist = o.dist_goal;
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
  out.cost = o.cost;
  out.violations = o.violations;
  // a finishing env continues from its new layout: pose = layout, theta = 0 (R = I exactly: gx_kinematics(0)
  // returns (1, 0)), qvel = ctrl = 0; lag heading, prev_dist, steps and the done flag stay the step's
  // (engine.py:702-731 returns the un-stepped data and never clears _done)
  if (out.done > 0.0f) {
    gx_ = sp.goal.x; gy_ = sp.goal.y;
    d.qx = sp.robot.x; d.qy = sp.robot.y; d.qth = 0.0f; d.vx = 0.0f; d.vy = 0.0f; d.vw = 0.0f;
    ox = sp.obj.x; oy = sp.obj.y;
    px = d.qx; py = d.qy; r00 = 1.0f; r10 = 0.0f;
    r.c0 = 0.0f; r.c1 = 0.0f; r.c2 = 0.0f;
    if (live) {
      if (sub == 0) s.goal[i] = sp.goal;
      if (sub >= 1 && sub <= 8) {
        float2* pair = reinterpret_cast<float2*>(s.hz + (size_t)((sub - 1) >> 1) * s.cap + i);
        pair[(sub - 1) & 1] = sp.obj;
      }
    }
  }
  gx_lidar_lanes(p, sub, 0xffffffffu, gmask, ox, oy, px, py, r00, r10, o);
  return r;
}

// state write-back by three different lanes of the env (one 16-B or 8-B store each)
__device__ __forceinline__ void gx_store_dyn_lanes(const GxState& s, int i, int sub, const GxDyn& d) {
  if (sub == 0) s.d0[i] = make_float4(d.qx, d.qy, d.qth, d.vx);
  if (sub == 1) s.d1[i] = make_float4(d.vy, d.vw, d.hc, d.hs);
  if (sub == 2) s.d2[i] = gx_pack_d2(d);
}

#endif  // GX_LANES_CUH_
