// gx_device.cuh -- device-side transition function of the batched env step.
//
// What the reference computes with jax.jit(jax.vmap(mjx_step)) (engine.py:659-700)
// as dozens of XLA kernels, each intermediate round-tripping HBM, is here ONE
// device function operating on registers: action -> ctrl (convert_action,
// engine.py:672-685), mjx.step for the Point model (forward + implicit-damping
// Euler), lagged-pose observation (engine.py:738-778, 817-900), reward / done
// (:787-802), cost (:804-811), NaN/Inf guard (:696-699), timeout + step counter
// (:492-493).
//
// Arithmetic order is CANONICAL: every fp32 expression below is evaluated in the
// same order as oracle/gx_oracle.c (a line-by-line restatement of the reference),
// with no fused multiply-add except the explicit gx_fma() inside gx_math.h, so
// results are bit-identical to the oracle.  Compile with -fmad=false.
#ifndef GX_DEVICE_CUH_
#define GX_DEVICE_CUH_

#include <cuda_runtime.h>
#include <stdint.h>

#include "gx_math.h"
#include "gx_params.h"
#include "gx_philox.h"

// ---------------------------------------------------------------------------
// HBM state, structure of arrays.  `cap` = env_num rounded up to 32 so that
// every warp's slice of every array starts 128-B aligned.
//   d0  float4  qpos.x qpos.y qpos.theta qvel.x
//   d1  float4  qvel.y qvel.theta lag_cos lag_sin    (lag = R[0,0], R[1,0] of the
//                                                     last forward, engine.py:676)
//   d2  float4  prev_dist, steps, age_lo, age_hi
//               (sign bit of steps = Engine._done; steps == 0 whenever done,
//                engine.py:493, so -0.0f encodes it.  age = number of step() calls
//                this env has seen: the counter that keys its Philox draws.  It is
//                the same number for every env -- the reference keeps ONE key and
//                splits it once per step (engine.py:431) -- but it is stored per env
//                so that no launch carries a host-side count and no kernel needs a
//                grid-wide counter update: CUDA-graph replay stays correct for free.)
//   goal float2
//   hz  float4[nh2][cap]   two hazards per float4, pair-major (coalesced per pair)
//   ep  float4  ep_ret ep_cost ep_len cum_cost       (only touched when stats on)
//   hx, hf      history read by ego_vel_acc only (observe_vel / observe_acc, default off)
//   fin float4[2][cap]  per-env sums over FINISHED episodes: {n, ret, cost, len},
//               {ret^2, cost^2, -, -}  (read-modify-written only on a done step)
// Dynamic part 48 B, static part 8 + 8*nh B per env.
// ---------------------------------------------------------------------------
struct GxState {
  float4* d0;
  float4* d1;
  float4* d2;
  float2* goal;
  float4* hz;
  float4* ep;
  float4* fin;
  float4* hx;     // xpos_xy of Engine._data, xpos_xy of Engine._last_data  (observe_vel/acc only)
  uint32_t* hf;   // bit 0: Engine._last_done, bits 1-2: history depth       (observe_vel/acc only)
  int cap;
  int nh2;  // ceil(nh / 2)
};

// Programmatic dependent launch (guardx.cu: gx_launch): wait for every grid before this one in the stream to complete and
// flush -- BEFORE the first access to global memory --, then let the grid after this one be set up.  Both are no-ops
// for a launch without the attribute.
// Order measured at 4096 envs (us per step, no PDL 5.63): wait then release 5.37; release right before the stores 5.88;
// release BEFORE the wait (several grids pile up on the SMs) 6.90.
__device__ __forceinline__ void gx_pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

struct GxDyn {
  float qx, qy, qth, vx, vy, vw;  // qpos, qvel (post-step values after a step)
  float hc, hs;                   // lagged heading (cos, sin)
  float prev_dist;                // goal_pos(last_data)
  float steps;                    // Engine._steps
  float done;                     // Engine._done (0/1)
  uint64_t age;                   // step() calls seen so far (keys the RNG draws)
};

__device__ __forceinline__ void gx_load_dyn(const GxState& s, int i, GxDyn& d) {
  float4 a = s.d0[i];
  float4 b = s.d1[i];
  float4 c = s.d2[i];
  d.qx = a.x; d.qy = a.y; d.qth = a.z; d.vx = a.w;
  d.vy = b.x; d.vw = b.y; d.hc = b.z; d.hs = b.w;
  d.prev_dist = c.x;
  d.done = (__float_as_uint(c.y) >> 31) ? 1.0f : 0.0f;
  d.steps = fabsf(c.y);
  d.age = (uint64_t)__float_as_uint(c.z) | ((uint64_t)__float_as_uint(c.w) << 32);
}

__device__ __forceinline__ float4 gx_pack_d2(const GxDyn& d) {
  return make_float4(d.prev_dist, d.done > 0.0f ? -d.steps : d.steps, __uint_as_float((uint32_t)d.age),
                     __uint_as_float((uint32_t)(d.age >> 32)));
}

__device__ __forceinline__ void gx_store_dyn(const GxState& s, int i, const GxDyn& d) {
  s.d0[i] = make_float4(d.qx, d.qy, d.qth, d.vx);
  s.d1[i] = make_float4(d.vy, d.vw, d.hc, d.hs);
  s.d2[i] = gx_pack_d2(d);
}

// mjx kinematics of the robot body (quaternion path, A4): R[0,0], R[1,0].
__device__ __forceinline__ void gx_kinematics(float th, float& r00, float& r10) {
  float half = th * 0.5f;
  float s, c;
  gx_sincos(half, &s, &c);
  float n = gx_sqrt(c * c + s * s);
  float w = gx_div(c, n), z = gx_div(s, n);
  float ww = w * w, zz = z * z, wz = w * z;
  r00 = ww - zz;
  r10 = 2.0f * wz;
}

__device__ __forceinline__ void gx_solve_robot_block(float d0, float d1, float mi, float a, float b,
                                                     float f0, float f1, float f2, float& x0,
                                                     float& x1, float& x2) {
  float t0 = gx_div(f0, d0), t1 = gx_div(f1, d1);
  float schur = mi - (gx_div(a * a, d0) + gx_div(b * b, d1));
  float alpha = gx_div(f2 - (a * t0 + b * t1), schur);
  x0 = t0 - gx_div(a * alpha, d0);
  x1 = t1 - gx_div(b * alpha, d1);
  x2 = alpha;
}

// FAST (template flag used by the hot kernels): the launch code has checked that
// ctrllimited, forcelimited, lidar_alias and the exponential lidar are all on and
// physics_steps_per_control_step == 1 (the Goal_Point_8Hazards defaults), so those
// uniform branches disappear at compile time.
template <bool FAST>
__device__ __forceinline__ float gx_actuator(const gx_params& p, int i, float ctrl, float v) {
  float u = ctrl;
  if (FAST || p.ctrllimited) u = gx_clip(u, p.ctrl_lo, p.ctrl_hi);
  float act_vel = p.gear[i] * v;
  float force = p.gain * u + p.bias2 * act_vel;
  if (FAST || p.forcelimited) force = gx_clip(force, p.force_lo, p.force_hi);
  return p.gear[i] * force;
}

// One mjx.step for the Point model.  On return (px,py,r00,r10) is the pose
// snapshot of the forward pass (PRE-step qpos) and d.q*/d.v* are post-step.
template <bool FAST = false>
__device__ __forceinline__ void gx_mjx_step_point(const gx_params& p, GxDyn& d, float c0, float c1,
                                                  float c2, float& px, float& py, float& r00,
                                                  float& r10, float* qacc) {
  gx_kinematics(d.qth, r00, r10);
  px = d.qx;
  py = d.qy;
  float m02 = -(p.mc * r10), m12 = p.mc * r00;
  float w2 = d.vw * d.vw;
  float b0 = -(p.mc * r00) * w2, b1 = -(p.mc * r10) * w2;
  float f0 = (-(p.damp[0] * d.vx) - b0) + gx_actuator<FAST>(p, 0, c0, d.vx);
  float f1 = (-(p.damp[1] * d.vy) - b1) + gx_actuator<FAST>(p, 1, c1, d.vy);
  float f2 = (-(p.damp[2] * d.vw) - 0.0f) + gx_actuator<FAST>(p, 2, c2, d.vw);
  if (qacc) gx_solve_robot_block(p.mass, p.mass, p.inertia, m02, m12, f0, f1, f2, qacc[0], qacc[1], qacc[2]);
  float a0, a1, a2;
  gx_solve_robot_block(p.mh0, p.mh1, p.mh2, m02, m12, f0, f1, f2, a0, a1, a2);
  d.vx = d.vx + p.dt * a0;
  d.vy = d.vy + p.dt * a1;
  d.vw = d.vw + p.dt * a2;
  d.qx = d.qx + p.dt * d.vx;
  d.qy = d.qy + p.dt * d.vy;
  d.qth = d.qth + p.dt * d.vw;
}

// One lidar return (engine.py:877-899) for an object at (ox,oy) seen from pose
// (px,py,R): bin index as the reference computes it (may be == nb), the sensor
// reading and its two aliased neighbours.
struct GxRay {
  int bin;
  float sensor, up, down;  // obs[bin], obs[bin+1], obs[bin-1] contributions
};

__device__ __forceinline__ void gx_ego(float ox, float oy, float px, float py, float r00, float r10,
                                       float& ex, float& ey) {
  float dx = ox - px, dy = oy - py;
  ex = dx * r00 + dy * r10;
  ey = dx * (-r10) + dy * r00;
}

template <bool FAST = false>
__device__ __forceinline__ GxRay gx_lidar_ray(const gx_params& p, float ex, float ey) {
  GxRay r;
  float dist = gx_sqrt(ex * ex + ey * ey);
  float angle = gx_atan2(ey, ex);
  if (angle < 0.0f) angle = angle + GX_TWO_PI_F;
  float q = gx_div(angle, p.bin_size);
  r.bin = (q >= 0.0f && q < 1.0e9f) ? (int)q : 0;
  float bin_angle = p.bin_size * (float)r.bin;
  if (FAST || p.use_exp) {
    r.sensor = gx_exp(p.neg_exp_gain * dist);
  } else {
    r.sensor = gx_div(gx_max(0.0f, p.lidar_max_dist - dist), p.lidar_max_dist);
  }
  float alias = gx_div(angle - bin_angle, p.bin_size);
  r.up = alias * r.sensor;
  r.down = (1.0f - alias) * r.sensor;
  return r;
}

// Scatter one ray into a thread-private row of nb bins held in shared memory,
// with JAX's index semantics (gather clamps, out-of-bounds scatter is dropped).
// Returns nonzero if a non-finite value was written.
__device__ __forceinline__ int gx_lidar_scatter(const gx_params& p, const GxRay& r, float* bins) {
  const int nb = p.nb;
  int bad = 0;
  if (r.bin < nb) {
    float v = gx_max(bins[r.bin], r.sensor);
    bins[r.bin] = v;
    bad |= gx_notfinite(v);
  }
  if (p.alias) {
    int bp = r.bin + 1; if (bp >= nb) bp -= nb; if (bp >= nb) bp -= nb;  // (bin+1) % nb, bin <= nb
    int bm = r.bin - 1; if (bm < 0) bm += nb;
    float v = gx_max(bins[bp], r.up);
    bins[bp] = v;
    bad |= gx_notfinite(v);
    v = gx_max(bins[bm], r.down);
    bins[bm] = v;
    bad |= gx_notfinite(v);
  }
  return bad;
}

// ---------------------------------------------------------------------------
// Observation row builder for one env per thread.  `row` points at the
// thread's obs_dim floats inside the warp's shared-memory tile.
// Geometry comes from the lagged pose (px,py,r00,r10); qpos/qvel are d's
// current values (engine.py:754-766).  Returns the NaN/Inf flag of the row and
// fills dist_goal, cost, violations, bins (optional).
// ---------------------------------------------------------------------------
// |pose|, |object| below this: squares of ego components stay below FLT_MAX with a wide margin, so every
// distance / angle / sensor reading of the observation is finite by construction (fused auto-reset paths)
#define GX_SANE_BOUND 1.0e15f

struct GxObsOut {
  float dist_goal, cost;
  int violations, bad;
};

template <int NH_CT>
__device__ __forceinline__ GxObsOut gx_obs_row(const gx_params& p, const GxState& s, int i,
                                               float gx_, float gy_, const float4* hz_regs,
                                               const GxDyn& d, float c0, float c1, float c2,
                                               float px, float py, float r00, float r10, float* row,
                                               int32_t* aux_bins) {
  GxObsOut o;
  o.bad = 0;
  const int nb = p.nb;
  // vel / acc default to 0 = obs(data, None, None, None, None) (engine.py:656, 726);
  // gx_vel_acc overwrites them on a step
  if (p.off_acc >= 0) { row[p.off_acc] = 0.0f; row[p.off_acc + 1] = 0.0f; }
  if (p.off_vel >= 0) { row[p.off_vel] = 0.0f; row[p.off_vel + 1] = 0.0f; }
  if (p.off_ctrl >= 0) {
    row[p.off_ctrl + 0] = c0; row[p.off_ctrl + 1] = c1; row[p.off_ctrl + 2] = c2;
    o.bad |= gx_notfinite(c0) | gx_notfinite(c1) | gx_notfinite(c2);
  }
  float gex, gey;
  gx_ego(gx_, gy_, px, py, r00, r10, gex, gey);
  if (p.off_compass >= 0) {
    row[p.off_compass + 0] = gex; row[p.off_compass + 1] = gey;
    o.bad |= gx_notfinite(gex) | gx_notfinite(gey);
  }
  // goal_pos (engine.py:780-785)
  {
    float dx = gx_ - px, dy = gy_ - py;
    o.dist_goal = gx_sqrt(dx * dx + dy * dy);
  }
  GxRay gr = gx_lidar_ray(p, gex, gey);
  if (aux_bins) aux_bins[0] = gr.bin;
  if (p.off_glidar >= 0) {
    float* bins = row + p.off_glidar;
    for (int k = 0; k < nb; ++k) bins[k] = 0.0f;
    o.bad |= gx_lidar_scatter(p, gr, bins);
  }
  float* hb = p.off_hlidar >= 0 ? row + p.off_hlidar : nullptr;
  if (hb) for (int k = 0; k < nb; ++k) hb[k] = 0.0f;
  float cost = 0.0f;
  int viol = 0;
  const int nh = NH_CT ? NH_CT : p.nh;
  // one hazard: lidar return (engine.py:872-899) and cost term (:808-810)
  auto hazard = [&](float hx, float hy, int idx) {
    float ex, ey;
    gx_ego(hx, hy, px, py, r00, r10, ex, ey);
    if (hb) {
      GxRay r = gx_lidar_ray(p, ex, ey);
      if (aux_bins) aux_bins[1 + idx] = r.bin;
      o.bad |= gx_lidar_scatter(p, r, hb);
    } else if (aux_bins) {
      aux_bins[1 + idx] = gx_lidar_ray(p, ex, ey).bin;
    }
    float dx = hx - px, dy = hy - py;
    float dist = gx_sqrt(dx * dx + dy * dy);
    float term = p.hazards_size - gx_min(dist, p.hazards_size);
    if (term > 0.0f) ++viol;
    cost = cost + term;
  };
  if constexpr (NH_CT != 0) {
    // NOT unrolled over the hazard pairs: two copies of the ray code instead of one per hazard.  Unrolled, this kernel was
    // 9184 instructions (147 KB: more than the instruction cache holds; ncu: 2 cycles of `no_instruction` stall per issued
    // instruction); rolled, the hazards are indexed in a 64-byte per-thread array that lives in local memory (L1 hits).
    // Fresh batches, us per step at 16 384 / 65 536 / 262 144 / 2^20 envs: unrolled 8.1 / 11.2 / 29.5 / 99.3, rolled
    // 7.7 / 10.9 / 28.1 / 87.9 (one hazard per iteration: 8.1 / 11.3 / 29.1 / 90.9).  -DGX_HZ_UNROLLED restores the old loop.
#if defined(GX_HZ_UNROLLED)
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int k = 0; k < (NH_CT + 1) / 2; ++k) {
      const float4 h = hz_regs[k];
      hazard(h.x, h.y, 2 * k);
      if (2 * k + 1 < NH_CT) hazard(h.z, h.w, 2 * k + 1);
    }
  } else {
#pragma unroll 1
    for (int k = 0; k < s.nh2; ++k) {
      const float4 h = s.hz[(size_t)k * s.cap + i];
      hazard(h.x, h.y, 2 * k);
      if (2 * k + 1 < nh) hazard(h.z, h.w, 2 * k + 1);
    }
  }
  o.cost = cost;
  o.violations = viol;
  if (p.off_qpos >= 0) {
    row[p.off_qpos + 0] = d.qx; row[p.off_qpos + 1] = d.qy; row[p.off_qpos + 2] = d.qth;
    o.bad |= gx_notfinite(d.qx) | gx_notfinite(d.qy) | gx_notfinite(d.qth);
  }
  if (p.off_qvel >= 0) {
    row[p.off_qvel + 0] = d.vx; row[p.off_qvel + 1] = d.vy; row[p.off_qvel + 2] = d.vw;
    o.bad |= gx_notfinite(d.vx) | gx_notfinite(d.vy) | gx_notfinite(d.vw);
  }
  return o;
}

// ego_vel_acc (engine.py:902-929): finite-difference velocity / acceleration of
// the robot xpos in the robot frame, with the done-masking and the None-ness of
// last_done / last_last_done on the first two steps.  Also performs update_data's
// history shift for these fields (engine.py:426-430).  Returns the NaN/Inf flag.
__device__ __forceinline__ int gx_vel_acc(const gx_params& p, const GxState& s, int i, float px, float py,
                                          float r00, float r10, float last_done, float* row) {
  const float4 h = s.hx[i];
  const uint32_t f = s.hf[i];
  const float last_last_done = (f & 1u) ? 1.0f : 0.0f;
  const int hist = (int)((f >> 1) & 3u);
  const float lx = h.x, ly = h.y;  // last_data.xpos (the data this step started from)
  const float llx = hist >= 2 ? h.z : lx, lly = hist >= 2 ? h.w : ly;
  float plx = px, ply = py, pllx = px, plly = py;
  if (hist >= 1) {
    if (!(last_done > 0.0f)) { plx = lx; ply = ly; }
    if (hist >= 2) {
      if (last_last_done + last_done > 0.0f) { pllx = plx; plly = ply; }
      else { pllx = llx; plly = lly; }
    }
  }
  const float vwx = gx_div(px - plx, p.ctrl_dt), vwy = gx_div(py - ply, p.ctrl_dt);
  const float lvx = gx_div(plx - pllx, p.ctrl_dt), lvy = gx_div(ply - plly, p.ctrl_dt);
  const float awx = gx_div(vwx - lvx, p.ctrl_dt), awy = gx_div(vwy - lvy, p.ctrl_dt);
  const float vx = vwx * r00 + vwy * r10, vy = vwx * (-r10) + vwy * r00;
  const float ax = awx * r00 + awy * r10, ay = awx * (-r10) + awy * r00;
  int bad = 0;
  if (p.off_acc >= 0) { row[p.off_acc] = ax; row[p.off_acc + 1] = ay; bad |= gx_notfinite(ax) | gx_notfinite(ay); }
  if (p.off_vel >= 0) { row[p.off_vel] = vx; row[p.off_vel + 1] = vy; bad |= gx_notfinite(vx) | gx_notfinite(vy); }
  s.hx[i] = make_float4(px, py, lx, ly);
  s.hf[i] = (last_done > 0.0f ? 1u : 0u) | ((uint32_t)(hist < 2 ? hist + 1 : 2) << 1);
  return bad;
}

struct GxStepOut {
  float reward, done, cost;
  int flags, violations;
};

// The whole transition for one env (mjx_step + Engine.step tail).
template <int NH_CT>
__device__ __forceinline__ GxStepOut gx_step_env(const gx_params& p, const GxState& s, int i,
                                                 GxDyn& d, float gx_, float gy_,
                                                 const float4* hz_regs, float a0, float a1,
                                                 float* row, float* qacc, int32_t* aux_bins) {
  // convert_action (engine.py:672-685)
  float c0 = d.hc * a0, c1 = d.hs * a0, c2 = a1;
  float last_dist = d.prev_dist, last_done = d.done;
  float px, py, r00, r10;
  for (int k = 0; k < p.substeps; ++k) gx_mjx_step_point(p, d, c0, c1, c2, px, py, r00, r10, qacc);
  GxObsOut o = gx_obs_row<NH_CT>(p, s, i, gx_, gy_, hz_regs, d, c0, c1, c2, px, py, r00, r10, row, aux_bins);
  if ((p.off_vel >= 0) | (p.off_acc >= 0)) o.bad |= gx_vel_acc(p, s, i, px, py, r00, r10, last_done, row);
  // reward_done (engine.py:787-802)
  float dist = o.dist_goal;
  float last = last_done > 0.0f ? dist : last_dist;
  float d_dist = last - dist;
  GxStepOut out;
  out.reward = d_dist * p.reward_distance;
  out.done = dist < p.goal_size ? 1.0f : 0.0f;
  out.flags = out.done > 0.0f ? GX_AUX_GOAL : 0;
  if (gx_abs(d_dist) > 1.0f) { out.done = 1.0f; out.reward = 0.0f; out.flags |= GX_AUX_JUMP; }
  if (o.bad) { out.reward = 0.0f; out.done = 1.0f; out.flags |= GX_AUX_GUARD; }
  if (d.steps > p.num_steps) { out.done = 1.0f; out.flags |= GX_AUX_TIMEOUT; }
  d.steps = out.done > 0.0f ? 0.0f : d.steps + 1.0f;
  d.done = out.done;
  d.prev_dist = dist;
  d.hc = r00;
  d.hs = r10;
  d.age += 1;  // update_data splits the key once per step (engine.py:431)
  out.cost = o.cost;
  out.violations = o.violations;
  return out;
}

// Observation of a freshly placed env (mjx_reset engine.py:644-657, and the
// obs_reset of mjx_reset_done engine.py:719-726): pose = layout, theta = 0,
// qvel = ctrl = 0.  The throw-away step of reset_done leaves qpos/qvel exactly
// unchanged for the Point (zero force on a body at rest), so one function serves both.
template <int NH_CT>
__device__ __forceinline__ float gx_reset_row(const gx_params& p, const GxState& s, int i,
                                              const GxDyn& d, float gx_, float gy_,
                                              const float4* hz_regs, float* row) {
  float r00, r10;
  gx_kinematics(d.qth, r00, r10);
  GxObsOut o = gx_obs_row<NH_CT>(p, s, i, gx_, gy_, hz_regs, d, 0.0f, 0.0f, 0.0f, d.qx, d.qy, r00, r10, row, nullptr);
  return o.dist_goal;
}

// Cooperative copy of a warp's obs tile (rows x obs_dim floats, contiguous in
// shared memory) to global memory.  A full warp's tile is 128*obs_dim bytes and
// starts 128-B aligned, so it goes out as float4.
__device__ __forceinline__ void gx_store_tile(const float* tile, float* gdst, int rows, int obs_dim,
                                              int lane) {
  __syncwarp();
  const int total = rows * obs_dim;
  if (rows == 32 && ((reinterpret_cast<uintptr_t>(gdst) & 15u) == 0)) {
    const float4* src4 = reinterpret_cast<const float4*>(tile);
    float4* dst4 = reinterpret_cast<float4*>(gdst);
    const int n4 = total >> 2;
    for (int k = lane; k < n4; k += 32) dst4[k] = src4[k];
  } else {
    for (int k = lane; k < total; k += 32) gdst[k] = tile[k];
  }
  __syncwarp();
}

// The same tile as ONE asynchronous bulk copy (cp.async.bulk shared -> global, the non-tensor TMA path): the
// 32 x obs_dim floats of a full warp leave with a single instruction instead of 344 LDS.128 + 344 STG.128 (30 % of
// the step kernel's instruction stream at 2^20 envs, ncu).  Needs 16-B aligned source, destination and size.
// Every lane has written its row with ordinary stores: fence them towards the async proxy, meet, then lane 0 issues
// the copy and waits until the source has been READ (the block's shared memory may be handed to the next block as
// soon as the last thread exits).
__device__ __forceinline__ void gx_store_tile_bulk(const float* tile, float* gdst, unsigned bytes, int lane) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    const unsigned src = (unsigned)__cvta_generic_to_shared(tile);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}

__device__ __forceinline__ void gx_store_tile_any(const float* tile, float* gdst, int rows, int obs_dim, int lane, bool bulk) {
  const unsigned bytes = (unsigned)(rows * obs_dim) * 4u;
  if (bulk && rows == 32 && (bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(gdst) & 15u) == 0) &&
      ((__cvta_generic_to_shared(tile) & 15u) == 0))
    gx_store_tile_bulk(tile, gdst, bytes, lane);
  else
    gx_store_tile(tile, gdst, rows, obs_dim, lane);
}

// ---------------------------------------------------------------------------
// shared by the Point kernels (guardx.cu) and the articulated kernels (gx_art_*.cu)
// ---------------------------------------------------------------------------
extern __shared__ __align__(16) float gx_smem[];

// layout row of env i: caller-supplied, or a pool slot drawn with the env's own counter
__device__ __forceinline__ const float* gx_pick_layout(const gx_params& p, const float* pool,
                                                       const float* layout, int i, uint64_t t_global,
                                                       uint32_t stream) {
  if (layout) return layout + (size_t)i * p.lay_dim;
  const uint64_t g = p.env_id_offset + (uint64_t)i;
  gx_u32x4 r = gx_philox4x32_10((uint32_t)g, (uint32_t)t_global, (uint32_t)(t_global >> 32), stream, p.seed_lo, p.seed_hi);
  const uint32_t idx = gx_randint(r.v[0], (uint32_t)p.pool_size);
  return pool + (size_t)idx * p.lay_dim;
}

// a_t ~ U(-1,1) for actuator a of env g at step `age` (the synthetic rollouts of
// BASELINE.json; same stream as the Point kernels: word 3 = GX_STREAM_ACTION + block)
__device__ __forceinline__ float gx_art_synth_action(const gx_params& p, uint64_t g, uint64_t age, int a) {
  gx_u32x4 u = gx_philox4x32_10((uint32_t)g, (uint32_t)age, (uint32_t)(age >> 32), GX_STREAM_ACTION + (uint32_t)(a >> 2),
                                p.seed_lo, p.seed_hi);
  return gx_uniform(u.v[a & 3], -1.0f, 1.0f);
}

// hazard h of env i: from a layout row (while re-placing) or from the state arrays
__device__ __forceinline__ void gx_art_hazard(const GxState& s, const float* lay, int i, int h, float& hx, float& hy) {
  if (lay) { hx = lay[4 + 2 * h]; hy = lay[5 + 2 * h]; return; }
  const float4 v = s.hz[(size_t)(h >> 1) * s.cap + i];
  hx = (h & 1) ? v.z : v.x;
  hy = (h & 1) ? v.w : v.y;
}

// ---- episode statistics (ppo_train.py:370-376, 405-411 on the host) --------
// Per-env accumulators only -- no atomics and no grid-wide reduction in the step
// (4096 same-address atomics per launch cost more than the step itself).  The
// sums over envs are taken, deterministically, when the statistics are read.
// Split in two so that latency-bound kernels can issue the accumulator load at their top (it is
// the only load a step would otherwise issue at its very end, a full DRAM round trip on the
// critical path of every launch).  The sums over finished episodes are updated with
// fire-and-forget vector reductions (red.global.add.v4.f32): each env's slot has a single writer
// per launch, so the result is the same single fp32 addition, and nothing waits for it.
__device__ __forceinline__ float4 gx_stats_load(const GxState& s, int i) { return s.ep[i]; }

__device__ __forceinline__ void gx_stats_commit(const GxState& s, int i, float4 e, float reward, float cost, float done) {
  e.x += reward; e.y += cost; e.z += 1.0f; e.w += cost;
  if (done > 0.0f) {
    atomicAdd(&s.fin[i], make_float4(1.0f, e.x, e.y, e.z));
    atomicAdd(&s.fin[(size_t)s.cap + i], make_float4(e.x * e.x, e.y * e.y, 0.0f, 0.0f));
    e.x = 0.f; e.y = 0.f; e.z = 0.f;
  }
  s.ep[i] = e;
}

__device__ __forceinline__ void gx_stats_step(const GxState& s, int i, float reward, float cost, float done) {
  gx_stats_commit(s, i, gx_stats_load(s, i), reward, cost, done);
}

#endif  // GX_DEVICE_CUH_
