// gx_task.cuh -- device side of the config-3 extension (include/guardx_task.h): Button / Push tasks, the Car robot,
// vases, pillars, gremlins.  NO REFERENCE, PARITY UNPINNED: the semantics are the specification in DESIGN.md section 8;
// the checker is oracle/task_oracle.py (independent numpy restatement of that specification).
//
// One thread per env, the whole transition fused in one kernel like the Point path (gx_device.cuh), whose conventions it
// keeps: geometry for observation / reward / cost is the pose snapshot of the step's forward pass (PRE-step pose,
// engine.py:676-680, 754-758), qpos / qvel in the observation are post-step, the episode ends at the goal, the dense
// hazard cost, the |d dist| > 1 guard, the NaN / Inf guard, the timeout, the per-env step count that keys every
// Philox draw.  State is structure-of-arrays in one buffer: slot k of env i at S[k * cap + i]:
//    0..5   qx qy qth vx vy vw          6,7  lagged heading (cos, sin)      8  prev task distance   9  prev robot-box distance
//    10     done    11  steps           12,13 age (bit patterns)             14 goal button index    15 (spare)
//    16 + 2 o, 17 + 2 o   xy of object o (placement order: goal, box, buttons, hazards, vases, pillars, gremlin CENTRES)
//    16 + 2 nobj + g      phase of gremlin g at episode time 0
// gxt_get_state / gxt_set_state rows have the same order (age slots are skipped on set).
// Layout rows (pool, caller layouts): xy per object in placement order, robot xy, gremlin phases.
#ifndef GX_TASK_CUH_
#define GX_TASK_CUH_

#include "gx_device.cuh"
#include "../../include/guardx_task.h"

#define GXT_SLOT_OBJ 16
#define GX_STREAM_BUTTON 0x42544e21u /* goal button draw */

struct gxt_params {
  gx_params pt;                 // Point dynamics + lidar constants (nb, alias, use_exp, neg_exp_gain, bin_size, lidar_max_dist), seeds
  int32_t n, cap, task, robot, nobj, ng, obs_dim, state_dim, lay_dim, pool_size, tries, nu;
  int32_t first[GXT_NCLASS], count[GXT_NCLASS];
  float size[GXT_NCLASS];
  float lo[GXT_NCLASS][2], hi[GXT_NCLASS][2], keep[GXT_NCLASS], margin, min_robot_goal;
  float num_steps, reward_distance, reward_box_dist, reward_box_goal, hazards_cost_scale;
  float vases_contact_cost, pillars_cost, buttons_cost, gremlins_contact_cost, gremlins_dist_threshold, gremlins_dist_cost, gremlins_travel;
  float goal_size, dt;
  int32_t off_box_compass, off_box_lidar, off_buttons_lidar, off_ctrl, off_goal_compass, off_goal_lidar, off_gremlins_lidar,
      off_hazards_lidar, off_pillars_lidar, off_qpos, off_qvel, off_vases_lidar;
  float car_mass, car_inertia, car_wheel_r, car_half_track, car_gear, car_mh_lin, car_mh_ang;
  uint64_t cls_bits[GXT_MAX_OBJECTS / 16];   // class of object o in 4 bits: (cls_bits[o >> 4] >> (4 * (o & 15))) & 15 (filled by gxt_lower)
};

__host__ __device__ inline int gxt_class_scan(const gxt_params& p, int o) {
  int c = 0;
  for (int k = 1; k < GXT_ROBOT; ++k) c = (o >= p.first[k] && p.count[k] > 0) ? k : c;
  return c;
}
// class of object o: a table lookup (the scan over the classes was 4 % of the lanes kernel's instructions)
__device__ __forceinline__ int gxt_class_of_obj(const gxt_params& p, int o) {
  const uint64_t w = (o >> 4) == 0 ? p.cls_bits[0] : ((o >> 4) == 1 ? p.cls_bits[1] : ((o >> 4) == 2 ? p.cls_bits[2] : p.cls_bits[3]));
  return (int)((w >> (4 * (o & 15))) & 15u);
}

// ---- robots ----------------------------------------------------------------------------------------------------
// Car: planar differential drive.  ctrl = wheel motor commands (left, right) in [-1, 1]; torque = gear * ctrl per wheel;
// traction F = (tl + tr) / r, yaw torque N = b (tr - tl) / r; no lateral slip (the lateral velocity is removed);
// semi-implicit Euler with implicit viscous damping, as mjx's euler does for joint damping.
__device__ __forceinline__ void gxt_car_step(const gxt_params& p, GxDyn& d, float a0, float a1, float& c0, float& c1, float& px,
                                             float& py, float& r00, float& r10) {
  c0 = gx_clip(a0, -1.0f, 1.0f);
  c1 = gx_clip(a1, -1.0f, 1.0f);
  gx_kinematics(d.qth, r00, r10);
  px = d.qx; py = d.qy;
  const float tl = p.car_gear * c0, tr = p.car_gear * c1;
  const float F = gx_div(tl + tr, p.car_wheel_r);
  const float N = gx_div(p.car_half_track * (tr - tl), p.car_wheel_r);
  const float u = d.vx * r00 + d.vy * r10;                       // forward speed, body frame
  const float un = gx_div(p.car_mass * u + p.dt * F, p.car_mh_lin);
  const float wn = gx_div(p.car_inertia * d.vw + p.dt * N, p.car_mh_ang);
  d.vx = un * r00; d.vy = un * r10; d.vw = wn;
  d.qx = d.qx + p.dt * d.vx;
  d.qy = d.qy + p.dt * d.vy;
  d.qth = d.qth + p.dt * d.vw;
}

// ---- observation -----------------------------------------------------------------------------------------------
struct GxtGeom {   // pose snapshot of the forward pass + where the objects are at that time
  float px, py, r00, r10, time;   // time = steps * dt (gremlin phase)
};

__device__ __forceinline__ void gxt_obj_xy(const gxt_params& p, const float* __restrict__ S, int i, int o, float& x, float& y) {
  x = S[(size_t)(GXT_SLOT_OBJ + 2 * o) * p.cap + i];
  y = S[(size_t)(GXT_SLOT_OBJ + 2 * o + 1) * p.cap + i];
}

// gremlin g: circle of radius gremlins_travel about its centre, 1 rad / s
__device__ __forceinline__ void gxt_gremlin_xy(const gxt_params& p, const float* __restrict__ S, int i, int g, float time, float& x, float& y) {
  float cx, cy;
  gxt_obj_xy(p, S, i, p.first[GXT_GREMLIN] + g, cx, cy);
  const float ph = S[(size_t)(GXT_SLOT_OBJ + 2 * p.nobj + g) * p.cap + i] + time;
  float sn, cs;
  gx_sincos(ph, &sn, &cs);
  x = cx + p.gremlins_travel * cs;
  y = cy + p.gremlins_travel * sn;
}

// lidar of one class into `bins` (nb floats of the row): zero, then one ray per object (engine.py:846-900)
__device__ __forceinline__ int gxt_lidar_class(const gxt_params& p, const float* __restrict__ S, int i, int cls, const GxtGeom& g, float* bins) {
  const int nb = p.pt.nb;
  int bad = 0;
  for (int k = 0; k < nb; ++k) bins[k] = 0.0f;
#pragma unroll 1
  for (int k = 0; k < p.count[cls]; ++k) {
    float ox, oy, ex, ey;
    if (cls == GXT_GREMLIN) gxt_gremlin_xy(p, S, i, k, g.time, ox, oy);
    else gxt_obj_xy(p, S, i, p.first[cls] + k, ox, oy);
    gx_ego(ox, oy, g.px, g.py, g.r00, g.r10, ex, ey);
    const GxRay r = gx_lidar_ray(p.pt, ex, ey);
    bad |= gx_lidar_scatter(p.pt, r, bins);
  }
  return bad;
}

struct GxtObs {
  float dist_task, dist_box, cost;
  int violations, contacts, bad, goal_bin;
};

// the task's target: goal (goal / push) or the goal button (button)
__device__ __forceinline__ void gxt_target_xy(const gxt_params& p, const float* __restrict__ S, int i, float goal_button, float& x, float& y) {
  if (p.task == GXT_TASK_BUTTON) {
    int b = (int)goal_button;
    b = b < 0 ? 0 : (b >= p.count[GXT_BUTTON] ? p.count[GXT_BUTTON] - 1 : b);
    gxt_obj_xy(p, S, i, p.first[GXT_BUTTON] + b, x, y);
  } else {
    gxt_obj_xy(p, S, i, p.first[GXT_GOAL], x, y);
  }
}

// observation row + task distances + cost, all from the pose snapshot `g`; qpos / qvel are d's (post-step) values
__device__ __forceinline__ GxtObs gxt_obs_row(const gxt_params& p, const float* __restrict__ S, int i, const GxtGeom& g, const GxDyn& d,
                                              float goal_button, float c0, float c1, float c2, float* row) {
  GxtObs o;
  o.bad = 0; o.violations = 0; o.contacts = 0; o.goal_bin = 0;
  const int nb = p.pt.nb;
  const float rr = p.size[GXT_ROBOT];
  float tx, ty;
  gxt_target_xy(p, S, i, goal_button, tx, ty);
  float bx = 0.0f, by = 0.0f;
  if (p.count[GXT_BOX] > 0) gxt_obj_xy(p, S, i, p.first[GXT_BOX], bx, by);
  // ctrl
  if (p.off_ctrl >= 0) {
    row[p.off_ctrl] = c0; row[p.off_ctrl + 1] = c1;
    o.bad |= gx_notfinite(c0) | gx_notfinite(c1);
    if (p.nu == 3) { row[p.off_ctrl + 2] = c2; o.bad |= gx_notfinite(c2); }
  }
  // goal compass / lidar (the goal button in the button task)
  {
    float ex, ey;
    gx_ego(tx, ty, g.px, g.py, g.r00, g.r10, ex, ey);
    if (p.off_goal_compass >= 0) { row[p.off_goal_compass] = ex; row[p.off_goal_compass + 1] = ey; o.bad |= gx_notfinite(ex) | gx_notfinite(ey); }
    const GxRay r = gx_lidar_ray(p.pt, ex, ey);
    o.goal_bin = r.bin;
    if (p.off_goal_lidar >= 0) {
      float* bins = row + p.off_goal_lidar;
      for (int k = 0; k < nb; ++k) bins[k] = 0.0f;
      o.bad |= gx_lidar_scatter(p.pt, r, bins);
    }
  }
  if (p.count[GXT_BOX] > 0) {
    float ex, ey;
    gx_ego(bx, by, g.px, g.py, g.r00, g.r10, ex, ey);
    if (p.off_box_compass >= 0) { row[p.off_box_compass] = ex; row[p.off_box_compass + 1] = ey; o.bad |= gx_notfinite(ex) | gx_notfinite(ey); }
    if (p.off_box_lidar >= 0) {
      float* bins = row + p.off_box_lidar;
      for (int k = 0; k < nb; ++k) bins[k] = 0.0f;
      o.bad |= gx_lidar_scatter(p.pt, gx_lidar_ray(p.pt, ex, ey), bins);
    }
  }
  if (p.off_buttons_lidar >= 0) o.bad |= gxt_lidar_class(p, S, i, GXT_BUTTON, g, row + p.off_buttons_lidar);
  if (p.off_gremlins_lidar >= 0) o.bad |= gxt_lidar_class(p, S, i, GXT_GREMLIN, g, row + p.off_gremlins_lidar);
  if (p.off_hazards_lidar >= 0) o.bad |= gxt_lidar_class(p, S, i, GXT_HAZARD, g, row + p.off_hazards_lidar);
  if (p.off_pillars_lidar >= 0) o.bad |= gxt_lidar_class(p, S, i, GXT_PILLAR, g, row + p.off_pillars_lidar);
  if (p.off_vases_lidar >= 0) o.bad |= gxt_lidar_class(p, S, i, GXT_VASE, g, row + p.off_vases_lidar);
  if (p.off_qpos >= 0) {
    row[p.off_qpos] = d.qx; row[p.off_qpos + 1] = d.qy; row[p.off_qpos + 2] = d.qth;
    o.bad |= gx_notfinite(d.qx) | gx_notfinite(d.qy) | gx_notfinite(d.qth);
  }
  if (p.off_qvel >= 0) {
    row[p.off_qvel] = d.vx; row[p.off_qvel + 1] = d.vy; row[p.off_qvel + 2] = d.vw;
    o.bad |= gx_notfinite(d.vx) | gx_notfinite(d.vy) | gx_notfinite(d.vw);
  }
  // task distances
  {
    float dx = tx - g.px, dy = ty - g.py;
    const float dist_robot_target = gx_sqrt(dx * dx + dy * dy);
    o.dist_task = dist_robot_target;
    o.dist_box = 0.0f;
    if (p.task == GXT_TASK_PUSH) {
      float ex = bx - g.px, ey = by - g.py;
      o.dist_box = gx_sqrt(ex * ex + ey * ey);
      float fx = tx - bx, fy = ty - by;
      o.dist_task = gx_sqrt(fx * fx + fy * fy);       // box -> goal
    }
  }
  // cost, class by class in placement order
  float cost = 0.0f;
#pragma unroll 1
  for (int k = 0; k < p.count[GXT_BUTTON]; ++k) {       // touching a button that is not the goal button
    if (p.task == GXT_TASK_BUTTON && k == (int)goal_button) continue;
    float ox, oy;
    gxt_obj_xy(p, S, i, p.first[GXT_BUTTON] + k, ox, oy);
    float dx = ox - g.px, dy = oy - g.py;
    if (gx_sqrt(dx * dx + dy * dy) < rr + p.size[GXT_BUTTON]) { cost = cost + p.buttons_cost; ++o.contacts; }
  }
#pragma unroll 1
  for (int k = 0; k < p.count[GXT_HAZARD]; ++k) {       // dense, as the JAX Engine (engine.py:804-811)
    float ox, oy;
    gxt_obj_xy(p, S, i, p.first[GXT_HAZARD] + k, ox, oy);
    float dx = ox - g.px, dy = oy - g.py;
    const float dist = gx_sqrt(dx * dx + dy * dy);
    const float term = p.size[GXT_HAZARD] - gx_min(dist, p.size[GXT_HAZARD]);
    if (term > 0.0f) ++o.violations;
    cost = cost + p.hazards_cost_scale * term;
  }
#pragma unroll 1
  for (int k = 0; k < p.count[GXT_VASE]; ++k) {
    float ox, oy;
    gxt_obj_xy(p, S, i, p.first[GXT_VASE] + k, ox, oy);
    float dx = ox - g.px, dy = oy - g.py;
    if (gx_sqrt(dx * dx + dy * dy) < rr + p.size[GXT_VASE]) { cost = cost + p.vases_contact_cost; ++o.contacts; }
  }
#pragma unroll 1
  for (int k = 0; k < p.count[GXT_PILLAR]; ++k) {
    float ox, oy;
    gxt_obj_xy(p, S, i, p.first[GXT_PILLAR] + k, ox, oy);
    float dx = ox - g.px, dy = oy - g.py;
    if (gx_sqrt(dx * dx + dy * dy) < rr + p.size[GXT_PILLAR]) { cost = cost + p.pillars_cost; ++o.contacts; }
  }
#pragma unroll 1
  for (int k = 0; k < p.count[GXT_GREMLIN]; ++k) {
    float ox, oy;
    gxt_gremlin_xy(p, S, i, k, g.time, ox, oy);
    float dx = ox - g.px, dy = oy - g.py;
    const float dist = gx_sqrt(dx * dx + dy * dy);
    if (dist < rr + p.size[GXT_GREMLIN]) { cost = cost + p.gremlins_contact_cost; ++o.contacts; }
    const float gap = dist - p.size[GXT_GREMLIN];
    if (gap < p.gremlins_dist_threshold) cost = cost + p.gremlins_dist_cost * (p.gremlins_dist_threshold - gap);
  }
  o.cost = cost;
  return o;
}

// ---- contacts after the integration (quasi-static) ---------------------------------------------------
// pillars are immovable: the robot is projected out of them and loses its inward normal velocity; vases and the box are
// light: the robot pushes them out of its way along the contact normal (their new positions go back to the state).
__device__ __forceinline__ void gxt_resolve_contacts(const gxt_params& p, float* __restrict__ S, int i, GxDyn& d) {
  const float rr = p.size[GXT_ROBOT];
#pragma unroll 1
  for (int k = 0; k < p.count[GXT_PILLAR]; ++k) {
    float ox, oy;
    gxt_obj_xy(p, S, i, p.first[GXT_PILLAR] + k, ox, oy);
    float dx = d.qx - ox, dy = d.qy - oy;
    const float dist = gx_sqrt(dx * dx + dy * dy);
    const float lim = rr + p.size[GXT_PILLAR];
    if (dist < lim && dist > 0.0f) {
      const float nx = gx_div(dx, dist), ny = gx_div(dy, dist);
      d.qx = ox + nx * lim; d.qy = oy + ny * lim;
      const float vn = d.vx * nx + d.vy * ny;
      if (vn < 0.0f) { d.vx = d.vx - vn * nx; d.vy = d.vy - vn * ny; }
    }
  }
  for (int pass = 0; pass < 2; ++pass) {
    const int cls = pass == 0 ? GXT_BOX : GXT_VASE;
#pragma unroll 1
    for (int k = 0; k < p.count[cls]; ++k) {
      const int o = p.first[cls] + k;
      float ox, oy;
      gxt_obj_xy(p, S, i, o, ox, oy);
      float dx = ox - d.qx, dy = oy - d.qy;
      const float dist = gx_sqrt(dx * dx + dy * dy);
      const float lim = rr + p.size[cls];
      if (dist < lim && dist > 0.0f) {
        const float nx = gx_div(dx, dist), ny = gx_div(dy, dist);
        S[(size_t)(GXT_SLOT_OBJ + 2 * o) * p.cap + i] = d.qx + nx * lim;
        S[(size_t)(GXT_SLOT_OBJ + 2 * o + 1) * p.cap + i] = d.qy + ny * lim;
      }
    }
  }
}

// ---- 8 lanes per env (small and medium batches) -------------------------------------------------------------------
// At BASELINE configs[2]'s size (16 384 envs) one thread per env leaves less than one warp per scheduler and every env is a
// ~4 k-instruction dependent chain (18 objects x lidar ray + cost).  Here lane l of an env's 8 lanes owns objects l, l + 8, ...:
// it reads the object once, casts its ray into the class's lidar (shared-memory row, float max as an integer atomicMax:
// every contribution is >= +0, NaN is canonicalised to the positive pattern, which compares above every finite value),
// adds its cost term, and -- for a vase or the box -- pushes it out of the robot's way.  Cost / counts are combined with a
// fixed butterfly over the 8 lanes (same value on every lane).
#ifndef GXT_LPE
#define GXT_LPE 8
#endif

__device__ __forceinline__ unsigned gxt_grp_mask() { return ((1u << GXT_LPE) - 1u) << (threadIdx.x & 31 & ~(GXT_LPE - 1)); }

__device__ __forceinline__ void gxt_scatter_atomic(const gx_params& p, const GxRay& r, float* bins, int& bad) {
  const int nb = p.nb;
  auto put = [&](int b, float v) {
    bad |= gx_notfinite(v);
    const int bits = (v != v) ? 0x7fc00000 : __float_as_int(v);
    atomicMax(reinterpret_cast<int*>(bins + b), bits);
  };
  if (r.bin < nb) put(r.bin, r.sensor);
  if (p.alias) {
    int bp = r.bin + 1; if (bp >= nb) bp -= nb; if (bp >= nb) bp -= nb;
    int bm = r.bin - 1; if (bm < 0) bm += nb;
    put(bp, r.up);
    put(bm, r.down);
  }
}

__device__ __forceinline__ float gxt_sum8(float x, unsigned mask) {
  if (GXT_LPE > 8) x += __shfl_xor_sync(mask, x, 8);
  x += __shfl_xor_sync(mask, x, 4);
  x += __shfl_xor_sync(mask, x, 2);
  x += __shfl_xor_sync(mask, x, 1);
  return x;
}
__device__ __forceinline__ int gxt_sum8i(int x, unsigned mask) {
  if (GXT_LPE > 8) x += __shfl_xor_sync(mask, x, 8);
  x += __shfl_xor_sync(mask, x, 4);
  x += __shfl_xor_sync(mask, x, 2);
  x += __shfl_xor_sync(mask, x, 1);
  return x;
}

// index of the task's target object: the goal button (button task) or the goal
__device__ __forceinline__ int gxt_target_index(const gxt_params& p, float goal_button) {
  return p.task == GXT_TASK_BUTTON ? p.first[GXT_BUTTON] + min(max((int)goal_button, 0), p.count[GXT_BUTTON] - 1) : p.first[GXT_GOAL];
}

// task distances (every lane, same numbers) from where target and box are in `ob`
__device__ __forceinline__ void gxt_task_dists(const gxt_params& p, const float* ob, int tgt, const GxtGeom& g, GxtObs& o) {
  const float tx = ob[2 * tgt], ty = ob[2 * tgt + 1];
  const float dx = tx - g.px, dy = ty - g.py;
  o.dist_task = gx_sqrt(dx * dx + dy * dy);
  o.dist_box = 0.0f;
  if (p.task == GXT_TASK_PUSH) {
    const float bx = ob[2 * p.first[GXT_BOX]], by = ob[2 * p.first[GXT_BOX] + 1];
    const float ex = bx - g.px, ey = by - g.py;
    o.dist_box = gx_sqrt(ex * ex + ey * ey);
    const float fx = tx - bx, fy = ty - by;
    o.dist_task = gx_sqrt(fx * fx + fy * fy);       // box -> goal
  }
}

// where object j is at the snapshot's time (a gremlin circles its centre at radius gremlins_travel, 1 rad / s)
__device__ __forceinline__ void gxt_obj_pos(const gxt_params& p, const float* ob, int j, int cls, float time, float& ox, float& oy) {
  ox = ob[2 * j]; oy = ob[2 * j + 1];
  if (cls == GXT_GREMLIN) {
    float sn, cs;
    gx_sincos(ob[2 * p.nobj + (j - p.first[GXT_GREMLIN])] + time, &sn, &cs);
    ox = ox + p.gremlins_travel * cs;
    oy = oy + p.gremlins_travel * sn;
  }
}

// cost term of object j (world-frame distance to the snapshot position, as engine.py:804-811 does for hazards)
__device__ __forceinline__ void gxt_obj_cost(const gxt_params& p, int j, int cls, int tgt, float ox, float oy, const GxtGeom& g, float& cost,
                                             int& viol, int& contacts) {
  const float rr = p.size[GXT_ROBOT];
  const float dx = ox - g.px, dy = oy - g.py;
  const float dist = gx_sqrt(dx * dx + dy * dy);
  // Branch-free over the classes (the lanes of an env hold objects of different classes: an if-chain would run class by
  // class).  Every `cost = cost + x` below adds exactly what the chain added, or nothing: `cost` is updated through selects,
  // never by adding a zero, so the bits are the same.
  const float hterm = p.size[GXT_HAZARD] - gx_min(dist, p.size[GXT_HAZARD]);
  const bool is_h = cls == GXT_HAZARD;
  viol += (is_h && hterm > 0.0f) ? 1 : 0;
  cost = is_h ? cost + p.hazards_cost_scale * hterm : cost;
  const bool touch_b = cls == GXT_BUTTON && !(p.task == GXT_TASK_BUTTON && j == tgt) && dist < rr + p.size[GXT_BUTTON];
  const bool touch_v = cls == GXT_VASE && dist < rr + p.size[GXT_VASE];
  const bool touch_p = cls == GXT_PILLAR && dist < rr + p.size[GXT_PILLAR];
  const bool is_g = cls == GXT_GREMLIN;
  const bool touch_g = is_g && dist < rr + p.size[GXT_GREMLIN];
  cost = touch_b ? cost + p.buttons_cost : cost;
  cost = touch_v ? cost + p.vases_contact_cost : cost;
  cost = touch_p ? cost + p.pillars_cost : cost;
  cost = touch_g ? cost + p.gremlins_contact_cost : cost;
  contacts += (touch_b || touch_v || touch_p || touch_g) ? 1 : 0;
  const float gap = dist - p.size[GXT_GREMLIN];
  cost = (is_g && gap < p.gremlins_dist_threshold) ? cost + p.gremlins_dist_cost * (p.gremlins_dist_threshold - gap) : cost;
}

// One observation / cost pass of an env by its 8 lanes from the pose snapshot `g`.  `row`: the env's row in shared memory
// (lidar ranges must be zero on entry, see gxt_zero_lidars).  `push` (may be null): the post-step robot state; vases and
// the box that overlap it are pushed out (written back to S) AFTER their ray has been cast from where they were.
// `ob`: the env's object coordinates staged in shared memory (xy per object, then the gremlin phases; gxt_stage_objects).
__device__ __forceinline__ GxtObs gxt_obs_lanes(const gxt_params& p, float* __restrict__ S, const float* ob, int i, int l, unsigned mask,
                                                const GxtGeom& g, float goal_button, float* row, const GxDyn* push) {
  GxtObs o;
  o.goal_bin = 0;
  int bad = 0, viol = 0, contacts = 0;
  float cost = 0.0f;
  const float rr = p.size[GXT_ROBOT];
  const int tgt = gxt_target_index(p, goal_button);
  int goal_bin = 0;
  // task distances from where target and box are NOW: the loop below may push the box
  gxt_task_dists(p, ob, tgt, g, o);
  __syncwarp(mask);
#pragma unroll 1
  for (int j = l; j < p.nobj; j += GXT_LPE) {
    const int cls = gxt_class_of_obj(p, j);
    float ox, oy;
    gxt_obj_pos(p, ob, j, cls, g.time, ox, oy);
    float ex, ey;
    gx_ego(ox, oy, g.px, g.py, g.r00, g.r10, ex, ey);
    const GxRay r = gx_lidar_ray(p.pt, ex, ey);
    // (selects, not a switch: the lanes of an env hold objects of different classes, and a divergent switch runs case by case)
    int off = -1;
    off = cls == GXT_BOX ? p.off_box_lidar : off;
    off = cls == GXT_BUTTON ? p.off_buttons_lidar : off;
    off = cls == GXT_HAZARD ? p.off_hazards_lidar : off;
    off = cls == GXT_VASE ? p.off_vases_lidar : off;
    off = cls == GXT_PILLAR ? p.off_pillars_lidar : off;
    off = cls == GXT_GREMLIN ? p.off_gremlins_lidar : off;
    if (off >= 0) gxt_scatter_atomic(p.pt, r, row + off, bad);
    if (j == tgt) {
      goal_bin = r.bin;
      if (p.off_goal_compass >= 0) { row[p.off_goal_compass] = ex; row[p.off_goal_compass + 1] = ey; bad |= gx_notfinite(ex) | gx_notfinite(ey); }
      if (p.off_goal_lidar >= 0) gxt_scatter_atomic(p.pt, r, row + p.off_goal_lidar, bad);
    }
    if (cls == GXT_BOX && p.off_box_compass >= 0) { row[p.off_box_compass] = ex; row[p.off_box_compass + 1] = ey; bad |= gx_notfinite(ex) | gx_notfinite(ey); }
    gxt_obj_cost(p, j, cls, tgt, ox, oy, g, cost, viol, contacts);
    // quasi-static push by the post-step robot
    if (push && (cls == GXT_VASE || cls == GXT_BOX)) {
      const float qx = ox - push->qx, qy = oy - push->qy;
      const float dq = gx_sqrt(qx * qx + qy * qy);
      const float lim = rr + p.size[cls];
      if (dq < lim && dq > 0.0f) {
        const float nx = gx_div(qx, dq), ny = gx_div(qy, dq);
        S[(size_t)(GXT_SLOT_OBJ + 2 * j) * p.cap + i] = push->qx + nx * lim;
        S[(size_t)(GXT_SLOT_OBJ + 2 * j + 1) * p.cap + i] = push->qy + ny * lim;
      }
    }
  }
  o.cost = gxt_sum8(cost, mask);
  o.violations = gxt_sum8i(viol, mask);
  o.contacts = gxt_sum8i(contacts, mask);
  o.bad = gxt_sum8i(bad, mask) != 0;
  o.goal_bin = gxt_sum8i(goal_bin, mask);          // exactly one lane owns the target
  return o;
}

// The cost half of gxt_obs_lanes alone (no lidar, no row, no pushes): cost / violations / contacts from the pose snapshot,
// with the same per-lane object order and the same butterfly, so the sums are the same bits; plus the target's lidar bin
// (aux).  For an env that finishes in this step: its observation is the reset observation, only these leave the step pose.
__device__ __forceinline__ void gxt_cost_lanes(const gxt_params& p, const float* ob, int l, unsigned mask, const GxtGeom& g, float goal_button,
                                               GxtObs& o) {
  int viol = 0, contacts = 0, goal_bin = 0;
  float cost = 0.0f;
  const int tgt = gxt_target_index(p, goal_button);
#pragma unroll 1
  for (int j = l; j < p.nobj; j += GXT_LPE) {
    const int cls = gxt_class_of_obj(p, j);
    float ox, oy;
    gxt_obj_pos(p, ob, j, cls, g.time, ox, oy);
    if (j == tgt) {
      float ex, ey;
      gx_ego(ox, oy, g.px, g.py, g.r00, g.r10, ex, ey);
      goal_bin = gx_lidar_ray(p.pt, ex, ey).bin;
    }
    gxt_obj_cost(p, j, cls, tgt, ox, oy, g, cost, viol, contacts);
  }
  o.cost = gxt_sum8(cost, mask);
  o.violations = gxt_sum8i(viol, mask);
  o.contacts = gxt_sum8i(contacts, mask);
  o.goal_bin = gxt_sum8i(goal_bin, mask);
}

// the robot against the pillars (sequential: each projection moves the robot), every lane computes the same numbers
__device__ __forceinline__ void gxt_resolve_pillars(const gxt_params& p, const float* ob, GxDyn& d) {
  const float rr = p.size[GXT_ROBOT];
#pragma unroll 1
  for (int k = 0; k < p.count[GXT_PILLAR]; ++k) {
    const float ox = ob[2 * (p.first[GXT_PILLAR] + k)], oy = ob[2 * (p.first[GXT_PILLAR] + k) + 1];
    float dx = d.qx - ox, dy = d.qy - oy;
    const float dist = gx_sqrt(dx * dx + dy * dy);
    const float lim = rr + p.size[GXT_PILLAR];
    if (dist < lim && dist > 0.0f) {
      const float nx = gx_div(dx, dist), ny = gx_div(dy, dist);
      d.qx = ox + nx * lim; d.qy = oy + ny * lim;
      const float vn = d.vx * nx + d.vy * ny;
      if (vn < 0.0f) { d.vx = d.vx - vn * nx; d.vy = d.vy - vn * ny; }
    }
  }
}

// stage the env's object coordinates and gremlin phases in shared memory: the 8 lanes issue their loads together, every
// later access (lidar loop, pillars, target, box) is a shared-memory read instead of a dependent global load
// Returns 1 when every staged value of this lane is finite and of sane magnitude (then, with a sane pose, every ego vector,
// distance, angle and sensor reading of the observation is finite by construction: the NaN / Inf guard cannot fire on them).
__device__ __forceinline__ int gxt_stage_objects(const gxt_params& p, const float* __restrict__ S, int i, int l, float* ob) {
  const int cnt = 2 * p.nobj + p.ng;
  int ok = 1;
  for (int k = l; k < cnt; k += GXT_LPE) {
    const float v = S[(size_t)(GXT_SLOT_OBJ + k) * p.cap + i];
    ob[k] = v;
    ok &= gx_abs(v) < GX_SANE_BOUND;
  }
  return ok;
}

// zero the lidar ranges of the env's row (its 8 lanes stride over the row)
__device__ __forceinline__ void gxt_zero_lidars(const gxt_params& p, float* row, int l) {
  const int nb = p.pt.nb;
  const int offs[7] = {p.off_box_lidar, p.off_buttons_lidar, p.off_goal_lidar, p.off_gremlins_lidar, p.off_hazards_lidar,
                       p.off_pillars_lidar, p.off_vases_lidar};
#pragma unroll
  for (int c = 0; c < 7; ++c)
    if (offs[c] >= 0) for (int k = l; k < nb; k += GXT_LPE) row[offs[c] + k] = 0.0f;
}

// ---- placement ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ const float* gxt_pick_layout(const gxt_params& p, const float* pool, const float* layout, int i, uint64_t age,
                                                        uint32_t stream) {
  if (layout) return layout + (size_t)i * p.lay_dim;
  const uint64_t g = p.pt.env_id_offset + (uint64_t)i;
  const gx_u32x4 r = gx_philox4x32_10((uint32_t)g, (uint32_t)age, (uint32_t)(age >> 32), stream, p.pt.seed_lo, p.pt.seed_hi);
  return pool + (size_t)gx_randint(r.v[0], (uint32_t)p.pool_size) * p.lay_dim;
}

// layout2qpos + the goal-button draw; leaves lag heading, prev distances, steps and done alone (engine.py:702-731)
__device__ __forceinline__ void gxt_place(const gxt_params& p, float* __restrict__ S, int i, const float* lay, GxDyn& d, float& goal_button,
                                          uint64_t age) {
  for (int k = 0; k < 2 * p.nobj; ++k) S[(size_t)(GXT_SLOT_OBJ + k) * p.cap + i] = lay[k];
  d.qx = lay[2 * p.nobj]; d.qy = lay[2 * p.nobj + 1];
  d.qth = 0.0f; d.vx = 0.0f; d.vy = 0.0f; d.vw = 0.0f;
  for (int g = 0; g < p.ng; ++g) S[(size_t)(GXT_SLOT_OBJ + 2 * p.nobj + g) * p.cap + i] = lay[2 * p.nobj + 2 + g];
  goal_button = 0.0f;
  if (p.task == GXT_TASK_BUTTON) {
    const uint64_t gid = p.pt.env_id_offset + (uint64_t)i;
    const gx_u32x4 r = gx_philox4x32_10((uint32_t)gid, (uint32_t)age, (uint32_t)(age >> 32), GX_STREAM_BUTTON, p.pt.seed_lo, p.pt.seed_hi);
    goal_button = (float)gx_randint(r.v[0], (uint32_t)p.count[GXT_BUTTON]);
  }
  S[(size_t)14 * p.cap + i] = goal_button;
}

#endif  // GX_TASK_CUH_
