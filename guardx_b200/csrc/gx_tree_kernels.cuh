// gx_tree_kernels.cuh -- Engine.reset / step / reset_done for general-tree robots (Doggo).
// One thread per env (gx_tree.cuh); joint state SoA  tq[(which * W + k) * cap + env]  with
// W = nq + 2 nv: which 0 = (qpos | qvel | qacc_warmstart), which 1 = scratch copy for the throw-away
// step of reset_done.  The goal / hazard bodies keep their positions in GxState.goal / .hz and their
// velocities in tw[k * cap + env] (they move when the defaults class gives their slides a stiffness).
// Egocentric transforms use the full 3-D xmat (engine.py:817-844): a tilted torso mixes z in.
#ifndef GX_TREE_KERNELS_CUH_
#define GX_TREE_KERNELS_CUH_

#include "gx_art_launch.h"
#include "gx_device.cuh"
#include "gx_tree.cuh"
#include "gx_tree_launch.h"

// matmul(world_3vec, robot_mat)[:2] with world_3vec = (ox - px, oy - py, 0 - pz)   (engine.py:817-826)
__device__ __forceinline__ void gx_ego3(float ox, float oy, const GxTreeSnap& sn, float& ex, float& ey) {
  const float dx = ox - sn.pos[0], dy = oy - sn.pos[1], dz = 0.0f - sn.pos[2];
  ex = dx * sn.mat[0] + dy * sn.mat[3] + dz * sn.mat[6];
  ey = dx * sn.mat[1] + dy * sn.mat[4] + dz * sn.mat[7];
}

struct GxTreeObs { float dist_goal, cost; int bad, violations; };

__device__ __forceinline__ GxTreeObs gx_tree_obs_row(const gx_params& p, const GxState& s, const GxTreeModel& m, int i,
                                                     float gx_, float gy_, const float* hz /* [2*nh] thread-local */,
                                                     const GxTreeSnap& sn, const float* ctrl, const float* q, const float* v,
                                                     size_t ss, float* row, int32_t* aux_row) {
  GxTreeObs o;
  int bad = 0;
  if (p.off_acc >= 0) { row[p.off_acc] = 0.0f; row[p.off_acc + 1] = 0.0f; }
  if (p.off_vel >= 0) { row[p.off_vel] = 0.0f; row[p.off_vel + 1] = 0.0f; }
  if (p.off_ctrl >= 0) for (int k = 0; k < m.nu; ++k) { row[p.off_ctrl + k] = ctrl[k]; bad |= gx_notfinite(ctrl[k]); }
  if (p.off_qpos >= 0) for (int k = 0; k < m.nq; ++k) { const float x = q[(size_t)k * ss]; row[p.off_qpos + k] = x; bad |= gx_notfinite(x); }
  if (p.off_qvel >= 0) for (int k = 0; k < m.nv; ++k) { const float x = v[(size_t)k * ss]; row[p.off_qvel + k] = x; bad |= gx_notfinite(x); }
  float gex, gey;
  gx_ego3(gx_, gy_, sn, gex, gey);
  if (p.off_compass >= 0) { row[p.off_compass] = gex; row[p.off_compass + 1] = gey; bad |= gx_notfinite(gex) | gx_notfinite(gey); }
  {
    const float dx = gx_ - sn.pos[0], dy = gy_ - sn.pos[1];
    o.dist_goal = gx_sqrt(dx * dx + dy * dy);
  }
  const GxRay gr = gx_lidar_ray(p, gex, gey);
  if (aux_row) aux_row[2] = gr.bin;
  if (p.off_glidar >= 0) {
    for (int k = 0; k < p.nb; ++k) row[p.off_glidar + k] = 0.0f;
    bad |= gx_lidar_scatter(p, gr, row + p.off_glidar);
  }
  if (p.off_hlidar >= 0) for (int k = 0; k < p.nb; ++k) row[p.off_hlidar + k] = 0.0f;
  float cost = 0.0f;
  int viol = 0;
  for (int h = 0; h < p.nh; ++h) {
    const float hx = hz[2 * h], hy = hz[2 * h + 1];
    float ex, ey;
    gx_ego3(hx, hy, sn, ex, ey);
    const GxRay r = gx_lidar_ray(p, ex, ey);
    if (aux_row) aux_row[3 + h] = r.bin;
    if (p.off_hlidar >= 0) bad |= gx_lidar_scatter(p, r, row + p.off_hlidar);
    const float dx = hx - sn.pos[0], dy = hy - sn.pos[1];
    const float dist = gx_sqrt(dx * dx + dy * dy);
    const float term = p.hazards_size - gx_min(dist, p.hazards_size);
    if (term > 0.0f) ++viol;
    cost = cost + term;
  }
  o.cost = cost; o.violations = viol; o.bad = bad;
  return o;
}

// ego_vel_acc (engine.py:902-929) with the full 3-D xpos / xmat of a general-tree root: finite differences of
// the root position over Engine._data / _last_data / _last_last_data with the done masking, rotated into the
// robot frame (world_vec @ xmat, first two components).  x / y history and the flags live in s.hx / s.hf as for
// the other robots (gx_vel_acc), the z history in zh[0] (last) and zh[cap] (last-last).
__device__ __forceinline__ int gx_vel_acc3(const gx_params& p, const GxState& s, float* zh, size_t cap, int i,
                                           const GxTreeSnap& sn, float last_done, float* row) {
  const float4 h = s.hx[i];
  const uint32_t f = s.hf[i];
  const float last_last_done = (f & 1u) ? 1.0f : 0.0f;
  const int hist = (int)((f >> 1) & 3u);
  const float cur[3] = {sn.pos[0], sn.pos[1], sn.pos[2]};
  const float l[3] = {h.x, h.y, zh[0]};
  const float ll[3] = {hist >= 2 ? h.z : l[0], hist >= 2 ? h.w : l[1], hist >= 2 ? zh[cap] : l[2]};
  float vw[3], aw[3];
  for (int k = 0; k < 3; ++k) {
    float pl = cur[k], pll = cur[k];
    if (hist >= 1) {
      if (!(last_done > 0.0f)) pl = l[k];
      if (hist >= 2) pll = (last_last_done + last_done > 0.0f) ? pl : ll[k];
    }
    vw[k] = gx_div(cur[k] - pl, p.ctrl_dt);
    const float lv = gx_div(pl - pll, p.ctrl_dt);
    aw[k] = gx_div(vw[k] - lv, p.ctrl_dt);
  }
  const float* R = sn.mat;
  const float vx = vw[0] * R[0] + vw[1] * R[3] + vw[2] * R[6], vy = vw[0] * R[1] + vw[1] * R[4] + vw[2] * R[7];
  const float ax = aw[0] * R[0] + aw[1] * R[3] + aw[2] * R[6], ay = aw[0] * R[1] + aw[1] * R[4] + aw[2] * R[7];
  int bad = 0;
  if (p.off_acc >= 0) { row[p.off_acc] = ax; row[p.off_acc + 1] = ay; bad |= gx_notfinite(ax) | gx_notfinite(ay); }
  if (p.off_vel >= 0) { row[p.off_vel] = vx; row[p.off_vel + 1] = vy; bad |= gx_notfinite(vx) | gx_notfinite(vy); }
  s.hx[i] = make_float4(cur[0], cur[1], l[0], l[1]);
  zh[cap] = l[2]; zh[0] = cur[2];
  s.hf[i] = (last_done > 0.0f ? 1u : 0u) | ((uint32_t)(hist < 2 ? hist + 1 : 2) << 1);
  return bad;
}

// one semi-implicit Euler step of a goal / hazard slide dof (mass + armature M, damping d, stiffness k,
// spring reference 0): the same arithmetic mjx.step applies to them inside the full model
__device__ __forceinline__ void gx_world_dof(const GxTreeModel& m, float M, float& q, float& v) {
  const float f = -m.world_stiffness * q - m.world_damping * v;
  const float a = gx_div(f, gx_fma(m.dt, m.world_damping, M));
  v = gx_fma(m.dt, a, v);
  q = gx_fma(m.dt, v, q);
}

template <int MODE, bool AUTORESET>
__global__ void __launch_bounds__(64) gx_tree_kernel(const __grid_constant__ gx_params p, const GxState s,
                                                     const GxTreeDev a, const __grid_constant__ GxTreeModel mdl,
                                                     const float* __restrict__ pool,
                                                     const float* __restrict__ layout,
                                                     const float* __restrict__ action, const float* obs_in,
                                                     float* obs, float* __restrict__ act_out,
                                                     float* __restrict__ reward, float* __restrict__ done,
                                                     float* __restrict__ cost, float* __restrict__ qacc_out,
                                                     int32_t* __restrict__ aux, int stats_on) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int wbase = i - lane;
  if (wbase >= p.n) return;
  const int rows = min(32, p.n - wbase);
  const bool live = i < p.n;
  // The model travels as a kernel parameter: every read of it is a uniform constant-bank load instead of a global
  // load that competes with the scratch for the small L1 (the model is read ~2 k times per env-step).
  const GxTreeModel& m = mdl;
  const size_t cap = (size_t)s.cap;
  const int W = a.nq + 2 * a.nv;
  // The observation row is written in place (no staging tile: one 340-byte row per env and per multi-ms step).
  float* row = obs + (size_t)i * p.obs_dim;
  // matrix scratch (Hessian / Cholesky workspace, solve operand, qfrc_constraint, packed mass matrix): shared
  // memory, element k of this thread at sm[k * blockDim.x] -- every load of the factorisation is on a dependent
  // chain, and a shared-memory round trip is a fifth of an L2 one.  Doggo: 418 floats per env, 53.5 KB per 32-env
  // block, four blocks per SM.  (Leaving M in the global scratch -- 7 blocks per SM -- measured slower once the
  // constraint rows went through registers: gx_tree_step takes M's pointer / stride separately for that experiment.)
  float* sm = gx_smem + threadIdx.x;
  const size_t ms = blockDim.x;
  float* mm = sm + (size_t)GXT_MOFF(a.nv) * ms;
  const size_t mms = ms;

  if (MODE == GX_ART_RESET_DONE) {
    if (obs_in != obs) {
      const int total = rows * p.obs_dim;
      const float* src = obs_in + (size_t)wbase * p.obs_dim;
      float* dst = obs + (size_t)wbase * p.obs_dim;
      for (int k = lane; k < total; k += 32) dst[k] = src[k];
      __syncwarp();
    }
  }
  if (live) {
    float* q = a.tq + i;                       // element k at q[k * cap]
    float* v = q + (size_t)a.nq * cap;
    float* w = v + (size_t)a.nv * cap;
    float* q2 = a.tq + (size_t)W * cap + i;    // scratch copy (throw-away step)
    float* v2 = q2 + (size_t)a.nq * cap;
    float* w2 = v2 + (size_t)a.nv * cap;
    float* wv = a.tw + i;
    float* zh = a.tw + (size_t)m.nworld * cap + i;   // z history of ego_vel_acc
    // global scratch: one chunk of (floats per env) x 32 per block, element k of lane l at chunk[k * 32 + l] -- coalesced,
    // and the stride is a compile-time constant (loads at different constant offsets can pass earlier stores)
    float* chunk = a.scratch + (size_t)blockIdx.x * a.scratch_floats * GXT_DEVICE_MS;
    float* sc = chunk + lane;
    float* jrows = chunk + 4 * lane;            // this env's slot in the float4-grouped constraint rows (region GXT_J)
    const float4 c2 = s.d2[i];
    float prev_dist = c2.x;
    float was_done = (__float_as_uint(c2.y) >> 31) ? 1.0f : 0.0f;
    float steps = fabsf(c2.y);
    uint64_t age = (uint64_t)__float_as_uint(c2.z) | ((uint64_t)__float_as_uint(c2.w) << 32);
    float2 goal = s.goal[i];
    float hz[2 * GX_MAX_HAZARDS];
    for (int h = 0; h < p.nh; ++h) gx_art_hazard(s, nullptr, i, h, hz[2 * h], hz[2 * h + 1]);
    float ctrl[GX_TREE_MAXU];
    for (int k = 0; k < m.nu; ++k) ctrl[k] = 0.0f;
    GxTreeSnap sn;
    float o_reward = 0.0f, o_done = 0.0f, o_cost = 0.0f;

    // layout2qpos (engine.py:623-639): goal / hazards to their slides (velocity 0), robot to (x, y) of its
    // planar joints or to (x, y, z_height, 1, 0, 0, 0) of its free joint, every other qpos 0, qvel 0
    auto place = [&](const float* lay, float* qd, float* vd) {
      goal = make_float2(lay[0], lay[1]);
      s.goal[i] = goal;
      for (int k = 0; k < s.nh2; ++k) {
        float4 h;
        h.x = lay[4 + 4 * k]; h.y = lay[5 + 4 * k];
        h.z = (2 * k + 1 < p.nh) ? lay[6 + 4 * k] : 0.0f;
        h.w = (2 * k + 1 < p.nh) ? lay[7 + 4 * k] : 0.0f;
        s.hz[(size_t)k * cap + i] = h;
      }
      for (int h = 0; h < p.nh; ++h) { hz[2 * h] = lay[4 + 2 * h]; hz[2 * h + 1] = lay[5 + 2 * h]; }
      for (int k = 0; k < m.nworld; ++k) wv[(size_t)k * cap] = 0.0f;
      for (int k = 0; k < m.nq; ++k) qd[(size_t)k * cap] = 0.0f;
      for (int k = 0; k < m.nv; ++k) vd[(size_t)k * cap] = 0.0f;
      qd[(size_t)m.root_xadr * cap] = lay[2]; qd[(size_t)m.root_yadr * cap] = lay[3];
      if (m.root_free) { qd[2 * cap] = m.body[0].pos[2]; qd[3 * cap] = 1.0f; }
    };
    // the goal / hazard slides advance with every mjx.step; the observation of that step sees their
    // PRE-step positions (xpos comes from the forward pass), so integrate after the row is built
    auto advance_world = [&]() {
      if (m.world_stiffness == 0.0f) return;
      float vx = wv[0], vy = wv[cap];
      gx_world_dof(m, m.world_mass[0], goal.x, vx); gx_world_dof(m, m.world_mass[0], goal.y, vy);
      wv[0] = vx; wv[cap] = vy;
      for (int h = 0; h < p.nh; ++h) {
        float hvx = wv[(size_t)(2 + 2 * h) * cap], hvy = wv[(size_t)(3 + 2 * h) * cap];
        gx_world_dof(m, m.world_mass[1], hz[2 * h], hvx); gx_world_dof(m, m.world_mass[1], hz[2 * h + 1], hvy);
        wv[(size_t)(2 + 2 * h) * cap] = hvx; wv[(size_t)(3 + 2 * h) * cap] = hvy;
      }
    };
    auto store_world = [&]() {
      s.goal[i] = goal;
      for (int k = 0; k < s.nh2; ++k) {
        float4 h;
        h.x = hz[4 * k]; h.y = hz[4 * k + 1];
        h.z = (2 * k + 1 < p.nh) ? hz[4 * k + 2] : 0.0f;
        h.w = (2 * k + 1 < p.nh) ? hz[4 * k + 3] : 0.0f;
        s.hz[(size_t)k * cap + i] = h;
      }
    };

    if (MODE == GX_ART_RESET) {
      const float* lay = gx_pick_layout(p, pool, layout, i, age, GX_STREAM_PICK ^ 1u);
      place(lay, q, v);
      for (int k = 0; k < m.nv; ++k) w[(size_t)k * cap] = 0.0f;     // mjx_reset starts from the pristine mjx_data
      gx_tree_step(m, q, v, w, cap, ctrl, sc, GXT_DEVICE_MS, jrows, sm, ms, mm, mms, true, q, v, w, sn, nullptr);
      GxTreeObs o = gx_tree_obs_row(p, s, m, i, goal.x, goal.y, hz, sn, ctrl, q, v, cap, row, nullptr);
      prev_dist = o.dist_goal;
      steps = 0.0f;
      if (s.ep) s.ep[i] = make_float4(0.f, 0.f, 0.f, s.ep[i].w);
      if ((p.off_vel >= 0) | (p.off_acc >= 0)) {   // Engine._data.xpos <- reset pose; _last_data stays
        float4 h = s.hx[i];
        h.x = sn.pos[0]; h.y = sn.pos[1];
        s.hx[i] = h;
        zh[0] = sn.pos[2];
      }
    } else if (MODE == GX_ART_RESET_DONE) {
      if (was_done > 0.0f) {
        const float* lay = gx_pick_layout(p, pool, layout, i, age, GX_STREAM_PICK);
        place(lay, q, v);
        // throw-away step(s) on a scratch copy: the state keeps the un-stepped pose and the stale warm start
        const float* qs = q; const float* vs = v; const float* ws = w;
        for (int k = 0; k < p.substeps; ++k) {
          gx_tree_step(m, qs, vs, ws, cap, ctrl, sc, GXT_DEVICE_MS, jrows, sm, ms, mm, mms, false, q2, v2, w2, sn, nullptr);
          qs = q2; vs = v2; ws = w2;
        }
        gx_tree_obs_row(p, s, m, i, goal.x, goal.y, hz, sn, ctrl, q2, v2, cap, row, nullptr);
      }
    } else {
      const uint64_t g = p.env_id_offset + (uint64_t)i;
      for (int k = 0; k < m.nu; ++k) {
        ctrl[k] = action ? action[(size_t)i * m.nu + k] : gx_art_synth_action(p, g, age, k);
        if (act_out) act_out[(size_t)i * m.nu + k] = ctrl[k];
      }
      const float last_dist = prev_dist, last_done = was_done;
      float gpre_x = goal.x, gpre_y = goal.y, hpre[2 * GX_MAX_HAZARDS];
      for (int k = 0; k < p.substeps; ++k) {
        gpre_x = goal.x; gpre_y = goal.y;
        for (int h = 0; h < 2 * p.nh; ++h) hpre[h] = hz[h];
        gx_tree_step(m, q, v, w, cap, ctrl, sc, GXT_DEVICE_MS, jrows, sm, ms, mm, mms, false, q, v, w, sn, nullptr);
        advance_world();
      }
      if (qacc_out) for (int k = 0; k < m.nv; ++k) qacc_out[(size_t)i * m.nv + k] = w[(size_t)k * cap];
      int32_t* aux_row = aux ? aux + (size_t)i * p.aux_dim : nullptr;
      GxTreeObs o = gx_tree_obs_row(p, s, m, i, gpre_x, gpre_y, hpre, sn, ctrl, q, v, cap, row, aux_row);
      if ((p.off_vel >= 0) | (p.off_acc >= 0)) o.bad |= gx_vel_acc3(p, s, zh, cap, i, sn, last_done, row);
      const float dist = o.dist_goal;
      const float last = last_done > 0.0f ? dist : last_dist;
      const float d_dist = last - dist;
      o_reward = d_dist * p.reward_distance;
      o_done = dist < p.goal_size ? 1.0f : 0.0f;
      int flags = o_done > 0.0f ? GX_AUX_GOAL : 0;
      if (gx_abs(d_dist) > 1.0f) { o_done = 1.0f; o_reward = 0.0f; flags |= GX_AUX_JUMP; }
      if (o.bad) { o_reward = 0.0f; o_done = 1.0f; flags |= GX_AUX_GUARD; }
      if (steps > p.num_steps) { o_done = 1.0f; flags |= GX_AUX_TIMEOUT; }
      if (aux_row) { aux_row[0] = flags; aux_row[1] = o.violations; }
      steps = o_done > 0.0f ? 0.0f : steps + 1.0f;
      was_done = o_done;
      prev_dist = dist;
      age += 1;
      o_cost = o.cost;
      if (stats_on) gx_stats_step(s, i, o_reward, o_cost, o_done);
      store_world();
      if (AUTORESET && o_done > 0.0f) {
        const float* lay = gx_pick_layout(p, pool, nullptr, i, age, GX_STREAM_PICK);
        place(lay, q, v);
        for (int k = 0; k < m.nu; ++k) ctrl[k] = 0.0f;
        const float* qs = q; const float* vs = v; const float* ws = w;
        for (int k = 0; k < p.substeps; ++k) {
          gx_tree_step(m, qs, vs, ws, cap, ctrl, sc, GXT_DEVICE_MS, jrows, sm, ms, mm, mms, false, q2, v2, w2, sn, nullptr);
          qs = q2; vs = v2; ws = w2;
        }
        gx_tree_obs_row(p, s, m, i, goal.x, goal.y, hz, sn, ctrl, q2, v2, cap, row, nullptr);
      }
      reward[i] = o_reward; done[i] = o_done; cost[i] = o_cost;
    }
    s.d2[i] = make_float4(prev_dist, was_done > 0.0f ? -steps : steps, __uint_as_float((uint32_t)age),
                          __uint_as_float((uint32_t)(age >> 32)));
  }
}

#endif  // GX_TREE_KERNELS_CUH_
