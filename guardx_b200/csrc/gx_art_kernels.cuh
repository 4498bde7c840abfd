// gx_art_kernels.cuh -- Engine.reset / step / reset_done for articulated robots
// (included by guardx.cu after the Point kernels; shares GxState's env-level arrays,
// the layout pool, the Philox streams and the lidar / reward / cost device functions).
//
// Thread mapping: LANES consecutive threads per env (lane l owns leg l, gx_art.cuh),
// 32/LANES envs per warp; the warp's observation rows are assembled in a shared-memory
// tile and leave as coalesced 128-bit stores.  The joint state lives in one SoA array
//   aq[(which * nv + dof) * cap + env],  which = 0 qpos, 1 qvel, 2 qacc_warmstart.
#ifndef GX_ART_KERNELS_CUH_
#define GX_ART_KERNELS_CUH_

#include "gx_art.cuh"
#include "gx_art_launch.h"
#include "gx_device.cuh"


template <int LANES>
__device__ __forceinline__ unsigned gx_art_mask() {
  if (LANES == 32) return 0xffffffffu;
  const unsigned lw = threadIdx.x & 31u;
  return ((1u << LANES) - 1u) << (lw & ~(unsigned)(LANES - 1));
}

struct GxArtObs { float dist_goal, cost; int bad, violations; };

// One ray into the env's shared lidar row by atomic max on the float's bit pattern (readings are >= +0, for which the integer
// order is the float order; NaN is written as the canonical quiet NaN, which is larger than every finite pattern, and
// flagged).  Same index semantics as gx_lidar_scatter (gather clamps, out-of-bounds scatter is dropped).
__device__ __forceinline__ int gx_art_scatter_atomic(const gx_params& p, const GxRay& r, float* bins) {
  const int nb = p.nb;
  int bad = 0;
  auto put = [&](int b, float v) {
    bad |= gx_notfinite(v);
    atomicMax(reinterpret_cast<int*>(bins + b), (v != v) ? 0x7fc00000 : __float_as_int(v));
  };
  if (r.bin < nb) put(r.bin, r.sensor);
  if (p.alias) {
    int bp = r.bin + 1; if (bp >= nb) bp -= nb; if (bp >= nb) bp -= nb;
    int bm = r.bin - 1; if (bm < 0) bm += nb;
    put(bp, r.up);
    put(bm, r.down);
  }
  return bad;
}

// Observation row of one env, written by its LANES lanes into `row` (shared memory):
// geometry from the pose snapshot (engine.py:754-758, 817-900), qpos / qvel = (q, v) given
// (post-step values on a step, engine.py:759-762), ctrl = the unclipped action.
template <int LANES, int NLEG, int K, int NB>
__device__ __forceinline__ GxArtObs gx_art_obs_row(const gx_params& p, const GxState& s, const GxArtModel& m, int i,
                                                   int lane, unsigned mask, const float* lay, float gx_, float gy_,
                                                   const GxArtSnap& sn, const float* cr, const float (*cl)[K],
                                                   const float* qr, const float* vr, const float (*ql)[K],
                                                   const float (*vl)[K], float* row, int32_t* aux_row) {
  constexpr int LPL = NLEG / LANES;
  GxArtObs o;
  int bad = 0;
  const int nb = p.nb;
  // zero both lidars cooperatively
  if (p.off_glidar >= 0) for (int k = lane; k < nb; k += LANES) row[p.off_glidar + k] = 0.0f;
  if (p.off_hlidar >= 0) for (int k = lane; k < nb; k += LANES) row[p.off_hlidar + k] = 0.0f;
  // root entries: lane 0; leg entries: owning lane
  if (lane == 0) {
    // vel / acc default to 0 = obs(data, None, None, None, None) (engine.py:656, 726); a step overwrites them
    if (p.off_acc >= 0) { row[p.off_acc] = 0.0f; row[p.off_acc + 1] = 0.0f; }
    if (p.off_vel >= 0) { row[p.off_vel] = 0.0f; row[p.off_vel + 1] = 0.0f; }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int a = m.root.dof[r].act;
      if (p.off_ctrl >= 0 && a >= 0) { row[p.off_ctrl + a] = cr[r]; bad |= gx_notfinite(cr[r]); }
      if (p.off_qpos >= 0) { row[p.off_qpos + m.root_qadr + r] = qr[r]; bad |= gx_notfinite(qr[r]); }
      if (p.off_qvel >= 0) { row[p.off_qvel + m.root_qadr + r] = vr[r]; bad |= gx_notfinite(vr[r]); }
    }
  }
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
    const int leg = j * LANES + lane;
    const GxArtSeg& sg = m.leg[leg];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k < sg.nk) {
        const int a = sg.dof[k].act, idx = m.leg_qadr[leg] + k;
        if (p.off_ctrl >= 0 && a >= 0) { row[p.off_ctrl + a] = cl[j][k]; bad |= gx_notfinite(cl[j][k]); }
        if (p.off_qpos >= 0) { row[p.off_qpos + idx] = ql[j][k]; bad |= gx_notfinite(ql[j][k]); }
        if (p.off_qvel >= 0) { row[p.off_qvel + idx] = vl[j][k]; bad |= gx_notfinite(vl[j][k]); }
      }
    }
  }
  float gex, gey;
  gx_ego(gx_, gy_, sn.px, sn.py, sn.r00, sn.r10, gex, gey);
  {
    const float dx = gx_ - sn.px, dy = gy_ - sn.py;
    o.dist_goal = gx_sqrt(dx * dx + dy * dy);
  }
  __syncwarp(mask);
  if (lane == 0) {
    if (p.off_compass >= 0) { row[p.off_compass] = gex; row[p.off_compass + 1] = gey; bad |= gx_notfinite(gex) | gx_notfinite(gey); }
    const GxRay gr = gx_lidar_ray(p, gex, gey);
    if (aux_row) aux_row[2] = gr.bin;
    if (p.off_glidar >= 0) bad |= gx_lidar_scatter(p, gr, row + p.off_glidar);
  }
  // hazards: lane l takes hazards l, l + LANES, ...; all lanes cast their rays at the same time and scatter into the shared
  // row with atomic max (the updates are max, so their order cannot change the result).  The scatters used to be
  // serialised lane by lane, i.e. the lidar ran with one lane in LANES active: 14 % of the ant's step.
  float cost = 0.0f, viol = 0.0f;
#pragma unroll 1
  for (int h = lane; h < p.nh; h += LANES) {
    float hx, hy, ex, ey;
    gx_art_hazard(s, lay, i, h, hx, hy);
    gx_ego(hx, hy, sn.px, sn.py, sn.r00, sn.r10, ex, ey);
    if (p.off_hlidar >= 0 || aux_row) {
      const GxRay r = gx_lidar_ray(p, ex, ey);
      if (aux_row) aux_row[3 + h] = r.bin;
      if (p.off_hlidar >= 0) bad |= gx_art_scatter_atomic(p, r, row + p.off_hlidar);
    }
    const float dx = hx - sn.px, dy = hy - sn.py;
    const float dist = gx_sqrt(dx * dx + dy * dy);
    const float term = p.hazards_size - gx_min(dist, p.hazards_size);
    if (term > 0.0f) viol += 1.0f;
    cost = cost + term;
  }
  __syncwarp(mask);
  o.cost = gx_art_sum<LANES>(cost, mask);
  o.violations = (int)gx_art_sum<LANES>(viol, mask);
  o.bad = gx_art_or<LANES>(bad, mask);
  return o;
}

// MINB = resident 128-thread blocks per SM the register allocation aims for: 2 (no spills,
// ~250 registers) is fastest while the batch leaves SMs idle anyway; 4 (128 registers, some
// spills to L1) wins by ~30 % once the grid covers the machine several times (measured:
// ant, 65 536 envs, 512 -> 393 us per step).
template <int LANES, int NLEG, int K, int NB, int MODE, bool AUTORESET, int MINB>
__global__ void __launch_bounds__(128, MINB) gx_art_kernel(const __grid_constant__ gx_params p, const GxState s,
                                                     const GxArtDev a, const float* __restrict__ pool,
                                                     const float* __restrict__ layout,
                                                     const float* __restrict__ action, const float* obs_in,
                                                     float* obs, float* __restrict__ act_out,
                                                     float* __restrict__ reward, float* __restrict__ done,
                                                     float* __restrict__ cost, float* __restrict__ qacc_out, int32_t* __restrict__ aux,
                                                     int stats_on) {
  constexpr int LPL = NLEG / LANES;
  constexpr int EPW = 32 / LANES;
  const int lw = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lane = lw % LANES, ew = lw / LANES;
  const int wbase = (int)(((size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / LANES);
  if (wbase >= p.n) return;
  const int rows = min(EPW, p.n - wbase);
  const int env = wbase + ew;
  const bool live = env < p.n;
  const int i = live ? env : p.n - 1;
  const unsigned mask = gx_art_mask<LANES>();
  const GxArtModel& m = *a.m;
  const int nv = a.nv, nu = a.nu;
  const size_t cap = (size_t)s.cap;
  float* tile = gx_smem + (size_t)warp * EPW * p.obs_dim;
  float* row = tile + (size_t)ew * p.obs_dim;

  // ---- env-level state
  const float4 c2 = s.d2[i];
  float prev_dist = c2.x;
  float was_done = (__float_as_uint(c2.y) >> 31) ? 1.0f : 0.0f;
  float steps = fabsf(c2.y);
  uint64_t age = (uint64_t)__float_as_uint(c2.z) | ((uint64_t)__float_as_uint(c2.w) << 32);

  if (MODE == GX_ART_RESET_DONE) {
    // rows of envs that are not done pass through (Engine._obs); done envs overwrite theirs below
    if (obs_in != obs) {
      const int total = rows * p.obs_dim;
      const float* src = obs_in + (size_t)wbase * p.obs_dim;
      float* dst = obs + (size_t)wbase * p.obs_dim;
      for (int k = lw; k < total; k += 32) dst[k] = src[k];
      __syncwarp();
    }
    if (!(was_done > 0.0f) || !live) return;
  }

  // ---- joint state
  GxArtLane<LANES, NLEG, K, NB> st;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const size_t d = (size_t)(m.root_qadr + r);
    st.qr[r] = a.aq[d * cap + i];
    st.vr[r] = a.aq[(nv + d) * cap + i];
    st.wr[r] = a.aq[(2 * nv + d) * cap + i];
  }
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
    const int leg = j * LANES + lane;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const bool on = k < m.leg[leg].nk;
      const size_t d = (size_t)(m.leg_qadr[leg] + k);
      st.ql[j][k] = on ? a.aq[d * cap + i] : 0.0f;
      st.vl[j][k] = on ? a.aq[(nv + d) * cap + i] : 0.0f;
      st.wl[j][k] = on ? a.aq[(2 * nv + d) * cap + i] : 0.0f;
    }
  }
  float2 goal = s.goal[i];
  float cr[3] = {0.f, 0.f, 0.f}, cl[LPL][K];
#pragma unroll
  for (int j = 0; j < LPL; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) cl[j][k] = 0.0f;
  float qro[3], vro[3], qlo[LPL][K], vlo[LPL][K];
  GxArtSnap sn;
  float o_reward = 0.0f, o_done = 0.0f, o_cost = 0.0f;

  // place the env on a layout row: goal / hazards into the state arrays (lane 0), robot
  // root x, y from the layout, every other qpos 0, qvel 0 (layout2qpos, engine.py:623-639)
  auto place = [&](const float* lay) {
    goal = make_float2(lay[0], lay[1]);
    if (lane == 0 && live) {
      s.goal[i] = goal;
      for (int k = 0; k < s.nh2; ++k) {
        float4 h;
        h.x = lay[4 + 4 * k]; h.y = lay[5 + 4 * k];
        h.z = (2 * k + 1 < p.nh) ? lay[6 + 4 * k] : 0.0f;
        h.w = (2 * k + 1 < p.nh) ? lay[7 + 4 * k] : 0.0f;
        s.hz[(size_t)k * cap + i] = h;
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) { st.qr[r] = (r == m.root_x) ? lay[2] : ((r == m.root_y) ? lay[3] : 0.0f); st.vr[r] = 0.0f; }
#pragma unroll
    for (int j = 0; j < LPL; ++j)
#pragma unroll
      for (int k = 0; k < K; ++k) { st.ql[j][k] = 0.0f; st.vl[j][k] = 0.0f; }
  };

  // joint state -> HBM.  The state is stepped IN PLACE and stored as soon as it is final (after the step proper; after the
  // re-placement for a throw-away step, whose warm start stays the stale one), so no second copy of it is live across
  // the solver -- 15 registers less for the ant in every build.
  auto store_state = [&](bool with_w) {
    if (!live) return;
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const size_t d = (size_t)(m.root_qadr + r);
        a.aq[d * cap + i] = st.qr[r];
        a.aq[(nv + d) * cap + i] = st.vr[r];
        if (with_w) a.aq[(2 * nv + d) * cap + i] = st.wr[r];
      }
    }
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
      const int leg = j * LANES + lane;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (k < m.leg[leg].nk) {
          const size_t d = (size_t)(m.leg_qadr[leg] + k);
          a.aq[d * cap + i] = st.ql[j][k];
          a.aq[(nv + d) * cap + i] = st.vl[j][k];
          if (with_w) a.aq[(2 * nv + d) * cap + i] = st.wl[j][k];
        }
      }
    }
  };

  if (MODE == GX_ART_RESET) {
    // Engine.reset (engine.py:454-467): mjx_reset starts from the pristine mjx_data
    // (qacc_warmstart 0), runs mjx.forward, observes; steps <- 0; done/key stay stale.
    const float* lay = gx_pick_layout(p, pool, layout, i, age, GX_STREAM_PICK ^ 1u);
    place(lay);
#pragma unroll
    for (int r = 0; r < 3; ++r) st.wr[r] = 0.0f;
#pragma unroll
    for (int j = 0; j < LPL; ++j)
#pragma unroll
      for (int k = 0; k < K; ++k) st.wl[j][k] = 0.0f;
    gx_art_step<LANES, NLEG, K, NB>(m, lane, mask, st, nullptr, nullptr, false, true, false, sn, qro, vro, qlo, vlo, nullptr);
    store_state(true);
    GxArtObs o = gx_art_obs_row<LANES, NLEG, K, NB>(p, s, m, i, lane, mask, lay, goal.x, goal.y, sn, cr, cl, qro, vro, qlo, vlo, row, nullptr);
    prev_dist = o.dist_goal;
    steps = 0.0f;
    if (lane == 0 && live && s.ep) s.ep[i] = make_float4(0.f, 0.f, 0.f, s.ep[i].w);
    if (lane == 0 && live && ((p.off_vel >= 0) | (p.off_acc >= 0))) {   // Engine._data.xpos <- reset pose; _last_data stays
      float4 h = s.hx[i];
      h.x = sn.px; h.y = sn.py;
      s.hx[i] = h;
    }
  } else if (MODE == GX_ART_RESET_DONE) {
    // Engine.reset_done for a done env (engine.py:702-731): re-place, ctrl 0, one throw-away
    // mjx.step for the observation; the state keeps the UN-stepped pose, the stale warm start,
    // prev_dist, steps and the done flag.
    const float* lay = gx_pick_layout(p, pool, layout, i, age, GX_STREAM_PICK);
    place(lay);
    store_state(false);   // the un-stepped pose is the state; the step below is for the observation only
    for (int k = 0; k < p.substeps; ++k)
      gx_art_step<LANES, NLEG, K, NB>(m, lane, mask, st, nullptr, nullptr, true, false, false, sn, qro, vro, qlo, vlo, nullptr);
    gx_art_obs_row<LANES, NLEG, K, NB>(p, s, m, i, lane, mask, lay, goal.x, goal.y, sn, cr, cl, qro, vro, qlo, vlo, row, nullptr);
  } else {
    // Engine.step (engine.py:469-495) + mjx_step (:659-700)
    const uint64_t g = p.env_id_offset + (uint64_t)i;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int ai = m.root.dof[r].act;
      if (ai >= 0) {
        cr[r] = action ? action[(size_t)i * nu + ai] : gx_art_synth_action(p, g, age, ai);
        if (act_out && live && lane == 0) act_out[(size_t)i * nu + ai] = cr[r];   // actuated root dof (none of the shipped xmls has one)
      }
    }
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
      const GxArtSeg& sg = m.leg[j * LANES + lane];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const int ai = k < sg.nk ? sg.dof[k].act : -1;
        if (ai >= 0) {
          cl[j][k] = action ? action[(size_t)i * nu + ai] : gx_art_synth_action(p, g, age, ai);
          if (act_out && live) act_out[(size_t)i * nu + ai] = cl[j][k];
        }
      }
    }
    const float last_dist = prev_dist, last_done = was_done;
    // pass 0: the step.  pass 1 (AUTORESET and done): Engine.reset_done's throw-away step on the
    // re-placed env.  One loop body = one copy of the dynamics in the instruction stream.
    const float* lay = nullptr;
#pragma unroll 1
    for (int pass = 0; pass < (AUTORESET ? 2 : 1); ++pass) {
      if (pass == 1) {
        if (!(o_done > 0.0f)) break;
        lay = gx_pick_layout(p, pool, nullptr, i, age, GX_STREAM_PICK);
        place(lay);
#pragma unroll
        for (int r = 0; r < 3; ++r) cr[r] = 0.0f;
#pragma unroll
        for (int j = 0; j < LPL; ++j)
#pragma unroll
          for (int k = 0; k < K; ++k) cl[j][k] = 0.0f;
      }
      if (pass == 1) store_state(false);   // re-placed, un-stepped pose; the warm start stays pass 0's (engine.py:719-731)
      for (int k = 0; k < p.substeps; ++k)
        gx_art_step<LANES, NLEG, K, NB>(m, lane, mask, st, cr, cl, true, false, false, sn, qro, vro, qlo, vlo, nullptr);
      if (pass == 0) {
        store_state(true);
        if (qacc_out && live) {   // info['obs']['qacc'] (engine.py:763-764): the solver's qacc of the last forward
          if (lane == 0)
#pragma unroll
            for (int r = 0; r < 3; ++r) qacc_out[(size_t)i * nv + m.root_qadr + r] = st.wr[r];
#pragma unroll
          for (int j = 0; j < LPL; ++j)
#pragma unroll
            for (int k = 0; k < K; ++k)
              if (k < m.leg[j * LANES + lane].nk) qacc_out[(size_t)i * nv + m.leg_qadr[j * LANES + lane] + k] = st.wl[j][k];
        }
      }
      int32_t* aux_row = (aux && live && pass == 0) ? aux + (size_t)i * p.aux_dim : nullptr;
      GxArtObs o = gx_art_obs_row<LANES, NLEG, K, NB>(p, s, m, i, lane, mask, lay, goal.x, goal.y, sn, cr, cl, qro, vro, qlo, vlo, row, aux_row);
      if (pass == 0 && ((p.off_vel >= 0) | (p.off_acc >= 0))) {
        // ego_vel_acc (engine.py:902-929) + update_data's history shift, as for the Point
        int vb = 0;
        if (lane == 0 && live) vb = gx_vel_acc(p, s, i, sn.px, sn.py, sn.r00, sn.r10, last_done, row);
        o.bad |= gx_art_or<LANES>(vb, mask);
      }
      if (pass == 0) {
        // reward_done (engine.py:787-802), guards (:696-699), timeout + counter (:492-493)
        const float dist = o.dist_goal;
        const float last = last_done > 0.0f ? dist : last_dist;
        const float d_dist = last - dist;
        o_reward = d_dist * p.reward_distance;
        o_done = dist < p.goal_size ? 1.0f : 0.0f;
        int flags = o_done > 0.0f ? GX_AUX_GOAL : 0;
        if (gx_abs(d_dist) > 1.0f) { o_done = 1.0f; o_reward = 0.0f; flags |= GX_AUX_JUMP; }
        if (o.bad) { o_reward = 0.0f; o_done = 1.0f; flags |= GX_AUX_GUARD; }
        if (steps > p.num_steps) { o_done = 1.0f; flags |= GX_AUX_TIMEOUT; }
        if (aux_row && lane == 0) { aux_row[0] = flags; aux_row[1] = o.violations; }
        steps = o_done > 0.0f ? 0.0f : steps + 1.0f;
        was_done = o_done;
        prev_dist = dist;
        age += 1;
        o_cost = o.cost;
        if (stats_on && lane == 0 && live) gx_stats_step(s, i, o_reward, o_cost, o_done);
      }
    }
  }

  // ---- write back
  if (live && lane == 0) {
    s.d2[i] = make_float4(prev_dist, was_done > 0.0f ? -steps : steps, __uint_as_float((uint32_t)age),
                          __uint_as_float((uint32_t)(age >> 32)));
    if (MODE == GX_ART_STEP) { reward[i] = o_reward; done[i] = o_done; cost[i] = o_cost; }
  }
  if (MODE == GX_ART_RESET_DONE) {
    // only the done envs' rows are (over)written
    __syncwarp(mask);
    float* dst = obs + (size_t)env * p.obs_dim;
    for (int k = lane; k < p.obs_dim; k += LANES) dst[k] = row[k];
    return;
  }
  __syncwarp();
  {
    const int total = rows * p.obs_dim;
    float* dst = obs + (size_t)wbase * p.obs_dim;
    if ((total & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
      const float4* src4 = reinterpret_cast<const float4*>(tile);
      float4* dst4 = reinterpret_cast<float4*>(dst);
      for (int k = lw; k < (total >> 2); k += 32) dst4[k] = src4[k];
    } else {
      for (int k = lw; k < total; k += 32) dst[k] = tile[k];
    }
  }
}

// Host launcher for one robot shape: picks the mode / auto-reset instance and, for the step,
// the register budget by grid size: 2 resident blocks per SM (no spills) up to 3 waves, MINB_BIG
// beyond -- measured per shape at 65 536 envs: ant 2/4/6/8/10/12 -> 512/348/333/313/368/417 us (8 wins:
// 64 registers, spills served by L1, 32 warps per SM hide the dependent-instruction latency); walker
// 4/6/8 -> 599/729/807 us; swimmer indifferent.
template <int L, int NL, int K, int NB, int MINB_BIG>
static cudaError_t gx_art_launch_shape(const GxArtLaunch& a) {
  const int n = a.p->n, block = 128;
  const int grid = (n * L + block - 1) / block;
  const size_t sh = (size_t)(block / L) * a.p->obs_dim * sizeof(float);
#define GX_ART_GO(MODE, AR, MB)                                                                              \
  do {                                                                                                       \
    auto kern = gx_art_kernel<L, NL, K, NB, MODE, AR, MB>;                                                   \
    if (sh > 48 * 1024) { /* large lidars on a 1-lane robot: opt in to more dynamic shared memory */        \
      cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);     \
      if (e_ != cudaSuccess) return e_;                                                                      \
    }                                                                                                        \
    kern<<<grid, block, sh, a.stream>>>(*a.p, *a.s, *a.art, a.pool, a.layout, a.action, a.obs_in, a.obs,     \
                                        a.act_out, a.reward, a.done, a.cost, a.qacc, a.aux, a.stats_on);     \
  } while (0)
  if (a.mode == GX_ART_RESET) GX_ART_GO(GX_ART_RESET, false, 2);
  else if (a.mode == GX_ART_RESET_DONE) GX_ART_GO(GX_ART_RESET_DONE, false, 2);
  else if (grid > 3 * a.sm_count) { if (a.autoreset) GX_ART_GO(GX_ART_STEP, true, MINB_BIG); else GX_ART_GO(GX_ART_STEP, false, MINB_BIG); }
  else { if (a.autoreset) GX_ART_GO(GX_ART_STEP, true, 2); else GX_ART_GO(GX_ART_STEP, false, 2); }
#undef GX_ART_GO
  return cudaGetLastError();
}

#endif  // GX_ART_KERNELS_CUH_
