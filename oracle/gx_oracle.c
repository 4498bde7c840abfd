/* gx_oracle.c -- CPU restatement of guardX's batched safe_rl_envs step.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under guardx_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs do, as the checker and as the CPU
 * baseline.  The product (libguardx.so) has no CPU path.
 *
 * What it restates (paths relative to /root/reference/safe_rl_envs/safe_rl_envs/):
 *   envs/engine.py:426-431, 454-505   Engine.update_data / reset / step / reset_done
 *   envs/engine.py:533-657            placements, sample_layout, layout2qpos, mjx_reset
 *   envs/engine.py:659-731            mjx_step, convert_action, mjx_reset_done
 *   envs/engine.py:738-929            obs, goal_pos, reward_done, cost, ego_xy,
 *                                     obs_compass, obs_lidar
 *   envs/world.py:104-126, 310-333    which bodies/joints the compiled model has
 *   xmls/point.xml:1-41               the Point robot
 * and, for the physics the reference delegates to an ABSENT third-party
 * dependency -- `mujoco` + `mujoco_mjx`, un-pinned in requirements.txt:3-4, by
 * API use 3.1.2-3.1.x (mjx.put_model/put_data engine.py:231-232,
 * mjx.device_get_into engine.py:1038) -- the published MuJoCo/MJX algorithm
 * (forward kinematics, composite inertia, bias forces, passive damping,
 * affine-bias actuators with ctrl/force clamps, semi-implicit Euler with
 * implicit joint damping) specialised to the model World.build() emits for
 * xmls/point.xml (assumption tags A1-A4).
 *
 * PARITY STATUS: the reference ships no test, golden vector or fixture for this
 * path and cannot be executed in this image (no jax / mujoco wheels).  So:
 *   - obs / lidar / compass / reward / done / cost / layout logic
 *     (engine.py:546-621, 738-900) is PINNED against the reference's own source
 *     executed under a numpy stand-in for jax.numpy (tools/make_golden.py ->
 *     tests/golden/refsrc_*.npz, tests/test_oracle_golden.py);
 *   - the mjx.step / mjx.forward physics (engine.py:655, 668, 721) is
 *     **parity unpinned**: it is a restatement of the published algorithm under
 *     assumptions A1-A4, cross-checked only against the independent dense
 *     numpy restatement of mjx.step on the compiled Point model
 *     (tests/test_art_cpu.py::test_numpy_mjx_step_equals_closed_form_point_oracle).
 *     tests/tools/dump_reference_fixture.py regenerates a true fixture on any
 *     machine that has jax + mujoco.
 *
 * Arithmetic: fp32 throughout, compiled -ffp-contract=off, elementary functions
 * from guardx_b200/csrc/gx_math.h (shared with the kernels so that integer
 * outputs can be compared bit-exactly; see that header).  RNG: Philox4x32-10,
 * keyed as documented in DESIGN.md (the reference's threefry streams are not
 * reproduced; parity injects layouts and actions).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#include "../include/guardx.h"
#include "../guardx_b200/csrc/gx_math.h"
#include "../guardx_b200/csrc/gx_philox.h"
#include "../guardx_b200/csrc/gx_params.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define GXO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ types */

/* mjx.Data restricted to what Engine reads back (engine.py:754-766, 781-784,
   806-807): qpos/qvel/ctrl/qacc of the robot and the robot's xpos/xmat from the
   most recent forward pass.  Goal and hazard bodies never move (no actuator, no
   contact, no gravity component on their x/y slides; engine.py:365-366, 378-379,
   world.py:316-323), so their xpos is their layout xy. */
typedef struct {
  float qpos[3], qvel[3], ctrl[3], qacc[3];
  float xpos[3];     /* robot body position of the last forward (pre-step qpos) */
  float xmat[2][2];  /* upper-left 2x2 of the robot xmat of the last forward    */
  float goal[2];
  float hazards[GX_MAX_HAZARDS][2];
  /* Engine-level history (engine.py:426-431, 489-493) */
  float last_dist_goal; /* goal_pos(last_data), i.e. of the forward before the last */
  float done;           /* Engine._done  */
  /* history only ego_vel_acc reads (engine.py:902-929) */
  float last_xpos[2];   /* Engine._last_data.xpos[robot] */
  float last_done;      /* Engine._last_done */
  int hist;             /* 0: _done is None (no step yet); 1: _last_done is None; 2: neither */
  float steps;          /* Engine._steps (fp32 in the reference, engine.py:463) */
  /* episode accumulators (host side of ppo_train.py:370-376) */
  float ep_ret, ep_cost;
  int32_t ep_len;
} gxo_world;

struct gxo_env {
  gx_config cfg;
  gx_params p;
  int n, nh, nb, obs_dim, lay_dim;
  gxo_world* w;
  float* pool;       /* [pool_size, lay_dim] */
  int32_t* pool_tries;
  int pool_size;
  uint64_t epoch;    /* number of pool generations so far */
  uint64_t t_global; /* number of step() calls so far */
  int have_reset;
  int threads;
  double stats[8];
};
typedef struct gxo_env gxo_env;

static __thread char gxo_err[256];
GXO_API const char* gxo_last_error(void) { return gxo_err; }

/* ------------------------------------------------- mjx.forward / mjx.step */

/* mjx kinematics for the robot body: three joints on one body, slide-x, slide-y
   (world-frame axes because they precede the hinge), hinge-z.  Quaternion path
   as MJX does it: axis_angle_to_quat -> quat_mul with the identity body quat ->
   normalize -> quat_to_mat.  (A4) */
static void mjx_kinematics(const gx_params* m, const float qpos[3], float xpos[3],
                           float xmat[2][2]) {
  float half = qpos[2] * 0.5f;
  float s, c;
  gx_sincos(half, &s, &c);
  float n = gx_sqrt(c * c + s * s);
  float w = gx_div(c, n), z = gx_div(s, n);
  float ww = w * w, zz = z * z, wz = w * z;
  xmat[0][0] = ww - zz;   /* q00 + q11 - q22 - q33 */
  xmat[1][0] = 2.0f * wz; /* 2 (q12 + q03) */
  xmat[0][1] = -xmat[1][0];
  xmat[1][1] = ww - zz;
  xpos[0] = qpos[0];
  xpos[1] = qpos[1];
  xpos[2] = m->body_z;
}

/* Solve [[d0,0,a],[0,d1,b],[a,b,mi]] x = f by block elimination (the robot block
   of the dense mass matrix; the other 18 DoF are decoupled and static). */
static void solve_robot_block(float d0, float d1, float mi, float a, float b, const float f[3],
                              float x[3]) {
  float t0 = gx_div(f[0], d0), t1 = gx_div(f[1], d1);
  float schur = mi - (gx_div(a * a, d0) + gx_div(b * b, d1));
  float alpha = gx_div(f[2] - (a * t0 + b * t1), schur);
  x[0] = t0 - gx_div(a * alpha, d0);
  x[1] = t1 - gx_div(b * alpha, d1);
  x[2] = alpha;
}

/* One mjx.step(model, data) for the Point model (call sites engine.py:668, 721).
   forward: kinematics, mass matrix, bias, passive, actuation, qacc (no active
   constraint rows: A2); then Euler with implicit damping. */
static void mjx_step_point(const gx_params* m, gxo_world* d) {
  mjx_kinematics(m, d->qpos, d->xpos, d->xmat);
  float cth = d->xmat[0][0], sth = d->xmat[1][0];
  /* mass matrix robot block: M = [[m,0,-mc sin],[0,m,mc cos],[.,.,I_joint]] */
  float m02 = -(m->mc * sth), m12 = m->mc * cth;
  /* bias (centrifugal) c(q,v) */
  float w2 = d->qvel[2] * d->qvel[2];
  float bias[3] = {-(m->mc * cth) * w2, -(m->mc * sth) * w2, 0.0f};
  float qfrc[3];
  for (int i = 0; i < 3; ++i) {
    float passive = -(m->damp[i] * d->qvel[i]);
    /* actuation: clamp ctrl, affine bias on actuator velocity, clamp force */
    float u = d->ctrl[i];
    if (m->ctrllimited) u = gx_clip(u, m->ctrl_lo, m->ctrl_hi);
    float act_vel = m->gear[i] * d->qvel[i];
    float force = m->gain * u + m->bias2 * act_vel;
    if (m->forcelimited) force = gx_clip(force, m->force_lo, m->force_hi);
    float actuator = m->gear[i] * force;
    qfrc[i] = (passive - bias[i]) + actuator; /* qfrc_smooth */
  }
  solve_robot_block(m->mass, m->mass, m->inertia, m02, m12, qfrc, d->qacc);
  /* Euler, implicit in damping: (M + dt diag(damping)) qacc_e = qfrc_smooth */
  float qacc_e[3];
  solve_robot_block(m->mh0, m->mh1, m->mh2, m02, m12, qfrc, qacc_e);
  for (int i = 0; i < 3; ++i) {
    d->qvel[i] = d->qvel[i] + m->dt * qacc_e[i];
    d->qpos[i] = d->qpos[i] + m->dt * d->qvel[i];
  }
}

/* ------------------------------------------------- observation functions */

/* ego_xy (engine.py:817-826) / obs_compass (:834-844): matmul(world_3vec,
   robot_mat)[:2], a row vector times R.  The z row of R is (0,0,1) for a planar
   robot, so the z component of world_3vec contributes an exact zero. */
static void ego_xy(const gxo_world* d, const float pos[2], float out[2]) {
  float dx = pos[0] - d->xpos[0], dy = pos[1] - d->xpos[1];
  out[0] = dx * d->xmat[0][0] + dy * d->xmat[1][0];
  out[1] = dx * d->xmat[0][1] + dy * d->xmat[1][1];
}

/* obs_lidar for one object (engine.py:872-899).  Returns the bin index as the
   reference computes it (may equal nb at the fp32 wrap edge). */
static int lidar_from_ego(const gx_params* p, const float z[2], float* obs);
static int obs_lidar_one(const gx_params* p, const gxo_world* d, const float pos[2], float* obs) {
  float z[2];
  ego_xy(d, pos, z);
  return lidar_from_ego(p, z, obs);
}
/* the binning part of obs_lidar (engine.py:877-899) for an egocentric vector z */
static int lidar_from_ego(const gx_params* p, const float z[2], float* obs) {
  const int nb = p->nb;
  float dist = gx_sqrt(z[0] * z[0] + z[1] * z[1]);
  float angle = gx_atan2(z[1], z[0]);
  if (angle < 0.0f) angle = angle + GX_TWO_PI_F; /* jnp `%` on [-pi,pi] */
  float q = gx_div(angle, p->bin_size);
  int bin = (q >= 0.0f && q < 1.0e9f) ? (int)q : 0; /* NaN: deterministic 0 */
  float bin_angle = p->bin_size * (float)bin;
  float sensor;
  if (p->use_exp) {
    sensor = gx_exp(p->neg_exp_gain * dist);
  } else {
    sensor = gx_div(gx_max(0.0f, p->lidar_max_dist - dist), p->lidar_max_dist);
  }
  /* JAX indexing: gather clamps, scatter out of bounds is dropped */
  int gbin = bin < nb ? bin : nb - 1;
  float cur = gx_max(obs[gbin], sensor);
  if (bin < nb) obs[bin] = cur;
  if (p->alias) {
    float alias = gx_div(angle - bin_angle, p->bin_size);
    int bp = (bin + 1) % nb, bm = (bin - 1 + nb) % nb;
    obs[bp] = gx_max(obs[bp], alias * sensor);
    obs[bm] = gx_max(obs[bm], (1.0f - alias) * sensor);
  }
  return bin;
}

/* ego_vel_acc (engine.py:902-929).  cur/last/last_last are robot xpos xy of
   data / last_data / last_last_data; have_last / have_last_last say whether
   last_done / last_last_done are not None. */
static void ego_vel_acc(const gx_params* p, const gxo_world* d, const float cur[2], const float last[2],
                        const float last_last[2], int have_last, int have_last_last, float last_done,
                        float last_last_done, float vel[2], float acc[2]) {
  float pl[2] = {cur[0], cur[1]}, pll[2] = {cur[0], cur[1]};
  if (have_last) {
    if (!(last_done > 0.0f)) { pl[0] = last[0]; pl[1] = last[1]; }
    if (have_last_last) {
      if (last_last_done + last_done > 0.0f) { pll[0] = pl[0]; pll[1] = pl[1]; }
      else { pll[0] = last_last[0]; pll[1] = last_last[1]; }
    }
  }
  float vw[2], lvw[2], aw[2];
  for (int k = 0; k < 2; ++k) {
    vw[k] = gx_div(cur[k] - pl[k], p->ctrl_dt);
    lvw[k] = gx_div(pl[k] - pll[k], p->ctrl_dt);
    aw[k] = gx_div(vw[k] - lvw[k], p->ctrl_dt);
  }
  /* matmul(vec, robot_mat)[:2]; the z component of the differences is exactly 0 */
  vel[0] = vw[0] * d->xmat[0][0] + vw[1] * d->xmat[1][0];
  vel[1] = vw[0] * d->xmat[0][1] + vw[1] * d->xmat[1][1];
  acc[0] = aw[0] * d->xmat[0][0] + aw[1] * d->xmat[1][0];
  acc[1] = aw[0] * d->xmat[0][1] + aw[1] * d->xmat[1][1];
}

/* goal_pos (engine.py:780-785) */
static float goal_pos(const gxo_world* d) {
  float dx = d->goal[0] - d->xpos[0], dy = d->goal[1] - d->xpos[1];
  return gx_sqrt(dx * dx + dy * dy);
}

/* Engine.obs (engine.py:738-778): flat concat in sorted key order. */
static void engine_obs(const gx_params* p, const gxo_world* d, float* obs, int32_t* bins,
                       const float* vel, const float* acc) {
  for (int i = 0; i < p->obs_dim; ++i) obs[i] = 0.0f;
  /* vel / acc default to 0: obs(data, None, None, None, None) (engine.py:656, 726) */
  if (p->off_acc >= 0 && acc) { obs[p->off_acc] = acc[0]; obs[p->off_acc + 1] = acc[1]; }
  if (p->off_vel >= 0 && vel) { obs[p->off_vel] = vel[0]; obs[p->off_vel + 1] = vel[1]; }
  if (p->off_ctrl >= 0)
    for (int i = 0; i < 3; ++i) obs[p->off_ctrl + i] = d->ctrl[i];
  if (p->off_compass >= 0) ego_xy(d, d->goal, obs + p->off_compass);
  float scratch[GX_MAX_LIDAR_BINS];
  {
    float* gl = p->off_glidar >= 0 ? obs + p->off_glidar : scratch;
    for (int i = 0; i < p->nb; ++i) gl[i] = 0.0f;
    int b = obs_lidar_one(p, d, d->goal, gl);
    if (bins) bins[0] = b;
  }
  {
    float* hl = p->off_hlidar >= 0 ? obs + p->off_hlidar : scratch;
    for (int i = 0; i < p->nb; ++i) hl[i] = 0.0f;
    for (int h = 0; h < p->nh; ++h) {
      int b = obs_lidar_one(p, d, d->hazards[h], hl);
      if (bins) bins[1 + h] = b;
    }
  }
  if (p->off_qpos >= 0)
    for (int i = 0; i < 3; ++i) obs[p->off_qpos + i] = d->qpos[i];
  if (p->off_qvel >= 0)
    for (int i = 0; i < 3; ++i) obs[p->off_qvel + i] = d->qvel[i];
}

/* cost (engine.py:804-811): sum_h (size - min(dist_h, size)), summed in order. */
static float engine_cost(const gx_params* p, const gxo_world* d, int* violations) {
  float size = p->hazards_size, c = 0.0f;
  int v = 0;
  for (int h = 0; h < p->nh; ++h) {
    float dx = d->hazards[h][0] - d->xpos[0], dy = d->hazards[h][1] - d->xpos[1];
    float dist = gx_sqrt(dx * dx + dy * dy);
    float term = size - gx_min(dist, size);
    if (term > 0.0f) ++v;
    c = c + term;
  }
  if (violations) *violations = v;
  return c;
}

/* ------------------------------------------------------ mjx_step (659-700) */

static void engine_step_one(const gx_params* p, gxo_world* d, const float* action, float* obs,
                            float* reward, float* done, float* cost, float* qacc,
                            int32_t* aux) {
  /* convert_action (engine.py:672-685): R is data.xmat of the previous forward */
  d->ctrl[0] = d->xmat[0][0] * action[0];
  d->ctrl[1] = d->xmat[1][0] * action[0];
  d->ctrl[2] = action[1];
  /* update_data (engine.py:426-431): last_data <- data, last_done <- done */
  float last_dist = d->last_dist_goal; /* goal_pos(last_data) uses last_data.xpos */
  float last_done = d->done;
  const float last_last_done = d->last_done;
  const float x_last[2] = {d->xpos[0], d->xpos[1]};           /* last_data <- data */
  const float x_last_last[2] = {d->last_xpos[0], d->last_xpos[1]}; /* last_last_data <- last_data */
  const int hist = d->hist;
  /* lax.scan(mjx.step, n) (engine.py:689) */
  for (int k = 0; k < p->substeps; ++k) mjx_step_point(p, d);
  int32_t bins[1 + GX_MAX_HAZARDS];
  float vel[2], acc[2];
  {
    const float cur[2] = {d->xpos[0], d->xpos[1]};
    /* obs(): last_last_data = last_data when it is None (engine.py:749-750) */
    ego_vel_acc(p, d, cur, x_last, hist >= 2 ? x_last_last : x_last, hist >= 1, hist >= 2, last_done,
                last_last_done, vel, acc);
  }
  engine_obs(p, d, obs, bins, vel, acc);
  /* reward_done (engine.py:787-802) */
  float dist = goal_pos(d);
  float last = last_done > 0.0f ? dist : last_dist;
  float d_dist = last - dist;
  float r = d_dist * p->reward_distance;
  float dn = dist < p->goal_size ? 1.0f : 0.0f;
  int flags = dn > 0.0f ? GX_AUX_GOAL : 0;
  if (gx_abs(d_dist) > 1.0f) { dn = 1.0f; r = 0.0f; flags |= GX_AUX_JUMP; }
  int violations;
  float c = engine_cost(p, d, &violations);
  /* NaN/Inf guard (engine.py:696-699) */
  int bad = 0;
  for (int i = 0; i < p->obs_dim; ++i) bad |= gx_notfinite(obs[i]);
  if (bad) { r = 0.0f; dn = 1.0f; flags |= GX_AUX_GUARD; }
  /* Engine.step tail (engine.py:492-493) */
  if (d->steps > p->num_steps) { dn = 1.0f; flags |= GX_AUX_TIMEOUT; }
  d->steps = dn > 0.0f ? 0.0f : d->steps + 1.0f;
  d->done = dn;
  d->last_dist_goal = dist; /* this data becomes last_data on the next call */
  d->last_xpos[0] = x_last[0]; d->last_xpos[1] = x_last[1];
  d->last_done = last_done;
  d->hist = hist < 2 ? hist + 1 : 2;
  *reward = r; *done = dn; *cost = c;
  if (qacc) { qacc[0] = d->qacc[0]; qacc[1] = d->qacc[1]; qacc[2] = d->qacc[2]; }
  if (aux) {
    aux[0] = flags; aux[1] = violations;
    for (int i = 0; i < 1 + p->nh; ++i) aux[2 + i] = bins[i];
  }
}

/* --------------------------------------- layout (engine.py:533-639) ------ */

static int class_of(int obj, int nh) {
  /* placement order goal, hazard0.., robot (engine.py:538-542) */
  return obj == 0 ? GX_CLS_GOAL : (obj <= nh ? GX_CLS_HAZARD : GX_CLS_ROBOT);
}

/* draw_placement with placements=None (engine.py:600-601, 617-621): uniform in
   the extents shrunk by keepout.  One Philox call gives x and y. */
static void draw_placement(const gx_params* p, int cls, uint32_t k0, uint32_t k1, uint32_t slot,
                           uint32_t attempt, uint32_t obj, uint32_t t, float xy[2]) {
  gx_u32x4 r = gx_philox4x32_10(slot, attempt, (obj << 8) | t, GX_STREAM_LAYOUT, k0, k1);
  xy[0] = gx_uniform(r.v[0], p->lo[cls][0], p->hi[cls][0]);
  xy[1] = gx_uniform(r.v[1], p->lo[cls][1], p->hi[cls][1]);
}

/* sample_layout (engine.py:546-572).  out: goal xy, robot xy, hazards xy (the
   C-ABI layout row order).  Returns success. */
static int sample_layout(const gx_params* p, uint32_t k0, uint32_t k1, uint32_t slot,
                         uint32_t attempt, float* out) {
  float placed[2 + GX_MAX_HAZARDS][2];
  const int nh = p->nh, nobj = nh + 2;
  int success = 1;
  for (int o = 0; o < nobj; ++o) {
    int cls = class_of(o, nh);
    int conflicted = 1;
    float xy[2] = {-1e30f, -1e30f}; /* stands for the reference's -inf placeholder */
    for (int t = 0; t < p->tries; ++t) {
      float cur[2];
      draw_placement(p, cls, k0, k1, slot, attempt, (uint32_t)o, (uint32_t)t, cur);
      int valid = 1; /* placement_is_valid (engine.py:549-555) */
      for (int q = 0; q < o; ++q) {
        float dx = cur[0] - placed[q][0], dy = cur[1] - placed[q][1];
        float dist = gx_sqrt(dx * dx + dy * dy);
        if (dist < p->keepsum[class_of(q, nh)][cls]) valid = 0;
      }
      if (valid) { xy[0] = cur[0]; xy[1] = cur[1]; conflicted = 0; }
    }
    placed[o][0] = xy[0]; placed[o][1] = xy[1];
    if (conflicted) success = 0;
  }
  {
    float dx = placed[nobj - 1][0] - placed[0][0], dy = placed[nobj - 1][1] - placed[0][1];
    if (gx_sqrt(dx * dx + dy * dy) < p->min_robot_goal) success = 0;
  }
  out[0] = placed[0][0]; out[1] = placed[0][1];
  out[2] = placed[nobj - 1][0]; out[3] = placed[nobj - 1][1];
  for (int h = 0; h < nh; ++h) { out[4 + 2 * h] = placed[1 + h][0]; out[5 + 2 * h] = placed[1 + h][1]; }
  return success;
}

/* reset_layout (engine.py:433-444): the reference filters 1e6 candidates; here
   every pool slot consumes candidates until one succeeds, which yields the same
   distribution (iid accepted layouts). */
static void reset_layout(gxo_env* e) {
  gx_u32x4 key = gx_philox4x32_10((uint32_t)e->epoch, (uint32_t)(e->epoch >> 32), 0u,
                                  GX_STREAM_LAYOUT, e->p.seed_lo, e->p.seed_hi);
#pragma omp parallel for schedule(dynamic, 64) num_threads(e->threads)
  for (int j = 0; j < e->pool_size; ++j) {
    uint32_t attempt = 0;
    const uint32_t max_attempts = 1u << 16; /* then give up: tries = -1 (engine.py:444 would assert) */
    while (attempt < max_attempts &&
           !sample_layout(&e->p, key.v[0], key.v[1], (uint32_t)j, attempt,
                          e->pool + (size_t)j * e->lay_dim))
      ++attempt;
    e->pool_tries[j] = attempt < max_attempts ? (int32_t)(attempt + 1) : -1;
  }
  e->epoch += 1;
}

/* get_layout (engine.py:446-452): one pool index per env, with replacement. */
static const float* get_layout(const gxo_env* e, int i, uint32_t stream) {
  uint64_t g = e->p.env_id_offset + (uint64_t)i;
  gx_u32x4 r = gx_philox4x32_10((uint32_t)g, (uint32_t)e->t_global,
                                (uint32_t)(e->t_global >> 32), stream, e->p.seed_lo, e->p.seed_hi);
  uint32_t idx = gx_randint(r.v[0], (uint32_t)e->pool_size);
  return e->pool + (size_t)idx * e->lay_dim;
}

/* layout2qpos (engine.py:623-639) */
static void layout2qpos(const gxo_env* e, gxo_world* d, const float* lay) {
  d->goal[0] = lay[0]; d->goal[1] = lay[1];
  d->qpos[0] = lay[2]; d->qpos[1] = lay[3]; d->qpos[2] = 0.0f;
  for (int h = 0; h < e->nh; ++h) { d->hazards[h][0] = lay[4 + 2 * h]; d->hazards[h][1] = lay[5 + 2 * h]; }
}

/* mjx_reset (engine.py:644-657): qvel = ctrl = 0, mjx.forward, obs. */
static void mjx_reset_one(const gxo_env* e, gxo_world* d, const float* lay, float* obs) {
  layout2qpos(e, d, lay);
  for (int i = 0; i < 3; ++i) { d->qvel[i] = 0.0f; d->ctrl[i] = 0.0f; d->qacc[i] = 0.0f; }
  mjx_kinematics(&e->p, d->qpos, d->xpos, d->xmat);
  engine_obs(&e->p, d, obs, NULL, NULL, NULL);
  d->last_dist_goal = goal_pos(d);
}

/* mjx_reset_done (engine.py:702-731) for ONE env whose done > 0: masked
   re-placement, a throw-away mjx.step on a copy to produce obs_reset, the
   returned data is the UN-stepped one (its xpos/xmat stay stale). */
static void mjx_reset_done_one(const gxo_env* e, gxo_world* d, const float* lay, float* obs) {
  layout2qpos(e, d, lay);
  for (int i = 0; i < 3; ++i) { d->qvel[i] = 0.0f; d->ctrl[i] = 0.0f; }
  gxo_world tmp = *d;
  for (int k = 0; k < e->p.substeps; ++k) mjx_step_point(&e->p, &tmp);
  engine_obs(&e->p, &tmp, obs, NULL, NULL, NULL);
  /* d->xpos / d->xmat / d->last_dist_goal deliberately untouched */
}

/* ------------------------------------------------------------- C API ---- */

GXO_API int gxo_create(const gx_config* cfg, gxo_env** out) {
  gx_params p;
  memset(&p, 0, sizeof p);
  const char* why = gx_lower_config(cfg, &p);
  if (!why && cfg->robot != GX_ROBOT_POINT) why = "this C oracle restates the Point only; articulated robots: oracle/art_oracle.py";
  if (why || !out) {
    snprintf(gxo_err, sizeof gxo_err, "gxo_create: %s", why ? why : "null out");
    return GX_EINVAL;
  }
  gxo_env* e = (gxo_env*)calloc(1, sizeof *e);
  e->cfg = *cfg; e->p = p;
  e->n = p.n; e->nh = p.nh; e->nb = p.nb; e->obs_dim = p.obs_dim; e->lay_dim = p.lay_dim;
  e->pool_size = p.pool_size;
  e->w = (gxo_world*)calloc((size_t)e->n, sizeof(gxo_world));
  e->pool = (float*)calloc((size_t)e->pool_size * e->lay_dim, sizeof(float));
  e->pool_tries = (int32_t*)calloc((size_t)e->pool_size, sizeof(int32_t));
  e->threads = 1;
  /* xmat of a never-forwarded mjx.Data is what put_data copies from MjData: identity */
  for (int i = 0; i < e->n; ++i) { e->w[i].xmat[0][0] = 1.0f; e->w[i].xmat[1][1] = 1.0f; }
  *out = e;
  return GX_OK;
}

GXO_API int gxo_destroy(gxo_env* e) {
  if (!e) return GX_OK;
  free(e->w); free(e->pool); free(e->pool_tries); free(e);
  return GX_OK;
}

GXO_API void gxo_default_config(gx_config* c) { gx_fill_default_config(c); }
GXO_API int gxo_config_size(void) { return (int)sizeof(gx_config); }
GXO_API void gxo_set_threads(gxo_env* e, int t) { e->threads = t < 1 ? 1 : t; }
GXO_API int gxo_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
GXO_API int gxo_obs_dim(const gxo_env* e) { return e->obs_dim; }
GXO_API int gxo_act_dim(const gxo_env* e) { (void)e; return 2; }
GXO_API int gxo_env_num(const gxo_env* e) { return e->n; }
GXO_API int gxo_pool_size(const gxo_env* e) { return e->pool_size; }
GXO_API int gxo_aux_dim(const gxo_env* e) { return e->p.aux_dim; }
GXO_API int gxo_state_dim(const gxo_env* e) { return e->p.state_dim; }

static void clear_episode(gxo_world* w) { w->ep_ret = w->ep_cost = 0.0f; w->ep_len = 0; }

GXO_API int gxo_reset(gxo_env* e, float* obs) {
  reset_layout(e);
#pragma omp parallel for schedule(static) num_threads(e->threads)
  for (int i = 0; i < e->n; ++i) {
    mjx_reset_one(e, &e->w[i], get_layout(e, i, GX_STREAM_PICK ^ 1u), obs + (size_t)i * e->obs_dim);
    e->w[i].steps = 0.0f;
    clear_episode(&e->w[i]);
  }
  e->have_reset = 1;
  return GX_OK;
}

GXO_API int gxo_reset_with_layout(gxo_env* e, const float* layout, float* obs) {
#pragma omp parallel for schedule(static) num_threads(e->threads)
  for (int i = 0; i < e->n; ++i) {
    mjx_reset_one(e, &e->w[i], layout + (size_t)i * e->lay_dim, obs + (size_t)i * e->obs_dim);
    e->w[i].steps = 0.0f;
    clear_episode(&e->w[i]);
  }
  e->have_reset = 1;
  return GX_OK;
}

GXO_API int gxo_step(gxo_env* e, const float* action, float* obs, float* reward, float* done,
                     float* cost, float* qacc, int32_t* aux) {
  if (!e->have_reset) { snprintf(gxo_err, sizeof gxo_err, "step before reset"); return GX_ESTATE; }
  const int ad = e->p.aux_dim;
#pragma omp parallel for schedule(static) num_threads(e->threads)
  for (int i = 0; i < e->n; ++i)
    engine_step_one(&e->p, &e->w[i], action + 2 * (size_t)i, obs + (size_t)i * e->obs_dim,
                    reward + i, done + i, cost + i, qacc ? qacc + 3 * (size_t)i : NULL,
                    aux ? aux + (size_t)ad * i : NULL);
  /* episode bookkeeping (what ppo_train.py:370-376, 405-411 does on the host),
     serial so that the double sums are order-stable */
  for (int i = 0; i < e->n; ++i) {
    gxo_world* w = &e->w[i];
    w->ep_ret += reward[i]; w->ep_cost += cost[i]; w->ep_len += 1;
    e->stats[6] += cost[i]; e->stats[7] += 1.0;
    if (done[i] > 0.0f) {
      double* s = e->stats;
      s[0] += 1.0; s[1] += w->ep_ret; s[2] += w->ep_cost; s[3] += w->ep_len;
      s[4] += (double)w->ep_ret * w->ep_ret; s[5] += (double)w->ep_cost * w->ep_cost;
      clear_episode(w);
    }
  }
  e->t_global += 1;
  return GX_OK;
}

GXO_API int gxo_reset_done(gxo_env* e, const float* obs_in, float* obs_out, const float* layout) {
  if (!e->have_reset) { snprintf(gxo_err, sizeof gxo_err, "reset_done before reset"); return GX_ESTATE; }
#pragma omp parallel for schedule(static) num_threads(e->threads)
  for (int i = 0; i < e->n; ++i) {
    float* o = obs_out + (size_t)i * e->obs_dim;
    if (e->w[i].done > 0.0f) {
      const float* lay = layout ? layout + (size_t)i * e->lay_dim : get_layout(e, i, GX_STREAM_PICK);
      mjx_reset_done_one(e, &e->w[i], lay, o);
    } else if (obs_in != obs_out) {
      memcpy(o, obs_in + (size_t)i * e->obs_dim, sizeof(float) * e->obs_dim);
    }
  }
  return GX_OK;
}

GXO_API int gxo_get_state(gxo_env* e, float* state) {
  const int sd = e->p.state_dim;
  for (int i = 0; i < e->n; ++i) {
    const gxo_world* w = &e->w[i];
    float* s = state + (size_t)i * sd;
    for (int k = 0; k < 3; ++k) { s[k] = w->qpos[k]; s[3 + k] = w->qvel[k]; }
    s[6] = w->xmat[0][0]; s[7] = w->xmat[1][0]; s[8] = w->last_dist_goal; s[9] = w->done;
    s[10] = w->steps; s[11] = w->goal[0]; s[12] = w->goal[1];
    for (int h = 0; h < e->nh; ++h) { s[13 + 2 * h] = w->hazards[h][0]; s[14 + 2 * h] = w->hazards[h][1]; }
    float* x = s + 13 + 2 * e->nh;
    /* the history columns are only kept (and only meaningful) with observe_vel / observe_acc */
    const int va = e->p.off_vel >= 0 || e->p.off_acc >= 0;
    x[0] = va ? w->xpos[0] : 0.0f; x[1] = va ? w->xpos[1] : 0.0f;
    x[2] = va ? w->last_xpos[0] : 0.0f; x[3] = va ? w->last_xpos[1] : 0.0f;
    x[4] = va ? w->last_done : 0.0f; x[5] = va ? (float)w->hist : 0.0f;
  }
  return GX_OK;
}

GXO_API int gxo_set_state(gxo_env* e, const float* state) {
  const int sd = e->p.state_dim;
  for (int i = 0; i < e->n; ++i) {
    gxo_world* w = &e->w[i];
    const float* s = state + (size_t)i * sd;
    for (int k = 0; k < 3; ++k) { w->qpos[k] = s[k]; w->qvel[k] = s[3 + k]; }
    w->xmat[0][0] = s[6]; w->xmat[1][0] = s[7]; w->xmat[0][1] = -s[7]; w->xmat[1][1] = s[6];
    const float* x = s + 13 + 2 * e->nh;
    w->xpos[0] = x[0]; w->xpos[1] = x[1]; w->xpos[2] = e->p.body_z;
    w->last_xpos[0] = x[2]; w->last_xpos[1] = x[3]; w->last_done = x[4]; w->hist = (int)x[5];
    w->last_dist_goal = s[8]; w->done = s[9]; w->steps = s[10];
    w->goal[0] = s[11]; w->goal[1] = s[12];
    for (int h = 0; h < e->nh; ++h) { w->hazards[h][0] = s[13 + 2 * h]; w->hazards[h][1] = s[14 + 2 * h]; }
  }
  e->have_reset = 1;
  return GX_OK;
}

/* counters that key the Philox draws; lets a test put both implementations at
   the same point of the stream */
GXO_API void gxo_set_counters(gxo_env* e, uint64_t t_global, uint64_t epoch) { e->t_global = t_global; e->epoch = epoch; }

GXO_API int gxo_pool_read(gxo_env* e, float* pool, int32_t* tries) {
  if (pool) memcpy(pool, e->pool, sizeof(float) * (size_t)e->pool_size * e->lay_dim);
  if (tries) memcpy(tries, e->pool_tries, sizeof(int32_t) * (size_t)e->pool_size);
  return GX_OK;
}

/* Synthetic rollout as trpo/myTest.py:28-49 drives the reference (a ~ U(-1,1)),
   with the PPO loop's reset_done after every step (ppo_train.py:387-417). */
GXO_API int gxo_rollout(gxo_env* e, int32_t T, float* obs, float* act, float* reward, float* cost,
                        float* done) {
  if (!e->have_reset) { snprintf(gxo_err, sizeof gxo_err, "rollout before reset"); return GX_ESTATE; }
  const int n = e->n, od = e->obs_dim;
  float* a = (float*)malloc(sizeof(float) * 2 * (size_t)n);
  float* o = (float*)malloc(sizeof(float) * (size_t)od * n);
  float* r = (float*)malloc(sizeof(float) * 3 * (size_t)n);
  for (int t = 0; t < T; ++t) {
#pragma omp parallel for schedule(static) num_threads(e->threads)
    for (int i = 0; i < n; ++i) {
      uint64_t g = e->p.env_id_offset + (uint64_t)i;
      gx_u32x4 u = gx_philox4x32_10((uint32_t)g, (uint32_t)e->t_global,
                                    (uint32_t)(e->t_global >> 32), GX_STREAM_ACTION,
                                    e->p.seed_lo, e->p.seed_hi);
      a[2 * i] = gx_uniform(u.v[0], -1.0f, 1.0f);
      a[2 * i + 1] = gx_uniform(u.v[1], -1.0f, 1.0f);
    }
    gxo_step(e, a, o, r, r + n, r + 2 * n, NULL, NULL);
    gxo_reset_done(e, o, o, NULL);
    if (act) memcpy(act + (size_t)t * n * 2, a, sizeof(float) * 2 * (size_t)n);
    if (obs) memcpy(obs + (size_t)t * n * od, o, sizeof(float) * (size_t)od * n);
    if (reward) memcpy(reward + (size_t)t * n, r, sizeof(float) * n);
    if (done) memcpy(done + (size_t)t * n, r + n, sizeof(float) * n);
    if (cost) memcpy(cost + (size_t)t * n, r + 2 * n, sizeof(float) * n);
  }
  free(a); free(o); free(r);
  return GX_OK;
}

GXO_API int gxo_stats_read(gxo_env* e, double* out8) { memcpy(out8, e->stats, sizeof e->stats); return GX_OK; }
GXO_API int gxo_stats_reset(gxo_env* e) { memset(e->stats, 0, sizeof e->stats); return GX_OK; }

/* ------------------------------------------- test hooks for the shared math */
GXO_API void gxo_v_sincos(const float* x, float* s, float* c, int n) { for (int i = 0; i < n; ++i) gx_sincos(x[i], s + i, c + i); }
GXO_API void gxo_v_atan2(const float* y, const float* x, float* o, int n) { for (int i = 0; i < n; ++i) o[i] = gx_atan2(y[i], x[i]); }
GXO_API void gxo_v_exp(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) o[i] = gx_exp(x[i]); }
GXO_API void gxo_v_max(const float* a, const float* b, float* o, int n) { for (int i = 0; i < n; ++i) o[i] = gx_max(a[i], b[i]); }
GXO_API void gxo_v_min(const float* a, const float* b, float* o, int n) { for (int i = 0; i < n; ++i) o[i] = gx_min(a[i], b[i]); }
GXO_API void gxo_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  gx_u32x4 r = gx_philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = r.v[i];
}

/* ---- single-env pieces, for the golden-vector tests (tests/golden) ------- */
static void world_from_pose(gxo_world* d, const float* xpos3, const float* xmat9) {
  memset(d, 0, sizeof *d);
  d->xpos[0] = xpos3[0]; d->xpos[1] = xpos3[1]; d->xpos[2] = xpos3[2];
  d->xmat[0][0] = xmat9[0]; d->xmat[0][1] = xmat9[1]; d->xmat[1][0] = xmat9[3]; d->xmat[1][1] = xmat9[4];
}
/* obs_lidar(data, positions) (engine.py:846-900) */
GXO_API int gxo_piece_lidar(const gx_config* cfg, const float* xpos3, const float* xmat9,
                            const float* positions, int npos, float* obs, int32_t* bins) {
  gx_params p; memset(&p, 0, sizeof p);
  if (gx_lower_config(cfg, &p)) return GX_EINVAL;
  gxo_world d; world_from_pose(&d, xpos3, xmat9);
  for (int i = 0; i < p.nb; ++i) obs[i] = 0.0f;
  for (int k = 0; k < npos; ++k) { int b = obs_lidar_one(&p, &d, positions + 2 * k, obs); if (bins) bins[k] = b; }
  return GX_OK;
}
/* ego_xy / obs_compass with the FULL 3-D robot_mat (engine.py:817-826: matmul(world_3vec, robot_mat)[:2],
   world_3vec = (x, y, 0) - robot_pos): what a free-joint robot (Doggo) needs -- a tilted torso mixes z in */
static void ego_xy3(const float* xpos3, const float* xmat9, const float pos[2], float out[2]) {
  float dx = pos[0] - xpos3[0], dy = pos[1] - xpos3[1], dz = 0.0f - xpos3[2];
  out[0] = dx * xmat9[0] + dy * xmat9[3] + dz * xmat9[6];
  out[1] = dx * xmat9[1] + dy * xmat9[4] + dz * xmat9[7];
}
GXO_API int gxo_piece_lidar3(const gx_config* cfg, const float* xpos3, const float* xmat9,
                             const float* positions, int npos, float* obs, int32_t* bins) {
  gx_params p; memset(&p, 0, sizeof p);
  if (gx_lower_config(cfg, &p)) return GX_EINVAL;
  for (int i = 0; i < p.nb; ++i) obs[i] = 0.0f;
  for (int k = 0; k < npos; ++k) {
    float z[2];
    ego_xy3(xpos3, xmat9, positions + 2 * k, z);
    int b = lidar_from_ego(&p, z, obs);
    if (bins) bins[k] = b;
  }
  return GX_OK;
}
GXO_API int gxo_piece_compass3(const float* xpos3, const float* xmat9, const float* pos, float* out2) {
  ego_xy3(xpos3, xmat9, pos, out2);
  return GX_OK;
}
/* obs_compass (engine.py:834-844) */
GXO_API int gxo_piece_compass(const float* xpos3, const float* xmat9, const float* pos, float* out2) {
  gxo_world d; world_from_pose(&d, xpos3, xmat9);
  ego_xy(&d, pos, out2);
  return GX_OK;
}
/* reward_done (engine.py:787-802): robot/goal xy now and at last_data */
GXO_API int gxo_piece_reward_done(const gx_config* cfg, const float* robot_xy, const float* goal_xy,
                                  const float* last_robot_xy, const float* last_goal_xy,
                                  float last_done, float* reward, float* done) {
  gx_params p; memset(&p, 0, sizeof p);
  if (gx_lower_config(cfg, &p)) return GX_EINVAL;
  gxo_world d; memset(&d, 0, sizeof d);
  d.xpos[0] = robot_xy[0]; d.xpos[1] = robot_xy[1]; d.goal[0] = goal_xy[0]; d.goal[1] = goal_xy[1];
  gxo_world l; memset(&l, 0, sizeof l);
  l.xpos[0] = last_robot_xy[0]; l.xpos[1] = last_robot_xy[1]; l.goal[0] = last_goal_xy[0]; l.goal[1] = last_goal_xy[1];
  float dist = goal_pos(&d);
  float last = last_done > 0.0f ? dist : goal_pos(&l);
  float d_dist = last - dist;
  float r = d_dist * p.reward_distance;
  float dn = dist < p.goal_size ? 1.0f : 0.0f;
  if (gx_abs(d_dist) > 1.0f) { dn = 1.0f; r = 0.0f; }
  *reward = r; *done = dn;
  return GX_OK;
}
/* ego_vel_acc (engine.py:902-929); have: 0 = last_done None, 1 = last_last_done None, 2 = neither */
GXO_API int gxo_piece_vel_acc(const gx_config* cfg, const float* xmat9, const float* cur, const float* last,
                              const float* last_last, int have, float last_done, float last_last_done,
                              float* vel, float* acc) {
  gx_params p; memset(&p, 0, sizeof p);
  if (gx_lower_config(cfg, &p)) return GX_EINVAL;
  const float x3[3] = {cur[0], cur[1], 0.1f};
  gxo_world d; world_from_pose(&d, x3, xmat9);
  ego_vel_acc(&p, &d, cur, last, last_last, have >= 1, have >= 2, last_done, last_last_done, vel, acc);
  return GX_OK;
}
/* cost (engine.py:804-811) */
GXO_API int gxo_piece_cost(const gx_config* cfg, const float* robot_xy, const float* hazards_xy, float* cost) {
  gx_params p; memset(&p, 0, sizeof p);
  if (gx_lower_config(cfg, &p)) return GX_EINVAL;
  gxo_world d; memset(&d, 0, sizeof d);
  d.xpos[0] = robot_xy[0]; d.xpos[1] = robot_xy[1];
  for (int h = 0; h < p.nh; ++h) { d.hazards[h][0] = hazards_xy[2 * h]; d.hazards[h][1] = hazards_xy[2 * h + 1]; }
  *cost = engine_cost(&p, &d, NULL);
  return GX_OK;
}
/* placement validity + acceptance of sample_layout (engine.py:546-572) given the
   DRAWS (so the golden test can feed the reference's own draws): draws is
   [nobj, tries, 2] in placement order; out as gxo_sample_layout. */
GXO_API int gxo_piece_layout_from_draws(const gx_config* cfg, const float* draws, float* out, int* success) {
  gx_params p; memset(&p, 0, sizeof p);
  if (gx_lower_config(cfg, &p)) return GX_EINVAL;
  float placed[2 + GX_MAX_HAZARDS][2];
  const int nh = p.nh, nobj = nh + 2;
  int ok = 1;
  for (int o = 0; o < nobj; ++o) {
    int cls = class_of(o, nh), conflicted = 1;
    float xy[2] = {-1e30f, -1e30f};
    for (int t = 0; t < p.tries; ++t) {
      const float* cur = draws + ((size_t)o * p.tries + t) * 2;
      int valid = 1;
      for (int q = 0; q < o; ++q) {
        float dx = cur[0] - placed[q][0], dy = cur[1] - placed[q][1];
        if (gx_sqrt(dx * dx + dy * dy) < p.keepsum[class_of(q, nh)][cls]) valid = 0;
      }
      if (valid) { xy[0] = cur[0]; xy[1] = cur[1]; conflicted = 0; }
    }
    placed[o][0] = xy[0]; placed[o][1] = xy[1];
    if (conflicted) ok = 0;
  }
  float dx = placed[nobj - 1][0] - placed[0][0], dy = placed[nobj - 1][1] - placed[0][1];
  if (gx_sqrt(dx * dx + dy * dy) < p.min_robot_goal) ok = 0;
  out[0] = placed[0][0]; out[1] = placed[0][1]; out[2] = placed[nobj - 1][0]; out[3] = placed[nobj - 1][1];
  for (int h = 0; h < nh; ++h) { out[4 + 2 * h] = placed[1 + h][0]; out[5 + 2 * h] = placed[1 + h][1]; }
  *success = ok;
  return GX_OK;
}
GXO_API int gxo_sample_layout(const gx_config* cfg, uint32_t k0, uint32_t k1, uint32_t slot,
                              uint32_t attempt, float* out) {
  gx_params p; memset(&p, 0, sizeof p);
  if (gx_lower_config(cfg, &p)) return GX_EINVAL;
  return sample_layout(&p, k0, k1, slot, attempt, out);
}

/* ---- PPOBufferX.finish_path (safe_rl_libX/ppo/ppo_train.py:61-113) restated:
   per env, per path slice [start, end) closed by done (or by the epoch end),
   deltas in fp32, discount_cumsum as scipy.signal.lfilter does it (fp64 direct
   form), rewards-to-go over rews + [last_val] with the last element dropped.
   Time-major [T, n] buffers. */
static void discount_cumsum_rev(const double* x, int len, double discount, double* y) {
  double acc = 0.0; /* lfilter([1],[1,-discount], x[::-1])[::-1] */
  for (int k = len - 1; k >= 0; --k) { acc = x[k] + discount * acc; y[k] = acc; }
}
GXO_API int gxo_gae(const float* reward, const float* value, const float* done, const float* last_value,
                    int T, int n, double gamma, double lam, float* adv, float* ret) {
  double* x = (double*)malloc(sizeof(double) * (size_t)(T + 1) * 2);
  double* y = x + (T + 1);
  const float gf = (float)gamma;
  for (int i = 0; i < n; ++i) {
    int start = 0;
    for (int t = 0; t < T; ++t) {
      const int closed = done[(size_t)t * n + i] > 0.0f;
      if (!closed && t != T - 1) continue;
      const int end = t + 1, len = end - start;
      const float lv = closed ? 0.0f : (last_value ? last_value[i] : 0.0f);
      for (int k = 0; k < len; ++k) { /* deltas = rews[:-1] + gamma * vals[1:] - vals[:-1] */
        const size_t q = (size_t)(start + k) * n + i;
        const float vn = (k + 1 < len) ? value[q + n] : lv;
        float g = gf * vn;
        float d = reward[q] + g;
        d = d - value[q];
        x[k] = (double)d;
      }
      discount_cumsum_rev(x, len, gamma * lam, y);
      for (int k = 0; k < len; ++k) adv[(size_t)(start + k) * n + i] = (float)y[k];
      for (int k = 0; k < len; ++k) x[k] = (double)reward[(size_t)(start + k) * n + i];
      x[len] = (double)lv;
      discount_cumsum_rev(x, len + 1, gamma, y);
      for (int k = 0; k < len; ++k) ret[(size_t)(start + k) * n + i] = (float)y[k];
      start = end;
    }
  }
  free(x);
  return GX_OK;
}
