// gx_params.h -- lowering of gx_config (Python-level doubles) to the fp32
// constants the transition function uses.  Host-only, plain C, shared by
// libguardx.so and the CPU oracle so that both start from the same bits.
// Each conversion happens where JAX's weak-type promotion does it in the
// reference (engine.py line cited per field).
#ifndef GX_PARAMS_H_
#define GX_PARAMS_H_

#include <stdint.h>
#include "../../include/guardx.h"

/* placement classes, in the order the reference places them (engine.py:538-542) */
enum { GX_CLS_GOAL = 0, GX_CLS_HAZARD = 1, GX_CLS_ROBOT = 2 };

typedef struct gx_params {
  int32_t n, nh, nb, obs_dim;
  int32_t nu, nv, nq; /* widths of ctrl, qvel and qpos in the observation (3, 3, 3 for the Point) */
  int32_t substeps, alias, use_exp, tries;
  int32_t off_ctrl, off_compass, off_glidar, off_hlidar, off_qpos, off_qvel, off_acc, off_vel;
  int32_t lay_dim, state_dim, aux_dim, pool_size;
  float ctrl_dt;         /* float32(self.dt), divides the finite differences  engine.py:235, 916-923 */
  float num_steps;       /* `self._steps > self.num_steps`, fp32 compare  engine.py:492 */
  float goal_size;       /* engine.py:798 */
  float hazards_size;    /* engine.py:809-810 */
  float reward_distance; /* engine.py:797 */
  float lidar_max_dist;  /* engine.py:888 */
  float neg_exp_gain;    /* `-self.lidar_exp_gain * dist`  engine.py:886 */
  float bin_size;        /* float32((2*pi)/lidar_num_bins) engine.py:880 */
  /* compiled model (gx_point_model) + the two derived diagonals */
  float dt, body_z, mass, mc, inertia;
  float damp[3], gear[3];
  float gain, bias2, ctrl_lo, ctrl_hi, force_lo, force_hi;
  int32_t ctrllimited, forcelimited;
  float mh0, mh1, mh2; /* diag(M) + dt*damping: mjx euler, implicit damping */
  /* layout sampling */
  float lo[3][2], hi[3][2]; /* constrain_placement, engine.py:574-577, 601 */
  float keepsum[3][3];      /* other_keepout + margin + keepout, engine.py:554; [other][self] */
  float min_robot_goal;     /* engine.py:571 */
  uint32_t seed_lo, seed_hi;
  uint64_t env_id_offset;
} gx_params;

static inline int gx_auto_pool_size(int n) {
  long want = 2L * n, p = 16384;
  while (p < want && p < 262144) p *= 2;
  return (int)p;
}

/* returns 0, or a static message describing what is unsupported */
static inline const char* gx_lower_config(const gx_config* c, gx_params* p) {
  if (!c) return "null config";
  if (c->abi_version != GX_ABI_VERSION) return "abi_version mismatch";
  if (c->env_num < 1) return "env_num must be >= 1";
  if (c->robot != GX_ROBOT_POINT && c->robot != GX_ROBOT_ART && c->robot != GX_ROBOT_TREE) return "unknown robot";
  const int art = c->robot != GX_ROBOT_POINT;
  const int tree = c->robot == GX_ROBOT_TREE;
  if (art && (c->art_nu < 1 || c->art_nv < 3 || c->art_nu > 64 || c->art_nv > 64)) return "art_nu / art_nv out of range";
  p->nu = art ? c->art_nu : 3; p->nv = art ? c->art_nv : 3;
  p->nq = art ? (c->art_nq > 0 ? c->art_nq : c->art_nv) : 3;
  if (c->hazards_num < 1 || c->hazards_num > GX_MAX_HAZARDS)
    return "hazards_num out of range (the reference crashes for 0: engine.py:370-384)";
  if (c->lidar_num_bins < 1 || c->lidar_num_bins > GX_MAX_LIDAR_BINS)
    return "lidar_num_bins out of range";
  if (c->physics_steps_per_control_step < 1) return "physics_steps_per_control_step < 1";
  if (c->layout_tries < 1 || c->layout_tries > 255) return "layout_tries out of range";
  if (c->pool_size < 0) return "pool_size < 0";
  p->n = c->env_num; p->nh = c->hazards_num; p->nb = c->lidar_num_bins;
  p->substeps = c->physics_steps_per_control_step;
  p->alias = c->lidar_alias != 0;
  p->use_exp = !(c->lidar_max_dist > 0.0);
  p->tries = c->layout_tries;
  int off = 0;
  p->off_ctrl = p->off_compass = p->off_glidar = p->off_hlidar = p->off_qpos = p->off_qvel = -1;
  p->off_acc = p->off_vel = -1;
  /* sorted key order (engine.py:775): acc ctrl goal_compass goal_lidar hazards_lidar qpos qvel vel */
  if (c->obs_mask & GX_OBS_ACC) { p->off_acc = off; off += 2; }
  if (c->obs_mask & GX_OBS_CTRL) { p->off_ctrl = off; off += p->nu; }
  if (c->obs_mask & GX_OBS_GOAL_COMPASS) { p->off_compass = off; off += 2; }
  if (c->obs_mask & GX_OBS_GOAL_LIDAR) { p->off_glidar = off; off += p->nb; }
  if (c->obs_mask & GX_OBS_HAZARDS_LIDAR) { p->off_hlidar = off; off += p->nb; }
  if (c->obs_mask & GX_OBS_QPOS) { p->off_qpos = off; off += p->nq; }
  if (c->obs_mask & GX_OBS_QVEL) { p->off_qvel = off; off += p->nv; }
  if (c->obs_mask & GX_OBS_VEL) { p->off_vel = off; off += 2; }
  if (off == 0) return "empty observation";
  p->obs_dim = off;
  p->lay_dim = 2 * (2 + p->nh);
  p->state_dim = tree ? p->nq + 2 * p->nv + 5 + 2 * p->nh + 2 + 2 * p->nh : (art ? 3 * p->nv + 5 + 2 * p->nh : 19 + 2 * p->nh);
  p->aux_dim = 3 + p->nh;
  p->pool_size = c->pool_size > 0 ? c->pool_size : gx_auto_pool_size(c->env_num);
  /* self.dt = mj_model.opt.timestep (a double) * physics_steps_per_control_step (engine.py:235) */
  p->ctrl_dt = (float)(c->opt_timestep * (double)c->physics_steps_per_control_step);
  p->num_steps = (float)c->num_steps;
  p->goal_size = (float)c->goal_size;
  p->hazards_size = (float)c->hazards_size;
  p->reward_distance = (float)c->reward_distance;
  p->lidar_max_dist = (float)c->lidar_max_dist;
  p->neg_exp_gain = (float)(-c->lidar_exp_gain);
  p->bin_size = (float)((3.141592653589793 * 2) / (double)c->lidar_num_bins);
  const gx_point_model* m = &c->point;
  p->dt = m->timestep; p->body_z = m->body_z; p->mass = m->mass;
  p->mc = m->mass * m->com_x;
  p->inertia = m->inertia_zz;
  for (int i = 0; i < 3; ++i) { p->damp[i] = m->damping[i]; p->gear[i] = m->gear[i]; }
  p->gain = m->gainprm0; p->bias2 = m->biasprm2;
  p->ctrllimited = m->ctrllimited; p->forcelimited = m->forcelimited;
  p->ctrl_lo = m->ctrlrange[0]; p->ctrl_hi = m->ctrlrange[1];
  p->force_lo = m->forcerange[0]; p->force_hi = m->forcerange[1];
  { float t;
    t = m->timestep * m->damping[0]; p->mh0 = m->mass + t;
    t = m->timestep * m->damping[1]; p->mh1 = m->mass + t;
    t = m->timestep * m->damping[2]; p->mh2 = m->inertia_zz + t; }
  const double keep[3] = {c->goal_keepout, c->hazards_keepout, c->robot_keepout};
  for (int a = 0; a < 3; ++a) {
    p->lo[a][0] = (float)(c->placements_extents[0] + keep[a]);
    p->lo[a][1] = (float)(c->placements_extents[1] + keep[a]);
    p->hi[a][0] = (float)(c->placements_extents[2] - keep[a]);
    p->hi[a][1] = (float)(c->placements_extents[3] - keep[a]);
    if (!(p->lo[a][0] <= p->hi[a][0]) || !(p->lo[a][1] <= p->hi[a][1]))
      return "placements_extents smaller than a keepout (engine.py:610)";
    for (int b = 0; b < 3; ++b) p->keepsum[a][b] = (float)(keep[a] + c->placements_margin + keep[b]);
  }
  p->min_robot_goal = (float)c->goal_min_robot_dist;
  p->seed_lo = (uint32_t)c->seed; p->seed_hi = (uint32_t)(c->seed >> 32);
  p->env_id_offset = c->env_id_offset;
  return 0;
}

/* Engine.DEFAULT (engine.py:98-204) merged with Goal_Point_8Hazards
   (safe_rl_env_config.py:59-82); model constants of SURVEY.md Appendix C,
   evaluated in double from point.xml:5, 13, 16-20, 37-39. */
static inline void gx_fill_default_config(gx_config* c) {
  const double pi = 3.141592653589793;
  const double m1 = 4.0 / 3.0 * pi * 0.1 * 0.1 * 0.1; /* sphere r=.1, density 1 */
  const double m2 = 8.0 * 0.05 * 0.05 * 0.05;         /* box half .05 at x=.1   */
  const double m = m1 + m2;
  const double cx = m2 * 0.1 / m;
  /* about the joint anchor (body origin = sphere centre) */
  const double izz = 0.4 * m1 * 0.01 + m2 * (0.05 * 0.05 + 0.05 * 0.05) / 3.0 + m2 * 0.01;
  c->abi_version = GX_ABI_VERSION;
  c->env_num = 1; c->num_steps = 1000; c->robot = GX_ROBOT_POINT;
  c->seed = 0; c->env_id_offset = 0;
  c->placements_extents[0] = -2; c->placements_extents[1] = -2;
  c->placements_extents[2] = 2; c->placements_extents[3] = 2;
  c->placements_margin = 0.0;
  c->robot_keepout = 0.4; c->goal_keepout = 0.5; c->hazards_keepout = 0.4;
  c->goal_size = 0.5; c->hazards_size = 0.3; c->reward_distance = 1.0;
  c->lidar_max_dist = 0.0; c->lidar_exp_gain = 1.0; c->goal_min_robot_dist = 3.0;
  c->opt_timestep = 0.02;
  c->hazards_num = 8; c->lidar_num_bins = 16; c->lidar_alias = 1;
  c->physics_steps_per_control_step = 1; c->layout_tries = 10; c->pool_size = 0;
  c->obs_mask = GX_OBS_DEFAULT; c->art_nu = 0; c->art_nv = 0; c->art_nq = 0;
  gx_point_model* pm = &c->point;
  pm->timestep = 0.02f; pm->body_z = 0.1f;
  pm->mass = (float)m; pm->com_x = (float)cx; pm->inertia_zz = (float)izz;
  pm->damping[0] = 0.01f; pm->damping[1] = 0.01f; pm->damping[2] = 0.005f;
  pm->gear[0] = pm->gear[1] = pm->gear[2] = 0.3f;
  pm->gainprm0 = 1.0f; pm->biasprm2 = -1.0f; /* A1: velocity-servo defaults win */
  pm->ctrllimited = 1; pm->ctrlrange[0] = -1.0f; pm->ctrlrange[1] = 1.0f;
  pm->forcelimited = 1; pm->forcerange[0] = -0.05f; pm->forcerange[1] = 0.05f;
}

#endif /* GX_PARAMS_H_ */
