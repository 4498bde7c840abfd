/* guardx_task.h -- C ABI of the config-3 extension: Button / Push tasks, the Car robot, vases, pillars and gremlins
 * (BASELINE.json configs[2]).
 *
 * NO REFERENCE, PARITY UNPINNED BY CONSTRUCTION.  The reference's JAX Engine implements only the goal task with hazards
 * (safe_rl_envs/safe_rl_envs/envs/engine.py:335-384): there is no button / push task, no vase, pillar or gremlin and no
 * Car robot anywhere under /root/reference (world.py:55 names an xmls/car.xml that does not exist; engine.py:32-65 keeps
 * only the colour / group constants of these objects; safe_rl_env_config.py's Push_* entries carry keys Engine.parse
 * rejects, engine.py:326-328).  What is implemented here is the SPECIFICATION written down in DESIGN.md section 8, which
 * follows upstream Safety-Gym's semantics where they are recalled and the JAX Engine's conventions (lagged pose for
 * observation / reward / cost, episode ends at the goal, dense hazard cost, no goal bonus, NaN guard, timeout) everywhere
 * else.  The checker is oracle/task_oracle.py, an independent numpy restatement of that specification.
 *
 * One handle = one batch of envs on one device; all pointers are DEVICE pointers; every entry point enqueues on
 * `stream` and returns without synchronising (gxt_create and gxt_destroy excepted); 0 = ok, negative = error with the
 * message in gx_last_error() (include/guardx.h).
 */
#ifndef GUARDX_TASK_H_
#define GUARDX_TASK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GXT_ABI_VERSION 1
#define GXT_MAX_PER_CLASS 16
#define GXT_MAX_OBJECTS 64

enum { GXT_TASK_GOAL = 0, GXT_TASK_BUTTON = 1, GXT_TASK_PUSH = 2 };
enum { GXT_ROBOT_POINT = 0, GXT_ROBOT_CAR = 1 };
/* object classes, in placement order (the robot is placed last, as engine.py:538-542 does) */
enum { GXT_GOAL = 0, GXT_BOX = 1, GXT_BUTTON = 2, GXT_HAZARD = 3, GXT_VASE = 4, GXT_PILLAR = 5, GXT_GREMLIN = 6, GXT_ROBOT = 7, GXT_NCLASS = 8 };

/* observation components, flat layout in sorted key order (engine.py:773-777):
 * box_compass box_lidar buttons_lidar ctrl goal_compass goal_lidar gremlins_lidar hazards_lidar pillars_lidar qpos qvel vases_lidar */
enum {
  GXT_OBS_BOX_COMPASS = 1 << 0, GXT_OBS_BOX_LIDAR = 1 << 1, GXT_OBS_BUTTONS_LIDAR = 1 << 2, GXT_OBS_CTRL = 1 << 3,
  GXT_OBS_GOAL_COMPASS = 1 << 4, GXT_OBS_GOAL_LIDAR = 1 << 5, GXT_OBS_GREMLINS_LIDAR = 1 << 6, GXT_OBS_HAZARDS_LIDAR = 1 << 7,
  GXT_OBS_PILLARS_LIDAR = 1 << 8, GXT_OBS_QPOS = 1 << 9, GXT_OBS_QVEL = 1 << 10, GXT_OBS_VASES_LIDAR = 1 << 11
};

typedef struct gxt_config {
  int32_t abi_version, env_num, num_steps, task, robot;
  uint64_t seed, env_id_offset;
  int32_t count[GXT_NCLASS];      /* objects per class (goal 1; box 0/1; robot 1) */
  double size[GXT_NCLASS];        /* radius of the class's disc (robot: contact radius) */
  double keepout[GXT_NCLASS];     /* placement keep-out radius */
  double placements_extents[4], placements_margin, goal_min_robot_dist;
  int32_t layout_tries, pool_size;
  /* rewards / costs */
  double reward_distance, reward_box_dist, reward_box_goal;
  double hazards_cost_scale;      /* 1: dense hazard cost as the JAX Engine (engine.py:804-811) */
  double vases_contact_cost, pillars_cost, buttons_cost;
  double gremlins_contact_cost, gremlins_dist_threshold, gremlins_dist_cost, gremlins_travel;
  /* lidar */
  int32_t lidar_num_bins, lidar_alias;
  double lidar_max_dist, lidar_exp_gain;
  uint32_t obs_mask;
  /* Point robot: the compiled constants of xmls/point.xml (same meaning as gx_point_model in guardx.h) */
  float pt_timestep, pt_mass, pt_com_x, pt_inertia_zz, pt_damping[3], pt_gear[3], pt_gain, pt_bias2, pt_ctrl[2], pt_force[2];
  /* Car robot (planar differential drive): chassis mass, yaw inertia, wheel radius, half track, motor gear,
   * linear / angular viscous damping, timestep */
  float car_mass, car_inertia, car_wheel_r, car_half_track, car_gear, car_damp_lin, car_damp_ang, car_timestep;
} gxt_config;

typedef struct gxt_env gxt_env;

int gxt_config_size(void);                  /* sizeof(gxt_config) as the library sees it */
void gxt_default_config(gxt_config* cfg);   /* Button task, Car, 4 buttons, 4 hazards, 4 vases, 2 pillars, 4 gremlins */
int gxt_create(const gxt_config* host_cfg, int device, gxt_env** out);
int gxt_destroy(gxt_env* e);
int gxt_obs_dim(const gxt_env* e);
int gxt_act_dim(const gxt_env* e);
int gxt_state_dim(const gxt_env* e);
int gxt_layout_dim(const gxt_env* e);       /* floats per layout row: xy per object in placement order, robot xy, gremlin phases */
int gxt_pool_size(const gxt_env* e);
int gxt_obs_offset(const gxt_env* e, uint32_t component);

/* Engine.reset: regenerate the layout pool (rejection sampling, Philox), place every env, observation [n, obs_dim].
 * layout (optional, [n, layout_dim]) replaces the pool draw. */
int gxt_reset(gxt_env* e, const float* layout, float* obs, void* stream);
/* Engine.step (+ Engine.reset_done fused when autoreset != 0).  action [n, act_dim]; outputs obs [n, obs_dim],
 * reward / done / cost [n]; aux (optional int32 [n, 4]): done-reason flags (GX_AUX_*), hazard violations,
 * contact count (vases + pillars + gremlins + wrong buttons), goal-lidar bin. */
int gxt_step(gxt_env* e, const float* action, float* obs, float* reward, float* done, float* cost, int32_t* aux,
             int32_t autoreset, void* stream);
/* Engine.reset_done: re-place the envs whose done flag is set; obs_out = obs_in with their rows replaced. */
int gxt_reset_done(gxt_env* e, const float* obs_in, float* obs_out, const float* layout, void* stream);
/* state rows [n, state_dim] (parity injection / checkpoints); layout documented in csrc/gx_task.cuh */
int gxt_get_state(gxt_env* e, float* state, void* stream);
int gxt_set_state(gxt_env* e, const float* state, void* stream);
int gxt_pool_read(gxt_env* e, float* pool, int32_t* tries, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GUARDX_TASK_H_ */
