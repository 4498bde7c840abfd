/* guardx.h -- C ABI of libguardx.so: the batched safe_rl_envs environment step
 * (guardX `Engine.reset / step / reset_done`) as hand-written sm_100a CUDA.
 *
 * This is the drop-in boundary.  The reference has no FFI: its
 * boundary is the Python class `Engine` in
 * safe_rl_envs/safe_rl_envs/envs/engine.py, which hands torch tensors to
 * jax.jit(jax.vmap(...)) closures through DLPack (engine_utils.py:22-27, 55-64).
 * Each entry point below names the reference method it replaces; the Python
 * class `guardx_b200.Engine` binds them with ctypes and `tensor.data_ptr()`
 * (INTEGRATION.md shows the stub a maintainer would add to engine.py).
 *
 * Conventions
 *  - plain C, no torch types; every array pointer is a DEVICE pointer unless the
 *    parameter name starts with `h_` (pinned or pageable host memory);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *    all work is enqueued on it and nothing synchronises unless stated;
 *  - every function returns 0 on success, a negative GX_E* code otherwise, and
 *    leaves a message for gx_last_error() (thread-local);
 *  - a handle owns one batch of `env_num` environments on one GPU and is not
 *    thread-safe (same as the reference object).
 *  - all float arrays are fp32, row-major, batch dimension first.
 */
#ifndef GUARDX_H_
#define GUARDX_H_

#include <stdint.h>

#include "guardx_art.h"
#include "guardx_tree.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GX_ABI_VERSION 3

#define GX_MAX_HAZARDS 32    /* hazards_num upper bound of this build */
#define GX_MAX_LIDAR_BINS 64 /* lidar_num_bins upper bound of this build */

enum {
  GX_OK = 0,
  GX_EINVAL = -1,   /* bad argument / unsupported configuration */
  GX_ECUDA = -2,    /* CUDA runtime error (message has the cudaError string) */
  GX_ENOMEM = -3,
  GX_ESTATE = -4    /* call order violated (e.g. step before reset) */
};

/* robot_base (engine.py:113).  Only robots whose dynamics this build implements. */
enum {
  GX_ROBOT_POINT = 0, /* xmls/point.xml: closed-form 3-DoF dynamics (gx_point_model) */
  GX_ROBOT_ART = 1,   /* articulated planar-root robot (ant / walker / swimmer): gx_create_art + guardx_art.h */
  GX_ROBOT_TREE = 2   /* general kinematic tree, free-joint root, capsule contacts (doggo): gx_create_tree + guardx_tree.h */
};

/* observation components, flat order = sorted key order (engine.py:773-777) */
enum {
  GX_OBS_ACC = 1 << 0,           /* observe_acc   (engine.py:127) */
  GX_OBS_CTRL = 1 << 1,          /* observe_ctrl  (engine.py:128) */
  GX_OBS_GOAL_COMPASS = 1 << 2,  /* observe_goal_comp  (engine.py:120) */
  GX_OBS_GOAL_LIDAR = 1 << 3,    /* observe_goal_lidar (engine.py:119) */
  GX_OBS_HAZARDS_LIDAR = 1 << 4, /* observe_hazards    (engine.py:121) */
  GX_OBS_QPOS = 1 << 5,          /* observe_qpos  (engine.py:123) */
  GX_OBS_QVEL = 1 << 6,          /* observe_qvel  (engine.py:124) */
  GX_OBS_VEL = 1 << 7            /* observe_vel   (engine.py:126) */
};
#define GX_OBS_DEFAULT                                                              \
  (GX_OBS_CTRL | GX_OBS_GOAL_COMPASS | GX_OBS_GOAL_LIDAR | GX_OBS_HAZARDS_LIDAR | \
   GX_OBS_QPOS | GX_OBS_QVEL)

/* Compiled constants of the MuJoCo model that World.build() emits for the Point
 * robot (world.py:104-333 + xmls/point.xml).  The reference obtains them from
 * mujoco.MjModel.from_xml_string (engine.py:228); no MuJoCo compiler exists in
 * this image, so gx_default_config() fills the hand-derived values of SURVEY.md
 * Appendix C.  They are data, not code: a maintainer with MuJoCo can overwrite
 * them from the real MjModel before gx_create(). */
typedef struct gx_point_model {
  float timestep;      /* opt.timestep                       point.xml:3   0.02 */
  float body_z;        /* body 'robot' pos z                 point.xml:13  0.1  */
  float mass;          /* body_mass[robot] (sphere + box, density 1)            */
  float com_x;         /* body_ipos[robot].x                                    */
  float inertia_zz;    /* I_zz about the hinge axis THROUGH THE JOINT ANCHOR    */
  float damping[3];    /* dof_damping robot_x, robot_y, robot_z point.xml:16-18 */
  float gear[3];       /* actuator_gear[:,0]                 point.xml:37-39    */
  float gainprm0;      /* actuator_gainprm[:,0]  (1: default class, A1)         */
  float biasprm2;      /* actuator_biasprm[:,2]  (-1 velocity servo; 0 = motor) */
  int32_t ctrllimited; /* actuator_ctrllimited                                  */
  float ctrlrange[2];
  int32_t forcelimited;
  float forcerange[2];
} gx_point_model;

/* The subset of Engine.DEFAULT (engine.py:98-204) that changes the transition
 * function.  Keys the JAX Engine accepts but ignores (sensors_*, constrain_*,
 * goal_mode, reward_goal, lidar_type, render_*, device_id, ...) are parsed and
 * dropped on the Python side.
 *
 * Python-float keys are carried as DOUBLES, exactly as the reference holds them:
 * the library converts to fp32 at the points where JAX's weak-type promotion
 * does (e.g. `other_keepout + placements_margin + keepout` at engine.py:554 is
 * summed in double and only then compared with an fp32 distance; `(2*pi)/bins`
 * at engine.py:880 is a double divided once and then rounded). */
typedef struct gx_config {
  int32_t abi_version;            /* GX_ABI_VERSION */
  int32_t env_num;                /* engine.py:101 */
  int32_t num_steps;              /* engine.py:99  */
  int32_t robot;                  /* GX_ROBOT_*    */
  uint64_t seed;                  /* '_seed' engine.py:203 */
  uint64_t env_id_offset;         /* global id of local env 0 (multi-GPU sharding) */
  double placements_extents[4];   /* xmin ymin xmax ymax, engine.py:103 */
  double placements_margin;       /* engine.py:104 */
  double robot_keepout;           /* engine.py:112 */
  double goal_keepout;            /* engine.py:166 */
  double hazards_keepout;         /* engine.py:198 */
  double goal_size;               /* engine.py:167 */
  double hazards_size;            /* engine.py:199 */
  double reward_distance;         /* engine.py:174 */
  double lidar_max_dist;          /* engine.py:150; <= 0 means None (exponential lidar) */
  double lidar_exp_gain;          /* engine.py:151 */
  double goal_min_robot_dist;     /* the literal 3.0 at engine.py:571 */
  double opt_timestep;            /* mj_model.opt.timestep as a double (point.xml:3); Engine.dt = this * substeps, engine.py:235 */
  int32_t hazards_num;            /* engine.py:195, 1..GX_MAX_HAZARDS */
  int32_t lidar_num_bins;         /* engine.py:148, 1..GX_MAX_LIDAR_BINS */
  int32_t lidar_alias;            /* engine.py:153 */
  int32_t physics_steps_per_control_step; /* engine.py:202 */
  int32_t layout_tries;           /* the literal range(10) at engine.py:562 */
  int32_t pool_size;              /* layout pool entries; 0 = auto (replaces the 1e6-candidate pool, engine.py:263) */
  uint32_t obs_mask;              /* GX_OBS_* */
  int32_t art_nu;                 /* GX_ROBOT_ART: robot.nu (width of 'ctrl' in the observation, engine.py:402-403) */
  int32_t art_nv;                 /* GX_ROBOT_ART / TREE: robot.nv (width of 'qvel', engine.py:400-401) */
  int32_t art_nq;                 /* GX_ROBOT_TREE: robot.nq (width of 'qpos'; = art_nv for GX_ROBOT_ART, 0 = same) */
  gx_point_model point;
} gx_config;

typedef struct gx_env gx_env;

/* ---- library ---------------------------------------------------------- */
int gx_abi_version(void);
int gx_config_size(void); /* sizeof(gx_config), for binding self-checks */
const char* gx_last_error(void);
/* Engine.DEFAULT + Goal_Point_8Hazards (safe_rl_env_config.py:59-82) + Appendix C. */
void gx_default_config(gx_config* cfg);

/* ---- lifetime: Engine.__init__ (engine.py:207-316) ---------------------- */
int gx_create(const gx_config* h_cfg, int device, gx_env** out);
/* Same for cfg->robot == GX_ROBOT_ART: h_model is the compiled robot (guardx_art.h), what
 * mujoco.MjModel.from_xml_string + mjx.put_model produce in the reference (engine.py:228-231).
 * Every entry point below then works on that robot: action [env_num, nu], qacc
 * [env_num, nv], state rows as documented at gx_get_state, aux rows as for the Point. */
int gx_create_art(const gx_config* h_cfg, const GxArtModel* h_model, int device, gx_env** out);
/* Same for cfg->robot == GX_ROBOT_TREE (Doggo): h_model is the compiled general-tree robot (guardx_tree.h).
 * First version of this path: one thread per env with a global-memory scratch buffer; state rows
 * qpos[nq] qvel[nv] qacc_warmstart[nv] prev_dist prev_done steps goal[2] hazards[2*nh] world_vel[2+2*nh]
 * (the goal / hazard slides move when the robot's defaults class gives them a stiffness). */
int gx_create_tree(const gx_config* h_cfg, const GxTreeModel* h_model, int device, gx_env** out);
int gx_destroy(gx_env* env);

/* ---- spaces: build_observation_space (engine.py:386-418), action space
 *      (engine.py:291-297) ------------------------------------------------ */
int gx_obs_dim(const gx_env* env);
int gx_act_dim(const gx_env* env);
int gx_env_num(const gx_env* env);
int gx_pool_size(const gx_env* env);
/* offset of a GX_OBS_* component inside the flat observation, -1 if absent */
int gx_obs_offset(const gx_env* env, uint32_t component);
/* h_low/h_high: act_dim floats each (+-inf when not ctrllimited) */
int gx_action_bounds(const gx_env* env, float* h_low, float* h_high);

/* ---- Engine.reset (engine.py:454-467): regenerate the layout pool
 *      (reset_layout :433-444), draw one layout per env (get_layout :446-452),
 *      mjx_reset (:644-657).  obs: [env_num, obs_dim]. ---------------------- */
int gx_reset(gx_env* env, float* obs, void* stream);

/* Same, but with caller-supplied layouts instead of the pool: layout is
 * [env_num, 2*(2+hazards_num)] = goal xy, robot xy, hazard0 xy, ...  This is
 * Engine._reset(layout) (engine.py:269, 460) and is how parity fixtures inject
 * the reference's own layouts. */
int gx_reset_with_layout(gx_env* env, const float* layout, float* obs, void* stream);

/* ---- Engine.step (engine.py:469-495) + mjx_step (:659-700).
 *      action [env_num, act_dim]; obs [env_num, obs_dim]; reward, done, cost
 *      [env_num]; done is fp32 in {0,1}.  Optional (NULL to skip):
 *      qacc [env_num, 3]   info['obs']['qacc']  (engine.py:763-764)
 *      aux  [env_num, gx_aux_dim()] int32 = flags (GX_AUX_*), number of hazards
 *           in violation (cost terms > 0), goal lidar bin, hazard lidar bins
 *           [hazards_num] -- the integer outputs BASELINE.json asks to be
 *           bit-exact (bins as the reference computes them, i.e. possibly ==
 *           lidar_num_bins at the fp32 wrap edge). ----------------------- */
int gx_step(gx_env* env, const float* action, float* obs, float* reward, float* done,
            float* cost, float* qacc, int32_t* aux, void* stream);

/* Fused gx_step + in-place gx_reset_done in ONE launch ("termination plus
 * auto-reset", BASELINE.json): reward/done/cost are the step's; the obs row of
 * an env that finished is its reset observation -- bit-identical to calling
 * gx_step and then gx_reset_done(obs, obs, NULL).  The reference needs two jitted
 * calls and a full extra physics step for this (engine.py:497-505, 719-724). */
int gx_step_autoreset(gx_env* env, const float* action, float* obs, float* reward, float* done,
                      float* cost, void* stream);

int gx_aux_dim(const gx_env* env); /* 3 + hazards_num */
enum {
  GX_AUX_GUARD = 1,   /* NaN/Inf in obs -> reward 0, done 1 (engine.py:696-699) */
  GX_AUX_TIMEOUT = 2, /* steps > num_steps (engine.py:492) */
  GX_AUX_GOAL = 4,    /* dist_goal < goal_size (engine.py:798) */
  GX_AUX_JUMP = 8     /* |d_dist| > 1 (engine.py:800-801) */
};

/* ---- Engine.reset_done (engine.py:497-505) + mjx_reset_done (:702-731).
 *      obs_out rows of done envs get the reset observation, the other rows are
 *      copied from obs_in (the previous step's observation, Engine._obs).
 *      obs_in == obs_out is allowed (only done rows are written).
 *      layout: NULL = draw from the pool, else [env_num, 2*(2+hazards_num)]
 *      (rows of envs that are not done are ignored). ---------------------- */
int gx_reset_done(gx_env* env, const float* obs_in, float* obs_out, const float* layout,
                  void* stream);

/* ---- state injection / checkpoint (no reference counterpart: the reference
 *      state is the mjx.Data pytree held in Engine._data).  One row per env:
 *      qpos[3] qvel[3] lag_cos lag_sin prev_dist prev_done steps goal[2]
 *      hazards[2*hazards_num], then the history that only observe_vel /
 *      observe_acc read (engine.py:902-929): xpos_xy of Engine._data, xpos_xy of
 *      Engine._last_data, Engine._last_done, history depth (0,1,2 = how many of
 *      _last_done/_last_last_done are not None).  gx_state_dim() floats per row.
 *      GX_ROBOT_ART rows: qpos[nv] qvel[nv] qacc_warmstart[nv] prev_dist prev_done steps
 *      goal[2] hazards[2*hazards_num]. */
int gx_state_dim(const gx_env* env);
int gx_get_state(gx_env* env, float* state, void* stream);
int gx_set_state(gx_env* env, const float* state, void* stream);

/* ---- the two counters that key every random draw: t_global (env steps taken)
 *      and epoch (layout pools generated).  They live in DEVICE memory and are
 *      advanced by the kernels themselves, so gx_reset / gx_step* /
 *      gx_reset_done / gx_rollout never bake a host-side count into a launch and
 *      can be captured in a CUDA graph and replayed.  These two calls copy them
 *      and synchronise `stream` (checkpointing, test alignment). ------------- */
int gx_set_counters(gx_env* env, uint64_t t_global, uint64_t epoch, void* stream);
int gx_get_counters(gx_env* env, uint64_t* h_t_global, uint64_t* h_epoch, void* stream);

/* ---- layout pool (reset_layout, engine.py:433-444) ------------------------
 * pool: [pool_size, 2*(2+hazards_num)] device floats; tries: [pool_size] int32,
 * the number of candidates each slot consumed (acceptance-rate evidence). */
int gx_pool_read(gx_env* env, float* pool, int32_t* tries, void* stream);

/* ---- synthetic random-action rollout: T control steps of
 *        a_t ~ U(-1,1)^act_dim (Philox, keyed seed/global env id/step counter,
 *        as trpo/myTest.py:31 drives the reference), step, and -- as the PPO
 *        loop does (ppo_train.py:387-417) -- reset_done after every step on
 *        which any env of the batch is done.
 *      All outputs optional (NULL to skip):  obs [T, env_num, obs_dim] is the
 *      observation the NEXT action is chosen from (post reset_done), act
 *      [T, env_num, act_dim], reward/cost/done [T, env_num]. ---------------- */
int gx_rollout(gx_env* env, int32_t T, float* obs, float* act, float* reward, float* cost,
               float* done, void* stream);

/* ---- episode statistics accumulated on device by gx_step/gx_rollout since
 *      the last gx_stats_reset: out[8] doubles = n_episodes, sum_ret, sum_cost,
 *      sum_len, sum_ret^2, sum_cost^2, total_cost, total_env_steps.  Replaces the
 *      host-side accumulation of ppo_train.py:370-376 and feeds the NCCL
 *      all-reduce that replaces mpi_statistics_scalar (mpi_tools.py:70-92). */
int gx_stats_read(gx_env* env, double* out8, void* stream);
/* zeroes the sums and (first call) switches accumulation on; until then
 * gx_step / gx_rollout do not touch the statistics buffers at all */
int gx_stats_reset(gx_env* env, void* stream);

/* ---- host-buffer convenience (the end-to-end path of bench.py): H2D of the
 *      action, step, optional reset_done, D2H of obs/reward/done/cost, and a
 *      stream synchronise.  When all five buffers are pinned AND mapped (cudaHostAlloc,
 *      cudaHostRegister, torch pin_memory) the copies are skipped: one launch reads
 *      the actions from and writes the outputs to host memory directly. -------- */
int gx_step_host(gx_env* env, const float* h_action, float* h_obs, float* h_reward,
                 float* h_done, float* h_cost, int32_t auto_reset_done, void* stream);

/* ---- host-buffer rollout: T x (gx_step_host with auto_reset_done) for actions
 *      that are known up front (the random-action rollouts of trpo/myTest.py:31
 *      and of BASELINE.json's configs), pipelined: while chunk c steps on the
 *      device, chunk c-1's outputs drain to the host and chunk c+1's actions go
 *      up.  h_action [T, env_num, act_dim]; h_obs [T, env_num, obs_dim];
 *      h_reward/h_done/h_cost [T, env_num]; all pinned host memory.  Returns
 *      after everything has landed in host memory.  Bit-identical to T calls of
 *      gx_step_host. ------------------------------------------------------- */
int gx_rollout_host(gx_env* env, int32_t T, const float* h_action, float* h_obs, float* h_reward,
                    float* h_done, float* h_cost, void* stream);
/* Which route gx_rollout_host settled on: 0 = DMA ring (copy engine), 1 = zero-copy (kernels store straight into the
 * mapped host buffers), -1 = still measuring (the first four calls time both).  rates2 (host, may be NULL) receives
 * the env-steps/s measured for the two routes.  GX_ROLLOUT_HOST_ROUTE=dma|zc in the environment forces one. */
int gx_rollout_host_route(const gx_env* e, double* rates2);

/* ---- rollout-side op: GAE-lambda advantages and rewards-to-go
 *      over a time-major rollout, replacing PPOBufferX.finish_path
 *      (safe_rl_libX/ppo/ppo_train.py:61-113: a per-env Python loop over
 *      scipy.signal.lfilter with .cpu() round trips) by one streaming kernel.
 *      reward, value, done: [T, n] (done[t] = 1 closes a path AFTER step t; the
 *      reference bootstraps every closed path and the epoch end with 0,
 *      ppo_train.py:388-409, so V(s_T) = last_value or 0 when NULL).
 *        delta_t = (r_t + gamma * V_{t+1} * (1 - done_t)) - V_t        (fp32)
 *        adv_t   = delta_t + gamma*lam * adv_{t+1} * (1 - done_t)      (fp64 scan, as lfilter)
 *        ret_t   = r_t + gamma * ret_{t+1} * (1 - done_t)              (fp64 scan)
 *      adv, ret: [T, n] fp32.  No env handle: works on any buffers. ---------- */
int gx_gae(const float* reward, const float* value, const float* done, const float* last_value,
           int32_t T, int32_t n, double gamma, double lam, float* adv, float* ret, void* stream);

/* ---- rollout-side op: done-compaction.  idx[0 .. *count) = ascending indices of
 *      the envs with done > 0 (idx has room for n entries), *count = how many -- replaces the
 *      host-side np.where(done == 1) of ppo_train.py:403-411 / scpo/scpo.py:722.  No env handle. */
int gx_compact_done(const float* done, int32_t n, int32_t* idx, int32_t* count, void* stream);

/* ---- the policy half of a PPO / PPO-Lagrangian rollout step in one launch ------------------
 * Replaces `ac.step(o)` + the rollout-buffer stores of safe_rl_libX/ppo/ppo_train.py:364-379 (actor-critic of
 * ppo_core.py:116-160: three fp32 tanh MLPs obs -> 64 -> 64 -> {act_dim | 1 | 1}, state-independent log-std,
 * Gaussian sample + log-probability; the third net is the cost critic of TRPO-Lag).
 *   weights   device, gx_policy_weight_floats(e) floats, packed as csrc/gx_policy.cuh documents
 *   obs       device [env_num, obs_dim] (what the env kernels wrote)
 *   act       device [env_num, act_dim]  sampled action   (feed it to gx_step / gx_step_autoreset)
 *   val, cval, logp   device [env_num]
 *   mu_out    device [env_num, act_dim] or NULL: the actor mean (tests)
 * The exploration noise is a pure function of (noise_seed, config seed, global env id, the env's step count): the call
 * can be captured in a CUDA graph and replayed.  Call before the env step of the same control step. */
int64_t gx_policy_weight_floats(const gx_env* e);
int gx_policy_step(gx_env* e, const float* weights, const float* obs, uint64_t noise_seed, float* act, float* val,
                   float* cval, float* logp, float* mu_out, void* stream);

/* ---- a whole PPO / PPO-Lagrangian rollout in ONE persistent launch ------------------------------
 * Replaces the T-step loop of safe_rl_libX/ppo/ppo_train.py:364-417 ({ a, v, logp = ac.step(o); o, r, d, info =
 * env.step(a); buf.store(...) } with the env's auto-reset) for the default Point task shape (obs 43, act 2:
 * gx_policy_rollout_supported(e) != 0): every block owns 32 envs, keeps their state in registers and the three
 * networks in shared memory, and alternates policy and env phases T times (csrc/gx_policy.cuh).
 *   weights   device, as for gx_policy_step, 16-byte aligned
 *   obs       device [T + 1, env_num, 43]: row 0 must hold the current observation (gx_reset's, or the last row of
 *             the previous rollout); rows 1..T are written (row t + 1 = observation after step t, reset rows for
 *             envs that finished in step t)
 *   act       device [T, env_num, 2];  reward, done, cost, val, cval, logp: device [T, env_num]
 * Results are bit-identical to T x { gx_policy_step; gx_step_autoreset } with the same noise_seed.  Episode statistics
 * (gx_stats_*) are accumulated when switched on.  Graph-capturable. */
int gx_policy_rollout_supported(const gx_env* e);
int gx_policy_rollout(gx_env* e, const float* weights, int32_t T, uint64_t noise_seed, float* obs, float* act, float* reward,
                      float* done, float* cost, float* val, float* cval, float* logp, void* stream);

/* ---- diagnostics: evaluate the shared fp32 math (csrc/gx_math.h) on the
 *      device, so tests can prove host/device bit-equality of the header.
 *      fn: 0 sincos(a)->out,out2   1 atan2(a,b)->out   2 exp(a)->out
 *          3 max(a,b)->out   4 min(a,b)->out.
 *      a, b, out, out2: device arrays of n floats (b/out2 may be NULL). ------ */
int gx_debug_math(int fn, const float* a, const float* b, float* out, float* out2, int n,
                  void* stream);
/* diagnostics: the SM clock (MHz) a kernel sees on `stream`'s device, measured over
 * `spin_cycles` cycles of one spinning thread against %globaltimer; *out_mhz is a DEVICE float */
int gx_debug_sm_clock(float* out_mhz, int32_t spin_cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GUARDX_H_ */
