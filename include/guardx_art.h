/* guardx_art.h -- compiled model of an articulated planar-root robot (Ant, Walker,
 * Swimmer) as the CUDA kernels of libguardx.so consume it: what mjx.put_model
 * (engine.py:231) uploads in the reference, restricted to the robot (the goal/hazard
 * bodies World.build adds, world.py:310-326, are static and decoupled).
 *
 * Topology: STAR -- one root body with three planar root joints and up to
 * GX_ART_MAXLEG serial chains, each with at most one sphere that can touch the floor
 * plane z = 0.  guardx_b200/art_model.py fills this struct from the model compiler
 * (guardx_b200/mjcf.py); a maintainer with MuJoCo can fill it from a real MjModel.
 * All values fp32 (mjx casts the double-precision MjModel the same way). */
#ifndef GUARDX_ART_H_
#define GUARDX_ART_H_

#include <stdint.h>

#define GX_ART_MAXLEG 4
#define GX_ART_MAXK 5  /* dofs per chain */
#define GX_ART_MAXB 3  /* bodies per chain */
#define GX_ART_MINVAL 1e-15f

enum { GX_ART_SLIDE = 0, GX_ART_HINGE = 1 };

typedef struct GxArtBody {
  float pos[3], quat[4], ipos[3];
  float inert[6]; /* about the COM, body axes: xx yy zz xy xz yz */
  float mass;
  int32_t ndof;
} GxArtBody;

typedef struct GxArtDof {
  int32_t type, limited, act, ctrllimited, forcelimited, pad0;
  float axis[3], jpos[3];
  float damping, armature, stiffness, springref, qpos0;
  float range[2], margin, invw, k, b, imp[5]; /* limit row: dof_invweight0, solref -> k,b, solimp */
  float gear, gain, biasv, ctrl_lo, ctrl_hi, force_lo, force_hi;
} GxArtDof;

typedef struct GxArtContact { /* one sphere against the floor plane z = 0 */
  int32_t on, body;
  float pos[3], radius, margin, mu, invw, k, b, imp[5];
} GxArtContact;

typedef struct GxArtSeg {
  int32_t nb, nk;
  int32_t dofbody[GX_ART_MAXK];
  GxArtBody body[GX_ART_MAXB];
  GxArtDof dof[GX_ART_MAXK];
  GxArtContact con;
} GxArtSeg;

typedef struct GxArtModel {
  int32_t nleg, nv, nu, max_iter;
  int32_t root_qadr, leg_qadr[GX_ART_MAXLEG]; /* first qpos/dof index of each segment */
  int32_t root_x, root_y;                      /* which root dof is joint 'robot_x' / 'robot_y' (layout2qpos, engine.py:623-639) */
  float dt, grav_z, tol_scale; /* tol_scale = opt.tolerance * meaninertia * max(1, nv_full) */
  GxArtSeg root;
  GxArtSeg leg[GX_ART_MAXLEG];
} GxArtModel;


#endif /* GUARDX_ART_H_ */
