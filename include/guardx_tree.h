/* guardx_tree.h -- compiled model of a GENERAL kinematic tree robot (free-joint or jointed root,
 * hinge / slide joints, sphere / capsule geoms against the floor plane and against each other):
 * the Doggo path of libguardx.so.  What mjx.put_model (engine.py:231) uploads in the reference,
 * restricted to the robot; the goal / hazard bodies World.build adds (world.py:310-326) are carried
 * as 2 + 2*hazards_num independent damped-spring slide dofs (`world_*`), because Doggo's defaults
 * class gives them a stiffness.  guardx_b200/tree_model.py fills this struct from the model
 * compiler (guardx_b200/mjcf.py).  All values fp32. */
#ifndef GUARDX_TREE_H_
#define GUARDX_TREE_H_

#include <stdint.h>

#define GX_TREE_MAXB 16    /* robot bodies */
#define GX_TREE_MAXJ 16    /* joints */
#define GX_TREE_MAXV 20    /* dofs */
#define GX_TREE_MAXQ 21    /* qpos */
#define GX_TREE_MAXU 16    /* actuators */
#define GX_TREE_MAXG 20    /* colliding geoms */
#define GX_TREE_MAXPAIR 96 /* contact candidates (geom pairs) */
#define GX_TREE_MAXCON 24  /* simultaneously active contacts kept per env */

enum { GX_TJ_FREE = 0, GX_TJ_SLIDE = 2, GX_TJ_HINGE = 3 };
enum { GX_TG_SPHERE = 2, GX_TG_CAPSULE = 3 };
enum { GX_TP_PLANE_SPHERE = 0, GX_TP_PLANE_CAPSULE = 1, GX_TP_CAPSULE_CAPSULE = 2 };

typedef struct GxTreeBody {
  int32_t parent;  /* -1: the world */
  int32_t jntadr, jntnum, dofadr, dofnum;
  uint32_t dofmask; /* dofs that move this body (bit i = dof i) */
  float pos[3], quat[4], ipos[3];
  float inert[6];  /* about the COM, body axes: xx yy zz xy xz yz */
  float mass;
} GxTreeBody;

typedef struct GxTreeJoint {
  int32_t type, body, qposadr, dofadr, limited, pad;
  float axis[3], pos[3];
  float qpos0, stiffness, springref, range[2], margin;
  float invw, k, b, imp[5]; /* limit row: dof_invweight0, solref -> k, b, solimp */
} GxTreeJoint;

typedef struct GxTreeDof {
  int32_t body, parent, act, ctrllimited, forcelimited, pad;
  float damping, armature;
  float gear, gain, biasv, ctrl_lo, ctrl_hi, force_lo, force_hi;
} GxTreeDof;

typedef struct GxTreeGeom {
  int32_t type, body;
  float pos[3], mat[9]; /* frame in the body */
  float radius, halflen;
} GxTreeGeom;

typedef struct GxTreePair { /* g1 = -1: the floor plane z = 0 */
  int32_t type, g1, g2, pad;
  float margin, mu, invw, k, b, imp[5];
} GxTreePair;

typedef struct GxTreeModel {
  int32_t nb, njnt, nq, nv, nu, ngeom, npair, max_iter;
  int32_t root_free;        /* 1: body 0 has a free joint */
  int32_t nworld;           /* 2 + 2 * hazards_num static-body slide dofs */
  int32_t root_xadr, root_yadr; /* qpos index the layout's robot x / y go to (layout2qpos, engine.py:623-639) */
  float dt, grav_z, tol_scale;
  float world_mass[2], world_damping, world_stiffness; /* [0] goal, [1] hazards: body mass + armature */
  GxTreeBody body[GX_TREE_MAXB];
  GxTreeJoint jnt[GX_TREE_MAXJ];
  GxTreeDof dof[GX_TREE_MAXV];
  GxTreeGeom geom[GX_TREE_MAXG];
  GxTreePair pair[GX_TREE_MAXPAIR];
} GxTreeModel;

#endif /* GUARDX_TREE_H_ */
