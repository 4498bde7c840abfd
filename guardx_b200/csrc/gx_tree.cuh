// gx_tree.cuh -- mjx.step for a GENERAL kinematic tree (free-joint or jointed root, hinge / slide
// joints, capsule / sphere contacts with the floor and with each other): the Doggo path.
//
// ONE THREAD PER ENV (the dense 19 x 19 solve does not split over a fixed lane count the way the star robots do).
// Per-env data: the big arrays (constraint Jacobian, body quantities, solver vectors) in a scratch buffer in global
// memory, interleaved over the 32 envs of a block (element k of lane l at chunk[k * stride + l], so a warp's
// accesses coalesce; constraint rows and per-row scalars in 128-bit groups); the Hessian / Cholesky workspace, the
// solve operand, qfrc_constraint and the mass matrix in a small matrix scratch (shared memory in the kernels).
// Both strides are compile-time constants on the device: with runtime strides the compiler cannot reorder a load
// past a store that might alias it, which serialised every read-modify-write chain.
// The algorithm is the dense pipeline of oracle/mjx_numpy.py -- kinematics, composite inertias (frame centred on
// the root body), RNE bias, Cholesky, limit and pyramidal contact rows for the ACTIVE candidates only, Newton with
// warm start and exact line search, implicit-damping Euler, quaternion integration.
//
// Like gx_art.cuh this header also compiles as plain C++ (tests/emu/) for CPU-side checks.
#ifndef GX_TREE_CUH_
#define GX_TREE_CUH_

#include "../../include/guardx_tree.h"
#include "gx_art.cuh"   // small vector / spatial algebra, gx_art_imp, gx_rcp, gx_rsqrt

#define GX_TREE_NROW (GX_TREE_MAXJ + 4 * GX_TREE_MAXCON)

// scratch layout (float offsets); NVX = GX_TREE_MAXV
enum {
  GXT_J = 0,
  GXT_RD = GXT_J + GX_TREE_NROW * GX_TREE_MAXV,   // row D
  GXT_RA = GXT_RD + GX_TREE_NROW,                 // row aref
  GXT_RJ = GXT_RA + GX_TREE_NROW,                 // row J*qacc - aref
  GXT_RV = GXT_RJ + GX_TREE_NROW,                 // row J*search
  GXT_CDOF = GXT_RV + GX_TREE_NROW,
  GXT_CDD = GXT_CDOF + 6 * GX_TREE_MAXV,
  GXT_XPOS = GXT_CDD + 6 * GX_TREE_MAXV,
  GXT_XQUAT = GXT_XPOS + 3 * GX_TREE_MAXB,
  GXT_XMAT = GXT_XQUAT + 4 * GX_TREE_MAXB,
  GXT_ANCH = GXT_XMAT + 9 * GX_TREE_MAXB,
  GXT_AXIS = GXT_ANCH + 3 * GX_TREE_MAXJ,
  GXT_CIN = GXT_AXIS + 3 * GX_TREE_MAXJ,
  GXT_CRB = GXT_CIN + 10 * GX_TREE_MAXB,
  GXT_CVEL = GXT_CRB + 10 * GX_TREE_MAXB,
  GXT_CACC = GXT_CVEL + 6 * GX_TREE_MAXB,
  GXT_CFRC = GXT_CACC + 6 * GX_TREE_MAXB,
  GXT_FS = GXT_CFRC + 6 * GX_TREE_MAXB,           // qfrc_smooth
  GXT_AS = GXT_FS + GX_TREE_MAXV,                 // qacc_smooth
  GXT_QA = GXT_AS + GX_TREE_MAXV,                 // qacc
  GXT_MA = GXT_QA + GX_TREE_MAXV,                 // M*qacc
  GXT_G = GXT_MA + GX_TREE_MAXV,                  // gradient / search direction
  GXT_MV = GXT_G + GX_TREE_MAXV,
  GXT_FC = GXT_MV + GX_TREE_MAXV,                 // qfrc_constraint
  GXT_T = GXT_FC + GX_TREE_MAXV,                  // temp vector
  GXT_AL = GXT_T + GX_TREE_MAXV,                  // compact list of the active rows (indices)
  GXT_SIZE = GXT_AL + GX_TREE_NROW
};

struct GxTreeSnap { float pos[3], mat[9]; };   // root xpos / xmat of the forward pass (pre-step pose)

#define TS(off) sc[(size_t)(off) * stride]

// The constraint rows J (region GXT_J = the first NROW * MAXV floats of every env's scratch) are stored in groups
// of four dofs: group g of row r of this env at jrows4[(r * MAXV / 4 + g) * stride], so a row is five 128-bit loads
// (20 scalar loads with their address arithmetic were 18 % of the kernel's instructions).
struct alignas(16) gxt_f4 { float x, y, z, w; };
#define GXT_DEVICE_MS 32   /* threads per block of the kernels = lanes the matrix scratch is interleaved over */
#define JQ(r, g) reinterpret_cast<gxt_f4*>(jrows)[(size_t)((r) * (GX_TREE_MAXV / 4) + (g)) * stride]
#define JE(r, i) (&JQ(r, (i) >> 2).x)[(i) & 3]
// Per-row scalars (regions GXT_RD .. GXT_RV, 4 * NROW floats) as one group per row: x = D, y = aref,
// z = J qacc - aref, w = J search -- the line-search loops read a row with one 128-bit load.
#define RQ(r) reinterpret_cast<gxt_f4*>(jrows + (size_t)GXT_RD * stride)[(size_t)(r) * stride]

// The arrays every inner loop of the Cholesky, of the solves and of the Hessian update walks through with
// dependent loads live in a SEPARATE, small matrix scratch (`sm`, stride `ms`): shared memory in the kernel (an
// L2 round trip per dependent load made the first version 5 ms per step), lower triangles packed.
// Layout for a model with nv dofs, tri = nv (nv + 1) / 2:  H [tri] | XS [nv] | FC [nv] | M [tri]
// (sized by the model, not by GX_TREE_MAXV: 418 floats per env for Doggo).  M is addressed through its own
// pointer / stride (`mm`, `mms`) so that a launch can place it elsewhere (global memory was tried: slower).
#define GXT_XS (tri)                  /* operand vector of the solves, of M*x and of J*x */
#define GXT_JR (tri + nv)             /* qfrc_constraint (accumulated row by row) */
#define FCS(i) sm[(size_t)(GXT_JR + (i)) * ms]
#define GXT_MOFF(nv) ((nv) * ((nv) + 1) / 2 + 2 * (nv))
#define GXT_SM_FLOATS_OF(nv, with_m) ((nv) * ((nv) + 1) / 2 * ((with_m) ? 2 : 1) + 2 * (nv))
#define GXT_SM_FLOATS GXT_SM_FLOATS_OF(GX_TREE_MAXV, 1)
#define MM(i, j) mm[(size_t)((i) * ((i) + 1) / 2 + (j)) * mms]             /* j <= i */
#define HH(i, j) sm[(size_t)((i) * ((i) + 1) / 2 + (j)) * ms]              /* j <= i */
#define XS(i) sm[(size_t)(GXT_XS + (i)) * ms]

GX_HD void gxt_get(const float* sc, size_t stride, int off, int n, float* out) {
  for (int k = 0; k < n; ++k) out[k] = TS(off + k);
}
GX_HD void gxt_put(float* sc, size_t stride, int off, int n, const float* in) {
  for (int k = 0; k < n; ++k) TS(off + k) = in[k];
}

// s - sum_k a[k] b[k]; elements `ms` apart.  Two accumulators (even / odd k): with one or two warps per
// scheduler the kernel runs at the latency of its dependent chains, so the inner products are written for
// instruction-level parallelism -- independent chains, loads of a batch issued before its FMAs.
GX_HD float gxt_dotsub(const float* a, const float* b, size_t ms, int n, float s) {
  float s1 = 0.f;
  int k = 0;
  for (; k + 1 < n; k += 2) {
    const float a0 = a[(size_t)k * ms], a1 = a[(size_t)(k + 1) * ms];
    const float b0 = b[(size_t)k * ms], b1 = b[(size_t)(k + 1) * ms];
    s = gx_fma(-a0, b0, s);
    s1 = gx_fma(-a1, b1, s1);
  }
  if (k < n) s = gx_fma(-a[(size_t)k * ms], b[(size_t)k * ms], s);
  return s + s1;
}
// in-place Cholesky of the packed lower triangle H (n x n) at sm, inverse diagonal stored.  Left-looking,
// two rows at a time (they share the loads of row j), two accumulators per row.
GX_HD void gxt_chol(float* sm, size_t ms, int n) {
  float* rj = sm;                                          // row j of H
  for (int j = 0; j < n; ++j) {
    const float inv = gx_rsqrt(gxt_dotsub(rj, rj, ms, j, rj[(size_t)j * ms]));
    rj[(size_t)j * ms] = inv;
    float* ri = rj + (size_t)(j + 1) * ms;                 // row i = j + 1 ...
    int i = j + 1;
    for (; i + 1 < n; i += 2) {
      float* r2 = ri + (size_t)(i + 1) * ms;               // row i + 1
      float a0 = ri[(size_t)j * ms], a1 = 0.f, b0 = r2[(size_t)j * ms], b1 = 0.f;
      int k = 0;
      for (; k + 1 < j; k += 2) {
        const float l0 = rj[(size_t)k * ms], l1 = rj[(size_t)(k + 1) * ms];
        const float x0 = ri[(size_t)k * ms], x1 = ri[(size_t)(k + 1) * ms];
        const float y0 = r2[(size_t)k * ms], y1 = r2[(size_t)(k + 1) * ms];
        a0 = gx_fma(-x0, l0, a0); a1 = gx_fma(-x1, l1, a1);
        b0 = gx_fma(-y0, l0, b0); b1 = gx_fma(-y1, l1, b1);
      }
      if (k < j) {
        const float l0 = rj[(size_t)k * ms];
        a0 = gx_fma(-ri[(size_t)k * ms], l0, a0);
        b0 = gx_fma(-r2[(size_t)k * ms], l0, b0);
      }
      ri[(size_t)j * ms] = (a0 + a1) * inv;
      r2[(size_t)j * ms] = (b0 + b1) * inv;
      ri = r2 + (size_t)(i + 2) * ms;
    }
    if (i < n) ri[(size_t)j * ms] = gxt_dotsub(ri, rj, ms, j, ri[(size_t)j * ms]) * inv;
    rj += (size_t)(j + 1) * ms;
  }
}
// solve L L^T x = b in place, x = XS(0..n)
GX_HD void gxt_chol_solve(float* sm, size_t ms, int n) {
  const int tri = n * (n + 1) / 2;
  const float* H = sm;
  float* x = sm + (size_t)GXT_XS * ms;
  const float* ri = H;
  for (int i = 0; i < n; ++i) {
    x[(size_t)i * ms] = gxt_dotsub(ri, x, ms, i, x[(size_t)i * ms]) * ri[(size_t)i * ms];
    ri += (size_t)(i + 1) * ms;
  }
  for (int i = n - 1; i >= 0; --i) {
    float s = x[(size_t)i * ms], s1 = 0.f;
    const float* dg = H + (size_t)(i * (i + 1) / 2 + i) * ms;   // H(i, i)
    const float* p = dg + (size_t)(i + 1) * ms;                 // H(i + 1, i)
    int k = i + 1;
    for (; k + 1 < n; k += 2) {
      const float* p1 = p + (size_t)(k + 1) * ms;               // H(k + 1, i)
      const float h0 = *p, h1 = *p1;
      s = gx_fma(-h0, x[(size_t)k * ms], s);
      s1 = gx_fma(-h1, x[(size_t)(k + 1) * ms], s1);
      p = p1 + (size_t)(k + 2) * ms;
    }
    if (k < n) s = gx_fma(-*p, x[(size_t)k * ms], s);
    x[(size_t)i * ms] = (s + s1) * *dg;
  }
}
// One constraint row (global scratch) into registers: every load is in flight before the first use, so a row
// costs ONE L2 round trip (loaded a few elements at a time each batch paid its own ~400 cycles; the ncu source
// page showed the first FMA after each batch as the hottest instructions of the kernel).  Compile-time indices.
GX_HD void gxt_row_load(const float* jrows, size_t stride, int r, float (&row)[GX_TREE_MAXV]) {
  const gxt_f4* p = reinterpret_cast<const gxt_f4*>(jrows) + (size_t)(r * (GX_TREE_MAXV / 4)) * stride;
#pragma unroll
  for (int g = 0; g < GX_TREE_MAXV / 4; ++g) {
    const gxt_f4 v = p[(size_t)g * stride];
    row[4 * g] = v.x; row[4 * g + 1] = v.y; row[4 * g + 2] = v.z; row[4 * g + 3] = v.w;
  }
}
// s + sum_i J[r][i] x[i]; x: matrix scratch
GX_HD float gxt_rowdot(const float* jrows, size_t stride, int r, const float* x, size_t ms, int n, float s) {
  float row[GX_TREE_MAXV], s1 = 0.f;
  gxt_row_load(jrows, stride, r, row);
#pragma unroll
  for (int i = 0; i < GX_TREE_MAXV; i += 2) {
    if (i < n) s = gx_fma(row[i], x[(size_t)i * ms], s);
    if (i + 1 < n) s1 = gx_fma(row[i + 1], x[(size_t)(i + 1) * ms], s1);
  }
  return s + s1;
}
// row i of the symmetric packed matrix A (stride as) times x (stride ms), j ascending
GX_HD float gxt_symrow(const float* A, size_t as, const float* x, size_t ms, int n, int i) {
  const float* r = A + (size_t)(i * (i + 1) / 2) * as;
  float s = 0.f, s1 = 0.f;
  int j = 0;
  for (; j + 1 <= i; j += 2) {
    const float a0 = r[(size_t)j * as], a1 = r[(size_t)(j + 1) * as];
    s = gx_fma(a0, x[(size_t)j * ms], s);
    s1 = gx_fma(a1, x[(size_t)(j + 1) * ms], s1);
  }
  if (j <= i) s = gx_fma(r[(size_t)j * as], x[(size_t)j * ms], s);
  const float* p = r + (size_t)(2 * i + 1) * as;               // A(i + 1, i)
  j = i + 1;
  for (; j + 1 < n; j += 2) {
    const float* p1 = p + (size_t)(j + 1) * as;
    const float a0 = *p, a1 = *p1;
    s = gx_fma(a0, x[(size_t)j * ms], s);
    s1 = gx_fma(a1, x[(size_t)(j + 1) * ms], s1);
    p = p1 + (size_t)(j + 2) * as;
  }
  if (j < n) s = gx_fma(*p, x[(size_t)j * ms], s);
  return s + s1;
}
// H := M (packed triangles), loads in batches of 4
GX_HD void gxt_copy_tri(const float* mm, size_t mms, float* sm, size_t ms, int tri) {
  for (int k = 0; k < tri; k += 4) {
    const float a0 = mm[(size_t)k * mms];
    const float a1 = k + 1 < tri ? mm[(size_t)(k + 1) * mms] : 0.f;
    const float a2 = k + 2 < tri ? mm[(size_t)(k + 2) * mms] : 0.f;
    const float a3 = k + 3 < tri ? mm[(size_t)(k + 3) * mms] : 0.f;
    sm[(size_t)k * ms] = a0;
    if (k + 1 < tri) sm[(size_t)(k + 1) * ms] = a1;
    if (k + 2 < tri) sm[(size_t)(k + 2) * ms] = a2;
    if (k + 3 < tri) sm[(size_t)(k + 3) * ms] = a3;
  }
}

// closest points of two segments (oracle/mjx_numpy.py closest_segment_to_segment_points)
GX_HD void gxt_seg_point(const float* a, const float* b, const float* pt, float* out) {
  float ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ap[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  float t = gx_dot3(ap, ab) * gx_rcp(gx_dot3(ab, ab) + 1e-6f);
  t = gx_clip(t, 0.0f, 1.0f);
  out[0] = gx_fma(t, ab[0], a[0]); out[1] = gx_fma(t, ab[1], a[1]); out[2] = gx_fma(t, ab[2], a[2]);
}
GX_HD void gxt_seg_seg(const float* a0, const float* a1, const float* b0, const float* b1, float* pa, float* pb) {
  float da[3] = {a1[0] - a0[0], a1[1] - a0[1], a1[2] - a0[2]}, db[3] = {b1[0] - b0[0], b1[1] - b0[1], b1[2] - b0[2]};
  const float la = gx_sqrt(gx_dot3(da, da)), lb = gx_sqrt(gx_dot3(db, db));
  const float ia = gx_rcp(la), ib = gx_rcp(lb);
  float dira[3] = {da[0] * ia, da[1] * ia, da[2] * ia}, dirb[3] = {db[0] * ib, db[1] * ib, db[2] * ib};
  const float ha = 0.5f * la, hb = 0.5f * lb;
  float am[3], bm[3], tr[3];
  for (int k = 0; k < 3; ++k) { am[k] = gx_fma(dira[k], ha, a0[k]); bm[k] = gx_fma(dirb[k], hb, b0[k]); tr[k] = am[k] - bm[k]; }
  const float d_a = gx_dot3(dira, tr), d_b = gx_dot3(dirb, tr), dd = gx_dot3(dira, dirb);
  float ta = (-d_a + d_b * dd) * gx_rcp((1.0f - dd * dd) + 1e-6f);
  float tb = d_b + ta * dd;
  ta = gx_clip(ta, -ha, ha);
  tb = gx_clip(tb, -hb, hb);
  float ba[3], bb[3], na[3], nb[3];
  for (int k = 0; k < 3; ++k) { ba[k] = gx_fma(dira[k], ta, am[k]); bb[k] = gx_fma(dirb[k], tb, bm[k]); }
  gxt_seg_point(a0, a1, bb, na);
  gxt_seg_point(b0, b1, ba, nb);
  float d1 = 0.f, d2 = 0.f;
  for (int k = 0; k < 3; ++k) { d1 += (na[k] - bb[k]) * (na[k] - bb[k]); d2 += (nb[k] - ba[k]) * (nb[k] - ba[k]); }
  for (int k = 0; k < 3; ++k) { pa[k] = d1 < d2 ? na[k] : ba[k]; pb[k] = d1 < d2 ? bb[k] : nb[k]; }
}
GX_HD void gxt_make_frame(const float* n, float* t1, float* t2) {
  const bool usey = gx_abs(n[1]) < 0.5f;
  float ref[3] = {0.f, usey ? 1.f : 0.f, usey ? 0.f : 1.f};
  const float d = gx_dot3(n, ref);
  float t[3] = {ref[0] - n[0] * d, ref[1] - n[1] * d, ref[2] - n[2] * d};
  const float inv = gx_rsqrt(gx_dot3(t, t));
  t1[0] = t[0] * inv; t1[1] = t[1] * inv; t1[2] = t[2] * inv;
  gx_cross(n, t1, t2);
}

// One mjx.step (or mjx.forward when forward_only).  q [nq], v [nv], warm [nv] are read with stride
// `ss`; results go to qo, vo (post-step, same stride; untouched when forward_only) and wo (the
// solver's qacc).  ctrl [nu] is thread-local.  sc: scratch, GXT_SIZE floats with stride `stride`;
// sm: matrix scratch (H | XS | JR), stride `ms`; mm: packed mass matrix, stride `mms`.
GX_HD void gx_tree_step(const GxTreeModel& m, const float* q, const float* v, const float* warm, size_t ss,
                        const float* ctrl, float* sc, size_t stride_arg, float* jrows, float* sm, size_t ms_arg, float* mm, size_t mms_arg, bool forward_only, float* qo, float* vo,
                        float* wo, GxTreeSnap& snap, int* iters_out) {
  const int nv = m.nv, tri = nv * (nv + 1) / 2;
#ifdef __CUDA_ARCH__
  // One warp per block: the matrix scratch is interleaved over exactly 32 lanes.  A compile-time stride turns every
  // shared-memory access into base + immediate, and -- what matters -- lets the compiler see that hrow[j] and hrow[j']
  // never alias, so the loads of the unrolled Hessian update / Cholesky are no longer serialised behind the stores.
  constexpr size_t ms = GXT_DEVICE_MS, mms = GXT_DEVICE_MS, stride = GXT_DEVICE_MS;   // (the global scratch too: one chunk per block)
  (void)ms_arg; (void)mms_arg; (void)stride_arg;
#else
  const size_t ms = ms_arg, mms = mms_arg, stride = stride_arg;
#endif
  // ---------------- kinematics ----------------
  for (int b = 0; b < m.nb; ++b) {
    const GxTreeBody& B = m.body[b];
    float pp[3] = {0.f, 0.f, 0.f}, pq[4] = {1.f, 0.f, 0.f, 0.f}, R[9], t[3], pos[3], quat[4];
    if (B.parent >= 0) { gxt_get(sc, stride, GXT_XPOS + 3 * B.parent, 3, pp); gxt_get(sc, stride, GXT_XQUAT + 4 * B.parent, 4, pq); }
    gx_quat2mat(pq, R);
    gx_rot(R, B.pos, t);
    pos[0] = pp[0] + t[0]; pos[1] = pp[1] + t[1]; pos[2] = pp[2] + t[2];
    gx_quat_mul(pq, B.quat, quat);
    for (int k = 0; k < B.jntnum; ++k) {
      const int j = B.jntadr + k;
      const GxTreeJoint& Jn = m.jnt[j];
      float anchor[3], axis[3];
      if (Jn.type == GX_TJ_FREE) {
        for (int a = 0; a < 3; ++a) pos[a] = q[(size_t)(Jn.qposadr + a) * ss];
        for (int a = 0; a < 4; ++a) quat[a] = q[(size_t)(Jn.qposadr + 3 + a) * ss];
        const float nr = gx_rsqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
        for (int a = 0; a < 4; ++a) quat[a] *= nr;
        for (int a = 0; a < 3; ++a) { anchor[a] = pos[a]; axis[a] = 0.f; }
      } else {
        gx_quat2mat(quat, R);
        gx_rot(R, Jn.pos, t);
        anchor[0] = pos[0] + t[0]; anchor[1] = pos[1] + t[1]; anchor[2] = pos[2] + t[2];
        gx_rot(R, Jn.axis, axis);
        const float dq = q[(size_t)Jn.qposadr * ss] - Jn.qpos0;
        if (Jn.type == GX_TJ_SLIDE) {
          for (int a = 0; a < 3; ++a) pos[a] = gx_fma(axis[a], dq, pos[a]);
        } else {
          float s, c, ql[4], qn[4];
          gx_sincos(0.5f * dq, &s, &c);
          ql[0] = c; ql[1] = Jn.axis[0] * s; ql[2] = Jn.axis[1] * s; ql[3] = Jn.axis[2] * s;
          gx_quat_mul(quat, ql, qn);
          for (int a = 0; a < 4; ++a) quat[a] = qn[a];
          gx_quat2mat(quat, R);
          gx_rot(R, Jn.pos, t);
          for (int a = 0; a < 3; ++a) pos[a] = anchor[a] - t[a];
        }
      }
      gxt_put(sc, stride, GXT_ANCH + 3 * j, 3, anchor);
      gxt_put(sc, stride, GXT_AXIS + 3 * j, 3, axis);
    }
    const float nr = gx_rsqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
    for (int a = 0; a < 4; ++a) quat[a] *= nr;
    gx_quat2mat(quat, R);
    gxt_put(sc, stride, GXT_XPOS + 3 * b, 3, pos);
    gxt_put(sc, stride, GXT_XQUAT + 4 * b, 4, quat);
    gxt_put(sc, stride, GXT_XMAT + 9 * b, 9, R);
  }
  float c[3];
  gxt_get(sc, stride, GXT_XPOS, 3, c);     // reference point of all spatial quantities: the root body
  gxt_get(sc, stride, GXT_XPOS, 3, snap.pos);
  gxt_get(sc, stride, GXT_XMAT, 9, snap.mat);
  // ---------------- cdof, velocities, spatial inertias, RNE body forces (forward pass) ----------------
  for (int b = 0; b < m.nb; ++b) {
    const GxTreeBody& B = m.body[b];
    float cvel[6] = {0, 0, 0, 0, 0, 0}, cacc[6] = {0, 0, 0, 0, 0, -m.grav_z}, R[9], xp[3];
    if (B.parent >= 0) { gxt_get(sc, stride, GXT_CVEL + 6 * B.parent, 6, cvel); gxt_get(sc, stride, GXT_CACC + 6 * B.parent, 6, cacc); }
    gxt_get(sc, stride, GXT_XMAT + 9 * b, 9, R);
    gxt_get(sc, stride, GXT_XPOS + 3 * b, 3, xp);
    for (int k = 0; k < B.jntnum; ++k) {
      const int j = B.jntadr + k;
      const GxTreeJoint& Jn = m.jnt[j];
      if (Jn.type == GX_TJ_FREE) {
        const int i0 = Jn.dofadr;
        float cd[6][6];
        for (int a = 0; a < 3; ++a) {
          for (int e = 0; e < 6; ++e) cd[a][e] = 0.f;
          cd[a][3 + a] = 1.f;
          float ax[3] = {R[a], R[3 + a], R[6 + a]}, off[3] = {c[0] - xp[0], c[1] - xp[1], c[2] - xp[2]};
          cd[3 + a][0] = ax[0]; cd[3 + a][1] = ax[1]; cd[3 + a][2] = ax[2];
          gx_cross(ax, off, cd[3 + a] + 3);
        }
        // mj_comVel: translations have cdof_dot = 0; the three rotations all see the velocity after them
        for (int a = 0; a < 3; ++a) {
          const float vi = v[(size_t)(i0 + a) * ss];
          for (int e = 0; e < 6; ++e) { cvel[e] = gx_fma(cd[a][e], vi, cvel[e]); TS(GXT_CDD + 6 * (i0 + a) + e) = 0.f; }
        }
        float dd[3][6];
        for (int a = 0; a < 3; ++a) gx_cross_motion(cvel, cd[3 + a], dd[a]);
        for (int a = 0; a < 3; ++a) {
          const float vi = v[(size_t)(i0 + 3 + a) * ss];
          for (int e = 0; e < 6; ++e) { cvel[e] = gx_fma(cd[3 + a][e], vi, cvel[e]); cacc[e] = gx_fma(dd[a][e], vi, cacc[e]); TS(GXT_CDD + 6 * (i0 + 3 + a) + e) = dd[a][e]; }
        }
        for (int a = 0; a < 6; ++a) gxt_put(sc, stride, GXT_CDOF + 6 * (i0 + a), 6, cd[a]);
      } else {
        const int i = Jn.dofadr;
        float cd[6], axis[3], anchor[3], dd[6];
        gxt_get(sc, stride, GXT_AXIS + 3 * j, 3, axis);
        gxt_get(sc, stride, GXT_ANCH + 3 * j, 3, anchor);
        if (Jn.type == GX_TJ_SLIDE) {
          cd[0] = cd[1] = cd[2] = 0.f; cd[3] = axis[0]; cd[4] = axis[1]; cd[5] = axis[2];
        } else {
          float off[3] = {c[0] - anchor[0], c[1] - anchor[1], c[2] - anchor[2]};
          cd[0] = axis[0]; cd[1] = axis[1]; cd[2] = axis[2];
          gx_cross(axis, off, cd + 3);
        }
        gx_cross_motion(cvel, cd, dd);
        const float vi = v[(size_t)i * ss];
        for (int e = 0; e < 6; ++e) { cacc[e] = gx_fma(dd[e], vi, cacc[e]); cvel[e] = gx_fma(cd[e], vi, cvel[e]); }
        gxt_put(sc, stride, GXT_CDOF + 6 * i, 6, cd);
        gxt_put(sc, stride, GXT_CDD + 6 * i, 6, dd);
      }
    }
    gxt_put(sc, stride, GXT_CVEL + 6 * b, 6, cvel);
    gxt_put(sc, stride, GXT_CACC + 6 * b, 6, cacc);
    // spatial inertia about c
    float I[10], t[3], d[3], A[9];
    gx_rot(R, B.ipos, t);
    d[0] = xp[0] + t[0] - c[0]; d[1] = xp[1] + t[1] - c[1]; d[2] = xp[2] + t[2] - c[2];
    const float* Jb = B.inert;
    for (int r = 0; r < 3; ++r) {
      A[3 * r + 0] = gx_fma(R[3 * r + 2], Jb[4], gx_fma(R[3 * r + 1], Jb[3], R[3 * r] * Jb[0]));
      A[3 * r + 1] = gx_fma(R[3 * r + 2], Jb[5], gx_fma(R[3 * r + 1], Jb[1], R[3 * r] * Jb[3]));
      A[3 * r + 2] = gx_fma(R[3 * r + 2], Jb[2], gx_fma(R[3 * r + 1], Jb[5], R[3 * r] * Jb[4]));
    }
    const float ms = B.mass, dd2 = gx_dot3(d, d);
    I[0] = gx_dot3(A + 0, R + 0) + ms * (dd2 - d[0] * d[0]);
    I[1] = gx_dot3(A + 3, R + 3) + ms * (dd2 - d[1] * d[1]);
    I[2] = gx_dot3(A + 6, R + 6) + ms * (dd2 - d[2] * d[2]);
    I[3] = gx_dot3(A + 0, R + 3) - ms * d[0] * d[1];
    I[4] = gx_dot3(A + 0, R + 6) - ms * d[0] * d[2];
    I[5] = gx_dot3(A + 3, R + 6) - ms * d[1] * d[2];
    I[6] = ms * d[0]; I[7] = ms * d[1]; I[8] = ms * d[2]; I[9] = ms;
    gxt_put(sc, stride, GXT_CIN + 10 * b, 10, I);
    gxt_put(sc, stride, GXT_CRB + 10 * b, 10, I);
    float Ia[6], Iv[6], x[6], f[6];
    gx_inert_mul(I, cacc, Ia);
    gx_inert_mul(I, cvel, Iv);
    gx_cross_force(cvel, Iv, x);
    for (int e = 0; e < 6; ++e) f[e] = Ia[e] + x[e];
    gxt_put(sc, stride, GXT_CFRC + 6 * b, 6, f);
  }
  // ---------------- backward pass: composite inertias and accumulated forces ----------------
  for (int b = m.nb - 1; b >= 1; --b) {
    const int p = m.body[b].parent;
    if (p < 0) continue;
    // all 32 loads first (a read-modify-write per element would wait for each load: the compiler may not move a
    // load above a store that could alias it)
    float cb[10], cp[10], fb[6], fp[6];
    gxt_get(sc, stride, GXT_CRB + 10 * b, 10, cb); gxt_get(sc, stride, GXT_CRB + 10 * p, 10, cp);
    gxt_get(sc, stride, GXT_CFRC + 6 * b, 6, fb); gxt_get(sc, stride, GXT_CFRC + 6 * p, 6, fp);
    for (int e = 0; e < 10; ++e) cp[e] += cb[e];
    for (int e = 0; e < 6; ++e) fp[e] += fb[e];
    gxt_put(sc, stride, GXT_CRB + 10 * p, 10, cp);
    gxt_put(sc, stride, GXT_CFRC + 6 * p, 6, fp);
  }
  // ---------------- mass matrix, smooth forces ----------------
  for (int i = 0; i < nv; ++i) for (int j = 0; j <= i; ++j) MM(i, j) = 0.f;
  for (int i = 0; i < nv; ++i) {
    const GxTreeDof& D = m.dof[i];
    float I[10], cd[6], f[6], F[6];
    gxt_get(sc, stride, GXT_CRB + 10 * D.body, 10, I);
    gxt_get(sc, stride, GXT_CDOF + 6 * i, 6, cd);
    gx_inert_mul(I, cd, f);
    for (int j = i; j >= 0; j = m.dof[j].parent) {
      float cj[6];
      gxt_get(sc, stride, GXT_CDOF + 6 * j, 6, cj);
      const float mij = gx_dot6(cj, f);
      MM(i, j) = mij;
    }
    MM(i, i) += D.armature;
    gxt_get(sc, stride, GXT_CFRC + 6 * D.body, 6, F);
    float fs = -(D.damping * v[(size_t)i * ss]) - gx_dot6(cd, F);
    if (D.act >= 0) {
      float u = ctrl[D.act];
      if (D.ctrllimited) u = gx_clip(u, D.ctrl_lo, D.ctrl_hi);
      float force = D.gain * u + D.biasv * (D.gear * v[(size_t)i * ss]);
      if (D.forcelimited) force = gx_clip(force, D.force_lo, D.force_hi);
      fs += D.gear * force;
    }
    TS(GXT_FS + i) = fs;
  }
  for (int j = 0; j < m.njnt; ++j) {
    const GxTreeJoint& Jn = m.jnt[j];
    if (Jn.type != GX_TJ_FREE && Jn.stiffness != 0.0f)
      TS(GXT_FS + Jn.dofadr) -= Jn.stiffness * (q[(size_t)Jn.qposadr * ss] - Jn.springref);
  }
  // ---------------- qacc_smooth ----------------
  gxt_copy_tri(mm, mms, sm, ms, tri);
  for (int i = 0; i < nv; ++i) XS(i) = TS(GXT_FS + i);
  gxt_chol(sm, ms, nv);
  gxt_chol_solve(sm, ms, nv);
  for (int i = 0; i < nv; ++i) TS(GXT_AS + i) = XS(i);
  // ---------------- constraint rows: limits, then the active contacts ----------------
  int nrow = 0;
  for (int j = 0; j < m.njnt; ++j) {
    const GxTreeJoint& Jn = m.jnt[j];
    if (Jn.type == GX_TJ_FREE || !Jn.limited) continue;
    const float qj = q[(size_t)Jn.qposadr * ss];
    const float dmin = qj - Jn.range[0], dmax = Jn.range[1] - qj;
    const float pos = gx_min(dmin, dmax) - Jn.margin;
    if (!(pos < 0.0f)) continue;
    const float sgn = dmin < dmax ? 1.0f : -1.0f;
    const float imp = gx_art_imp(Jn.imp, pos);
    const float Rr = gx_max(gx_div(Jn.invw * (1.0f - imp), imp), GX_ART_MINVAL);
    for (int g = 0; g < GX_TREE_MAXV / 4; ++g) JQ(nrow, g) = gxt_f4{0.f, 0.f, 0.f, 0.f};
    JE(nrow, Jn.dofadr) = sgn;
    RQ(nrow) = gxt_f4{gx_rcp(Rr), -Jn.b * (sgn * v[(size_t)Jn.dofadr * ss]) - Jn.k * imp * pos, 0.f, 0.f};
    ++nrow;
  }
  int ncon = 0;
  for (int pi = 0; pi < m.npair; ++pi) {
    const GxTreePair& Pr = m.pair[pi];
    const GxTreeGeom& G2 = m.geom[Pr.g2];
    float R2[9], x2[3], t[3], p2[3], ax2[3];
    gxt_get(sc, stride, GXT_XMAT + 9 * G2.body, 9, R2);
    gxt_get(sc, stride, GXT_XPOS + 3 * G2.body, 3, x2);
    gx_rot(R2, G2.pos, t);
    p2[0] = x2[0] + t[0]; p2[1] = x2[1] + t[1]; p2[2] = x2[2] + t[2];
    { float lz[3] = {G2.mat[2], G2.mat[5], G2.mat[8]}; gx_rot(R2, lz, ax2); }
    // up to two contacts (plane-capsule: one per end)
    float cdist[2], cpos[2][3], nrm[3], t1[3], t2[3];
    int nc = 0, b1 = -1;
    if (Pr.type == GX_TP_PLANE_SPHERE || Pr.type == GX_TP_PLANE_CAPSULE) {
      nrm[0] = 0.f; nrm[1] = 0.f; nrm[2] = 1.f;
      if (Pr.type == GX_TP_PLANE_SPHERE) {
        gxt_make_frame(nrm, t1, t2);
        cdist[0] = p2[2] - G2.radius;
        cpos[0][0] = p2[0]; cpos[0][1] = p2[1]; cpos[0][2] = p2[2] - (G2.radius + 0.5f * cdist[0]);
        nc = 1;
      } else {
        // tangent 1 along the capsule axis projected into the plane (collision_primitive.plane_capsule)
        float bv[3] = {ax2[0], ax2[1], 0.f};
        const float bn = gx_sqrt(bv[0] * bv[0] + bv[1] * bv[1]);
        if (bn < 0.5f) { t1[0] = 0.f; t1[1] = 1.f; t1[2] = 0.f; }
        else { const float ib = gx_rcp(bn); t1[0] = bv[0] * ib; t1[1] = bv[1] * ib; t1[2] = 0.f; }
        gx_cross(nrm, t1, t2);
        for (int e = 0; e < 2; ++e) {
          const float sg = e == 0 ? G2.halflen : -G2.halflen;
          float ep[3] = {gx_fma(ax2[0], sg, p2[0]), gx_fma(ax2[1], sg, p2[1]), gx_fma(ax2[2], sg, p2[2])};
          cdist[e] = ep[2] - G2.radius;
          cpos[e][0] = ep[0]; cpos[e][1] = ep[1]; cpos[e][2] = ep[2] - (G2.radius + 0.5f * cdist[e]);
        }
        nc = 2;
      }
    } else {
      const GxTreeGeom& G1 = m.geom[Pr.g1];
      b1 = G1.body;
      float R1[9], x1[3], p1[3], ax1[3];
      gxt_get(sc, stride, GXT_XMAT + 9 * G1.body, 9, R1);
      gxt_get(sc, stride, GXT_XPOS + 3 * G1.body, 3, x1);
      gx_rot(R1, G1.pos, t);
      p1[0] = x1[0] + t[0]; p1[1] = x1[1] + t[1]; p1[2] = x1[2] + t[2];
      { float lz[3] = {G1.mat[2], G1.mat[5], G1.mat[8]}; gx_rot(R1, lz, ax1); }
      float a0[3], a1[3], c0[3], c1[3], pa[3], pb[3];
      for (int e = 0; e < 3; ++e) {
        a0[e] = gx_fma(-ax1[e], G1.halflen, p1[e]); a1[e] = gx_fma(ax1[e], G1.halflen, p1[e]);
        c0[e] = gx_fma(-ax2[e], G2.halflen, p2[e]); c1[e] = gx_fma(ax2[e], G2.halflen, p2[e]);
      }
      gxt_seg_seg(a0, a1, c0, c1, pa, pb);
      float dv[3] = {pb[0] - pa[0], pb[1] - pa[1], pb[2] - pa[2]};
      const float dn = gx_sqrt(gx_dot3(dv, dv));
      if (dn == 0.0f) { nrm[0] = 1.f; nrm[1] = 0.f; nrm[2] = 0.f; }
      else { const float id = gx_rcp(dn); nrm[0] = dv[0] * id; nrm[1] = dv[1] * id; nrm[2] = dv[2] * id; }
      cdist[0] = dn - (G1.radius + G2.radius);
      for (int e = 0; e < 3; ++e) cpos[0][e] = gx_fma(nrm[e], G1.radius + 0.5f * cdist[0], pa[e]);
      gxt_make_frame(nrm, t1, t2);
      nc = 1;
    }
    for (int e = 0; e < nc; ++e) {
      const float pos = cdist[e] - Pr.margin;
      if (!(pos < 0.0f) || ncon >= GX_TREE_MAXCON) continue;
      ++ncon;
      const uint32_t m2 = m.body[G2.body].dofmask, m1 = b1 >= 0 ? m.body[b1].dofmask : 0u;
      const float off[3] = {cpos[e][0] - c[0], cpos[e][1] - c[1], cpos[e][2] - c[2]};
      const float imp = gx_art_imp(Pr.imp, pos);
      const float Rr = gx_max(gx_div(Pr.invw * (1.0f - imp), imp), GX_ART_MINVAL);
      const float Dv = gx_rcp(Rr), kp = Pr.k * imp * pos;
      float vel[4] = {0.f, 0.f, 0.f, 0.f};
      for (int i = 0; i < nv; ++i) {
        const int s2 = (m2 >> i) & 1u, s1 = (m1 >> i) & 1u;
        float jn = 0.f, j1 = 0.f, j2 = 0.f;
        if (s2 != s1) {
          float cd[6], x[3];
          gxt_get(sc, stride, GXT_CDOF + 6 * i, 6, cd);
          gx_cross(cd, off, x);
          float jp[3] = {cd[3] + x[0], cd[4] + x[1], cd[5] + x[2]};
          const float sgn = s2 ? 1.0f : -1.0f;
          jn = sgn * gx_dot3(nrm, jp); j1 = sgn * gx_dot3(t1, jp); j2 = sgn * gx_dot3(t2, jp);
        }
        const float r0 = gx_fma(Pr.mu, j1, jn), r1 = gx_fma(-Pr.mu, j1, jn), r2 = gx_fma(Pr.mu, j2, jn), r3 = gx_fma(-Pr.mu, j2, jn);
        JE(nrow + 0, i) = r0; JE(nrow + 1, i) = r1; JE(nrow + 2, i) = r2; JE(nrow + 3, i) = r3;
        const float vi = v[(size_t)i * ss];
        vel[0] = gx_fma(r0, vi, vel[0]); vel[1] = gx_fma(r1, vi, vel[1]); vel[2] = gx_fma(r2, vi, vel[2]); vel[3] = gx_fma(r3, vi, vel[3]);
      }
      for (int i = nv; i < GX_TREE_MAXV; ++i) { JE(nrow + 0, i) = 0.f; JE(nrow + 1, i) = 0.f; JE(nrow + 2, i) = 0.f; JE(nrow + 3, i) = 0.f; }
      for (int r = 0; r < 4; ++r) RQ(nrow + r) = gxt_f4{Dv, -Pr.b * vel[r] - kp, 0.f, 0.f};
      nrow += 4;
    }
  }
  // ---------------- solver.solve: Newton on the primal cost (dense) ----------------
  int iters = 0;
  for (int i = 0; i < nv; ++i) { TS(GXT_QA + i) = TS(GXT_AS + i); FCS(i) = 0.f; }
  if (nrow > 0) {
    // cost of a candidate (and its M*qacc, J*qacc - aref): 0 = smooth, 1 = warm start
    float cost_s = 0.f;
    for (int cnd = 0; cnd < 3; ++cnd) {
      const bool usew = cnd == 1;
      for (int i = 0; i < nv; ++i) { const float qa = usew ? warm[(size_t)i * ss] : TS(GXT_AS + i); TS(GXT_QA + i) = qa; XS(i) = qa; }
      float cost = 0.f;
      for (int i = 0; i < nv; ++i) {
        float s = 0.f;
        if (usew) s = gxt_symrow(mm, mms, sm + (size_t)GXT_XS * ms, ms, nv, i);
        else s = TS(GXT_FS + i);
        TS(GXT_MA + i) = s;
        cost = gx_fma(0.5f * (s - TS(GXT_FS + i)), TS(GXT_QA + i) - TS(GXT_AS + i), cost);
      }
      for (int r = 0; r < nrow; ++r) {
        const gxt_f4 rq = RQ(r);
        const float s = gxt_rowdot(jrows, stride, r, &XS(0), ms, nv, -rq.y);
        RQ(r).z = s;
        if (s < 0.0f) cost = gx_fma(0.5f * rq.x * s, s, cost);
      }
      if (cnd == 0) cost_s = cost;
      else if (cnd == 1 && cost < cost_s) break;
    }
    bool finish = false;
    for (int trip = 0; trip < m.max_iter + 2; ++trip) {
      // forces and gradient at the current point
      float gn = 0.f;
      // active rows (J qacc - aref < 0): compact index list, then qfrc_constraint = J^T f row by row
      int nact = 0;
      for (int r = 0; r < nrow; r += 4) {
        const float j0 = RQ(r).z;
        const float j1 = r + 1 < nrow ? RQ(r + 1).z : 0.f;
        const float j2 = r + 2 < nrow ? RQ(r + 2).z : 0.f;
        const float j3 = r + 3 < nrow ? RQ(r + 3).z : 0.f;
        if (j0 < 0.0f) { TS(GXT_AL + nact) = (float)r; ++nact; }
        if (j1 < 0.0f) { TS(GXT_AL + nact) = (float)(r + 1); ++nact; }
        if (j2 < 0.0f) { TS(GXT_AL + nact) = (float)(r + 2); ++nact; }
        if (j3 < 0.0f) { TS(GXT_AL + nact) = (float)(r + 3); ++nact; }
      }
      for (int i = 0; i < nv; ++i) FCS(i) = 0.f;
      for (int k = 0; k < nact; ++k) {
        const int r = (int)TS(GXT_AL + k);
        float row[GX_TREE_MAXV];
        gxt_row_load(jrows, stride, r, row);
        const gxt_f4 rq = RQ(r);
        const float force = -rq.x * rq.z;
#pragma unroll
        for (int i = 0; i < GX_TREE_MAXV; ++i) if (i < nv) FCS(i) = gx_fma(row[i], force, FCS(i));
      }
      for (int i = 0; i < nv; ++i) {
        const float g = FCS(i) - (TS(GXT_MA + i) - TS(GXT_FS + i));   // -gradient
        XS(i) = g;                                                    // right-hand side of the Newton solve
        gn = gx_fma(g, g, gn);
      }
      if (finish || !(gx_sqrt(gn) >= m.tol_scale) || iters >= m.max_iter) break;
      // H = M + J^T D J over the active rows; search = H^-1 (-grad)
      gxt_copy_tri(mm, mms, sm, ms, tri);
      for (int k = 0; k < nact; ++k) {
        const int r = (int)TS(GXT_AL + k);
        float row[GX_TREE_MAXV];
        gxt_row_load(jrows, stride, r, row);
        const float Dr = RQ(r).x;
#pragma unroll
        for (int i = 0; i < GX_TREE_MAXV; ++i) {
          if (i < nv) {
            const float di = Dr * row[i];
            if (di != 0.0f) {
              float* hrow = sm + (size_t)(i * (i + 1) / 2) * ms;
#pragma unroll
              for (int j = 0; j <= i; ++j) hrow[(size_t)j * ms] = gx_fma(di, row[j], hrow[(size_t)j * ms]);
            }
          }
        }
      }
      gxt_chol(sm, ms, nv);
      gxt_chol_solve(sm, ms, nv);
      // line-search quadratics
      float g0 = 0.f, h0 = 0.f, ssq = 0.f;
      for (int i = 0; i < nv; ++i) {
        const float s = gxt_symrow(mm, mms, sm + (size_t)GXT_XS * ms, ms, nv, i);
        TS(GXT_MV + i) = s;
        const float si = XS(i);
        g0 = gx_fma(si, TS(GXT_MA + i) - TS(GXT_FS + i), g0);
        h0 = gx_fma(si, s, h0);
        ssq = gx_fma(si, si, ssq);
      }
      for (int r = 0; r < nrow; ++r) RQ(r).w = gxt_rowdot(jrows, stride, r, &XS(0), ms, nv, 0.f);
      float alpha = 0.f, lo = 0.f, hi = 3.0e38f, slope0 = 0.f;
      for (int ls = 0; ls < 40; ++ls) {
        float d1 = gx_fma(alpha, h0, g0), d2 = h0;
        for (int r = 0; r < nrow; ++r) {
          const gxt_f4 rq = RQ(r);
          const float jv = rq.w, x = gx_fma(alpha, jv, rq.z);
          if (x < 0.0f) { const float dj = rq.x * jv; d1 = gx_fma(dj, x, d1); d2 = gx_fma(dj, jv, d2); }
        }
        if (ls == 0) slope0 = d1;
        if (!(d1 < 0.0f || d1 > 0.0f)) break;
        if (d1 < 0.0f) lo = alpha; else hi = alpha;
        float nxt = alpha - d1 * gx_rcp(d2);
        const bool bad = !(nxt > lo && nxt < hi);
        if (bad) nxt = hi < 3.0e38f ? 0.5f * (lo + hi) : 2.0f * gx_max(lo, 1.0f);
        int changed = bad ? 1 : 0;
        for (int r = 0; r < nrow && !changed; ++r) {
          const gxt_f4 rq = RQ(r);
          const float jv = rq.w, jr = rq.z;
          changed |= (gx_fma(alpha, jv, jr) < 0.0f) != (gx_fma(nxt, jv, jr) < 0.0f);
        }
        alpha = nxt;
        if (!changed) break;
      }
      float qq = 0.f;
      for (int i = 0; i < nv; ++i) {
        const float qa = gx_fma(alpha, XS(i), TS(GXT_QA + i));
        TS(GXT_QA + i) = qa;
        TS(GXT_MA + i) = gx_fma(alpha, TS(GXT_MV + i), TS(GXT_MA + i));
        qq = gx_fma(qa, qa, qq);
      }
      for (int r = 0; r < nrow; ++r) { const gxt_f4 rq = RQ(r); RQ(r).z = gx_fma(alpha, rq.w, rq.z); }
      ++iters;
      const float improvement = -0.5f * alpha * slope0;   // decrease predicted by the line-search model (see gx_art.cuh)
      if (!(improvement >= m.tol_scale) || !(alpha * alpha * ssq > 1e-12f * qq)) finish = true;
    }
  }
  if (iters_out) *iters_out = iters;
  for (int i = 0; i < nv; ++i) wo[(size_t)i * ss] = TS(GXT_QA + i);
  if (forward_only) return;
  // ---------------- euler: (M + dt*damping) qacc = qfrc_smooth + qfrc_constraint ----------------
  gxt_copy_tri(mm, mms, sm, ms, tri);
  for (int i = 0; i < nv; ++i) {
    HH(i, i) = gx_fma(m.dt, m.dof[i].damping, HH(i, i));
    XS(i) = TS(GXT_FS + i) + FCS(i);
  }
  gxt_chol(sm, ms, nv);
  gxt_chol_solve(sm, ms, nv);
  for (int i = 0; i < nv; ++i) vo[(size_t)i * ss] = gx_fma(m.dt, XS(i), v[(size_t)i * ss]);
  for (int j = 0; j < m.njnt; ++j) {
    const GxTreeJoint& Jn = m.jnt[j];
    if (Jn.type == GX_TJ_FREE) {
      const int a = Jn.qposadr, i = Jn.dofadr;
      for (int e = 0; e < 3; ++e) qo[(size_t)(a + e) * ss] = gx_fma(m.dt, vo[(size_t)(i + e) * ss], q[(size_t)(a + e) * ss]);
      float w[3] = {vo[(size_t)(i + 3) * ss], vo[(size_t)(i + 4) * ss], vo[(size_t)(i + 5) * ss]};
      const float wn = gx_sqrt(gx_dot3(w, w));
      float ql[4] = {1.f, 0.f, 0.f, 0.f}, qc[4], qn[4];
      if (wn > 0.0f) {
        float s, cs;
        gx_sincos(0.5f * wn * m.dt, &s, &cs);
        const float iw = gx_rcp(wn);
        ql[0] = cs; ql[1] = w[0] * iw * s; ql[2] = w[1] * iw * s; ql[3] = w[2] * iw * s;
      }
      for (int e = 0; e < 4; ++e) qc[e] = q[(size_t)(a + 3 + e) * ss];
      gx_quat_mul(qc, ql, qn);
      const float nr = gx_rsqrt(qn[0] * qn[0] + qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3]);
      for (int e = 0; e < 4; ++e) qo[(size_t)(a + 3 + e) * ss] = qn[e] * nr;
    } else {
      qo[(size_t)Jn.qposadr * ss] = gx_fma(m.dt, vo[(size_t)Jn.dofadr * ss], q[(size_t)Jn.qposadr * ss]);
    }
  }
}

#undef TS
#undef MM
#undef HH
#endif  // GX_TREE_CUH_
