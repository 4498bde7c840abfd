// gx_art.cuh -- mjx.step for articulated planar-root robots (Ant, Walker, Swimmer),
// written for a STAR topology: one root body (planar root joints) and NLEG serial
// chains ("legs") hanging off it.  Every robot the reference's JAX Engine runs has
// this shape (xmls/{swimmer,ant,walker}.xml; Doggo too, with a free-joint root).
//
// What the reference does (engine.py:668 -> mjx.step, third party, absent; restated in
// oracle/mjx_numpy.py from the published pipeline) on a dense 29..31-DoF system -- 18 of
// the DoF being the static goal/hazard slides World.build adds -- is done here on the
// robot block only, with LANES threads per env, lane l owning leg l:
//   * kinematics, composite inertias, velocities and RNE forces run leg-parallel, in a
//     common frame centred on the root body (fp32-safe: offsets <= 0.4 m);
//   * the mass matrix has ARROWHEAD structure (root block + one block per leg + root-leg
//     couplings, no leg-leg terms); so does the Newton Hessian H = M + J^T D J, because
//     a leg's limit/contact rows touch only the root and that leg.  Every linear solve
//     (qacc_smooth, each Newton direction, the implicit-damping Euler solve) is a
//     per-lane Cholesky of the leg block plus a 3x3 Schur complement assembled with 9
//     sub-warp shuffle reductions -- no dense nv x nv factorisation, no shared memory;
//   * constraint rows live in the registers of the lane that owns them; the solver's
//     scalars (cost, gradient norm, line-search derivatives) are butterfly-reduced over
//     the env's lanes, which leaves bit-identical values on every lane, so all lanes of
//     an env take the same branches.
// Solver: Newton with warm start and an exact line search on the piecewise-quadratic
// 1-D cost (same algorithm as the oracle; see oracle/mjx_numpy.py for why MJX's
// bracketing search is not reproduced).
//
// The header also compiles as plain C++ (LANES = 1, reductions are identities): the
// tests build it with g++ to compare the arithmetic with the numpy oracle before any
// GPU time is spent (tests/emu/).  That build is test scaffolding, never shipped.
#ifndef GX_ART_CUH_
#define GX_ART_CUH_

#include <stdint.h>

#include "gx_math.h"

#include "../../include/guardx_art.h"

struct GxArtDev {
  const GxArtModel* m;  // device copy of the model
  float* aq;            // joint state SoA [3 * nv, cap]: qpos | qvel | qacc_warmstart
  int nv, nu;
};

// ---------------------------------------------------------------------------
// lane collectives over the LANES threads of one env
// ---------------------------------------------------------------------------
template <int LANES>
GX_HD float gx_art_sum(float x, unsigned mask) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = LANES >> 1; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
#else
  (void)mask;
#endif
  return x;
}
template <int LANES>
GX_HD int gx_art_or(int x, unsigned mask) {
#if defined(__CUDA_ARCH__)
#pragma unroll
  for (int o = LANES >> 1; o > 0; o >>= 1) x |= __shfl_xor_sync(mask, x, o);
#else
  (void)mask;
#endif
  return x;
}

// Correctly rounded reciprocal / reciprocal square root in ONE instruction sequence of ~4 ops
// (MUFU + Newton fix-up) instead of an IEEE division (~10 ops) or sqrt followed by a division.
// The articulated path is compared with its oracle under a tolerance, so the host build may use
// the plain expressions.
GX_HD float gx_rcp(float x) {
#if defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
GX_HD float gx_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
  return __frsqrt_rn(x);
#else
  return 1.0f / __builtin_sqrtf(x);
#endif
}

// ---------------------------------------------------------------------------
// small vector / spatial algebra (motion vectors [ang; lin], forces [torque; force],
// both about the env's reference point, world axes)
// ---------------------------------------------------------------------------
GX_HD void gx_cross(const float* a, const float* b, float* o) {
  o[0] = a[1] * b[2] - a[2] * b[1];
  o[1] = a[2] * b[0] - a[0] * b[2];
  o[2] = a[0] * b[1] - a[1] * b[0];
}
GX_HD float gx_dot3(const float* a, const float* b) { return gx_fma(a[2], b[2], gx_fma(a[1], b[1], a[0] * b[0])); }
GX_HD float gx_dot6(const float* a, const float* b) { return gx_dot3(a, b) + gx_dot3(a + 3, b + 3); }
GX_HD void gx_quat_mul(const float* a, const float* b, float* o) {
  o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}
GX_HD void gx_quat2mat(const float* q, float* R) {
  const float w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2.0f * (x * y - w * z); R[2] = 2.0f * (x * z + w * y);
  R[3] = 2.0f * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2.0f * (y * z - w * x);
  R[6] = 2.0f * (x * z - w * y); R[7] = 2.0f * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
GX_HD void gx_rot(const float* R, const float* v, float* o) {
  o[0] = gx_fma(R[2], v[2], gx_fma(R[1], v[1], R[0] * v[0]));
  o[1] = gx_fma(R[5], v[2], gx_fma(R[4], v[1], R[3] * v[0]));
  o[2] = gx_fma(R[8], v[2], gx_fma(R[7], v[1], R[6] * v[0]));
}
// spatial inertia I = {Ixx Iyy Izz Ixy Ixz Iyz, m*d (3), m} about the reference point
GX_HD void gx_inert_mul(const float* I, const float* v, float* f) {
  float t[3];
  gx_cross(I + 6, v + 3, t);
  f[0] = gx_fma(I[4], v[2], gx_fma(I[3], v[1], I[0] * v[0])) + t[0];
  f[1] = gx_fma(I[5], v[2], gx_fma(I[1], v[1], I[3] * v[0])) + t[1];
  f[2] = gx_fma(I[2], v[2], gx_fma(I[5], v[1], I[4] * v[0])) + t[2];
  gx_cross(I + 6, v, t);
  f[3] = I[9] * v[3] - t[0];
  f[4] = I[9] * v[4] - t[1];
  f[5] = I[9] * v[5] - t[2];
}
GX_HD void gx_cross_motion(const float* v, const float* w, float* o) {
  float a[3], b[3];
  gx_cross(v, w, o);
  gx_cross(v, w + 3, a);
  gx_cross(v + 3, w, b);
  o[3] = a[0] + b[0]; o[4] = a[1] + b[1]; o[5] = a[2] + b[2];
}
GX_HD void gx_cross_force(const float* v, const float* f, float* o) {
  float a[3], b[3];
  gx_cross(v, f, a);
  gx_cross(v + 3, f + 3, b);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  gx_cross(v, f + 3, o + 3);
}

// solver impedance (mj_makeImpedance / mjx constraint._kbi), power == 2 is the common case
GX_HD float gx_art_imp(const float* s /*dmin dmax width mid power*/, float pos) {
  const float x = gx_div(gx_abs(pos), s[2]);
  if (x > 1.0f) return s[1];
  float y;
  if (s[4] == 2.0f) y = x < s[3] ? gx_div(x * x, s[3]) : 1.0f - gx_div((1.0f - x) * (1.0f - x), 1.0f - s[3]);
  else if (x < s[3]) y = gx_div(1.0f, powf(s[3], s[4] - 1.0f)) * powf(x, s[4]);
  else y = 1.0f - gx_div(1.0f, powf(1.0f - s[3], s[4] - 1.0f)) * powf(1.0f - x, s[4]);
  return gx_clip(s[0] + y * (s[1] - s[0]), s[0], s[1]);
}

// ---------------------------------------------------------------------------
// per-segment work arrays (registers after unrolling)
// ---------------------------------------------------------------------------
template <int K, int NB>
struct GxArtSegWork {
  float bp[NB][3], bR[NB][9];      // body poses
  float anchor[K][3], axis[K][3];  // joint anchors / axes, world
  float cdof[K][6];
  float crb[NB][10];               // body spatial inertia, then composite (suffix sum)
  float cfrc[NB][6];               // body RNE force, then accumulated (suffix sum)
};

// kinematics of one serial segment: bodies in order, joints applied in order (mjx
// smooth.kinematics).  (pp, pq): parent position / quaternion; on return the pose of
// the last body.
template <int K, int NB>
GX_HD void gx_art_seg_kin(const GxArtSeg& sg, const float* q, float* pp, float* pq, GxArtSegWork<K, NB>& w) {
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b < sg.nb) {
      const GxArtBody& B = sg.body[b];
      float R[9], t[3], pos[3], quat[4];
      gx_quat2mat(pq, R);
      gx_rot(R, B.pos, t);
      pos[0] = pp[0] + t[0]; pos[1] = pp[1] + t[1]; pos[2] = pp[2] + t[2];
      gx_quat_mul(pq, B.quat, quat);
      // every (body, dof) pair is unrolled with a predicate so that all indices are
      // compile-time constants and the work arrays stay in registers
#pragma unroll
      for (int di = 0; di < K; ++di) {
        if (di < sg.nk && sg.dofbody[di] == b) {
          const GxArtDof& D = sg.dof[di];
          gx_quat2mat(quat, R);
          gx_rot(R, D.jpos, t);
          w.anchor[di][0] = pos[0] + t[0]; w.anchor[di][1] = pos[1] + t[1]; w.anchor[di][2] = pos[2] + t[2];
          gx_rot(R, D.axis, w.axis[di]);
          const float dq = q[di] - D.qpos0;
          if (D.type == GX_ART_SLIDE) {
            pos[0] = gx_fma(w.axis[di][0], dq, pos[0]);
            pos[1] = gx_fma(w.axis[di][1], dq, pos[1]);
            pos[2] = gx_fma(w.axis[di][2], dq, pos[2]);
          } else {
            float s, c, ql[4], qn[4];
            gx_sincos(0.5f * dq, &s, &c);
            ql[0] = c; ql[1] = D.axis[0] * s; ql[2] = D.axis[1] * s; ql[3] = D.axis[2] * s;
            gx_quat_mul(quat, ql, qn);
            quat[0] = qn[0]; quat[1] = qn[1]; quat[2] = qn[2]; quat[3] = qn[3];
            gx_quat2mat(quat, R);
            gx_rot(R, D.jpos, t);
            pos[0] = w.anchor[di][0] - t[0]; pos[1] = w.anchor[di][1] - t[1]; pos[2] = w.anchor[di][2] - t[2];
          }
        }
      }
      const float nrm = gx_rsqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
      quat[0] *= nrm; quat[1] *= nrm; quat[2] *= nrm; quat[3] *= nrm;
      gx_quat2mat(quat, w.bR[b]);
      w.bp[b][0] = pos[0]; w.bp[b][1] = pos[1]; w.bp[b][2] = pos[2];
      pp[0] = pos[0]; pp[1] = pos[1]; pp[2] = pos[2];
      pq[0] = quat[0]; pq[1] = quat[1]; pq[2] = quat[2]; pq[3] = quat[3];
    }
  }
}

// cdof, spatial inertias, velocities and RNE body forces of one segment about the
// reference point c (mjx smooth.com_pos / com_vel / rne, restricted to a chain).
// (cvel, cacc): parent's spatial velocity / acceleration; on return the last body's.
template <int K, int NB>
GX_HD void gx_art_seg_dyn(const GxArtSeg& sg, const float* c, const float* v, float* cvel, float* cacc,
                          GxArtSegWork<K, NB>& w) {
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b < sg.nb) {
      const GxArtBody& B = sg.body[b];
#pragma unroll
      for (int di = 0; di < K; ++di) {
        if (di < sg.nk && sg.dofbody[di] == b) {
          float* cd = w.cdof[di];
          if (sg.dof[di].type == GX_ART_SLIDE) {
            cd[0] = 0.f; cd[1] = 0.f; cd[2] = 0.f;
            cd[3] = w.axis[di][0]; cd[4] = w.axis[di][1]; cd[5] = w.axis[di][2];
          } else {
            float off[3] = {c[0] - w.anchor[di][0], c[1] - w.anchor[di][1], c[2] - w.anchor[di][2]};
            cd[0] = w.axis[di][0]; cd[1] = w.axis[di][1]; cd[2] = w.axis[di][2];
            gx_cross(w.axis[di], off, cd + 3);
          }
          float dd[6];
          gx_cross_motion(cvel, cd, dd);  // cdof_dot
          const float vi = v[di];
#pragma unroll
          for (int k = 0; k < 6; ++k) { cacc[k] = gx_fma(dd[k], vi, cacc[k]); }
#pragma unroll
          for (int k = 0; k < 6; ++k) { cvel[k] = gx_fma(cd[k], vi, cvel[k]); }
        }
      }
      // spatial inertia about c: R I R^T + m (|d|^2 1 - d d^T), m d, m
      float* I = w.crb[b];
      const float* R = w.bR[b];
      float t[3], d[3];
      gx_rot(R, B.ipos, t);
      d[0] = w.bp[b][0] + t[0] - c[0]; d[1] = w.bp[b][1] + t[1] - c[1]; d[2] = w.bp[b][2] + t[2] - c[2];
      // A = R * Ib (Ib symmetric), Iw = A * R^T
      const float* J = B.inert;
      float A[9];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        A[3 * r + 0] = gx_fma(R[3 * r + 2], J[4], gx_fma(R[3 * r + 1], J[3], R[3 * r] * J[0]));
        A[3 * r + 1] = gx_fma(R[3 * r + 2], J[5], gx_fma(R[3 * r + 1], J[1], R[3 * r] * J[3]));
        A[3 * r + 2] = gx_fma(R[3 * r + 2], J[2], gx_fma(R[3 * r + 1], J[5], R[3 * r] * J[4]));
      }
      const float m = B.mass, dd2 = gx_dot3(d, d);
      I[0] = gx_dot3(A + 0, R + 0) + m * (dd2 - d[0] * d[0]);
      I[1] = gx_dot3(A + 3, R + 3) + m * (dd2 - d[1] * d[1]);
      I[2] = gx_dot3(A + 6, R + 6) + m * (dd2 - d[2] * d[2]);
      I[3] = gx_dot3(A + 0, R + 3) - m * d[0] * d[1];
      I[4] = gx_dot3(A + 0, R + 6) - m * d[0] * d[2];
      I[5] = gx_dot3(A + 3, R + 6) - m * d[1] * d[2];
      I[6] = m * d[0]; I[7] = m * d[1]; I[8] = m * d[2]; I[9] = m;
      // RNE body force: I a + v x* (I v)
      float Ia[6], Iv[6], x[6];
      gx_inert_mul(I, cacc, Ia);
      gx_inert_mul(I, cvel, Iv);
      gx_cross_force(cvel, Iv, x);
#pragma unroll
      for (int k = 0; k < 6; ++k) w.cfrc[b][k] = Ia[k] + x[k];
    }
  }
  // suffix sums along the chain: composite inertia and accumulated force
#pragma unroll
  for (int b = NB - 2; b >= 0; --b) {
    if (b + 1 < sg.nb) {
#pragma unroll
      for (int k = 0; k < 10; ++k) w.crb[b][k] += w.crb[b + 1][k];
#pragma unroll
      for (int k = 0; k < 6; ++k) w.cfrc[b][k] += w.cfrc[b + 1][k];
    }
  }
}

// passive + actuator force of one dof (mjx passive, fwd_actuation; joint transmission)
GX_HD float gx_art_dof_force(const GxArtDof& D, float q, float v, float ctrl) {
  float f = -(D.damping * v) - D.stiffness * (q - D.springref);
  if (D.act >= 0) {
    float u = ctrl;
    if (D.ctrllimited) u = gx_clip(u, D.ctrl_lo, D.ctrl_hi);
    float force = D.gain * u + D.biasv * (D.gear * v);
    if (D.forcelimited) force = gx_clip(force, D.force_lo, D.force_hi);
    f += D.gear * force;
  }
  return f;
}

// in-place Cholesky (lower) of the leading n x n block of a K x K matrix; rows >= n are
// treated as identity so that the fixed-size loops stay harmless.  The diagonal is stored
// INVERTED (1 / L_ii): every later use of it is a division, and an IEEE division is ~10
// instructions on this machine -- ncu counted ~460 of them per env-step before this.
template <int K>
GX_HD void gx_art_chol(float (*A)[K], int n) {
#pragma unroll
  for (int j = 0; j < K; ++j) {
    if (j < n) {
      float d = A[j][j];
#pragma unroll
      for (int k = 0; k < K; ++k) if (k < j) d = gx_fma(-A[j][k], A[j][k], d);
      const float inv = gx_rsqrt(d);
      A[j][j] = inv;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        if (i > j && i < n) {
          float s = A[i][j];
#pragma unroll
          for (int k = 0; k < K; ++k) if (k < j) s = gx_fma(-A[i][k], A[j][k], s);
          A[i][j] = s * inv;
        }
      }
    }
  }
}
template <int K>
GX_HD void gx_art_chol_solve(const float (*L)[K], int n, float* x) {
#pragma unroll
  for (int i = 0; i < K; ++i) {
    if (i < n) {
      float s = x[i];
#pragma unroll
      for (int k = 0; k < K; ++k) if (k < i) s = gx_fma(-L[i][k], x[k], s);
      x[i] = s * L[i][i];
    }
  }
#pragma unroll
  for (int i = K - 1; i >= 0; --i) {
    if (i < n) {
      float s = x[i];
#pragma unroll
      for (int k = 0; k < K; ++k) if (k > i && k < n) s = gx_fma(-L[k][i], x[k], s);
      x[i] = s * L[i][i];
    }
  }
}

// ---------------------------------------------------------------------------
// one env, one lane
// ---------------------------------------------------------------------------
template <int LANES, int NLEG, int K, int NB>
struct GxArtLane {
  static constexpr int LPL = NLEG / LANES;  // legs per lane
  static_assert(NLEG % LANES == 0, "legs must divide over lanes");
  // dynamic state: root dofs replicated on every lane, leg dofs on the owning lane
  float qr[3], vr[3], wr[3];
  float ql[LPL][K], vl[LPL][K], wl[LPL][K];
};

template <int K>
struct GxArtRows {     // constraint rows owned by one leg
  float ls[K];         // limit row sign (0 = inactive): J = ls * e_i
  float lD[K], laref[K];
  float cj[3][3 + K];  // contact normal / tangent-1 / tangent-2 Jacobians over (root 3, chain K)
  float cD, caref[4], cmu;
  int con;             // contact active
};

// A (K x K lower incl. diag), C (K x 3), local root contribution Arr (6: 00 10 11 20 21 22):
// the arrowhead system  [Mrr + sum Arr , C^T ; C , A] [xr; xl] = [br; bl]
template <int LANES, int NLEG, int K>
GX_HD void gx_art_arrow_solve(float (*A)[NLEG / LANES][K][K], const float (*C)[K][3], const float (*Arr)[6],
                              const float* Mrr /*6*/, const int* nk, float (*bl)[K], float* br, unsigned mask) {
  constexpr int LPL = NLEG / LANES;
  float Y[LPL][3][K];
  float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
    gx_art_chol<K>((*A)[j], nk[j]);
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int i = 0; i < K; ++i) Y[j][r][i] = i < nk[j] ? C[j][i][r] : 0.0f;
      gx_art_chol_solve<K>((*A)[j], nk[j], Y[j][r]);
    }
    gx_art_chol_solve<K>((*A)[j], nk[j], bl[j]);
    // Schur: S -= C^T Y ; g -= C^T z
    int e = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s <= r; ++s, ++e) {
        float t = Arr ? Arr[j][e] : 0.0f;
#pragma unroll
        for (int i = 0; i < K; ++i) if (i < nk[j]) t = gx_fma(-C[j][i][r], Y[j][s][i], t);
        acc[e] += t;
      }
      float g = 0.0f;
#pragma unroll
      for (int i = 0; i < K; ++i) if (i < nk[j]) g = gx_fma(-C[j][i][r], bl[j][i], g);
      acc[6 + r] += g;
    }
  }
#pragma unroll
  for (int e = 0; e < 9; ++e) acc[e] = gx_art_sum<LANES>(acc[e], mask);
  float S[3][3];
  S[0][0] = Mrr[0] + acc[0];
  S[1][0] = Mrr[1] + acc[1]; S[1][1] = Mrr[2] + acc[2];
  S[2][0] = Mrr[3] + acc[3]; S[2][1] = Mrr[4] + acc[4]; S[2][2] = Mrr[5] + acc[5];
  S[0][1] = S[0][2] = S[1][2] = 0.0f;
  br[0] += acc[6]; br[1] += acc[7]; br[2] += acc[8];
  gx_art_chol<3>(S, 3);
  gx_art_chol_solve<3>(S, 3, br);
#pragma unroll
  for (int j = 0; j < LPL; ++j)
#pragma unroll
    for (int i = 0; i < K; ++i)
      if (i < nk[j]) bl[j][i] = bl[j][i] - (Y[j][0][i] * br[0] + Y[j][1][i] * br[1] + Y[j][2][i] * br[2]);
}

struct GxArtSnap {     // what the env logic reads from the forward pass (PRE-step pose)
  float px, py, r00, r10;
};

// One mjx.step.  ctrl: this lane's actuator inputs, cr for root dofs, cl per leg dof.
// commit: write the post-step q, v back into `st` (false = reset_done's throw-away step);
// `qo`/`vo` always receive the post-step values of this lane's dofs (root: qro/vro).
// forward_only: skip the integration (mjx.forward at reset): q, v unchanged.
// The solver's qacc is always stored as the new warm start unless keep_warm.
template <int LANES, int NLEG, int K, int NB>
GX_HD void gx_art_step(const GxArtModel& m, int lane, unsigned mask, GxArtLane<LANES, NLEG, K, NB>& st,
                       const float* cr, const float (*cl)[K], bool commit, bool forward_only, bool keep_warm,
                       GxArtSnap& snap, float* qro, float* vro, float (*qlo)[K], float (*vlo)[K], int* iters_out) {
  constexpr int LPL = NLEG / LANES;
  // ---------------- root kinematics (replicated) ----------------
  GxArtSegWork<3, 1> rw;
  float pp[3] = {0.f, 0.f, 0.f}, pq[4] = {1.f, 0.f, 0.f, 0.f};
  gx_art_seg_kin<3, 1>(m.root, st.qr, pp, pq, rw);
  const float c[3] = {pp[0], pp[1], pp[2]};
  snap.px = c[0]; snap.py = c[1]; snap.r00 = rw.bR[0][0]; snap.r10 = rw.bR[0][3];
  float cvel[6] = {0, 0, 0, 0, 0, 0}, cacc[6] = {0, 0, 0, 0, 0, -m.grav_z};
  gx_art_seg_dyn<3, 1>(m.root, c, st.vr, cvel, cacc, rw);
  // ---------------- legs ----------------
  GxArtSegWork<K, NB> lw[LPL];
  int nk[LPL];
  float cp_pose[LPL][12];  // pose of the contact body
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
    const int leg = j * LANES + lane;
    const GxArtSeg& sg = m.leg[leg];
    nk[j] = sg.nk;
    float lp[3] = {pp[0], pp[1], pp[2]}, lq[4] = {pq[0], pq[1], pq[2], pq[3]};
    gx_art_seg_kin<K, NB>(sg, st.ql[j], lp, lq, lw[j]);
    float lv[6], la[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) { lv[k] = cvel[k]; la[k] = cacc[k]; }
    gx_art_seg_dyn<K, NB>(sg, c, st.vl[j], lv, la, lw[j]);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b == sg.con.body) {
#pragma unroll
        for (int k = 0; k < 3; ++k) cp_pose[j][k] = lw[j].bp[b][k];
#pragma unroll
        for (int k = 0; k < 9; ++k) cp_pose[j][3 + k] = lw[j].bR[b][k];
      }
  }
  // ---------------- root composite inertia / force: own + sum over legs ----------------
  float rcrb[10], rfrc[6];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < LPL; ++j) s += lw[j].crb[0][k];
    rcrb[k] = rw.crb[0][k] + gx_art_sum<LANES>(s, mask);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < LPL; ++j) s += lw[j].cfrc[0][k];
    rfrc[k] = rw.cfrc[0][k] + gx_art_sum<LANES>(s, mask);
  }
  // ---------------- mass matrix (arrowhead) and smooth forces ----------------
  float Mrr[6], fr[3];
  {
    int e = 0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float f[6];
      gx_inert_mul(rcrb, rw.cdof[r], f);
#pragma unroll
      for (int s = 0; s <= r; ++s, ++e) Mrr[e] = gx_dot6(rw.cdof[s], f);
      Mrr[e - 1] += m.root.dof[r].armature;
      const float bias = gx_dot6(rw.cdof[r], rfrc);
      fr[r] = gx_art_dof_force(m.root.dof[r], st.qr[r], st.vr[r], cr ? cr[r] : 0.0f) - bias;
    }
  }
  float Mll[LPL][K][K], C[LPL][K][3], fl[LPL][K];
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
    const GxArtSeg& sg = m.leg[j * LANES + lane];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      if (i < sg.nk) {
        float f[6];
        const int b = sg.dofbody[i];
        // crb[b] selected by a compile-time-unrolled scan (keeps the arrays in registers)
        float I[10];
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) if (bb == b) {
#pragma unroll
          for (int k = 0; k < 10; ++k) I[k] = lw[j].crb[bb][k];
        }
        gx_inert_mul(I, lw[j].cdof[i], f);
#pragma unroll
        for (int s = 0; s < K; ++s) Mll[j][i][s] = (s <= i) ? gx_dot6(lw[j].cdof[s], f) : 0.0f;
        Mll[j][i][i] += sg.dof[i].armature;
#pragma unroll
        for (int r = 0; r < 3; ++r) C[j][i][r] = gx_dot6(rw.cdof[r], f);
        float F[6];
#pragma unroll
        for (int bb = 0; bb < NB; ++bb) if (bb == b) {
#pragma unroll
          for (int k = 0; k < 6; ++k) F[k] = lw[j].cfrc[bb][k];
        }
        fl[j][i] = gx_art_dof_force(sg.dof[i], st.ql[j][i], st.vl[j][i], cl ? cl[j][i] : 0.0f) - gx_dot6(lw[j].cdof[i], F);
      } else {
#pragma unroll
        for (int s = 0; s < K; ++s) Mll[j][i][s] = (s == i) ? 1.0f : 0.0f;
#pragma unroll
        for (int r = 0; r < 3; ++r) C[j][i][r] = 0.0f;
        fl[j][i] = 0.0f;
      }
    }
  }
  // ---------------- constraint rows (mjx make_constraint) ----------------
  GxArtRows<K> rows[LPL];
  int any_row = 0;
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
    const GxArtSeg& sg = m.leg[j * LANES + lane];
    GxArtRows<K>& R = rows[j];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      R.ls[i] = 0.0f; R.lD[i] = 0.0f; R.laref[i] = 0.0f;
      if (i < sg.nk && sg.dof[i].limited) {
        const GxArtDof& D = sg.dof[i];
        const float q = st.ql[j][i];
        const float dmin = q - D.range[0], dmax = D.range[1] - q;
        const float pos = gx_min(dmin, dmax) - D.margin;
        if (pos < 0.0f) {
          const float sgn = dmin < dmax ? 1.0f : -1.0f;
          const float imp = gx_art_imp(D.imp, pos);
          const float Rr = gx_max(gx_div(D.invw * (1.0f - imp), imp), GX_ART_MINVAL);
          R.ls[i] = sgn;
          R.lD[i] = gx_rcp(Rr);
          R.laref[i] = -D.b * (sgn * st.vl[j][i]) - D.k * imp * pos;
          any_row = 1;
        }
      }
    }
    R.con = 0; R.cD = 0.0f; R.cmu = sg.con.mu;
#pragma unroll
    for (int k = 0; k < 4; ++k) R.caref[k] = 0.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int i = 0; i < 3 + K; ++i) R.cj[a][i] = 0.0f;
    if (sg.con.on) {
      float t[3], sp[3];
      gx_rot(cp_pose[j] + 3, sg.con.pos, t);
      sp[0] = cp_pose[j][0] + t[0]; sp[1] = cp_pose[j][1] + t[1]; sp[2] = cp_pose[j][2] + t[2];
      const float dist = sp[2] - sg.con.radius;      // floor: z = 0, normal +z
      const float pos = dist - sg.con.margin;
      if (pos < 0.0f) {
        // contact point and its offset from the reference point
        const float off[3] = {sp[0] - c[0], sp[1] - c[1], (sp[2] - (sg.con.radius + 0.5f * dist)) - c[2]};
        // frame: normal z, tangents y and z x y = -x
        float vel[3] = {0.f, 0.f, 0.f};
        auto col = [&](const float* cd, int idx, float qv) {
          float x[3];
          gx_cross(cd, off, x);
          const float jx = cd[3] + x[0], jy = cd[4] + x[1], jz = cd[5] + x[2];
          R.cj[0][idx] = jz; R.cj[1][idx] = jy; R.cj[2][idx] = -jx;
          vel[0] = gx_fma(jz, qv, vel[0]); vel[1] = gx_fma(jy, qv, vel[1]); vel[2] = gx_fma(-jx, qv, vel[2]);
        };
#pragma unroll
        for (int r = 0; r < 3; ++r) col(rw.cdof[r], r, st.vr[r]);
#pragma unroll
        for (int i = 0; i < K; ++i) if (i < sg.nk && sg.dofbody[i] <= sg.con.body) col(lw[j].cdof[i], 3 + i, st.vl[j][i]);
        const float imp = gx_art_imp(sg.con.imp, pos);
        const float Rr = gx_max(gx_div(sg.con.invw * (1.0f - imp), imp), GX_ART_MINVAL);
        R.cD = gx_rcp(Rr);
        const float kp = sg.con.k * imp * pos, mu = sg.con.mu;
        R.caref[0] = -sg.con.b * (vel[0] + mu * vel[1]) - kp;
        R.caref[1] = -sg.con.b * (vel[0] - mu * vel[1]) - kp;
        R.caref[2] = -sg.con.b * (vel[0] + mu * vel[2]) - kp;
        R.caref[3] = -sg.con.b * (vel[0] - mu * vel[2]) - kp;
        R.con = 1;
        any_row = 1;
      }
    }
  }
  any_row = gx_art_or<LANES>(any_row, mask);
  // ---------------- one phase loop for every linear solve ----------------
  // qacc_smooth = M^-1 qfrc_smooth, then (if any constraint row is active) solver.solve:
  // Newton on the primal cost with warm start, then euler's implicit-damping solve
  // (M + dt*damping) qacc = qfrc_smooth + qfrc_constraint.  All three are "assemble an
  // arrowhead matrix, factor, solve"; keeping them in ONE non-unrolled loop keeps ONE copy
  // of the assembly / Cholesky / Schur code in the instruction stream (the fully inlined
  // version was 55 k SASS instructions and stalled 28 % of the time on instruction fetch).
  float sr[3] = {0.f, 0.f, 0.f}, sl[LPL][K];      // qacc_smooth
  float ar[3] = {0.f, 0.f, 0.f}, al[LPL][K];      // qacc
  float fcr[3] = {0.f, 0.f, 0.f}, fcl[LPL][K];    // qfrc_constraint
  float Mar[3] = {0.f, 0.f, 0.f}, Mal[LPL][K], jl[LPL][K], jc[LPL][4];   // M*qacc, J*qacc - aref
  float er[3] = {0.f, 0.f, 0.f}, el[LPL][K];      // euler qacc
#pragma unroll
  for (int j = 0; j < LPL; ++j) {
#pragma unroll
    for (int i = 0; i < K; ++i) { sl[j][i] = 0.f; al[j][i] = 0.f; fcl[j][i] = 0.f; Mal[j][i] = 0.f; jl[j][i] = 0.f; el[j][i] = 0.f; }
#pragma unroll
    for (int k = 0; k < 4; ++k) jc[j][k] = 0.f;
  }
  // M*x for the arrowhead matrix (root part reduced over lanes)
  auto mul_m = [&](const float* xr, const float (*xl)[K], float* yr, float (*yl)[K]) {
    float t[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < LPL; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float s = 0.0f;
        if (i < nk[j]) {
#pragma unroll
          for (int k = 0; k < K; ++k) if (k < nk[j]) s = gx_fma(k <= i ? Mll[j][i][k] : Mll[j][k][i], xl[j][k], s);
#pragma unroll
          for (int r = 0; r < 3; ++r) { s = gx_fma(C[j][i][r], xr[r], s); t[r] = gx_fma(C[j][i][r], xl[j][i], t[r]); }
        }
        yl[j][i] = s;
      }
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = gx_art_sum<LANES>(t[r], mask);
    yr[0] = t[0] + (Mrr[0] * xr[0] + Mrr[1] * xr[1] + Mrr[3] * xr[2]);
    yr[1] = t[1] + (Mrr[1] * xr[0] + Mrr[2] * xr[1] + Mrr[4] * xr[2]);
    yr[2] = t[2] + (Mrr[3] * xr[0] + Mrr[4] * xr[1] + Mrr[5] * xr[2]);
  };
  // J*x for this lane's rows: limits K, contact 4
  auto mul_j = [&](int j, const float* xr, const float* xl, float* ylim, float* ycon) {
    const GxArtRows<K>& R = rows[j];
#pragma unroll
    for (int i = 0; i < K; ++i) ylim[i] = R.ls[i] * xl[i];
    float n = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int r = 0; r < 3; ++r) { n = gx_fma(R.cj[0][r], xr[r], n); t1 = gx_fma(R.cj[1][r], xr[r], t1); t2 = gx_fma(R.cj[2][r], xr[r], t2); }
#pragma unroll
    for (int i = 0; i < K; ++i) { n = gx_fma(R.cj[0][3 + i], xl[i], n); t1 = gx_fma(R.cj[1][3 + i], xl[i], t1); t2 = gx_fma(R.cj[2][3 + i], xl[i], t2); }
    ycon[0] = n + R.cmu * t1; ycon[1] = n - R.cmu * t1; ycon[2] = n + R.cmu * t2; ycon[3] = n - R.cmu * t2;
  };
  enum { PH_SMOOTH = 0, PH_NEWTON = 1, PH_EULER = 2 };
  int mode = PH_SMOOTH, iters = 0;
  bool finish = false;
#pragma unroll 1
  for (int trip = 0; trip < m.max_iter + 4; ++trip) {
    float A[LPL][K][K], Cc[LPL][K][3], Arr[LPL][6], bl[LPL][K], br[3];
    // ---- Newton: constraint forces and gradient at the current qacc
    if (mode == PH_NEWTON) {
      float gr_loc[3] = {0.f, 0.f, 0.f}, gn_loc = 0.0f;
#pragma unroll
      for (int j = 0; j < LPL; ++j) {
        const GxArtRows<K>& R = rows[j];
#pragma unroll
        for (int i = 0; i < K; ++i) fcl[j][i] = (jl[j][i] < 0.0f) ? R.ls[i] * (-R.lD[i] * jl[j][i]) : 0.0f;
        if (R.con) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (jc[j][k] < 0.0f) {
              const float force = -R.cD * jc[j][k];
              const float sg_ = (k & 1) ? -R.cmu : R.cmu;
              const int ta = 1 + (k >> 1);
#pragma unroll
              for (int r = 0; r < 3; ++r) gr_loc[r] = gx_fma(gx_fma(sg_, R.cj[ta][r], R.cj[0][r]), force, gr_loc[r]);
#pragma unroll
              for (int i = 0; i < K; ++i) fcl[j][i] = gx_fma(gx_fma(sg_, R.cj[ta][3 + i], R.cj[0][3 + i]), force, fcl[j][i]);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < K; ++i) {
          bl[j][i] = i < nk[j] ? fcl[j][i] - (Mal[j][i] - fl[j][i]) : 0.0f;   // -gradient
          gn_loc = gx_fma(bl[j][i], bl[j][i], gn_loc);
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        fcr[r] = gx_art_sum<LANES>(gr_loc[r], mask);
        br[r] = fcr[r] - (Mar[r] - fr[r]);
      }
      const float gnorm = gx_sqrt(gx_art_sum<LANES>(gn_loc, mask) + br[0] * br[0] + br[1] * br[1] + br[2] * br[2]);
      // converged, out of iterations, or NaN (singular model state: the env's NaN guard fires)
      if (finish || !(gnorm >= m.tol_scale) || iters >= m.max_iter) {
        if (forward_only) break;
        mode = PH_EULER;
      }
    }
    // ---- assemble the arrowhead system of this phase
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
      const GxArtSeg& sg = m.leg[j * LANES + lane];
      const GxArtRows<K>& R = rows[j];
#pragma unroll
      for (int e = 0; e < 6; ++e) Arr[j][e] = 0.0f;
#pragma unroll
      for (int i = 0; i < K; ++i) {
#pragma unroll
        for (int s2 = 0; s2 < K; ++s2) A[j][i][s2] = Mll[j][i][s2];
#pragma unroll
        for (int r = 0; r < 3; ++r) Cc[j][i][r] = C[j][i][r];
        if (mode == PH_SMOOTH) bl[j][i] = fl[j][i];
        if (mode == PH_EULER) {
          bl[j][i] = fl[j][i] + fcl[j][i];
          if (i < nk[j]) A[j][i][i] = gx_fma(m.dt, sg.dof[i].damping, A[j][i][i]);
        }
        if (mode == PH_NEWTON && jl[j][i] < 0.0f) A[j][i][i] += R.lD[i];   // ls^2 = 1 on active rows
      }
      if (mode == PH_NEWTON && R.con) {
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          if (jc[j][k] < 0.0f) {
            const float sg_ = (k & 1) ? -R.cmu : R.cmu;
            const int ta = 1 + (k >> 1);
            float row[3 + K];
#pragma unroll
            for (int i = 0; i < 3 + K; ++i) row[i] = gx_fma(sg_, ta == 1 ? R.cj[1][i] : R.cj[2][i], R.cj[0][i]);
            int e = 0;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int s2 = 0; s2 <= r; ++s2, ++e) Arr[j][e] = gx_fma(R.cD * row[r], row[s2], Arr[j][e]);
#pragma unroll
            for (int i = 0; i < K; ++i) {
              const float di = R.cD * row[3 + i];
#pragma unroll
              for (int s2 = 0; s2 < K; ++s2) if (s2 <= i) A[j][i][s2] = gx_fma(di, row[3 + s2], A[j][i][s2]);
#pragma unroll
              for (int r = 0; r < 3; ++r) Cc[j][i][r] = gx_fma(di, row[r], Cc[j][i][r]);
            }
          }
        }
      }
      if (mode == PH_EULER && lane == 0 && j == 0) {   // root damping enters the Schur sum once
        Arr[j][0] = m.dt * m.root.dof[0].damping;
        Arr[j][2] = m.dt * m.root.dof[1].damping;
        Arr[j][5] = m.dt * m.root.dof[2].damping;
      }
    }
    if (mode == PH_SMOOTH) { br[0] = fr[0]; br[1] = fr[1]; br[2] = fr[2]; }
    if (mode == PH_EULER) { br[0] = fr[0] + fcr[0]; br[1] = fr[1] + fcr[1]; br[2] = fr[2] + fcr[2]; }
    gx_art_arrow_solve<LANES, NLEG, K>(&A, Cc, Arr, Mrr, nk, bl, br, mask);
    // ---- what the solution is, per phase
    if (mode == PH_EULER) {
#pragma unroll
      for (int r = 0; r < 3; ++r) er[r] = br[r];
#pragma unroll
      for (int j = 0; j < LPL; ++j)
#pragma unroll
        for (int i = 0; i < K; ++i) el[j][i] = bl[j][i];
      break;
    }
    if (mode == PH_SMOOTH) {
#pragma unroll
      for (int r = 0; r < 3; ++r) { sr[r] = br[r]; ar[r] = br[r]; }
#pragma unroll
      for (int j = 0; j < LPL; ++j)
#pragma unroll
        for (int i = 0; i < K; ++i) { sl[j][i] = bl[j][i]; al[j][i] = bl[j][i]; }
      if (!any_row) {
        if (forward_only) break;
        mode = PH_EULER;
        continue;
      }
      // warm start (mjx solver.solve): start from qacc_warmstart iff its cost beats qacc_smooth's.
      // Candidates in the order smooth, warm, smooth-again so that the quantities derived from the
      // winner (M*qacc, J*qacc - aref) are the ones left behind by the last evaluation.  At
      // qacc_smooth the Gauss term vanishes identically and M*qacc = qfrc_smooth, so that candidate
      // needs no matrix-vector product.
      float cost_s = 0.0f;
#pragma unroll 1
      for (int cnd = 0; cnd < 3; ++cnd) {
        const bool usew = cnd == 1;
#pragma unroll
        for (int r = 0; r < 3; ++r) ar[r] = usew ? st.wr[r] : sr[r];
#pragma unroll
        for (int j = 0; j < LPL; ++j)
#pragma unroll
          for (int i = 0; i < K; ++i) al[j][i] = usew ? st.wl[j][i] : sl[j][i];
        if (usew) {
          mul_m(ar, al, Mar, Mal);
        } else {
#pragma unroll
          for (int r = 0; r < 3; ++r) Mar[r] = fr[r];
#pragma unroll
          for (int j = 0; j < LPL; ++j)
#pragma unroll
            for (int i = 0; i < K; ++i) Mal[j][i] = fl[j][i];
        }
        float local = 0.0f;
#pragma unroll
        for (int j = 0; j < LPL; ++j) {
          mul_j(j, ar, al[j], jl[j], jc[j]);
#pragma unroll
          for (int i = 0; i < K; ++i) {
            jl[j][i] -= rows[j].laref[i];
            if (jl[j][i] < 0.0f) local = gx_fma(0.5f * rows[j].lD[i] * jl[j][i], jl[j][i], local);
            local = gx_fma(0.5f * (Mal[j][i] - fl[j][i]), al[j][i] - sl[j][i], local);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            jc[j][k] -= rows[j].caref[k];
            if (rows[j].con && jc[j][k] < 0.0f) local = gx_fma(0.5f * rows[j].cD * jc[j][k], jc[j][k], local);
          }
        }
        float cost = gx_art_sum<LANES>(local, mask);
#pragma unroll
        for (int r = 0; r < 3; ++r) cost = gx_fma(0.5f * (Mar[r] - fr[r]), ar[r] - sr[r], cost);
        if (cnd == 0) cost_s = cost;
        else if (cnd == 1 && cost < cost_s) break;       // warm start wins: its derived quantities are current
      }
      mode = PH_NEWTON;
      continue;
    }
    // ---- Newton: (br, bl) is the search direction; exact line search, then move
    float mvr[3], mvl[LPL][K], jvl[LPL][K], jvc[LPL][4];
    mul_m(br, bl, mvr, mvl);
    float g0 = 0.0f, h0 = 0.0f;
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
      mul_j(j, br, bl[j], jvl[j], jvc[j]);
#pragma unroll
      for (int i = 0; i < K; ++i) {
        g0 = gx_fma(bl[j][i], Mal[j][i] - fl[j][i], g0);
        h0 = gx_fma(bl[j][i], mvl[j][i], h0);
      }
    }
    g0 = gx_art_sum<LANES>(g0, mask);
    h0 = gx_art_sum<LANES>(h0, mask);
#pragma unroll
    for (int r = 0; r < 3; ++r) { g0 = gx_fma(br[r], Mar[r] - fr[r], g0); h0 = gx_fma(br[r], mvr[r], h0); }
    // exact minimiser of the piecewise-quadratic 1-D cost: Newton on its derivative, bracketed
    float alpha = 0.0f, lo = 0.0f, hi = 3.0e38f, slope0 = 0.0f;
#pragma unroll 1
    for (int ls = 0; ls < 40; ++ls) {
      float d1 = 0.0f, d2 = 0.0f;
#pragma unroll
      for (int j = 0; j < LPL; ++j) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          const float x = gx_fma(alpha, jvl[j][i], jl[j][i]);
          if (x < 0.0f) { const float dj = rows[j].lD[i] * jvl[j][i]; d1 = gx_fma(dj, x, d1); d2 = gx_fma(dj, jvl[j][i], d2); }
        }
        if (rows[j].con) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float x = gx_fma(alpha, jvc[j][k], jc[j][k]);
            if (x < 0.0f) { const float dj = rows[j].cD * jvc[j][k]; d1 = gx_fma(dj, x, d1); d2 = gx_fma(dj, jvc[j][k], d2); }
          }
        }
      }
      d1 = gx_art_sum<LANES>(d1, mask) + gx_fma(alpha, h0, g0);
      d2 = gx_art_sum<LANES>(d2, mask) + h0;
      if (ls == 0) slope0 = d1;
      if (!(d1 < 0.0f || d1 > 0.0f)) break;   // exact minimiser, or NaN
      if (d1 < 0.0f) lo = alpha; else hi = alpha;
      float nxt = alpha - d1 * gx_rcp(d2);
      const bool bad = !(nxt > lo && nxt < hi);
      if (bad) nxt = hi < 3.0e38f ? 0.5f * (lo + hi) : 2.0f * gx_max(lo, 1.0f);
      // converged when the step lands in the same active set
      int changed = bad ? 1 : 0;
#pragma unroll
      for (int j = 0; j < LPL; ++j) {
#pragma unroll
        for (int i = 0; i < K; ++i)
          changed |= ((gx_fma(alpha, jvl[j][i], jl[j][i]) < 0.0f) != (gx_fma(nxt, jvl[j][i], jl[j][i]) < 0.0f)) && rows[j].lD[i] > 0.0f;
        if (rows[j].con) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            changed |= (gx_fma(alpha, jvc[j][k], jc[j][k]) < 0.0f) != (gx_fma(nxt, jvc[j][k], jc[j][k]) < 0.0f);
        }
      }
      changed = gx_art_or<LANES>(changed, mask);
      alpha = nxt;
      if (!changed) break;
    }
    float ss = 0.0f, qq = 0.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r) { ar[r] = gx_fma(alpha, br[r], ar[r]); Mar[r] = gx_fma(alpha, mvr[r], Mar[r]); }
#pragma unroll
    for (int j = 0; j < LPL; ++j) {
#pragma unroll
      for (int i = 0; i < K; ++i) {
        al[j][i] = gx_fma(alpha, bl[j][i], al[j][i]);
        Mal[j][i] = gx_fma(alpha, mvl[j][i], Mal[j][i]);
        jl[j][i] = gx_fma(alpha, jvl[j][i], jl[j][i]);
        ss = gx_fma(bl[j][i], bl[j][i], ss);
        qq = gx_fma(al[j][i], al[j][i], qq);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) jc[j][k] = gx_fma(alpha, jvc[j][k], jc[j][k]);
    }
    ++iters;
    // mjx stops on cost(prev) - cost(now) < tolerance (scaled).  In fp32 that difference of two
    // O(1e3) numbers is rounding noise, so the decrease is taken from the line-search model
    // instead: the area under the (piecewise linear) directional derivative, -alpha*f'(0)/2.
    const float improvement = -0.5f * alpha * slope0;
    // ... and a step below fp32 resolution of qacc cannot improve anything (near-singular mass
    // matrices: the planar roots of ant / walker degenerate at yaw = +-90 deg, where the x and y
    // slides align)
    ss = gx_art_sum<LANES>(ss, mask) + br[0] * br[0] + br[1] * br[1] + br[2] * br[2];
    qq = gx_art_sum<LANES>(qq, mask) + ar[0] * ar[0] + ar[1] * ar[1] + ar[2] * ar[2];
    if (!(improvement >= m.tol_scale) || !(alpha * alpha * ss > 1e-12f * qq)) finish = true;
  }
  if (iters_out) *iters_out = iters;
  if (!keep_warm) {
#pragma unroll
    for (int r = 0; r < 3; ++r) st.wr[r] = ar[r];
#pragma unroll
    for (int j = 0; j < LPL; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) st.wl[j][i] = al[j][i];
  }
  if (forward_only) {
#pragma unroll
    for (int r = 0; r < 3; ++r) { qro[r] = st.qr[r]; vro[r] = st.vr[r]; }
#pragma unroll
    for (int j = 0; j < LPL; ++j)
#pragma unroll
      for (int i = 0; i < K; ++i) { qlo[j][i] = st.ql[j][i]; vlo[j][i] = st.vl[j][i]; }
    return;
  }
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    vro[r] = gx_fma(m.dt, er[r], st.vr[r]);
    qro[r] = gx_fma(m.dt, vro[r], st.qr[r]);
    if (commit) { st.vr[r] = vro[r]; st.qr[r] = qro[r]; }
  }
#pragma unroll
  for (int j = 0; j < LPL; ++j)
#pragma unroll
    for (int i = 0; i < K; ++i) {
      vlo[j][i] = gx_fma(m.dt, el[j][i], st.vl[j][i]);
      qlo[j][i] = gx_fma(m.dt, vlo[j][i], st.ql[j][i]);
      if (commit) { st.vl[j][i] = vlo[j][i]; st.ql[j][i] = qlo[j][i]; }
    }
}

#endif  // GX_ART_CUH_
