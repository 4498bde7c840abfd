// gx_math.h -- deterministic fp32 elementary functions shared by host and device.
//
// Every function here is built only from IEEE-754 correctly rounded primitives
// (+, -, *, /, sqrt, fma, rint, floor) in a FIXED order, so the same inputs give
// the same bits under gcc (-ffp-contract=off) and nvcc (-fmad=false, default
// -prec-div=true -prec-sqrt=true -ftz=false).  This is what lets the integer
// outputs of the env step (done flags, lidar bin indices, violation flags,
// reset masks) be compared bit-exactly between the CUDA kernels and the CPU
// oracle over a free-running horizon: libm atan2f/expf/sinf and libdevice's
// differ by 1-2 ulp, and the Point robot's bang-bang servo amplifies that into
// different trajectories within a few hundred steps ("numerics").
//
// Accuracy (tests/test_gx_math.py measures it against float64): sin/cos <= 2 ulp
// for |x| < 1e4, atan2 <= 2.5 ulp, exp <= 2 ulp.  The reference's own elementary
// functions are whatever XLA emits, so there is no bit-level target to hit there.
//
// Contractions are explicit: gx_fma() is the only place a fused multiply-add may
// happen.  Never compile a translation unit that includes this header with
// -ffast-math / --use_fast_math / -fmad=true.
#ifndef GX_MATH_H_
#define GX_MATH_H_

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GX_HD __host__ __device__ __forceinline__
#else
#define GX_HD static inline
#endif

#define GX_PI_F 3.14159274101257324f      /* float32(pi)   0x40490fdb */
#define GX_PIO2_F 1.57079637050628662f    /* float32(pi/2) 0x3fc90fdb */
#define GX_TWO_PI_F 6.28318548202514648f  /* float32(2pi)  0x40c90fdb */

GX_HD float gx_fma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return __builtin_fmaf(a, b, c);
#endif
}

GX_HD float gx_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(x);
#else
  return __builtin_sqrtf(x);
#endif
}

GX_HD float gx_div(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}

GX_HD float gx_abs(float x) { return fabsf(x); }

GX_HD uint32_t gx_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c;
  c.f = f;
  return c.u;
#endif
}

GX_HD float gx_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c;
  c.u = u;
  return c.f;
#endif
}

GX_HD int gx_isnan(float x) { return x != x; }
/* true for +-inf and NaN */
GX_HD int gx_notfinite(float x) { return (gx_f2u(x) & 0x7f800000u) == 0x7f800000u; }

/* jnp.maximum / jnp.minimum semantics: NaN in either operand wins; zeros are
   ordered -0 < +0 (IEEE 754-2019 maximum/minimum).  On the device this is ONE
   instruction (PTX max.NaN.f32 / min.NaN.f32 -> FMNMX.NAN); the host spells the
   same function out.  Which NaN comes back is unspecified (tests treat all NaNs
   as equal). */
GX_HD float gx_max(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  float m = a > b ? a : b;
  if (a == b) m = gx_u2f(gx_f2u(a) & gx_f2u(b)); /* (+0,-0) -> +0 */
  m = (b != b) ? b : m;
  return (a != a) ? a : m;
#endif
}
GX_HD float gx_min(float a, float b) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
#else
  float m = a < b ? a : b;
  if (a == b) m = gx_u2f(gx_f2u(a) | gx_f2u(b)); /* (+0,-0) -> -0 */
  m = (b != b) ? b : m;
  return (a != a) ? a : m;
#endif
}
/* jnp.clip semantics (NaN stays NaN). */
GX_HD float gx_clip(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* sin and cos of x, Cody-Waite 3-term reduction by pi/2 with fma, Cephes
   single-precision kernels on [-pi/4, pi/4].  |x| up to ~1e4 keeps full
   accuracy; larger finite x is still deterministic and bounded, inf/NaN -> NaN. */
GX_HD void gx_sincos(float x, float* sn, float* cs) {
  float k = rintf(x * 0.636619746685028076f); /* x * 2/pi */
  float r = gx_fma(k, -1.57079637050628662f, x);
  r = gx_fma(k, 4.37113882867379289e-8f, r);
  r = gx_fma(k, 1.71512451000588188e-15f, r);
  /* quadrant = k mod 4 computed in float (exact for |k| < 2^24; garbage but
     deterministic beyond, where the reduction has lost all meaning anyway) */
  float qf = k - 4.0f * floorf(k * 0.25f);
  int q = (qf >= 0.0f && qf < 4.0f) ? (int)qf : 0;
  float z = r * r;
  float ps = gx_fma(-1.9515295891e-4f, z, 8.3321608736e-3f);
  ps = gx_fma(ps, z, -1.6666654611e-1f);
  ps = gx_fma(ps * z, r, r);
  float pc = gx_fma(2.443315711809948e-5f, z, -1.388731625493765e-3f);
  pc = gx_fma(pc, z, 4.166664568298827e-2f);
  pc = gx_fma(pc * z, z, gx_fma(-0.5f, z, 1.0f));
  float s, c;
  if (q & 1) { s = pc; c = ps; } else { s = ps; c = pc; }
  if (q & 2) s = -s;
  if ((q + 1) & 2) c = -c;
  *sn = s;
  *cs = c;
}

/* atan2(y, x) in (-pi, pi].  atan on [0,1] by an 8-term odd minimax polynomial,
   then octant fix-ups with single correctly rounded subtractions. */
GX_HD float gx_atan2(float y, float x) {
  float ax = fabsf(x), ay = fabsf(y);
  float mx = ax > ay ? ax : ay;
  float mn = ax > ay ? ay : ax;
  float a = (mx == 0.0f) ? 0.0f : gx_div(mn, mx);
  float s = a * a;
  float r = 0.00282363896f;
  r = gx_fma(r, s, -0.0159569028f);
  r = gx_fma(r, s, 0.0425049886f);
  r = gx_fma(r, s, -0.0748900920f);
  r = gx_fma(r, s, 0.106347933f);
  r = gx_fma(r, s, -0.142027363f);
  r = gx_fma(r, s, 0.199926957f);
  r = gx_fma(r, s, -0.333331018f);
  r = r * s;
  r = gx_fma(r, a, a);
  if (ay > ax) r = GX_PIO2_F - r;
  if (x < 0.0f) r = GX_PI_F - r;
  if (x != x || y != y) r = x + y; /* NaN in, NaN out */
  return (gx_f2u(y) & 0x80000000u) ? -r : r;
}

/* exp(x).  k = rint(x*log2e), two-term ln2 reduction, degree-6 polynomial,
   result scaled in two steps so that denormal results round once-ish and
   identically on both sides. */
GX_HD float gx_exp(float x) {
  /* branch-free: evaluate on the clamped argument, patch the special cases with
     selects at the end (NaN falls through every comparison and is restored last) */
  float xc = x < -104.0f ? -104.0f : x;
  xc = xc > 88.7228394f ? 88.7228394f : xc;
  float k = rintf(xc * 1.44269502162933350f);
  float r = gx_fma(k, -0.693147182464599609f, xc);
  r = gx_fma(k, 1.90465421212593356e-9f, r);
  float p = 1.37805939e-3f;
  p = gx_fma(p, r, 8.37312452e-3f);
  p = gx_fma(p, r, 4.16695364e-2f);
  p = gx_fma(p, r, 1.66664720e-1f);
  p = gx_fma(p, r, 4.99999851e-1f);
  p = gx_fma(p, r, 1.0f);
  p = gx_fma(p, r, 1.0f);
  float kk = (k == k) ? k : 0.0f; /* NaN -> 0 before the int conversion */
  int ki = (int)kk;               /* |k| <= 151 here */
  int k1 = ki / 2, k2 = ki - k1;
  float s1 = gx_u2f((uint32_t)(k1 + 127) << 23);
  float s2 = gx_u2f((uint32_t)(k2 + 127) << 23);
  float res = (p * s1) * s2;
  res = (x > 88.7228394f) ? gx_u2f(0x7f800000u) : res;
  res = (x < -104.0f) ? 0.0f : res;
  return (x != x) ? x : res;
}

#endif /* GX_MATH_H_ */
