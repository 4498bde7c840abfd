// gx_philox.h -- Philox4x32-10 counter-based RNG (Salmon et al., SC'11), host+device.
//
// The reference draws layouts with JAX threefry keys (engine.py:216, 431, 447,
// 563, 618).  This build uses Philox instead (BASELINE.json north_star), keyed so
// that every draw is a pure function of (seed, stream, indices): results do not
// depend on grid shape, on the number of GPUs, or on evaluation order.  RNG
// streams are therefore NOT comparable with the reference; parity injects
// layouts/actions.  Known-answer vectors from the Random123
// distribution are checked in tests/test_philox.py.
#ifndef GX_PHILOX_H_
#define GX_PHILOX_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define GXP_HD __host__ __device__ __forceinline__
#else
#define GXP_HD static inline
#endif

typedef struct { uint32_t v[4]; } gx_u32x4;

GXP_HD uint32_t gx_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

GXP_HD gx_u32x4 gx_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                 uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = gx_mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = gx_mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  gx_u32x4 out;
  out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
  return out;
}

/* 24-bit uniform in [0,1): exact in fp32, no rounding involved. */
GXP_HD float gx_u01(uint32_t bits) { return (float)(bits >> 8) * 5.9604644775390625e-8f; }

/* uniform in [lo, hi): lo + u*(hi-lo), two separate roundings (no fma). */
GXP_HD float gx_uniform(uint32_t bits, float lo, float hi) {
  float w = hi - lo;
  float t = gx_u01(bits) * w;
  return lo + t;
}

/* unbiased-enough integer in [0, n): high word of bits*n (Lemire, no rejection). */
GXP_HD uint32_t gx_randint(uint32_t bits, uint32_t n) { return gx_mulhi32(bits, n); }

/* Stream ids: the c3 word of the Philox counter.  Keeps draws for different
   purposes disjoint under one seed. */
#define GX_STREAM_LAYOUT 0x4c41594fu /* layout pool candidates */
#define GX_STREAM_PICK 0x5049434bu   /* pool index per env per reset */
#define GX_STREAM_ACTION 0x41435421u /* synthetic U(-1,1) actions */

#endif /* GX_PHILOX_H_ */
