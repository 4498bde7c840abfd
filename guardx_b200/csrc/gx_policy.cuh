// gx_policy.cuh -- the policy half of a PPO / PPO-Lagrangian rollout step in ONE kernel.
//
// The reference's rollout step (safe_rl_libX/ppo/ppo_train.py:364-379, ppo_core.py:116-160) is
//   a, v, logp = ac.step(o)      three fp32 tanh MLPs obs -> 64 -> 64 -> {act_dim | 1 | 1} (actor mean, value critic,
//                                cost critic: TRPO-Lag's vc), a state-independent log-std, a Gaussian sample and its
//                                log-probability  -- ~30 small torch / cuBLAS kernels per step
//   buf.store(o, a, r, v, logp)  five more copies
// followed by env.step.  Here the whole `ac.step` + the buffer writes are one launch that reads the observation row
// the env kernel wrote (directly in the rollout buffer) and writes act / val / cval / logp into the rollout buffer's
// rows; the env kernel that follows reads `act` from there.  A rollout step is two launches.
//
// fp32 FFMA only (the reference policy is fp32; the largest contraction is a [32 x 64] x [64 x 64] tile per net: a
// tensor-core path would need tf32 / bf16 operands, i.e. other numbers than the reference's).  Work split: one block
// per 32 envs, 384 threads = 192 hidden units (3 nets x 64) x 2 halves of 16 envs.  A thread keeps ITS unit's weight
// column in registers (obs_dim, then 64 values) and walks its 16 envs in two groups of 8 independent accumulators;
// the activations of the 32 envs live in shared memory and are read as warp-wide broadcasts (all lanes of a warp
// belong to the same net and read the same activation vector), so the inner loop is 1 LDS.128 + 2 FFMA2 (packed
// fp32 pairs, fma.rn.f32x2) per env per four inputs.
//
// Packed weights (device floats, written by guardx_b200/ppo_lag.py:pack_policy; K1 = obs_dim rounded up to 4):
//   W1 [K1][192]   W1[k][net*64+u] = layer-0 weight of unit u of net `net` for input k (rows k >= obs_dim are 0)
//   B1 [192]
//   W2 [3][64][64] W2[net][k][u]
//   B2 [192]
//   W3 [A+2][64]   rows 0..A-1: actor mean head (net 0), row A: value head (net 1), row A+1: cost-value head (net 2)
//   B3 [A+2]
//   LS [A]         log_std
// nets: 0 = actor mean (mu_net), 1 = value critic (v), 2 = cost critic (vc).
//
// Exploration noise: a = mu + exp(log_std) * z, z ~ N(0,1) by Box-Muller on Philox4x32-10 keyed by
// (noise seed ^ env seed, global env id, the env's own step count `age`, action block) -- a pure function of
// device-resident counters, so the launch can be captured in a CUDA graph and replayed, and results do not depend on
// the number of GPUs.  (The reference draws from torch's global generator; streams are not comparable.)
// logp = sum_a [ -(a - mu)^2 / (2 sigma^2) - log_std - log sqrt(2 pi) ], exactly Normal.log_prob of the STORED action.
#ifndef GX_POLICY_CUH_
#define GX_POLICY_CUH_

#include "gx_device.cuh"

#define GX_POL_HID 64
#define GX_POL_UNITS 192   // 3 nets x 64
#define GX_POL_ENVS 32     // envs per block
#define GX_POL_THREADS 384
#define GX_POL_GROUPS (GX_POL_ENVS * GX_POL_UNITS / GX_POL_THREADS / 8)   // groups of 8 envs per thread
#define GX_POL_MAX_ACT 16
#define GX_STREAM_POLICY 0x504f4c49u /* exploration noise */

__host__ __device__ inline size_t gx_policy_weight_count(int obs_dim, int act_dim) {
  const int k1 = (obs_dim + 3) & ~3;
  return (size_t)k1 * GX_POL_UNITS + GX_POL_UNITS + 3 * GX_POL_HID * GX_POL_HID + GX_POL_UNITS +
         (size_t)(act_dim + 2) * GX_POL_HID + (act_dim + 2) + act_dim;
}

__host__ __device__ inline size_t gx_policy_smem_floats(int k1) {
  return (size_t)GX_POL_ENVS * k1 + 2 * (size_t)GX_POL_ENVS * GX_POL_UNITS + 2 * (size_t)GX_POL_ENVS * (GX_POL_MAX_ACT + 2) +
         3 * (size_t)GX_POL_HID * GX_POL_HID + GX_POL_UNITS + (size_t)(GX_POL_MAX_ACT + 2) * (GX_POL_HID + 1) + GX_POL_MAX_ACT + 8;
}

// packed fp32 pair arithmetic (Blackwell FFMA2): d = a * b + c on both halves, round to nearest, one issue slot
__device__ __forceinline__ unsigned long long gx_pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long gx_fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float gx_sum2(unsigned long long v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo + hi;
}

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the SFU (ex2 + rcp): absolute error ~2e-7, which is what the 1e-5 bound on the heads
// needs (libdevice's tanhf is ~55 instructions, 17 % of the kernel); exact limits at +-inf, NaN propagates
__device__ __forceinline__ float gx_tanh_fast(float x) {
  const float t = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, t + 1.0f);
}

// one layer for this thread's unit: 16 envs in two groups of 8 accumulator PAIRS (even-k and odd-k partial sums of one
// env share a 64-bit register: the weights w[2m], w[2m+1] and the activations x[2m], x[2m+1] are adjacent already, so the
// inner loop is 1 LDS.128 + 2 FFMA2 per env per four inputs); `in` rows have `stride` floats, the unit's inputs start at
// `in + off`; K is a multiple of 4
template <int K>
__device__ __forceinline__ void gx_policy_layer(const float (&w)[K], float bias, const float* in, int stride, int off, int env0,
                                                float* out, int j) {
  unsigned long long w2[K / 2];
#pragma unroll
  for (int m = 0; m < K / 2; ++m) w2[m] = gx_pack2(w[2 * m], w[2 * m + 1]);
#pragma unroll 1
  for (int grp = 0; grp < GX_POL_GROUPS; ++grp) {
    unsigned long long acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = gx_pack2(bias, 0.0f);
    const float* base = in + (size_t)(env0 + grp * 8) * stride + off;
#pragma unroll
    for (int k4 = 0; k4 < K / 4; ++k4) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const ulonglong2 x = *reinterpret_cast<const ulonglong2*>(base + e * stride + 4 * k4);
        acc[e] = gx_fma2(w2[2 * k4], x.x, acc[e]);
        acc[e] = gx_fma2(w2[2 * k4 + 1], x.y, acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) out[(size_t)(env0 + grp * 8 + e) * GX_POL_UNITS + j] = gx_tanh_fast(gx_sum2(acc[e]));
  }
}

__device__ __forceinline__ void gx_cp_async4(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gx_cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}

// N(0,1) draw for action dimension `a` of global env `gid` at step count (age_lo, age_hi): Box-Muller on Philox words
// (0,1) -> normals 0,1 and (2,3) -> normals 2,3 of action block a >> 2
__device__ __forceinline__ float gx_policy_noise(unsigned long long gid, unsigned age_lo, unsigned age_hi, int a, unsigned k0, unsigned k1) {
  const gx_u32x4 r = gx_philox4x32_10((unsigned)gid, age_lo, age_hi ^ (unsigned)(gid >> 32), GX_STREAM_POLICY + (unsigned)(a >> 2), k0, k1);
  const int pair = (a >> 1) & 1;
  const float u1 = 1.0f - gx_u01(r.v[2 * pair]);           // (0, 1]
  const float u2 = gx_u01(r.v[2 * pair + 1]);
  const float rad = sqrtf(-2.0f * logf(u1));
  float sn, cs;
  sincospif(2.0f * u2, &sn, &cs);
  return rad * ((a & 1) ? sn : cs);
}

// Phases, each ended by one barrier; everything a later phase needs from global memory is requested at the very top
// (cp.async into shared memory, or into registers), so no phase but the first waits for L2 / HBM:
//   0  cp.async: observation tile, W2, W3 | B3 | log_std;  registers: first-layer column, the (env, action) noise draws
//   1  layer 1 -> h1        2  layer 2 (column from shared memory) -> h2        3  heads -> head
//   4  sample, per-dimension log-prob terms, act / mu stores        5  per-env logp / val / cval stores
template <int K1>
__global__ void __launch_bounds__(GX_POL_THREADS, 1)
gx_policy_kernel(const float* __restrict__ W, const float* __restrict__ obs, int n, int obs_dim, int A, const float4* __restrict__ d2,
                 unsigned long long env_id_offset, unsigned k0, unsigned k1, float* __restrict__ act, float* __restrict__ val,
                 float* __restrict__ cval, float* __restrict__ logp, float* __restrict__ mu_out) {
  extern __shared__ __align__(16) float gx_pol_smem[];
  float* xs = gx_pol_smem;                                   // [32][K1]   observations
  float* h1 = xs + GX_POL_ENVS * K1;                         // [32][192]  first hidden layer, all three nets
  float* h2 = h1 + GX_POL_ENVS * GX_POL_UNITS;               // [32][192]
  float* head = h2 + GX_POL_ENVS * GX_POL_UNITS;             // [32][A+2]  mu | v | vc
  float* lp = head + GX_POL_ENVS * (GX_POL_MAX_ACT + 2);     // [32][A]    per-dimension log-prob terms
  float* w2s = lp + GX_POL_ENVS * (GX_POL_MAX_ACT + 2);      // [3][64][64] second-layer weights
  float* w3s = w2s + 3 * GX_POL_HID * GX_POL_HID;            // B2 [192] | W3 [A+2][64] | B3 [A+2] | LS [A]
  const int tid = threadIdx.x, j = tid % GX_POL_UNITS, half = tid / GX_POL_UNITS;
  const int e_blk = blockIdx.x * GX_POL_ENVS;
  const int OUT = A + 2;
  const float* W1 = W;
  const float* B1 = W1 + (size_t)K1 * GX_POL_UNITS;
  const float* W2 = B1 + GX_POL_UNITS;
  const float* B2 = W2 + 3 * GX_POL_HID * GX_POL_HID;
  const int tail = GX_POL_UNITS + OUT * GX_POL_HID + OUT + A;   // floats from B2 to the end of the pack
  // ---- phase 0 --------------------------------------------------------------------------------------------------
  for (int idx = tid; idx < GX_POL_ENVS * K1; idx += GX_POL_THREADS) {
    const int e = idx / K1, k = idx - e * K1, g = e_blk + e;
    if (g < n && k < obs_dim) gx_cp_async4(xs + idx, obs + (size_t)g * obs_dim + k);
    else xs[idx] = 0.0f;
  }
  for (int c = tid; c < 3 * GX_POL_HID * GX_POL_HID / 4; c += GX_POL_THREADS) gx_cp_async16(w2s + 4 * c, W2 + 4 * c);
  if ((((size_t)(B2 - W)) & 3) == 0) {
    for (int c = tid; c < tail / 4; c += GX_POL_THREADS) gx_cp_async16(w3s + 4 * c, B2 + 4 * c);
    for (int c = (tail & ~3) + tid; c < tail; c += GX_POL_THREADS) gx_cp_async4(w3s + c, B2 + c);
  } else {
    for (int c = tid; c < tail; c += GX_POL_THREADS) gx_cp_async4(w3s + c, B2 + c);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  float w1[K1];
#pragma unroll
  for (int k = 0; k < K1; ++k) w1[k] = W1[(size_t)k * GX_POL_UNITS + j];
  const float b1 = B1[j];
  // exploration noise of this thread's (env, action) pairs: independent of the networks, so it is drawn while the loads fly
  float z[2] = {0.0f, 0.0f};
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int idx = tid + q * GX_POL_THREADS;
    if (idx < GX_POL_ENVS * A) {
      const int e = idx / A, a = idx - e * A, g = e_blk + e;
      if (g < n) {
        const float4 c = d2[g];
        z[q] = gx_policy_noise(env_id_offset + (unsigned long long)g, __float_as_uint(c.z), __float_as_uint(c.w), a, k0, k1);
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  // ---- phase 1, 2 -----------------------------------------------------------------------------------------------
  gx_policy_layer<K1>(w1, b1, xs, K1, 0, half * 8 * GX_POL_GROUPS, h1, j);
  const int net = j / GX_POL_HID, u = j - net * GX_POL_HID;
  float w2[GX_POL_HID];
#pragma unroll
  for (int k = 0; k < GX_POL_HID; ++k) w2[k] = w2s[(net * GX_POL_HID + k) * GX_POL_HID + u];
  const float b2 = w3s[j];
  __syncthreads();
  gx_policy_layer<GX_POL_HID>(w2, b2, h1, GX_POL_UNITS, net * GX_POL_HID, half * 8 * GX_POL_GROUPS, h2, j);
  __syncthreads();
  // ---- phase 3: heads, one 64-long dot product per (env, output) ---------------------------------------------------
  const float* W3 = w3s + GX_POL_UNITS;
  const float* B3 = W3 + OUT * GX_POL_HID;
  const float* LS = B3 + OUT;
  for (int idx = tid; idx < GX_POL_ENVS * OUT; idx += GX_POL_THREADS) {
    const int e = idx / OUT, o = idx - e * OUT;
    const int hn = o < A ? 0 : o - A + 1;
    const float4* hv = reinterpret_cast<const float4*>(h2 + (size_t)e * GX_POL_UNITS + hn * GX_POL_HID);
    const float* wv = W3 + o * GX_POL_HID;
    float acc0 = B3[o], acc1 = 0.0f;
#pragma unroll
    for (int k4 = 0; k4 < GX_POL_HID / 4; ++k4) {
      const float4 x = hv[k4];
      acc0 = __fmaf_rn(wv[4 * k4 + 0], x.x, acc0); acc1 = __fmaf_rn(wv[4 * k4 + 1], x.y, acc1);
      acc0 = __fmaf_rn(wv[4 * k4 + 2], x.z, acc0); acc1 = __fmaf_rn(wv[4 * k4 + 3], x.w, acc1);
    }
    head[e * (GX_POL_MAX_ACT + 2) + o] = acc0 + acc1;
  }
  __syncthreads();
  // ---- phase 4: Gaussian sample per (env, action dimension) -------------------------------------------------------
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int idx = tid + q * GX_POL_THREADS;
    if (idx < GX_POL_ENVS * A) {
      const int e = idx / A, a = idx - e * A, g = e_blk + e;
      if (g < n) {
        const float mu = head[e * (GX_POL_MAX_ACT + 2) + a];
        const float ls = LS[a], sd = expf(ls);
        const float av = __fmaf_rn(sd, z[q], mu);
        const float dlt = av - mu;
        lp[e * GX_POL_MAX_ACT + a] = -(dlt * dlt) / (2.0f * sd * sd) - ls - 0.91893853320467274f;
        act[(size_t)g * A + a] = av;
        if (mu_out) mu_out[(size_t)g * A + a] = mu;
      }
    }
  }
  __syncthreads();
  // ---- phase 5 ----------------------------------------------------------------------------------------------------
  if (tid < GX_POL_ENVS) {
    const int g = e_blk + tid;
    if (g < n) {
      float s = 0.0f;
      for (int a = 0; a < A; ++a) s += lp[tid * GX_POL_MAX_ACT + a];   // fixed order: deterministic
      logp[g] = s;
      val[g] = head[tid * (GX_POL_MAX_ACT + 2) + A];
      cval[g] = head[tid * (GX_POL_MAX_ACT + 2) + A + 1];
    }
  }
}

#endif  // GX_POLICY_CUH_
