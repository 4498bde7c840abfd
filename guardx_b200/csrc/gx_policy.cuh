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
// belong to the same net and read the same activation vector), so the inner loop is 1 LDS.128 per 4 FFMA per env.
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
#define GX_POL_MAX_ACT 16
#define GX_STREAM_POLICY 0x504f4c49u /* exploration noise */

__host__ __device__ inline size_t gx_policy_weight_count(int obs_dim, int act_dim) {
  const int k1 = (obs_dim + 3) & ~3;
  return (size_t)k1 * GX_POL_UNITS + GX_POL_UNITS + 3 * GX_POL_HID * GX_POL_HID + GX_POL_UNITS +
         (size_t)(act_dim + 2) * GX_POL_HID + (act_dim + 2) + act_dim;
}

__host__ __device__ inline size_t gx_policy_smem_floats(int k1) {
  return (size_t)GX_POL_ENVS * k1 + 2 * (size_t)GX_POL_ENVS * GX_POL_UNITS + 2 * (size_t)GX_POL_ENVS * (GX_POL_MAX_ACT + 2);
}

// one layer for this thread's unit: 16 envs in two groups of 8 accumulators; `in` rows have `stride` floats, the
// unit's inputs start at `in + off`; K is a multiple of 4
template <int K>
__device__ __forceinline__ void gx_policy_layer(const float (&w)[K], float bias, const float* in, int stride, int off, int env0,
                                                float* out, int j) {
#pragma unroll 1
  for (int grp = 0; grp < 2; ++grp) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = bias;
    const float* base = in + (size_t)(env0 + grp * 8) * stride + off;
#pragma unroll
    for (int k4 = 0; k4 < K / 4; ++k4) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float4 x = *reinterpret_cast<const float4*>(base + e * stride + 4 * k4);
        acc[e] = __fmaf_rn(w[4 * k4 + 0], x.x, acc[e]);
        acc[e] = __fmaf_rn(w[4 * k4 + 1], x.y, acc[e]);
        acc[e] = __fmaf_rn(w[4 * k4 + 2], x.z, acc[e]);
        acc[e] = __fmaf_rn(w[4 * k4 + 3], x.w, acc[e]);
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) out[(size_t)(env0 + grp * 8 + e) * GX_POL_UNITS + j] = tanhf(acc[e]);
  }
}

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
  const int tid = threadIdx.x, j = tid % GX_POL_UNITS, half = tid / GX_POL_UNITS;
  const int e_blk = blockIdx.x * GX_POL_ENVS;
  const int OUT = A + 2;
  const float* W1 = W;
  const float* B1 = W1 + (size_t)K1 * GX_POL_UNITS;
  const float* W2 = B1 + GX_POL_UNITS;
  const float* B2 = W2 + 3 * GX_POL_HID * GX_POL_HID;
  const float* W3 = B2 + GX_POL_UNITS;
  const float* B3 = W3 + (size_t)OUT * GX_POL_HID;
  const float* LS = B3 + OUT;
  // this thread's first-layer column (coalesced across j) while the observation tile is staged
  float w1[K1];
#pragma unroll
  for (int k = 0; k < K1; ++k) w1[k] = W1[(size_t)k * GX_POL_UNITS + j];
  const float b1 = B1[j];
  for (int idx = tid; idx < GX_POL_ENVS * K1; idx += GX_POL_THREADS) {
    const int e = idx / K1, k = idx - e * K1, g = e_blk + e;
    xs[idx] = (g < n && k < obs_dim) ? obs[(size_t)g * obs_dim + k] : 0.0f;
  }
  __syncthreads();
  gx_policy_layer<K1>(w1, b1, xs, K1, 0, half * 16, h1, j);
  const int net = j / GX_POL_HID, u = j - net * GX_POL_HID;
  float w2[GX_POL_HID];
#pragma unroll
  for (int k = 0; k < GX_POL_HID; ++k) w2[k] = W2[((size_t)net * GX_POL_HID + k) * GX_POL_HID + u];
  const float b2 = B2[j];
  __syncthreads();
  gx_policy_layer<GX_POL_HID>(w2, b2, h1, GX_POL_UNITS, net * GX_POL_HID, half * 16, h2, j);
  __syncthreads();
  // heads: one 64-long dot product per (env, output)
  for (int idx = tid; idx < GX_POL_ENVS * OUT; idx += GX_POL_THREADS) {
    const int e = idx / OUT, o = idx - e * OUT;
    const int hn = o < A ? 0 : o - A + 1;
    const float4* hv = reinterpret_cast<const float4*>(h2 + (size_t)e * GX_POL_UNITS + hn * GX_POL_HID);
    const float4* wv = reinterpret_cast<const float4*>(W3 + (size_t)o * GX_POL_HID);
    float acc = B3[o];
#pragma unroll
    for (int k4 = 0; k4 < GX_POL_HID / 4; ++k4) {
      const float4 x = hv[k4], w = __ldg(wv + k4);
      acc = __fmaf_rn(w.x, x.x, acc); acc = __fmaf_rn(w.y, x.y, acc);
      acc = __fmaf_rn(w.z, x.z, acc); acc = __fmaf_rn(w.w, x.w, acc);
    }
    head[e * (GX_POL_MAX_ACT + 2) + o] = acc;
  }
  __syncthreads();
  // Gaussian sample per (env, action dimension)
  for (int idx = tid; idx < GX_POL_ENVS * A; idx += GX_POL_THREADS) {
    const int e = idx / A, a = idx - e * A, g = e_blk + e;
    if (g >= n) continue;
    const float mu = head[e * (GX_POL_MAX_ACT + 2) + a];
    const float4 c = d2[g];
    const unsigned long long gid = env_id_offset + (unsigned long long)g;
    // counter: (global env id lo, age lo, age hi ^ env id hi, stream + action block)
    const gx_u32x4 r = gx_philox4x32_10((unsigned)gid, __float_as_uint(c.z), __float_as_uint(c.w) ^ (unsigned)(gid >> 32),
                                        GX_STREAM_POLICY + (unsigned)(a >> 2), k0, k1);
    // Box-Muller: words (0,1) -> normals 0,1; words (2,3) -> normals 2,3
    const int pair = (a >> 1) & 1;
    const float u1 = 1.0f - gx_u01(r.v[2 * pair]);           // (0, 1]
    const float u2 = gx_u01(r.v[2 * pair + 1]);
    const float rad = sqrtf(-2.0f * logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    const float z = rad * ((a & 1) ? sn : cs);
    const float ls = LS[a], sd = expf(ls);
    const float av = __fmaf_rn(sd, z, mu);
    const float dlt = av - mu;
    lp[e * GX_POL_MAX_ACT + a] = -(dlt * dlt) / (2.0f * sd * sd) - ls - 0.91893853320467274f;
    act[(size_t)g * A + a] = av;
    if (mu_out) mu_out[(size_t)g * A + a] = mu;
  }
  __syncthreads();
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
