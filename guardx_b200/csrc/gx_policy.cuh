// gx_policy.cuh -- the policy half of a PPO / PPO-Lagrangian rollout step in ONE kernel (gx_policy_kernel), and a whole
// rollout -- T x { policy, env step } -- in ONE persistent launch (gx_policy_rollout_kernel).
//
// The reference's rollout step (safe_rl_libX/ppo/ppo_train.py:364-379, ppo_core.py:116-160) is
//   a, v, logp = ac.step(o)      three fp32 tanh MLPs obs -> 64 -> 64 -> {act_dim | 1 | 1} (actor mean, value critic,
//                                cost critic: TRPO-Lag's vc), a state-independent log-std, a Gaussian sample and its
//                                log-probability  -- ~30 small torch / cuBLAS kernels per step
//   buf.store(o, a, r, v, logp)  five more copies
// followed by env.step.  gx_policy_kernel is the whole `ac.step` + the buffer writes as one launch that reads the
// observation row the env kernel wrote (directly in the rollout buffer) and writes act / val / cval / logp into the
// rollout buffer's rows; the env kernel that follows reads `act` from there.
//
// fp32 FFMA only (the reference policy is fp32; the largest contraction is a [32 x 64] x [64 x 64] tile per net: a
// tensor-core path would need tf32 / bf16 operands, i.e. other numbers than the reference's).
//
// What bounds a tiny fp32 MLP on an SM, measured (ncu, profiles/):
//  * the shared-memory return path: the LSU delivers ONE warp-wide 4-byte operand per clock (128 B / clk / SM; an
//    LDS.128 is 4 wavefronts, broadcast or not) while the four schedulers can retire four warp-wide FMAs.  The first
//    version kept a unit's weight column in registers and read every activation once per FMA pair -- 4 B per FMA,
//    18 k LSU cycles per 32-env block, which WAS its 9-10 us.  Hence register tiles: a thread owns an 8 unit x 4 env
//    tile (32 accumulator pairs) and per four inputs loads 8 weight quads + 4 activation quads (12 LDS.128) for
//    128 FMAs = 1.5 B per FMA;
//    Six warps per 32 envs: warp = (net, half of the block's envs).  Measured alternatives, all within 5 %: an 8 x 8 tile
//    on three warps (1 B per FMA, 232 registers; slower), 8 x 2 tiles on twelve warps (GX_POL_E = 2), the reduction split
//    over warp PAIRS that exchange partial sums through shared memory (12 warps).  What the layers cost is invariant under
//    the tiling: ~5.7 k cycles of FFMA2 issue per scheduler (FFMA2 issues once per 2.2 cycles per scheduler below 8
//    warps per scheduler, i.e. at the scalar FMA rate: tools/micro/ffma2.cu) PLUS ~5-6 k cycles of LDS.128 returns
//    (1.5-2 cycles each SM-wide at 6-12 warps: tools/micro/lds128.cu) -- the two add up instead of overlapping, which
//    points at the register file's bandwidth (an FFMA2 moves eight registers per lane, an LDS.128 four) as the shared
//    resource;
//  * and the activations: one tanh per 54 FMAs is as many instructions as the matrix products unless it is five
//    instructions (gx_tanh_fast).
// Lane mapping inside a warp tile: lane = 8 unit groups x 4 env groups; unit u = i * 8 + ug, env e = i * 4 + eg, so a
// quarter-warp reads 128 contiguous bytes of weights and ONE broadcast activation quad: no bank conflicts.
// The weights arrive as bulk-async copies (cp.async.bulk global -> shared, the non-tensor TMA path) on an mbarrier.
//
// Packed weights (device floats, written by guardx_b200/ppo_lag.py:pack_policy; K1 = obs_dim rounded up to 4):
//   W1 [3][K1 / 4][64][4]  W1[net][k / 4][u][k & 3] = layer-0 weight of unit u of net `net` for input k (inputs k >= obs_dim
//                          are 0): the layout the kernel keeps in shared memory, a thread's four weights are one LDS.128
//   W2 [3][16][64][4]      W2[net][k / 4][u][k & 3]
//   B1 [3][64]
//   B2 [3][64]
//   W3 [A+2][64]           rows 0..A-1: actor mean head (net 0), row A: value head (net 1), row A+1: cost-value head (net 2)
//   B3 [A+2]
//   LS [A]                 log_std
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
#include "gx_lanes.cuh"

#define GX_POL_HID 64
#define GX_POL_UNITS 192   // 3 nets x 64
#define GX_POL_ENVS 32     // envs per block
#ifndef GX_POL_E
#define GX_POL_E 4         // envs per thread tile (2 or 4); a warp tile is 4 E envs x 64 units.  4 -> 6 warps, 2 -> 12 warps: same speed (measured)
#endif
#define GX_POL_EPW (4 * GX_POL_E)                     // envs per warp tile
#define GX_POL_SUB (GX_POL_ENVS / GX_POL_EPW)         // warp tiles per net
#define GX_POL_WARPS (3 * GX_POL_SUB)                 // warp = (net, slice of the block's envs)
#define GX_POL_THREADS (32 * GX_POL_WARPS)
#define GX_POL_HS 72       // row stride of the hidden activations in shared memory (conflict-free tile stores)
#define GX_POL_MAX_ACT 16
#define GX_STREAM_POLICY 0x504f4c49u /* exploration noise */

__host__ __device__ inline size_t gx_policy_weight_count(int obs_dim, int act_dim) {
  const int k1 = (obs_dim + 3) & ~3;
  return (size_t)k1 * GX_POL_UNITS + GX_POL_UNITS + 3 * GX_POL_HID * GX_POL_HID + GX_POL_UNITS +
         (size_t)(act_dim + 2) * GX_POL_HID + (act_dim + 2) + act_dim;
}

// shared memory: xs [32][K1] | h1 [3][32][HS] | h2 [3][32][HS] | w1s [3][K1/4][64][4] | w2s [3][16][64][4] |
//                tail = B1 | B2 | W3 | B3 | LS (as packed) | mbarrier
__host__ __device__ inline size_t gx_policy_tail_floats(int A) { return 2 * GX_POL_UNITS + (size_t)(A + 2) * (GX_POL_HID + 1) + A; }
__host__ __device__ inline size_t gx_policy_smem_floats(int k1) {
  return (size_t)GX_POL_ENVS * k1 + 2 * 3 * (size_t)GX_POL_ENVS * GX_POL_HS + (size_t)k1 * GX_POL_UNITS + 3 * (size_t)GX_POL_HID * GX_POL_HID +
         ((gx_policy_tail_floats(GX_POL_MAX_ACT) + 3) & ~(size_t)3) + 8;
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

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the SFU, five instructions (FMUL, MUFU.EX2, FADD, MUFU.RCP, FFMA): absolute error
// ~2e-7, which is what the 1e-5 bound on the heads needs.  The activations are as many instructions as the matrix
// products here (one tanh per 54 FMAs): libdevice's tanhf is ~55 instructions, __expf + __fdividef with their range
// handling 28 (45 % of this kernel's executed instructions, ncu).  Exact limits: exp2 -> inf gives rcp 0 -> 1,
// exp2 -> 0 gives 1 - 2 = -1; NaN propagates.
__device__ __forceinline__ float gx_tanh_fast(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(x * 2.88539008177792681f));   // exp(2x) = 2^(2x log2 e)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return __fmaf_rn(-2.0f, r, 1.0f);
}

// One layer of one net for a warp's 16 env x 64 unit tile; this thread: 8 units x 4 envs.  Accumulator PAIRS: the
// even-k and the odd-k partial sum of one (env, unit) share a 64-bit register -- the weights w[2m], w[2m+1] and the
// activations x[2m], x[2m+1] are adjacent in shared memory already, so a quad of inputs is 2 FFMA2 per (env, unit) with
// no packing work.  ws: this net's weights [K / 4][64][4]; bs: its biases [64]; `in` rows have `stride` floats; out rows
// GX_POL_HS.  K is a multiple of 4.  The tile is produced and consumed by this warp alone: callers separate the layers
// with __syncwarp.
template <int K, int E>
__device__ __forceinline__ void gx_policy_layer(const float* ws, const float* bs, const float* in, int stride, float* out, int lane) {
  const int eg = lane >> 3, ug = lane & 7;
  unsigned long long acc[E][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float b = bs[i * 8 + ug];
#pragma unroll
    for (int e = 0; e < E; ++e) acc[e][i] = gx_pack2(b, 0.0f);
  }
  const float* wp = ws + ug * 4;
  const float* xp = in + (size_t)eg * stride;
#pragma unroll 2
  for (int k4 = 0; k4 < K / 4; ++k4) {
    ulonglong2 x[E];
#pragma unroll
    for (int e = 0; e < E; ++e) x[e] = *reinterpret_cast<const ulonglong2*>(xp + (e * 4) * stride + 4 * k4);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const ulonglong2 w = *reinterpret_cast<const ulonglong2*>(wp + (k4 * GX_POL_HID + i * 8) * 4);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e][i] = gx_fma2(w.x, x[e].x, acc[e][i]);
#pragma unroll
      for (int e = 0; e < E; ++e) acc[e][i] = gx_fma2(w.y, x[e].y, acc[e][i]);
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e)
#pragma unroll
    for (int i = 0; i < 8; ++i) out[(size_t)(e * 4 + eg) * GX_POL_HS + i * 8 + ug] = gx_tanh_fast(gx_sum2(acc[e][i]));
}

// mbarrier + bulk-async copy (global -> shared), the non-tensor TMA path
__device__ __forceinline__ unsigned gx_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gx_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gx_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void gx_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gx_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gx_bulk_g2s(float* smem_dst, const float* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(gx_smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(gx_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void gx_mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok)
                 : "r"(gx_smem_u32(bar)), "r"(parity)
                 : "memory");
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

// 64-long head dot product of output o for the hidden row hv (two interleaved accumulators)
__device__ __forceinline__ float gx_policy_head(const float* W3, const float* B3, const float4* hv, int o) {
  const float* wv = W3 + o * GX_POL_HID;
  float acc0 = B3[o], acc1 = 0.0f;
#pragma unroll
  for (int k4 = 0; k4 < GX_POL_HID / 4; ++k4) {
    const float4 x = hv[k4];
    acc0 = __fmaf_rn(wv[4 * k4 + 0], x.x, acc0); acc1 = __fmaf_rn(wv[4 * k4 + 1], x.y, acc1);
    acc0 = __fmaf_rn(wv[4 * k4 + 2], x.z, acc0); acc1 = __fmaf_rn(wv[4 * k4 + 3], x.w, acc1);
  }
  return acc0 + acc1;
}

// Phases:
//   0  thread 0: two bulk copies (W1, W2 blocks) onto the mbarrier; all threads: cp.async of the observation tile and the
//      small tail (biases, head rows, log_std); the lanes that will sample draw the exploration noise of the first action
//      block while the loads fly; ONE block barrier (tile + tail visible), then the weights' mbarrier
//   1  layer 1 -> h1        2  layer 2 -> h2   (per warp, __syncwarp between)
//   3  lane = env (16 per warp): head dot products; the actor warps sample, write act / mu / logp; the critic warps write
//      val / cval
template <int K1>
__global__ void __launch_bounds__(GX_POL_THREADS, 1)
gx_policy_kernel(const float* __restrict__ W, const float* __restrict__ obs, int n, int obs_dim, int A, const float4* __restrict__ d2,
                 unsigned long long env_id_offset, unsigned k0, unsigned k1, float* __restrict__ act, float* __restrict__ val,
                 float* __restrict__ cval, float* __restrict__ logp, float* __restrict__ mu_out) {
  extern __shared__ __align__(16) float gx_pol_smem[];
  float* xs = gx_pol_smem;                                   // [32][K1]      observations
  float* h1 = xs + GX_POL_ENVS * K1;                         // [3][32][HS]   first hidden layer
  float* h2 = h1 + 3 * GX_POL_ENVS * GX_POL_HS;              // [3][32][HS]
  float* w1s = h2 + 3 * GX_POL_ENVS * GX_POL_HS;             // [3][K1 / 4][64][4]
  float* w2s = w1s + K1 * GX_POL_UNITS;                      // [3][16][64][4]
  float* tl = w2s + 3 * GX_POL_HID * GX_POL_HID;             // B1 [192] | B2 [192] | W3 [A+2][64] | B3 [A+2] | LS [A]
  const int OUT = A + 2;
  const int tail = (int)gx_policy_tail_floats(A);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(tl + ((gx_policy_tail_floats(GX_POL_MAX_ACT) + 3) & ~(size_t)3));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int net = warp / GX_POL_SUB, half = warp % GX_POL_SUB;   // this warp: net `net`, envs [half * EPW, half * EPW + EPW) of the block
  const int e_blk = blockIdx.x * GX_POL_ENVS;
  const float* W1 = W;
  const float* W2 = W1 + (size_t)K1 * GX_POL_UNITS;
  const float* TL = W2 + 3 * GX_POL_HID * GX_POL_HID;
  // ---- phase 0 --------------------------------------------------------------------------------------------------
  if (tid == 0) {
    gx_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    constexpr unsigned b1 = K1 * GX_POL_UNITS * sizeof(float), b2 = 3 * GX_POL_HID * GX_POL_HID * sizeof(float);
    gx_mbar_expect_tx(bar, b1 + b2);
    gx_bulk_g2s(w1s, W1, b1, bar);
    gx_bulk_g2s(w2s, W2, b2, bar);
  }
  for (int idx = tid; idx < GX_POL_ENVS * K1; idx += GX_POL_THREADS) {
    const int e = idx / K1, k = idx - e * K1, g = e_blk + e;
    if (g < n && k < obs_dim) gx_cp_async4(xs + idx, obs + (size_t)g * obs_dim + k);
    else xs[idx] = 0.0f;
  }
  for (int c = tid; c < tail / 4; c += GX_POL_THREADS) gx_cp_async16(tl + 4 * c, TL + 4 * c);   // TL is 16-byte aligned: every block before it is a multiple of 4 floats
  for (int c = (tail & ~3) + tid; c < tail; c += GX_POL_THREADS) gx_cp_async4(tl + c, TL + c);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // exploration noise of this lane's env, first block of four action dimensions: independent of the networks, so it is
  // drawn while the loads fly (further blocks, act_dim > 4, are drawn in the sampling loop)
  const bool mine = lane < GX_POL_EPW;                 // phase 3: lane = env of this warp's slice
  const int le = half * GX_POL_EPW + (mine ? lane : 0);
  const int g = e_blk + le;
  unsigned age_lo = 0, age_hi = 0;
  float z0[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (net == 0 && mine && g < n) {
    const float4 c = d2[g];
    age_lo = __float_as_uint(c.z); age_hi = __float_as_uint(c.w);
#pragma unroll
    for (int a = 0; a < 4; ++a)
      if (a < A) z0[a] = gx_policy_noise(env_id_offset + (unsigned long long)g, age_lo, age_hi, a, k0, k1);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  gx_mbar_wait(bar, 0);
  // ---- phase 1, 2: this warp's net and envs ------------------------------------------------------------------------
  float* h1n = h1 + ((size_t)net * GX_POL_ENVS + half * GX_POL_EPW) * GX_POL_HS;
  float* h2n = h2 + ((size_t)net * GX_POL_ENVS + half * GX_POL_EPW) * GX_POL_HS;
  gx_policy_layer<K1, GX_POL_E>(w1s + (size_t)net * K1 * GX_POL_HID, tl + net * GX_POL_HID, xs + (size_t)half * GX_POL_EPW * K1, K1, h1n, lane);
  __syncwarp();
  gx_policy_layer<GX_POL_HID, GX_POL_E>(w2s + (size_t)net * GX_POL_HID * GX_POL_HID, tl + GX_POL_UNITS + net * GX_POL_HID, h1n, GX_POL_HS, h2n, lane);
  __syncwarp();
  // ---- phase 3: lane = env; one 64-long dot product per output ---------------------------------------------------------
  if (!mine) return;
  const float* W3 = tl + 2 * GX_POL_UNITS;
  const float* B3 = W3 + OUT * GX_POL_HID;
  const float* LS = B3 + OUT;
  const float4* hv = reinterpret_cast<const float4*>(h2n + (size_t)lane * GX_POL_HS);
  if (net != 0) {
    const float v = gx_policy_head(W3, B3, hv, A + net - 1);
    if (g < n) (net == 1 ? val : cval)[g] = v;
    return;
  }
  // Gaussian sample per action dimension, log-probability summed in a fixed order (deterministic)
  float s = 0.0f;
  for (int a = 0; a < A; ++a) {
    const float mu = gx_policy_head(W3, B3, hv, a);
    float z;
    if (a < 4) z = a == 0 ? z0[0] : (a == 1 ? z0[1] : (a == 2 ? z0[2] : z0[3]));
    else z = g < n ? gx_policy_noise(env_id_offset + (unsigned long long)g, age_lo, age_hi, a, k0, k1) : 0.0f;
    const float ls = LS[a], sd = expf(ls);
    const float av = __fmaf_rn(sd, z, mu);
    const float dlt = av - mu;
    s += -(dlt * dlt) / (2.0f * sd * sd) - ls - 0.91893853320467274f;
    if (g < n) {
      act[(size_t)g * A + a] = av;
      if (mu_out) mu_out[(size_t)g * A + a] = mu;
    }
  }
  if (g < n) logp[g] = s;
}

// ---------------------------------------------------------------------------------------------------------------------
// A whole PPO rollout (ppo_train.py:364-417: T x { ac.step(o); env.step(a); buf.store }) as ONE persistent launch, for
// the default Point shape (the shape the 16-lanes-per-env kernels handle: obs 43, act 2).
//
// Envs are independent and the policy is evaluated per env, so a block that owns 32 envs never needs another block: it
// keeps its envs' dynamic state in registers, the three networks' weights in shared memory (loaded ONCE, two bulk-async
// copies), and alternates two phases per control step, each closed by one barrier -- of the GROUP: the 16 warps form
// GX_PR_GROUPS groups that own 32 / GX_PR_GROUPS envs each and run out of phase (see the main loop).  Per group:
//   policy  (every thread first parks its env state in shared memory) the group's first three warps evaluate the three
//           MLPs for the group's observation rows that sit in shared memory (gx_policy_layer) and leave the actor means
//           in shared memory, the two value heads in the rollout buffer; meanwhile its fourth warp draws the step's
//           exploration noise (Philox + Box-Muller, lane = env)
//   env     all warps of the group, 16 lanes per env: Gaussian sample + log-prob, then gx_step_autoreset_lanes (the single-step
//           kernel's own device function); the new observation row goes to the rollout buffer AND to shared memory, where
//           the next policy phase reads it
// Against the 2T-launch graph this removes, per step: two launch / drain gaps, the reload of 82 KB of weights per block, and
// the HBM round trips of state, observation and action.  Every number is produced by the same expressions as in
// gx_policy_kernel + gx_step_lanes_kernel, so the rollout buffer is bit-identical (tests/test_ppo_lag.py).
#define GX_PR_ENVS 32
#define GX_PR_THREADS 512
#define GX_PR_K1 44
#ifndef GX_PR_GROUPS
#define GX_PR_GROUPS 4     // groups of warps that run their envs' steps out of phase (1, 2 or 4)
#endif
#ifndef GX_PR_E
#define GX_PR_E (GX_PR_GROUPS == 4 ? 2 : 4)   // envs per thread tile: a group of 8 envs is one 8-env warp tile per net
#endif

__host__ __device__ inline size_t gx_policy_rollout_smem_floats() {
  return (size_t)GX_PR_ENVS * GX_PR_K1 + 2 * 3 * (size_t)GX_PR_ENVS * GX_POL_HS + (size_t)GX_PR_K1 * GX_POL_UNITS + 3 * (size_t)GX_POL_HID * GX_POL_HID +
         ((gx_policy_tail_floats(2) + 3) & ~(size_t)3) + 8 + 2 * GX_PR_ENVS /* mu */ + 2 * GX_PR_ENVS /* z */ + 2 * GX_PR_ENVS /* age */ +
         18 * GX_PR_THREADS /* env state parked during the policy phase */;
}

template <bool STATS>
__global__ void __launch_bounds__(GX_PR_THREADS, 1)
gx_policy_rollout_kernel(const __grid_constant__ gx_params p, const GxState s, const float* __restrict__ pool, const float* __restrict__ W, int T,
                         unsigned k0, unsigned k1, float* __restrict__ obs, float2* __restrict__ act, float* __restrict__ reward,
                         float* __restrict__ done, float* __restrict__ cost, float* __restrict__ val, float* __restrict__ cval,
                         float* __restrict__ logp) {
  constexpr int K1 = GX_PR_K1, A = 2, OUT = A + 2, E = GX_PR_E, EPW = 4 * E;
  static_assert(GX_PR_ENVS == GX_POL_ENVS, "the rollout kernel uses the policy kernel's warp layout");
  extern __shared__ __align__(16) float gx_pol_smem[];
  float* xs = gx_pol_smem;                                   // [32][K1]      observation rows (column 43 stays 0)
  float* h1 = xs + GX_PR_ENVS * K1;                          // [3][32][HS]
  float* h2 = h1 + 3 * GX_PR_ENVS * GX_POL_HS;               // [3][32][HS]
  float* w1s = h2 + 3 * GX_PR_ENVS * GX_POL_HS;              // [3][K1 / 4][64][4]
  float* w2s = w1s + K1 * GX_POL_UNITS;                      // [3][16][64][4]
  float* tl = w2s + 3 * GX_POL_HID * GX_POL_HID;             // B1 | B2 | W3 | B3 | LS
  constexpr int tail = 2 * GX_POL_UNITS + OUT * (GX_POL_HID + 1) + A;
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(tl + ((tail + 3) & ~3));
  float* mu_s = reinterpret_cast<float*>(bar) + 8;           // [32][2] actor means of the current step
  float* z_s = mu_s + 2 * GX_PR_ENVS;                        // [32][2] exploration noise of the current step
  unsigned* age_s = reinterpret_cast<unsigned*>(z_s + 2 * GX_PR_ENVS);   // [32][2] step counts (lo, hi) that key the noise
  float* park = reinterpret_cast<float*>(age_s + 2 * GX_PR_ENVS) + threadIdx.x;   // [18][512] this thread's env state while it evaluates the policy
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sub = tid & 15, el = tid >> 4;                   // env phase: 16 lanes per env
  const int e_blk = blockIdx.x * GX_PR_ENVS;
  const int env = e_blk + el, n = p.n;
  const bool live = env < n;
  const int i = live ? env : n - 1;
  const unsigned gmask = gx_grp_mask();
  const float* W2 = W + (size_t)K1 * GX_POL_UNITS;
  const float* TL = W2 + 3 * GX_POL_HID * GX_POL_HID;
  // ---- once: weights, tail, first observation rows, env state ---------------------------------------------------------
  if (tid == 0) {
    gx_mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    constexpr unsigned b1 = K1 * GX_POL_UNITS * sizeof(float), b2 = 3 * GX_POL_HID * GX_POL_HID * sizeof(float);
    gx_mbar_expect_tx(bar, b1 + b2);
    gx_bulk_g2s(w1s, W, b1, bar);
    gx_bulk_g2s(w2s, W2, b2, bar);
  }
  for (int c = tid; c < tail; c += GX_PR_THREADS) tl[c] = TL[c];
  for (int idx = tid; idx < GX_PR_ENVS * K1; idx += GX_PR_THREADS) {
    const int e = idx / K1, k = idx - e * K1, g = e_blk + e;
    xs[idx] = (g < n && k < 43) ? obs[(size_t)g * 43 + k] : 0.0f;
  }
  GxDyn d;
  gx_load_dyn(s, i, d);
  const float2 g0 = s.goal[i];
  float gx_ = g0.x, gy_ = g0.y, ox, oy;
  gx_load_object(s, i, sub, gx_, gy_, ox, oy);
  if (sub == 0) { age_s[2 * el] = (unsigned)d.age; age_s[2 * el + 1] = (unsigned)(d.age >> 32); }
  __syncthreads();
  gx_mbar_wait(bar, 0);
  const float* W3 = tl + 2 * GX_POL_UNITS;
  const float* B3 = W3 + OUT * GX_POL_HID;
  const float* LS = B3 + OUT;
  // Groups: with GX_PR_GROUPS == 2 the block's warps 0-7 / 8-15 own envs 0-15 / 16-31 and run their steps HALF A STEP APART
  // (group 1 starts when group 0 has finished its first policy phase; after that nothing couples them: each group has
  // its own named barrier), so that one group's env phase -- latency-bound, light -- fills the issue slots that the
  // other group's layers -- register-file-bound -- leave.
  constexpr int NG = GX_PR_GROUPS, GW = (GX_PR_THREADS / 32) / NG, GE = GX_PR_ENVS / NG, SUBG = GE / EPW;
  static_assert(GE % EPW == 0 && 3 * SUBG < GW, "a group needs 3 * SUBG policy warps and one noise warp");
  const int grp = warp / GW, wl = warp % GW, e0 = grp * GE;
  auto group_sync = [&]() {
    if (NG == 1) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(GW * 32) : "memory");
  };
  // stagger: group g starts when group g - 1 has finished its first policy phase (barrier NG + g, both groups' threads)
  if (NG > 1 && grp > 0) asm volatile("bar.sync %0, %1;" ::"r"(NG + grp), "r"(2 * GW * 32) : "memory");
  for (int t = 0; t < T; ++t) {
    const size_t row = (size_t)t * (size_t)n;
    // the env state of this thread is dead weight during the policy phase: park it in shared memory so that the layers'
    // register tiles have the whole budget (128 registers at 512 threads; 12.4 -> 11.9 us per step)
    {
      const float v[18] = {d.qx, d.qy, d.qth, d.vx, d.vy, d.vw, d.hc, d.hs, d.prev_dist, d.steps, d.done, __uint_as_float((unsigned)d.age),
                           __uint_as_float((unsigned)(d.age >> 32)), gx_, gy_, ox, oy, 0.0f};
#pragma unroll
      for (int k = 0; k < 17; ++k) park[k * GX_PR_THREADS] = v[k];
    }
    // ---- policy phase ------------------------------------------------------------------------------------------------
    if (wl < 3 * SUBG) {
      const int net = wl / SUBG, r0 = e0 + (wl % SUBG) * EPW;      // this warp: net `net`, envs [r0, r0 + EPW) of the block
      float* h1n = h1 + ((size_t)net * GX_PR_ENVS + r0) * GX_POL_HS;
      float* h2n = h2 + ((size_t)net * GX_PR_ENVS + r0) * GX_POL_HS;
#ifndef GX_PR_SKIP_POLICY
      gx_policy_layer<K1, E>(w1s + (size_t)net * K1 * GX_POL_HID, tl + net * GX_POL_HID, xs + (size_t)r0 * K1, K1, h1n, lane);
      __syncwarp();
      gx_policy_layer<GX_POL_HID, E>(w2s + (size_t)net * GX_POL_HID * GX_POL_HID, tl + GX_POL_UNITS + net * GX_POL_HID, h1n, GX_POL_HS, h2n, lane);
      __syncwarp();
#endif
      // heads: lane = (action dimension, env of the warp's slice) for the actor warps, the first EPW lanes for the critics
      const int ll = lane % EPW, o = lane / EPW;
      const int le = r0 + ll, g = e_blk + le;
      const float4* hv = reinterpret_cast<const float4*>(h2n + (size_t)ll * GX_POL_HS);
      if (net == 0) {
        if (o < A) mu_s[2 * le + o] = gx_policy_head(W3, B3, hv, o);
      } else if (o == 0) {
        const float v = gx_policy_head(W3, B3, hv, A + net - 1);
        if (g < n) (net == 1 ? val : cval)[row + g] = v;
      }
    } else if (wl == 3 * SUBG && lane < GE) {
      // exploration noise of the group's env `lane` for this step: the two normals of Philox words (0, 1), as gx_policy_noise draws them
      const int le = e0 + lane, g = e_blk + le;
      float z0 = 0.0f, z1 = 0.0f;
      if (g < n) {
        const unsigned long long gid = p.env_id_offset + (unsigned long long)g;
        const gx_u32x4 r = gx_philox4x32_10((unsigned)gid, age_s[2 * le], age_s[2 * le + 1] ^ (unsigned)(gid >> 32), GX_STREAM_POLICY,
                                            p.seed_lo ^ k0, p.seed_hi ^ k1);
        const float u1 = 1.0f - gx_u01(r.v[0]);
        const float u2 = gx_u01(r.v[1]);
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        z0 = rad * cs; z1 = rad * sn;
      }
      z_s[2 * le] = z0; z_s[2 * le + 1] = z1;
    }
    group_sync();
    if (NG > 1 && grp + 1 < NG && t == 0) asm volatile("bar.arrive %0, %1;" ::"r"(NG + grp + 1), "r"(2 * GW * 32) : "memory");   // releases the next group
    {
      float v[17];
#pragma unroll
      for (int k = 0; k < 17; ++k) v[k] = park[k * GX_PR_THREADS];
      d.qx = v[0]; d.qy = v[1]; d.qth = v[2]; d.vx = v[3]; d.vy = v[4]; d.vw = v[5]; d.hc = v[6]; d.hs = v[7]; d.prev_dist = v[8];
      d.steps = v[9]; d.done = v[10];
      d.age = (uint64_t)__float_as_uint(v[11]) | ((uint64_t)__float_as_uint(v[12]) << 32);
      gx_ = v[13]; gy_ = v[14]; ox = v[15]; oy = v[16];
    }
    // ---- env phase: sample (all 16 lanes of the env compute the same numbers), step, outputs --------------------------------
    float av[2], lsum = 0.0f;
#pragma unroll
    for (int a = 0; a < A; ++a) {
      const float mu = mu_s[2 * el + a], z = z_s[2 * el + a];
      const float ls = LS[a], sd = expf(ls);
      av[a] = __fmaf_rn(sd, z, mu);
      const float dlt = av[a] - mu;
      lsum += -(dlt * dlt) / (2.0f * sd * sd) - ls - 0.91893853320467274f;
    }
#ifdef GX_PR_SKIP_ENV
    GxLaneStep r; r.obs.glidar = av[0]; r.obs.hlidar = av[1]; r.obs.cx = lsum; r.obs.cy = 0.f; r.c0 = r.c1 = r.c2 = 0.f; r.out.reward = r.out.done = r.out.cost = 0.f;
    d.age += 1;
#else
    const float* lay = gx_pick_layout(p, pool, nullptr, i, d.age + 1, GX_STREAM_PICK);
    const GxSpecLayout sp = gx_spec_layout(lay, sub);
    GxLaneStep r = gx_step_autoreset_lanes(p, s, i, sub, gmask, live, lay, sp, d, gx_, gy_, ox, oy, av[0], av[1], nullptr);
#endif
    gx_store_row_lanes(xs + (size_t)el * K1, sub, r.obs, r.c0, r.c1, r.c2, d);
    if (sub == 0) { age_s[2 * el] = (unsigned)d.age; age_s[2 * el + 1] = (unsigned)(d.age >> 32); }
    if (live) {
      const size_t k = row + (size_t)i;
      gx_store_row_lanes(obs + (k + (size_t)n) * 43, sub, r.obs, r.c0, r.c1, r.c2, d);
      if (sub == 3) reward[k] = r.out.reward;
      if (sub == 4) done[k] = r.out.done;
      if (sub == 5) cost[k] = r.out.cost;
      if (sub == 10) act[k] = make_float2(av[0], av[1]);
      if (sub == 11) logp[k] = lsum;
      if (STATS && sub == 9) gx_stats_step(s, i, r.out.reward, r.out.cost, r.out.done);
    }
    group_sync();
  }
  if (live) gx_store_dyn_lanes(s, i, sub, d);
}

#endif  // GX_POLICY_CUH_
