// General-tree (Doggo) kernels: see gx_tree.cuh / gx_tree_kernels.cuh.
#include "gx_tree_kernels.cuh"

cudaError_t gx_tree_launch(const GxTreeLaunch& a) {
  // 32-env blocks; Hessian workspace + mass matrix in shared memory (53.5 KB per block for Doggo, four blocks per SM)
  static_assert(GXT_DEVICE_MS == 32, "gx_tree kernels run one 32-lane warp per block");
  const int n = a.p->n, block = GXT_DEVICE_MS, grid = (n + block - 1) / block;
  const size_t sh = (size_t)block * GXT_SM_FLOATS_OF(a.t->nv, 1) * sizeof(float);
  if (a.mode == GX_ART_RESET) return gx_tree_go_reset(a, grid, block, sh);
  if (a.mode == GX_ART_RESET_DONE) return gx_tree_go_reset_done(a, grid, block, sh);
  return a.autoreset ? gx_tree_go_step_autoreset(a, grid, block, sh) : gx_tree_go_step(a, grid, block, sh);
}

// ---- state injection / checkpoint: one row per env =
//      qpos[nq] qvel[nv] qacc_warmstart[nv] prev_dist prev_done steps goal[2] hazards[2*nh] world_vel[2+2*nh]
__global__ void gx_tree_get_state_kernel(const __grid_constant__ gx_params p, const GxState s, const GxTreeDev a, float* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  float* o = out + (size_t)i * p.state_dim;
  const size_t cap = (size_t)s.cap;
  const int W = a.nq + 2 * a.nv;
  for (int k = 0; k < W; ++k) o[k] = a.tq[(size_t)k * cap + i];
  o += W;
  const float4 c = s.d2[i];
  o[0] = c.x; o[1] = (__float_as_uint(c.y) >> 31) ? 1.0f : 0.0f; o[2] = fabsf(c.y);
  const float2 g = s.goal[i];
  o[3] = g.x; o[4] = g.y;
  for (int h = 0; h < p.nh; ++h) gx_art_hazard(s, nullptr, i, h, o[5 + 2 * h], o[6 + 2 * h]);
  for (int k = 0; k < 2 + 2 * p.nh; ++k) o[5 + 2 * p.nh + k] = a.tw[(size_t)k * cap + i];
}

__global__ void gx_tree_set_state_kernel(const __grid_constant__ gx_params p, const GxState s, const GxTreeDev a, const float* in) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const float* o = in + (size_t)i * p.state_dim;
  const size_t cap = (size_t)s.cap;
  const int W = a.nq + 2 * a.nv;
  for (int k = 0; k < W; ++k) a.tq[(size_t)k * cap + i] = o[k];
  o += W;
  const float4 c = s.d2[i];
  s.d2[i] = make_float4(o[0], o[1] > 0.0f ? -o[2] : o[2], c.z, c.w);
  s.goal[i] = make_float2(o[3], o[4]);
  for (int k = 0; k < s.nh2; ++k) {
    float4 h;
    h.x = o[5 + 4 * k]; h.y = o[6 + 4 * k];
    h.z = (2 * k + 1 < p.nh) ? o[7 + 4 * k] : 0.0f;
    h.w = (2 * k + 1 < p.nh) ? o[8 + 4 * k] : 0.0f;
    s.hz[(size_t)k * cap + i] = h;
  }
  for (int k = 0; k < 2 + 2 * p.nh; ++k) a.tw[(size_t)k * cap + i] = o[5 + 2 * p.nh + k];
}


void gx_tree_get_state(const gx_params& p, const GxState& s, const GxTreeDev& t, float* out, cudaStream_t st) {
  gx_tree_get_state_kernel<<<(p.n + 127) / 128, 128, 0, st>>>(p, s, t, out);
}
void gx_tree_set_state(const gx_params& p, const GxState& s, const GxTreeDev& t, const float* in, cudaStream_t st) {
  gx_tree_set_state_kernel<<<(p.n + 127) / 128, 128, 0, st>>>(p, s, t, in);
}
int gx_tree_scratch_floats(void) { return (int)(GXT_SIZE + GX_TREE_MAXV * (GX_TREE_MAXV + 1) / 2); }
