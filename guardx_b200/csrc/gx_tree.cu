// General-tree (Doggo) kernels: see gx_tree.cuh / gx_tree_kernels.cuh.
#include "gx_tree_kernels.cuh"

cudaError_t gx_tree_launch(const GxTreeLaunch& a) {
  const int n = a.p->n, block = 64, grid = (n + block - 1) / block;
  const size_t sh = (size_t)block * a.p->obs_dim * sizeof(float);
#define GX_TREE_GO(MODE, AR)                                                                                  \
  do {                                                                                                        \
    auto kern = gx_tree_kernel<MODE, AR>;                                                                     \
    if (sh > 48 * 1024) {                                                                                     \
      cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);      \
      if (e_ != cudaSuccess) return e_;                                                                       \
    }                                                                                                         \
    kern<<<grid, block, sh, a.stream>>>(*a.p, *a.s, *a.t, a.pool, a.layout, a.action, a.obs_in, a.obs,        \
                                        a.act_out, a.reward, a.done, a.cost, a.qacc, a.aux, a.stats_on);      \
  } while (0)
  if (a.mode == GX_ART_RESET) GX_TREE_GO(GX_ART_RESET, false);
  else if (a.mode == GX_ART_RESET_DONE) GX_TREE_GO(GX_ART_RESET_DONE, false);
  else if (a.autoreset) GX_TREE_GO(GX_ART_STEP, true);
  else GX_TREE_GO(GX_ART_STEP, false);
#undef GX_TREE_GO
  return cudaGetLastError();
}

void gx_tree_get_state(const gx_params& p, const GxState& s, const GxTreeDev& t, float* out, cudaStream_t st) {
  gx_tree_get_state_kernel<<<(p.n + 127) / 128, 128, 0, st>>>(p, s, t, out);
}
void gx_tree_set_state(const gx_params& p, const GxState& s, const GxTreeDev& t, const float* in, cudaStream_t st) {
  gx_tree_set_state_kernel<<<(p.n + 127) / 128, 128, 0, st>>>(p, s, t, in);
}
int gx_tree_scratch_floats(void) { return (int)GXT_SIZE; }
