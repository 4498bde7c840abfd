// General-tree (Doggo) kernels, one instantiation per translation unit: Engine.reset.
#include "gx_tree_kernels.cuh"

cudaError_t gx_tree_go_reset(const GxTreeLaunch& a, int grid, int block, size_t sh) {
  auto kern = gx_tree_kernel<GX_ART_RESET, false>;
  if (sh > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (e != cudaSuccess) return e;
  }
  // Small grids leave most of the 256 KB SM array to L1: the global scratch (constraint rows, body frames) is what
  // the dependent chains wait on, and with four blocks' worth of shared memory carved out only ~40 KB of L1 is left.
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, gx_tree_carveout(grid, sh));
  kern<<<grid, block, sh, a.stream>>>(*a.p, *a.s, *a.t, *a.host_model, a.pool, a.layout, a.action, a.obs_in, a.obs, a.act_out, a.reward,
                                      a.done, a.cost, a.qacc, a.aux, a.stats_on);
  return cudaGetLastError();
}
