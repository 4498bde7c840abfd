// General-tree (Doggo) kernels: see gx_tree.cuh / gx_tree_kernels.cuh.
#include "gx_tree_kernels.cuh"

#include <stdlib.h>

static int gx_tree_shared_limit() {
  static int limit = -1;
  if (limit < 0) {
    const char* o = getenv("GUARDX_TREE_SHARED_LIMIT");   // tuning override (tools/bench_art_quick.py)
    limit = o ? atoi(o) : 0x7fffffff;
  }
  return limit;
}

cudaError_t gx_tree_launch(const GxTreeLaunch& a) {
  // The mass matrix / factorisation workspace lives in shared memory (32-env blocks, 53.5 KB for Doggo, four per
  // SM).  Measured on B200: faster than the global-scratch build at every batch size (4096 envs 2.4 vs 5.4 ms,
  // 65 536 envs 10.7 vs 11.7 ms per step); GUARDX_TREE_SHARED_LIMIT=0 selects the global build for A/B runs.
  const int n = a.p->n;
  const int mat_shared = n <= gx_tree_shared_limit();
  const int block = mat_shared ? 32 : 64, grid = (n + block - 1) / block;
  const size_t sh = mat_shared ? (size_t)block * GXT_SM_FLOATS_OF(a.t->nv) * sizeof(float) : 0;
#define GX_TREE_GO(MODE, AR)                                                                                  \
  do {                                                                                                        \
    if (mat_shared) GX_TREE_GO2(MODE, AR, true); else GX_TREE_GO2(MODE, AR, false);                           \
  } while (0)
#define GX_TREE_GO2(MODE, AR, MSH)                                                                            \
  do {                                                                                                        \
    auto kern = gx_tree_kernel<MODE, AR, MSH>;                                                                \
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
#undef GX_TREE_GO2
  return cudaGetLastError();
}

void gx_tree_get_state(const gx_params& p, const GxState& s, const GxTreeDev& t, float* out, cudaStream_t st) {
  gx_tree_get_state_kernel<<<(p.n + 127) / 128, 128, 0, st>>>(p, s, t, out);
}
void gx_tree_set_state(const gx_params& p, const GxState& s, const GxTreeDev& t, const float* in, cudaStream_t st) {
  gx_tree_set_state_kernel<<<(p.n + 127) / 128, 128, 0, st>>>(p, s, t, in);
}
int gx_tree_scratch_floats(void) { return (int)(GXT_SIZE + GXT_SM_FLOATS); }
