// General-tree (Doggo) kernels: see gx_tree.cuh / gx_tree_kernels.cuh.
#include "gx_tree_kernels.cuh"

#include <stdlib.h>

static int gx_tree_shared_limit() {
  static int limit = -1;
  if (limit < 0) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char* o = getenv("GUARDX_TREE_SHARED_LIMIT");   // tuning override (tools/bench_art_quick.py)
    limit = o ? atoi(o) : sms * 128;
  }
  return limit;
}

cudaError_t gx_tree_launch(const GxTreeLaunch& a) {
  // 32-env blocks; Hessian workspace in shared memory.  Up to one wave of four blocks per SM the mass matrix is in
  // shared memory too (lowest latency per env); beyond, it stays in the global scratch and seven blocks share an SM.
  const int n = a.p->n;
  const int mat_shared = n <= gx_tree_shared_limit();
  const int block = 32, grid = (n + block - 1) / block;
  const size_t sh = (size_t)block * GXT_SM_FLOATS_OF(a.t->nv, mat_shared) * sizeof(float);
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
int gx_tree_scratch_floats(void) { return (int)(GXT_SIZE + GX_TREE_MAXV * (GX_TREE_MAXV + 1) / 2); }
