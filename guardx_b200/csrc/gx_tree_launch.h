// gx_tree_launch.h -- host-side interface between guardx.cu (C ABI) and gx_tree.cu (internal).
#ifndef GX_TREE_LAUNCH_H_
#define GX_TREE_LAUNCH_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/guardx_tree.h"

struct gx_params;
struct GxState;

struct GxTreeDev {
  const GxTreeModel* m;
  float* tq;       // [2 * (nq + 2 nv), cap]
  float* tw;       // [nworld + 2, cap] goal / hazard slide velocities, z history of ego_vel_acc
  float* scratch;  // one chunk [scratch_floats, 32] per 32-env block
  int nq, nv, nu;
  int scratch_floats;   // floats of global scratch per env (gx_tree_scratch_floats())
};

struct GxTreeLaunch {
  const gx_params* p;
  const GxState* s;
  const GxTreeDev* t;
  const GxTreeModel* host_model;   // passed to the kernels BY VALUE (11 KB of kernel parameters = constant bank)
  const float *pool, *layout, *action, *obs_in;
  float *obs, *act_out, *reward, *done, *cost, *qacc;
  int32_t* aux;
  int stats_on, mode, autoreset;
  cudaStream_t stream;
};
cudaError_t gx_tree_launch(const GxTreeLaunch& a);
// one kernel instantiation per translation unit (gx_tree_k*.cu), so that `make -j` compiles them in parallel
// shared-memory carve-out (percent of the SM array) that fits the blocks one SM will actually hold
static inline int gx_tree_carveout(int grid, size_t sh) {
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  int per_sm = (grid + sms - 1) / sms;
  const int fit = (int)((227 * 1024) / (sh + 1024));
  if (per_sm > fit) per_sm = fit;
  if (per_sm < 1) per_sm = 1;
  const int pct = (int)((100 * (size_t)per_sm * (sh + 1024) + 228 * 1024 - 1) / (228 * 1024));
  return pct > 100 ? 100 : pct;
}
cudaError_t gx_tree_go_reset(const GxTreeLaunch& a, int grid, int block, size_t sh);
cudaError_t gx_tree_go_reset_done(const GxTreeLaunch& a, int grid, int block, size_t sh);
cudaError_t gx_tree_go_step(const GxTreeLaunch& a, int grid, int block, size_t sh);
cudaError_t gx_tree_go_step_autoreset(const GxTreeLaunch& a, int grid, int block, size_t sh);

void gx_tree_get_state(const gx_params& p, const GxState& s, const GxTreeDev& t, float* out, cudaStream_t st);
void gx_tree_set_state(const gx_params& p, const GxState& s, const GxTreeDev& t, const float* in, cudaStream_t st);
int gx_tree_scratch_floats(void);

#endif  // GX_TREE_LAUNCH_H_
