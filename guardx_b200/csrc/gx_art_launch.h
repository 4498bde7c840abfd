// gx_art_launch.h -- host-side interface between guardx.cu (C ABI) and the per-robot
// translation units that hold the articulated kernels (internal to libguardx.so).
#ifndef GX_ART_LAUNCH_H_
#define GX_ART_LAUNCH_H_

#include <cuda_runtime.h>
#include <stdint.h>

enum { GX_ART_RESET = 0, GX_ART_STEP = 1, GX_ART_RESET_DONE = 2 };

struct gx_params;
struct GxState;
struct GxArtDev;

struct GxArtLaunch {
  const gx_params* p;
  const GxState* s;
  const GxArtDev* art;
  const float *pool, *layout, *action, *obs_in;
  float *obs, *act_out, *reward, *done, *cost, *qacc;
  int32_t* aux;
  int stats_on, mode /* GX_ART_RESET / STEP / RESET_DONE */, autoreset, sm_count;
  cudaStream_t stream;
};

cudaError_t gx_art_launch_ant(const GxArtLaunch& a);      // 4 legs x <= 2 dofs x <= 3 bodies, 4 lanes per env
cudaError_t gx_art_launch_walker(const GxArtLaunch& a);   // 2 legs x <= 5 dofs x <= 3 bodies, 2 lanes per env
cudaError_t gx_art_launch_swimmer(const GxArtLaunch& a);  // 1 chain x <= 2 dofs x <= 2 bodies, 1 lane per env

#endif  // GX_ART_LAUNCH_H_
