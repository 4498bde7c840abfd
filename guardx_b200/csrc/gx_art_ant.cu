// Articulated step kernels for the ant-shaped robots (see gx_art.cuh / gx_art_kernels.cuh).
#include "gx_art_kernels.cuh"

#ifndef GX_ANT_MINB
#define GX_ANT_MINB 8
#endif
cudaError_t gx_art_launch_ant(const GxArtLaunch& a) { return gx_art_launch_shape<4, 4, 2, 3, GX_ANT_MINB>(a); }
