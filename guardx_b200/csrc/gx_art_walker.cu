// Articulated step kernels for the walker-shaped robots (see gx_art.cuh / gx_art_kernels.cuh).
#include "gx_art_kernels.cuh"

#ifndef GX_WALKER_MINB
#define GX_WALKER_MINB 4
#endif
cudaError_t gx_art_launch_walker(const GxArtLaunch& a) { return gx_art_launch_shape<2, 2, 5, 3, GX_WALKER_MINB>(a); }
