// Articulated step kernels for the swimmer-shaped robots (see gx_art.cuh / gx_art_kernels.cuh).
#include "gx_art_kernels.cuh"

cudaError_t gx_art_launch_swimmer(const GxArtLaunch& a) { return gx_art_launch_shape<1, 1, 2, 2, 4>(a); }
