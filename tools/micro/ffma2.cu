// Micro-benchmark: issue rate of packed fp32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
template <int MODE, int CH>
__global__ void k(unsigned long long* out, long long* cyc, int iters) {
  unsigned long long a[CH]; float f[2 * CH];
  for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x + i; f[2 * i] = threadIdx.x + i; f[2 * i + 1] = i; }
  const unsigned long long m = 0x3f8000003f800000ull + threadIdx.x;
  const float fm = 1.0f + threadIdx.x * 1e-9f;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < CH; ++i) a[i] = fma2(a[i], m, a[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 2 * CH; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[i]) : "f"(f[i]), "f"(fm));
    }
  }
  long long t1 = clock64();
  unsigned long long s = 0; float fs = 0;
  for (int i = 0; i < CH; ++i) { s += a[i]; fs += f[2 * i] + f[2 * i + 1]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (unsigned long long)fs;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  unsigned long long* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int threads : {32, 128, 256, 512, 1024}) {
    long long c0, c1;
    k<0, 16><<<1, threads>>>(out, cyc, iters); cudaMemcpy(&c0, cyc, 8, cudaMemcpyDeviceToHost);
    k<1, 16><<<1, threads>>>(out, cyc, iters); cudaMemcpy(&c1, cyc, 8, cudaMemcpyDeviceToHost);
    const double w = threads / 32.0;
    printf("threads %4d (%.0f warps/SM): FFMA2 %.2f cycles per warp-instr per scheduler-share (%.1f FMA/clk/SM) | FFMA %.2f (%.1f FMA/clk/SM)\n", threads, w,
           (double)c0 / (iters * 16), w * 16 * iters * 64.0 / c0, (double)c1 / (iters * 32), w * 32 * iters * 32.0 / c1);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
