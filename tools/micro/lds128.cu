// Micro-benchmark: shared-memory -> register bandwidth of LDS.128 / LDS.64 / LDS.32 on sm_100a for the access patterns of
// gx_policy_layer (quarter-warp broadcast: 8 lanes read one 16-byte quad; and 8 lanes reading 128 contiguous bytes).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds128 lds128.cu && ./lds128
#include <cstdio>
#include <cuda_runtime.h>
template <int W, int PAT>   // W: bytes per lane (4, 8, 16); PAT 0: quarter-warp broadcast (4 addresses per warp), 1: each quarter reads 8 x W contiguous bytes (same for all quarters), 2: all 32 lanes distinct
__global__ void k(float* out, long long* cyc, int iters) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int idx;
  if (PAT == 0) idx = (lane >> 3) * 44;            // floats: 4 rows 44 floats apart (the activation tile)
  else if (PAT == 1) idx = (lane & 7) * (W / 4);   // 8 lanes contiguous, quarters identical (the weight tile)
  else idx = lane * (W / 4);
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm + idx + (warp & 3) * 256);
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const unsigned a = base + ((it + u) & 7) * 1024;
      if (W == 16) { float x, y, z, w; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(a)); acc += x + y + z + w; }
      else if (W == 8) { float x, y; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(a)); acc += x + y; }
      else { float x; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(a)); acc += x; }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int W, int PAT>
void run(const char* name, float* out, long long* cyc) {
  const int iters = 2048;
  cudaFuncSetAttribute(k<W, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int threads : {32, 128, 192, 512}) {
    long long c;
    k<W, PAT><<<1, threads, 64 * 1024>>>(out, cyc, iters);
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double inst = (double)iters * 16 * (threads / 32);
    printf("%-34s %4d threads: %.2f cycles per warp-LDS (SM-wide), %.1f B/clk/SM into registers\n", name, threads, c / inst, inst * 32 * W / c);
  }
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  run<16, 0>("LDS.128 quarter-warp broadcast", out, cyc);
  run<16, 1>("LDS.128 8 lanes contiguous x4 same", out, cyc);
  run<16, 2>("LDS.128 32 lanes distinct", out, cyc);
  run<8, 0>("LDS.64 quarter-warp broadcast", out, cyc);
  run<4, 0>("LDS.32 quarter-warp broadcast", out, cyc);
  run<4, 2>("LDS.32 32 lanes distinct", out, cyc);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
