// Issue rate of fp64 / shuffle / fp32 instructions on one SM of a B200: W warps, each with ILP independent chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tput_bench tput_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP, int ILP>
__global__ void tput(double* outd, long long* cyc, double a) {
  double x[ILP], y = 1.0000001;
  float f[ILP], g = 0.99999f;
#pragma unroll
  for (int k = 0; k < ILP; ++k) { x[k] = a + threadIdx.x * 1e-9 + k; f[k] = (float)x[k]; }
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < N; ++i) {
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
      if (OP == 0) x[k] = fma(x[k], y, y);
      if (OP == 1) x[k] = x[k] + y;
      if (OP == 2) f[k] = fmaf(f[k], g, g);
      if (OP == 3) f[k] = __shfl_sync(0xffffffffu, f[k], (threadIdx.x + 31) & 31);
      if (OP == 4) { x[k] = x[k] * y; f[k] = __shfl_sync(0xffffffffu, f[k], (threadIdx.x + 31) & 31); }
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int k = 0; k < ILP; ++k) s += x[k] + f[k];
  outd[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* d; long long* c;
  cudaMalloc(&d, 1024 * 8); cudaMallocManaged(&c, 8);
  const char* names[] = {"DFMA", "DADD", "FFMA", "SHFL", "DMUL+SHFL"};
#define RUN(OP, W) { tput<OP, 8><<<1, 32 * W>>>(d, c, 1.0); tput<OP, 8><<<1, 32 * W>>>(d, c, 1.0); cudaDeviceSynchronize(); \
    printf("%-10s warps=%2d ILP=8: %6.2f cycles per warp-instruction per warp, %6.2f cycles per warp-instruction on the SM\n", names[OP], W, (double)c[0] / (N * 8.0 * (OP == 4 ? 2 : 1)), (double)c[0] / (N * 8.0 * W * (OP == 4 ? 2 : 1))); }
  RUN(0, 1) RUN(0, 4) RUN(0, 8) RUN(0, 16)
  RUN(1, 1) RUN(1, 4) RUN(1, 16)
  RUN(2, 1) RUN(2, 4) RUN(2, 16)
  RUN(3, 1) RUN(3, 4) RUN(3, 16)
  RUN(4, 1) RUN(4, 4)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
