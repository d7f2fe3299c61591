// Dependent-issue latency of the instructions a CTC lattice step is made of (one warp, B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_bench lat_bench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int OP>
__global__ void chain(double* outd, float* outf, long long* cyc, double a, float fa) {
  double x = a + threadIdx.x * 1e-9, y = 1.0000001;
  float f = fa + threadIdx.x * 1e-6f, g = 0.99999f;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = x + y;                                   // DADD
    if (OP == 1) x = x * y;                                   // DMUL
    if (OP == 2) x = fma(x, y, y);                            // DFMA
    if (OP == 3) f = f + g;                                   // FADD
    if (OP == 4) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f));
    if (OP == 5) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f));
    if (OP == 6) f = __shfl_up_sync(0xffffffffu, f, 1);
    if (OP == 7) { f = __shfl_up_sync(0xffffffffu, f, 1); f = f + g; }
    if (OP == 8) { x = __hiloint2double(__shfl_up_sync(0xffffffffu, __double2hiint(x), 1), __shfl_up_sync(0xffffffffu, __double2loint(x), 1)); x = x + y; x = x * y; }
    if (OP == 9) { f = fmaxf(f, g); g = fminf(g, f) + 1e-7f; }
    if (OP == 10) { x = x + y; x = x * y; }                   // DADD -> DMUL
    if (OP == 11) { f = __shfl_up_sync(0xffffffffu, f, 1); float m = fmaxf(f, g), lo = fminf(f, g); float e; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(lo - m)); float l; asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.f + e)); f = m + l + 0.001f; }   // log-domain step
    if (OP == 12) { double d = (double)f; x = x + d; f = (float)x; }   // F2F both ways + DADD
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[OP] = t1 - t0;
  outd[threadIdx.x] = x; outf[threadIdx.x] = f + g;
}
int main() {
  double* d; float* f; long long* c;
  cudaMalloc(&d, 32 * 8); cudaMalloc(&f, 32 * 4); cudaMallocManaged(&c, 16 * 8);
  const char* names[] = {"DADD", "DMUL", "DFMA", "FADD", "MUFU.EX2", "MUFU.LG2", "SHFL", "SHFL+FADD", "2xSHFL+DADD+DMUL",
                         "FMNMX x2 + FADD", "DADD+DMUL", "log-domain step (SHFL,max,min,sub,ex2,add,lg2,add,add)", "F2F.64.32+DADD+F2F.32.64"};
#define RUN(K) chain<K><<<1, 32>>>(d, f, c, 1.0, 0.5f); chain<K><<<1, 32>>>(d, f, c, 1.0, 0.5f);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12)
  cudaDeviceSynchronize();
  for (int k = 0; k < 13; ++k) printf("%-60s %7.1f cycles per iteration\n", names[k], (double)c[k] / N);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
