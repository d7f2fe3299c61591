// Fused gradient-clip + Adam/AdamW step: replaces, for the parameters of
// the path, torch.nn.utils.clip_grad_norm_ (train.py:553), the grad-norm logging loop with one
// `.item()` sync per parameter (train.py:555-560) and optimizer.step() (train.py:112-137,
// 563-566).  Two kernels per parameter tensor and no host synchronisation: the global norm stays
// on the device and the update kernel derives the clip coefficient from it.
#include "sc_common.cuh"

namespace sc {

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ acc) {
  float s = 0.f;
  const int64_t n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    s = fmaf(g[i], g[i], s);
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = red[threadIdx.x];
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
    if (threadIdx.x == 0) atomicAdd(acc, (double)t);
  }
}

// clip_coef = min(1, max_norm / (sqrt(sumsq) + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__device__ __forceinline__ float clip_coef(const double* sumsq, float max_norm) {
  if (sumsq == nullptr || max_norm <= 0.f) return 1.f;
  const float c = max_norm / ((float)sqrt(*sumsq) + 1e-6f);
  return c < 1.f ? c : 1.f;
}

__global__ void __launch_bounds__(256)
scale_grads_kernel(float* __restrict__ g, int64_t n, const double* __restrict__ sumsq, float max_norm) {
  const float c = clip_coef(sumsq, max_norm);
  if (c >= 1.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) g[i] *= c;
}

// decoupled = 1: AdamW (p *= 1 - lr*wd);  decoupled = 0: Adam with L2 (g += wd*p)
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                 const double* __restrict__ sumsq, float max_norm, int decoupled) {
  const float c = clip_coef(sumsq, max_norm);
  const float step = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pi = p[i];
    float gi = g[i] * c;
    if (decoupled) pi *= 1.f - lr * wd; else gi = fmaf(wd, pi, gi);
    const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    p[i] = pi - step * mi / (sqrtf(vi) / bc2_sqrt + eps);
  }
}

// Lion (Chen et al. 2023, "Symbolic Discovery of Optimization Algorithms", Algorithm 2; the rule the
// absent `lion_pytorch.Lion` of train.py:125-131 implements): decoupled decay, the SIGN of the
// beta1-interpolated momentum as the update, momentum tracked with beta2.  One state tensor.
__global__ void __launch_bounds__(256)
lion_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, int64_t n, float lr,
                 float b1, float b2, float wd, const double* __restrict__ sumsq, float max_norm) {
  const float c = clip_coef(sumsq, max_norm);
  const float keep = 1.f - lr * wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * c;
    const float mi = m[i];
    const float u = fmaf(b1, mi, (1.f - b1) * gi);
    const float sgn = (u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f);
    p[i] = fmaf(-lr, sgn, p[i] * keep);
    m[i] = fmaf(b2, mi, (1.f - b2) * gi);
  }
}

}  // namespace sc

using namespace sc;

static unsigned opt_grid(int64_t n) { return (unsigned)min((int64_t)148 * 8, cdiv(n, 1024) > 0 ? cdiv(n, 1024) : 1); }

extern "C" int sc_sumsq_accum(const float* g, int64_t n, double* acc, void* stream) {
  SC_CHECK_ARG(n >= 0 && acc, SC_E_BADARG);
  if (n == 0) return 0;
  SC_CHECK_ARG(g && ((uintptr_t)g & 15) == 0, SC_E_ALIGN);
  sumsq_kernel<<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(g, n, acc);
  SC_LAUNCH_RET();
}

extern "C" int sc_scale_grads(float* g, int64_t n, const double* sumsq, float max_norm, void* stream) {
  SC_CHECK_ARG(n >= 0 && sumsq, SC_E_BADARG);
  if (n == 0) return 0;
  SC_CHECK_ARG(g, SC_E_BADARG);
  scale_grads_kernel<<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(g, n, sumsq, max_norm);
  SC_LAUNCH_RET();
}

extern "C" int sc_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1,
                            float beta2, float eps, float weight_decay, int64_t step, const double* sumsq,
                            float max_norm, int decoupled, void* stream) {
  SC_CHECK_ARG(n >= 0 && step >= 1, SC_E_BADARG);
  if (n == 0) return 0;
  SC_CHECK_ARG(p && g && m && v, SC_E_BADARG);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  adam_step_kernel<<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1,
      bc2s, sumsq, max_norm, decoupled);
  SC_LAUNCH_RET();
}

extern "C" int sc_lion_step(float* p, const float* g, float* m, int64_t n, float lr, float beta1, float beta2,
                            float weight_decay, const double* sumsq, float max_norm, void* stream) {
  SC_CHECK_ARG(n >= 0, SC_E_BADARG);
  if (n == 0) return 0;
  SC_CHECK_ARG(p && g && m, SC_E_BADARG);
  lion_step_kernel<<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, n, lr, beta1, beta2, weight_decay, sumsq,
      max_norm);
  SC_LAUNCH_RET();
}
