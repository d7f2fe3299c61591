// Fused gradient-clip + Adam/AdamW step: replaces, for the parameters of
// the path, torch.nn.utils.clip_grad_norm_ (train.py:553), the grad-norm logging loop with one
// `.item()` sync per parameter (train.py:555-560) and optimizer.step() (train.py:112-137,
// 563-566).  Two kernels per parameter tensor and no host synchronisation: the global norm stays
// on the device and the update kernel derives the clip coefficient from it.
#include "sc_common.cuh"
#include <string.h>
#include <stdlib.h>

namespace sc {

// Sum of squares of g[tid::nth] over [0, n).  A gradient may start anywhere on a 4-byte boundary (a view into a
// bucket): the elements in front of the first 16-byte boundary and the n % 4 tail go through scalar loads, the
// body through float4.
__device__ __forceinline__ float sumsq_range(const float* __restrict__ g, int64_t n, int64_t tid, int64_t nth) {
  int64_t head = (int64_t)((16 - ((uintptr_t)g & 15)) & 15) / 4;
  if (head > n) head = n;
  const int64_t n4 = (n - head) / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g + head);
  float s = 0.f;
  for (int64_t i = tid; i < n4; i += nth) {
    const float4 v = g4[i];
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  for (int64_t i = tid; i < head; i += nth) s = fmaf(g[i], g[i], s);
  for (int64_t i = head + n4 * 4 + tid; i < n; i += nth) s = fmaf(g[i], g[i], s);
  return s;
}

// block of 256 threads: one double atomic per block
__device__ __forceinline__ void block_add_to(double* __restrict__ acc, float s) {
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

__global__ void __launch_bounds__(256)
sumsq_kernel(const float* __restrict__ g, int64_t n, double* __restrict__ acc) {
  block_add_to(acc, sumsq_range(g, n, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x));
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
struct AdamArgs { float lr, b1, b2, eps, wd, step, bc2_sqrt, c; int decoupled; };

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamArgs& a) {
  float pi = p;
  float gi = g * a.c;
  if (a.decoupled) pi *= 1.f - a.lr * a.wd; else gi = fmaf(a.wd, pi, gi);
  const float mi = fmaf(a.b1, m, (1.f - a.b1) * gi);
  const float vi = fmaf(a.b2, v, (1.f - a.b2) * gi * gi);
  m = mi; v = vi;
  p = pi - a.step * mi / (sqrtf(vi) / a.bc2_sqrt + a.eps);
}

// VEC = 1: all four arrays are 16-byte aligned, the body moves float4 words (seven 128-bit accesses per
// four elements) and the scalar loop only sees the n % 4 tail; VEC = 0: scalar throughout.
template <int VEC>
__device__ __forceinline__ void adam_range(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                           float* __restrict__ v, int64_t n, const AdamArgs& a, int64_t tid, int64_t nth) {
  int64_t done = 0;
  if (VEC) {
    const int64_t n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(p); const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
    for (int64_t i = tid; i < n4; i += nth) {
      float4 pp = p4[i], mm = m4[i], vv = v4[i];
      const float4 gg = g4[i];
      adam_elem(pp.x, gg.x, mm.x, vv.x, a); adam_elem(pp.y, gg.y, mm.y, vv.y, a);
      adam_elem(pp.z, gg.z, mm.z, vv.z, a); adam_elem(pp.w, gg.w, mm.w, vv.w, a);
      p4[i] = pp; m4[i] = mm; v4[i] = vv;
    }
    done = n4 * 4;
  }
  for (int64_t i = done + tid; i < n; i += nth) adam_elem(p[i], g[i], m[i], v[i], a);
}

template <int VEC>
__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 int64_t n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
                 const double* __restrict__ sumsq, float max_norm, int decoupled) {
  const AdamArgs a{lr, b1, b2, eps, wd, lr / bc1, bc2_sqrt, clip_coef(sumsq, max_norm), decoupled};
  adam_range<VEC>(p, g, m, v, n, a, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// Lion (Chen et al. 2023, "Symbolic Discovery of Optimization Algorithms", Algorithm 2; the rule the
// absent `lion_pytorch.Lion` of train.py:125-131 implements): decoupled decay, the SIGN of the
// beta1-interpolated momentum as the update, momentum tracked with beta2.  One state tensor.
struct LionArgs { float lr, b1, b2, keep, c; };

__device__ __forceinline__ void lion_elem(float& p, float g, float& m, const LionArgs& a) {
  const float gi = g * a.c;
  const float mi = m;
  const float u = fmaf(a.b1, mi, (1.f - a.b1) * gi);
  const float sgn = (u > 0.f) ? 1.f : ((u < 0.f) ? -1.f : 0.f);
  p = fmaf(-a.lr, sgn, p * a.keep);
  m = fmaf(a.b2, mi, (1.f - a.b2) * gi);
}

template <int VEC>
__device__ __forceinline__ void lion_range(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                           int64_t n, const LionArgs& a, int64_t tid, int64_t nth) {
  int64_t done = 0;
  if (VEC) {
    const int64_t n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(p); const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    for (int64_t i = tid; i < n4; i += nth) {
      float4 pp = p4[i], mm = m4[i];
      const float4 gg = g4[i];
      lion_elem(pp.x, gg.x, mm.x, a); lion_elem(pp.y, gg.y, mm.y, a);
      lion_elem(pp.z, gg.z, mm.z, a); lion_elem(pp.w, gg.w, mm.w, a);
      p4[i] = pp; m4[i] = mm;
    }
    done = n4 * 4;
  }
  for (int64_t i = done + tid; i < n; i += nth) lion_elem(p[i], g[i], m[i], a);
}

template <int VEC>
__global__ void __launch_bounds__(256)
lion_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, int64_t n, float lr,
                 float b1, float b2, float wd, const double* __restrict__ sumsq, float max_norm) {
  const LionArgs a{lr, b1, b2, 1.f - lr * wd, clip_coef(sumsq, max_norm)};
  lion_range<VEC>(p, g, m, n, a, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// ---- multi-tensor variants: up to SC_MT tensors per launch, the table travels in the kernel parameters
// (no device-side table to keep in step with gradients that zero_grad(set_to_none=True) re-allocates);
// every tensor gets blocks in proportion to its size (one per SC_MT_CHUNK elements, at least one): blk0 is the
// running block count, a block finds its tensor by walking it (<= 32 compares on kernel parameters) and then
// strides over that tensor with the bodies above, as one of the tensor's own blocks.
constexpr int SC_MT = 32;
constexpr int64_t SC_MT_CHUNK = 16384;
struct MultiDesc {
  float* p[SC_MT]; const float* g[SC_MT]; float* m[SC_MT]; float* v[SC_MT]; int64_t n[SC_MT];
  unsigned blk0[SC_MT + 1]; int count;
};

struct MultiSlot { int t; int64_t tid, nth; };
__device__ __forceinline__ MultiSlot multi_slot(const MultiDesc& d) {
  int t = 0;
  while (t + 1 < d.count && blockIdx.x >= d.blk0[t + 1]) ++t;
  const int64_t lb = blockIdx.x - d.blk0[t], nb = d.blk0[t + 1] - d.blk0[t];
  return MultiSlot{t, lb * blockDim.x + threadIdx.x, nb * blockDim.x};
}

__device__ __forceinline__ bool dev_aligned16(const void* a, const void* b, const void* c, const void* d) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}

__global__ void __launch_bounds__(256)
sumsq_multi_kernel(const __grid_constant__ MultiDesc d, double* __restrict__ acc) {
  const MultiSlot sl = multi_slot(d);
  block_add_to(acc, sumsq_range(d.g[sl.t], d.n[sl.t], sl.tid, sl.nth));
}

__global__ void __launch_bounds__(256)
adam_multi_kernel(const __grid_constant__ MultiDesc d, float lr, float b1, float b2, float eps, float wd, float bc1,
                  float bc2_sqrt, const double* __restrict__ sumsq, float max_norm, int decoupled) {
  const MultiSlot sl = multi_slot(d);
  const int t = sl.t;
  const AdamArgs a{lr, b1, b2, eps, wd, lr / bc1, bc2_sqrt, clip_coef(sumsq, max_norm), decoupled};
  const int64_t tid = sl.tid, nth = sl.nth;
  if (dev_aligned16(d.p[t], d.g[t], d.m[t], d.v[t])) adam_range<1>(d.p[t], d.g[t], d.m[t], d.v[t], d.n[t], a, tid, nth);
  else adam_range<0>(d.p[t], d.g[t], d.m[t], d.v[t], d.n[t], a, tid, nth);
}

__global__ void __launch_bounds__(256)
lion_multi_kernel(const __grid_constant__ MultiDesc d, float lr, float b1, float b2, float wd,
                  const double* __restrict__ sumsq, float max_norm) {
  const MultiSlot sl = multi_slot(d);
  const int t = sl.t;
  const LionArgs a{lr, b1, b2, 1.f - lr * wd, clip_coef(sumsq, max_norm)};
  const int64_t tid = sl.tid, nth = sl.nth;
  if (dev_aligned16(d.p[t], d.g[t], d.m[t], d.m[t])) lion_range<1>(d.p[t], d.g[t], d.m[t], d.n[t], a, tid, nth);
  else lion_range<0>(d.p[t], d.g[t], d.m[t], d.n[t], a, tid, nth);
}

// gradient (fp32) <-> slice of a flat communication buffer (fp32 or bf16), all tensors of a bucket in one launch.
// p[] carries the slice pointer, g[] the gradient.  DIRECTION 0: g -> slice, 1: slice*scale -> g.
template <typename TF, int DIRECTION>
__global__ void __launch_bounds__(256)
grads_flat_multi_kernel(const __grid_constant__ MultiDesc d, float scale) {
  const MultiSlot sl = multi_slot(d);
  const int t = sl.t;
  TF* flat = reinterpret_cast<TF*>(d.p[t]);
  float* g = const_cast<float*>(d.g[t]);
  for (int64_t i = sl.tid; i < d.n[t]; i += sl.nth) {
    if (DIRECTION == 0) st_f(flat + i, g[i]);
    else g[i] = scale * ld_f(flat + i);
  }
}

}  // namespace sc

using namespace sc;

static bool aligned16(const void* a, const void* b, const void* c, const void* d) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}
static unsigned opt_grid(int64_t n) { return (unsigned)min((int64_t)148 * 8, cdiv(n, 1024) > 0 ? cdiv(n, 1024) : 1); }

extern "C" int sc_sumsq_accum(const float* g, int64_t n, double* acc, void* stream) {
  SC_CHECK_ARG(n >= 0 && acc, SC_E_BADARG);
  if (n == 0) return 0;
  SC_CHECK_ARG(g && ((uintptr_t)g & 3) == 0, SC_E_ALIGN);
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
  // float4 body for 16-byte aligned tensors (B200, r02: per-tensor AdamW step of the cfg2 parameter set 0.69 -> 0.57 ms,
  // tests/test_gpu_zz_optim_ext.py green); SC_OPT_VEC=0 selects the scalar body (A/B runs).
  static const bool vec = [] { const char* e = getenv("SC_OPT_VEC"); return !(e && e[0] == '0'); }();
  if (vec && aligned16(p, g, m, v))
    adam_step_kernel<1><<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
        bc1, bc2s, sumsq, max_norm, decoupled);
  else
    adam_step_kernel<0><<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay,
        bc1, bc2s, sumsq, max_norm, decoupled);
  SC_LAUNCH_RET();
}

extern "C" int sc_lion_step(float* p, const float* g, float* m, int64_t n, float lr, float beta1, float beta2,
                            float weight_decay, const double* sumsq, float max_norm, void* stream) {
  SC_CHECK_ARG(n >= 0, SC_E_BADARG);
  if (n == 0) return 0;
  SC_CHECK_ARG(p && g && m, SC_E_BADARG);
  if (aligned16(p, g, m, m))
    lion_step_kernel<1><<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, n, lr, beta1, beta2, weight_decay, sumsq,
        max_norm);
  else
    lion_step_kernel<0><<<opt_grid(n), 256, 0, (cudaStream_t)stream>>>(p, g, m, n, lr, beta1, beta2, weight_decay, sumsq,
        max_norm);
  SC_LAUNCH_RET();
}

// ---- multi-tensor entry points.  p/g/m/v/n are HOST arrays (of device pointers / element counts).
template <class Launch>
static int multi_chunks(float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* n,
                        int64_t count, Launch&& launch) {
  for (int64_t base = 0; base < count;) {
    MultiDesc d;
    memset(&d, 0, sizeof(d));
    int k = 0;
    for (; base < count && k < SC_MT; ++base) {
      if (n[base] < 0) return SC_E_BADARG;
      if (n[base] == 0) continue;
      if (!g[base] || (p && !p[base]) || (m && !m[base]) || (v && !v[base])) return SC_E_BADARG;
      if (((uintptr_t)g[base] & 3) || (p && ((uintptr_t)p[base] & 3))) return SC_E_ALIGN;
      d.g[k] = g[base]; d.n[k] = n[base];
      d.p[k] = p ? p[base] : nullptr; d.m[k] = m ? m[base] : nullptr; d.v[k] = v ? v[base] : nullptr;
      ++k;
    }
    if (k == 0) continue;
    d.count = k;
    d.blk0[0] = 0;
    for (int i = 0; i < k; ++i) d.blk0[i + 1] = d.blk0[i] + (unsigned)cdiv(d.n[i], SC_MT_CHUNK);
    launch(d, dim3(d.blk0[k]));
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

extern "C" int sc_sumsq_accum_multi(const float* const* g, const int64_t* n, int64_t count, double* acc, void* stream) {
  SC_CHECK_ARG(count >= 0 && acc && (count == 0 || (g && n)), SC_E_BADARG);
  return multi_chunks(nullptr, g, nullptr, nullptr, n, count, [&](const MultiDesc& d, dim3 grid) {
    sumsq_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, acc);
  });
}

extern "C" int sc_adam_step_multi(float* const* p, const float* const* g, float* const* m, float* const* v,
                                  const int64_t* n, int64_t count, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, int64_t step, const double* sumsq, float max_norm, int decoupled,
                                  void* stream) {
  SC_CHECK_ARG(count >= 0 && step >= 1 && (count == 0 || (p && g && m && v && n)), SC_E_BADARG);
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  return multi_chunks(p, g, m, v, n, count, [&](const MultiDesc& d, dim3 grid) {
    adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, sumsq,
        max_norm, decoupled);
  });
}

extern "C" int sc_lion_step_multi(float* const* p, const float* const* g, float* const* m, const int64_t* n,
                                  int64_t count, float lr, float beta1, float beta2, float weight_decay,
                                  const double* sumsq, float max_norm, void* stream) {
  SC_CHECK_ARG(count >= 0 && (count == 0 || (p && g && m && n)), SC_E_BADARG);
  return multi_chunks(p, g, m, nullptr, n, count, [&](const MultiDesc& d, dim3 grid) {
    lion_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, lr, beta1, beta2, weight_decay, sumsq, max_norm);
  });
}

// Data-parallel gradient exchange (dp.StreamDataParallel): pack the fp32 gradients of a bucket into their slices of
// one flat communication buffer (fp32, or bf16 = half the NVLink payload), and the reverse with the averaging factor.
// g / flat_slices / n are HOST arrays (device pointers, element counts); one launch per 32 tensors.
extern "C" int sc_grads_pack_multi(const float* const* g, void* const* flat_slices, const int64_t* n, int64_t count,
                                   int flat_dtype, void* stream) {
  SC_CHECK_ARG(count >= 0 && (count == 0 || (g && flat_slices && n)), SC_E_BADARG);
  SC_CHECK_ARG(flat_dtype == SC_F32 || flat_dtype == SC_BF16, SC_E_DTYPE);
  return multi_chunks(reinterpret_cast<float* const*>(flat_slices), g, nullptr, nullptr, n, count, [&](const MultiDesc& d, dim3 grid) {
    if (flat_dtype == SC_F32) grads_flat_multi_kernel<float, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(d, 1.f);
    else grads_flat_multi_kernel<bf16, 0><<<grid, 256, 0, (cudaStream_t)stream>>>(d, 1.f);
  });
}

extern "C" int sc_grads_unpack_multi(float* const* g, const void* const* flat_slices, const int64_t* n, int64_t count,
                                     int flat_dtype, float scale, void* stream) {
  SC_CHECK_ARG(count >= 0 && (count == 0 || (g && flat_slices && n)), SC_E_BADARG);
  SC_CHECK_ARG(flat_dtype == SC_F32 || flat_dtype == SC_BF16, SC_E_DTYPE);
  return multi_chunks(reinterpret_cast<float* const*>(const_cast<void* const*>(flat_slices)), g, nullptr, nullptr, n, count,
                      [&](const MultiDesc& d, dim3 grid) {
    if (flat_dtype == SC_F32) grads_flat_multi_kernel<float, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(d, scale);
    else grads_flat_multi_kernel<bf16, 1><<<grid, 256, 0, (cudaStream_t)stream>>>(d, scale);
  });
}
