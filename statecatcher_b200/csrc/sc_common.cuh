// Shared device helpers for the statecatcher_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/statecatcher_b200.h"

#define SC_CHECK_ARG(cond, code) do { if (!(cond)) return (code); } while (0)
#define SC_LAUNCH_RET() do { cudaError_t e__ = cudaGetLastError(); return (int)e__; } while (0)

namespace sc {

typedef __nv_bfloat16 bf16;

__host__ __device__ inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

template <typename T> struct DT;
template <> struct DT<float> { static constexpr int code = SC_F32; };
template <> struct DT<bf16>  { static constexpr int code = SC_BF16; };

// ---- scalar load/store with conversion to/from fp32 -------------------------------
__device__ __forceinline__ float ld_f(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_f(const bf16* p) { return __bfloat162float(__ldg(p)); }
__device__ __forceinline__ void st_f(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_f(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

// ---- vector of N consecutive elements <-> float[N] ---------------------------------
// N*sizeof(T) must be 4, 8 or 16 bytes and the address aligned to it.
template <int BYTES> struct RawVec;
template <> struct RawVec<4>  { typedef uint32_t type; };
template <> struct RawVec<8>  { typedef uint2 type; };
template <> struct RawVec<16> { typedef uint4 type; };

template <typename T, int N> struct Vec {
  typedef typename RawVec<sizeof(T) * N>::type raw_t;
  raw_t raw;
};

// streaming (read-once) global load: read-only path, do not allocate in L1
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t* p) {
  uint32_t r; asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p)); return r;
}
__device__ __forceinline__ uint2 ldg_stream(const uint2* p) {
  uint2 r; asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p)); return r;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];"
                        : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
__device__ __forceinline__ void stg_stream(uint32_t* p, uint32_t v) {
  asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg_stream(uint2* p, uint2 v) {
  asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1,%2};" :: "l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void stg_stream(uint4* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename T, int N>
__device__ __forceinline__ Vec<T, N> vload(const T* p) {
  Vec<T, N> v;
  v.raw = ldg_stream(reinterpret_cast<const typename Vec<T, N>::raw_t*>(p));
  return v;
}
template <typename T, int N>
__device__ __forceinline__ void vstore(T* p, const Vec<T, N>& v) {
  stg_stream(reinterpret_cast<typename Vec<T, N>::raw_t*>(p), v.raw);
}

template <int N>
__device__ __forceinline__ void unpack(const Vec<float, N>& v, float (&f)[N]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v.raw);
#pragma unroll
  for (int i = 0; i < N; ++i) f[i] = __uint_as_float(w[i]);
}
template <int N>
__device__ __forceinline__ void unpack(const Vec<bf16, N>& v, float (&f)[N]) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&v.raw);
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    f[2 * i]     = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
template <int N>
__device__ __forceinline__ Vec<float, N> pack(const float (&f)[N], float*) {
  Vec<float, N> v;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v.raw);
#pragma unroll
  for (int i = 0; i < N; ++i) w[i] = __float_as_uint(f[i]);
  return v;
}
template <int N>
__device__ __forceinline__ Vec<bf16, N> pack(const float (&f)[N], bf16*) {
  Vec<bf16, N> v;
  uint32_t* w = reinterpret_cast<uint32_t*>(&v.raw);
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&h);
  }
  return v;
}

// ---- activations ------------------------------------------------------------------
// PRECISE=true : ex2/rcp based, abs error ~1e-7 (fp32 parity path, rtol 1e-4 contract)
// PRECISE=false: one MUFU.TANH each (abs error ~5e-4, below bf16 resolution of the outputs);
//                keeps the bf16 scan off the MUFU roofline (3 instead of 6 MUFU / element).
__device__ __forceinline__ float tanh_approx(float x) {
  float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
template <bool PRECISE> __device__ __forceinline__ float sigmoidf_(float x) {
  if (PRECISE) return __fdividef(1.0f, 1.0f + __expf(-x));
  return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f);
}
template <bool PRECISE> __device__ __forceinline__ float tanhf_(float x) {
  if (PRECISE) {
    // 1 - 2/(e^{2x}+1); saturates correctly for |x| large (e^{2x} -> inf or 0)
    float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
  }
  return tanh_approx(x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// log(sum_i exp(x_i)) of one row, computed by one warp (every lane gets the result).
// Rows whose 16-byte vectors fit four per lane (V <= 1024 bf16 / 512 fp32) are pulled into
// registers with all loads in flight at once, then max and exp-sum are two passes over
// registers — ~5 instructions per element instead of the online form's per-chunk rescaling.
template <typename T>
__device__ __forceinline__ float warp_row_lse(const T* __restrict__ x, int V, int lane) {
  constexpr int VW = 16 / (int)sizeof(T);
  const float NEGINF = -INFINITY;
  float m = NEGINF, ssum = 0.f;
  const bool vec_ok = (V % VW == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  if (vec_ok && V <= 4 * 32 * VW) {
    Vec<T, VW> raw[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = (k * 32 + lane) * VW;
      if (i < V) raw[k].raw = __ldg(reinterpret_cast<const uint4*>(x + i));
    }
    // only the RAW 16-byte words stay live across the two passes (16 registers instead of 16 + 32 unpacked
    // floats): these kernels are bound by bytes in flight per SM, i.e. by how many warps fit
    if constexpr (sizeof(T) == 2) {
      // bf16: the row maximum straight on the packed words (HMNMX2.BF16, 4 per 8 elements instead of
      // 8 unpacks + 8 FMNMX); the maximum of bf16 values is a bf16 value, so nothing is lost
      __nv_bfloat162 mm = __float2bfloat162_rn(NEGINF);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((k * 32 + lane) * VW < V) {
          const __nv_bfloat162* w = reinterpret_cast<const __nv_bfloat162*>(&raw[k].raw);
#pragma unroll
          for (int j = 0; j < 4; ++j) mm = __hmax2(mm, w[j]);
        }
      }
      m = fmaxf(__low2float(mm), __high2float(mm));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if ((k * 32 + lane) * VW < V) {
          float f[VW];
          unpack(raw[k], f);
#pragma unroll
          for (int j = 0; j < VW; ++j) m = fmaxf(m, f[j]);
        }
      }
    }
    m = warp_max(m);
    const float mc = (m == NEGINF) ? 0.f : m;                 // all -inf row: exp(-inf - 0) = 0, lse = -inf
    const float mc2 = mc * 1.4426950408889634f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if ((k * 32 + lane) * VW < V) {
        float f[VW];
        unpack(raw[k], f);
#pragma unroll
        for (int j = 0; j < VW; ++j) {                        // exp(f - m) as one FFMA + one MUFU (no range fix-ups: the argument is <= 0)
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(f[j], 1.4426950408889634f, -mc2)));
          ssum += e;
        }
      }
    }
    return mc + __logf(warp_sum(ssum));
  }
  if (vec_ok) {
    for (int i = lane * VW; i < V; i += 32 * VW) {
      float f[VW];
      Vec<T, VW> raw; raw.raw = __ldg(reinterpret_cast<const uint4*>(x + i));
      unpack(raw, f);
      float mm = f[0];
#pragma unroll
      for (int j = 1; j < VW; ++j) mm = fmaxf(mm, f[j]);
      const float nm = fmaxf(m, mm);
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < VW; ++j) acc += __expf(f[j] - nm);
      ssum = ssum * __expf(m - nm) + acc;
      m = nm;
    }
  } else {
    for (int i = lane; i < V; i += 32) {
      const float f = ld_f(x + i);
      const float nm = fmaxf(m, f);
      ssum = ssum * __expf(m - nm) + __expf(f - nm);
      m = nm;
    }
  }
  const float gm = warp_max(m);
  ssum = (m == NEGINF) ? 0.f : ssum * __expf(m - gm);
  return gm + __logf(warp_sum(ssum));
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace sc
