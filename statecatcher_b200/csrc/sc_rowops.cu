// Row-wise helpers around the projections: dtype cast, bias-gradient column sums and the
// LayerNorm forward/backward of lucyrnn.py:17-20 (nn.LayerNorm(H), eps=1e-5) used when
// config.layer_norm=True.  All HBM-bound, coalesced along the contiguous dimension.
#include "sc_common.cuh"
#include "sc_tma.cuh"
#include <stdlib.h>

namespace sc {

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ src, int64_t lds, TD* __restrict__ dst, int64_t ldd,
                            int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    st_f(dst + r * ldd + c, ld_f(src + r * lds + c));
  }
}

// four consecutive elements per thread (16-byte fp32 / 8-byte bf16 accesses, no 64-bit division per element):
// the weight and input casts of a training step (26 launches) were 0.49 ms of it at one element per thread
template <typename TS, typename TD>
__global__ void cast_vec4_kernel(const TS* __restrict__ src, int64_t lds, TD* __restrict__ dst, int64_t ldd,
                                 int64_t rows, int cols4) {
  const int64_t n = rows * cols4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols4;
    const int c = (int)(i - r * cols4) * 4;
    float f[4];
    unpack(vload<TS, 4>(src + r * lds + c), f);
    vstore<TD, 4>(dst + r * ldd + c, pack(f, (TD*)nullptr));
  }
}

// dst[r, :] = src[r, :] * float(mask[r])  —  model.py:377 `feats * mask.unsqueeze(-1).float()`
// (a true multiply, so -0.0 / NaN behave as upstream)
template <typename T>
__global__ void mask_rows_kernel(const T* __restrict__ src, int64_t lds, const uint8_t* __restrict__ mask,
                                 T* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    st_f(dst + r * ldd + c, ld_f(src + r * lds + c) * (mask[r] ? 1.f : 0.f));
  }
}

// hi = bf16(x), lo = bf16(x - hi): a two-term bf16 expansion (16 mantissa bits) of an fp32
// matrix, written side by side as [rows, 2*cols] so one bf16 tensor-core GEMM over the
// doubled reduction dimension reproduces the fp32 product to ~1e-5.
__global__ void split_bf16_kernel(const float* __restrict__ src, int64_t lds, bf16* __restrict__ dst, int64_t ldd,
                                  int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    const float x = src[r * lds + c];
    const bf16 hi = __float2bfloat16_rn(x);
    dst[r * ldd + c] = hi;
    dst[r * ldd + cols + c] = __float2bfloat16_rn(x - __bfloat162float(hi));
  }
}

// fp32 matrix -> SIX bf16 blocks for the tensor-core evaluation of an fp32 product.  x = hi + mid + lo to 2^-25 |x|
// (three bf16 terms = 24 mantissa bits), and
//   a.w = a_hi w_hi + a_hi w_mid + a_mid w_hi + a_hi w_lo + a_mid w_mid + a_lo w_hi   (+ terms below 2^-24 of the product)
// which ONE bf16 GEMM (fp32 accumulation in TMEM) over a six-fold reduction dimension computes when the blocks of the
// two operands are laid out in matching order: pattern 0 = [lo hi mid mid hi hi] (left operand),
// pattern 1 = [hi lo mid hi mid hi] (right operand).  (Two terms per operand — three products — were measured first:
// 1e-5 relative, five of the fp32 parity tests outside their rtol 1e-4 bounds; r02.)
// block_stride: element offset between consecutive blocks in dst (cols for blocks side by side in a row of width
// >= 6*cols; rows*ldd for blocks stacked vertically).
// four columns per thread (16-byte load, two 8-byte stores): the scalar form above took 24 us for the [5120 x 1032] fp32
// gradient of projection folding, 3x its traffic (r02)
__global__ void split_bf16_vec4_kernel(const float* __restrict__ src, int64_t lds, bf16* __restrict__ dst, int64_t ldd,
                                       int rows, int cols4) {
  const unsigned n = (unsigned)rows * (unsigned)cols4;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const unsigned r = i / (unsigned)cols4, c = (i - r * (unsigned)cols4) * 4u;
    const float4 x = __ldg(reinterpret_cast<const float4*>(src + (int64_t)r * lds + c));
    const __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
    const __nv_bfloat162 l0 = __floats2bfloat162_rn(x.x - __low2float(h0), x.y - __high2float(h0));
    const __nv_bfloat162 l1 = __floats2bfloat162_rn(x.z - __low2float(h1), x.w - __high2float(h1));
    uint2 hv, lv;
    hv.x = *reinterpret_cast<const uint32_t*>(&h0); hv.y = *reinterpret_cast<const uint32_t*>(&h1);
    lv.x = *reinterpret_cast<const uint32_t*>(&l0); lv.y = *reinterpret_cast<const uint32_t*>(&l1);
    bf16* d = dst + (int64_t)r * ldd + c;
    *reinterpret_cast<uint2*>(d) = hv;
    *reinterpret_cast<uint2*>(d + 4 * (int64_t)cols4) = lv;
  }
}
__global__ void split6_bf16_kernel(const float* __restrict__ src, int64_t lds, bf16* __restrict__ dst, int64_t ldd,
                                   int64_t rows, int64_t cols, int pattern, int64_t block_stride) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i % cols;
    const float x = src[r * lds + c];
    const bf16 hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);
    const bf16 mid = __float2bfloat16_rn(r1);
    const bf16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    bf16* d = dst + r * ldd + c;
    // ascending magnitude: lo.hi, hi.lo, mid.mid, mid.hi, hi.mid, hi.hi — the tensor core TRUNCATES when it adds an MMA's
    // products into the fp32 accumulator (loss 2^-24 |acc| per add, one-sided): with the small terms first, only the last
    // K/16 adds happen at full magnitude (measured r02, K = 1024: 2e-5 relative with hi.hi first)
    if (pattern == 0) {
      d[0] = lo; d[block_stride] = hi; d[2 * block_stride] = mid; d[3 * block_stride] = mid; d[4 * block_stride] = hi; d[5 * block_stride] = hi;
    } else {
      d[0] = hi; d[block_stride] = lo; d[2 * block_stride] = mid; d[3 * block_stride] = hi; d[4 * block_stride] = mid; d[5 * block_stride] = hi;
    }
  }
}

__global__ void zero_kernel(float* __restrict__ p, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = 0.f;
}

// out[n] += sum over a chunk of rows; one thread per column, coalesced across columns
// (r02, measured and dropped: a bf16 form with 16-byte loads — a thread owns 8 adjacent columns, 8 rows in flight, 592 blocks
// of 1024 columns — took 0.202 ms for the [192000 x 1024] dlogits of a cfg2 step against 0.128 ms for this kernel,
// profiles/r02_call78.sh; not investigated further: the pass is 0.4 % of the step.)
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ X, int64_t ldx, float* __restrict__ out,
                              int64_t M, int64_t N, int64_t rows_per_block) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int64_t m = m0;
  for (; m + 3 < m1; m += 4) {
    a0 += ld_f(X + m * ldx + n); a1 += ld_f(X + (m + 1) * ldx + n);
    a2 += ld_f(X + (m + 2) * ldx + n); a3 += ld_f(X + (m + 3) * ldx + n);
  }
  for (; m < m1; ++m) a0 += ld_f(X + m * ldx + n);
  atomicAdd(out + n, (a0 + a1) + (a2 + a3));
}

constexpr int LN_WARPS = 4;
constexpr float LN_EPS = 1e-5f;

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const T* __restrict__ X, int64_t ldx, const float* __restrict__ w,
                     const float* __restrict__ b, T* __restrict__ Y, int64_t ldy,
                     float* __restrict__ mean, float* __restrict__ rstd, int64_t M, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const T* x = X + row * ldx;
  float s = 0.f;
  for (int i = lane; i < H; i += 32) s += ld_f(x + i);
  const float mu = warp_sum(s) / (float)H;
  float v = 0.f;
  for (int i = lane; i < H; i += 32) { const float d = ld_f(x + i) - mu; v = fmaf(d, d, v); }
  const float rs = rsqrtf(warp_sum(v) / (float)H + LN_EPS);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  T* y = Y + row * ldy;
  for (int i = lane; i < H; i += 32) st_f(y + i, (ld_f(x + i) - mu) * rs * w[i] + b[i]);
}

// Backward: dx per row (two warp reductions) and dw/db column sums.  Each lane owns the same
// columns (lane + 32*j) for every row it visits, so the column sums live in registers (NJ per
// lane) and touch shared/global memory once per block — the first version used two shared-memory
// atomics per element and ran at 0.8 TB/s.  NJ = ceil(H/32) <= LN_MAXJ; larger H takes the
// generic atomic kernel below.
constexpr int LN_MAXJ = 32;

template <typename T, int NJ>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_reg_kernel(const T* __restrict__ dY, int64_t lddy, const T* __restrict__ X, int64_t ldx,
                         const float* __restrict__ w, const float* __restrict__ mean,
                         const float* __restrict__ rstd, T* __restrict__ dX, int64_t lddx,
                         float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dxs, int64_t M, int H,
                         int64_t rows_per_block) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  float aw[NJ], ab[NJ], ax[NJ], wv[NJ];             // ax: column sums of dX (the bias gradient of the projection that produced X)
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    aw[j] = 0.f; ab[j] = 0.f; ax[j] = 0.f;
    const int i = lane + 32 * j;
    wv[j] = (i < H) ? w[i] : 0.f;
  }
  for (int64_t row = m0 + warp; row < m1; row += LN_WARPS) {
    const T* x = X + row * ldx;
    const T* dy = dY + row * lddy;
    const float mu = mean[row], rs = rstd[row];
    float dyv[NJ], xh[NJ];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int i = lane + 32 * j;
      dyv[j] = (i < H) ? ld_f(dy + i) : 0.f;
      xh[j] = (i < H) ? (ld_f(x + i) - mu) * rs : 0.f;
      const float g = dyv[j] * wv[j];
      s1 += g; s2 = fmaf(g, xh[j], s2);
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    T* dx = dX + row * lddx;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int i = lane + 32 * j;
      const float o = rs * (dyv[j] * wv[j] - s1 - xh[j] * s2);
      if (i < H) { st_f(dx + i, o); ax[j] += o; }
      aw[j] = fmaf(dyv[j], xh[j], aw[j]);
      ab[j] += dyv[j];
    }
  }
  // block partials through shared memory, then one global atomic per column per block
  extern __shared__ float sm[];
  float* sdw = sm;
  float* sdb = sm + H;
  float* sdx = sm + 2 * H;
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int i = lane + 32 * j;
    if (i < H) { atomicAdd(sdw + i, aw[j]); atomicAdd(sdb + i, ab[j]); if (dxs) atomicAdd(sdx + i, ax[j]); }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dw + i, sdw[i]);
    atomicAdd(db + i, sdb[i]);
    if (dxs) atomicAdd(dxs + i, sdx[i]);
  }
}

// Vectorised variant for H % 256 == 0: a lane owns 8 consecutive columns per 256-column
// group (16-byte loads for bf16, 2 x 16 for fp32), NV = H/256 groups.
template <typename T> __device__ __forceinline__ void ld8(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void ld8<bf16>(const bf16* p, float (&f)[8]) {
  Vec<bf16, 8> v; v.raw = __ldg(reinterpret_cast<const uint4*>(p)); unpack(v, f);
}
template <> __device__ __forceinline__ void ld8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <typename T> __device__ __forceinline__ void st8(T* p, const float (&f)[8]);
template <> __device__ __forceinline__ void st8<bf16>(bf16* p, const float (&f)[8]) {
  const Vec<bf16, 8> v = pack(f, (bf16*)nullptr);
  *reinterpret_cast<uint4*>(p) = v.raw;
}
template <> __device__ __forceinline__ void st8<float>(float* p, const float (&f)[8]) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// Vectorised forward for H % 256 == 0, H <= 1024 (same lane/column map as the backward below): the row is read once
// with 16-byte loads and stays in registers for the mean, the variance and the output (the kernel above reads it
// three times with 2-byte loads: 0.24 ms per [192000 x 1024] bf16 call against 0.12 ms of traffic).
template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_vec_kernel(const T* __restrict__ X, int64_t ldx, const float* __restrict__ w,
                         const float* __restrict__ b, T* __restrict__ Y, int64_t ldy,
                         float* __restrict__ mean, float* __restrict__ rstd, int64_t M, int H) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
  if (row >= M) return;
  const T* x = X + row * ldx;
  float xv[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j) ld8<T>(x + (lane + 32 * j) * 8, xv[j]);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) s += xv[j][e];
  const float mu = warp_sum(s) / (float)H;
  float v = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) { xv[j][e] -= mu; v = fmaf(xv[j][e], xv[j][e], v); }
  const float rs = rsqrtf(warp_sum(v) / (float)H + LN_EPS);
  if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
  T* y = Y + row * ldy;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float wv[8], bv[8], o[8];
    ld8<float>(w + (lane + 32 * j) * 8, wv);
    ld8<float>(b + (lane + 32 * j) * 8, bv);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = xv[j][e] * rs * wv[e] + bv[e];
    st8<T>(y + (lane + 32 * j) * 8, o);
  }
}

// (r02, measured and dropped: a persistent form of the forward — 3-4 blocks per SM, each warp walking rows W apart with
// the rows of the next two visits requested ahead as raw 16-byte words, 128 registers — is bit-identical and no faster:
// 3.54-3.56 ms for the 18 calls of a cfg2 LayerNorm step against 3.47-3.52 ms, alternating same-box runs,
// profiles/r02_call69.sh.  Bytes in flight are not what holds this kernel at ~4 TB/s inside the power-capped step.)
// (r02, measured and dropped: reading the row twice — a first sweep for the two row sums, a second for dX and the
// column sums — with the gain vector re-read per row brings the kernel from 255 to 168 registers and 3 blocks per SM
// instead of 2, and is SLOWER: 18 calls of the cfg2 LayerNorm step 7.6 ms against 6.1 ms.)
template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_vec_kernel(const T* __restrict__ dY, int64_t lddy, const T* __restrict__ X, int64_t ldx,
                         const float* __restrict__ w, const float* __restrict__ mean,
                         const float* __restrict__ rstd, T* __restrict__ dX, int64_t lddx,
                         float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dxs, int64_t M, int H,
                         int64_t rows_per_block) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  float aw[NV][8], ab[NV][8], ax[NV][8], wv[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      aw[j][e] = 0.f; ab[j][e] = 0.f; ax[j][e] = 0.f;
      wv[j][e] = w[(lane + 32 * j) * 8 + e];
    }
  for (int64_t row = m0 + warp; row < m1; row += LN_WARPS) {
    const T* x = X + row * ldx;
    const T* dy = dY + row * lddy;
    const float mu = mean[row], rs = rstd[row];
    float dyv[NV][8], xh[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      ld8<T>(dy + (lane + 32 * j) * 8, dyv[j]);
      ld8<T>(x + (lane + 32 * j) * 8, xh[j]);
    }
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        xh[j][e] = (xh[j][e] - mu) * rs;
        const float g = dyv[j][e] * wv[j][e];
        s1 += g; s2 = fmaf(g, xh[j][e], s2);
      }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    T* dx = dX + row * lddx;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o[e] = rs * (dyv[j][e] * wv[j][e] - s1 - xh[j][e] * s2);
        aw[j][e] = fmaf(dyv[j][e], xh[j][e], aw[j][e]);
        ab[j][e] += dyv[j][e];
        ax[j][e] += o[e];
      }
      st8<T>(dx + (lane + 32 * j) * 8, o);
    }
  }
  extern __shared__ float sm[];
  float* sdw = sm;
  float* sdb = sm + H;
  float* sdx = sm + 2 * H;
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(sdw + (lane + 32 * j) * 8 + e, aw[j][e]);
      atomicAdd(sdb + (lane + 32 * j) * 8 + e, ab[j][e]);
      if (dxs) atomicAdd(sdx + (lane + 32 * j) * 8 + e, ax[j][e]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dw + i, sdw[i]);
    atomicAdd(db + i, sdb[i]);
    if (dxs) atomicAdd(dxs + i, sdx[i]);
  }
}

// Staged form of the kernel above: what bounds it is BYTES IN FLIGHT — 255 registers allow 8 warps per SM, each with one
// 4 KB row pair in flight: 32 KB per SM over ~1.3 us of latency = 3.6 TB/s, the measured rate (and a warp-pair form
// with 12 warps of half rows, 24 KB in flight, measured slower: r02).  Here every warp owns a ring of LN_RING rows in
// shared memory filled by bulk async copies issued LN_RING-1 rows ahead, so the bytes in flight no longer cost registers.
// (Reading the staged row twice instead of holding it, 167 registers and three blocks per SM: 5.7-5.8 vs 5.5-5.6 ms. Not kept.)
#ifndef SC_LN_RING
#define SC_LN_RING 3
#endif
constexpr int LN_RING = SC_LN_RING;
template <typename T, int NV>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_stage_kernel(const T* __restrict__ dY, int64_t lddy, const T* __restrict__ X, int64_t ldx,
                           const float* __restrict__ w, const float* __restrict__ mean,
                           const float* __restrict__ rstd, T* __restrict__ dX, int64_t lddx,
                           float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dxs, int64_t M, int H,
                           int64_t rows_per_block) {
  constexpr int ROWB = NV * 256 * (int)sizeof(T);       // bytes of one row of one tensor
  extern __shared__ __align__(128) uint8_t dyn[];       // [3H floats: block partials][per warp: LN_RING x {dy row, x row}]
  __shared__ __align__(8) uint64_t bars[LN_WARPS][LN_RING];
  float* sm = reinterpret_cast<float*>(dyn);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* ring = dyn + ((3 * (size_t)H * sizeof(float) + 127) & ~(size_t)127) + (size_t)warp * LN_RING * 2 * ROWB;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  const int64_t first = m0 + warp;
  const int nk = first < m1 ? (int)((m1 - first + LN_WARPS - 1) / LN_WARPS) : 0;   // rows of this warp
  if (lane == 0) {
    for (int i = 0; i < LN_RING; ++i) mbar_init(smem_u32(&bars[warp][i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](int k) {                              // lane 0 only
    const int slot = k % LN_RING;
    const int64_t row = first + (int64_t)k * LN_WARPS;
    const uint32_t bar = smem_u32(&bars[warp][slot]);
    mbar_expect_tx(bar, 2 * ROWB);
    bulk_load_1d(smem_u32(ring + (size_t)slot * 2 * ROWB), dY + row * lddy, ROWB, bar);
    bulk_load_1d(smem_u32(ring + (size_t)slot * 2 * ROWB + ROWB), X + row * ldx, ROWB, bar);
  };
  if (lane == 0)
    for (int k = 0; k < LN_RING - 1 && k < nk; ++k) issue(k);
  float aw[NV][8], ab[NV][8], ax[NV][8], wv[NV][8];
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      aw[j][e] = 0.f; ab[j][e] = 0.f; ax[j][e] = 0.f;
      wv[j][e] = w[(lane + 32 * j) * 8 + e];
    }
  float mu_n = nk > 0 ? mean[first] : 0.f, rs_n = nk > 0 ? rstd[first] : 0.f;
  for (int k = 0; k < nk; ++k) {
    const int64_t row = first + (int64_t)k * LN_WARPS;
    __syncwarp();                                        // every lane has left the slot the next copy overwrites
    if (lane == 0 && k + LN_RING - 1 < nk) issue(k + LN_RING - 1);
    const float mu = mu_n, rs = rs_n;
    if (k + 1 < nk) { mu_n = mean[row + LN_WARPS]; rs_n = rstd[row + LN_WARPS]; }   // one row ahead
    mbar_wait(smem_u32(&bars[warp][k % LN_RING]), (uint32_t)((k / LN_RING) & 1));
    const T* dy = reinterpret_cast<const T*>(ring + (size_t)(k % LN_RING) * 2 * ROWB);
    const T* x = reinterpret_cast<const T*>(ring + (size_t)(k % LN_RING) * 2 * ROWB + ROWB);
    float dyv[NV][8], xh[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if constexpr (sizeof(T) == 2) {
        Vec<T, 8> a, b;
        a.raw = *reinterpret_cast<const uint4*>(dy + (lane + 32 * j) * 8);
        b.raw = *reinterpret_cast<const uint4*>(x + (lane + 32 * j) * 8);
        unpack(a, dyv[j]); unpack(b, xh[j]);
      } else {
        const float4* pa = reinterpret_cast<const float4*>(dy + (lane + 32 * j) * 8);
        const float4* pb = reinterpret_cast<const float4*>(x + (lane + 32 * j) * 8);
        const float4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];
        dyv[j][0] = a0.x; dyv[j][1] = a0.y; dyv[j][2] = a0.z; dyv[j][3] = a0.w; dyv[j][4] = a1.x; dyv[j][5] = a1.y; dyv[j][6] = a1.z; dyv[j][7] = a1.w;
        xh[j][0] = b0.x; xh[j][1] = b0.y; xh[j][2] = b0.z; xh[j][3] = b0.w; xh[j][4] = b1.x; xh[j][5] = b1.y; xh[j][6] = b1.z; xh[j][7] = b1.w;
      }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        xh[j][e] = (xh[j][e] - mu) * rs;
        const float g = dyv[j][e] * wv[j][e];
        s1 += g; s2 = fmaf(g, xh[j][e], s2);
      }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    T* dx = dX + row * lddx;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        o[e] = rs * (dyv[j][e] * wv[j][e] - s1 - xh[j][e] * s2);
        aw[j][e] = fmaf(dyv[j][e], xh[j][e], aw[j][e]);
        ab[j][e] += dyv[j][e];
        ax[j][e] += o[e];
      }
      st8<T>(dx + (lane + 32 * j) * 8, o);
    }
  }
  float* sdw = sm;
  float* sdb = sm + H;
  float* sdx = sm + 2 * H;
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NV; ++j)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(sdw + (lane + 32 * j) * 8 + e, aw[j][e]);
      atomicAdd(sdb + (lane + 32 * j) * 8 + e, ab[j][e]);
      if (dxs) atomicAdd(sdx + (lane + 32 * j) * 8 + e, ax[j][e]);
    }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dw + i, sdw[i]);
    atomicAdd(db + i, sdb[i]);
    if (dxs) atomicAdd(dxs + i, sdx[i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const T* __restrict__ dY, int64_t lddy, const T* __restrict__ X, int64_t ldx,
                     const float* __restrict__ w, const float* __restrict__ mean,
                     const float* __restrict__ rstd, T* __restrict__ dX, int64_t lddx,
                     float* __restrict__ dw, float* __restrict__ db, float* __restrict__ dxs, int64_t M, int H,
                     int64_t rows_per_block) {
  extern __shared__ float sm[];        // dw[H], db[H], dxsum[H] block partials
  float* sdw = sm;
  float* sdb = sm + H;
  float* sdx = sm + 2 * H;
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  for (int64_t row = m0 + warp; row < m1; row += LN_WARPS) {
    const T* x = X + row * ldx;
    const T* dy = dY + row * lddy;
    const float mu = mean[row], rs = rstd[row];
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < H; i += 32) {
      const float g = ld_f(dy + i) * w[i];
      const float xh = (ld_f(x + i) - mu) * rs;
      s1 += g; s2 = fmaf(g, xh, s2);
    }
    s1 = warp_sum(s1) / (float)H;
    s2 = warp_sum(s2) / (float)H;
    T* dx = dX + row * lddx;
    for (int i = lane; i < H; i += 32) {
      const float dyi = ld_f(dy + i);
      const float xh = (ld_f(x + i) - mu) * rs;
      const float o = rs * (dyi * w[i] - s1 - xh * s2);
      st_f(dx + i, o);
      atomicAdd(sdw + i, dyi * xh);
      atomicAdd(sdb + i, dyi);
      if (dxs) atomicAdd(sdx + i, o);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    atomicAdd(dw + i, sdw[i]);
    atomicAdd(db + i, sdb[i]);
    if (dxs) atomicAdd(dxs + i, sdx[i]);
  }
}

template <typename T>
static void launch_ln_bwd(const void* dY, int64_t lddy, const void* X, int64_t ldx, const float* w, const float* mean,
                          const float* rstd, void* dX, int64_t lddx, float* dw, float* db, float* dxs, int64_t M, int H,
                          int64_t blocks, int64_t rpb, size_t smem, cudaStream_t st) {
#define SC_LN_BWD(NJ) do { \
    if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_reg_kernel<T, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    layernorm_bwd_reg_kernel<T, NJ><<<(unsigned)blocks, LN_WARPS * 32, smem, st>>>((const T*)dY, lddy, (const T*)X, ldx, w, mean, rstd, \
        (T*)dX, lddx, dw, db, dxs, M, H, rpb); } while (0)
  const bool vec_ok = (H % 256 == 0) && H <= 1024 && (lddy % 8 == 0) && (ldx % 8 == 0) && (lddx % 8 == 0) &&
                      aligned16(dY) && aligned16(X) && aligned16(dX);
#define SC_LN_BWD_V(NV) do { \
    if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_vec_kernel<T, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    layernorm_bwd_vec_kernel<T, NV><<<(unsigned)blocks, LN_WARPS * 32, smem, st>>>((const T*)dY, lddy, (const T*)X, ldx, w, mean, rstd, \
        (T*)dX, lddx, dw, db, dxs, M, H, rpb); } while (0)
  static const bool stage_off = [] { const char* e = getenv("SC_LN_BWD_STAGE"); return e && e[0] == '0'; }();   // A/B switch
#define SC_LN_BWD_S(NV) do { \
    const size_t sm2 = ((smem + 127) & ~(size_t)127) + (size_t)LN_WARPS * LN_RING * 2 * (NV) * 256 * sizeof(T); \
    if (sm2 > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_stage_kernel<T, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2); \
    layernorm_bwd_stage_kernel<T, NV><<<(unsigned)blocks, LN_WARPS * 32, sm2, st>>>((const T*)dY, lddy, (const T*)X, ldx, w, mean, rstd, \
        (T*)dX, lddx, dw, db, dxs, M, H, rpb); } while (0)
  if (vec_ok && !stage_off) {
    switch (H / 256) {
      case 1: SC_LN_BWD_S(1); return;
      case 2: SC_LN_BWD_S(2); return;
      case 3: SC_LN_BWD_S(3); return;
      default: SC_LN_BWD_S(4); return;
    }
  }
#undef SC_LN_BWD_S
  if (vec_ok) {
    switch (H / 256) {
      case 1: SC_LN_BWD_V(1); return;
      case 2: SC_LN_BWD_V(2); return;
      case 3: SC_LN_BWD_V(3); return;
      default: SC_LN_BWD_V(4); return;
    }
  }
#undef SC_LN_BWD_V
  const int nj = (H + 31) / 32;
  if (nj <= 2) SC_LN_BWD(2);
  else if (nj <= 8) SC_LN_BWD(8);
  else if (nj <= 16) SC_LN_BWD(16);
  else if (nj <= LN_MAXJ) SC_LN_BWD(32);
  else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(layernorm_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    layernorm_bwd_kernel<T><<<(unsigned)blocks, LN_WARPS * 32, smem, st>>>((const T*)dY, lddy, (const T*)X, ldx, w, mean, rstd,
        (T*)dX, lddx, dw, db, dxs, M, H, rpb);
  }
#undef SC_LN_BWD
}

}  // namespace sc

using namespace sc;

extern "C" int sc_cast(const void* src, int64_t lds, int src_dtype, void* dst, int64_t ldd, int dst_dtype,
                       int64_t rows, int64_t cols, void* stream) {
  SC_CHECK_ARG(rows >= 0 && cols >= 0, SC_E_BADARG);
  if (rows * cols == 0) return 0;
  SC_CHECK_ARG(src && dst, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const int ssz = src_dtype == SC_F32 ? 4 : 2, dsz = dst_dtype == SC_F32 ? 4 : 2;
  if ((src_dtype == SC_F32 || src_dtype == SC_BF16) && (dst_dtype == SC_F32 || dst_dtype == SC_BF16) && cols % 4 == 0 &&
      lds % 4 == 0 && ldd % 4 == 0 && ((uintptr_t)src % (4 * ssz)) == 0 && ((uintptr_t)dst % (4 * dsz)) == 0 && cols < ((int64_t)1 << 31)) {
    const int c4 = (int)(cols / 4);
    const unsigned vb = (unsigned)min((int64_t)148 * 16, cdiv(rows * c4, 256));
    if (src_dtype == SC_F32 && dst_dtype == SC_BF16) cast_vec4_kernel<float, bf16><<<vb, 256, 0, st>>>((const float*)src, lds, (bf16*)dst, ldd, rows, c4);
    else if (src_dtype == SC_BF16 && dst_dtype == SC_F32) cast_vec4_kernel<bf16, float><<<vb, 256, 0, st>>>((const bf16*)src, lds, (float*)dst, ldd, rows, c4);
    else if (src_dtype == SC_F32) cast_vec4_kernel<float, float><<<vb, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, rows, c4);
    else cast_vec4_kernel<bf16, bf16><<<vb, 256, 0, st>>>((const bf16*)src, lds, (bf16*)dst, ldd, rows, c4);
    SC_LAUNCH_RET();
  }
  const unsigned blocks = (unsigned)min((int64_t)148 * 16, cdiv(rows * cols, 256));
  if (src_dtype == SC_F32 && dst_dtype == SC_BF16)
    cast_kernel<float, bf16><<<blocks, 256, 0, st>>>((const float*)src, lds, (bf16*)dst, ldd, rows, cols);
  else if (src_dtype == SC_BF16 && dst_dtype == SC_F32)
    cast_kernel<bf16, float><<<blocks, 256, 0, st>>>((const bf16*)src, lds, (float*)dst, ldd, rows, cols);
  else if (src_dtype == SC_F32 && dst_dtype == SC_F32)
    cast_kernel<float, float><<<blocks, 256, 0, st>>>((const float*)src, lds, (float*)dst, ldd, rows, cols);
  else if (src_dtype == SC_BF16 && dst_dtype == SC_BF16)
    cast_kernel<bf16, bf16><<<blocks, 256, 0, st>>>((const bf16*)src, lds, (bf16*)dst, ldd, rows, cols);
  else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}

extern "C" int sc_mask_rows(const void* src, int64_t lds, int dtype, const uint8_t* mask, void* dst, int64_t ldd,
                            int64_t rows, int64_t cols, void* stream) {
  SC_CHECK_ARG(rows >= 0 && cols >= 0, SC_E_BADARG);
  if (rows * cols == 0) return 0;
  SC_CHECK_ARG(src && dst && mask, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)min((int64_t)148 * 16, cdiv(rows * cols, 256));
  if (dtype == SC_F32) mask_rows_kernel<float><<<blocks, 256, 0, st>>>((const float*)src, lds, mask, (float*)dst, ldd, rows, cols);
  else if (dtype == SC_BF16) mask_rows_kernel<bf16><<<blocks, 256, 0, st>>>((const bf16*)src, lds, mask, (bf16*)dst, ldd, rows, cols);
  else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}

extern "C" int sc_colsum(const void* X, int64_t ldx, int dtype, float* out, int64_t M, int64_t N,
                         int accumulate, void* stream) {
  SC_CHECK_ARG(M >= 0 && N >= 0, SC_E_BADARG);
  if (N == 0) return 0;
  SC_CHECK_ARG(out && (M == 0 || X), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) zero_kernel<<<(unsigned)cdiv(N, 256), 256, 0, st>>>(out, N);
  if (M > 0) {
    const int64_t col_blocks = cdiv(N, 128);
    int64_t row_blocks = cdiv(148 * 8, col_blocks);
    if (row_blocks > cdiv(M, 64)) row_blocks = cdiv(M, 64);
    const int64_t rpb = cdiv(M, row_blocks);
    dim3 grid((unsigned)col_blocks, (unsigned)cdiv(M, rpb));
    if (dtype == SC_F32) colsum_kernel<float><<<grid, 128, 0, st>>>((const float*)X, ldx, out, M, N, rpb);
    else if (dtype == SC_BF16) colsum_kernel<bf16><<<grid, 128, 0, st>>>((const bf16*)X, ldx, out, M, N, rpb);
    else return SC_E_DTYPE;
  }
  SC_LAUNCH_RET();
}

extern "C" int sc_layernorm_fwd(const void* X, int64_t ldx, const float* w, const float* b,
                                void* Y, int64_t ldy, float* mean, float* rstd,
                                int64_t M, int64_t H, int dtype, void* stream) {
  SC_CHECK_ARG(M >= 0 && H > 0 && H < (1 << 24), SC_E_BADARG);
  if (M == 0) return 0;
  SC_CHECK_ARG(X && w && b && Y && mean && rstd, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)cdiv(M, LN_WARPS);
  SC_CHECK_ARG(dtype == SC_F32 || dtype == SC_BF16, SC_E_DTYPE);
  const bool vec_ok = (H % 256 == 0) && H <= 1024 && (ldx % 8 == 0) && (ldy % 8 == 0) && aligned16(X) && aligned16(Y) &&
                      aligned16(w) && aligned16(b);
#define SC_LN_FWD_V(TT, NV) layernorm_fwd_vec_kernel<TT, NV><<<blocks, LN_WARPS * 32, 0, st>>>((const TT*)X, ldx, w, b, (TT*)Y, ldy, mean, rstd, M, (int)H)
#define SC_LN_FWD_T(TT) do { \
    if (!vec_ok) layernorm_fwd_kernel<TT><<<blocks, LN_WARPS * 32, 0, st>>>((const TT*)X, ldx, w, b, (TT*)Y, ldy, mean, rstd, M, (int)H); \
    else if (H == 256) SC_LN_FWD_V(TT, 1); else if (H == 512) SC_LN_FWD_V(TT, 2); else if (H == 768) SC_LN_FWD_V(TT, 3); else SC_LN_FWD_V(TT, 4); } while (0)
  if (dtype == SC_F32) SC_LN_FWD_T(float); else SC_LN_FWD_T(bf16);
#undef SC_LN_FWD_T
#undef SC_LN_FWD_V
  SC_LAUNCH_RET();
}

extern "C" int sc_layernorm_bwd(const void* dY, int64_t lddy, const void* X, int64_t ldx, const float* w,
                                const float* mean, const float* rstd, void* dX, int64_t lddx,
                                float* dw, float* db, float* dxsum, int64_t M, int64_t H, int dtype, void* stream) {
  SC_CHECK_ARG(M >= 0 && H > 0 && H <= 24 * 1024, SC_E_BADARG);
  if (M == 0) return 0;
  SC_CHECK_ARG(dY && X && w && mean && rstd && dX && dw && db, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t blocks = 148 * 4;
  if (blocks > cdiv(M, LN_WARPS)) blocks = cdiv(M, LN_WARPS);
  const int64_t rpb = cdiv(M, blocks);
  blocks = cdiv(M, rpb);
  const size_t smem = 3 * (size_t)H * sizeof(float);
  if (dtype == SC_F32) launch_ln_bwd<float>(dY, lddy, X, ldx, w, mean, rstd, dX, lddx, dw, db, dxsum, M, (int)H, blocks, rpb, smem, st);
  else if (dtype == SC_BF16) launch_ln_bwd<bf16>(dY, lddy, X, ldx, w, mean, rstd, dX, lddx, dw, db, dxsum, M, (int)H, blocks, rpb, smem, st);
  else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}

extern "C" int sc_split_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols,
                             void* stream) {
  SC_CHECK_ARG(rows >= 0 && cols >= 0, SC_E_BADARG);
  if (rows * cols == 0) return 0;
  SC_CHECK_ARG(src && dst && ldd >= 2 * cols, SC_E_BADARG);
  if (cols % 4 == 0 && lds % 4 == 0 && ldd % 4 == 0 && aligned16(src) && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) &&
      rows * (cols / 4) < ((int64_t)1 << 31)) {
    const unsigned blocks = (unsigned)min((int64_t)148 * 16, cdiv(rows * (cols / 4), 256));
    split_bf16_vec4_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, (bf16*)dst, ldd, (int)rows, (int)(cols / 4));
    SC_LAUNCH_RET();
  }
  const unsigned blocks = (unsigned)min((int64_t)148 * 16, cdiv(rows * cols, 256));
  split_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, (bf16*)dst, ldd, rows, cols);
  SC_LAUNCH_RET();
}

extern "C" int sc_split6_bf16(const float* src, int64_t lds, void* dst, int64_t ldd, int64_t rows, int64_t cols,
                              int pattern, int64_t block_stride, void* stream) {
  SC_CHECK_ARG(rows >= 0 && cols >= 0 && (pattern == 0 || pattern == 1) && block_stride > 0, SC_E_BADARG);
  if (rows * cols == 0) return 0;
  SC_CHECK_ARG(src && dst && ldd >= cols, SC_E_BADARG);
  const unsigned blocks = (unsigned)min((int64_t)148 * 16, cdiv(rows * cols, 256));
  split6_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, lds, (bf16*)dst, ldd, rows, cols, pattern, block_stride);
  SC_LAUNCH_RET();
}
