// K2, TMA-staged variant: the fused LucyRNN scan (forward and reverse-time backward) with
// gate rows landed in shared memory by the TMA engine instead of register prefetch.
//
// Why (measured on B200, profiles/r01_*): the register-prefetch kernels in sc_scan.cu can
// only keep ~5 MB of loads in flight chip-wide (landing buffers compete with everything else
// for registers), which caps them at 51 % (fwd) / 25 % (bwd) of HBM peak.  Here each CTA owns
// CB adjacent channels of one stream for the whole segment and a ring of NST shared-memory
// stages; one thread issues cp.async.bulk.tensor boxes of [TC timesteps x CB channels] per
// gate, completion is signalled on an mbarrier, and all threads only ever read shared
// memory.  ~100 KB per SM are in flight regardless of occupancy, rows are fetched as
// CB*e-byte bursts, and S and h never leave registers.
//
// Same math, same C-ABI and same checkpoints as sc_scan.cu; the
// dispatcher in sc_scan.cu picks this path when H and the strides satisfy TMA alignment.
#include "sc_common.cuh"
#include "sc_tma.cuh"
#include <stdlib.h>

namespace sc {

constexpr int TC = SC_SCAN_CKPT;     // timesteps per stage == checkpoint interval
constexpr int CB = 256;              // channels per CTA (bf16: 512-byte rows)
// channels per thread (VEC) is a template parameter: 1 doubles the resident warps per SM (16
// instead of 8), 2 halves the instruction count per channel

template <typename T> struct TmaType;
template <> struct TmaType<bf16>  { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16; };
template <> struct TmaType<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };

// [rows, cols] row-major (row stride ld elements) -> boxes of [box_rows x CB], no swizzle
template <typename T>
static bool make_scan_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(T)};
  cuuint32_t box[2] = {(cuuint32_t)CB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, TmaType<T>::v, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// shared-memory element access: thread owns channels [VEC*tid, VEC*tid+VEC) of a [TC][CB] box
__device__ __forceinline__ void lds2(const bf16* row, int tid, float (&f)[2]) {
  const uint32_t w = reinterpret_cast<const uint32_t*>(row)[tid];
  f[0] = __uint_as_float(w << 16);
  f[1] = __uint_as_float(w & 0xffff0000u);
}
__device__ __forceinline__ void lds2(const float* row, int tid, float (&f)[2]) {
  const float2 w = reinterpret_cast<const float2*>(row)[tid];
  f[0] = w.x; f[1] = w.y;
}
__device__ __forceinline__ void lds2(const bf16* row, int tid, float (&f)[1]) {
  f[0] = __uint_as_float((uint32_t)reinterpret_cast<const uint16_t*>(row)[tid] << 16);
}
__device__ __forceinline__ void lds2(const float* row, int tid, float (&f)[1]) { f[0] = row[tid]; }
template <typename T> __device__ __forceinline__ void stg_vec(T* p, const float (&f)[2]) { vstore<T, 2>(p, pack(f, (T*)nullptr)); }
template <typename T> __device__ __forceinline__ void stg_vec(T* p, const float (&f)[1]) { st_f(p, f[0]); }

// ------------------------------------------------------------------ forward ----------
template <typename T, int VEC, int NST, bool TRAIN, bool PRECISE>
__global__ void __launch_bounds__(CB / VEC)
lucy_scan_fwd_tma_kernel(const __grid_constant__ CUtensorMap mapG, const float* __restrict__ h0,
                         const float* __restrict__ s0, T* __restrict__ Hout, int64_t ldh,
                         float* __restrict__ hT, float* __restrict__ sT, float* __restrict__ Sckpt,
                         int Tn, int H, int cblocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int BOX = TC * CB * (int)sizeof(T);          // one gate box
  constexpr int STAGE = 5 * BOX;
  __shared__ __align__(8) uint64_t bars[NST];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / cblocks;
  const int c0 = (blockIdx.x % cblocks) * CB;
  const int ch = c0 + tid * VEC;
  const bool live = ch < H;                              // H is even; VEC channels live together
  const int nchunk = (Tn + TC - 1) / TC;
  const uint32_t sbase = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto issue = [&](int chunk) {
    const int st = chunk % NST;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_expect_tx(bar, STAGE);
    const int row = b * Tn + chunk * TC;
#pragma unroll
    for (int g = 0; g < 5; ++g) tma_load_2d(sbase + st * STAGE + g * BOX, &mapG, bar, g * H + c0, row);
  };
  if (tid == 0)
    for (int c = 0; c < NST - 1 && c < nchunk; ++c) issue(c);

  float S[VEC], h[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    h[i] = live ? h0[(int64_t)b * H + ch + i] : 0.f;
    S[i] = (TRAIN || !live) ? 0.f : s0[(int64_t)b * H + ch + i];
  }
  T* ho = Hout + (int64_t)b * Tn * ldh + ch;

  for (int c = 0; c < nchunk; ++c) {
    __syncthreads();                                     // everyone is done with chunk c-1's stage
    if (tid == 0 && c + NST - 1 < nchunk) issue(c + NST - 1);
    mbar_wait(smem_u32(&bars[c % NST]), (uint32_t)((c / NST) & 1));
    const T* st = reinterpret_cast<const T*>(smem + (c % NST) * STAGE);
    if (live && Sckpt != nullptr) {
      float* ck = Sckpt + ((int64_t)b * nchunk + c) * H + ch;
#pragma unroll
      for (int i = 0; i < VEC; ++i) ck[i] = S[i];
    }
    const int t0 = c * TC;
#pragma unroll
    for (int u = 0; u < TC; ++u) {
      if (t0 + u < Tn) {
        float z[VEC], k[VEC], v[VEC], p[VEC], q[VEC], out[VEC];
        lds2(st + (SC_GATE_Z * TC + u) * CB, tid, z);
        lds2(st + (SC_GATE_K * TC + u) * CB, tid, k);
        lds2(st + (SC_GATE_V * TC + u) * CB, tid, v);
        lds2(st + (SC_GATE_P * TC + u) * CB, tid, p);
        lds2(st + (SC_GATE_Q * TC + u) * CB, tid, q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float d = sigmoidf_<PRECISE>(q[i]);
          const float kv = k[i] * v[i];
          S[i] = fmaf(d, S[i], kv);
          const float sp = TRAIN ? fmaf(d, S[i], kv) : S[i];
          const float cc = tanhf_<PRECISE>(p[i] + sp);
          const float zh = sigmoidf_<PRECISE>(z[i]);
          h[i] = fmaf(zh, h[i] - cc, cc);
          out[i] = h[i];
        }
        if (live) stg_vec<T>(ho + (int64_t)(t0 + u) * ldh, out);
      }
    }
  }
  if (live) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      hT[(int64_t)b * H + ch + i] = h[i];
      if (!TRAIN && sT != nullptr) sT[(int64_t)b * H + ch + i] = S[i];
    }
  }
}

// ------------------------------------------------------------------ backward ---------
// Stage = 5 gate boxes + dHout box + Hout box shifted one row back (h_{t-1}).
template <typename T, int VEC, int NST, bool TRAIN, bool PRECISE>
__global__ void __launch_bounds__(CB / VEC)
lucy_scan_bwd_tma_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapH,
                         const __grid_constant__ CUtensorMap mapDH, const float* __restrict__ h0,
                         const float* __restrict__ Sckpt, T* __restrict__ dG, int64_t lddg,
                         float* __restrict__ dbias, int Tn, int H, int cblocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int BOX = TC * CB * (int)sizeof(T);
  constexpr int STAGE = 7 * BOX;
  __shared__ __align__(8) uint64_t bars[NST];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / cblocks;
  const int c0 = (blockIdx.x % cblocks) * CB;
  const int ch = c0 + tid * VEC;
  const bool live = ch < H;
  const int nchunk = (Tn + TC - 1) / TC;
  const uint32_t sbase = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // chunks are visited last-to-first; "it" counts visits
  auto issue = [&](int it) {
    const int chunk = nchunk - 1 - it;
    const int st = it % NST;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_expect_tx(bar, STAGE);
    const int row = b * Tn + chunk * TC;
#pragma unroll
    for (int g = 0; g < 5; ++g) tma_load_2d(sbase + st * STAGE + g * BOX, &mapG, bar, g * H + c0, row);
    tma_load_2d(sbase + st * STAGE + 5 * BOX, &mapDH, bar, c0, row);
    tma_load_2d(sbase + st * STAGE + 6 * BOX, &mapH, bar, c0, row - 1);   // h_{t-1}; row -1 is OOB -> zeros
  };
  if (tid == 0)
    for (int it = 0; it < NST - 1 && it < nchunk; ++it) issue(it);

  float gz[VEC], ds[VEC], acc[5][VEC], hfirst[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    gz[i] = 0.f; ds[i] = 0.f;
    hfirst[i] = live ? h0[(int64_t)b * H + ch + i] : 0.f;
#pragma unroll
    for (int g = 0; g < 5; ++g) acc[g][i] = 0.f;
  }
  T* dg = dG + (int64_t)b * Tn * lddg + ch;
  float Snext[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) Snext[i] = live ? Sckpt[((int64_t)b * nchunk + nchunk - 1) * H + ch + i] : 0.f;

  for (int it = 0; it < nchunk; ++it) {
    const int chunk = nchunk - 1 - it;
    __syncthreads();
    if (tid == 0 && it + NST - 1 < nchunk) issue(it + NST - 1);
    // checkpoint of THIS interval was fetched one iteration ago; fetch the next one now so its
    // HBM latency hides behind this interval's work (it was 24 % of the stall samples)
    float Sin[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      Sin[i] = Snext[i];
      Snext[i] = (live && chunk > 0) ? Sckpt[((int64_t)b * nchunk + chunk - 1) * H + ch + i] : 0.f;
    }
    mbar_wait(smem_u32(&bars[it % NST]), (uint32_t)((it / NST) & 1));
    const T* st = reinterpret_cast<const T*>(smem + (it % NST) * STAGE);
    const int t0 = chunk * TC;
    // pass 1: recompute S_t and d_t across the interval
    float Sl[TC][VEC], dl[TC][VEC];
    {
      float S[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) S[i] = Sin[i];
#pragma unroll
      for (int u = 0; u < TC; ++u) {
        float k[VEC], v[VEC], q[VEC];
        lds2(st + (SC_GATE_K * TC + u) * CB, tid, k);
        lds2(st + (SC_GATE_V * TC + u) * CB, tid, v);
        lds2(st + (SC_GATE_Q * TC + u) * CB, tid, q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float d = sigmoidf_<PRECISE>(q[i]);
          S[i] = fmaf(d, S[i], k[i] * v[i]);
          Sl[u][i] = S[i];
          dl[u][i] = d;
        }
      }
    }
    // pass 2: reverse time
#pragma unroll
    for (int u = TC - 1; u >= 0; --u) {
      const int t = t0 + u;
      if (t < Tn) {
        float z[VEC], k[VEC], v[VEC], p[VEC], go[VEC], hp[VEC];
        lds2(st + (SC_GATE_Z * TC + u) * CB, tid, z);
        lds2(st + (SC_GATE_K * TC + u) * CB, tid, k);
        lds2(st + (SC_GATE_V * TC + u) * CB, tid, v);
        lds2(st + (SC_GATE_P * TC + u) * CB, tid, p);
        lds2(st + (5 * TC + u) * CB, tid, go);
        lds2(st + (6 * TC + u) * CB, tid, hp);
        if (t == 0) {
#pragma unroll
          for (int i = 0; i < VEC; ++i) hp[i] = hfirst[i];
        }
        float dz[VEC], dk[VEC], dv[VEC], dp[VEC], dq[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float d = dl[u][i];
          const float kv = k[i] * v[i];
          const float St = Sl[u][i];
          const float Sp = (u > 0) ? Sl[u - 1][i] : Sin[i];
          const float sp = TRAIN ? fmaf(d, St, kv) : St;
          const float cc = tanhf_<PRECISE>(p[i] + sp);
          const float zh = sigmoidf_<PRECISE>(z[i]);
          const float gam = go[i] + gz[i];
          const float omz = 1.f - zh;
          const float da = gam * omz * fmaf(-cc, cc, 1.f);
          dz[i] = gam * (hp[i] - cc) * zh * omz;
          gz[i] = zh * gam;
          dp[i] = da;
          float sig, dkv, dd;
          if (TRAIN) {
            sig = fmaf(d, da, ds[i]);
            dkv = da + sig;
            dd = fmaf(St, da, Sp * sig);
          } else {
            sig = da + ds[i];
            dkv = sig;
            dd = Sp * sig;
          }
          ds[i] = d * sig;
          dk[i] = dkv * v[i];
          dv[i] = dkv * k[i];
          dq[i] = dd * d * (1.f - d);
          acc[SC_GATE_Z][i] += dz[i]; acc[SC_GATE_K][i] += dk[i]; acc[SC_GATE_V][i] += dv[i];
          acc[SC_GATE_P][i] += dp[i]; acc[SC_GATE_Q][i] += dq[i];
        }
        if (live) {
          T* row = dg + (int64_t)t * lddg;
          stg_vec<T>(row + (int64_t)SC_GATE_Z * H, dz);
          stg_vec<T>(row + (int64_t)SC_GATE_K * H, dk);
          stg_vec<T>(row + (int64_t)SC_GATE_V * H, dv);
          stg_vec<T>(row + (int64_t)SC_GATE_P * H, dp);
          stg_vec<T>(row + (int64_t)SC_GATE_Q * H, dq);
        }
      }
    }
  }
  if (live && dbias != nullptr) {
#pragma unroll
    for (int g = 0; g < 5; ++g)
#pragma unroll
      for (int i = 0; i < VEC; ++i) atomicAdd(dbias + (int64_t)g * H + ch + i, acc[g][i]);
  }
}

// ------------------------------------------------------------------ host -------------
template <typename T>
static bool tma_ok(const void* G, int64_t ldg, const void* a, int64_t lda, const void* b, int64_t ldb, int64_t H) {
  constexpr int per16 = 16 / (int)sizeof(T);
  if (H % per16 != 0) return false;
  if (ldg % per16 || lda % per16 || ldb % per16) return false;
  if (!aligned16(G) || !aligned16(a) || !aligned16(b)) return false;
  return get_encode() != nullptr;
}

// SC_SCAN_VEC=1|2 in the environment overrides the channels-per-thread choice (A/B measurements)
static int scan_vec_choice() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("SC_SCAN_VEC"); v = (e && e[0] == '1') ? 1 : (e && e[0] == '2') ? 2 : 0; }
  return v;
}

template <typename T, int VEC, bool PRECISE>
static int scan_fwd_tma(const void* G, int64_t ldg, const float* h0, const float* s0, void* Hout, int64_t ldh,
                        float* hT, float* sT, float* Sckpt, int64_t B, int64_t Tn, int64_t H, int train,
                        cudaStream_t st) {
  constexpr int NST = 4;
  constexpr int smem = NST * 5 * TC * CB * (int)sizeof(T);
  CUtensorMap mapG;
  if (!make_scan_map<T>(&mapG, G, B * Tn, 5 * H, ldg, TC)) return SC_E_UNSUP;
  const int cblocks = (int)cdiv(H, CB);
  const unsigned grid = (unsigned)(B * cblocks);
  auto kt = lucy_scan_fwd_tma_kernel<T, VEC, NST, true, PRECISE>;
  auto ks = lucy_scan_fwd_tma_kernel<T, VEC, NST, false, PRECISE>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  if (train) kt<<<grid, CB / VEC, smem, st>>>(mapG, h0, s0, (T*)Hout, ldh, hT, sT, Sckpt, (int)Tn, (int)H, cblocks);
  else       ks<<<grid, CB / VEC, smem, st>>>(mapG, h0, s0, (T*)Hout, ldh, hT, sT, Sckpt, (int)Tn, (int)H, cblocks);
  SC_LAUNCH_RET();
}

template <typename T, int VEC, bool PRECISE>
static int scan_bwd_tma(const void* G, int64_t ldg, const void* Hout, int64_t ldh, const float* h0,
                        const float* Sckpt, const void* dHout, int64_t lddh, void* dG, int64_t lddg,
                        float* dbias, int64_t B, int64_t Tn, int64_t H, int train, cudaStream_t st) {
  constexpr int NST = 3;
  constexpr int smem = NST * 7 * TC * CB * (int)sizeof(T);
  CUtensorMap mapG, mapH, mapDH;
  if (!make_scan_map<T>(&mapG, G, B * Tn, 5 * H, ldg, TC) || !make_scan_map<T>(&mapH, Hout, B * Tn, H, ldh, TC) ||
      !make_scan_map<T>(&mapDH, dHout, B * Tn, H, lddh, TC))
    return SC_E_UNSUP;
  const int cblocks = (int)cdiv(H, CB);
  const unsigned grid = (unsigned)(B * cblocks);
  auto kt = lucy_scan_bwd_tma_kernel<T, VEC, NST, true, PRECISE>;
  auto ks = lucy_scan_bwd_tma_kernel<T, VEC, NST, false, PRECISE>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(kt, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  if (train) kt<<<grid, CB / VEC, smem, st>>>(mapG, mapH, mapDH, h0, Sckpt, (T*)dG, lddg, dbias, (int)Tn, (int)H, cblocks);
  else       ks<<<grid, CB / VEC, smem, st>>>(mapG, mapH, mapDH, h0, Sckpt, (T*)dG, lddg, dbias, (int)Tn, (int)H, cblocks);
  SC_LAUNCH_RET();
}

// entry points used by the dispatcher in sc_scan.cu; return SC_E_UNSUP when TMA cannot serve
int scan_fwd_tma_dispatch(const void* G, int64_t ldg, const float* h0, const float* s0, void* Hout, int64_t ldh,
                          float* hT, float* sT, float* Sckpt, int64_t B, int64_t T, int64_t H, int dtype,
                          int train, cudaStream_t st) {
  if (T == 0 || B * T >= ((int64_t)1 << 31) - TC) return SC_E_UNSUP;
  if (dtype == SC_BF16) {
    if (!tma_ok<bf16>(G, ldg, Hout, ldh, Hout, ldh, H)) return SC_E_UNSUP;
    if (scan_vec_choice() == 1) return scan_fwd_tma<bf16, 1, false>(G, ldg, h0, s0, Hout, ldh, hT, sT, Sckpt, B, T, H, train, st);
    return scan_fwd_tma<bf16, 2, false>(G, ldg, h0, s0, Hout, ldh, hT, sT, Sckpt, B, T, H, train, st);
  }
  if (dtype == SC_F32) {
    if (!tma_ok<float>(G, ldg, Hout, ldh, Hout, ldh, H)) return SC_E_UNSUP;
    return scan_fwd_tma<float, 2, true>(G, ldg, h0, s0, Hout, ldh, hT, sT, Sckpt, B, T, H, train, st);
  }
  return SC_E_DTYPE;
}

int scan_bwd_tma_dispatch(const void* G, int64_t ldg, const void* Hout, int64_t ldh, const float* h0,
                          const float* Sckpt, const void* dHout, int64_t lddh, void* dG, int64_t lddg,
                          float* dbias, int64_t B, int64_t T, int64_t H, int dtype, int train, cudaStream_t st) {
  if (T == 0 || B * T >= ((int64_t)1 << 31) - TC) return SC_E_UNSUP;
  if (dtype == SC_BF16) {
    if (!tma_ok<bf16>(G, ldg, Hout, ldh, dHout, lddh, H) || (lddg % 2) || ((uintptr_t)dG & 3)) return SC_E_UNSUP;
    if (scan_vec_choice() == 1) return scan_bwd_tma<bf16, 1, false>(G, ldg, Hout, ldh, h0, Sckpt, dHout, lddh, dG, lddg, dbias, B, T, H, train, st);
    return scan_bwd_tma<bf16, 2, false>(G, ldg, Hout, ldh, h0, Sckpt, dHout, lddh, dG, lddg, dbias, B, T, H, train, st);
  }
  if (dtype == SC_F32) {
    if (!tma_ok<float>(G, ldg, Hout, ldh, dHout, lddh, H) || (lddg % 2) || ((uintptr_t)dG & 7)) return SC_E_UNSUP;
    return scan_bwd_tma<float, 2, true>(G, ldg, Hout, ldh, h0, Sckpt, dHout, lddh, dG, lddg, dbias, B, T, H, train, st);
  }
  return SC_E_DTYPE;
}

}  // namespace sc
