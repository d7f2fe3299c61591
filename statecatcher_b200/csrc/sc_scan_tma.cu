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
#include <initializer_list>

namespace sc {

constexpr int TC = SC_SCAN_CKPT;     // timesteps per stage == checkpoint interval
// ring depths of the 16-bit kernels (stages of TC rows).  r02, deeper rings everywhere (fwd 5, bwd 4, split scans 6-8) in
// alternating same-box steps: forward 2.50 vs 2.43-2.50 ms, backward 5.20-5.24 vs 5.26-5.36, split scans unchanged — the
// ~100 KB per SM these depths keep in flight already cover the latency
#ifndef SC_NST_FWD
#define SC_NST_FWD 4
#endif
#ifndef SC_NST_BWD
#define SC_NST_BWD 3
#endif
#ifndef SC_SCAN_BWD_ROWS
#define SC_SCAN_BWD_ROWS 8                   // rows per stage of the fused backward (16-bit rows); 16 -> two stages of 56 KB: measured, no gain
                                             // (5.10-5.11 against 5.02-5.03 ms per step, alternating same-box runs; the split kernels of the
                                             // LayerNorm path did gain from 16/32-row stages - they carry 2-3 boxes per stage, this one seven)
#endif
#ifndef SC_SSCAN_ROWS
#define SC_SSCAN_ROWS 16                     // rows per stage of the S-scan kernels (16-bit rows)
#endif
#ifndef SC_HSCAN_FWD_ROWS
#define SC_HSCAN_FWD_ROWS 32                 // rows per stage of the h-scan kernels (16-bit rows)
#endif
#ifndef SC_HSCAN_BWD_ROWS
#define SC_HSCAN_BWD_ROWS 16
#endif
#ifndef SC_SCAN_CB
#define SC_SCAN_CB 256
#endif
constexpr int CB = SC_SCAN_CB;       // channels per CTA (bf16: 512-byte rows).  r02, -DSC_SCAN_CB=128 (512 CTAs of 2 warps, 5 per SM) against 256
                                     // in alternating same-box steps: fwd 2.53 vs 2.46-2.74 ms, bwd 5.30-5.33 vs 5.14-5.18 ms per step: 256 stays
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

// the five gate blocks of G [rows, 5H] as ONE tensor [rows][5][H] -> a single box of [box_rows][5 gates][CB] per interval
// instead of five (r02: every cp.async.bulk issue costs its thread ~100 ns — measured on the CTC recursion kernel — and
// the issuing thread here is one of the compute warps, which the rest of the block then waits for at the barrier).
// Same-box alternating steps: forward 2.383 -> 2.340 ms per step; the backward (7 -> 3 issues) went 5.015 -> 5.06 and keeps
// its five 2-D boxes.
template <typename T>
static bool make_scan_map_gates(CUtensorMap* m, const void* ptr, int64_t rows, int64_t H, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[3] = {(cuuint64_t)H, 5, (cuuint64_t)rows};
  cuuint64_t strides[2] = {(cuuint64_t)H * sizeof(T), (cuuint64_t)ld * sizeof(T)};
  cuuint32_t box[3] = {(cuuint32_t)CB, 5, (cuuint32_t)box_rows};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, TmaType<T>::v, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
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

// Tried and rejected (r01): packed fp32 pairs (FADD2/FMUL2/FFMA2 over a thread's two channels; same
// IEEE operations, bit-identical output, 14 % / 21 % fewer SASS instructions in fwd / bwd).  Measured
// inside the power-capped step, same box, alternating runs: scan fwd 2.83 -> 2.97 ms, bwd 5.01 -> 5.12 ms
// per step.  With ~7 warps per SM the scans are bound by the dependent chain of a warp's timestep, and
// the packed instructions lengthen that chain more than the saved issue slots shorten it.
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
    tma_load_3d(sbase + st * STAGE, &mapG, bar, c0, 0, row);      // [TC rows][5 gates][CB] in one box
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
    T* hp = ho + (int64_t)t0 * ldh;                      // running output pointer (one 64-bit add per step)
    auto step = [&](int u) {
      float z[VEC], k[VEC], v[VEC], p[VEC], q[VEC], out[VEC];
      lds2(st + (u * 5 + SC_GATE_Z) * CB, tid, z);
      lds2(st + (u * 5 + SC_GATE_K) * CB, tid, k);
      lds2(st + (u * 5 + SC_GATE_V) * CB, tid, v);
      lds2(st + (u * 5 + SC_GATE_P) * CB, tid, p);
      lds2(st + (u * 5 + SC_GATE_Q) * CB, tid, q);
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
      if (live) stg_vec<T>(hp, out);
      hp += ldh;
    };
    if (t0 + TC <= Tn) {                                 // full interval: no per-step bounds test
      {
        // The same operations per element in the same order (bit-identical), arranged by what depends on what: only
        // S_t = d S_{t-1} + kv and h_t = zh (h_{t-1} - c) + c are carried from step to step; the gates' activations
        // and the candidate c_t = tanh(p + S') are not.  Written step by step (`step` above) a warp — there are only
        // ~7 per SM, one or two per scheduler — serialises ~46 instructions with their LDS / MUFU latencies: measured
        // 234 cycles per step, fast enough for HBM only at the full clock (alone 0.99 of the copy peak, inside the
        // power-capped step 0.75).  Here all 8 steps' loads and activations go first (40 LDS and 32 MUFU in flight
        // together), then the S chain, the 16 tanh, the h chain and the stores: 47 registers instead of 32, inside
        // the step 2.87 -> 2.35 ms per 6 layers = 0.75 -> 0.92 of the copy peak (r02, alternating same-box runs).
        // (left to itself ptxas sinks every activation back to its use to save registers and arrives at the same
        // step-by-step schedule; the warp-level barrier below is a scheduling fence it does not move code across)
        float d[TC][VEC], kv[TC][VEC], zh[TC][VEC], pa[TC][VEC];
#pragma unroll
        for (int u = 0; u < TC; ++u) {
          float z[VEC], k[VEC], v[VEC], q[VEC];
          lds2(st + (u * 5 + SC_GATE_Z) * CB, tid, z);
          lds2(st + (u * 5 + SC_GATE_K) * CB, tid, k);
          lds2(st + (u * 5 + SC_GATE_V) * CB, tid, v);
          lds2(st + (u * 5 + SC_GATE_P) * CB, tid, pa[u]);
          lds2(st + (u * 5 + SC_GATE_Q) * CB, tid, q);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            d[u][i] = sigmoidf_<PRECISE>(q[i]);
            kv[u][i] = k[i] * v[i];
            zh[u][i] = sigmoidf_<PRECISE>(z[i]);
          }
        }
        __syncwarp();                                    // the scheduling fence
#pragma unroll
        for (int u = 0; u < TC; ++u)
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            S[i] = fmaf(d[u][i], S[i], kv[u][i]);
            pa[u][i] += TRAIN ? fmaf(d[u][i], S[i], kv[u][i]) : S[i];
          }
#pragma unroll
        for (int u = 0; u < TC; ++u)
#pragma unroll
          for (int i = 0; i < VEC; ++i) pa[u][i] = tanhf_<PRECISE>(pa[u][i]);
#pragma unroll
        for (int u = 0; u < TC; ++u) {
          float out[VEC];
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            h[i] = fmaf(zh[u][i], h[i] - pa[u][i], pa[u][i]);
            out[i] = h[i];
          }
          if (live) stg_vec<T>(hp, out);
          hp += ldh;
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < TC; ++u)
        if (t0 + u < Tn) step(u);
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
// (r02, measured and dropped: a warp-specialised form — a fifth warp issuing the copies, the checkpoint row arriving with
// the stage, per-warp full/empty mbarriers instead of the block barrier; ncu had 8 % of this kernel's stall samples on
// the block barrier and 11 % on the first use of the prefetched checkpoint — bit-identical, 5.37-5.42 ms per step against
// 5.03-5.09 in alternating same-box runs.)
template <typename T, int VEC, int NST, bool TRAIN, bool PRECISE, int ROWS>
__global__ void __launch_bounds__(CB / VEC)
lucy_scan_bwd_tma_kernel(const __grid_constant__ CUtensorMap mapG, const __grid_constant__ CUtensorMap mapH,
                         const __grid_constant__ CUtensorMap mapDH, const float* __restrict__ h0,
                         const float* __restrict__ Sckpt, T* __restrict__ dG, int64_t lddg,
                         float* __restrict__ dbias, int Tn, int H, int cblocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int BOX = ROWS * CB * (int)sizeof(T);         // stages of ROWS rows, walked last 8-step interval first
  constexpr int STAGE = 7 * BOX;
  __shared__ __align__(8) uint64_t bars[NST];
  const int tid = threadIdx.x;
  const int b = blockIdx.x / cblocks;
  const int c0 = (blockIdx.x % cblocks) * CB;
  const int ch = c0 + tid * VEC;
  const bool live = ch < H;
  const int nchunk = (Tn + TC - 1) / TC;                 // 8-step intervals (the checkpoints' index)
  const int nstage = (Tn + ROWS - 1) / ROWS;
  const uint32_t sbase = smem_u32(smem);

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // chunks are visited last-to-first; "it" counts visits
  auto issue = [&](int it) {
    const int chunk = nstage - 1 - it;
    const int st = it % NST;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_expect_tx(bar, STAGE);
    const int row = b * Tn + chunk * ROWS;
#pragma unroll
    for (int g = 0; g < 5; ++g) tma_load_2d(sbase + st * STAGE + g * BOX, &mapG, bar, g * H + c0, row);   // (one 3-D box of the five gates, as in the forward kernel: 5.06 vs 5.01 ms per step here)
    tma_load_2d(sbase + st * STAGE + 5 * BOX, &mapDH, bar, c0, row);
    tma_load_2d(sbase + st * STAGE + 6 * BOX, &mapH, bar, c0, row - 1);   // h_{t-1}; row -1 is OOB -> zeros
  };
  if (tid == 0)
    for (int it = 0; it < NST - 1 && it < nstage; ++it) issue(it);

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

  for (int it = 0; it < nstage; ++it) {
    const int stage_idx = nstage - 1 - it;
    __syncthreads();
    if (tid == 0 && it + NST - 1 < nstage) issue(it + NST - 1);
    mbar_wait(smem_u32(&bars[it % NST]), (uint32_t)((it / NST) & 1));
#pragma unroll 1
   for (int sub = ROWS / TC - 1; sub >= 0; --sub) {
    const int chunk = stage_idx * (ROWS / TC) + sub;
    if (chunk >= nchunk) continue;
    // checkpoint of THIS interval was fetched one interval ago; fetch the next one now so its
    // HBM latency hides behind this interval's work (it was 24 % of the stall samples)
    float Sin[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      Sin[i] = Snext[i];
      Snext[i] = (live && chunk > 0) ? Sckpt[((int64_t)b * nchunk + chunk - 1) * H + ch + i] : 0.f;
    }
    const T* st = reinterpret_cast<const T*>(smem + (it % NST) * STAGE) + (size_t)sub * TC * CB;   // row u of box g: st + (g * ROWS + u) * CB
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
        lds2(st + (SC_GATE_K * ROWS + u) * CB, tid, k);
        lds2(st + (SC_GATE_V * ROWS + u) * CB, tid, v);
        lds2(st + (SC_GATE_Q * ROWS + u) * CB, tid, q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float d = sigmoidf_<PRECISE>(q[i]);
          S[i] = fmaf(d, S[i], k[i] * v[i]);
          Sl[u][i] = S[i];
          dl[u][i] = d;
        }
      }
    }
    // pass 2: reverse time.  Five running output pointers (one per gate block) step back by one
    // row per timestep instead of rebuilding 64-bit addresses for every store.
    T* rz = dg + (int64_t)(t0 + TC - 1) * lddg;
    T* rk = rz + (int64_t)SC_GATE_K * H;
    T* rv = rz + (int64_t)SC_GATE_V * H;
    T* rp = rz + (int64_t)SC_GATE_P * H;
    T* rq = rz + (int64_t)SC_GATE_Q * H;
    auto step = [&](int u) {
      float z[VEC], k[VEC], v[VEC], p[VEC], go[VEC], hp[VEC];
      lds2(st + (SC_GATE_Z * ROWS + u) * CB, tid, z);
      lds2(st + (SC_GATE_K * ROWS + u) * CB, tid, k);
      lds2(st + (SC_GATE_V * ROWS + u) * CB, tid, v);
      lds2(st + (SC_GATE_P * ROWS + u) * CB, tid, p);
      lds2(st + (5 * ROWS + u) * CB, tid, go);
      lds2(st + (6 * ROWS + u) * CB, tid, hp);
      if (t0 + u == 0) {
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
        stg_vec<T>(rz, dz); stg_vec<T>(rk, dk); stg_vec<T>(rv, dv); stg_vec<T>(rp, dp); stg_vec<T>(rq, dq);
      }
    };
    // (r02: arranging pass 2 like the forward kernel — every step's activations and products first, behind a scheduling
    // fence, then the carried chain — was measured inside the step at 5.17 ms against 4.97 ms as it stands: 168 registers
    // instead of 80, and this kernel already sits at 0.87 of the copy peak.  Not kept.)
    if (t0 + TC <= Tn) {
#pragma unroll
      for (int u = TC - 1; u >= 0; --u) {
        step(u);
        rz -= lddg; rk -= lddg; rv -= lddg; rp -= lddg; rq -= lddg;
      }
    } else {
#pragma unroll
      for (int u = TC - 1; u >= 0; --u) {
        if (t0 + u < Tn) step(u);
        rz -= lddg; rk -= lddg; rv -= lddg; rp -= lddg; rq -= lddg;
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

// ------------------------------------------------------------------ split scans ------
// TMA-staged versions of the general-path kernels of sc_scan.cu (layer_norm=True /
// fused_ops=False): same thread/CTA mapping and stage ring as the fused kernels above.
constexpr int SV = 2;                 // channels per thread in the split kernels
constexpr int SPLIT_THREADS = CB / SV;

struct ScanBlock {
  int tid, b, c0, ch, nchunk;
  bool live;
  uint32_t sbase;
};
__device__ __forceinline__ ScanBlock scan_block(int Tn, int H, int cblocks, const uint8_t* smem) {
  ScanBlock k;
  k.tid = threadIdx.x;
  k.b = blockIdx.x / cblocks;
  k.c0 = (blockIdx.x % cblocks) * CB;
  k.ch = k.c0 + k.tid * SV;
  k.live = k.ch < H;
  k.nchunk = (Tn + TC - 1) / TC;
  k.sbase = smem_u32(smem);
  return k;
}
template <int NST>
__device__ __forceinline__ void scan_bars_init(uint64_t* bars, int tid) {
  if (tid == 0) {
    for (int s = 0; s < NST; ++s) mbar_init(smem_u32(&bars[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
}

// S scan + second application.  A = addend + s'.  Learned decay only (prefix_sum keeps the
// simple kernel).  Stage = k, v, q, addend boxes.
template <typename T, int NST, bool TRAIN, bool PRECISE, int ROWS>
__global__ void __launch_bounds__(SPLIT_THREADS)
sscan_fwd_tma_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV,
                     const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapAdd,
                     const float* __restrict__ s0, T* __restrict__ A, int64_t lda, float* __restrict__ S_all,
                     float* __restrict__ sT, int Tn, int H, int cblocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int BOX = ROWS * CB * (int)sizeof(T);        // stages of ROWS rows, walked as 8-step groups (see the h scan)
  constexpr int STAGE = 4 * BOX;
  __shared__ __align__(8) uint64_t bars[NST];
  const ScanBlock k = scan_block(Tn, H, cblocks, smem);   // k.nchunk: 8-step intervals (the checkpoints' index)
  const int nstage = (Tn + ROWS - 1) / ROWS;
  scan_bars_init<NST>(bars, k.tid);
  auto issue = [&](int chunk) {
    const int st = chunk % NST;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_expect_tx(bar, STAGE);
    const int row = k.b * Tn + chunk * ROWS;
    const uint32_t d = k.sbase + st * STAGE;
    tma_load_2d(d, &mapK, bar, k.c0, row);
    tma_load_2d(d + BOX, &mapV, bar, k.c0, row);
    tma_load_2d(d + 2 * BOX, &mapQ, bar, k.c0, row);
    tma_load_2d(d + 3 * BOX, &mapAdd, bar, k.c0, row);
  };
  if (k.tid == 0)
    for (int c = 0; c < NST - 1 && c < nstage; ++c) issue(c);
  float S[SV];
#pragma unroll
  for (int i = 0; i < SV; ++i) S[i] = (TRAIN || !k.live) ? 0.f : s0[(int64_t)k.b * H + k.ch + i];
  T* ao = A + (int64_t)k.b * Tn * lda + k.ch;
  for (int sc = 0; sc < nstage; ++sc) {
    __syncthreads();
    if (k.tid == 0 && sc + NST - 1 < nstage) issue(sc + NST - 1);
    mbar_wait(smem_u32(&bars[sc % NST]), (uint32_t)((sc / NST) & 1));
#pragma unroll 1
   for (int sub = 0; sub < ROWS / TC; ++sub) {
    const int c = sc * (ROWS / TC) + sub;                 // 8-step interval
    const int t0 = c * TC;
    if (t0 >= Tn) break;
    const T* st = reinterpret_cast<const T*>(smem + (sc % NST) * STAGE) + (size_t)sub * TC * CB;   // row u of box g: st + (g * ROWS + u) * CB
    // S entering the interval: all the backward needs (it recomputes inside an interval, as the fused kernel does);
    // r01 saved every S_t as an fp32 row: 4 bytes per element written here and 4.5 read back
    if (k.live) *reinterpret_cast<float2*>(S_all + ((int64_t)k.b * k.nchunk + c) * H + k.ch) = make_float2(S[0], S[1]);
    if (t0 + TC <= Tn) {
      // full interval: loads, decay activations and k*v of all steps first (scheduling fence), then the S chain
      float d[TC][SV], kv[TC][SV], ad[TC][SV];
#pragma unroll
      for (int u = 0; u < TC; ++u) {
        float kk[SV], vv[SV], qq[SV];
        lds2(st + (0 * ROWS + u) * CB, k.tid, kk);
        lds2(st + (1 * ROWS + u) * CB, k.tid, vv);
        lds2(st + (2 * ROWS + u) * CB, k.tid, qq);
        lds2(st + (3 * ROWS + u) * CB, k.tid, ad[u]);
#pragma unroll
        for (int i = 0; i < SV; ++i) { d[u][i] = sigmoidf_<PRECISE>(qq[i]); kv[u][i] = kk[i] * vv[i]; }
      }
      __syncwarp();
#pragma unroll
      for (int u = 0; u < TC; ++u) {
        float out[SV];
#pragma unroll
        for (int i = 0; i < SV; ++i) {
          S[i] = fmaf(d[u][i], S[i], kv[u][i]);
          out[i] = ad[u][i] + (TRAIN ? fmaf(d[u][i], S[i], kv[u][i]) : S[i]);
        }
        if (k.live) stg_vec<T>(ao + (int64_t)(t0 + u) * lda, out);
      }
      continue;
    }
#pragma unroll
    for (int u = 0; u < TC; ++u) {
      if (t0 + u < Tn) {
        float kk[SV], vv[SV], qq[SV], ad[SV], out[SV];
        lds2(st + (0 * ROWS + u) * CB, k.tid, kk);
        lds2(st + (1 * ROWS + u) * CB, k.tid, vv);
        lds2(st + (2 * ROWS + u) * CB, k.tid, qq);
        lds2(st + (3 * ROWS + u) * CB, k.tid, ad);
#pragma unroll
        for (int i = 0; i < SV; ++i) {
          const float d = sigmoidf_<PRECISE>(qq[i]);
          const float kv = kk[i] * vv[i];
          S[i] = fmaf(d, S[i], kv);
          out[i] = ad[i] + (TRAIN ? fmaf(d, S[i], kv) : S[i]);
        }
        if (k.live) stg_vec<T>(ao + (int64_t)(t0 + u) * lda, out);
      }
    }
   }
  }
  if (k.live && !TRAIN && sT != nullptr) {
#pragma unroll
    for (int i = 0; i < SV; ++i) sT[(int64_t)k.b * H + k.ch + i] = S[i];
  }
}

// Reverse-time adjoint of the S scan.  Stage = k, v, q, dA boxes; S_t of an interval is recomputed forwards from the
// interval's checkpoint (fetched one interval ahead), then the interval is walked backwards.
template <typename T, int NST, bool TRAIN, bool PRECISE, int ROWS>
__global__ void __launch_bounds__(SPLIT_THREADS)
sscan_bwd_tma_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV,
                     const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapDA,
                     const float* __restrict__ Sck,
                     T* __restrict__ dk, T* __restrict__ dv, T* __restrict__ dq, int64_t lddg,
                     float* __restrict__ dsum, int Tn, int H, int cblocks) {
  // dsum (may be null): [3][H] column sums of dk, dv, dq over all rows, added atomically — the bias gradients of the
  // gate projection, which otherwise cost a pass over the whole gradient tensor (sc_colsum)
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int BOX = ROWS * CB * (int)sizeof(T);        // stages of ROWS rows, walked last 8-step group first
  constexpr int STAGE = 4 * BOX;
  __shared__ __align__(8) uint64_t bars[NST];
  const ScanBlock k = scan_block(Tn, H, cblocks, smem);   // k.nchunk: 8-step intervals (the checkpoints' index)
  const int nstage = (Tn + ROWS - 1) / ROWS;
  scan_bars_init<NST>(bars, k.tid);
  auto issue = [&](int it) {
    const int chunk = nstage - 1 - it;
    const int st = it % NST;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_expect_tx(bar, STAGE);
    const int row = k.b * Tn + chunk * ROWS;
    const uint32_t d = k.sbase + st * STAGE;
    tma_load_2d(d, &mapK, bar, k.c0, row);
    tma_load_2d(d + BOX, &mapV, bar, k.c0, row);
    tma_load_2d(d + 2 * BOX, &mapQ, bar, k.c0, row);
    tma_load_2d(d + 3 * BOX, &mapDA, bar, k.c0, row);
  };
  if (k.tid == 0)
    for (int it = 0; it < NST - 1 && it < nstage; ++it) issue(it);
  float ds[SV], acc[3][SV], Snext[SV];
#pragma unroll
  for (int i = 0; i < SV; ++i) {
    ds[i] = 0.f;
    acc[0][i] = 0.f; acc[1][i] = 0.f; acc[2][i] = 0.f;
    Snext[i] = k.live ? Sck[((int64_t)k.b * k.nchunk + k.nchunk - 1) * H + k.ch + i] : 0.f;
  }
  const int64_t obase = (int64_t)k.b * Tn * lddg + k.ch;
  for (int it = 0; it < nstage; ++it) {
    const int stage_idx = nstage - 1 - it;
    __syncthreads();
    if (k.tid == 0 && it + NST - 1 < nstage) issue(it + NST - 1);
    mbar_wait(smem_u32(&bars[it % NST]), (uint32_t)((it / NST) & 1));
#pragma unroll 1
   for (int sub = ROWS / TC - 1; sub >= 0; --sub) {
    const int chunk = stage_idx * (ROWS / TC) + sub;      // 8-step interval
    if (chunk >= k.nchunk) continue;
    float Sin[SV];
#pragma unroll
    for (int i = 0; i < SV; ++i) {
      Sin[i] = Snext[i];
      Snext[i] = (k.live && chunk > 0) ? Sck[((int64_t)k.b * k.nchunk + chunk - 1) * H + k.ch + i] : 0.f;
    }
    const T* st = reinterpret_cast<const T*>(smem + (it % NST) * STAGE) + (size_t)sub * TC * CB;   // row u of box g: st + (g * ROWS + u) * CB
    const int t0 = chunk * TC;
    // pass 1: every load of the interval, d_t, and S_t forwards from the checkpoint (rows past the segment's end
    // belong to the next stream or are zero-filled: computed, never stored)
    float Sl[TC][SV], dl[TC][SV], kk[TC][SV], vv[TC][SV], da[TC][SV];
    {
      float S[SV];
#pragma unroll
      for (int i = 0; i < SV; ++i) S[i] = Sin[i];
#pragma unroll
      for (int u = 0; u < TC; ++u) {
        float qq[SV];
        lds2(st + (0 * ROWS + u) * CB, k.tid, kk[u]);
        lds2(st + (1 * ROWS + u) * CB, k.tid, vv[u]);
        lds2(st + (2 * ROWS + u) * CB, k.tid, qq);
        lds2(st + (3 * ROWS + u) * CB, k.tid, da[u]);
#pragma unroll
        for (int i = 0; i < SV; ++i) dl[u][i] = sigmoidf_<PRECISE>(qq[i]);
      }
      __syncwarp();                                    // scheduling fence: loads and activations stay in front of the chains
#pragma unroll
      for (int u = 0; u < TC; ++u)
#pragma unroll
        for (int i = 0; i < SV; ++i) {
          S[i] = fmaf(dl[u][i], S[i], kk[u][i] * vv[u][i]);
          Sl[u][i] = S[i];
        }
    }
    // pass 2: reverse time
#pragma unroll
    for (int u = TC - 1; u >= 0; --u) {
      const int t = t0 + u;
      if (t < Tn) {
        float ok[SV], ov[SV], oq[SV];
#pragma unroll
        for (int i = 0; i < SV; ++i) {
          const float d = dl[u][i];
          const float St = Sl[u][i];
          const float Sp = (u > 0) ? Sl[u - 1][i] : Sin[i];
          float sig, dkv, dd;
          if (TRAIN) {
            sig = fmaf(d, da[u][i], ds[i]);
            dkv = da[u][i] + sig;
            dd = fmaf(St, da[u][i], Sp * sig);
          } else {
            sig = da[u][i] + ds[i];
            dkv = sig;
            dd = Sp * sig;
          }
          ds[i] = d * sig;
          ok[i] = dkv * vv[u][i];
          ov[i] = dkv * kk[u][i];
          oq[i] = dd * d * (1.f - d);
          acc[0][i] += ok[i]; acc[1][i] += ov[i]; acc[2][i] += oq[i];
        }
        if (k.live) {
          const int64_t o = obase + (int64_t)t * lddg;
          stg_vec<T>(dk + o, ok);
          stg_vec<T>(dv + o, ov);
          stg_vec<T>(dq + o, oq);
        }
      }
    }
   }
  }
  if (k.live && dsum != nullptr) {
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int i = 0; i < SV; ++i) atomicAdd(dsum + (int64_t)g * H + k.ch + i, acc[g][i]);
  }
}

// h scan.  Stage = An, Zn boxes of TCH rows.  A stage of 8 rows made this kernel pay a block barrier, an mbarrier
// wait and two copy issues per ~150 instructions of work (0.36 ms per [192000 x 1024] call = 3.3 TB/s of 1.18 GB, and
// a deeper ring changed nothing): stages of TCH rows (32 for 16-bit rows) are walked as 8-step groups.
template <typename T, int NST, bool PRECISE, int TCH>
__global__ void __launch_bounds__(SPLIT_THREADS)
hscan_fwd_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapZ,
                     const float* __restrict__ h0, T* __restrict__ Hout, int64_t ldh, float* __restrict__ hT,
                     int Tn, int H, int cblocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int BOX = TCH * CB * (int)sizeof(T);
  constexpr int STAGE = 2 * BOX;
  __shared__ __align__(8) uint64_t bars[NST];
  ScanBlock k = scan_block(Tn, H, cblocks, smem);
  k.nchunk = (Tn + TCH - 1) / TCH;
  scan_bars_init<NST>(bars, k.tid);
  auto issue = [&](int chunk) {
    const int st = chunk % NST;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_expect_tx(bar, STAGE);
    const int row = k.b * Tn + chunk * TCH;
    tma_load_2d(k.sbase + st * STAGE, &mapA, bar, k.c0, row);
    tma_load_2d(k.sbase + st * STAGE + BOX, &mapZ, bar, k.c0, row);
  };
  if (k.tid == 0)
    for (int c = 0; c < NST - 1 && c < k.nchunk; ++c) issue(c);
  float h[SV];
#pragma unroll
  for (int i = 0; i < SV; ++i) h[i] = k.live ? h0[(int64_t)k.b * H + k.ch + i] : 0.f;
  T* ho = Hout + (int64_t)k.b * Tn * ldh + k.ch;
  for (int c = 0; c < k.nchunk; ++c) {
    __syncthreads();
    if (k.tid == 0 && c + NST - 1 < k.nchunk) issue(c + NST - 1);
    mbar_wait(smem_u32(&bars[c % NST]), (uint32_t)((c / NST) & 1));
    const T* st = reinterpret_cast<const T*>(smem + (c % NST) * STAGE);
#pragma unroll 1
    for (int sub = 0; sub < TCH / TC; ++sub) {
      const int t0 = c * TCH + sub * TC;
      if (t0 >= Tn) break;
      const T* sa = st + (size_t)sub * TC * CB;          // An rows of this group
      const T* sz = st + (size_t)(TCH + sub * TC) * CB;  // Zn rows
      if (t0 + TC <= Tn) {
        // full group: every step's loads and activations first, behind a scheduling fence, then the carried chain
        // (same operations in the same order per element; see lucy_scan_fwd_tma_kernel)
        float cc[TC][SV], zh[TC][SV];
#pragma unroll
        for (int u = 0; u < TC; ++u) {
          float an[SV], zn[SV];
          lds2(sa + u * CB, k.tid, an);
          lds2(sz + u * CB, k.tid, zn);
#pragma unroll
          for (int i = 0; i < SV; ++i) { cc[u][i] = tanhf_<PRECISE>(an[i]); zh[u][i] = sigmoidf_<PRECISE>(zn[i]); }
        }
        __syncwarp();
#pragma unroll
        for (int u = 0; u < TC; ++u) {
          float out[SV];
#pragma unroll
          for (int i = 0; i < SV; ++i) { h[i] = fmaf(zh[u][i], h[i] - cc[u][i], cc[u][i]); out[i] = h[i]; }
          if (k.live) stg_vec<T>(ho + (int64_t)(t0 + u) * ldh, out);
        }
      } else {
#pragma unroll
        for (int u = 0; u < TC; ++u) {
          if (t0 + u < Tn) {
            float an[SV], zn[SV], out[SV];
            lds2(sa + u * CB, k.tid, an);
            lds2(sz + u * CB, k.tid, zn);
#pragma unroll
            for (int i = 0; i < SV; ++i) {
              const float cc = tanhf_<PRECISE>(an[i]);
              const float zh = sigmoidf_<PRECISE>(zn[i]);
              h[i] = fmaf(zh, h[i] - cc, cc);
              out[i] = h[i];
            }
            if (k.live) stg_vec<T>(ho + (int64_t)(t0 + u) * ldh, out);
          }
        }
      }
    }
  }
  if (k.live) {
#pragma unroll
    for (int i = 0; i < SV; ++i) hT[(int64_t)k.b * H + k.ch + i] = h[i];
  }
}

// Reverse-time adjoint of the h scan.  Stage = An, Zn, dHout boxes + Hout shifted one row back, TCH rows each, walked
// last 8-step group first.
template <typename T, int NST, bool PRECISE, int TCH>
__global__ void __launch_bounds__(SPLIT_THREADS)
hscan_bwd_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapZ,
                     const __grid_constant__ CUtensorMap mapDH, const __grid_constant__ CUtensorMap mapH,
                     const float* __restrict__ h0, T* __restrict__ dAn, int64_t lddan, T* __restrict__ dZn,
                     int64_t lddzn, int Tn, int H, int cblocks) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int BOX = TCH * CB * (int)sizeof(T);
  constexpr int STAGE = 4 * BOX;
  __shared__ __align__(8) uint64_t bars[NST];
  ScanBlock k = scan_block(Tn, H, cblocks, smem);
  k.nchunk = (Tn + TCH - 1) / TCH;
  scan_bars_init<NST>(bars, k.tid);
  auto issue = [&](int it) {
    const int chunk = k.nchunk - 1 - it;
    const int st = it % NST;
    const uint32_t bar = smem_u32(&bars[st]);
    mbar_expect_tx(bar, STAGE);
    const int row = k.b * Tn + chunk * TCH;
    const uint32_t d = k.sbase + st * STAGE;
    tma_load_2d(d, &mapA, bar, k.c0, row);
    tma_load_2d(d + BOX, &mapZ, bar, k.c0, row);
    tma_load_2d(d + 2 * BOX, &mapDH, bar, k.c0, row);
    tma_load_2d(d + 3 * BOX, &mapH, bar, k.c0, row - 1);
  };
  if (k.tid == 0)
    for (int it = 0; it < NST - 1 && it < k.nchunk; ++it) issue(it);
  float gz[SV], hfirst[SV];
#pragma unroll
  for (int i = 0; i < SV; ++i) {
    gz[i] = 0.f;
    hfirst[i] = k.live ? h0[(int64_t)k.b * H + k.ch + i] : 0.f;
  }
  T* oa = dAn + (int64_t)k.b * Tn * lddan + k.ch;
  T* oz = dZn + (int64_t)k.b * Tn * lddzn + k.ch;
  for (int it = 0; it < k.nchunk; ++it) {
    const int chunk = k.nchunk - 1 - it;
    __syncthreads();
    if (k.tid == 0 && it + NST - 1 < k.nchunk) issue(it + NST - 1);
    mbar_wait(smem_u32(&bars[it % NST]), (uint32_t)((it / NST) & 1));
    const T* stg = reinterpret_cast<const T*>(smem + (it % NST) * STAGE);
#pragma unroll 1
    for (int sub = TCH / TC - 1; sub >= 0; --sub) {
      const int t0 = chunk * TCH + sub * TC;
      if (t0 >= Tn) continue;
      const T* st = stg + (size_t)sub * TC * CB;       // row u of box g of this group: st + (g * TCH + u) * CB
      if (t0 + TC <= Tn && t0 > 0) {
        // full group that does not hold frame 0: activations and their products of all steps first (scheduling
        // fence), then the carried chain gz -> gam; same operations in the same order per element
        float zh[TC][SV], t1[TC][SV], hmc[TC][SV], go[TC][SV];
#pragma unroll
        for (int u = TC - 1; u >= 0; --u) {
          float an[SV], zn[SV], hp[SV];
          lds2(st + u * CB, k.tid, an);
          lds2(st + (TCH + u) * CB, k.tid, zn);
          lds2(st + (2 * TCH + u) * CB, k.tid, go[u]);
          lds2(st + (3 * TCH + u) * CB, k.tid, hp);
#pragma unroll
          for (int i = 0; i < SV; ++i) {
            const float cc = tanhf_<PRECISE>(an[i]);
            zh[u][i] = sigmoidf_<PRECISE>(zn[i]);
            t1[u][i] = fmaf(-cc, cc, 1.f);
            hmc[u][i] = hp[i] - cc;
          }
        }
        __syncwarp();
#pragma unroll
        for (int u = TC - 1; u >= 0; --u) {
          float da[SV], dz[SV];
#pragma unroll
          for (int i = 0; i < SV; ++i) {
            const float gam = go[u][i] + gz[i];
            gz[i] = zh[u][i] * gam;
            const float omz = 1.f - zh[u][i];
            da[i] = gam * omz * t1[u][i];
            dz[i] = gam * hmc[u][i] * zh[u][i] * omz;
          }
          if (k.live) {
            stg_vec<T>(oa + (int64_t)(t0 + u) * lddan, da);
            stg_vec<T>(oz + (int64_t)(t0 + u) * lddzn, dz);
          }
        }
        continue;
      }
#pragma unroll
      for (int u = TC - 1; u >= 0; --u) {
        const int t = t0 + u;
        if (t < Tn) {
          float an[SV], zn[SV], go[SV], hp[SV], da[SV], dz[SV];
          lds2(st + u * CB, k.tid, an);
          lds2(st + (TCH + u) * CB, k.tid, zn);
          lds2(st + (2 * TCH + u) * CB, k.tid, go);
          lds2(st + (3 * TCH + u) * CB, k.tid, hp);
          if (t == 0) {
#pragma unroll
            for (int i = 0; i < SV; ++i) hp[i] = hfirst[i];
          }
#pragma unroll
          for (int i = 0; i < SV; ++i) {
            const float cc = tanhf_<PRECISE>(an[i]);
            const float zh = sigmoidf_<PRECISE>(zn[i]);
            const float gam = go[i] + gz[i];
            gz[i] = zh * gam;
            const float omz = 1.f - zh;
            da[i] = gam * omz * fmaf(-cc, cc, 1.f);
            dz[i] = gam * (hp[i] - cc) * zh * omz;
          }
          if (k.live) {
            stg_vec<T>(oa + (int64_t)t * lddan, da);
            stg_vec<T>(oz + (int64_t)t * lddzn, dz);
          }
        }
      }
    }
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
  static const int v = [] { const char* e = getenv("SC_SCAN_VEC"); return (e && e[0] == '1') ? 1 : (e && e[0] == '2') ? 2 : 0; }();
  return v;
}

template <typename T, int VEC, bool PRECISE>
static int scan_fwd_tma(const void* G, int64_t ldg, const float* h0, const float* s0, void* Hout, int64_t ldh,
                        float* hT, float* sT, float* Sckpt, int64_t B, int64_t Tn, int64_t H, int train,
                        cudaStream_t st) {
  constexpr int NST = sizeof(T) == 2 ? SC_NST_FWD : 4;     // ring depth (16-bit rows: see SC_NST_FWD above)
  constexpr int smem = NST * 5 * TC * CB * (int)sizeof(T);
  CUtensorMap mapG;
  if (!make_scan_map_gates<T>(&mapG, G, B * Tn, H, ldg, TC)) return SC_E_UNSUP;
  const int cblocks = (int)cdiv(H, CB);
  const unsigned grid = (unsigned)(B * cblocks);
  auto kt = lucy_scan_fwd_tma_kernel<T, VEC, NST, true, PRECISE>;
  auto ks = lucy_scan_fwd_tma_kernel<T, VEC, NST, false, PRECISE>;
  static std::atomic<uint64_t> attr_t{0}, attr_s{0};             // per (instantiation, device)
  {
    const cudaError_t e = train ? ensure_dyn_smem(kt, smem, attr_t) : ensure_dyn_smem(ks, smem, attr_s);
    if (e != cudaSuccess) return (int)e;
  }
  if (train) kt<<<grid, CB / VEC, smem, st>>>(mapG, h0, s0, (T*)Hout, ldh, hT, sT, Sckpt, (int)Tn, (int)H, cblocks);
  else       ks<<<grid, CB / VEC, smem, st>>>(mapG, h0, s0, (T*)Hout, ldh, hT, sT, Sckpt, (int)Tn, (int)H, cblocks);
  SC_LAUNCH_RET();
}

template <typename T, int VEC, bool PRECISE>
static int scan_bwd_tma(const void* G, int64_t ldg, const void* Hout, int64_t ldh, const float* h0,
                        const float* Sckpt, const void* dHout, int64_t lddh, void* dG, int64_t lddg,
                        float* dbias, int64_t B, int64_t Tn, int64_t H, int train, cudaStream_t st) {
  constexpr int ROWS = sizeof(T) == 2 ? SC_SCAN_BWD_ROWS : 8;      // rows per stage
  constexpr int NST = ROWS == 8 ? (sizeof(T) == 2 ? SC_NST_BWD : 3) : 2;
  constexpr int smem = NST * 7 * ROWS * CB * (int)sizeof(T);
  CUtensorMap mapG, mapH, mapDH;
  if (!make_scan_map<T>(&mapG, G, B * Tn, 5 * H, ldg, ROWS) || !make_scan_map<T>(&mapH, Hout, B * Tn, H, ldh, ROWS) ||
      !make_scan_map<T>(&mapDH, dHout, B * Tn, H, lddh, ROWS))
    return SC_E_UNSUP;
  const int cblocks = (int)cdiv(H, CB);
  const unsigned grid = (unsigned)(B * cblocks);
  auto kt = lucy_scan_bwd_tma_kernel<T, VEC, NST, true, PRECISE, ROWS>;
  auto ks = lucy_scan_bwd_tma_kernel<T, VEC, NST, false, PRECISE, ROWS>;
  static std::atomic<uint64_t> attr_t{0}, attr_s{0};             // per (instantiation, device)
  {
    const cudaError_t e = train ? ensure_dyn_smem(kt, smem, attr_t) : ensure_dyn_smem(ks, smem, attr_s);
    if (e != cudaSuccess) return (int)e;
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


// ---- split-scan dispatch (SC_E_UNSUP -> caller uses the simple kernels) ----
template <typename T>
static bool split_ok(std::initializer_list<const void*> ptrs, std::initializer_list<int64_t> lds, int64_t H, int64_t B,
                     int64_t Tn) {
  constexpr int per16 = 16 / (int)sizeof(T);
  if (H % per16 != 0 || Tn == 0 || B * Tn >= ((int64_t)1 << 31) - TC - 1) return false;
  for (const void* p : ptrs) if (!aligned16(p)) return false;
  for (int64_t l : lds) if (l % per16) return false;
  return get_encode() != nullptr;
}
template <typename K>
static int set_smem(K kern, int smem) {
  return (int)cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
}

template <typename T, bool PRECISE>
static int sscan_fwd_tma_t(const void* k, const void* v, const void* q, int64_t ldg, const void* addend, int64_t ldadd,
                           const float* s0, void* A, int64_t lda, float* S_all, float* sT, int64_t B, int64_t Tn,
                           int64_t H, int train, cudaStream_t st) {
  if (!split_ok<T>({k, v, q, addend, A, S_all}, {ldg, ldadd, lda}, H, B, Tn)) return SC_E_UNSUP;
  constexpr int ROWS = sizeof(T) == 2 ? SC_SSCAN_ROWS : 8;        // rows per stage: 32 KB per stage either way
  constexpr int NST = 3;
  constexpr int smem = NST * 4 * ROWS * CB * (int)sizeof(T);
  CUtensorMap mk, mv, mq, ma;
  if (!make_scan_map<T>(&mk, k, B * Tn, H, ldg, ROWS) || !make_scan_map<T>(&mv, v, B * Tn, H, ldg, ROWS) ||
      !make_scan_map<T>(&mq, q, B * Tn, H, ldg, ROWS) || !make_scan_map<T>(&ma, addend, B * Tn, H, ldadd, ROWS))
    return SC_E_UNSUP;
  const int cblocks = (int)cdiv(H, CB);
  const unsigned grid = (unsigned)(B * cblocks);
  auto kt = sscan_fwd_tma_kernel<T, NST, true, PRECISE, ROWS>;
  auto ks = sscan_fwd_tma_kernel<T, NST, false, PRECISE, ROWS>;
  int e = set_smem(kt, smem); if (e) return e;
  e = set_smem(ks, smem); if (e) return e;
  if (train) kt<<<grid, SPLIT_THREADS, smem, st>>>(mk, mv, mq, ma, s0, (T*)A, lda, S_all, sT, (int)Tn, (int)H, cblocks);
  else       ks<<<grid, SPLIT_THREADS, smem, st>>>(mk, mv, mq, ma, s0, (T*)A, lda, S_all, sT, (int)Tn, (int)H, cblocks);
  SC_LAUNCH_RET();
}
int sscan_fwd_tma_dispatch(const void* k, const void* v, const void* q, int64_t ldg, const void* addend, int64_t ldadd,
                           const float* s0, void* A, int64_t lda, float* S_all, float* sT, int64_t B, int64_t T,
                           int64_t H, int dtype, int train, cudaStream_t st) {
  if (dtype == SC_BF16) return sscan_fwd_tma_t<bf16, false>(k, v, q, ldg, addend, ldadd, s0, A, lda, S_all, sT, B, T, H, train, st);
  if (dtype == SC_F32) return sscan_fwd_tma_t<float, true>(k, v, q, ldg, addend, ldadd, s0, A, lda, S_all, sT, B, T, H, train, st);
  return SC_E_DTYPE;
}

template <typename T, bool PRECISE>
static int sscan_bwd_tma_t(const void* k, const void* v, const void* q, int64_t ldg, const float* S_all, const float* s0,
                           const void* dA, int64_t ldda, void* dk, void* dv, void* dq, int64_t lddg, float* dsum, int64_t B,
                           int64_t Tn, int64_t H, int train, cudaStream_t st) {
  (void)s0;                                                      // interval 0's checkpoint IS the initial state
  if (!split_ok<T>({k, v, q, dA}, {ldg, ldda}, H, B, Tn) || (lddg % 2) ||
      (((uintptr_t)dk | (uintptr_t)dv | (uintptr_t)dq) & (2 * sizeof(T) - 1)) || ((uintptr_t)S_all & 7))
    return SC_E_UNSUP;
  constexpr int ROWS = sizeof(T) == 2 ? SC_SSCAN_ROWS : 8;        // rows per stage: 32 KB per stage either way
  constexpr int NST = 3;
  constexpr int smem = NST * 4 * ROWS * CB * (int)sizeof(T);
  CUtensorMap mk, mv, mq, mda;
  if (!make_scan_map<T>(&mk, k, B * Tn, H, ldg, ROWS) || !make_scan_map<T>(&mv, v, B * Tn, H, ldg, ROWS) ||
      !make_scan_map<T>(&mq, q, B * Tn, H, ldg, ROWS) || !make_scan_map<T>(&mda, dA, B * Tn, H, ldda, ROWS))
    return SC_E_UNSUP;
  const int cblocks = (int)cdiv(H, CB);
  const unsigned grid = (unsigned)(B * cblocks);
  auto kt = sscan_bwd_tma_kernel<T, NST, true, PRECISE, ROWS>;
  auto ks = sscan_bwd_tma_kernel<T, NST, false, PRECISE, ROWS>;
  int e = set_smem(kt, smem); if (e) return e;
  e = set_smem(ks, smem); if (e) return e;
  if (train) kt<<<grid, SPLIT_THREADS, smem, st>>>(mk, mv, mq, mda, S_all, (T*)dk, (T*)dv, (T*)dq, lddg, dsum, (int)Tn, (int)H, cblocks);
  else       ks<<<grid, SPLIT_THREADS, smem, st>>>(mk, mv, mq, mda, S_all, (T*)dk, (T*)dv, (T*)dq, lddg, dsum, (int)Tn, (int)H, cblocks);
  SC_LAUNCH_RET();
}
int sscan_bwd_tma_dispatch(const void* k, const void* v, const void* q, int64_t ldg, const float* S_all, const float* s0,
                           const void* dA, int64_t ldda, void* dk, void* dv, void* dq, int64_t lddg, float* dsum, int64_t B,
                           int64_t T, int64_t H, int dtype, int train, cudaStream_t st) {
  if (dtype == SC_BF16) return sscan_bwd_tma_t<bf16, false>(k, v, q, ldg, S_all, s0, dA, ldda, dk, dv, dq, lddg, dsum, B, T, H, train, st);
  if (dtype == SC_F32) return sscan_bwd_tma_t<float, true>(k, v, q, ldg, S_all, s0, dA, ldda, dk, dv, dq, lddg, dsum, B, T, H, train, st);
  return SC_E_DTYPE;
}

template <typename T, bool PRECISE>
static int hscan_fwd_tma_t(const void* An, int64_t ldan, const void* Zn, int64_t ldzn, const float* h0, void* Hout,
                           int64_t ldh, float* hT, int64_t B, int64_t Tn, int64_t H, cudaStream_t st) {
  if (!split_ok<T>({An, Zn}, {ldan, ldzn}, H, B, Tn) || (ldh % 2) || ((uintptr_t)Hout & (2 * sizeof(T) - 1))) return SC_E_UNSUP;
  constexpr int TCH = sizeof(T) == 2 ? SC_HSCAN_FWD_ROWS : 16;    // rows per stage: 32 KB per stage either way
  constexpr int NST = 3;
  constexpr int smem = NST * 2 * TCH * CB * (int)sizeof(T);
  CUtensorMap ma, mz;
  if (!make_scan_map<T>(&ma, An, B * Tn, H, ldan, TCH) || !make_scan_map<T>(&mz, Zn, B * Tn, H, ldzn, TCH)) return SC_E_UNSUP;
  const int cblocks = (int)cdiv(H, CB);
  auto kk = hscan_fwd_tma_kernel<T, NST, PRECISE, TCH>;
  int e = set_smem(kk, smem); if (e) return e;
  kk<<<(unsigned)(B * cblocks), SPLIT_THREADS, smem, st>>>(ma, mz, h0, (T*)Hout, ldh, hT, (int)Tn, (int)H, cblocks);
  SC_LAUNCH_RET();
}
int hscan_fwd_tma_dispatch(const void* An, int64_t ldan, const void* Zn, int64_t ldzn, const float* h0, void* Hout,
                           int64_t ldh, float* hT, int64_t B, int64_t T, int64_t H, int dtype, cudaStream_t st) {
  if (dtype == SC_BF16) return hscan_fwd_tma_t<bf16, false>(An, ldan, Zn, ldzn, h0, Hout, ldh, hT, B, T, H, st);
  if (dtype == SC_F32) return hscan_fwd_tma_t<float, true>(An, ldan, Zn, ldzn, h0, Hout, ldh, hT, B, T, H, st);
  return SC_E_DTYPE;
}

template <typename T, bool PRECISE>
static int hscan_bwd_tma_t(const void* An, int64_t ldan, const void* Zn, int64_t ldzn, const void* Hout, int64_t ldh,
                           const float* h0, const void* dHout, int64_t lddh, void* dAn, int64_t lddan, void* dZn,
                           int64_t lddzn, int64_t B, int64_t Tn, int64_t H, cudaStream_t st) {
  if (!split_ok<T>({An, Zn, Hout, dHout}, {ldan, ldzn, ldh, lddh}, H, B, Tn) || (lddan % 2) || (lddzn % 2) ||
      (((uintptr_t)dAn | (uintptr_t)dZn) & (2 * sizeof(T) - 1)))
    return SC_E_UNSUP;
  constexpr int TCH = sizeof(T) == 2 ? SC_HSCAN_BWD_ROWS : 8;     // rows per stage: 32 KB per stage either way (two CTAs per SM)
  constexpr int NST = 3;
  constexpr int smem = NST * 4 * TCH * CB * (int)sizeof(T);
  CUtensorMap ma, mz, mdh, mh;
  if (!make_scan_map<T>(&ma, An, B * Tn, H, ldan, TCH) || !make_scan_map<T>(&mz, Zn, B * Tn, H, ldzn, TCH) ||
      !make_scan_map<T>(&mdh, dHout, B * Tn, H, lddh, TCH) || !make_scan_map<T>(&mh, Hout, B * Tn, H, ldh, TCH))
    return SC_E_UNSUP;
  const int cblocks = (int)cdiv(H, CB);
  auto kk = hscan_bwd_tma_kernel<T, NST, PRECISE, TCH>;
  int e = set_smem(kk, smem); if (e) return e;
  kk<<<(unsigned)(B * cblocks), SPLIT_THREADS, smem, st>>>(ma, mz, mdh, mh, h0, (T*)dAn, lddan, (T*)dZn, lddzn, (int)Tn, (int)H, cblocks);
  SC_LAUNCH_RET();
}
int hscan_bwd_tma_dispatch(const void* An, int64_t ldan, const void* Zn, int64_t ldzn, const void* Hout, int64_t ldh,
                           const float* h0, const void* dHout, int64_t lddh, void* dAn, int64_t lddan, void* dZn,
                           int64_t lddzn, int64_t B, int64_t T, int64_t H, int dtype, cudaStream_t st) {
  if (dtype == SC_BF16) return hscan_bwd_tma_t<bf16, false>(An, ldan, Zn, ldzn, Hout, ldh, h0, dHout, lddh, dAn, lddan, dZn, lddzn, B, T, H, st);
  if (dtype == SC_F32) return hscan_bwd_tma_t<float, true>(An, ldan, Zn, ldzn, Hout, ldh, h0, dHout, lddh, dAn, lddan, dZn, lddzn, B, T, H, st);
  return SC_E_DTYPE;
}

}  // namespace sc
