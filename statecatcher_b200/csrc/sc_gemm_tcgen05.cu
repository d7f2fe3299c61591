// K1 (bf16 training path): persistent, warp-specialised tcgen05 + TMA GEMM for sm_100a.
//
// Replaces the cuBLAS calls behind nn.Linear forward/backward for input_proj, W_fused and
// output_proj (lucyrnn.py:15, 23, 85; 113, 116, 186).  One kernel template serves the three
// contractions of a linear layer, all written as  D[i,j] = sum_r A(i,r) * B(j,r):
//   fwd   : i=m, j=n, r=k   A = X  [m,k] (K-major)    B = W  [n,k] (K-major)   D = Y  (+bias[j])
//   dgrad : i=m, j=k, r=n   A = dY [m,n] (K-major)    B = W  [n,k] (MN-major)  D = dX
//   wgrad : i=n, j=k, r=m   A = dY [m,n] (MN-major)   B = X  [m,k] (MN-major)  D = dW (fp32, split-R)
// "MN-major" operands are consumed straight from their row-major activation layout through
// the UMMA descriptor's major bit — no transposed copies of activations are ever made.
//
// Structure (one persistent CTA per SM, 320 threads, clusters of two CTAs = one MMA pair):
//   warp 0   : TMA producer — cp.async.bulk.tensor 128B-swizzled tiles into the operand ring
//              (pair mode: 5 stages of [A 128x64 | half of B 128x64]; both CTAs' bytes are counted
//              on the leader's mbarrier)
//   warp 1   : MMA issuer   — one thread of the LEADER CTA issues tcgen05.mma.cta_group::2
//              (M=256 over the two SMs, N=256, K=16, bf16->fp32) into one of two 256-column TMEM
//              accumulator stages; tcgen05.commit (multicast to both CTAs) releases smem stages /
//              publishes the accumulator through mbarriers
//   warps 2-9: epilogue     — tcgen05.ld (32 lanes x 32 columns per instruction), bias add
//              (broadcast 16-byte loads through L1), bf16 [32 x 64] boxes of 128-byte rows through
//              shared memory to a TMA store (or fp32 TMA reduce-add for split-R wgrad); both CTAs'
//              epilogue warps release the accumulator on the leader's barrier
// The epilogue of tile i overlaps the main loop of tile i+1 (double-buffered TMEM).
#include "sc_common.cuh"
#include "sc_tma.cuh"
#include <stdlib.h>

namespace sc {

constexpr int TM = 128;          // tile rows   (UMMA M)
constexpr int TN = 256;          // tile cols   (UMMA N)
constexpr int TK = 64;           // reduction elements per stage = one 128-byte swizzle row
constexpr int UK = 16;           // UMMA K for 16-bit inputs
constexpr int A_BYTES = TM * TK * 2;              // 16 KB
constexpr int RING_BYTES = 192 * 1024;            // operand ring region: 4 stages of 48 KB; pair mode: 5 of 32 KB + 2nd set of store boxes
constexpr int ATOM_BYTES = TK * 128;              // one [TK x 64-element] MN-major box = 8 KB
constexpr int EPI_WARPS = 8;                     // two warps per TMEM lane quarter, 128 columns each
constexpr int GEMM_THREADS = 64 + EPI_WARPS * 32;
constexpr int EPI_STAGE_BYTES = 32 * 64 * 2;      // per epilogue warp: one [32 rows x 64 cols] bf16 box (128-byte rows) for the TMA store
constexpr int EPI_OFF = RING_BYTES;               // first set of store boxes: right behind the ring (1024-B aligned: SWIZZLE_128B)
constexpr int BAR_OFF = EPI_OFF + EPI_WARPS * EPI_STAGE_BYTES;   // mbarriers + TMEM pointer
static_assert(EPI_OFF % 1024 == 0 && EPI_STAGE_BYTES % 1024 == 0, "store boxes must be 1024-byte aligned (128-byte swizzle pattern)");
constexpr int SMEM_BYTES = 1024 /*align slack*/ + BAR_OFF + 256;
static_assert(SMEM_BYTES <= 232448, "227 KB of shared memory per CTA");
constexpr uint32_t TMEM_COLS = 512;
#ifndef SC_GEMM_EPI_BUFS
#define SC_GEMM_EPI_BUFS 2   // store boxes per epilogue warp in pair mode: the second set fills the 32 KB tail of the ring region
#endif

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout), 128-byte swizzle:
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4   [32,46) stride byte offset >> 4
//   [46,48) version = 1 (sm_100)   [61,64) layout type = 2 (SWIZZLE_128B)
// K-major tile  [rows][64 elems] : 8-row atoms of 1024 B stacked along rows -> SBO = 1024, LBO unused.
// MN-major tile : boxes of [TK rows][64 elems]; SBO = 1024 (next 8 reduction rows),
//                 LBO = ATOM_BYTES (next 64 elements along M/N = next TMA box).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct GemmParams {
  int64_t I, J, R;          // output rows, output cols, reduction length
  int tiles_i, tiles_j, splits, kb_per_split, kb_total;
  void* D; int64_t ldd;
  const float* bias;
};

// EPI: 0 = bf16 store (+bias), 1 = fp32 store (+bias), 2 = fp32 atomic accumulate (split-R)
// PAIR: the two CTAs of a cluster issue ONE tcgen05.mma.cta_group::2 (M = 256 across the two
// SMs, N = 256): each CTA stages its own 128 rows of A and only HALF of the B tile — the tensor
// cores read the other half from the peer's shared memory — so a stage is 32 KB instead of 48 KB
// (shared-memory write and operand-read traffic per flop drop by a third, five stages and the second set of store boxes fit in the ring region).
// !PAIR: each CTA issues its own M = 128 MMA on a full copy of the B tile (TMA multicast).
template <bool A_MN, bool B_MN, int EPI, bool PAIR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ CUtensorMap mapD, const GemmParams p) {
  constexpr int B_BYTES = (PAIR ? TN / 2 : TN) * TK * 2;            // 16 KB / 32 KB
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = PAIR ? 5 : 4;
  constexpr int EPI_BUFS = PAIR ? SC_GEMM_EPI_BUFS : 1;              // store boxes per epilogue warp (sets beyond the first: tail of the ring region)
  static_assert(STAGES * STAGE_BYTES + (EPI_BUFS - 1) * EPI_WARPS * EPI_STAGE_BYTES <= RING_BYTES, "ring region");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;       // SWIZZLE_128B: 1024-B aligned tiles
  const uint32_t bar0 = base + BAR_OFF;
  // barrier slots (8 B each): full[STAGES], empty[STAGES], tfull[2], tempty[2], then tmem ptr
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (STAGES + s); };
  auto tfull = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
  auto tempty = [&](int s) { return bar0 + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + BAR_OFF + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    // !PAIR: empty is released by the MMA commits of BOTH CTAs (each multicasts half of B into the other).
    // PAIR: the leader's single commit is multicast to both CTAs; the leader's tempty collects the
    // epilogue warps of both CTAs (the peer's full/tempty barriers are never waited on).
    for (int s = 0; s < STAGES; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), PAIR ? 1 : 2); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), PAIR ? 2 * EPI_WARPS : EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
  }
  if (warp == 1) { if (PAIR) tmem_alloc_pair(tmem_slot, TMEM_COLS); else tmem_alloc(tmem_slot, TMEM_COLS); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // peer's barriers are initialised before any multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // A cluster = 2 CTAs working on vertically adjacent tiles (ti = 2*pair + rank, same tj): they
  // share the B tile — each CTA fetches half of it and TMA-multicasts it into both — which cuts
  // L2->SM operand traffic per k-block from 48 KB to 32 KB per SM (the single-CTA version ran
  // at the ~12 TB/s L2 limit: 23 GB of operand reads per 2-TFLOP GEMM).
  const uint32_t rank = cluster_ctarank();
  const int pairs_i = (p.tiles_i + 1) / 2;
  const int64_t total = (int64_t)pairs_i * p.tiles_j * p.splits;
  const int64_t w0 = blockIdx.x >> 1, wstep = gridDim.x >> 1;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int64_t w = w0; w < total; w += wstep) {
        const int64_t ntile = (int64_t)pairs_i * p.tiles_j;
        const int split = (int)(w / ntile);          // split-major: pairs running together share the reduction range
        const int64_t tile = w % ntile;
        const int tj = (int)(tile % p.tiles_j), ti = 2 * (int)(tile / p.tiles_j) + (int)rank;
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty(stage), phase ^ 1);
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
          if (PAIR) {
            // both CTAs' loads are counted on the leader's barrier; only the leader expects them
            if (rank == 0) mbar_expect_tx(full(stage), 2 * STAGE_BYTES);
            if (!A_MN) {
              tma_load_2d_pair(sa, &mapA, full(stage), kb * TK, ti * TM);
            } else {
#pragma unroll
              for (int a = 0; a < TM / 64; ++a) tma_load_2d_pair(sa + a * ATOM_BYTES, &mapA, full(stage), ti * TM + a * 64, kb * TK);
            }
            // this CTA's half of the B tile (columns rank*128 .. +128 of the 256-wide tile)
            if (!B_MN) {
              tma_load_2d_pair(sb, &mapB, full(stage), kb * TK, tj * TN + (int)rank * (TN / 2));
            } else {
#pragma unroll
              for (int b = 0; b < TN / 128; ++b)
                tma_load_2d_pair(sb + b * ATOM_BYTES, &mapB, full(stage), tj * TN + ((int)rank * (TN / 128) + b) * 64, kb * TK);
            }
          } else {
          mbar_expect_tx(full(stage), STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sa, &mapA, full(stage), kb * TK, ti * TM);
          } else {
#pragma unroll
            for (int a = 0; a < TM / 64; ++a) tma_load_2d(sa + a * ATOM_BYTES, &mapA, full(stage), ti * TM + a * 64, kb * TK);
          }
          // this CTA's half of the B tile, multicast to both CTAs of the pair
          if (!B_MN) {
            tma_load_2d_mc(sb + rank * (B_BYTES / 2), &mapB, full(stage), kb * TK, tj * TN + (int)rank * (TN / 2), 0x3);
          } else {
#pragma unroll
            for (int b = 0; b < TN / 128; ++b) {
              const int bb = (int)rank * (TN / 128) + b;
              tma_load_2d_mc(sb + bb * ATOM_BYTES, &mapB, full(stage), tj * TN + bb * 64, kb * TK, 0x3);
            }
          }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && (!PAIR || rank == 0)) {
      // instruction descriptor: c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
      // a_major bit15, b_major bit16 (1 = MN-major), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                             ((uint32_t)(TN >> 3) << 17) | ((uint32_t)((PAIR ? 2 * TM : TM) >> 4) << 24);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t w = w0; w < total; w += wstep) {
        const int split = (int)(w / ((int64_t)pairs_i * p.tiles_j));
        const int kb0 = split * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
        mbar_wait(tempty(acc), acc_phase ^ 1);          // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * TN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full(stage), phase);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < TK / UK; ++k) {
            const uint64_t ad = A_MN ? make_desc(sa + k * UK * 128, ATOM_BYTES, 1024)
                                     : make_desc(sa + k * UK * 2, 0, 1024);
            const uint64_t bd = B_MN ? make_desc(sb + k * UK * 128, ATOM_BYTES, 1024)
                                     : make_desc(sb + k * UK * 2, 0, 1024);
            if (PAIR) umma_bf16_pair(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          // stage reusable (here AND in the peer) once these MMAs retire
          if (PAIR) umma_commit_pair_mc(empty(stage), 0x3); else umma_commit_mc(empty(stage), 0x3);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_pair_mc(tfull(acc), 0x3); else umma_commit(tfull(acc));   // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                             // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                   // which 128-column half of the tile
    // bias: every lane of a warp needs the same 32 values per chunk, so they come as broadcast 16-byte loads through L1
    // (they used to be staged in 2 KB of shared memory behind a named barrier; the store boxes need that room now)
    const bool bias_vec = (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0;
    auto add_bias = [&](uint32_t (&r)[32], int64_t col) {
      if (bias_vec && col + 32 <= p.J) {
        const float4* bv = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const float4 b4 = __ldg(bv + v);
          r[v * 4 + 0] = __float_as_uint(__uint_as_float(r[v * 4 + 0]) + b4.x);
          r[v * 4 + 1] = __float_as_uint(__uint_as_float(r[v * 4 + 1]) + b4.y);
          r[v * 4 + 2] = __float_as_uint(__uint_as_float(r[v * 4 + 2]) + b4.z);
          r[v * 4 + 3] = __float_as_uint(__uint_as_float(r[v * 4 + 3]) + b4.w);
        }
      } else {
#pragma unroll
        for (int v = 0; v < 32; ++v)
          if (col + v < p.J) r[v] = __float_as_uint(__uint_as_float(r[v]) + __ldg(p.bias + col + v));
      }
    };
    int acc = 0; uint32_t acc_phase = 0;
    uint32_t ebox = 0;                                  // store boxes issued by this warp (selects the staging buffer)
    for (int64_t w = w0; w < total; w += wstep) {
      const int64_t tile = w % ((int64_t)pairs_i * p.tiles_j);
      const int tj = (int)(tile % p.tiles_j), ti = 2 * (int)(tile / p.tiles_j) + (int)rank;
      const int64_t col0 = (int64_t)tj * TN;
      mbar_wait(tfull(acc), acc_phase);
      tc_fence_after();
      const int64_t row = (int64_t)ti * TM + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * TN;
      if (EPI == 0) {
        // bf16 output: bias, pack, stage the warp's [32 rows x 64 cols] box in shared memory and hand it to the TMA
        // engine — full 128-byte row segments (r01: 32 scattered 16-byte stores per instruction, 1.9 TB/s; r02 first
        // [32 x 32] boxes of 64-byte rows: a reduction-starved GEMM such as layer 0's K = 80 projection stored at 3.1
        // TB/s, ~97 ns per box per SM whatever the eight warps did — the store engine's per-box / per-row cost — so
        // the boxes carry twice the bytes now); the tensor map clips rows >= I and columns >= J.
#pragma unroll 1
        for (int cc = 0; cc < TN / 128; ++cc) {
          const int c = half * (TN / 64) + cc * 2;          // first of this box's two 32-column TMEM chunks
          uint32_t ra[32], rb[32];
          tmem_ld32(taddr + c * 32, ra);
          tmem_ld32(taddr + (c + 1) * 32, rb);
          tmem_ld_wait();
          const int64_t col = col0 + c * 32;
          if (p.bias != nullptr) { add_bias(ra, col); add_bias(rb, col + 32); }
          const uint32_t eslot = EPI_BUFS > 1 ? ebox % EPI_BUFS : 0u;
          const uint32_t stg = eslot ? base + STAGES * STAGE_BYTES + ((eslot - 1) * EPI_WARPS + (uint32_t)(warp - 2)) * EPI_STAGE_BYTES
                                     : base + EPI_OFF + (uint32_t)(warp - 2) * EPI_STAGE_BYTES;
          ++ebox;
          if (lane == 0) bulk_wait_read<EPI_BUFS - 1>();         // the box written EPI_BUFS stores ago has left shared memory
          __syncwarp();
          // 128-byte rows, SWIZZLE_128B (16-byte chunk ^= row & 7): the 8 lanes of a store phase land in 8 different
          // bank groups; unswizzled, rows 128 B apart would put all of them on the same four banks
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float f0 = __uint_as_float(v < 4 ? ra[v * 8 + 2 * e] : rb[(v - 4) * 8 + 2 * e]);
              const float f1 = __uint_as_float(v < 4 ? ra[v * 8 + 2 * e + 1] : rb[(v - 4) * 8 + 2 * e + 1]);
              __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
              pk[e] = *reinterpret_cast<uint32_t*>(&h);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 128 + ((v ^ (lane & 7)) * 16)), "r"(pk[0]), "r"(pk[1]),
                         "r"(pk[2]), "r"(pk[3]) : "memory");
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && col < p.J && (int64_t)ti * TM + q * 32 < p.I) {
            tma_store_2d(&mapD, stg, (int)col, ti * TM + q * 32);
            bulk_commit();
          }
        }
      } else {
#pragma unroll 1
      for (int cc = 0; cc < TN / 64; ++cc) {
        const int c = half * (TN / 64) + cc;
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        const int64_t col = col0 + c * 32;
        if (EPI == 2) {
          // split-R partial tile: fp32 TMA reduce-add of [32 rows x 16 cols] boxes (64-byte row
          // segments reduced in L2) instead of 32 scattered 4-byte atomics per lane per chunk
          const uint32_t stg = base + EPI_OFF + (uint32_t)(warp - 2) * EPI_STAGE_BYTES;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (lane == 0) bulk_wait_read0();
            __syncwarp();
#pragma unroll
            for (int v = 0; v < 4; ++v)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + lane * 64 + ((v ^ ((lane >> 1) & 3)) * 16)), "r"(r[hh * 16 + v * 4 + 0]),
                           "r"(r[hh * 16 + v * 4 + 1]), "r"(r[hh * 16 + v * 4 + 2]), "r"(r[hh * 16 + v * 4 + 3]) : "memory");
            fence_async_smem();
            __syncwarp();
            if (lane == 0 && col + hh * 16 < p.J && (int64_t)ti * TM + q * 32 < p.I) {
              tma_reduce_add_2d(&mapD, stg, (int)col + hh * 16, ti * TM + q * 32);
              bulk_commit();
            }
          }
        } else if (row < p.I && col < p.J) {
          // EPI == 1: fp32 output, 16-byte stores straight from registers
          if (p.bias != nullptr) add_bias(r, col);
          float* out = reinterpret_cast<float*>(p.D) + row * p.ldd + col;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            if (col + v * 4 < p.J) {                  // J is a multiple of 4
              *reinterpret_cast<float4*>(out + v * 4) =
                  make_float4(__uint_as_float(r[v * 4 + 0]), __uint_as_float(r[v * 4 + 1]),
                              __uint_as_float(r[v * 4 + 2]), __uint_as_float(r[v * 4 + 3]));
            }
          }
        }
      }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_leader(tempty(acc)); else mbar_arrive(tempty(acc)); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (EPI != 1 && lane == 0) bulk_wait0();          // all TMA stores / reductions of this warp have completed
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                   // no CTA leaves while the peer can still signal it
  if (warp == 1) {
    __syncwarp();
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side ----------
// 2-D bf16 tensor map over a row-major matrix [rows, cols] (cols contiguous, row stride ld
// elements) with a [box_rows x 64-element] box and 128-byte swizzle; OOB reads return zeros.
static bool make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// bf16 output [rows, cols] (row stride ld): [32 x 64] store boxes (128-byte rows), 128-byte swizzle
static bool make_store_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// fp32 output [rows, cols] (row stride ld): [32 x 16] reduce-add boxes (64-byte rows), 64-byte swizzle
static bool make_reduce_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {16, 32};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool tc_common_ok(const void* a, const void* b, const void* d, int64_t lda, int64_t ldb, int64_t ldd,
                         int64_t I, int64_t J, int64_t R, int d_align_elems, bool forced = false) {
  if (I < 1 || J < 8 || R < 8) return false;
  if ((lda % 8) || (ldb % 8) || (ldd % d_align_elems) || (J % 8)) return false;
  if (!aligned16(a) || !aligned16(b) || !aligned16(d)) return false;
  if (I >= ((int64_t)1 << 31) || J >= ((int64_t)1 << 31) || R >= ((int64_t)1 << 31)) return false;
  // worth the fixed cost only when a tile is reasonably filled — unless the caller insists (impl = 2: the fp32 six-block
  // path must take the SAME kernel for every M, or a segment cut differently would not reproduce the same bits)
  return (forced || I * J >= 64 * 64) && get_encode() != nullptr;
}

template <bool A_MN, bool B_MN, int EPI>
static int launch_tc(const void* A, int64_t lda, const void* B, int64_t ldb, void* D, int64_t ldd,
                     const float* bias, int64_t I, int64_t J, int64_t R, int splits, cudaStream_t st) {
  CUtensorMap mapA, mapB;
  bool ok = A_MN ? make_map(&mapA, A, R, I, lda, TK) : make_map(&mapA, A, I, R, lda, TM);
  ok = ok && (B_MN ? make_map(&mapB, B, R, J, ldb, TK) : make_map(&mapB, B, J, R, ldb, TN / 2));
  CUtensorMap mapD = mapA;                               // only dereferenced by the bf16 epilogue
  if (EPI == 0) ok = ok && make_store_map(&mapD, D, I, J, ldd);
  if (EPI == 2) ok = ok && make_reduce_map(&mapD, D, I, J, ldd);
  if (!ok) return SC_E_UNSUP;
  GemmParams p;
  p.I = I; p.J = J; p.R = R;
  p.tiles_i = (int)cdiv(I, TM); p.tiles_j = (int)cdiv(J, TN);
  p.kb_total = (int)cdiv(R, TK);
  if (splits < 1) splits = 1;
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (int)cdiv(p.kb_total, splits);
  p.splits = (int)cdiv(p.kb_total, p.kb_per_split);
  p.D = D; p.ldd = ldd; p.bias = bias;
  // CTA-pair MMA by default.  Timed alone at 192000 x 5120 x 1024: fwd 1.445 -> 1.367 ms, dgrad
  // 1.398 -> 1.328 ms, wgrad 1.392 -> 1.452 ms; inside the power-capped training step all three
  // gain (GEMM total 26.7 -> 25.3 ms per step: a third less shared-memory traffic per flop is
  // also less energy per flop).  SC_GEMM_PAIR=0 selects the multicast variant (A/B measurements).
  static const bool pair = [] { const char* e = getenv("SC_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  static std::atomic<uint64_t> attr_done{0};                      // per (instantiation, device); see ensure_dyn_smem
  {
    cudaError_t e = ensure_dyn_smem(pair ? gemm_tc_kernel<A_MN, B_MN, EPI, true> : gemm_tc_kernel<A_MN, B_MN, EPI, false>,
                                    SMEM_BYTES, attr_done);
    if (e != cudaSuccess) return (int)e;
  }
  const int64_t total = (int64_t)((p.tiles_i + 1) / 2) * p.tiles_j * p.splits;   // work items per CTA pair
  const int64_t clusters = total < num_sms() / 2 ? total : num_sms() / 2;
  const int grid = (int)(2 * clusters);
  if (pair) gemm_tc_kernel<A_MN, B_MN, EPI, true><<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(mapA, mapB, mapD, p);
  else gemm_tc_kernel<A_MN, B_MN, EPI, false><<<grid, GEMM_THREADS, SMEM_BYTES, st>>>(mapA, mapB, mapD, p);
  SC_LAUNCH_RET();
}

__global__ void zero2d_kernel(float* __restrict__ p, int64_t ld, int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[(i / cols) * ld + (i % cols)] = 0.f;
}

bool tc_gemm_fwd_ok(int64_t lda, int64_t ldw, int64_t ldy, int64_t M, int64_t N, int64_t K, int in_dtype,
                    int out_dtype, const void* A, const void* W, const void* Y, bool forced) {
  if (in_dtype != SC_BF16) return false;
  return tc_common_ok(A, W, Y, lda, ldw, ldy, M, N, K, out_dtype == SC_BF16 ? 8 : 4, forced);
}
int tc_gemm_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* Y, int64_t ldy,
                int64_t M, int64_t N, int64_t K, int out_dtype, cudaStream_t st) {
  if (out_dtype == SC_BF16) return launch_tc<false, false, 0>(A, lda, W, ldw, Y, ldy, bias, M, N, K, 1, st);
  return launch_tc<false, false, 1>(A, lda, W, ldw, Y, ldy, bias, M, N, K, 1, st);
}

bool tc_gemm_dgrad_ok(int64_t lddy, int64_t ldw, int64_t ldda, int64_t M, int64_t N, int64_t K, int in_dtype,
                      int out_dtype, const void* dY, const void* W, const void* dA, bool forced) {
  if (in_dtype != SC_BF16) return false;
  return tc_common_ok(dY, W, dA, lddy, ldw, ldda, M, K, N, out_dtype == SC_BF16 ? 8 : 4, forced);
}
int tc_gemm_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw, void* dA, int64_t ldda,
                  int64_t M, int64_t N, int64_t K, int out_dtype, cudaStream_t st) {
  // dA[m,k] = sum_n dY[m,n] W[n,k] : A = dY (K-major over n), B(j=k, r=n) = W[n,k] (MN-major)
  if (out_dtype == SC_BF16) return launch_tc<false, true, 0>(dY, lddy, W, ldw, dA, ldda, nullptr, M, K, N, 1, st);
  return launch_tc<false, true, 1>(dY, lddy, W, ldw, dA, ldda, nullptr, M, K, N, 1, st);
}

bool tc_gemm_wgrad_ok(int64_t lddy, int64_t lda, int64_t lddw, int64_t M, int64_t N, int64_t K, int in_dtype,
                      const void* dY, const void* A, const void* dW) {
  if (in_dtype != SC_BF16) return false;
  // N (rows of dW) comes from dY's columns via TMA: needs N % 8 for the map's 16-B row stride
  return (N % 8 == 0) && tc_common_ok(dY, A, dW, lddy, lda, lddw, N, K, M, 4) && M >= 64;
}
int tc_gemm_wgrad(const void* dY, int64_t lddy, const void* A, int64_t lda, float* dW, int64_t lddw,
                  int64_t M, int64_t N, int64_t K, int accumulate, cudaStream_t st) {
  // dW[n,k] = sum_m dY[m,n] A[m,k] : both operands MN-major, reduction over m, split across CTAs
  // Split the M-long reduction so the (tile pair x split) items fill the 74 CTA pairs evenly.
  // Items are ordered split-major: the pairs running at the same time work on (almost) all
  // tiles over the SAME reduction range, so every dY / X tile is fetched from HBM once and
  // served to its other users from L2; with that ordering more splits only cost the (TMA
  // reduce-add) epilogue, so the count is chosen to minimise last-wave waste.
  const int64_t pairs = cdiv(cdiv(N, TM), 2) * cdiv(K, TN);
  const int64_t kb = cdiv(M, TK);
  const int64_t clusters = num_sms() / 2;
  int64_t max_s = kb / 32;
  if (max_s < 1) max_s = 1;
  if (max_s > 32) max_s = 32;
  int64_t splits = 1;
  double best = 1e30;
  for (int64_t sidx = 1; sidx <= max_s; ++sidx) {
    const int64_t kps = cdiv(kb, sidx);
    const int64_t items = pairs * cdiv(kb, kps);
    const double cost = (double)cdiv(items, clusters) * ((double)kps + 8.0);   // rounds x (mainloop + ~8 k-blocks of epilogue)
    if (cost < best - 1e-9) { best = cost; splits = sidx; }
  }
  { const char* e = getenv("SC_WGRAD_SPLITS"); if (e) { splits = atoi(e); if (splits > kb / 8) splits = kb / 8; if (splits < 1) splits = 1; } }
  if (!accumulate)
    zero2d_kernel<<<(unsigned)min((int64_t)2048, cdiv(N * K, 256)), 256, 0, st>>>(dW, lddw, N, K);
  return launch_tc<true, true, 2>(dY, lddy, A, lda, dW, lddw, nullptr, N, K, M, (int)splits, st);
}

}  // namespace sc
