// K3: CTC loss and gradient, fused with log-softmax and its backward (sm_100a).
//
// Replaces model.py:70-71 (`enc_out.log_softmax(-1).transpose(0,1)` + nn.CTCLoss(blank=0,
// zero_infinity=True), train.py:142), i.e. ATen's log_softmax / ctc_loss_gpu /
// ctc_loss_backward_gpu chain.  Semantics: SURVEY.md Appendix B.
//
// Passes:
//   1. emissions (ctc_lse_gather_lin / ctc_lse_gather): one warp per frame (b,t<T_b): 128-bit coalesced sweep over
//      the V logits -> lse[b,t]; the frame's lattice emissions are gathered into a compact row while the row is
//      still hot in L1, so the serial recursions never touch the V-wide tensor.
//   2. lattice: for transcripts of up to 255 labels the alpha/beta recursions run in the LINEAR domain on fp64,
//      one warp per (utterance, direction) plus an I/O warp (sc_ctc_lin64.cuh; range checks flag what fp64 cannot
//      hold); flagged utterances and wider lattices take the log-domain kernels below (ctc_alpha_beta_wave2: pair
//      per thread, wavefront hand-off between warps, no block barrier per step; ctc_alpha_beta: one node per thread,
//      block barrier per step, any width).
//   3. gradient (ctc_grad): one warp per frame: occupancy = normalised alpha*beta over the lattice (pairs kept in
//      registers between the max, the sum and the scatter); dlogits = gout*scale_b*(softmax - occupancy), label
//      occupancies scattered with shared-memory atomics; exact zeros for t>=T_b and for infeasible utterances.
//   sc_ctc_head (opt-in) runs 1 and 3 on side streams under 2 cut into frame ranges.
// Tried and rejected (r01; (a) was overtaken in r02 by doing it in fp64): (a) a scaled linear-domain recursion, one warp per lattice with
// the column in registers — 6x shorter serial chain, but with T >> U the forward and backward
// masses sit at opposite ends of the lattice and their overlap (the occupancy) lies 2^-125
// and further below either column's maximum, outside fp32's range (measured on random
// logits, T=300, U=30); (b) log-domain with 2 or 4 nodes per thread in registers and shuffle
// exchange — 0.77 ms against 0.53 ms for one node per thread: fewer warps leave the
// MUFU/FMNMX latency chain of each step exposed; (c) two timesteps per block barrier with a
// 3-2-1 trapezoid of recomputed neighbour nodes — no change (0.54 ms): the two dependent
// log-sum-exps, not the barrier, are the chain.
// (d) pass 1 with the logits rows streamed through a shared-memory ring by bulk async copies
// (8 consumer warps + a producer warp per CTA, 96 KB in flight per CTA, labels prefetched while the
// row is awaited, emissions kept in registers): 0.192 ms against 0.185 ms for the register-staged
// kernel at 40 warps per SM — bytes in flight are not what holds pass 1 at ~2.9 TB/s; not kept.
// Algorithmic HBM bytes per frame: 3*V*e + 8*(2U+1)  (SURVEY.md 8d).
#include "sc_common.cuh"
#include "sc_tma.cuh"
#include <stdlib.h>
#include <vector>

namespace sc {

#define NEG_INF (-INFINITY)

__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == NEG_INF) return NEG_INF;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}
__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == NEG_INF) return NEG_INF;
  return m + __logf(__expf(a - m) + __expf(b - m));
}

__device__ __forceinline__ int64_t ext_label(const int64_t* tg, int s, int64_t blank) {
  return (s & 1) ? tg[s >> 1] : blank;
}

constexpr int CTC_WARPS = 8;
constexpr int CTC_RENORM = 16;
constexpr int CTC_EB = 16;       // rows of emissions per bulk-copied shared-memory block   // alpha/beta columns are re-centred every this many steps

// ---- pass 1 ------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(CTC_WARPS * 32, 5)
ctc_lse_gather_kernel(const T* __restrict__ logits, int64_t stride_b, int64_t stride_t,
                      const int64_t* __restrict__ targets, int64_t ldt,
                      const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                      int B, int Tn, int V, int Umax, int Smax, int64_t blank,
                      float* __restrict__ lse, float* __restrict__ lplat, float* __restrict__ cshift) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned nrows = (unsigned)B * (unsigned)Tn;           // B*T < 2^31 (checked by the host)
  // (grid-stride loop; launched with one warp per row — capping the grid at a few blocks per SM
  // measured slower: 0.148 -> 0.203 ms)
  for (unsigned row = blockIdx.x * CTC_WARPS + warp; row < nrows; row += gridDim.x * CTC_WARPS) {
    const int b = (int)(row / (unsigned)Tn), t = (int)(row - (unsigned)b * (unsigned)Tn);
    int64_t Tb = in_lens[b]; if (Tb > Tn) Tb = Tn;
    if (t >= Tb) continue;
    const int64_t U64 = tgt_lens[b];
    if (U64 < 0 || U64 > Umax) continue;                         // invalid length (torch raises on the host; lengths here may be device tensors): the lattice pass reports the utterance infeasible
    const T* x = logits + b * stride_b + t * stride_t;
    const float l = warp_row_lse<T>(x, V, lane);
    if (lane == 0) lse[row] = l;
    const int S = 2 * (int)U64 + 1;
    const int64_t* tg = targets + (int64_t)b * ldt;
    float* out = lplat + (int64_t)row * Smax;
    // a label outside [0, V) is probability zero (-inf emission): the utterance comes out infeasible
    auto emis = [&](int64_t lab) -> float { return (lab >= 0 && lab < V) ? (ld_f(x + lab) - l) * 1.4426950408889634f : NEG_INF; };
    // emissions in log2 units (the recursion runs on ex2/lg2 directly), shifted so that the
    // frame's largest lattice emission is 0: the recursion's values then drift by the gap between
    // the paths and the frame-wise best node (~2 per frame on random logits, ~0 on a trained
    // model) instead of by log2(V) per frame, which lets it re-centre 4x less often at the same
    // fp32 resolution.  The shift goes to cshift[b,t] and is added back into the likelihood.
    float c = NEG_INF;
    constexpr int NL = 10;                                       // lattice nodes per lane kept in registers (S <= 320)
    if (S <= 32 * NL) {
      // straight-line: the lane's emissions stay in registers between the max and the single store
      // (the two-pass loop below cost more instructions per frame than the V-wide log-sum-exp itself)
      float e[NL];
#pragma unroll
      for (int k = 0; k < NL; ++k) {
        const int s = lane + 32 * k;
        const int64_t lab = (s < S && (s & 1)) ? tg[s >> 1] : blank;
        e[k] = emis(lab);
        if (s < S) c = fmaxf(c, e[k]);
      }
      c = warp_max(c);
      if (!(c > -1e29f)) c = 0.f;                                // every lattice emission is -inf: leave the row alone
#pragma unroll
      for (int k = 0; k < NL; ++k)
        if (lane + 32 * k < S) out[lane + 32 * k] = fmaxf(e[k] - c, -1e30f);   // -inf logits become the recursion's finite dead value
    } else {
      for (int s = lane; s < S; s += 32) {
        const float e = emis(ext_label(tg, s, blank));
        out[s] = e;
        c = fmaxf(c, e);
      }
      c = warp_max(c);
      if (!(c > -1e29f)) c = 0.f;
      for (int s = lane; s < S; s += 32) out[s] = fmaxf(out[s] - c, -1e30f);   // each lane re-reads its own stores
    }
    if (lane == 0) cshift[row] = c;
  }
}

// ---- pass 2 ------------------------------------------------------------------------
// dir 0 = alpha (forward in t), dir 1 = beta (backward in t).  Both carry the emission of
// their own timestep through the recursion, as in the oracle (ctc_oracle.py) and ATen; alpha is
// written with it, beta WITHOUT it, so alpha+beta is the log-occupancy up to a per-frame constant
// and the gradient pass needs no emission gather.  Everything here is in
// LOG2 units (lplat is written pre-scaled by log2(e)) so the serial chain is
// LDS -> max -> ex2 -> add -> lg2 -> add -> STS -> barrier with no multiplies and no branches:
// the first version spent ~1300 cycles per timestep in ~90 dependent SASS instructions
// (profiles/r01_ncu_ctc_alpha_beta_hotlines_before.txt).
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// log2(2^a + 2^b + 2^c); all -inf in -> -inf out, without a branch
__device__ __forceinline__ float lse3_2(float a, float b, float c) {
  const float m = fmaxf(fmaxf(a, fmaxf(b, c)), -1e30f);
  return m + lg2f(ex2f(a - m) + ex2f(b - m) + ex2f(c - m));
}
constexpr float LOG2E = 1.4426950408889634f;
constexpr float CTC_DEAD = -1e30f;          // below CTC_DEAD_TEST a log2-domain value counts as probability zero
constexpr float CTC_DEAD_TEST = -1e29f;
constexpr int CTC_GRAD_ALL = 0, CTC_GRAD_SPEC = 1, CTC_GRAD_FIX = 2;   // modes of the gradient pass (sc_ctc_head)
constexpr float LN2 = 0.6931471805599453f;

// sum_t cshift[t], t < n, over the block in double; result valid in every thread.  Contains two
// block barriers: call it from uniform code.
__device__ __forceinline__ double block_shift_sum(const float* __restrict__ c, int n, double* dred) {
  double acc = 0.0;
  for (int t = threadIdx.x; t < n; t += blockDim.x) acc += (double)c[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = acc;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) r += dred[w];
  __syncthreads();
  return r;
}

__global__ void ctc_alpha_beta_kernel(const float* __restrict__ lplat, const float* __restrict__ cshift,
                                      const int64_t* __restrict__ targets, int64_t ldt,
                                      const int64_t* __restrict__ in_lens,
                                      const int64_t* __restrict__ tgt_lens,
                                      int Tn, int Umax, int Smax, int64_t blank,
                                      float* __restrict__ alpha, float* __restrict__ beta,
                                      float* __restrict__ nll) {
  extern __shared__ __align__(128) float sm[];   // 2 lines of (Smax + 4) floats (2 pad cells either side) + emission blocks
  __shared__ float red[32];
  __shared__ double dred[32];
  __shared__ __align__(8) uint64_t ebar[2];
  const int b = blockIdx.x, dir = blockIdx.y;
  int64_t Tb64 = in_lens[b]; if (Tb64 > Tn) Tb64 = Tn;
  const int Tb = (int)Tb64;
  const int64_t U64 = tgt_lens[b];
  const bool bad_len = U64 < 0 || U64 > Umax;
  const int U = bad_len ? 0 : (int)U64;
  const int S = 2 * U + 1;
  const int64_t* tg = targets + (int64_t)b * ldt;
  const int LINE = Smax + 4;
  if (Tb <= 0 || bad_len) {
    if (dir == 0 && threadIdx.x == 0) nll[b] = (U == 0 && !bad_len) ? 0.f : INFINITY;
    return;
  }
  const double shift_sum = dir == 0 ? block_shift_sum(cshift + (int64_t)b * Tn, Tb, dred) : 0.0;
  for (int i = threadIdx.x; i < 2 * LINE; i += blockDim.x) sm[i] = NEG_INF;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&ebar[0]), 1);
    mbar_init(smem_u32(&ebar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float* bufA = sm + 2;
  float* bufB = sm + LINE + 2;
  const float* lp_b = lplat + (int64_t)b * Tn * Smax;
  float* out_b = (dir == 0 ? alpha : beta) + (int64_t)b * Tn * Smax;
  const int t_first = dir == 0 ? 0 : Tb - 1;
  const int step = dir == 0 ? 1 : -1;
  const int nb = dir == 0 ? -1 : 1;   // neighbour direction: alpha looks at s-1,s-2; beta at s+1,s+2
  // skip transition (two nodes away) is a per-node constant: allowed iff l'_s is a label that
  // differs from the label two nodes away
  auto skip_ok = [&](int s) -> bool {
    if (!(s & 1)) return false;
    if (dir == 0) return s >= 2 && tg[s >> 1] != tg[(s >> 1) - 1];
    return s + 2 < S && tg[s >> 1] != tg[(s >> 1) + 1];
  };
  // Log-space values drift to magnitude ~T*log2(V); every CTC_RENORM steps the column is
  // re-centred on its maximum so fp32 keeps ~1e-6 absolute resolution for any T.  The removed
  // offsets only matter for the likelihood (summed in double); the gradient pass normalises
  // each frame's occupancies itself (sum_s 2^(alpha+beta-lp) is the same for every t).
  double csum = 0.0;
  auto block_max = [&](float m) -> float {
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    float r = NEG_INF;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) r = fmaxf(r, red[w]);
    return r;
  };
  const float* final_buf;
  if (S <= (int)blockDim.x) {
    // ---- fast path: one node per thread; emissions staged by bulk async copies ----
    // The emission stream lplat[b, t, :] is independent of the recursion.  Register prefetch
    // (LDG 8 steps ahead) stalls the chain anyway: the 6 scoreboard slots of a warp are
    // shared by LDS/MUFU/LDG, so a consumer ends up waiting on a much younger load (measured:
    // 1.41 ms with, 0.60 ms without the loads).  Instead CTC_EB consecutive rows (one
    // contiguous block of lplat) are landed in shared memory by cp.async.bulk on an mbarrier,
    // double-buffered, and the recursion only ever issues LDS.
    const int s = threadIdx.x;
    const bool has = s < S;
    const bool skip = has && skip_ok(s);
    float* ebuf = sm + 2 * LINE;                               // 2 buffers of CTC_EB rows x Smax
    const int nvis = (Tb + CTC_EB - 1) / CTC_EB;               // blocks of rows, visited in scan order
    auto blk_of = [&](int vi) { return dir == 0 ? vi : nvis - 1 - vi; };
    auto rows_of = [&](int blk) { const int r = Tb - blk * CTC_EB; return r < CTC_EB ? r : CTC_EB; };
    auto issue = [&](int vi) {
      const int blk = blk_of(vi);
      const uint32_t bytes = (uint32_t)rows_of(blk) * (uint32_t)Smax * 4u;
      const uint32_t bar = smem_u32(&ebar[vi & 1]);
      mbar_expect_tx(bar, bytes);
      bulk_load_1d(smem_u32(ebuf + (size_t)(vi & 1) * CTC_EB * Smax), lp_b + (int64_t)blk * CTC_EB * Smax, bytes, bar);
    };
    if (threadIdx.x == 0) {
      issue(0);
      if (nvis > 1) issue(1);
    }
    const int sidx = has ? s : 0;
    const int64_t stride = (int64_t)step * Smax;
    float* op = out_b + (int64_t)t_first * Smax + sidx;
    float* prev = bufA;
    float* cur = bufB;
    // Visit-structured loop: everything that happens once per 16-row emission block (barrier wait,
    // re-centring, recycling the buffer) sits outside the per-timestep path, which is then just
    // LDS x4 -> lse3 -> STS/STG -> barrier.
    for (int vi = 0; vi < nvis; ++vi) {
      mbar_wait(smem_u32(&ebar[vi & 1]), (uint32_t)((vi >> 1) & 1));
      const int rows = rows_of(blk_of(vi));
      // emission pointer of this thread's node for the first row visited, and its per-step stride
      const float* ep = ebuf + ((size_t)(vi & 1) * CTC_EB + (dir == 0 ? 0 : rows - 1)) * Smax + sidx;
      const int estride = dir == 0 ? Smax : -Smax;
      int pos = 0;
      if (vi == 0) {                                             // init column (step 0)
        float v = NEG_INF, pre = NEG_INF;
        if (has && (dir == 0 ? s < 2 : s >= S - 2)) { v = *ep; pre = 0.f; }
        if (has) { prev[s] = v; *op = dir == 0 ? v : pre; }
        __syncthreads();
        ep += estride;
        pos = 1;
      } else {
        // re-centre the last column once per visit (at most every CTC_EB steps)
        const float m = block_max(has ? prev[s] : NEG_INF);
        if (m > CTC_DEAD_TEST) {
          if (has) prev[s] -= m;                                  // rows already written keep their own offset
          csum += (double)m;
        }
        __syncthreads();
      }
      for (; pos < rows; ++pos) {
        const float e = *ep;
        ep += estride;
        op += stride;
        float pre = NEG_INF;
        if (has) pre = lse3_2(prev[s], prev[s + nb], skip ? prev[s + 2 * nb] : NEG_INF);
        const float v = pre + e;
        if (has) { cur[s] = v; *op = dir == 0 ? v : pre; }
        __syncthreads();
        float* tmp = prev; prev = cur; cur = tmp;
      }
      // every thread is past the barrier of this visit's last row: its buffer can be refilled
      if (threadIdx.x == 0 && vi + 2 < nvis) issue(vi + 2);
    }
    final_buf = prev;
  } else {
    // ---- general path: several nodes per thread ----
    // init column
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
      float v = NEG_INF, pre = NEG_INF;
      if (dir == 0 ? s < 2 : s >= S - 2) { v = lp_b[(int64_t)t_first * Smax + s]; pre = 0.f; }
      bufA[s] = v;
      out_b[(int64_t)t_first * Smax + s] = dir == 0 ? v : pre;
    }
    __syncthreads();
    float* prev = bufA;
    float* cur = bufB;
    for (int i = 1; i < Tb; ++i) {
      const int t = t_first + i * step;
      for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const float pre = lse3_2(prev[s], prev[s + nb], skip_ok(s) ? prev[s + 2 * nb] : NEG_INF);
        cur[s] = pre + lp_b[(int64_t)t * Smax + s];
        if (dir == 1) out_b[(int64_t)t * Smax + s] = pre;       // beta leaves without its frame's emission
      }
      if ((i % CTC_RENORM) == 0) {
        float m = NEG_INF;
        for (int s = threadIdx.x; s < S; s += blockDim.x) m = fmaxf(m, cur[s]);
        m = block_max(m);
        if (m > CTC_DEAD_TEST) {
          for (int s = threadIdx.x; s < S; s += blockDim.x) cur[s] -= m;
          csum += (double)m;
        }
      }
      if (dir == 0) for (int s = threadIdx.x; s < S; s += blockDim.x) out_b[(int64_t)t * Smax + s] = cur[s];
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    }
    final_buf = prev;
  }
  if (dir == 0 && threadIdx.x == 0) {
    const float ll2 = lse3_2(final_buf[S - 1], S > 1 ? final_buf[S - 2] : NEG_INF, NEG_INF);
    nll[b] = (ll2 < CTC_DEAD_TEST) ? INFINITY : (float)(-(csum + shift_sum + (double)ll2) * (double)LN2);   // +inf when infeasible
  }
}

}  // namespace sc
#include "sc_ctc_lin64.cuh"
namespace sc {

// ---- pass 2, wavefront variant (lattices up to 1024 nodes) ---------------------------------
// Same recursion, no block barrier per timestep.  Node values live in REGISTERS (thread i owns
// node i for alpha, node S-1-i for beta, so both directions only ever look at lower threads);
// the two neighbours come from __shfl_up, and only lanes 0/1 of a warp need anything from
// another warp: lanes 30/31 of warp w-1 publish {value, step tag} pairs into a shared-memory
// slot per timestep and lanes 0/1 of warp w poll that slot for the tag of the previous step.
// Dependencies only run from warp w-1 to warp w, so the warps settle into a skew of one
// shared-memory round trip and after that nobody waits: the per-step cost is one warp's own
// chain (SHFL -> min/max -> 2x ex2 -> lg2) instead of the slowest of 10 warps plus a barrier
// (measured: see DESIGN.md 3.3).  The block still meets once per emission block of `EB` rows
// (32-64 timesteps): there the column is re-centred on its maximum, the slots are recycled and
// the next emission block is requested.  log2(2^a+2^b+2^c) is taken as
// max + lg2(1 + 2^(mid-max) + 2^(min-max)): 3 MUFU ops instead of 4 — with 3 warps per
// scheduler the MUFU pipe (8 issue cycles per warp instruction) is what bounds a step.
// Dead nodes are the FINITE sentinel CTC_DEAD here (not -inf), so max - max never produces a NaN
// and the log-sum-exp needs no clamp: 4 min/max + 2 ex2 + 1 lg2.
__device__ __forceinline__ float lse3w(float a, float b, float c) {
  const float hi = fmaxf(a, b), lo = fminf(a, b);
  const float m = fmaxf(hi, c), mid = fminf(hi, c);
  return m + lg2f(1.f + ex2f(mid - m) + ex2f(lo - m));
}
__device__ __forceinline__ void slot_publish(uint32_t addr, float v, int tag) {
  asm volatile("st.volatile.shared.v2.b32 [%0], {%1, %2};" :: "r"(addr), "r"(__float_as_uint(v)), "r"(tag));
}
// value published by lane 31 of the producing warp, once it carries `tag`.  Straight-line when
// the slot is already there (the steady state); bounded spin, then trap, when it is not.
__device__ __forceinline__ float slot_poll(uint32_t addr, int tag) {
  uint32_t v;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 t, n;\n"
      "ld.volatile.shared.v2.b32 {%0, t}, [%1];\n"
      "setp.eq.s32 p, t, %2;\n"
      "@p bra.uni SC_SLOT_DONE;\n"
      "mov.b32 n, 0;\n"
      "SC_SLOT_SPIN:\n"
      "ld.volatile.shared.v2.b32 {%0, t}, [%1];\n"
      "setp.eq.s32 p, t, %2;\n"
      "@p bra.uni SC_SLOT_DONE;\n"
      "add.s32 n, n, 1;\n"
      "setp.lt.s32 p, n, 0x2000000;\n"
      "@p bra.uni SC_SLOT_SPIN;\n"
      "trap;\n"
      "SC_SLOT_DONE:\n"
      "}\n"
      : "=r"(v) : "r"(addr), "r"(tag));
  return __uint_as_float(v);
}

// ---- pass 2, wavefront variant with a (blank, label) PAIR per thread --------------------------
// Thread i owns blank node 2i and label node 2i+1 (alpha; mirrored node numbering for beta).
// The label looks at its own blank (same thread), so ONE shuffle per step — the previous thread's
// label — feeds both nodes; the two log-sum-exps are independent and interleave in the pipeline,
// and a 301-node lattice is 5 warps instead of 10: half the instructions per node and at most two
// warps per scheduler.
__device__ __forceinline__ float lse2w(float a, float b) {
  const float m = fmaxf(a, b), lo = fminf(a, b);
  return m + lg2f(1.f + ex2f(lo - m));
}

// FMT 0: emission rows of 2U+1 log2 values (pitch = Smax); FMT 1: the linear-domain words of sc_ctc_lin64.cuh
// (pitch LP; [0] blank, [1+u] label u) — in that format the kernel only recomputes the utterances the fp64
// recursion flagged (lossy[b] != 0) and leaves at once for the others.
template <int dir, int FMT>
__device__ __forceinline__ void
ctc_wave2_body(float* __restrict__ sm, const float* __restrict__ lplat, const float* __restrict__ cshift,
               const int64_t* __restrict__ targets, int64_t ldt,
               const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
               int Tn, int Umax, int Smax, int LP, int EB, const int* __restrict__ lossy,
               float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ nll) {
  __shared__ float red[2][32];
  __shared__ double dred[32];
  __shared__ float fin[2];
  __shared__ __align__(8) uint64_t ebar[2];
  const int b = blockIdx.x;
  int64_t Tb64 = in_lens[b]; if (Tb64 > Tn) Tb64 = Tn;
  const int Tb = (int)Tb64;
  if (FMT == 1 && !lossy[b]) return;                             // the fp64 recursion's result stands
  const int64_t U64 = tgt_lens[b];
  const bool bad_len = U64 < 0 || U64 > Umax;
  const int U = bad_len ? 0 : (int)U64;
  const int64_t* tg = targets + (int64_t)b * ldt;
  if (Tb <= 0 || bad_len) {
    if (dir == 0 && threadIdx.x == 0) nll[b] = (U == 0 && !bad_len) ? 0.f : INFINITY;
    return;
  }
  const int i = threadIdx.x, warp = i >> 5, lane = i & 31;
  const int nwarps = (int)(blockDim.x >> 5);
  const double shift_sum = dir == 0 ? block_shift_sum(cshift + (int64_t)b * Tn, Tb, dred) : 0.0;
  float* ebuf = sm;
  const int EP = FMT == 1 ? lin_row_pitch(U) : Smax;             // pitch of an emission row (FMT 1: this utterance's, inside a region of LP words per frame)
  int* slots = reinterpret_cast<int*>(sm + 2 * (size_t)EB * EP);     // [nwarps][EB+1] x {value, tag}
  for (int k = i; k < nwarps * (EB + 1) * 2; k += blockDim.x) slots[k] = -1;
  if (i == 0) {
    fin[0] = CTC_DEAD; fin[1] = CTC_DEAD;
    mbar_init(smem_u32(&ebar[0]), 1);
    mbar_init(smem_u32(&ebar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const bool hasb = i <= U, hasl = i < U;
  // memory positions of this thread's blank and label node
  const int sb = hasb ? (dir == 0 ? 2 * i : 2 * U - 2 * i) : 0;
  const int sl = hasl ? (dir == 0 ? 2 * i + 1 : 2 * U - 2 * i - 1) : 0;
  bool skip = false;                                             // label may also come from the previous thread's label
  if (hasl) {
    if (dir == 0) skip = i >= 1 && tg[i] != tg[i - 1];
    else { const int u = U - 1 - i; skip = u + 1 < U && tg[u] != tg[u + 1]; }
  }
  const bool producer = lane == 31 && warp + 1 < nwarps;
  const bool lane0 = lane == 0;
  const uint32_t my_slot = smem_u32(slots) + (uint32_t)(warp * (EB + 1) * 8);
  const uint32_t in_slot = smem_u32(slots) + (uint32_t)((warp > 0 ? warp - 1 : 0) * (EB + 1) * 8);
  // where the pair's emissions sit in a row, and how a stored word becomes a log2 value
  const int eb_pos = FMT == 1 ? 0 : sb;
  const int el_pos = FMT == 1 ? (hasl ? (dir == 0 ? 1 + i : U - i) : 0) : sl;
  auto ld_e = [](const float* q) -> float { return FMT == 1 ? lin_word_to_log2(__float_as_uint(*q)) : *q; };
  const float* lp_b = lplat + (int64_t)b * Tn * (FMT == 1 ? LP : Smax);
  float* out_b = (dir == 0 ? alpha : beta) + (int64_t)b * Tn * Smax;
  const int nvis = (Tb + EB - 1) / EB;
  auto blk_of = [&](int vi) { return dir == 0 ? vi : nvis - 1 - vi; };
  auto rows_of = [&](int blk) { const int r = Tb - blk * EB; return r < EB ? r : EB; };
  auto issue = [&](int vi) {
    const int blk = blk_of(vi);
    const uint32_t bytes = (uint32_t)rows_of(blk) * (uint32_t)EP * 4u;
    const uint32_t bar = smem_u32(&ebar[vi & 1]);
    mbar_expect_tx(bar, bytes);
    bulk_load_1d(smem_u32(ebuf + (size_t)(vi & 1) * EB * EP), lp_b + (int64_t)blk * EB * EP, bytes, bar);
  };
  if (i == 0) {
    issue(0);
    if (nvis > 1) issue(1);
  }
  const int t_first = dir == 0 ? 0 : Tb - 1;
  const int64_t stride = dir == 0 ? (int64_t)Smax : -(int64_t)Smax;
  const int estride = dir == 0 ? EP : -EP;
  float* opb = out_b + (int64_t)t_first * Smax + sb;
  float* opl = out_b + (int64_t)t_first * Smax + sl;
  double csum = 0.0;
  float pb = CTC_DEAD, pl = CTC_DEAD;                            // the pair's values after the last step
  int g = 0;
  for (int vi = 0; vi < nvis; ++vi) {
    int pos = 0;
    if (vi > 0) {
      float m = warp_max(hasb ? fmaxf(pb, pl) : CTC_DEAD);
      if (lane == 0) red[vi & 1][warp] = m;
      __syncthreads();
      m = CTC_DEAD;
      for (int w = 0; w < nwarps; ++w) m = fmaxf(m, red[vi & 1][w]);
      if (m > CTC_DEAD_TEST) { pb = fmaxf(pb - m, CTC_DEAD); pl = fmaxf(pl - m, CTC_DEAD); csum += (double)m; }
      if (i == 0 && vi + 1 < nvis) issue(vi + 1);
      if (producer) slot_publish(my_slot, pl, g - 1);
    }
    mbar_wait(smem_u32(&ebar[vi & 1]), (uint32_t)((vi >> 1) & 1));
    const int rows = rows_of(blk_of(vi));
    const float* erow = ebuf + ((size_t)(vi & 1) * EB + (dir == 0 ? 0 : rows - 1)) * EP;
    const float* epb = erow + eb_pos;
    const float* epl = erow + el_pos;
    if (vi == 0) {                                               // first column: nodes 0 and 1 of the scan order
      if (i == 0) { pb = ld_e(epb); pl = hasl ? ld_e(epl) : CTC_DEAD; }
      if (hasb) *opb = dir == 0 ? pb : (i == 0 ? 0.f : CTC_DEAD);
      if (hasl) *opl = dir == 0 ? pl : (i == 0 ? 0.f : CTC_DEAD);
      if (producer) slot_publish(my_slot + 8, pl, 0);
      epb += estride; epl += estride;
      pos = 1;
      g = 1;
    }
    uint32_t rd = in_slot + (uint32_t)(pos * 8);
    uint32_t wr = my_slot + (uint32_t)((pos + 1) * 8);
#pragma unroll 2
    for (; pos < rows; ++pos) {
      const float eb = ld_e(epb);
      const float el = hasl ? ld_e(epl) : CTC_DEAD;              // a missing label node stays dead
      epb += estride; epl += estride;
      opb += stride; opl += stride;
      float xl = __shfl_up_sync(0xffffffffu, pl, 1);
      float x = CTC_DEAD;
      if (warp > 0) x = slot_poll(rd, g - 1);
      xl = lane0 ? x : xl;
      const float preb = lse2w(pb, xl);
      const float prel = lse3w(pl, pb, skip ? xl : CTC_DEAD);
      pb = preb + eb;
      pl = prel + el;
      if (hasb) *opb = dir == 0 ? pb : preb;                     // beta leaves without its frame's emission
      if (hasl) *opl = dir == 0 ? pl : prel;
      if (producer) slot_publish(wr, pl, g);
      ++g; rd += 8; wr += 8;
    }
  }
  if (dir == 0) {
    if (i == U) fin[0] = pb;
    if (i == U - 1) fin[1] = pl;
    __syncthreads();
    if (i == 0) {
      const float ll2 = lse2w(fin[0], fin[1]);
      nll[b] = (ll2 < CTC_DEAD_TEST) ? INFINITY : (float)(-(csum + shift_sum + (double)ll2) * (double)LN2);
    }
  }
}

template <int FMT>
__global__ void __launch_bounds__(512, 1)
ctc_alpha_beta_wave2_kernel(const float* __restrict__ lplat, const float* __restrict__ cshift,
                            const int64_t* __restrict__ targets, int64_t ldt,
                            const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                            int Tn, int Umax, int Smax, int LP, int EB, const int* __restrict__ lossy,
                            float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ nll) {
  extern __shared__ __align__(128) float sm[];   // 2 emission blocks of EB rows, then nwarps x (EB+1) 8-byte slots
  if (blockIdx.y == 0) ctc_wave2_body<0, FMT>(sm, lplat, cshift, targets, ldt, in_lens, tgt_lens, Tn, Umax, Smax, LP, EB, lossy, alpha, beta, nll);
  else ctc_wave2_body<1, FMT>(sm, lplat, cshift, targets, ldt, in_lens, tgt_lens, Tn, Umax, Smax, LP, EB, lossy, alpha, beta, nll);
}

// loss = reduction over utterances with zero_infinity
__global__ void ctc_reduce_kernel(const float* __restrict__ nll, const int64_t* __restrict__ tgt_lens,
                                  int B, int reduction, float* __restrict__ loss) {
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += 32) {
    float v = nll[b];
    if (!isfinite(v)) v = 0.f;
    if (reduction == 1) {
      const float u = (float)tgt_lens[b];
      v = v / fmaxf(u, 1.f);
    }
    acc += v;
  }
  acc = warp_sum(acc);
  if (threadIdx.x == 0) *loss = (reduction == 1) ? acc / (float)B : acc;
}

// ---- pass 3 ------------------------------------------------------------------------
// Lattice nodes are taken as (blank 2u, label 2u+1) pairs, pair u on lane u%32 (one float2 load
// from alpha and one from beta per pair).  NP > 0: lattices up to 64*NP nodes, each lane keeps
// its NP pairs' log-occupancies in registers between the max, the normalising sum and the
// scatter.  NP == 0: any width, recomputed from (L1-hot) reloads in each sweep.
template <typename TI, typename TO, int NP>
__device__ __forceinline__ void
ctc_grad_row(float* __restrict__ sm, const unsigned row, const int warp, const int lane,
             const TI* __restrict__ logits, int64_t stride_b, int64_t stride_t,
             const int64_t* __restrict__ targets, int64_t ldt,
             const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
             int B, int Tn, int V, int Smax, int64_t blank,
             const float* __restrict__ lse, const float* __restrict__ alpha,
             const float* __restrict__ beta, const float* __restrict__ nll,
             const float* __restrict__ grad_out, int reduction,
             TO* __restrict__ dlogits, int64_t dstride_b, int64_t dstride_t, const int* __restrict__ lossy, int mode) {
  // mode CTC_GRAD_SPEC: the likelihood is not known yet (the recursions are still under way in other frames): the
  // length is validated here and the row is formed as if the utterance were feasible; CTC_GRAD_FIX puts that right
  const int b = (int)(row / (unsigned)Tn), t = (int)(row - (unsigned)b * (unsigned)Tn);
  int64_t Tb = in_lens[b]; if (Tb > Tn) Tb = Tn;
  TO* dx = dlogits + b * dstride_b + t * dstride_t;
  bool dead;
  if (mode == CTC_GRAD_SPEC) { const int64_t U64 = tgt_lens[b]; dead = U64 < 0 || 2 * U64 + 1 > Smax; }
  else dead = !isfinite(nll[b]);
  constexpr int VWI = 16 / sizeof(TI), VWO = 16 / sizeof(TO);
  const bool vec_out = (V % VWO == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0);
  if (t >= Tb || dead) {                              // exact zeros (App. B)
    if (vec_out) {
      for (int i = lane * VWO; i < V; i += 32 * VWO) *reinterpret_cast<uint4*>(dx + i) = make_uint4(0, 0, 0, 0);
    } else {
      for (int i = lane; i < V; i += 32) st_f(dx + i, 0.f);
    }
    return;
  }
  float* r = sm + (int64_t)warp * V;
  const TI* x = logits + b * stride_b + t * stride_t;
  const int U = (int)tgt_lens[b];                     // valid here: an invalid length left nll = inf above / was tested above
  const int64_t* tg = targets + (int64_t)b * ldt;
  const bool lin = lossy != nullptr && !lossy[b];     // rows in the linear-domain format (lossy == nullptr: log-domain launch)
  // pair u <= U exists; its label node only for u < U.  occupancy_s = 2^(alpha+beta)_s / sum_s'
  // (beta carries no emission; per-frame offsets from re-centring cancel in the normalisation)
  const float2* al2 = reinterpret_cast<const float2*>(alpha + (int64_t)row * Smax);
  const float2* be2 = reinterpret_cast<const float2*>(beta + (int64_t)row * Smax);
  const float l2 = lse[row] * LOG2E;
  // softmax row -> shared memory
  if ((V % VWI == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0)) {
    for (int i = lane * VWI; i < V; i += 32 * VWI) {
      float f[VWI];
      Vec<TI, VWI> raw; raw.raw = __ldg(reinterpret_cast<const uint4*>(x + i));
      unpack(raw, f);
#pragma unroll
      for (int j = 0; j < VWI; ++j) f[j] = ex2f(fmaf(f[j], LOG2E, -l2));
#pragma unroll
      for (int j = 0; j < VWI; j += 4) *reinterpret_cast<float4*>(r + i + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
  } else {
    for (int i = lane; i < V; i += 32) r[i] = ex2f(fmaf(ld_f(x + i), LOG2E, -l2));
  }
  float z = 0.f, bsum = 0.f, vmax = CTC_DEAD;
  if (NP > 0) {
    float wb[NP > 0 ? NP : 1], wl[NP > 0 ? NP : 1];
    bool live;
    if (lin) {
      // linear-domain rows (sc_ctc_lin64.cuh): each node is the high word of an fp64 value, alpha at its node
      // index, beta at the mirrored one.  occupancy ~ alpha*beta: exponent fields add, 20-bit mantissas multiply;
      // the frame's largest exponent sum is the common scale (per-column scale factors cancel with it).
      const uint2* aw = reinterpret_cast<const uint2*>(al2);
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(be2);
      int eb_[NP > 0 ? NP : 1], el_[NP > 0 ? NP : 1];
      int emax = -1;
      auto mant = [](uint32_t w) -> float { return __uint_as_float(0x3f800000u | ((w & 0xfffffu) << 3)); };
      // every word of the lane's pairs is requested before any is looked at (indices clamped into the row, the words
      // of absent nodes zeroed afterwards): written pair by pair with the loads inside `if (u <= U)` the compiler
      // cannot move them over the branches and the frame pays one global-memory round trip per pair — a third of
      // this kernel's stall samples sat on the first compare of a loaded word (r02, ncu source view)
      uint2 aw_[NP > 0 ? NP : 1];
      uint32_t cb_[NP > 0 ? NP : 1], cl_[NP > 0 ? NP : 1];
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        const int u = lane + 32 * kk;
        const int uc = u <= U ? u : U;                           // pair U always exists
        aw_[kk] = __ldg(aw + uc);
        cb_[kk] = __ldg(bw + 2 * (U - uc));                      // beta row mirrored: node s at 2U - s
        cl_[kk] = __ldg(bw + (uc < U ? 2 * (U - uc) - 1 : 0));
      }
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        const int u = lane + 32 * kk;
        const bool ob = u <= U && aw_[kk].x != 0u && cb_[kk] != 0u;
        const bool ol = u < U && aw_[kk].y != 0u && cl_[kk] != 0u;
        eb_[kk] = ob ? (int)(aw_[kk].x >> 20) + (int)(cb_[kk] >> 20) : -1;
        el_[kk] = ol ? (int)(aw_[kk].y >> 20) + (int)(cl_[kk] >> 20) : -1;
        wb[kk] = ob ? mant(aw_[kk].x) * mant(cb_[kk]) : 0.f;
        wl[kk] = ol ? mant(aw_[kk].y) * mant(cl_[kk]) : 0.f;
        emax = max(emax, max(eb_[kk], el_[kk]));
      }
      emax = __reduce_max_sync(0xffffffffu, emax);
      auto pow2 = [](int d) -> float { return d < -126 ? 0.f : __uint_as_float((uint32_t)(d + 127) << 23); };   // d <= 0
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        wb[kk] = eb_[kk] >= 0 ? wb[kk] * pow2(eb_[kk] - emax) : 0.f;
        wl[kk] = el_[kk] >= 0 ? wl[kk] * pow2(el_[kk] - emax) : 0.f;
        z += wb[kk] + wl[kk];
      }
      live = emax >= 0;
    } else {
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        const int u = lane + 32 * kk;
        wb[kk] = CTC_DEAD; wl[kk] = CTC_DEAD;
        if (u <= U) {
          const float2 a = __ldg(al2 + u), c = __ldg(be2 + u);
          wb[kk] = a.x + c.x;
          if (u < U) wl[kk] = a.y + c.y;
        }
        vmax = fmaxf(vmax, fmaxf(wb[kk], wl[kk]));
      }
      vmax = warp_max(vmax);
#pragma unroll
      for (int kk = 0; kk < NP; ++kk) {
        wb[kk] = ex2f(wb[kk] - vmax);
        wl[kk] = ex2f(wl[kk] - vmax);
        z += wb[kk] + wl[kk];
      }
      live = vmax > CTC_DEAD_TEST;
    }
    z = warp_sum(z);
    const float inv = (live && z > 0.f) ? 1.f / z : 0.f;
    __syncwarp();                                     // softmax row complete before the scatter
    // Every other lattice node is the blank: its contributions are summed in registers and
    // added once; label nodes scatter with shared-memory atomics (labels may repeat).
    int64_t lab[NP > 0 ? NP : 1];
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) { const int u = lane + 32 * kk; lab[kk] = u < U ? tg[u] : 0; }   // requested together, ahead of the atomics
#pragma unroll
    for (int kk = 0; kk < NP; ++kk) {
      const int u = lane + 32 * kk;
      bsum += wb[kk] * inv;
      if (u < U && wl[kk] != 0.f) atomicAdd(r + lab[kk], -wl[kk] * inv);   // a label outside the vocabulary has occupancy 0
    }
  } else {
    for (int u = lane; u <= U; u += 32) {
      const float2 a = __ldg(al2 + u), c = __ldg(be2 + u);
      vmax = fmaxf(vmax, a.x + c.x);
      if (u < U) vmax = fmaxf(vmax, a.y + c.y);
    }
    vmax = warp_max(vmax);
    for (int u = lane; u <= U; u += 32) {
      const float2 a = __ldg(al2 + u), c = __ldg(be2 + u);
      z += ex2f(a.x + c.x - vmax) + (u < U ? ex2f(a.y + c.y - vmax) : 0.f);
    }
    z = warp_sum(z);
    const float inv = (vmax > CTC_DEAD_TEST && z > 0.f) ? 1.f / z : 0.f;
    __syncwarp();
    for (int u = lane; u <= U; u += 32) {
      const float2 a = __ldg(al2 + u), c = __ldg(be2 + u);
      bsum += ex2f(a.x + c.x - vmax) * inv;
      if (u < U) atomicAdd(r + tg[u], -ex2f(a.y + c.y - vmax) * inv);
    }
  }
  bsum = warp_sum(bsum);
  __syncwarp();
  if (lane == 0) r[blank] -= bsum;
  __syncwarp();
  float scale;
  // grad_out == nullptr: unit upstream gradient (sc_ctc_head forms the gradient before autograd has one)
  if (reduction == 1) scale = (grad_out ? grad_out[0] : 1.f) / ((float)B * fmaxf((float)U, 1.f));
  else if (reduction == 2) scale = grad_out ? grad_out[0] : 1.f;
  else scale = grad_out ? grad_out[b] : 1.f;
  if (vec_out) {
    for (int i = lane * VWO; i < V; i += 32 * VWO) {
      float f[VWO];
#pragma unroll
      for (int j = 0; j < VWO; j += 4) {
        const float4 q4 = *reinterpret_cast<const float4*>(r + i + j);
        f[j] = scale * q4.x; f[j + 1] = scale * q4.y; f[j + 2] = scale * q4.z; f[j + 3] = scale * q4.w;
      }
      const Vec<TO, VWO> o = pack(f, (TO*)nullptr);
      *reinterpret_cast<uint4*>(dx + i) = o.raw;
    }
  } else {
    for (int i = lane; i < V; i += 32) st_f(dx + i, scale * r[i]);
  }
}


// Measured and dropped (r02, cfg2 shape, 0.30 ms = 3.8 TB/s of actual traffic as it stands; ncu: 46 % of DRAM peak,
// long-scoreboard stalls dominate, 45 % issue-active): (a) requesting the frame's lattice words and labels before the
// V-wide softmax pass (64 instead of 40 registers: 0.30 ms alone, 0.35 -> 0.40 ms inside the step); (b) a persistent grid
// whose warps keep the NEXT frame's logits / alpha / beta rows in flight as cp.async copies into a double-buffered
// per-warp stage (bit-identical, 13 KB of shared memory per warp -> 16 warps per SM): 0.39 ms.  A frame is ~800
// dependent-ish instructions for its warp; what hides that is the 43 resident warps per SM of this form, not a deeper
// load queue.  (c) the softmax row kept in REGISTERS: scale * softmax written straight from the loaded 16-byte words, the
// <= U + 1 entries that carry an occupancy rewritten afterwards as scale * (softmax - occupancy) from a per-warp table
// indexed by each label's first occurrence (640 B of shared memory per warp instead of 4 V; 64 registers): correct
// (112 CTC tests), but 0.36 ms against 0.28 ms alone and 0.46 against 0.36 ms inside the step — the ~10 scattered
// global accesses per lane (gather of the label logits, 2-byte stores into the finished row) cost more LSU time than
// the shared-memory row they replace.  Nor did it make the overlapped head pay (1.21 ms against 1.15 in step).
template <typename TI, typename TO, int NP>
__global__ void __launch_bounds__(CTC_WARPS * 32)
ctc_grad_kernel(const TI* __restrict__ logits, int64_t stride_b, int64_t stride_t,
                const int64_t* __restrict__ targets, int64_t ldt,
                const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                int B, int Tn, int V, int Smax, int64_t blank,
                const float* __restrict__ lse, const float* __restrict__ alpha,
                const float* __restrict__ beta, const float* __restrict__ nll,
                const float* __restrict__ grad_out, int reduction,
                TO* __restrict__ dlogits, int64_t dstride_b, int64_t dstride_t, const int* __restrict__ lossy,
                int t0, int t1, int mode) {
  // frames [t0, t1) of every utterance (the whole segment: 0, Tn).  CTC_GRAD_FIX: just the utterances flagged in
  // `lossy` or found infeasible (the rows of the others were written by earlier CTC_GRAD_SPEC launches and stand)
  extern __shared__ __align__(128) float sm[];       // per warp: V floats (softmax row, then the gradient row)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned wpb = blockDim.x >> 5;               // warps per block shrink for large V
  const unsigned span = (unsigned)(t1 - t0);
  const unsigned nidx = (unsigned)B * span;
  for (unsigned idx = blockIdx.x * wpb + warp; idx < nidx; idx += gridDim.x * wpb) {   // launched with one warp per row (a capped, persistent grid measured slower)
    const unsigned bb = idx / span;
    if (mode == CTC_GRAD_FIX && !((lossy != nullptr && lossy[bb]) || !isfinite(nll[bb]))) continue;
    const unsigned row = bb * (unsigned)Tn + (unsigned)t0 + (idx - bb * span);
    ctc_grad_row<TI, TO, NP>(sm, row, warp, lane, logits, stride_b, stride_t, targets, ldt, in_lens, tgt_lens, B, Tn, V,
                             Smax, blank, lse, alpha, beta, nll, grad_out, reduction, dlogits, dstride_b, dstride_t, lossy, mode);
    __syncwarp();                                     // the warp's shared-memory row is reused by its next frame
  }
}

}  // namespace sc

using namespace sc;

// Which lattice representation a call uses — decided from Umax alone, so the three entry points agree without
// sharing state: lattices of up to 32*LIN_MAXK - 1 labels run the fp64 linear-domain recursion (sc_ctc_lin64.cuh),
// larger ones the log-domain kernels.  SC_CTC_LIN=0 in the environment forces the log-domain path (A/B runs).
static bool ctc_use_lin(int64_t Umax) {
  static const bool off = [] { const char* e = getenv("SC_CTC_LIN"); return e && e[0] == '0'; }();
  return !off && Umax + 1 <= 32 * LIN_MAXK;
}
static int ctc_lin_pitch(int64_t Umax) { return lin_row_pitch((int)Umax); }   // words reserved per frame: the widest utterance's row (U+1 emissions, a zero word, padding to 32 K + 4)

// 4-byte words per frame the caller reserves for `lplat`: the lattice row (S = 2 Umax + 1 rounded up to 4) or, where the
// fp64 kernel serves, the emission row at ITS pitch if that is wider (narrow lattices)
extern "C" int64_t sc_ctc_lplat_pitch(int64_t Umax) {
  if (Umax < 0) return 4;
  const int64_t S = (2 * Umax + 1 + 7) & ~(int64_t)7;
  const int64_t LP = ctc_use_lin(Umax) ? ctc_lin_pitch(Umax) : 0;
  return S > LP ? S : LP;
}

extern "C" int64_t sc_ctc_workspace_bytes(int64_t B, int64_t T, int64_t Umax) {
  (void)Umax;
  if (B <= 0 || T < 0) return 16;
  return ctc_ws_bytes(B, T);
}

// pass 1 in the linear emission format over frames [t0, t1) of every utterance
static int launch_emissions_lin(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                                const int64_t* targets, int64_t ldt, const int64_t* in_lens, const int64_t* tgt_lens,
                                int64_t B, int64_t T, int64_t V, int64_t Umax, int64_t blank, float* lse, float* lplat,
                                float* cshift, int t0, int t1, cudaStream_t st) {
  if (t1 <= t0) return 0;
  const int LP = ctc_lin_pitch(Umax);
  const unsigned blocks = (unsigned)cdiv(B * (int64_t)(t1 - t0), CTC_WARPS);
#define SC_CTC_E(TT, NL) ctc_lse_gather_lin_kernel<TT, NL><<<blocks, CTC_WARPS * 32, 0, st>>>((const TT*)logits, stride_b, stride_t, \
      targets, ldt, in_lens, tgt_lens, (int)B, (int)T, (int)V, (int)Umax, LP, blank, lse, (uint32_t*)lplat, cshift, t0, t1)
  if (dtype == SC_F32) { if (Umax + 1 <= 160) SC_CTC_E(float, 5); else SC_CTC_E(float, 8); }
  else { if (Umax + 1 <= 160) SC_CTC_E(bf16, 5); else SC_CTC_E(bf16, 8); }
#undef SC_CTC_E
  SC_LAUNCH_RET();
}

extern "C" int sc_ctc_emissions(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                                const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                                const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                                int64_t blank, float* lse, float* lplat, float* cshift, void* stream) {
  SC_CHECK_ARG(B > 0 && T >= 0 && V > 0 && Umax >= 0 && blank >= 0 && blank < V, SC_E_BADARG);
  SC_CHECK_ARG(in_lens && tgt_lens && (Umax == 0 || targets), SC_E_BADARG);
  SC_CHECK_ARG(T == 0 || (logits && lse && lplat && cshift), SC_E_BADARG);
  SC_CHECK_ARG(B * T < ((int64_t)1 << 31) && V < (1 << 30) && Umax < (1 << 20), SC_E_SHAPE);
  SC_CHECK_ARG(dtype == SC_F32 || dtype == SC_BF16, SC_E_DTYPE);
  if (T == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int Smax = (int)((2 * Umax + 1 + 7) & ~(int64_t)7);      // row width of alpha/beta: whole 32-byte sectors
  if (ctc_use_lin(Umax))
    return launch_emissions_lin(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank,
                                lse, lplat, cshift, 0, (int)T, st);
  const unsigned blocks = (unsigned)cdiv(B * T, CTC_WARPS);
  if (dtype == SC_F32)
    ctc_lse_gather_kernel<float><<<blocks, CTC_WARPS * 32, 0, st>>>((const float*)logits, stride_b, stride_t,
        targets, ldt, in_lens, tgt_lens, (int)B, (int)T, (int)V, (int)Umax, Smax, blank, lse, lplat, cshift);
  else
    ctc_lse_gather_kernel<bf16><<<blocks, CTC_WARPS * 32, 0, st>>>((const bf16*)logits, stride_b, stride_t,
        targets, ldt, in_lens, tgt_lens, (int)B, (int)T, (int)V, (int)Umax, Smax, blank, lse, lplat, cshift);
  SC_LAUNCH_RET();
}

// Rows of emissions per block meeting of the wavefront kernel: as many as fit next to the
// other CTAs that have to share an SM (2B CTAs over the device), at most 64.
static int ctc_wave_rows(int pitch, int64_t B, int nwarps, size_t* smem_out) {
  const int sms = num_sms();
  int per_sm = (int)((2 * B + sms - 1) / sms);
  if (per_sm > 8) per_sm = 8;
  const size_t budget = (size_t)220 * 1024 / (size_t)per_sm - 2048;
  int forced = 0;
  if (const char* ev = getenv("SC_CTC_EB")) forced = atoi(ev);
  for (int eb = 64; eb >= 4; eb >>= 1) {
    if (forced > 0 && eb != forced) continue;
    const size_t need = 2 * (size_t)eb * pitch * sizeof(float) + (size_t)nwarps * (eb + 1) * 8 + 16;
    if (need <= budget || eb == 4 || forced > 0) { *smem_out = need; return eb; }
  }
  return 0;
}

template <int FMT>
static int launch_wave2(const float* lplat, const float* cshift, const int64_t* targets, int64_t ldt,
                        const int64_t* in_lens, const int64_t* tgt_lens, int64_t B, int64_t T, int64_t Umax,
                        int Smax, int LP, const int* lossy, float* alpha, float* beta, float* nll, cudaStream_t st) {
  const int wthreads = (((int)Umax + 1 + 31) / 32) * 32;
  size_t smem = 0;
  const int eb = ctc_wave_rows(FMT == 1 ? LP : Smax, B, wthreads / 32, &smem);
  SC_CHECK_ARG(eb > 0 && smem <= 220 * 1024, SC_E_SHAPE);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctc_alpha_beta_wave2_kernel<FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  ctc_alpha_beta_wave2_kernel<FMT><<<dim3((unsigned)B, 2), wthreads, smem, st>>>(lplat, cshift, targets, ldt, in_lens, tgt_lens,
      (int)T, (int)Umax, Smax, LP, eb, lossy, alpha, beta, nll);
  return 0;
}

static int launch_lin64(const float* lplat, const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                        const int64_t* tgt_lens, int64_t B, int64_t T, int64_t Umax, int Smax, int LP,
                        float* alpha, float* beta, float* nll, const CtcWs& w, int a0, int a1, int b0, int b1,
                        cudaStream_t st) {
  // alpha over frames [a0, a1), beta over [b0, b1) (a0, b0 multiples of LIN_EB); the whole segment: 0, T, 0, T
  const int Kmax = (int)((Umax + 32) >> 5);
  auto need = [&](int nb) { return (2 * (size_t)LIN_EB * (32 * Kmax + LIN_EPAD) + (size_t)nb * LIN_ROWS * 64 * Kmax) * sizeof(uint32_t); };
  const bool deep = need(LIN_NB) <= 200 * 1024;                  // the widest lattices (7 or 8 pairs per lane) get the shorter ring
  const size_t smem = need(deep ? LIN_NB : LIN_NB_WIDE);
  auto kern = deep ? ctc_lin64_kernel<LIN_NB> : ctc_lin64_kernel<LIN_NB_WIDE>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  kern<<<dim3((unsigned)B, 2), LIN_THREADS, smem, st>>>((const uint32_t*)lplat, targets, ldt, in_lens, tgt_lens,
      (int)T, (int)Umax, LP, Smax, alpha, beta, nll, w, a0, a1, b0, b1);
  return 0;
}

// after the fp64 recursions: likelihood + range check, then the log-domain recomputation of the utterances it flags
static int ctc_lin_tail(const float* lplat, const float* cshift, const int64_t* targets, int64_t ldt,
                        const int64_t* in_lens, const int64_t* tgt_lens, int64_t B, int64_t T, int64_t Umax,
                        int Smax, int LP, float* alpha, float* beta, float* nll, const CtcWs& w, cudaStream_t st) {
  int force = 0;
  if (const char* ev = getenv("SC_CTC_FORCE_LOSSY")) force = ev[0] == '1';     // tests: send every utterance down the recomputation path
  const unsigned nsb = 1 + (unsigned)cdiv(cdiv(T, LIN_SAMPLE), LIN_CHECK_THREADS / 32);
  ctc_lin64_check_kernel<<<dim3((unsigned)B, nsb), LIN_CHECK_THREADS, 0, st>>>(cshift, in_lens, tgt_lens, (int)T, (int)Umax, Smax, force, alpha, beta, nll, w);
  if (T > 0) return launch_wave2<1>(lplat, cshift, targets, ldt, in_lens, tgt_lens, B, T, Umax, Smax, LP, w.lossy, alpha, beta, nll, st);
  return 0;
}

extern "C" int sc_ctc_lattice(const float* lplat, const float* cshift, const int64_t* targets, int64_t ldt,
                              const int64_t* in_lens, const int64_t* tgt_lens, int64_t B, int64_t T,
                              int64_t Umax, int64_t blank, float* alpha, float* beta, float* nll,
                              float* loss, int reduction, void* ws, void* stream) {
  SC_CHECK_ARG(B > 0 && T >= 0 && Umax >= 0 && blank >= 0, SC_E_BADARG);
  SC_CHECK_ARG(in_lens && tgt_lens && nll && ws && (Umax == 0 || targets), SC_E_BADARG);
  SC_CHECK_ARG(reduction >= 0 && reduction <= 2 && (reduction == 0 || loss), SC_E_BADARG);
  SC_CHECK_ARG(T == 0 || (lplat && cshift && alpha && beta), SC_E_BADARG);
  SC_CHECK_ARG(B * T < ((int64_t)1 << 31) && Umax < (1 << 20), SC_E_SHAPE);
  SC_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 7) == 0, SC_E_ALIGN);
  cudaStream_t st = (cudaStream_t)stream;
  const int Smax = (int)((2 * Umax + 1 + 7) & ~(int64_t)7);
  int rc = 0;
  if (ctc_use_lin(Umax)) {
    // fp64 linear-domain recursion -> range check -> log-domain recomputation of the utterances it flags
    const CtcWs w = ctc_ws_carve(ws, B, T);
    const int LP = ctc_lin_pitch(Umax);
    rc = launch_lin64(lplat, targets, ldt, in_lens, tgt_lens, B, T, Umax, Smax, LP, alpha, beta, nll, w, 0, (int)T, 0, (int)T, st);
    if (rc) return rc;
    rc = ctc_lin_tail(lplat, cshift, targets, ldt, in_lens, tgt_lens, B, T, Umax, Smax, LP, alpha, beta, nll, w, st);
    if (rc) return rc;
  } else if (Smax <= 1024 && !(getenv("SC_CTC_WAVE") && getenv("SC_CTC_WAVE")[0] == '0')) {
    rc = launch_wave2<0>(lplat, cshift, targets, ldt, in_lens, tgt_lens, B, T, Umax, Smax, 0, nullptr, alpha, beta, nll, st);
    if (rc) return rc;
  } else {
    // block-barrier kernel, any lattice width: two recursion lines + (fast path, S <= 1024) two blocks of CTC_EB emission rows
    int threads = ((Smax + 31) / 32) * 32;
    if (threads > 1024) threads = 1024;
    const size_t smem = (2 * (size_t)(Smax + 4) + (Smax <= 1024 ? 2 * (size_t)CTC_EB * Smax : 0)) * sizeof(float);
    SC_CHECK_ARG(smem <= 200 * 1024, SC_E_SHAPE);
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(ctc_alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    ctc_alpha_beta_kernel<<<dim3((unsigned)B, 2), threads, smem, st>>>(lplat, cshift, targets, ldt, in_lens, tgt_lens,
        (int)T, (int)Umax, Smax, blank, alpha, beta, nll);
  }
  if (reduction != 0) ctc_reduce_kernel<<<1, 32, 0, st>>>(nll, tgt_lens, (int)B, reduction, loss);
  SC_LAUNCH_RET();
}

extern "C" int sc_ctc_fwd(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                          const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                          const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                          int64_t blank, float* lse, float* lplat, float* cshift, float* alpha, float* beta,
                          float* nll, float* loss, int reduction, void* ws, void* stream) {
  SC_CHECK_ARG(blank >= 0 && blank < V, SC_E_BADARG);
  const int rc = sc_ctc_emissions(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax,
                                  blank, lse, lplat, cshift, stream);
  if (rc) return rc;
  return sc_ctc_lattice(lplat, cshift, targets, ldt, in_lens, tgt_lens, B, T, Umax, blank, alpha, beta, nll, loss, reduction,
                        ws, stream);
}

template <typename TI, typename TO, int NP>
static int launch_ctc_grad(const void* logits, int64_t stride_b, int64_t stride_t,
                           const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                           const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int Smax,
                           int64_t blank, const float* lse, const float* alpha, const float* beta,
                           const float* nll, const float* grad_out, int reduction, void* dlogits,
                           int64_t dstride_b, int64_t dstride_t, const int* lossy, int t0, int t1, int mode,
                           cudaStream_t st) {
  // one shared-memory row of V floats per warp: fewer warps per block when the vocabulary is large
  const size_t per_warp = (size_t)V * sizeof(float);
  int warps = CTC_WARPS;
  while (warps > 1 && per_warp * warps > 200 * 1024) warps >>= 1;
  const size_t smem = per_warp * warps;
  if (smem > 200 * 1024) return SC_E_SHAPE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctc_grad_kernel<TI, TO, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  unsigned blocks = (unsigned)cdiv(B * (int64_t)(t1 - t0), warps);
  if (mode == CTC_GRAD_FIX && blocks > (unsigned)num_sms() * 4u) blocks = (unsigned)num_sms() * 4u;   // normally nothing to do: a small grid that strides
  ctc_grad_kernel<TI, TO, NP><<<blocks, warps * 32, smem, st>>>((const TI*)logits, stride_b, stride_t, targets, ldt,
      in_lens, tgt_lens, (int)B, (int)T, (int)V, Smax, blank, lse, alpha, beta, nll, grad_out, reduction,
      (TO*)dlogits, dstride_b, dstride_t, lossy, t0, t1, mode);
  SC_LAUNCH_RET();
}

template <typename TI, typename TO, typename... A>
static int ctc_grad_by_width(int Smax, A... a) {
  if (Smax <= 128) return launch_ctc_grad<TI, TO, 2>(a...);
  if (Smax <= 320) return launch_ctc_grad<TI, TO, 5>(a...);
  if (Smax <= 640) return launch_ctc_grad<TI, TO, 10>(a...);
  return launch_ctc_grad<TI, TO, 0>(a...);
}

static int ctc_bwd_range(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                                const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                                const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                                int64_t blank, const float* lse, const float* alpha, const float* beta,
                                const float* nll, const float* grad_out, int reduction,
                                void* dlogits, int64_t dstride_b, int64_t dstride_t, int out_dtype,
                                const void* ws, int64_t t0, int64_t t1, int mode, void* stream) {
  SC_CHECK_ARG(B > 0 && T >= 0 && V > 0 && Umax >= 0 && blank >= 0 && blank < V, SC_E_BADARG);
  SC_CHECK_ARG(t0 >= 0 && t1 <= T && t0 <= t1, SC_E_BADARG);
  if (t0 == t1) return 0;
  SC_CHECK_ARG(logits && in_lens && tgt_lens && lse && alpha && beta && nll && dlogits && ws, SC_E_BADARG);
  SC_CHECK_ARG(reduction >= 0 && reduction <= 2, SC_E_BADARG);
  SC_CHECK_ARG(B * T < ((int64_t)1 << 31), SC_E_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  const int Smax = (int)((2 * Umax + 1 + 7) & ~(int64_t)7);
  // rows are in the linear-domain format except for the utterances flagged in the workspace (same rule as the forward)
  const int* lossy = ctc_use_lin(Umax) ? ctc_ws_carve(const_cast<void*>(ws), B, T).lossy : nullptr;
#define SC_CTC_GRAD(TI, TO) ctc_grad_by_width<TI, TO>(Smax, logits, stride_b, stride_t, targets, ldt, in_lens, tgt_lens, \
    B, T, V, Smax, blank, lse, alpha, beta, nll, grad_out, reduction, dlogits, dstride_b, dstride_t, lossy, \
    (int)t0, (int)t1, mode, st)
  if (dtype == SC_F32 && out_dtype == SC_F32) return SC_CTC_GRAD(float, float);
  if (dtype == SC_BF16 && out_dtype == SC_BF16) return SC_CTC_GRAD(bf16, bf16);
  if (dtype == SC_F32 && out_dtype == SC_BF16) return SC_CTC_GRAD(float, bf16);
  if (dtype == SC_BF16 && out_dtype == SC_F32) return SC_CTC_GRAD(bf16, float);
#undef SC_CTC_GRAD
  return SC_E_DTYPE;
}

extern "C" int sc_ctc_bwd(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                          const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                          const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                          int64_t blank, const float* lse, const float* alpha, const float* beta,
                          const float* nll, const float* grad_out, int reduction,
                          void* dlogits, int64_t dstride_b, int64_t dstride_t, int out_dtype,
                          const void* ws, void* stream) {
  SC_CHECK_ARG(T >= 0, SC_E_BADARG);
  if (T == 0) return 0;
  SC_CHECK_ARG(grad_out, SC_E_BADARG);
  return ctc_bwd_range(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank, lse,
                          alpha, beta, nll, grad_out, reduction, dlogits, dstride_b, dstride_t, out_dtype, ws, 0, T, 0, stream);
}

// ---- the head in one call: loss AND gradient, the V-wide passes under the recursions ------------------------------
// The recursions are a chain of T dependent steps on 2B warps: while they run the device is otherwise idle.  alpha
// needs emissions from frame 0 upwards and beta from the last frame downwards, and a frame's gradient row can be formed
// as soon as both have passed it (its normaliser is the frame's own sum over the lattice, not the final likelihood).
// So the segment is cut into P chunks of whole emission blocks and the recursions run as P launches over frame ranges
// (the column travels through the workspace in fp64: the rows are bit-identical to one launch): phase p takes alpha
// through chunk p and beta through chunk P-1-p.  On a second stream the emission pass produces the chunks in the
// order the phases need them (both ends inwards); on a third the gradient pass takes the chunks both directions
// have crossed (the middle outwards) for a unit upstream gradient.  The recursion launches need 11 K registers and
// ~100 KB of shared memory on an SM the V-wide kernels would otherwise fill: they go to `stream_l`, which the caller
// creates with a HIGHER priority than the side streams, so that their blocks are placed first as the others'
// retire (measured on a B200 without it: no overlap at all, 0.79 ms against 0.82 ms one after the other).
// Ordering is by events only — no kernel waits on
// another — so it is safe under serialising tools and inside a stream capture (both side streams fork from and join
// `stream`).  What stays exposed: the first two chunks' emissions, the last two chunks' gradient rows, the range
// check.  Utterances the check flags (and infeasible ones, whose rows must be zero) are redone by a last gradient
// launch.  stream_l null: the recursions stay on `stream`.  stream_e / stream_g null, phases == 1 or a lattice the fp64 kernel does not take: the same passes, one
// after the other, on `stream`.
static cudaEvent_t ctc_head_event(int i) {
  thread_local std::vector<std::vector<cudaEvent_t>> pools;
  int dev = 0;
  cudaGetDevice(&dev);
  if ((int)pools.size() <= dev) pools.resize(dev + 1);
  auto& pool = pools[dev];
  while ((int)pool.size() <= i) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    pool.push_back(e);
  }
  return pool[i];
}
#define SC_CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

// number of recursion launches sc_ctc_head cuts a segment of T frames into (1: nothing to overlap).  phases <= 0: the
// library's choice (SC_CTC_PHASES in the environment overrides it)
extern "C" int64_t sc_ctc_head_phases(int64_t T, int64_t Umax, int64_t phases) {
  if (T <= 0 || Umax < 0 || !ctc_use_lin(Umax)) return 1;
  const int64_t nblk = cdiv(T, LIN_EB);
  int64_t P = phases;
  if (P <= 0) {
    if (const char* ev = getenv("SC_CTC_PHASES")) P = atoi(ev);
    if (P <= 0) P = nblk / 6;                                    // >= 6 emission blocks (384 frames) per launch
    if (P > 8) P = 8;
  }
  if (P > nblk) P = nblk;
  if (P < 2) return 1;
  const int64_t C = cdiv(nblk, P) * LIN_EB;                      // whole emission blocks per chunk
  return cdiv(T, C);
}

extern "C" int sc_ctc_head(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                           const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                           const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                           int64_t blank, float* lse, float* lplat, float* cshift, float* alpha, float* beta,
                           float* nll, float* loss, int reduction, void* ws,
                           void* dlogits, int64_t dstride_b, int64_t dstride_t, int out_dtype,
                           int64_t phases, void* stream, void* stream_l, void* stream_e, void* stream_g) {
  SC_CHECK_ARG(B > 0 && T >= 0 && V > 0 && Umax >= 0 && blank >= 0 && blank < V, SC_E_BADARG);
  SC_CHECK_ARG(reduction >= 1 && reduction <= 2 && loss && ws && nll && in_lens && tgt_lens, SC_E_BADARG);
  SC_CHECK_ARG(T == 0 || (logits && lse && lplat && cshift && alpha && beta && dlogits), SC_E_BADARG);
  SC_CHECK_ARG(B * T < ((int64_t)1 << 31) && V < (1 << 30) && Umax < (1 << 20), SC_E_SHAPE);
  SC_CHECK_ARG((dtype == SC_F32 || dtype == SC_BF16) && (out_dtype == SC_F32 || out_dtype == SC_BF16), SC_E_DTYPE);
  SC_CHECK_ARG(!(dtype == SC_BF16 && out_dtype == SC_F32), SC_E_DTYPE);
  cudaStream_t st = (cudaStream_t)stream, se = (cudaStream_t)stream_e, sg = (cudaStream_t)stream_g;
  cudaStream_t sl = stream_l ? (cudaStream_t)stream_l : st;      // the recursions' stream (a high-priority one: see above)
  const int64_t P = sc_ctc_head_phases(T, Umax, phases);
  const bool overlap = P >= 2 && se && sg && se != st && sg != st;
  if (!overlap) {
    int rc = sc_ctc_fwd(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank, lse,
                        lplat, cshift, alpha, beta, nll, loss, reduction, ws, stream);
    if (rc || T == 0) return rc;
    return ctc_bwd_range(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank, lse,
                         alpha, beta, nll, nullptr, reduction, dlogits, dstride_b, dstride_t, out_dtype, ws, 0, T,
                         CTC_GRAD_ALL, stream);
  }
  SC_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 7) == 0, SC_E_ALIGN);
  const int Smax = (int)((2 * Umax + 1 + 7) & ~(int64_t)7);
  const int LP = ctc_lin_pitch(Umax);
  const CtcWs w = ctc_ws_carve(ws, B, T);
  const int64_t C = cdiv(T, P * LIN_EB) * LIN_EB;                // frames per chunk: whole emission blocks (every chunk non-empty: sc_ctc_head_phases)
  auto edge = [&](int64_t k) -> int { return (int)(k * C < T ? k * C : T); };   // chunk k = frames [edge(k), edge(k+1))
  const int nE = (int)((P + 1) / 2);
  int nev = 0;
  cudaEvent_t ev0 = ctc_head_event(nev++);
  SC_CHECK_ARG(ev0, SC_E_BADARG);
  SC_CU(cudaEventRecord(ev0, st));                               // the logits are ready
  SC_CU(cudaStreamWaitEvent(se, ev0, 0));
  if (sl != st) SC_CU(cudaStreamWaitEvent(sl, ev0, 0));
  std::vector<cudaEvent_t> evE(nE);
  int rc = 0;
  for (int p = 0; p < nE; ++p) {                                 // both ends inwards
    const int q = (int)P - 1 - p;
    rc = launch_emissions_lin(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank,
                              lse, lplat, cshift, edge(p), edge(p + 1), se);
    if (rc) return rc;
    if (q != p) {
      rc = launch_emissions_lin(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank,
                                lse, lplat, cshift, edge(q), edge(q + 1), se);
      if (rc) return rc;
    }
    evE[p] = ctc_head_event(nev++);
    SC_CHECK_ARG(evE[p], SC_E_BADARG);
    SC_CU(cudaEventRecord(evE[p], se));
  }
  for (int p = 0; p < (int)P; ++p) {
    const int q = (int)P - 1 - p;
    if (p < nE) SC_CU(cudaStreamWaitEvent(sl, evE[p], 0));
    rc = launch_lin64(lplat, targets, ldt, in_lens, tgt_lens, B, T, Umax, Smax, LP, alpha, beta, nll, w,
                      edge(p), edge(p + 1), edge(q), edge(q + 1), sl);
    if (rc) return rc;
    if (p >= q) {                                                // both directions have crossed chunks p and q
      cudaEvent_t evL = ctc_head_event(nev++);
      SC_CHECK_ARG(evL, SC_E_BADARG);
      SC_CU(cudaEventRecord(evL, sl));
      SC_CU(cudaStreamWaitEvent(sg, evL, 0));
      rc = ctc_bwd_range(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank, lse,
                         alpha, beta, nll, nullptr, reduction, dlogits, dstride_b, dstride_t, out_dtype, ws,
                         edge(q), edge(q + 1), CTC_GRAD_SPEC, sg);
      if (rc) return rc;
      if (q != p) {
        rc = ctc_bwd_range(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank, lse,
                           alpha, beta, nll, nullptr, reduction, dlogits, dstride_b, dstride_t, out_dtype, ws,
                           edge(p), edge(p + 1), CTC_GRAD_SPEC, sg);
        if (rc) return rc;
      }
    }
  }
  rc = ctc_lin_tail(lplat, cshift, targets, ldt, in_lens, tgt_lens, B, T, Umax, Smax, LP, alpha, beta, nll, w, sl);
  if (rc) return rc;
  ctc_reduce_kernel<<<1, 32, 0, sl>>>(nll, tgt_lens, (int)B, reduction, loss);
  if (sl != st) {
    cudaEvent_t evT = ctc_head_event(nev++);
    SC_CHECK_ARG(evT, SC_E_BADARG);
    SC_CU(cudaEventRecord(evT, sl));
    SC_CU(cudaStreamWaitEvent(st, evT, 0));
  }
  cudaEvent_t evG = ctc_head_event(nev++);
  SC_CHECK_ARG(evG, SC_E_BADARG);
  SC_CU(cudaEventRecord(evG, sg));
  SC_CU(cudaStreamWaitEvent(st, evG, 0));
  return ctc_bwd_range(logits, stride_b, stride_t, dtype, targets, ldt, in_lens, tgt_lens, B, T, V, Umax, blank, lse,
                       alpha, beta, nll, nullptr, reduction, dlogits, dstride_b, dstride_t, out_dtype, ws, 0, T,
                       CTC_GRAD_FIX, stream);
}

// x *= *scale (device scalar), skipped entirely when *scale == 1: the gradient the overlapped head formed during the
// forward for a unit upstream gradient, brought to the upstream gradient autograd hands to the backward
template <typename T>
__global__ void ctc_scale_kernel(T* __restrict__ x, int64_t n, const float* __restrict__ scale) {
  const float sc = *scale;
  if (sc == 1.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st_f(x + i, ld_f(x + i) * sc);
}
extern "C" int sc_ctc_scale_grad(void* x, int dtype, int64_t n, const float* scale, void* stream) {
  SC_CHECK_ARG(n >= 0 && (n == 0 || (x && scale)), SC_E_BADARG);
  if (n == 0) return 0;
  const unsigned blocks = (unsigned)min((int64_t)148 * 16, cdiv(n, 256));
  if (dtype == SC_F32) ctc_scale_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((float*)x, n, scale);
  else if (dtype == SC_BF16) ctc_scale_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)x, n, scale);
  else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}
