// K3, pass 2 for lattices of up to 255 labels: the alpha/beta recursions in the LINEAR domain on fp64.
// Included by sc_ctc.cu (uses its helpers); replaces model.py:70-71's ctc_loss_gpu recursions.
//
// Why: the log-domain step is a dependent chain max -> ex2 -> add -> lg2 -> add per node; measured on a B200
// (profiles/micro/lat_bench.cu) that chain is 87 cycles for ONE two-term node with nothing else in the warp, and
// 285 cycles per timestep in the shipped pair-per-thread kernel once the cross-warp handoff, stores and address
// arithmetic share the in-order issue slot.  In the linear domain a step is alpha'(s) = (alpha(s) + alpha(s-1) +
// skip*alpha(s-2)) * p_t(s): DADD/DFMA/DMUL at 8.4 cycles each, no MUFU at all.  fp32 cannot hold it (with T >> U
// the forward and backward masses sit at opposite ends of the lattice; the nodes that carry the occupancy lie
// 2^-125 and further below a column's maximum — r01), a per-node software exponent costs more integer work than
// it saves (SC_CTC_WAVE=3/4 of r01, measured 0.73 / 1.00 ms against 0.46 ms: deleted), but fp64 IS a hardware
// (mantissa, 11-bit exponent) pair: a column kept with its maximum near 2^400 has 1400 binary orders below it.
//
// Shape: ONE WARP per (utterance, direction), no barrier and no polling anywhere.  Lane i owns the (blank, label)
// pairs c = 32 j + i, j < K, of the scan order (alpha: pair c = nodes 2c, 2c+1; beta: the mirrored lattice), so for
// a fixed j the warp's pairs are adjacent in memory and every row leaves through K coalesced 8-byte stores.  A label
// needs the previous pair's label: one lane rotation per j (lane 0 takes lane 31's value of j-1).  All K chains of
// a step are independent, so the in-order warp always has DP work to issue while a shuffle is in flight.
//
// Formats (4 bytes per node, as before):
//  * emissions  lplat[b,t,:] (pitch LP words): [0] = p(blank), [1+u] = p(label u), each the HIGH WORD of the fp64
//    value p = 2^(e - c_t) (c_t = the frame's largest lattice emission, so p <= 1; cshift[b,t] = c_t is added back
//    into the likelihood).  (hi, lo=0) is a valid double: no conversion instruction in the recursion.  20 mantissa
//    bits = a 4.8e-7 relative perturbation of each emission, the size of fp32 rounding of the logit itself; both
//    directions read the same words, so alpha, beta and the likelihood stay mutually consistent.  All blank nodes
//    share one emission: U+1 words per frame instead of 2U+1.
//  * alpha / beta rows: the high word of each node's fp64 value (alpha WITH its frame's emission, beta without), at
//    its SCAN-ORDER position (beta row reversed: node s sits at 2U - s).  The gradient pass adds exponents and
//    multiplies the 20-bit mantissas; per-column scale factors cancel in its per-frame normalisation.
//
// Range: every 8 steps the warp takes the column maximum (integer max of high words, one REDUX) and, when it has
// drifted by more than 2^48 from 2^900, rescales by an exact power of two (integer add on the exponent fields,
// the shift accumulated as an integer).  A column then has ~1700 binary orders below its maximum; measured need at
// T=3000, U=150, V=1024 on N(0, s^2) logits (the most occupied node against the column maximum): 741 (s=1),
// 944 (s=2), 1355 (s=3).  What fp64 cannot hold is detected, never silently accepted — ctc_lin64_check_kernel:
//  (a) the column maximum falling by more than 2^200 within 8 steps -> `danger` (the guaranteed range is gone);
//  (b) the two directions' likelihoods disagreeing (mass lost to underflow in one direction shows up here);
//  (c) at every 32nd frame, sum_s alpha_t(s) beta_t(s) — formed from the stored rows exactly the way the gradient
//      pass forms occupancies — must reproduce that likelihood to 1e-4 (log2): a flushed node that carried more
//      than that share of any sampled frame's mass fails it.
// Any of them sets lossy[b] = 1 and the log-domain pair-per-thread kernel recomputes exactly those utterances (it
// exits at once for the others); the gradient pass reads lossy[b] to know which row format it is looking at.
#pragma once

namespace sc {

constexpr int LIN_EB = 64;                 // emission rows per bulk-copied shared-memory block
constexpr int LIN_CHECK = 8;               // steps between range checks
constexpr int LIN_TGT = 1023 + 900;        // exponent field the column maximum is kept near (growth is < 2^13 per 8 steps: no overflow)
constexpr int LIN_HYST = 48;               // rescale when the maximum is further than 2^48 from the target
constexpr int LIN_DANGER = LIN_TGT - LIN_HYST - 200;   // column maximum this low at a check: it fell by > 2^200 within 8 steps
constexpr int LIN_MAXK = 8;                // pairs per lane: lattices of up to 32*8 - 1 labels
constexpr int LIN_SAMPLE = 32;             // check (c) looks at frames t = 15 mod 32

// layout of the opaque workspace `ws` (sc_ctc_workspace_bytes)
struct CtcWs {
  int* lossy;        // [B]      1: rows of this utterance are in the log-domain format (recomputed)
  int* danger;       // [B][2]   per direction: range event (a) above
  double* zl2;       // [B][2]   per direction: log2 of the shift-free likelihood (-inf when zero)
  int* accs;         // [B][2][NT] log2 of the scale taken out of the rows of frames [8k, 8k+7], per direction
  int NT;
  double* state;     // [B][2][2*LIN_MAXK+1][32] column of a direction between two launches over frame ranges
  int* istate;       // [B][2][2]  {acc, danger} of the same
};
__host__ __device__ inline int ctc_ws_nt(int64_t T) { return (int)(T / LIN_CHECK) + 2; }
__host__ __device__ inline CtcWs ctc_ws_carve(void* ws, int64_t B, int64_t T) {
  CtcWs w;
  char* p = reinterpret_cast<char*>(ws);
  w.NT = ctc_ws_nt(T);
  w.zl2 = reinterpret_cast<double*>(p);                          p += (size_t)B * 2 * sizeof(double);
  w.lossy = reinterpret_cast<int*>(p);                           p += (((size_t)B * 4 + 15) & ~(size_t)15);
  w.danger = reinterpret_cast<int*>(p);                          p += (((size_t)B * 8 + 15) & ~(size_t)15);
  w.accs = reinterpret_cast<int*>(p);                            p += (((size_t)B * 2 * w.NT * 4 + 15) & ~(size_t)15);
  w.state = reinterpret_cast<double*>(p);                        p += (size_t)B * 2 * (2 * 8 + 1) * 32 * sizeof(double);
  w.istate = reinterpret_cast<int*>(p);
  return w;
}
inline int64_t ctc_ws_bytes(int64_t B, int64_t T) {
  return (int64_t)(B * 16 + ((B * 4 + 15) & ~(int64_t)15) + ((B * 8 + 15) & ~(int64_t)15) +
                   ((B * 2 * ctc_ws_nt(T) * 4 + 15) & ~(int64_t)15) + B * 2 * (2 * 8 + 1) * 32 * 8 + B * 2 * 2 * 4 + 32);
}

// high word of 2^d (d <= 0, log2 units), mantissa rounded to 20 bits; 0 = probability zero
__device__ __forceinline__ uint32_t lin_word_of_log2(float d) {
  const bool dead = !(d > -1000.f);                              // (select, not a branch: the callers convert a handful of words per lane)
  const float dd = dead ? 0.f : d;
  const float fl = floorf(dd);
  const uint32_t mb = __float_as_uint(ex2f(dd - fl));            // 2^frac in [1, 2]: exponent field 127 (128 when it rounds to 2)
  const uint32_t w = (uint32_t)((896 + (int)fl) << 20) + ((mb + 4u) >> 3);   // (1023 - 127 + fl) << 20, + the float's own exponent/mantissa >> 3
  return dead ? 0u : w;
}
// the reverse, for the log-domain kernels reading the same emission words
__device__ __forceinline__ float lin_word_to_log2(uint32_t w) {
  if (w == 0u) return -1e30f;
  return (float)((int)(w >> 20) - 1023) + lg2f(__uint_as_float(0x3f800000u | ((w & 0xfffffu) << 3)));
}
// words per emission row of an utterance with U labels, in global and in shared memory: 32 K + LIN_EPAD, K = pairs per lane
// (>= U + 2: blank, labels, the zero word; a whole number of 32-byte sectors, so that rows do not straddle them in HBM)
constexpr int LIN_EPAD = 8;
__host__ __device__ __forceinline__ int lin_row_pitch(int U) { return 32 * ((U + 32) >> 5) + LIN_EPAD; }
__device__ __forceinline__ double lin_hi2d(uint32_t w) { return __hiloint2double((int)w, 0); }
__device__ __forceinline__ int lin_hi(double v) { return __double2hiint(v); }
__device__ __forceinline__ double lin_rot(double v, int src) {
  return __hiloint2double(__shfl_sync(0xffffffffu, __double2hiint(v), src), __shfl_sync(0xffffffffu, __double2loint(v), src));
}

// ---- pass 1, linear emission format -------------------------------------------------------------------------
template <typename T, int NL>
__global__ void __launch_bounds__(CTC_WARPS * 32, 5)
ctc_lse_gather_lin_kernel(const T* __restrict__ logits, int64_t stride_b, int64_t stride_t,
                          const int64_t* __restrict__ targets, int64_t ldt,
                          const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                          int B, int Tn, int V, int Umax, int LP, int64_t blank,
                          float* __restrict__ lse, uint32_t* __restrict__ lplat, float* __restrict__ cshift,
                          int t0, int t1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned span = (unsigned)(t1 - t0);                     // frames [t0, t1) of every utterance
  const unsigned nidx = (unsigned)B * span;
  for (unsigned idx = blockIdx.x * CTC_WARPS + warp; idx < nidx; idx += gridDim.x * CTC_WARPS) {
    const int b = (int)(idx / span), t = t0 + (int)(idx - (unsigned)b * span);
    const unsigned row = (unsigned)b * (unsigned)Tn + (unsigned)t;
    int64_t Tb = in_lens[b]; if (Tb > Tn) Tb = Tn;
    if (t >= Tb) continue;
    const int64_t U64 = tgt_lens[b];
    if (U64 < 0 || U64 > Umax) continue;                         // invalid length: the lattice pass reports the utterance infeasible
    const int U = (int)U64;
    const T* x = logits + b * stride_b + t * stride_t;
    const int64_t* tg = targets + (int64_t)b * ldt;
    int64_t lab[NL];                                             // requested before the V-wide pass: their latency hides under it
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const int idx = lane + 32 * k;                             // 0 = blank, 1 + u = label u
      lab[k] = (idx >= 1 && idx <= U) ? tg[idx - 1] : blank;
    }
    const float l = warp_row_lse<T>(x, V, lane);
    if (lane == 0) lse[row] = l;
    uint32_t* out = lplat + (int64_t)b * Tn * LP + (int64_t)t * lin_row_pitch(U);   // the pitch the recursion's shared-memory rows have
    float e[NL];
    float c = -INFINITY;
    // branch-free: the gathers (L1 hits: the row has just been read) are requested together, absent entries and labels
    // outside the vocabulary (probability zero) selected away afterwards
    float xv[NL];
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const bool ok = lane + 32 * k <= U && lab[k] >= 0 && lab[k] < V;
      xv[k] = ld_f(x + (ok ? lab[k] : 0));
    }
#pragma unroll
    for (int k = 0; k < NL; ++k) {
      const bool ok = lane + 32 * k <= U && lab[k] >= 0 && lab[k] < V;
      e[k] = ok ? (xv[k] - l) * 1.4426950408889634f : -INFINITY;
      c = fmaxf(c, e[k]);
    }
    c = warp_max(c);
    if (!(c > -1e29f)) c = 0.f;                                  // every lattice emission is -inf: leave the row dead
#pragma unroll
    for (int k = 0; k < NL; ++k)
      if (lane + 32 * k <= U) out[lane + 32 * k] = lin_word_of_log2(e[k] - c);
    if (lane == 0) { cshift[row] = c; out[U + 1] = 0u; }        // the row's zero word: what a missing label reads (writing the row's padding as well, so that whole sectors leave: 0.156 vs 0.152 ms — not kept)
  }
}

// ---- pass 2, linear domain ----------------------------------------------------------------------------------
// K = pairs per lane for THIS utterance (ceil((U+1)/32), chosen at run time by the kernel below).  Lane i owns the
// K CONSECUTIVE pairs c = K i + j of the scan order: a pair's label needs the previous pair's label of the
// PREVIOUS column, which for j > 0 is the lane's own register — one lane rotation per step (the previous lane's last
// label) feeds the whole lane, every other operand is already there.
// One warp alone on its scheduler pays every latency it exposes (measured: ~3 cycles per instruction however the
// step is arranged, a taken branch ~20), so the recursion warp issues nothing but the recursion: LIN_ROWS steps are
// unrolled into one branch-free group; the lane's LAST pair is computed first and its label rotated at once (the
// shuffle for the next step is in flight under the other pairs' DP work); emission words come from a shared-memory
// block with a COMPILE-TIME pitch (immediate offsets, no address arithmetic) and are turned into doubles one step
// ahead; the row goes to a staging buffer as plain 4-byte shared-memory stores of the registers that already hold
// the high words.  Everything else is a SECOND warp's job (warp specialisation, mbarrier hand-offs): it lands the
// emission rows (bulk async copies, row by row into the compile-time pitch) and copies finished staging batches to
// global memory with coalesced 16-byte stores.
// Per step and recursion lane: 5K DP instructions, K+1 LDS, K+1 moves, 2K STS, 2 shuffles, 2 selects.
// Measured on a B200 at cfg2 (B=64, T=3000, U<=150), lattice kernel alone (r02): log-domain pair-per-thread wavefront
// 455 us; fp64 single warp with interleaved pairs and direct 8-byte stores 410; consecutive pairs + bulk stores from
// the recursion warp 445-523; this warp-specialised form 360.  Tried and dropped: FOUR recursion warps (one per
// scheduler, 1-2 pairs per lane, last label handed from warp to warp through a polled shared-memory mailbox, column
// scale kept common without a barrier by acting on maxima posted one check point earlier) — bit-identical results,
// 620-800 us: the per-step mailbox round trip costs more than the 3/4 of the DP work it takes off a warp.
constexpr int LIN_ROWS = 8;                // steps per group = rows per staging batch
constexpr int LIN_NB = 8;                  // staging batches in the ring (power of two; LIN_NB_WIDE for the widest lattices, whose rows are larger).  Two were enough alone (0.39 ms) but a batch
                                           // is ~1 us of recursion: next to a kernel that loads the memory system the bulk stores take
                                           // longer than that to read their rows and the recursion waited for its buffer (measured, cfg2:
                                           // 0.58 ms beside a device copy, 0.56 beside the emission pass, 0.40 beside an issue-bound
                                           // kernel on an L2-resident tensor — profiles/r02_ctc_interference.txt)
constexpr int LIN_THREADS = 64;            // warp 0: recursion, warp 1: I/O

__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }

constexpr int LIN_NB_WIDE = 4;
struct LinBars {                            // shared-memory mbarriers of one CTA
  uint64_t efull[2], eempty[2];             // emission block landed / consumed
  uint64_t sfull[LIN_NB], sempty[LIN_NB];   // staging batch written / copied out
  int meta[LIN_NB][2];                      // per staging batch: rows, first frame
};

template <int K, int DIR>
struct Lin64 {
  static constexpr int SP = 64 * K;        // staging row pitch in words (>= 2*32*K)
  static constexpr int EPW = 32 * K + LIN_EPAD;   // emission row pitch in shared (and global) memory, words
  double bv[K], lv[K];
  double nbv;                              // previous lane's last label of the previous column (0 for lane 0)
  double skipf[K];
  int lane;
  int src;

  // emission doubles of the row at shared address `ea` (+ the lane's label offsets, already folded into eb[])
  __device__ __forceinline__ void load_em(const uint32_t (&eb)[K + 1], int rowoff, double (&p)[K + 1]) const {
#pragma unroll
    for (int j = 0; j <= K; ++j) p[j] = lin_hi2d(lds32(eb[j] + (uint32_t)rowoff));
  }
  // one column: consumes the emission doubles p (p[K] = blank), writes the lane's 2K row words at shared address sa
  __device__ __forceinline__ void step(const double (&p)[K + 1], uint32_t sa) {
    double nb_next = 0.0;
#pragma unroll
    for (int j = K - 1; j >= 0; --j) {                           // downwards: lv[j-1] is still the previous column's
      const double prev = j > 0 ? lv[j - 1] : nbv;
      const double sb = bv[j] + prev;
      const double sl = fma(skipf[j], prev, lv[j] + bv[j]);
      bv[j] = sb * p[K];
      lv[j] = sl * p[j];
      if (j == K - 1) nb_next = lin_rot(lv[K - 1], src);         // in flight under the remaining pairs
      sts32(sa + 8 * j, (uint32_t)lin_hi(DIR == 0 ? bv[j] : sb));       // beta leaves without its frame's emission
      sts32(sa + 8 * j + 4, (uint32_t)lin_hi(DIR == 0 ? lv[j] : sl));
    }
    nbv = lane == 0 ? 0.0 : nb_next;
  }
};

// group boundaries of the scan (both warps walk the same sequence): groups end where the range check sits — alpha
// after frames 7 mod 8, beta after frames 0 mod 8 — and never cross an emission block.
template <int DIR>
__device__ __forceinline__ int lin_group_len(int t, int rows) {
  int n = DIR == 0 ? LIN_ROWS - (t & (LIN_ROWS - 1)) : (t & (LIN_ROWS - 1)) + 1;
  return n < rows ? n : rows;
}

template <int K, int DIR, int NB>
__device__ __forceinline__ void
ctc_lin64_recursion(uint32_t* __restrict__ ebuf, uint32_t* __restrict__ stage, LinBars* bars,
                    const int64_t* __restrict__ tg, int Tb, int U, int lo, int hi,
                    int* __restrict__ accrec, double* __restrict__ zl2_out, int* __restrict__ danger_out,
                    double* __restrict__ state, int* __restrict__ istate) {
  // this launch walks frames [lo, hi) of the direction (lo a multiple of LIN_EB); the column is taken from / left in
  // `state` when the range does not start / end at the direction's first / last frame
  const bool first = DIR == 0 ? lo == 0 : hi == Tb;
  const bool last = DIR == 0 ? hi == Tb : lo == 0;
  constexpr int SP = Lin64<K, DIR>::SP, EPW = Lin64<K, DIR>::EPW;
  Lin64<K, DIR> L;
  const int lane = threadIdx.x;
  L.lane = lane;
  L.src = (lane + 31) & 31;
  uint32_t eoff[K + 1];                                          // byte offset of the lane's emission words inside a row
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int c = K * lane + j;
    const bool vl = c < U;
    const int u = DIR == 0 ? c : U - 1 - c;                      // label index in the transcript
    eoff[j] = 4u * (uint32_t)(vl ? 1 + u : U + 1);               // U+1: the row's zero word
    bool sk = false;
    if (vl && c >= 1) sk = tg[u] != tg[DIR == 0 ? u - 1 : u + 1];
    L.skipf[j] = sk ? 1.0 : 0.0;                                 // 1.0 when the label may also be entered from the previous pair's label
    L.bv[j] = 0.0; L.lv[j] = 0.0;
  }
  eoff[K] = 0u;                                                  // the blank's word
  L.nbv = 0.0;
  int acc = 1023 - LIN_TGT;                                      // log2 of the scale taken out of the column so far
  int danger = 0;
  if (first) {
    if (lane == 0) L.bv[0] = __hiloint2double(LIN_TGT << 20, 0); // virtual column before the first frame: all mass in front of node 0, at the target scale
  } else {
#pragma unroll
    for (int j = 0; j < K; ++j) { L.bv[j] = state[(2 * j) * 32 + lane]; L.lv[j] = state[(2 * j + 1) * 32 + lane]; }
    L.nbv = state[(2 * K) * 32 + lane];
    acc = istate[0]; danger = istate[1];
  }
  const int blk_lo = lo / LIN_EB, blk_hi = (hi - 1) / LIN_EB;
  const int nvis = blk_hi - blk_lo + 1;
  int t = DIR == 0 ? lo : hi - 1;
  if (first && lane == 0) accrec[t >> 3] = acc;                  // scale of the rows up to the first check
  const uint32_t ebuf_a = smem_u32(ebuf), stage_a = smem_u32(stage);
  int g = 0;                                                     // groups done
  for (int vi = 0; vi < nvis; ++vi) {
    mbar_wait(smem_u32(&bars->efull[vi & 1]), (uint32_t)((vi >> 1) & 1));
    const int blk = DIR == 0 ? blk_lo + vi : blk_hi - vi;
    const int r0 = max(lo, blk * LIN_EB) - blk * LIN_EB, r1 = min(hi, blk * LIN_EB + LIN_EB) - blk * LIN_EB;
    int rows = r1 - r0;
    int row = DIR == 0 ? r0 : r1 - 1;                            // row of the block the next step reads
    while (rows > 0) {
      const int n = lin_group_len<DIR>(t, rows);
      const bool aligned = DIR == 0 ? ((t + n - 1) & 7) == 7 : ((t - n + 1) & 7) == 0;
      const int batch = g & (NB - 1);
      if (g >= NB) mbar_wait(smem_u32(&bars->sempty[batch]), (uint32_t)(((g / NB) - 1) & 1));   // the copy engine has read this buffer's previous rows
      uint32_t eb[K + 1];
      const uint32_t rowbase = ebuf_a + 4u * (uint32_t)(((vi & 1) * LIN_EB + row) * EPW);
#pragma unroll
      for (int j = 0; j <= K; ++j) eb[j] = rowbase + eoff[j];
      const uint32_t sa = stage_a + 4u * (uint32_t)(batch * LIN_ROWS * SP + 2 * K * lane);
      constexpr int ESTEP = (DIR == 0 ? 4 : -4) * EPW;
      double p[K + 1], pn[K + 1];
      L.load_em(eb, 0, p);
      if (n == LIN_ROWS) {
#pragma unroll
        for (int s2 = 0; s2 < LIN_ROWS; ++s2) {
          if (s2 + 1 < LIN_ROWS) L.load_em(eb, (s2 + 1) * ESTEP, pn);    // one step ahead
          L.step(p, sa + 4u * (uint32_t)(s2 * SP));
#pragma unroll
          for (int j = 0; j <= K; ++j) p[j] = pn[j];
        }
      } else {
        for (int s2 = 0; s2 < n; ++s2) {
          if (s2 + 1 < n) L.load_em(eb, (s2 + 1) * ESTEP, pn);
          L.step(p, sa + 4u * (uint32_t)(s2 * SP));
#pragma unroll
          for (int j = 0; j <= K; ++j) p[j] = pn[j];
        }
      }
      __syncwarp();
      if (lane == 0) {
        bars->meta[batch][0] = n; bars->meta[batch][1] = t;
        mbar_arrive(smem_u32(&bars->sfull[batch]));              // release: the batch and its meta are visible to the I/O warp
      }
      if (aligned) {
        int m = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) m = max(m, max(lin_hi(L.bv[j]), lin_hi(L.lv[j])));
        m = __reduce_max_sync(0xffffffffu, m);
        const int e = m >> 20;
        if (e > 0) {
          if (e < LIN_DANGER) danger = 1;
          const int d = e - LIN_TGT;
          if (d > LIN_HYST || d < -LIN_HYST) {
            auto rescale = [&](double v) -> double {
              const int h = lin_hi(v);
              return ((h >> 20) > 0 && (h >> 20) - d > 0) ? __hiloint2double(h - (d << 20), __double2loint(v)) : 0.0;
            };
#pragma unroll
            for (int j = 0; j < K; ++j) { L.bv[j] = rescale(L.bv[j]); L.lv[j] = rescale(L.lv[j]); }
            L.nbv = rescale(L.nbv);
            acc += d;
          }
        }
        // rows of the next 8 frames in scan order carry this scale (alpha: frames t+1..t+8, beta: t-8..t-1)
        const int tl = DIR == 0 ? t + n - 1 : t - n + 1;
        if (lane == 0) { const int k = DIR == 0 ? (tl >> 3) + 1 : (tl >> 3) - 1; if (k >= 0) accrec[k] = acc; }
      }
      row += DIR == 0 ? n : -n;
      t += DIR == 0 ? n : -n;
      rows -= n;
      ++g;
    }
    __syncwarp();                                               // every lane is done with this visit's emission block
    if (lane == 0) mbar_arrive(smem_u32(&bars->eempty[vi & 1]));
  }
  if (!last) {                                                   // leave the column for the next launch
#pragma unroll
    for (int j = 0; j < K; ++j) { state[(2 * j) * 32 + lane] = L.bv[j]; state[(2 * j + 1) * 32 + lane] = L.lv[j]; }
    state[(2 * K) * 32 + lane] = L.nbv;
    if (lane == 0) { istate[0] = acc; istate[1] = danger; }
    return;
  }
  // likelihood from this direction: the last pair's blank and the label before it (both after their emission)
  double z = 0.0;
#pragma unroll
  for (int j = 0; j < K; ++j) {
    const int c = K * lane + j;
    if (c == U) z += L.bv[j];
    if (c == U - 1) z += L.lv[j];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) z += lin_rot(z, lane ^ o);
  if (lane == 0) {
    *zl2_out = z > 0.0 ? log2(z) + (double)acc : -INFINITY;
    *danger_out = danger;
  }
}

// the I/O warp: emission rows in (bulk async copies into the compile-time pitch), finished staging batches out
template <int DIR, int NB>
__device__ __forceinline__ void
ctc_lin64_io(uint32_t* __restrict__ ebuf, uint32_t* __restrict__ stage, LinBars* bars, const uint32_t* __restrict__ lp_b,
             int U, int lo, int hi, int K, int LP, int Smax, float* __restrict__ out_b) {
  const int lane = threadIdx.x & 31;
  const int EPW = 32 * K + LIN_EPAD, SP = 64 * K;
  const int blk_lo = lo / LIN_EB, blk_hi = (hi - 1) / LIN_EB;
  const int nvis = blk_hi - blk_lo + 1;
  const int rowchunks = (2 * U + 2 + 3) >> 2;                          // 16-byte chunks of an output row (nodes 0..2U, + the empty label slot)
  // The emission rows of an utterance lie in global memory with the SAME pitch as in shared memory (EPW words: the
  // emission pass writes them that way), so a block of up to LIN_EB rows is ONE bulk copy.  r02: issued row by row
  // (64 copies per block from one thread, ~100 ns each) the copies of block k+2 held up the stores of block k+1's
  // batches behind them until the staging ring was full: 0.36 -> 0.27 ms at cfg2, 0.30 -> 0.19 ms at U = 31.
  auto issue = [&](int vi) {                                           // lane 0 only
    const int blk = DIR == 0 ? blk_lo + vi : blk_hi - vi;
    const int r0 = max(lo, blk * LIN_EB) - blk * LIN_EB, r1 = min(hi, blk * LIN_EB + LIN_EB) - blk * LIN_EB;
    const uint32_t bar = smem_u32(&bars->efull[vi & 1]);
    const uint32_t bytes = (uint32_t)((r1 - r0) * EPW) * 4u;
    mbar_expect_tx(bar, bytes);
    bulk_load_1d(smem_u32(ebuf + (size_t)(vi & 1) * LIN_EB * EPW) + 4u * (uint32_t)(r0 * EPW),
                 lp_b + ((int64_t)blk * LIN_EB + r0) * EPW, bytes, bar);
  };
  if (lane == 0) {
    issue(0);
    if (nvis > 1) issue(1);
  }
  int t = DIR == 0 ? lo : hi - 1;
  int g = 0;
  for (int vi = 0; vi < nvis; ++vi) {
    const int blk = DIR == 0 ? blk_lo + vi : blk_hi - vi;
    int rows = min(hi, blk * LIN_EB + LIN_EB) - max(lo, blk * LIN_EB);
    while (rows > 0) {
      const int n = lin_group_len<DIR>(t, rows);
      const int batch = g & (NB - 1);
      mbar_wait(smem_u32(&bars->sfull[batch]), (uint32_t)((g / NB) & 1));
      if (lane == 0) {
        // the generic-proxy stores of the recursion warp -> the async proxy that reads them
        fence_async_smem();
        for (int r = 0; r < n; ++r)
          bulk_store_1d(out_b + (int64_t)(DIR == 0 ? t + r : t - r) * Smax,
                        smem_u32(stage + ((size_t)batch * LIN_ROWS + r) * SP), (uint32_t)rowchunks * 16u);
        bulk_commit();
        if (g >= NB - 2) {                                       // all but the NB-2 youngest batches have been read: free the oldest
          bulk_wait_read<NB - 2>();
          mbar_arrive(smem_u32(&bars->sempty[(g - (NB - 2)) & (NB - 1)]));
        }
      }
      t += DIR == 0 ? n : -n;
      rows -= n;
      ++g;
    }
    if (vi + 2 < nvis) {
      mbar_wait(smem_u32(&bars->eempty[vi & 1]), (uint32_t)((vi >> 1) & 1));   // the recursion warp has left this emission block
      if (lane == 0) issue(vi + 2);
    }
  }
  if (lane == 0) bulk_wait0();                                   // every row is written before the CTA retires
}

template <int DIR, int NB>
__device__ __forceinline__ void
ctc_lin64_body(uint32_t* __restrict__ lin_sm, LinBars* bars, const uint32_t* __restrict__ lplat,
               const int64_t* __restrict__ targets, int64_t ldt,
               const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
               int Tn, int Umax, int LP, int Smax, float* __restrict__ alpha, float* __restrict__ beta,
               float* __restrict__ nll, CtcWs ws, int f0, int f1) {
  // f0, f1: the frames [f0, f1) this launch covers in this direction (whole segment: 0, Tn)
  const int b = blockIdx.x, tid = threadIdx.x;
  int64_t Tb64 = in_lens[b]; if (Tb64 > Tn) Tb64 = Tn;
  const int Tb = (int)Tb64;
  const int64_t U64 = tgt_lens[b];
  const bool bad_len = U64 < 0 || U64 > Umax;
  const int U = bad_len ? 0 : (int)U64;
  if (Tb <= 0 || bad_len) {
    if (tid == 0 && (DIR == 0 ? f0 == 0 : f1 >= Tn)) {           // once, in the launch that holds the direction's first frame
      const bool ok = !bad_len && U == 0;                        // no frames, no labels: probability one
      ws.danger[2 * b + DIR] = 0;
      ws.zl2[2 * b + DIR] = ok ? 0.0 : -INFINITY;
      if (DIR == 0) { ws.lossy[b] = 0; nll[b] = ok ? 0.f : INFINITY; }
    }
    return;
  }
  const int lo = f0 < 0 ? 0 : f0, hi = f1 < Tb ? f1 : Tb;
  if (lo >= hi) return;                                          // none of this utterance's frames in the range
  if (tid == 0 && (DIR == 0 ? lo == 0 : hi == Tb)) { ws.danger[2 * b + DIR] = 0; if (DIR == 0) ws.lossy[b] = 0; }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&bars->efull[i]), 1); mbar_init(smem_u32(&bars->eempty[i]), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(smem_u32(&bars->sfull[i]), 1); mbar_init(smem_u32(&bars->sempty[i]), 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int K = (U + 32) >> 5;                                   // pairs per lane for this transcript
  const int Kmax = (Umax + 32) >> 5;
  uint32_t* ebuf = lin_sm;                                       // 2 emission blocks of LIN_EB rows, pitch 32K+LIN_EPAD words
  uint32_t* stage = lin_sm + 2 * (size_t)LIN_EB * (32 * Kmax + LIN_EPAD);   // LIN_NB staging batches of LIN_ROWS rows, pitch 64K words
  const uint32_t* lp_b = lplat + (int64_t)b * Tn * LP;           // the utterance's region (LP words per frame reserved); rows inside at pitch 32K+4
  float* out_b = (DIR == 0 ? alpha : beta) + (int64_t)b * Tn * Smax;
  if (tid >= 32) {
    ctc_lin64_io<DIR, NB>(ebuf, stage, bars, lp_b, U, lo, hi, K, LP, Smax, out_b);
    return;
  }
  const int64_t* tg = targets + (int64_t)b * ldt;
  int* accrec = ws.accs + (size_t)(2 * b + DIR) * ws.NT;
  double* zo = ws.zl2 + 2 * b + DIR;
  int* dg = ws.danger + 2 * b + DIR;
  double* state = ws.state + (size_t)(2 * b + DIR) * (2 * LIN_MAXK + 1) * 32;
  int* istate = ws.istate + (size_t)(2 * b + DIR) * 2;
#define SC_LIN_RUN(KK) ctc_lin64_recursion<KK, DIR, NB>(ebuf, stage, bars, tg, Tb, U, lo, hi, accrec, zo, dg, state, istate)
  switch (K) {
    case 1: SC_LIN_RUN(1); break;
    case 2: SC_LIN_RUN(2); break;
    case 3: SC_LIN_RUN(3); break;
    case 4: SC_LIN_RUN(4); break;
    case 5: SC_LIN_RUN(5); break;
    case 6: SC_LIN_RUN(6); break;
    case 7: SC_LIN_RUN(7); break;
    default: SC_LIN_RUN(8); break;
  }
#undef SC_LIN_RUN
}

template <int NB>
__global__ void __launch_bounds__(LIN_THREADS, 1)
ctc_lin64_kernel(const uint32_t* __restrict__ lplat, const int64_t* __restrict__ targets, int64_t ldt,
                 const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                 int Tn, int Umax, int LP, int Smax, float* __restrict__ alpha, float* __restrict__ beta,
                 float* __restrict__ nll, CtcWs ws, int a0, int a1, int b0, int b1) {
  // alpha walks frames [a0, a1) upwards, beta frames [b0, b1) downwards (a0, b0 multiples of LIN_EB; the whole segment: 0, Tn)
  extern __shared__ __align__(128) uint32_t lin_sm[];
  __shared__ __align__(8) LinBars bars;
  if (blockIdx.y == 0) ctc_lin64_body<0, NB>(lin_sm, &bars, lplat, targets, ldt, in_lens, tgt_lens, Tn, Umax, LP, Smax, alpha, beta, nll, ws, a0, a1);
  else ctc_lin64_body<1, NB>(lin_sm, &bars, lplat, targets, ldt, in_lens, tgt_lens, Tn, Umax, LP, Smax, alpha, beta, nll, ws, b0, b1);
}

// likelihood, and whether fp64 held everything (see the header of this file).  grid (B, 1 + samples/4): block
// (b, 0) does the per-utterance part (a), (b) and writes nll[b]; the warps of the other blocks take one sampled
// frame each for (c).  lossy[b] was zeroed by the lattice kernel; problems are OR-ed in.
constexpr int LIN_CHECK_THREADS = 128;
__global__ void __launch_bounds__(LIN_CHECK_THREADS)
ctc_lin64_check_kernel(const float* __restrict__ cshift, const int64_t* __restrict__ in_lens,
                       const int64_t* __restrict__ tgt_lens, int Tn, int Umax, int Smax, int force_lossy,
                       const float* __restrict__ alpha, const float* __restrict__ beta,
                       float* __restrict__ nll, CtcWs ws) {
  __shared__ double dred[LIN_CHECK_THREADS / 32];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int64_t Tb64 = in_lens[b]; if (Tb64 > Tn) Tb64 = Tn;
  const int Tb = (int)Tb64;
  const int64_t U64 = tgt_lens[b];
  if (Tb <= 0 || U64 < 0 || U64 > Umax) return;                  // nll already final, lossy[b] = 0
  const int U = (int)U64;
  const double za = ws.zl2[2 * b], zb = ws.zl2[2 * b + 1];
  const bool fa = za > -INFINITY, fb = zb > -INFINITY;
  if (blockIdx.y == 0) {
    double ssum = 0.0;
    const float* cs = cshift + (int64_t)b * Tn;
    for (int t = threadIdx.x; t < Tb; t += LIN_CHECK_THREADS) ssum += (double)cs[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ssum += lin_rot(ssum, lane ^ o);
    if (lane == 0) dred[warp] = ssum;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int w = 0; w < LIN_CHECK_THREADS / 32; ++w) tot += dred[w];
      int lossy = (ws.danger[2 * b] | ws.danger[2 * b + 1] | force_lossy) ? 1 : 0;       // (a)
      if (fa != fb || (fa && fabs(za - zb) > 1e-6)) lossy = 1;                            // (b) both directions must see the same likelihood
      if (lossy) atomicOr(ws.lossy + b, 1);
      nll[b] = fa ? (float)(-(za + tot) * 0.6931471805599453) : INFINITY;                 // recomputed later if the utterance ends up flagged
    }
    return;
  }
  // (c) sum_s alpha_t(s) beta_t(s) at a sampled frame, from the stored words
  if (!(fa && fb)) return;
  const int t = LIN_SAMPLE / 2 - 1 + LIN_SAMPLE * ((blockIdx.y - 1) * (LIN_CHECK_THREADS / 32) + warp);
  if (t >= Tb) return;
  const uint2* aw = reinterpret_cast<const uint2*>(alpha + ((int64_t)b * Tn + t) * Smax);
  const uint32_t* bw = reinterpret_cast<const uint32_t*>(beta + ((int64_t)b * Tn + t) * Smax);
  constexpr int NPC = LIN_MAXK;                                  // pairs per lane (U <= 32*LIN_MAXK - 1)
  uint32_t wa[NPC], wl[NPC], xb[NPC], xl[NPC];
  int emax = -1;
#pragma unroll
  for (int k = 0; k < NPC; ++k) {
    const int u = lane + 32 * k;
    wa[k] = 0u; wl[k] = 0u; xb[k] = 0u; xl[k] = 0u;
    if (u <= U) {
      const uint2 a = aw[u];
      wa[k] = a.x; wl[k] = a.y;
      xb[k] = bw[2 * (U - u)];
      if (u < U) xl[k] = bw[2 * (U - u) - 1];
      if (wa[k] != 0u && xb[k] != 0u) emax = max(emax, (int)(wa[k] >> 20) + (int)(xb[k] >> 20));
      if (wl[k] != 0u && xl[k] != 0u) emax = max(emax, (int)(wl[k] >> 20) + (int)(xl[k] >> 20));
    }
  }
  emax = __reduce_max_sync(0xffffffffu, emax);
  auto term = [&](uint32_t x, uint32_t y) -> double {
    if (x == 0u || y == 0u) return 0.0;
    const int d = (int)(x >> 20) + (int)(y >> 20) - emax;
    if (d < -60) return 0.0;
    return __hiloint2double((int)(0x3ff00000u | (x & 0xfffffu)), 0) * __hiloint2double((int)(0x3ff00000u | (y & 0xfffffu)), 0) *
           __hiloint2double((1023 + d) << 20, 0);
  };
  double acc = 0.0;
#pragma unroll
  for (int k = 0; k < NPC; ++k) acc += term(wa[k], xb[k]) + term(wl[k], xl[k]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += lin_rot(acc, lane ^ o);
  const int* aa = ws.accs + (size_t)(2 * b) * ws.NT;
  const int* ab = ws.accs + (size_t)(2 * b + 1) * ws.NT;
  const double zt = emax >= 0 ? log2(acc) + (double)(emax - 2046) + (double)aa[t >> 3] + (double)ab[t >> 3] : -INFINITY;
  if (lane == 0 && !(fabs(zt - za) <= 1e-4)) atomicOr(ws.lossy + b, 1);
}

}  // namespace sc
