// K4: RNN-T (transducer) loss and gradient — anti-diagonal wavefront alpha/beta (sm_100a).
//
// Replaces the external `warp_rnnt.RNNTLoss` the reference calls at model.py:97-105 with
// gather=True (only the blank and the label log-prob of every lattice node matter).  The
// reference call itself is unpinned; the algorithm is Graves 2012,
// restated in oracle/rnnt_oracle.py and cross-checked against torchaudio.
//
// Layout: every per-node array is stored SKEWED, [B][D = T+U1][U1p], row d = t+u holding the
// nodes of one anti-diagonal, so that the wavefront reads and writes one contiguous row per
// step (coalesced, and bulk-copyable):
//   eb[d][u] = log_probs[t,u,blank],  el[d][u] = log_probs[t,u,label_{u+1}]   (t = d-u)
// Passes:  (1) rnnt_gather: one thread per node pulls its two log-probs out of the V-wide row;
// (2) rnnt_alpha_beta: one CTA per (utterance, direction), node column u on thread u, the
// previous diagonal in a double-buffered shared-memory line, one __syncthreads per
// diagonal, emission rows landed by cp.async.bulk in blocks of RNNT_EB diagonals;
// (3) rnnt_grad: one thread per node scatters the two non-zero gradient entries.
#include "sc_common.cuh"
#include "sc_tma.cuh"

namespace sc {

// label count of utterance b, held inside the lattice the caller allocated (a length the host never saw — lengths may be
// device tensors — must not index past U1-1 columns; torch / warp_rnnt raise on the host for such input)
__device__ __forceinline__ int64_t rnnt_label_len(const int64_t* __restrict__ label_lens, int b, int U1) {
  const int64_t u = label_lens[b];
  return u < 0 ? 0 : (u > U1 - 1 ? U1 - 1 : u);
}

#define NEG_INF (-INFINITY)
constexpr int RNNT_EB = 16;

__device__ __forceinline__ float lse2n(float a, float b) {
  const float m = fmaxf(fmaxf(a, b), -1e30f);
  return m + __logf(__expf(a - m) + __expf(b - m));
}

__global__ void rnnt_gather_kernel(const float* __restrict__ lp, const int64_t* __restrict__ labels, int64_t ldl,
                                   const int64_t* __restrict__ frame_lens, const int64_t* __restrict__ label_lens,
                                   int B, int Tn, int U1, int V, int U1p, int64_t blank,
                                   const int64_t* __restrict__ row_offsets,
                                   float* __restrict__ eb, float* __restrict__ el) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * Tn * U1) return;
  const int u = (int)(i % U1);
  const int t = (int)((i / U1) % Tn);
  const int b = (int)(i / ((int64_t)U1 * Tn));
  int64_t Tb = frame_lens[b]; if (Tb > Tn) Tb = Tn;
  const int64_t Ub = rnnt_label_len(label_lens, b, U1);
  if (t >= Tb || u > Ub) return;
  // padded (B,T,U1,V) layout, or the compact packing of model.py:147-200: utterance b starts at
  // row_offsets[b] and holds T_b x (U_b+1) rows
  const int64_t r = row_offsets ? row_offsets[b] + (int64_t)t * (Ub + 1) + u : i;
  const float* row = lp + r * V;
  const int64_t D = Tn + U1;
  const int64_t o = ((int64_t)b * D + (t + u)) * U1p + u;
  eb[o] = __ldg(row + blank);
  el[o] = (u < Ub) ? __ldg(row + labels[(int64_t)b * ldl + u]) : NEG_INF;
}

// dir 0: alpha over diagonals 0..Tb+Ub-1;  dir 1: beta over diagonals Tb+Ub-1..0
__global__ void rnnt_alpha_beta_kernel(const float* __restrict__ eb, const float* __restrict__ el,
                                       const int64_t* __restrict__ frame_lens,
                                       const int64_t* __restrict__ label_lens, int Tn, int U1, int U1p,
                                       float* __restrict__ alpha, float* __restrict__ beta,
                                       float* __restrict__ nll) {
  extern __shared__ __align__(128) float sm[];   // 2 lines of (U1p + 4) + 2 x 2 blocks of RNNT_EB emission rows
  __shared__ __align__(8) uint64_t ebar[2];
  const int b = blockIdx.x, dir = blockIdx.y;
  int64_t Tb64 = frame_lens[b]; if (Tb64 > Tn) Tb64 = Tn;
  const int Tb = (int)Tb64, Ub = (int)rnnt_label_len(label_lens, b, U1);
  if (Tb <= 0) {
    if (dir == 0 && threadIdx.x == 0) nll[b] = 0.f;        // no frames: zero loss, zero gradient
    return;
  }
  const int LINE = U1p + 4;
  const int64_t D = Tn + U1;
  const int nd = Tb + Ub;                                    // live diagonals 0..nd-1
  const int u = threadIdx.x;
  for (int i = threadIdx.x; i < 2 * LINE; i += blockDim.x) sm[i] = NEG_INF;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&ebar[0]), 1);
    mbar_init(smem_u32(&ebar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float* lineA = sm + 2;
  float* lineB = sm + LINE + 2;
  float* stage = sm + 2 * LINE;                              // [2 buffers][2 arrays][RNNT_EB][U1p]
  const float* eb_b = eb + (int64_t)b * D * U1p;
  const float* el_b = el + (int64_t)b * D * U1p;
  float* out_b = (dir == 0 ? alpha : beta) + (int64_t)b * D * U1p;

  // emission rows are consumed in scan order in blocks of RNNT_EB diagonals
  const int nvis = (nd + RNNT_EB - 1) / RNNT_EB;
  auto blk_of = [&](int vi) { return dir == 0 ? vi : nvis - 1 - vi; };
  auto rows_of = [&](int blk) { const int r = nd - blk * RNNT_EB; return r < RNNT_EB ? r : RNNT_EB; };
  auto issue = [&](int vi) {
    const int blk = blk_of(vi);
    const uint32_t bytes = (uint32_t)rows_of(blk) * (uint32_t)U1p * 4u;
    const uint32_t bar = smem_u32(&ebar[vi & 1]);
    mbar_expect_tx(bar, 2 * bytes);
    float* dst = stage + (size_t)(vi & 1) * 2 * RNNT_EB * U1p;
    bulk_load_1d(smem_u32(dst), eb_b + (int64_t)blk * RNNT_EB * U1p, bytes, bar);
    bulk_load_1d(smem_u32(dst + RNNT_EB * U1p), el_b + (int64_t)blk * RNNT_EB * U1p, bytes, bar);
  };
  if (threadIdx.x == 0) {
    issue(0);
    if (nvis > 1) issue(1);
  }
  // row access by absolute diagonal index (must lie in the current or the previous visit's block)
  int vi = 0;
  mbar_wait(smem_u32(&ebar[0]), 0);
  auto row_ptr = [&](int d, int arr, int v) -> const float* {
    return stage + ((size_t)(v & 1) * 2 + arr) * RNNT_EB * U1p + (size_t)(d % RNNT_EB) * U1p;
  };
  const bool col_live = u <= Ub;
  float own = NEG_INF;                                       // this column's value on the previous diagonal
  float* prev = lineA;
  float* cur = lineB;
  if (dir == 0) {
    // alpha(t,u) = lse(alpha(t-1,u) + eb[t-1,u], alpha(t,u-1) + el[t,u-1]); both emissions sit on row d-1
    for (int d = 0; d < nd; ++d) {
      if (d >= 1 && (d - 1) % RNNT_EB == 0) {
        // step d reads row d-1, the first row of block k.  Everyone has passed the barrier of
        // step d-1 (last reader of block k-1), so that buffer can take block k+1.
        const int k = (d - 1) / RNNT_EB;
        if (k >= 1 && threadIdx.x == 0 && k + 1 < nvis) issue(k + 1);
        mbar_wait(smem_u32(&ebar[k & 1]), (uint32_t)((k >> 1) & 1));
        vi = k;
      }
      const int t = d - u;
      float v = NEG_INF;
      if (col_live && t >= 0 && t < Tb) {
        if (d == 0) {
          v = 0.f;
        } else {
          // selects, not arithmetic masking: slots of non-existent nodes are uninitialised
          const float down = (t > 0) ? own + row_ptr(d - 1, 0, vi)[u] : NEG_INF;
          const float left = (u > 0) ? prev[u - 1] + row_ptr(d - 1, 1, vi)[u - 1] : NEG_INF;
          v = lse2n(down, left);
        }
      }
      if (u < U1p) { cur[u] = v; out_b[(int64_t)d * U1p + u] = v; }
      own = v;
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    }
    if (threadIdx.x == 0) {
      // final node (Tb-1, Ub) lives on diagonal nd-1, column Ub: prev holds that diagonal
      const float a = prev[Ub];
      const float e = eb_b[(int64_t)(nd - 1) * U1p + Ub];
      nll[b] = -(a + e);
    }
  } else {
    // beta(t,u) = lse(beta(t+1,u) + eb[t,u], beta(t,u+1) + el[t,u]); emissions sit on the node's own row d
    for (int d = nd - 1; d >= 0; --d) {
      const int blk = d / RNNT_EB;
      const int v_of_blk = nvis - 1 - blk;
      if (v_of_blk != vi) {                                  // crossed into an older block of rows
        // the buffer of visit vi (block blk+1) is free after the barrier that ended step d+1
        if (threadIdx.x == 0 && vi + 2 < nvis) issue(vi + 2);
        vi = v_of_blk;
        mbar_wait(smem_u32(&ebar[vi & 1]), (uint32_t)((vi >> 1) & 1));
      }
      const int t = d - u;
      float v = NEG_INF;
      if (col_live && t >= 0 && t < Tb) {
        const float e_b = row_ptr(d, 0, vi)[u];
        if (t == Tb - 1 && u == Ub) {
          v = e_b;
        } else {
          const float e_l = row_ptr(d, 1, vi)[u];
          const float right = (u < Ub) ? prev[u + 1] + e_l : NEG_INF;
          const float down = (t + 1 < Tb) ? own + e_b : NEG_INF;
          v = lse2n(down, right);
        }
      }
      if (u < U1p) { cur[u] = v; out_b[(int64_t)d * U1p + u] = v; }
      own = v;
      __syncthreads();
      float* tmp = prev; prev = cur; cur = tmp;
    }
  }
}

__global__ void rnnt_grad_kernel(const float* __restrict__ eb, const float* __restrict__ el,
                                 const float* __restrict__ alpha, const float* __restrict__ beta,
                                 const float* __restrict__ nll, const float* __restrict__ grad_w,
                                 const int64_t* __restrict__ labels, int64_t ldl,
                                 const int64_t* __restrict__ frame_lens, const int64_t* __restrict__ label_lens,
                                 int B, int Tn, int U1, int V, int U1p, int64_t blank,
                                 const int64_t* __restrict__ row_offsets, float* __restrict__ grad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * Tn * U1) return;
  const int u = (int)(i % U1);
  const int t = (int)((i / U1) % Tn);
  const int b = (int)(i / ((int64_t)U1 * Tn));
  int64_t Tb = frame_lens[b]; if (Tb > Tn) Tb = Tn;
  const int64_t Ub = rnnt_label_len(label_lens, b, U1);
  if (t >= Tb || u > Ub) return;
  const int64_t D = Tn + U1;
  const int64_t base = (int64_t)b * D * U1p;
  const int64_t o = base + (int64_t)(t + u) * U1p + u;
  const float a = alpha[o];
  if (a == NEG_INF) return;
  const float ll = -nll[b];
  const float w = grad_w[b];
  const int64_t r = row_offsets ? row_offsets[b] + (int64_t)t * (Ub + 1) + u : i;
  float* g = grad + r * V;
  // blank: (t,u) -> (t+1,u); at the final node it terminates the path
  float nb = NEG_INF;
  if (t + 1 < Tb) nb = beta[base + (int64_t)(t + 1 + u) * U1p + u];
  else if (u == Ub) nb = 0.f;
  if (nb > NEG_INF) g[blank] = -w * __expf(a + eb[o] + nb - ll);
  if (u < Ub) {
    const float nl = beta[base + (int64_t)(t + u + 1) * U1p + (u + 1)];
    if (nl > NEG_INF) {
      const int64_t lab = labels[(int64_t)b * ldl + u];
      const float gv = -w * __expf(a + el[o] + nl - ll);
      if (lab == blank) g[blank] += gv; else g[lab] = gv;
    }
  }
}


// ================================================================== fused joint head ==
// Pieces of the chunked joint -> loss path (rnnt.py: RNNTFusedHead): the (B,T,U+1,V) logits
// of model.py:136-144 are never materialised as a whole — they are produced for a block of
// frames, reduced to the two log-probs each lattice node needs, and recomputed per block in
// the backward.

// joint[b,t,u,:] = tanh(enc[b,t,:] + pred[b,u,:])      (model.py:139-140)
// bf16 uses tanh.approx (rel. error 2^-11, below bf16's own rounding), fp32 the precise form.
template <typename T> struct JointTanh { static __device__ __forceinline__ float f(float x) { return tanhf_<true>(x); } };
template <> struct JointTanh<bf16> { static __device__ __forceinline__ float f(float x) { return tanhf_<false>(x); } };

// Vector kernel: one block per (b,t); a thread owns one 16-byte vector of the J channels and
// every (256/vectors-per-row)-th label row u, so enc[b,t,:] is loaded once per thread and the
// block streams U1 contiguous output rows.  No per-element index arithmetic.
template <typename T>
__global__ void __launch_bounds__(256)
joint_fwd_vec_kernel(const T* __restrict__ enc, int64_t enc_sb, int64_t enc_st,
                     const T* __restrict__ pred, int64_t pred_sb, int64_t pred_su,
                     T* __restrict__ out, int Tc, int U1, int J) {
  constexpr int VW = 16 / (int)sizeof(T);
  const int vpr = J / VW, upb = blockDim.x / vpr;
  const int jv = threadIdx.x % vpr, uq = threadIdx.x / vpr;
  if (uq >= upb) return;
  const int b = blockIdx.x / Tc, t = blockIdx.x - b * Tc;
  float fe[VW];
  { Vec<T, VW> ve; ve.raw = __ldg(reinterpret_cast<const uint4*>(enc + b * enc_sb + t * enc_st + jv * VW)); unpack(ve, fe); }
  const T* p = pred + b * pred_sb + jv * VW;
  T* o = out + (int64_t)blockIdx.x * U1 * J + jv * VW;
  for (int u = uq; u < U1; u += upb) {
    float fp[VW];
    Vec<T, VW> vp; vp.raw = __ldg(reinterpret_cast<const uint4*>(p + u * pred_su));
    unpack(vp, fp);
#pragma unroll
    for (int k = 0; k < VW; ++k) fp[k] = JointTanh<T>::f(fe[k] + fp[k]);
    const Vec<T, VW> vo = pack(fp, (T*)nullptr);
    *reinterpret_cast<uint4*>(o + (int64_t)u * J) = vo.raw;
  }
}

// Scalar fallback (odd J / unaligned): one thread per element.
template <typename T>
__global__ void __launch_bounds__(256)
joint_fwd_kernel(const T* __restrict__ enc, int64_t enc_sb, int64_t enc_st,
                 const T* __restrict__ pred, int64_t pred_sb, int64_t pred_su,
                 T* __restrict__ out, int B, int Tc, int U1, int J) {
  const int64_t total = (int64_t)B * Tc * U1 * J;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % J);
    const int64_t row = i / J;
    const int u = (int)(row % U1);
    const int t = (int)((row / U1) % Tc);
    const int b = (int)(row / ((int64_t)U1 * Tc));
    st_f(out + i, JointTanh<T>::f(ld_f(enc + b * enc_sb + t * enc_st + j) + ld_f(pred + b * pred_sb + u * pred_su + j)));
  }
}

// d_pre = dJ * (1 - joint^2), joint recomputed from enc/pred.
// REDUCE_U: d_enc[b,t,:] = sum_u d_pre        (one block per (b,t), written)
// else    : d_pred[b,u,:] += sum_t d_pre      (one block per (b,u), fp32 accumulate across chunks)
// Vector kernels: thread = (row group q, 16-byte channel vector jv); the block's row groups walk
// the reduced index in parallel (16-byte loads of dJ, several rows in flight per thread) and
// are combined through shared memory at the end.
template <typename T, bool REDUCE_U>
__global__ void __launch_bounds__(256)
joint_bwd_vec_kernel(const T* __restrict__ dJ, const T* __restrict__ enc, int64_t enc_sb, int64_t enc_st,
                     const T* __restrict__ pred, int64_t pred_sb, int64_t pred_su,
                     T* __restrict__ d_enc, int64_t denc_sb, int64_t denc_st,
                     float* __restrict__ d_pred, int Tc, int U1, int J) {
  constexpr int VW = 16 / (int)sizeof(T);
  extern __shared__ __align__(16) float red[];           // [groups][J]
  const int vpr = J / VW, groups = blockDim.x / vpr;
  const int jv = threadIdx.x % vpr, q = threadIdx.x / vpr;
  const int nfix = REDUCE_U ? Tc : U1;
  const int b = blockIdx.x / nfix, fixed = blockIdx.x - b * nfix;   // t (REDUCE_U) or u
  const int n = REDUCE_U ? U1 : Tc;
  float acc[VW];
#pragma unroll
  for (int k = 0; k < VW; ++k) acc[k] = 0.f;
  if (q < groups) {
    float fb[VW];
    {
      const T* bp = REDUCE_U ? enc + b * enc_sb + fixed * enc_st : pred + b * pred_sb + fixed * pred_su;
      Vec<T, VW> vb; vb.raw = __ldg(reinterpret_cast<const uint4*>(bp + jv * VW)); unpack(vb, fb);
    }
    for (int i = q; i < n; i += groups) {
      const int t = REDUCE_U ? fixed : i, u = REDUCE_U ? i : fixed;
      const T* op = REDUCE_U ? pred + b * pred_sb + u * pred_su : enc + b * enc_sb + t * enc_st;
      float fo[VW], fg[VW];
      Vec<T, VW> vo, vg;
      vo.raw = __ldg(reinterpret_cast<const uint4*>(op + jv * VW));
      vg.raw = __ldg(reinterpret_cast<const uint4*>(dJ + (((int64_t)b * Tc + t) * U1 + u) * J + jv * VW));
      unpack(vo, fo); unpack(vg, fg);
#pragma unroll
      for (int k = 0; k < VW; ++k) {
        const float jt = JointTanh<T>::f(fb[k] + fo[k]);
        acc[k] = fmaf(fg[k], 1.f - jt * jt, acc[k]);
      }
    }
#pragma unroll
    for (int k = 0; k < VW; ++k) red[q * J + jv * VW + k] = acc[k];
  }
  __syncthreads();
  // first group's threads sum the groups' partials for their vector
  if (q == 0) {
#pragma unroll
    for (int k = 0; k < VW; ++k) {
      float sum = 0.f;
      for (int g = 0; g < groups; ++g) sum += red[g * J + jv * VW + k];
      acc[k] = sum;
    }
    if (REDUCE_U) {
      const Vec<T, VW> vo = pack(acc, (T*)nullptr);
      *reinterpret_cast<uint4*>(d_enc + b * denc_sb + fixed * denc_st + jv * VW) = vo.raw;
    } else {
      float* dp = d_pred + ((int64_t)b * U1 + fixed) * J + jv * VW;
#pragma unroll
      for (int k = 0; k < VW; k += 4) {
        float4 c = *reinterpret_cast<float4*>(dp + k);
        c.x += acc[k]; c.y += acc[k + 1]; c.z += acc[k + 2]; c.w += acc[k + 3];
        *reinterpret_cast<float4*>(dp + k) = c;
      }
    }
  }
}

// (r02, measured and dropped: ONE pass over dJ — a block per (b, 8 frames), d_enc completed in the block, per-block partial
// d_pred rows through an fp32 workspace and a fixed-order reduction kernel, so dJ is read and the tanh evaluated once.
// Bit-identical d_enc, 128 registers, two 8-warp blocks per SM and 512 blocks for a 64-frame block of configs[3]:
// 0.49 ms per block against 0.40 ms for the two passes above (22.9 vs 18.8 ms per step, profiles/r02_call71.sh) — the
// 64 accumulator registers of the frame tile leave too few warps to cover the loads.)
template <typename T, bool REDUCE_U>
__global__ void joint_bwd_kernel(const T* __restrict__ dJ, const T* __restrict__ enc, int64_t enc_sb, int64_t enc_st,
                                 const T* __restrict__ pred, int64_t pred_sb, int64_t pred_su,
                                 T* __restrict__ d_enc, int64_t denc_sb, int64_t denc_st,
                                 float* __restrict__ d_pred, int B, int Tc, int U1, int J) {
  const int b = blockIdx.x / (REDUCE_U ? Tc : U1);
  const int fixed = blockIdx.x % (REDUCE_U ? Tc : U1);  // t (REDUCE_U) or u
  const int n = REDUCE_U ? U1 : Tc;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    float acc = 0.f;
    const float base = REDUCE_U ? ld_f(enc + b * enc_sb + fixed * enc_st + j) : ld_f(pred + b * pred_sb + fixed * pred_su + j);
    for (int i = 0; i < n; ++i) {
      const int t = REDUCE_U ? fixed : i, u = REDUCE_U ? i : fixed;
      const float other = REDUCE_U ? ld_f(pred + b * pred_sb + u * pred_su + j) : ld_f(enc + b * enc_sb + t * enc_st + j);
      const float jt = JointTanh<T>::f(base + other);
      const float g = ld_f(dJ + (((int64_t)b * Tc + t) * U1 + u) * J + j);
      acc = fmaf(g, 1.f - jt * jt, acc);
    }
    if (REDUCE_U) st_f(d_enc + b * denc_sb + fixed * denc_st + j, acc);
    else d_pred[((int64_t)b * U1 + fixed) * J + j] += acc;
  }
}

// rows of joint logits [B, Tc, U1, V] for frames [t0, t0+Tc): lse per node + the two log-probs
// into the skewed eb/el arrays
template <typename T>
__global__ void __launch_bounds__(256, 5)
rnnt_lse_gather_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, int64_t ldl,
                       const int64_t* __restrict__ frame_lens, const int64_t* __restrict__ label_lens,
                       int B, int Tn, int t0, int Tc, int U1, int V, int U1p, int64_t blank,
                       float* __restrict__ lse, float* __restrict__ eb, float* __restrict__ el) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned nrows = (unsigned)B * (unsigned)Tc * (unsigned)U1;   // < 2^31 (checked by the host)
  // (grid-stride loop, launched with one warp per row: a capped persistent grid measured slower)
  for (unsigned row = blockIdx.x * 8 + warp; row < nrows; row += gridDim.x * 8) {
    const unsigned bt = row / (unsigned)U1;
    const int u = (int)(row - bt * (unsigned)U1);
    const int b = (int)(bt / (unsigned)Tc);
    const int tc = (int)(bt - (unsigned)b * (unsigned)Tc);
    const int t = t0 + tc;
    int64_t Tb = frame_lens[b]; if (Tb > Tn) Tb = Tn;
    const int64_t Ub = rnnt_label_len(label_lens, b, U1);
    if (t >= Tb || u > Ub) continue;
    const T* x = logits + (int64_t)row * V;
    const float l = warp_row_lse<T>(x, V, lane);
    if (lane == 0) {
      const int64_t D = Tn + U1;
      const int64_t o = ((int64_t)b * D + (t + u)) * U1p + u;
      lse[((int64_t)b * Tn + t) * U1 + u] = l;
      eb[o] = ld_f(x + blank) - l;
      el[o] = (u < Ub) ? ld_f(x + labels[(int64_t)b * ldl + u]) - l : NEG_INF;
    }
  }
}

// per-node gradients of sum_b w_b*nll_b w.r.t. the blank / label LOG-PROB of the node
__global__ void rnnt_node_grad_kernel(const float* __restrict__ eb, const float* __restrict__ el,
                                      const float* __restrict__ alpha, const float* __restrict__ beta,
                                      const float* __restrict__ nll, const float* __restrict__ grad_w,
                                      const int64_t* __restrict__ frame_lens, const int64_t* __restrict__ label_lens,
                                      int B, int Tn, int U1, int U1p, float* __restrict__ gb, float* __restrict__ gl) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * Tn * U1) return;
  const int u = (int)(i % U1);
  const int t = (int)((i / U1) % Tn);
  const int b = (int)(i / ((int64_t)U1 * Tn));
  int64_t Tb = frame_lens[b]; if (Tb > Tn) Tb = Tn;
  const int64_t Ub = rnnt_label_len(label_lens, b, U1);
  float vb = 0.f, vl = 0.f;
  if (t < Tb && u <= Ub) {
    const int64_t D = Tn + U1;
    const int64_t base = (int64_t)b * D * U1p;
    const int64_t o = base + (int64_t)(t + u) * U1p + u;
    const float a = alpha[o];
    if (a > NEG_INF) {
      const float ll = -nll[b], w = grad_w[b];
      float nb = NEG_INF;
      if (t + 1 < Tb) nb = beta[base + (int64_t)(t + 1 + u) * U1p + u];
      else if (u == Ub) nb = 0.f;
      if (nb > NEG_INF) vb = -w * __expf(a + eb[o] + nb - ll);
      if (u < Ub) {
        const float nl = beta[base + (int64_t)(t + u + 1) * U1p + (u + 1)];
        if (nl > NEG_INF) vl = -w * __expf(a + el[o] + nl - ll);
      }
    }
  }
  gb[i] = vb;
  gl[i] = vl;
}

// dlogits[v] = gb*([v==blank]-p_v) + gl*([v==label]-p_v),  p = softmax(logits) of the node
// Vector kernel: persistent warps stride over the node rows; a lane owns the same NV 16-byte
// column vectors of every row it visits, so the column sums of dlogits (the gradient of the
// joiner's output bias) accumulate in registers and leave through one shared-memory and one
// global atomic per column per block — no second pass over the V-wide tensor.
template <typename T, int NV>
__global__ void __launch_bounds__(256)
rnnt_dlogits_vec_kernel(const T* __restrict__ logits, const float* __restrict__ lse, const float* __restrict__ gb,
                        const float* __restrict__ gl, const int64_t* __restrict__ labels, int64_t ldl,
                        const int64_t* __restrict__ label_lens, int B, int Tn, int t0, int Tc, int U1, int V,
                        int64_t blank, T* __restrict__ dlogits, float* __restrict__ dbias) {
  constexpr int VW = 16 / (int)sizeof(T);
  extern __shared__ __align__(16) float colacc[];        // [V]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (dbias != nullptr) {
    for (int i = threadIdx.x; i < V; i += blockDim.x) colacc[i] = 0.f;
    __syncthreads();
  }
  float cs[NV][VW];
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int j = 0; j < VW; ++j) cs[k][j] = 0.f;
  const unsigned nrows = (unsigned)B * (unsigned)Tc * (unsigned)U1;
  const unsigned wstep = gridDim.x * 8;
  for (unsigned row = blockIdx.x * 8 + warp; row < nrows; row += wstep) {
    const unsigned bt = row / (unsigned)U1;
    const int u = (int)(row - bt * (unsigned)U1);
    const int b = (int)(bt / (unsigned)Tc);
    const int tc = (int)(bt - (unsigned)b * (unsigned)Tc);
    const int64_t node = ((int64_t)b * Tn + t0 + tc) * U1 + u;
    const float vb = gb[node], vl = gl[node];
    T* d = dlogits + (int64_t)row * V;
    if (vb == 0.f && vl == 0.f) {                        // dead node (or zero weight): exact zeros
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int i = (k * 32 + lane) * VW;
        if (i < V) *reinterpret_cast<uint4*>(d + i) = make_uint4(0, 0, 0, 0);
      }
      continue;
    }
    const T* x = logits + (int64_t)row * V;
    const float l2 = lse[node] * 1.4426950408889634f;
    const float tot = vb + vl;
    const int lab = (u < rnnt_label_len(label_lens, b, U1)) ? (int)labels[(int64_t)b * ldl + u] : -1;
    Vec<T, VW> raw[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = (k * 32 + lane) * VW;
      if (i < V) raw[k].raw = __ldg(reinterpret_cast<const uint4*>(x + i));
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = (k * 32 + lane) * VW;
      if (i < V) {
        float g[VW];
        unpack(raw[k], g);
#pragma unroll
        for (int j = 0; j < VW; ++j) {                    // softmax prob = 2^(x*log2e - lse*log2e): one FFMA + one MUFU
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(g[j], 1.4426950408889634f, -l2)));
          g[j] = -tot * e;
        }
        const int ob = (int)blank - i, ol = lab - i;     // the two special columns, if this vector holds them
        if ((unsigned)ob < (unsigned)VW || (unsigned)ol < (unsigned)VW) {
#pragma unroll
          for (int j = 0; j < VW; ++j) {
            if (j == ob) g[j] += vb;
            if (j == ol) g[j] += vl;
          }
        }
        const Vec<T, VW> o = pack(g, (T*)nullptr);
        *reinterpret_cast<uint4*>(d + i) = o.raw;
#pragma unroll
        for (int j = 0; j < VW; ++j) cs[k][j] += g[j];
      }
    }
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = (k * 32 + lane) * VW;
      if (i < V) {
#pragma unroll
        for (int j = 0; j < VW; ++j) atomicAdd(colacc + i + j, cs[k][j]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < V; i += blockDim.x) atomicAdd(dbias + i, colacc[i]);
  }
}

// Scalar fallback (any V / alignment); the bias gradient is then a separate column-sum pass.
template <typename T>
__global__ void __launch_bounds__(256)
rnnt_dlogits_kernel(const T* __restrict__ logits, const float* __restrict__ lse, const float* __restrict__ gb,
                    const float* __restrict__ gl, const int64_t* __restrict__ labels, int64_t ldl,
                    const int64_t* __restrict__ label_lens, int B, int Tn, int t0, int Tc, int U1, int V,
                    int64_t blank, T* __restrict__ dlogits) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= (int64_t)B * Tc * U1) return;
  const int u = (int)(row % U1);
  const int tc = (int)((row / U1) % Tc);
  const int b = (int)(row / ((int64_t)U1 * Tc));
  const int64_t node = ((int64_t)b * Tn + t0 + tc) * U1 + u;
  const float vb = gb[node], vl = gl[node];
  T* d = dlogits + row * V;
  if (vb == 0.f && vl == 0.f) {                        // dead node (or zero weight): exact zeros
    for (int i = lane; i < V; i += 32) st_f(d + i, 0.f);
    return;
  }
  const T* x = logits + row * V;
  const float l = lse[node];
  const float tot = vb + vl;
  const int64_t lab = (u < rnnt_label_len(label_lens, b, U1)) ? labels[(int64_t)b * ldl + u] : -1;
  for (int i = lane; i < V; i += 32) {
    float g = -tot * __expf(ld_f(x + i) - l);
    if (i == blank) g += vb;
    if (i == lab) g += vl;
    st_f(d + i, g);
  }
}

}  // namespace sc

using namespace sc;

static bool rnnt_args_ok(int64_t B, int64_t T, int64_t U1, int64_t V, int64_t blank) {
  return B > 0 && T >= 0 && U1 >= 1 && U1 <= 1024 && V > 0 && blank >= 0 && blank < V &&
         B * T * U1 < ((int64_t)1 << 40) && B * (T + U1) * U1 < ((int64_t)1 << 31);
}

extern "C" int sc_rnnt_fwd(const float* log_probs, const int64_t* labels, int64_t ldl,
                           const int64_t* frame_lens, const int64_t* label_lens,
                           int64_t B, int64_t T, int64_t U1, int64_t V, int64_t blank,
                           const int64_t* row_offsets,
                           float* eb, float* el, float* alpha, float* beta, float* nll, void* stream) {
  SC_CHECK_ARG(rnnt_args_ok(B, T, U1, V, blank), SC_E_SHAPE);
  SC_CHECK_ARG(frame_lens && label_lens && nll && (U1 == 1 || labels), SC_E_BADARG);
  SC_CHECK_ARG(T == 0 || (log_probs && eb && el && alpha && beta), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const int U1p = (int)((U1 + 3) & ~(int64_t)3);
  if (T > 0) {
    const int64_t n = B * T * U1;
    rnnt_gather_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(log_probs, labels, ldl, frame_lens, label_lens,
        (int)B, (int)T, (int)U1, (int)V, U1p, blank, row_offsets, eb, el);
  }
  int threads = ((U1p + 31) / 32) * 32;
  const size_t smem = (2 * (size_t)(U1p + 4) + 4 * (size_t)RNNT_EB * U1p) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rnnt_alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  rnnt_alpha_beta_kernel<<<dim3((unsigned)B, 2), threads, smem, st>>>(eb, el, frame_lens, label_lens, (int)T, (int)U1,
      U1p, alpha, beta, nll);
  SC_LAUNCH_RET();
}

extern "C" int sc_rnnt_bwd(const int64_t* labels, int64_t ldl, const int64_t* frame_lens,
                           const int64_t* label_lens, int64_t B, int64_t T, int64_t U1, int64_t V,
                           int64_t blank, const int64_t* row_offsets, int64_t total_rows,
                           const float* eb, const float* el, const float* alpha,
                           const float* beta, const float* nll, const float* grad_w, float* grad,
                           void* stream) {
  SC_CHECK_ARG(rnnt_args_ok(B, T, U1, V, blank), SC_E_SHAPE);
  if (T == 0) return 0;
  SC_CHECK_ARG(frame_lens && label_lens && eb && el && alpha && beta && nll && grad_w && grad, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const int U1p = (int)((U1 + 3) & ~(int64_t)3);
  const int64_t rows = row_offsets ? total_rows : B * T * U1;
  SC_CHECK_ARG(rows >= 0, SC_E_BADARG);
  cudaError_t e = cudaMemsetAsync(grad, 0, sizeof(float) * (size_t)(rows * V), st);
  if (e != cudaSuccess) return (int)e;
  const int64_t n = B * T * U1;
  rnnt_grad_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(eb, el, alpha, beta, nll, grad_w, labels, ldl, frame_lens,
      label_lens, (int)B, (int)T, (int)U1, (int)V, U1p, blank, row_offsets, grad);
  SC_LAUNCH_RET();
}

// ---------------------------------------------------------------- fused joint head ABI
extern "C" int sc_joint_fwd(const void* enc, int64_t enc_sb, int64_t enc_st, const void* pred, int64_t pred_sb,
                            int64_t pred_su, void* out, int64_t B, int64_t Tc, int64_t U1, int64_t J, int dtype,
                            void* stream) {
  SC_CHECK_ARG(B > 0 && Tc >= 0 && U1 > 0 && J > 0 && B * Tc * U1 < ((int64_t)1 << 31), SC_E_SHAPE);
  if (Tc == 0) return 0;
  SC_CHECK_ARG(enc && pred && out, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = (((uintptr_t)enc | (uintptr_t)pred | (uintptr_t)out) & 15) == 0;
#define SC_JOINT(TT, VW_) do { \
    const bool vec = al && (J % (VW_) == 0) && (J / (VW_) <= 256) && (enc_sb % (VW_) == 0) && (enc_st % (VW_) == 0) && \
                     (pred_sb % (VW_) == 0) && (pred_su % (VW_) == 0); \
    if (vec) joint_fwd_vec_kernel<TT><<<(unsigned)(B * Tc), 256, 0, st>>>((const TT*)enc, enc_sb, enc_st, (const TT*)pred, pred_sb, pred_su, (TT*)out, (int)Tc, (int)U1, (int)J); \
    else joint_fwd_kernel<TT><<<(unsigned)min((int64_t)148 * 64, cdiv(B * Tc * U1 * J, 256)), 256, 0, st>>>((const TT*)enc, enc_sb, enc_st, (const TT*)pred, pred_sb, pred_su, (TT*)out, (int)B, (int)Tc, (int)U1, (int)J); } while (0)
  if (dtype == SC_BF16) SC_JOINT(bf16, 8);
  else if (dtype == SC_F32) SC_JOINT(float, 4);
  else return SC_E_DTYPE;
#undef SC_JOINT
  SC_LAUNCH_RET();
}

extern "C" int sc_joint_bwd(const void* dJ, const void* enc, int64_t enc_sb, int64_t enc_st, const void* pred,
                            int64_t pred_sb, int64_t pred_su, void* d_enc, int64_t denc_sb, int64_t denc_st,
                            float* d_pred, int64_t B, int64_t Tc, int64_t U1, int64_t J, int dtype, void* stream) {
  SC_CHECK_ARG(B > 0 && Tc >= 0 && U1 > 0 && J > 0, SC_E_SHAPE);
  if (Tc == 0) return 0;
  SC_CHECK_ARG(dJ && enc && pred && d_enc && d_pred, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = (((uintptr_t)dJ | (uintptr_t)enc | (uintptr_t)pred | (uintptr_t)d_enc | (uintptr_t)d_pred) & 15) == 0;
  const int threads = J >= 256 ? 256 : 128;
#define SC_JBWD(TT, VW_) do { \
    const bool vec = al && (J % (VW_) == 0) && (J / (VW_) <= 256) && (enc_sb % (VW_) == 0) && (enc_st % (VW_) == 0) && \
                     (pred_sb % (VW_) == 0) && (pred_su % (VW_) == 0) && (denc_sb % (VW_) == 0) && (denc_st % (VW_) == 0); \
    if (vec) { \
      const size_t smem = (size_t)(256 / (J / (VW_))) * J * sizeof(float); \
      if (smem > 48 * 1024) return SC_E_SHAPE; \
      joint_bwd_vec_kernel<TT, true><<<(unsigned)(B * Tc), 256, smem, st>>>((const TT*)dJ, (const TT*)enc, enc_sb, enc_st, (const TT*)pred, pred_sb, pred_su, (TT*)d_enc, denc_sb, denc_st, d_pred, (int)Tc, (int)U1, (int)J); \
      joint_bwd_vec_kernel<TT, false><<<(unsigned)(B * U1), 256, smem, st>>>((const TT*)dJ, (const TT*)enc, enc_sb, enc_st, (const TT*)pred, pred_sb, pred_su, (TT*)d_enc, denc_sb, denc_st, d_pred, (int)Tc, (int)U1, (int)J); \
    } else { \
      joint_bwd_kernel<TT, true><<<(unsigned)(B * Tc), threads, 0, st>>>((const TT*)dJ, (const TT*)enc, enc_sb, enc_st, (const TT*)pred, pred_sb, pred_su, (TT*)d_enc, denc_sb, denc_st, d_pred, (int)B, (int)Tc, (int)U1, (int)J); \
      joint_bwd_kernel<TT, false><<<(unsigned)(B * U1), threads, 0, st>>>((const TT*)dJ, (const TT*)enc, enc_sb, enc_st, (const TT*)pred, pred_sb, pred_su, (TT*)d_enc, denc_sb, denc_st, d_pred, (int)B, (int)Tc, (int)U1, (int)J); \
    } } while (0)
  if (dtype == SC_BF16) SC_JBWD(bf16, 8);
  else if (dtype == SC_F32) SC_JBWD(float, 4);
  else return SC_E_DTYPE;
#undef SC_JBWD
  SC_LAUNCH_RET();
}

extern "C" int sc_rnnt_lse_gather(const void* logits, int dtype, const int64_t* labels, int64_t ldl,
                                  const int64_t* frame_lens, const int64_t* label_lens, int64_t B, int64_t T,
                                  int64_t t0, int64_t Tc, int64_t U1, int64_t V, int64_t blank, float* lse,
                                  float* eb, float* el, void* stream) {
  SC_CHECK_ARG(rnnt_args_ok(B, T, U1, V, blank) && t0 >= 0 && Tc >= 0 && t0 + Tc <= T && B * Tc * U1 < ((int64_t)1 << 31), SC_E_SHAPE);
  if (Tc == 0) return 0;
  SC_CHECK_ARG(logits && frame_lens && label_lens && lse && eb && el && (U1 == 1 || labels), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const int U1p = (int)((U1 + 3) & ~(int64_t)3);
  const unsigned grid = (unsigned)cdiv(B * Tc * U1, 8);
  if (dtype == SC_BF16) rnnt_lse_gather_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)logits, labels, ldl, frame_lens, label_lens, (int)B, (int)T, (int)t0, (int)Tc, (int)U1, (int)V, U1p, blank, lse, eb, el);
  else if (dtype == SC_F32) rnnt_lse_gather_kernel<float><<<grid, 256, 0, st>>>((const float*)logits, labels, ldl, frame_lens, label_lens, (int)B, (int)T, (int)t0, (int)Tc, (int)U1, (int)V, U1p, blank, lse, eb, el);
  else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}

// alpha/beta + nll from already-filled eb/el (the second half of sc_rnnt_fwd)
extern "C" int sc_rnnt_lattice(const int64_t* frame_lens, const int64_t* label_lens, int64_t B, int64_t T, int64_t U1,
                               const float* eb, const float* el, float* alpha, float* beta, float* nll, void* stream) {
  SC_CHECK_ARG(rnnt_args_ok(B, T, U1, 1, 0), SC_E_SHAPE);
  SC_CHECK_ARG(frame_lens && label_lens && nll && (T == 0 || (eb && el && alpha && beta)), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const int U1p = (int)((U1 + 3) & ~(int64_t)3);
  int threads = ((U1p + 31) / 32) * 32;
  const size_t smem = (2 * (size_t)(U1p + 4) + 4 * (size_t)RNNT_EB * U1p) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(rnnt_alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  rnnt_alpha_beta_kernel<<<dim3((unsigned)B, 2), threads, smem, st>>>(eb, el, frame_lens, label_lens, (int)T, (int)U1,
      U1p, alpha, beta, nll);
  SC_LAUNCH_RET();
}

extern "C" int sc_rnnt_node_grads(const int64_t* frame_lens, const int64_t* label_lens, int64_t B, int64_t T,
                                  int64_t U1, const float* eb, const float* el, const float* alpha, const float* beta,
                                  const float* nll, const float* grad_w, float* gb, float* gl, void* stream) {
  SC_CHECK_ARG(rnnt_args_ok(B, T, U1, 1, 0), SC_E_SHAPE);
  if (T == 0) return 0;
  SC_CHECK_ARG(frame_lens && label_lens && eb && el && alpha && beta && nll && grad_w && gb && gl, SC_E_BADARG);
  const int U1p = (int)((U1 + 3) & ~(int64_t)3);
  rnnt_node_grad_kernel<<<(unsigned)cdiv(B * T * U1, 256), 256, 0, (cudaStream_t)stream>>>(eb, el, alpha, beta, nll, grad_w,
      frame_lens, label_lens, (int)B, (int)T, (int)U1, U1p, gb, gl);
  SC_LAUNCH_RET();
}

extern "C" int sc_rnnt_dlogits(const void* logits, int dtype, const float* lse, const float* gb, const float* gl,
                               const int64_t* labels, int64_t ldl, const int64_t* label_lens, int64_t B, int64_t T,
                               int64_t t0, int64_t Tc, int64_t U1, int64_t V, int64_t blank, void* dlogits,
                               float* dbias, void* stream) {
  SC_CHECK_ARG(rnnt_args_ok(B, T, U1, V, blank) && t0 >= 0 && Tc >= 0 && t0 + Tc <= T && B * Tc * U1 < ((int64_t)1 << 31), SC_E_SHAPE);
  if (Tc == 0) return 0;
  SC_CHECK_ARG(logits && lse && gb && gl && label_lens && dlogits && (U1 == 1 || labels), SC_E_BADARG);
  SC_CHECK_ARG(dtype == SC_BF16 || dtype == SC_F32, SC_E_DTYPE);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = B * Tc * U1;
  const int vw = dtype == SC_BF16 ? 8 : 4;
  const bool vec = (V % vw == 0) && V <= 1024 && ((((uintptr_t)logits | (uintptr_t)dlogits) & 15) == 0);
  if (vec) {
    // persistent: enough warps to fill the machine, few enough blocks that the per-block column
    // atomics stay negligible
    const unsigned grid = (unsigned)min(cdiv(rows, 8), (int64_t)num_sms() * 4);
    const size_t smem = (size_t)V * sizeof(float);
    if (dtype == SC_BF16)
      rnnt_dlogits_vec_kernel<bf16, 4><<<grid, 256, smem, st>>>((const bf16*)logits, lse, gb, gl, labels, ldl, label_lens,
          (int)B, (int)T, (int)t0, (int)Tc, (int)U1, (int)V, blank, (bf16*)dlogits, dbias);
    else
      rnnt_dlogits_vec_kernel<float, 8><<<grid, 256, smem, st>>>((const float*)logits, lse, gb, gl, labels, ldl, label_lens,
          (int)B, (int)T, (int)t0, (int)Tc, (int)U1, (int)V, blank, (float*)dlogits, dbias);
    SC_LAUNCH_RET();
  }
  const unsigned grid = (unsigned)cdiv(rows, 8);
  if (dtype == SC_BF16) rnnt_dlogits_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)logits, lse, gb, gl, labels, ldl, label_lens, (int)B, (int)T, (int)t0, (int)Tc, (int)U1, (int)V, blank, (bf16*)dlogits);
  else rnnt_dlogits_kernel<float><<<grid, 256, 0, st>>>((const float*)logits, lse, gb, gl, labels, ldl, label_lens, (int)B, (int)T, (int)t0, (int)Tc, (int)U1, (int)V, blank, (float*)dlogits);
  if (dbias != nullptr) return sc_colsum(dlogits, V, dtype, dbias, rows, V, 1, stream);
  SC_LAUNCH_RET();
}
