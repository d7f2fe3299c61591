// K3: CTC loss and gradient, fused with log-softmax and its backward (sm_100a).
//
// Replaces model.py:70-71 (`enc_out.log_softmax(-1).transpose(0,1)` + nn.CTCLoss(blank=0,
// zero_infinity=True), train.py:142), i.e. ATen's log_softmax / ctc_loss_gpu /
// ctc_loss_backward_gpu chain.  Semantics: SURVEY.md Appendix B.
//
// Passes (all fp32 arithmetic, log space):
//   1. ctc_lse_gather : one warp per frame (b,t<T_b): 128-bit coalesced sweep over the V
//      logits -> lse[b,t]; the 2U+1 lattice emissions logit[l'_s]-lse are gathered into the
//      compact lplat[b,t,s] while the row is still hot in L1, so the serial recursions never
//      touch the V-wide tensor.
//   2. ctc_alpha_beta : one CTA per (utterance, direction); lattice node s on thread s; the
//      previous column lives in a double-buffered shared-memory line, one __syncthreads per
//      timestep; the emission for step t+1 is prefetched before the barrier of step t.
//   3. ctc_grad       : one warp per frame: dlogits = gout*scale_b*(softmax - occupancy),
//      occupancy scattered with shared-memory atomics; exact zeros for t>=T_b and for
//      infeasible utterances (zero_infinity).
// Algorithmic HBM bytes per frame: 3*V*e + 8*(2U+1)  (SURVEY.md 8d).
#include "sc_common.cuh"

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
constexpr int CTC_PF = 8;        // emission prefetch depth of the alpha/beta recursion   // alpha/beta columns are re-centred every this many steps

// ---- pass 1 ------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(CTC_WARPS * 32)
ctc_lse_gather_kernel(const T* __restrict__ logits, int64_t stride_b, int64_t stride_t,
                      const int64_t* __restrict__ targets, int64_t ldt,
                      const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                      int B, int Tn, int V, int Smax, int64_t blank,
                      float* __restrict__ lse, float* __restrict__ lplat) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * CTC_WARPS + warp;
  if (row >= (int64_t)B * Tn) return;
  const int b = (int)(row / Tn), t = (int)(row % Tn);
  int64_t Tb = in_lens[b]; if (Tb > Tn) Tb = Tn;
  if (t >= Tb) return;
  const T* x = logits + b * stride_b + t * stride_t;
  // online max/sum, each lane over a strided slice
  float m = NEG_INF, ssum = 0.f;
  constexpr int VW = 16 / sizeof(T);
  const bool vec_ok = (V % VW == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
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
  ssum = (m == NEG_INF) ? 0.f : ssum * __expf(m - gm);
  const float gs = warp_sum(ssum);
  const float l = gm + __logf(gs);
  if (lane == 0) lse[row] = l;
  const int S = 2 * (int)tgt_lens[b] + 1;
  const int64_t* tg = targets + (int64_t)b * ldt;
  float* out = lplat + row * Smax;
  for (int s = lane; s < S; s += 32) out[s] = ld_f(x + ext_label(tg, s, blank)) - l;
}

// ---- pass 2 ------------------------------------------------------------------------
// dir 0 = alpha (forward in t), dir 1 = beta (backward in t).  Both include the emission of
// their own timestep, as in the oracle (ctc_oracle.py) and ATen.
__global__ void ctc_alpha_beta_kernel(const float* __restrict__ lplat,
                                      const int64_t* __restrict__ targets, int64_t ldt,
                                      const int64_t* __restrict__ in_lens,
                                      const int64_t* __restrict__ tgt_lens,
                                      int Tn, int Smax, int64_t blank,
                                      float* __restrict__ alpha, float* __restrict__ beta,
                                      float* __restrict__ nll) {
  extern __shared__ float sm[];       // 2 lines of (Smax + 4) floats, 2 pad cells either side
  __shared__ float red[32];
  const int b = blockIdx.x, dir = blockIdx.y;
  int64_t Tb64 = in_lens[b]; if (Tb64 > Tn) Tb64 = Tn;
  const int Tb = (int)Tb64;
  const int U = (int)tgt_lens[b];
  const int S = 2 * U + 1;
  const int64_t* tg = targets + (int64_t)b * ldt;
  const int LINE = Smax + 4;
  if (Tb <= 0) {
    if (dir == 0 && threadIdx.x == 0) nll[b] = (U == 0) ? 0.f : INFINITY;
    return;
  }
  // per-thread skip permission for each owned node (nodes tid, tid+blockDim, ...)
  for (int i = threadIdx.x; i < 2 * LINE; i += blockDim.x) sm[i] = NEG_INF;
  __syncthreads();
  float* bufA = sm + 2;
  float* bufB = sm + LINE + 2;
  const float* lp_b = lplat + (int64_t)b * Tn * Smax;
  float* out_b = (dir == 0 ? alpha : beta) + (int64_t)b * Tn * Smax;

  const int t_first = dir == 0 ? 0 : Tb - 1;
  const int step = dir == 0 ? 1 : -1;
  // init column
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    float v = NEG_INF;
    if (dir == 0) { if (s < 2) v = lp_b[(int64_t)t_first * Smax + s]; }
    else          { if (s >= S - 2) v = lp_b[(int64_t)t_first * Smax + s]; }
    bufA[s] = v;
    out_b[(int64_t)t_first * Smax + s] = v;
  }
  __syncthreads();
  float* prev = bufA;
  float* cur = bufB;
  // Log-space values drift to magnitude ~T*log(V); every CTC_RENORM steps the column is
  // re-centred on its maximum so fp32 keeps ~1e-6 absolute resolution for any T.  The removed
  // offsets only matter for the likelihood (summed in double); the gradient pass normalises
  // each frame's occupancies itself (sum_s exp(alpha+beta-lp) is the same for every t).
  double csum = 0.0;
  // one lattice step for node s with emission e (reads prev, writes cur)
  // skip transition s-2 -> s (alpha) / s+2 -> s (beta) is a per-node constant: allowed iff
  // l'_s is a label that differs from the label two nodes away
  auto skip_ok = [&](int s) -> bool {
    if (!(s & 1)) return false;
    if (dir == 0) return s >= 2 && tg[s >> 1] != tg[(s >> 1) - 1];
    return s + 2 < S && tg[s >> 1] != tg[(s >> 1) + 1];
  };
  // one lattice step for node s with emission e (reads prev, writes cur)
  auto node = [&](int s, float e, bool skip) {
    float v;
    if (dir == 0) {
      v = lse3(prev[s], prev[s - 1], skip ? prev[s - 2] : NEG_INF) + e;
    } else {
      // pad cells right of S-1 hold -inf, so s+1 / s+2 need no bounds test
      v = lse3(prev[s], prev[s + 1], skip ? prev[s + 2] : NEG_INF) + e;
    }
    cur[s] = v;
  };
  // re-centre the freshly written column (every CTC_RENORM steps), publish it, flip buffers
  auto finish = [&](int i, int t) {
    if ((i % CTC_RENORM) == 0) {
      float m = NEG_INF;
      for (int s = threadIdx.x; s < S; s += blockDim.x) m = fmaxf(m, cur[s]);   // own nodes only
      m = warp_max(m);
      if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
      __syncthreads();
      m = NEG_INF;
      for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) m = fmaxf(m, red[w]);
      if (m > NEG_INF) {
        for (int s = threadIdx.x; s < S; s += blockDim.x) cur[s] -= m;
        csum += (double)m;
      }
    }
    for (int s = threadIdx.x; s < S; s += blockDim.x) out_b[(int64_t)t * Smax + s] = cur[s];
    __syncthreads();
    float* tmp = prev; prev = cur; cur = tmp;
  };
  if (S <= (int)blockDim.x) {
    // one node per thread: the emission stream is independent of the recursion, so it is
    // prefetched CTC_PF steps ahead into registers and the serial chain never waits on HBM
    const int s = threadIdx.x;
    const bool has = s < S;
    const bool skip = has && skip_ok(s);
    float ring[CTC_PF];
#pragma unroll
    for (int j = 0; j < CTC_PF; ++j) {
      const int i = 1 + j;
      ring[j] = (has && i < Tb) ? __ldg(lp_b + (int64_t)(t_first + i * step) * Smax + s) : 0.f;
    }
    for (int i0 = 1; i0 < Tb; i0 += CTC_PF) {
#pragma unroll
      for (int j = 0; j < CTC_PF; ++j) {
        const int i = i0 + j;
        if (i < Tb) {
          const float e = ring[j];
          const int ip = i + CTC_PF;
          ring[j] = (has && ip < Tb) ? __ldg(lp_b + (int64_t)(t_first + ip * step) * Smax + s) : 0.f;
          if (has) node(s, e, skip);
          finish(i, t_first + i * step);
        }
      }
    }
  } else {
    for (int i = 1; i < Tb; ++i) {
      const int t = t_first + i * step;
      for (int s = threadIdx.x; s < S; s += blockDim.x) node(s, lp_b[(int64_t)t * Smax + s], skip_ok(s));
      finish(i, t);
    }
  }
  if (dir == 0 && threadIdx.x == 0) {
    const float ll = lse2(prev[S - 1], S > 1 ? prev[S - 2] : NEG_INF);
    nll[b] = (ll == NEG_INF) ? INFINITY : (float)(-(csum + (double)ll));   // +inf when infeasible
  }
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
template <typename TI, typename TO>
__global__ void __launch_bounds__(CTC_WARPS * 32)
ctc_grad_kernel(const TI* __restrict__ logits, int64_t stride_b, int64_t stride_t,
                const int64_t* __restrict__ targets, int64_t ldt,
                const int64_t* __restrict__ in_lens, const int64_t* __restrict__ tgt_lens,
                int B, int Tn, int V, int Smax, int64_t blank,
                const float* __restrict__ lse, const float* __restrict__ alpha,
                const float* __restrict__ beta, const float* __restrict__ nll,
                const float* __restrict__ grad_out, int reduction,
                TO* __restrict__ dlogits, int64_t dstride_b, int64_t dstride_t) {
  extern __shared__ float sm[];                     // CTC_WARPS rows of V floats
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * CTC_WARPS + warp;
  if (row >= (int64_t)B * Tn) return;
  const int b = (int)(row / Tn), t = (int)(row % Tn);
  int64_t Tb = in_lens[b]; if (Tb > Tn) Tb = Tn;
  TO* dx = dlogits + b * dstride_b + t * dstride_t;
  const float n = nll[b];
  if (t >= Tb || !isfinite(n)) {
    for (int i = lane; i < V; i += 32) st_f(dx + i, 0.f);
    return;
  }
  const TI* x = logits + b * stride_b + t * stride_t;
  float* r = sm + (int64_t)warp * V;
  const float l = lse[row];
  for (int i = lane; i < V; i += 32) r[i] = __expf(ld_f(x + i) - l);
  __syncwarp();
  const int U = (int)tgt_lens[b];
  const int S = 2 * U + 1;
  const int64_t* tg = targets + (int64_t)b * ldt;
  const float* al = alpha + row * Smax;
  const float* be = beta + row * Smax;
  // occupancy_s = exp(alpha+beta-lp)_s / sum_s' exp(alpha+beta-lp)_s'   (the sum is exp(-nll)
  // for every frame; normalising per frame keeps it exact under the re-centred alpha/beta)
  float vmax = NEG_INF;
  for (int s = lane; s < S; s += 32) {
    const float lp = ld_f(x + ext_label(tg, s, blank)) - l;
    vmax = fmaxf(vmax, al[s] + be[s] - lp);
  }
  vmax = warp_max(vmax);
  float zsum = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float lp = ld_f(x + ext_label(tg, s, blank)) - l;
    const float v = al[s] + be[s] - lp;
    if (v > NEG_INF) zsum += __expf(v - vmax);
  }
  zsum = warp_sum(zsum);
  const float inv = (zsum > 0.f) ? 1.f / zsum : 0.f;
  for (int s = lane; s < S; s += 32) {
    const int64_t lab = ext_label(tg, s, blank);
    const float lp = ld_f(x + lab) - l;
    const float v = al[s] + be[s] - lp;
    if (v > NEG_INF) atomicAdd(r + lab, -__expf(v - vmax) * inv);
  }
  __syncwarp();
  float scale;
  if (reduction == 1) scale = grad_out[0] / ((float)B * fmaxf((float)U, 1.f));
  else if (reduction == 2) scale = grad_out[0];
  else scale = grad_out[b];
  for (int i = lane; i < V; i += 32) st_f(dx + i, scale * r[i]);
}

}  // namespace sc

using namespace sc;

extern "C" int sc_ctc_fwd(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                          const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                          const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                          int64_t blank, float* lse, float* lplat, float* alpha, float* beta,
                          float* nll, float* loss, int reduction, void* stream) {
  SC_CHECK_ARG(B > 0 && T >= 0 && V > 0 && Umax >= 0 && blank >= 0 && blank < V, SC_E_BADARG);
  SC_CHECK_ARG(in_lens && tgt_lens && nll && (Umax == 0 || targets), SC_E_BADARG);
  SC_CHECK_ARG(reduction >= 0 && reduction <= 2 && (reduction == 0 || loss), SC_E_BADARG);
  SC_CHECK_ARG(T == 0 || (logits && lse && lplat && alpha && beta), SC_E_BADARG);
  SC_CHECK_ARG(B * T < ((int64_t)1 << 31) && V < (1 << 30) && Umax < (1 << 20), SC_E_SHAPE);
  cudaStream_t st = (cudaStream_t)stream;
  const int Smax = (int)(2 * Umax + 1);
  if (T > 0) {
    const unsigned blocks = (unsigned)cdiv(B * T, CTC_WARPS);
    if (dtype == SC_F32)
      ctc_lse_gather_kernel<float><<<blocks, CTC_WARPS * 32, 0, st>>>((const float*)logits, stride_b, stride_t,
          targets, ldt, in_lens, tgt_lens, (int)B, (int)T, (int)V, Smax, blank, lse, lplat);
    else if (dtype == SC_BF16)
      ctc_lse_gather_kernel<bf16><<<blocks, CTC_WARPS * 32, 0, st>>>((const bf16*)logits, stride_b, stride_t,
          targets, ldt, in_lens, tgt_lens, (int)B, (int)T, (int)V, Smax, blank, lse, lplat);
    else return SC_E_DTYPE;
  }
  int threads = ((Smax + 31) / 32) * 32;
  if (threads > 1024) threads = 1024;
  const size_t smem = 2 * (size_t)(Smax + 4) * sizeof(float);
  SC_CHECK_ARG(smem <= 200 * 1024, SC_E_SHAPE);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctc_alpha_beta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  ctc_alpha_beta_kernel<<<dim3((unsigned)B, 2), threads, smem, st>>>(lplat, targets, ldt, in_lens, tgt_lens,
      (int)T, Smax, blank, alpha, beta, nll);
  if (reduction != 0) ctc_reduce_kernel<<<1, 32, 0, st>>>(nll, tgt_lens, (int)B, reduction, loss);
  SC_LAUNCH_RET();
}

template <typename TI, typename TO>
static int launch_ctc_grad(const void* logits, int64_t stride_b, int64_t stride_t,
                           const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                           const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int Smax,
                           int64_t blank, const float* lse, const float* alpha, const float* beta,
                           const float* nll, const float* grad_out, int reduction, void* dlogits,
                           int64_t dstride_b, int64_t dstride_t, cudaStream_t st) {
  const size_t smem = (size_t)CTC_WARPS * V * sizeof(float);
  if (smem > 200 * 1024) return SC_E_SHAPE;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(ctc_grad_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  const unsigned blocks = (unsigned)cdiv(B * T, CTC_WARPS);
  ctc_grad_kernel<TI, TO><<<blocks, CTC_WARPS * 32, smem, st>>>((const TI*)logits, stride_b, stride_t, targets, ldt,
      in_lens, tgt_lens, (int)B, (int)T, (int)V, Smax, blank, lse, alpha, beta, nll, grad_out, reduction,
      (TO*)dlogits, dstride_b, dstride_t);
  SC_LAUNCH_RET();
}

extern "C" int sc_ctc_bwd(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                          const int64_t* targets, int64_t ldt, const int64_t* in_lens,
                          const int64_t* tgt_lens, int64_t B, int64_t T, int64_t V, int64_t Umax,
                          int64_t blank, const float* lse, const float* alpha, const float* beta,
                          const float* nll, const float* grad_out, int reduction,
                          void* dlogits, int64_t dstride_b, int64_t dstride_t, int out_dtype,
                          void* stream) {
  SC_CHECK_ARG(B > 0 && T >= 0 && V > 0 && Umax >= 0 && blank >= 0 && blank < V, SC_E_BADARG);
  if (T == 0) return 0;
  SC_CHECK_ARG(logits && in_lens && tgt_lens && lse && alpha && beta && nll && grad_out && dlogits, SC_E_BADARG);
  SC_CHECK_ARG(reduction >= 0 && reduction <= 2, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const int Smax = (int)(2 * Umax + 1);
#define SC_CTC_GRAD(TI, TO) launch_ctc_grad<TI, TO>(logits, stride_b, stride_t, targets, ldt, in_lens, tgt_lens, \
    B, T, V, Smax, blank, lse, alpha, beta, nll, grad_out, reduction, dlogits, dstride_b, dstride_t, st)
  if (dtype == SC_F32 && out_dtype == SC_F32) return SC_CTC_GRAD(float, float);
  if (dtype == SC_BF16 && out_dtype == SC_BF16) return SC_CTC_GRAD(bf16, bf16);
  if (dtype == SC_F32 && out_dtype == SC_BF16) return SC_CTC_GRAD(float, bf16);
  if (dtype == SC_BF16 && out_dtype == SC_F32) return SC_CTC_GRAD(bf16, float);
#undef SC_CTC_GRAD
  return SC_E_DTYPE;
}
