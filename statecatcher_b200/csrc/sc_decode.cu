// Greedy CTC decoding: replaces decoder.py:3-30 — torch.argmax plus a
// Python loop with one `.item()` device sync per token — by two kernels and a single
// device->host copy.  Integer work: results are bit-exact (ties in the argmax resolve to the
// lowest index, as torch.argmax does on the reference's CPU path).
#include "sc_common.cuh"

namespace sc {

// one warp per frame: argmax over V
template <typename T>
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const T* __restrict__ x, int64_t stride_b, int64_t stride_t, const int64_t* __restrict__ in_lens,
                   int B, int Tn, int V, int* __restrict__ pred) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row >= (int64_t)B * Tn) return;
  const int b = (int)(row / Tn), t = (int)(row % Tn);
  if (t >= in_lens[b]) return;
  const T* r = x + b * stride_b + t * stride_t;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < V; i += 32) {
    const float f = ld_f(r + i);
    if (f > best || (f == best && i < bi)) { best = f; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) pred[row] = (bi == 0x7fffffff) ? 0 : bi;
}

// one block per utterance: drop blanks and repeats, compact in order
__global__ void __launch_bounds__(256)
ctc_collapse_kernel(const int* __restrict__ pred, const int64_t* __restrict__ in_lens, int Tn, int64_t blank,
                    int64_t* __restrict__ out_tokens, int64_t* __restrict__ out_lens) {
  __shared__ int wsum[8];
  __shared__ int carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int64_t Tb = in_lens[b]; if (Tb > Tn) Tb = Tn;
  const int* p = pred + (int64_t)b * Tn;
  int64_t* out = out_tokens + (int64_t)b * Tn;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < Tb; t0 += 256) {
    const int t = t0 + tid;
    int tok = 0, keep = 0;
    if (t < Tb) {
      tok = p[t];
      keep = (tok != (int)blank) && (t == 0 || tok != p[t - 1]);
    }
    // block exclusive scan of keep
    int incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int base = carry;
    for (int w = 0; w < warp; ++w) base += wsum[w];
    if (keep) out[base + incl - 1] = tok;
    __syncthreads();
    if (tid == 255) carry = base + incl;
    __syncthreads();
  }
  if (tid == 0) out_lens[b] = carry;
}

}  // namespace sc

using namespace sc;

extern "C" int sc_ctc_greedy_decode(const void* logits, int64_t stride_b, int64_t stride_t, int dtype,
                                    const int64_t* in_lens, int64_t B, int64_t T, int64_t V, int64_t blank,
                                    int* pred, int64_t* out_tokens, int64_t* out_lens, void* stream) {
  SC_CHECK_ARG(B > 0 && T >= 0 && V > 0 && B * T < ((int64_t)1 << 31), SC_E_SHAPE);
  SC_CHECK_ARG(in_lens && out_lens && (T == 0 || (logits && pred && out_tokens)), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (T > 0) {
    const unsigned grid = (unsigned)cdiv(B * T, 8);
    if (dtype == SC_F32) argmax_rows_kernel<float><<<grid, 256, 0, st>>>((const float*)logits, stride_b, stride_t, in_lens, (int)B, (int)T, (int)V, pred);
    else if (dtype == SC_BF16) argmax_rows_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)logits, stride_b, stride_t, in_lens, (int)B, (int)T, (int)V, pred);
    else return SC_E_DTYPE;
  }
  ctc_collapse_kernel<<<(unsigned)B, 256, 0, st>>>(pred, in_lens, (int)T, blank, out_tokens, out_lens);
  SC_LAUNCH_RET();
}
