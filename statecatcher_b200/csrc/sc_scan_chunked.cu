// K2, few-streams variant: the fused LucyRNN forward scan parallelised over TIME.
//
// sc_lucy_scan_fwd gives every (stream, channel) one thread that walks the T timesteps of a
// segment in order: with 64 streams that is 65 536 chains, enough to saturate HBM, but a
// single live stream (streaming inference, configs[4]) leaves 4 CTAs walking 3000 dependent
// steps — 0.28 ms per layer of pure latency.  Both recurrences of the layer are affine in
// their state (lucyrnn.py:153-166 / 172-184),
//     S_t = d_t S_{t-1} + k_t v_t            h_t = z_t h_{t-1} + (1 - z_t) c_t,   c_t = tanh(p_t + S'_t),
// so a chunk of LC steps composes into one affine map (A, B) per channel regardless of the
// incoming state.  Three launches over a grid of (stream, 128-channel block, chunk):
//   1. chunk composites of the S recurrence                      (reads q, k, v)
//   2. S at each chunk start by a <= T/LC-step prefix over the composites (recomputed by every
//      CTA for itself), then the chunk's h-recurrence composites  (reads all five gates)
//   3. h at each chunk start likewise, then the chunk's outputs h_t, the checkpoints of S the
//      backward consumes, and the carried states.
// The gates are read three times instead of once, but for the shapes this path takes they
// are a few tens of MB and stay in L2; the arithmetic is re-associated across chunk
// boundaries only (products of at most LC decays), ~1e-6 relative in fp32.
#include "sc_common.cuh"

namespace sc {

constexpr int CH_LC = 64;            // timesteps per chunk
constexpr int CH_CB = 128;           // channels per CTA (one per thread)
constexpr int CH_CK = SC_SCAN_CKPT;

template <typename T, bool PRECISE>
struct GateRow {
  const T* g; int64_t H;
  __device__ __forceinline__ float z() const { return ld_f(g + (int64_t)SC_GATE_Z * H); }
  __device__ __forceinline__ float k() const { return ld_f(g + (int64_t)SC_GATE_K * H); }
  __device__ __forceinline__ float v() const { return ld_f(g + (int64_t)SC_GATE_V * H); }
  __device__ __forceinline__ float p() const { return ld_f(g + (int64_t)SC_GATE_P * H); }
  __device__ __forceinline__ float q() const { return ld_f(g + (int64_t)SC_GATE_Q * H); }
};

// work layout (fp32): [4][B][NC][H] = A_S, B_S, A_h, B_h
// PASS 1: S composites.  PASS 2: h composites.  PASS 3: outputs.
template <typename T, bool TRAIN, bool PRECISE, int PASS>
__global__ void __launch_bounds__(CH_CB)
lucy_scan_chunk_kernel(const T* __restrict__ G, int64_t ldg, const float* __restrict__ h0,
                       const float* __restrict__ s0, T* __restrict__ Hout, int64_t ldh,
                       float* __restrict__ hT, float* __restrict__ sT, float* __restrict__ Sckpt,
                       float* __restrict__ work, int B, int Tn, int H, int NC) {
  const int ch = blockIdx.x * CH_CB + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (ch >= H) return;
  const int64_t comp = (int64_t)B * NC * H;                     // one composite array
  float* AS = work, *BS = work + comp, *AH = work + 2 * comp, *BH = work + 3 * comp;
  const int64_t cbase = ((int64_t)b * NC) * H + ch;             // + chunk * H
  const int t0 = c * CH_LC, t1 = min(t0 + CH_LC, Tn);
  const T* g = G + ((int64_t)b * Tn + t0) * ldg + ch;
  if (PASS == 1) {
    float A = 1.f, Bv = 0.f;
    for (int t = t0; t < t1; ++t, g += ldg) {
      const GateRow<T, PRECISE> r{g, H};
      const float d = sigmoidf_<PRECISE>(r.q());
      A *= d;
      Bv = fmaf(d, Bv, r.k() * r.v());
    }
    AS[cbase + (int64_t)c * H] = A;
    BS[cbase + (int64_t)c * H] = Bv;
    return;
  }
  // S entering this chunk: prefix over the earlier chunks' composites
  float S = TRAIN ? 0.f : s0[(int64_t)b * H + ch];
  for (int cc = 0; cc < c; ++cc) S = fmaf(AS[cbase + (int64_t)cc * H], S, BS[cbase + (int64_t)cc * H]);
  if (PASS == 2) {
    float A = 1.f, Bv = 0.f;
    for (int t = t0; t < t1; ++t, g += ldg) {
      const GateRow<T, PRECISE> r{g, H};
      const float d = sigmoidf_<PRECISE>(r.q());
      const float kv = r.k() * r.v();
      S = fmaf(d, S, kv);
      const float sp = TRAIN ? fmaf(d, S, kv) : S;
      const float cv = tanhf_<PRECISE>(r.p() + sp);
      const float zh = sigmoidf_<PRECISE>(r.z());
      A *= zh;
      Bv = fmaf(zh, Bv - cv, cv);                              // same form as the sequential kernel's h update
    }
    AH[cbase + (int64_t)c * H] = A;
    BH[cbase + (int64_t)c * H] = Bv;
    if (!TRAIN && sT != nullptr && c == NC - 1) sT[(int64_t)b * H + ch] = S;
    return;
  }
  // PASS 3
  float h = h0[(int64_t)b * H + ch];
  for (int cc = 0; cc < c; ++cc) h = fmaf(AH[cbase + (int64_t)cc * H], h, BH[cbase + (int64_t)cc * H]);
  T* ho = Hout + ((int64_t)b * Tn + t0) * ldh + ch;
  const int nck = (Tn + CH_CK - 1) / CH_CK;
  for (int t = t0; t < t1; ++t, g += ldg, ho += ldh) {
    if (Sckpt != nullptr && (t % CH_CK) == 0) Sckpt[((int64_t)b * nck + t / CH_CK) * H + ch] = S;
    const GateRow<T, PRECISE> r{g, H};
    const float d = sigmoidf_<PRECISE>(r.q());
    const float kv = r.k() * r.v();
    S = fmaf(d, S, kv);
    const float sp = TRAIN ? fmaf(d, S, kv) : S;
    const float cv = tanhf_<PRECISE>(r.p() + sp);
    const float zh = sigmoidf_<PRECISE>(r.z());
    h = fmaf(zh, h - cv, cv);
    st_f(ho, h);
  }
  if (c == NC - 1) hT[(int64_t)b * H + ch] = h;
}

template <typename T, bool PRECISE>
static int launch_chunked(const void* G, int64_t ldg, const float* h0, const float* s0, void* Hout, int64_t ldh,
                          float* hT, float* sT, float* Sckpt, float* work, int64_t B, int64_t Tn, int64_t H, int train,
                          cudaStream_t st) {
  const int NC = (int)cdiv(Tn, CH_LC);
  const dim3 grid((unsigned)cdiv(H, CH_CB), (unsigned)NC, (unsigned)B);
#define SC_CHUNK(TR, PASS) lucy_scan_chunk_kernel<T, TR, PRECISE, PASS><<<grid, CH_CB, 0, st>>>( \
      (const T*)G, ldg, h0, s0, (T*)Hout, ldh, hT, sT, Sckpt, work, (int)B, (int)Tn, (int)H, NC)
  if (train) { SC_CHUNK(true, 1); SC_CHUNK(true, 2); SC_CHUNK(true, 3); }
  else { SC_CHUNK(false, 1); SC_CHUNK(false, 2); SC_CHUNK(false, 3); }
#undef SC_CHUNK
  SC_LAUNCH_RET();
}

}  // namespace sc

using namespace sc;

// Bytes of fp32 workspace the chunked scan needs, or 0 when the sequential kernel is the better
// choice (enough (stream, channel-block) CTAs to fill the machine, or a segment too short to cut).
extern "C" int64_t sc_lucy_scan_chunked_work_bytes(int64_t B, int64_t T, int64_t H) {
  if (B < 1 || T < 4 * CH_LC || H < 1) return 0;
  static const int force = [] { const char* e = getenv("SC_SCAN_CHUNKED"); return e ? (e[0] == '0' ? 0 : 1) : -1; }();
  if (force == 0) return 0;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // the sequential TMA kernel runs B * H/256 CTAs; below ~1/4 of the SMs it is latency-bound
  if (force != 1 && B * cdiv(H, 256) * 4 > sms) return 0;
  if (B > 65535 || cdiv(T, CH_LC) > 65535) return 0;
  return 4 * B * cdiv(T, CH_LC) * H * (int64_t)sizeof(float);
}

extern "C" int sc_lucy_scan_fwd_chunked(const void* G, int64_t ldg, const float* h0, const float* s0,
                                        void* Hout, int64_t ldh, float* hT, float* sT, float* Sckpt,
                                        float* work, int64_t B, int64_t T, int64_t H, int dtype, int train_mode,
                                        void* stream) {
  SC_CHECK_ARG(h0 && hT && work && (train_mode || s0), SC_E_BADARG);
  SC_CHECK_ARG(B > 0 && B <= 65535 && T > 0 && H > 0 && cdiv(T, CH_LC) <= 65535 && B * T < ((int64_t)1 << 31), SC_E_SHAPE);
  SC_CHECK_ARG(G && Hout, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == SC_BF16) return launch_chunked<bf16, false>(G, ldg, h0, s0, Hout, ldh, hT, sT, Sckpt, work, B, T, H, train_mode, st);
  if (dtype == SC_F32) return launch_chunked<float, true>(G, ldg, h0, s0, Hout, ldh, hT, sT, Sckpt, work, B, T, H, train_mode, st);
  return SC_E_DTYPE;
}
