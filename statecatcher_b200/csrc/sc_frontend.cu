// Frontend: the step in front of the hot path.
//
// Replaces make_frontend (model.py:250-279) — torchaudio MFCC(n_mfcc=80, dct ortho, log_mels) or
// MelSpectrogram + AmplitudeToDB(top_db=80) with n_fft = win = 400, hop = 160, 80 htk mel bands,
// center=False, power 2 — together with train.py:473-475's transpose to (B, T, 80), and
// compute_frame_mask + the in_lens line (train.py:296-306, 486-490).
//
// One fused kernel per segment batch, one CTA per 32 consecutive frames of a stream:
//   samples -> shared memory (each sample is read from HBM once although frames overlap 2.5x)
//   -> Hann window and TWO radix-2 folds of the real 400-point DFT (x[n] with x[400-n], then
//      n with 200-n): four dense ~100x100 real transforms instead of a 400x402 one, 40.6 k MAC
//      per frame instead of 160.8 k, fp32 FMA with 16 frames x 4 bins of accumulators per thread
//   -> |X|^2 -> 80 triangular mel bands (sparse: a band touches <= 32 bins) -> log
//   -> 80x80 DCT-II (mfcc) or 10 log10 + batch maximum (mel) -> (B, T, 80) fp32.
// The spectrogram, the mel spectrogram and the (B, 80, T) tensor torchaudio materialises never
// reach HBM: algorithmic bytes per frame = 160 samples x 4 B in + 80 x 4 B out = 960 B.
// Tensor cores are deliberately not used: the power spectrum of speech spans > 60 dB inside one
// frame and bf16/tf32 operands would put the weak bins' error at 1e-2 of their value.
#include "sc_common.cuh"
#include "sc_tma.cuh"

#include <math.h>

namespace sc {

constexpr int FE_NFFT = 400, FE_HOP = 160, FE_NMEL = 80, FE_NMFCC = 80, FE_NFREQ = 201;
constexpr int FE_FT = 32;              // frames per CTA
constexpr int FE_THREADS = 224;        // 7 warps: 4 transforms x 52 threads do the DFT stage, all 224 the rest
constexpr int FE_NB = 104;             // padded size of one folded transform (101 rounded up to a multiple of 8)
constexpr int FE_MAXW = 32;            // bins per mel band (at most)
constexpr int FE_VP = 36;              // row pitch (frames) of the folded inputs / power spectrum: 16-byte rows, 4-way instead of 32-way conflicts on the fold's stores
constexpr int FE_STAGE = 4 * 8 * FE_NB; // floats per basis stage: 8 rows of each of the 4 transforms
// table layout (floats)
constexpr int FE_OFF_WIN = 0;
constexpr int FE_OFF_BAS = 400;                                  // [13 stages][4 transforms][8 rows][104 bins]
constexpr int FE_OFF_MELW = FE_OFF_BAS + 4 * FE_NB * FE_NB;      // [80][32]
constexpr int FE_OFF_MELLO = FE_OFF_MELW + FE_NMEL * FE_MAXW;    // [80] int: first bin
constexpr int FE_OFF_MELCNT = FE_OFF_MELLO + FE_NMEL;            // [80] int: number of bins
constexpr int FE_OFF_DCT = FE_OFF_MELCNT + FE_NMEL;              // [80][80]
constexpr int FE_TABLE_LEN = FE_OFF_DCT + FE_NMEL * FE_NMFCC;
constexpr int FE_XS = (FE_FT - 1) * FE_HOP + FE_NFFT;            // samples one CTA needs

__device__ __forceinline__ unsigned enc_ordered(float f) {      // monotone float -> unsigned
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// mode 0: mfcc; mode 1: mel dB before the top_db floor (batch maximum -> gmax, ordered encoding)
__global__ void __launch_bounds__(FE_THREADS, 2)
frontend_kernel(const float* __restrict__ wav, int64_t ldw, int S, int T,
                const float* __restrict__ tab, int mode, const uint8_t* __restrict__ row_mask, int64_t ldmask,
                float* __restrict__ out, int64_t out_stride_b, unsigned* __restrict__ gmax) {
  extern __shared__ __align__(128) float sm[];
  __shared__ __align__(8) uint64_t bbar[2];
  float* xs = sm;                                   // [FE_XS] samples; later the log-mel rows L[80][32]
  float* V = sm + FE_XS;                            // [4][104][36] folded inputs; later the power spectrum P[2][104][36]
  float* BS = V + 4 * FE_NB * FE_VP;                // 2 stages x [4][8][104] basis rows
  const int b = blockIdx.y, t0 = blockIdx.x * FE_FT, tid = threadIdx.x, lane = tid & 31;
  const int nf = min(FE_FT, T - t0);                // live frames of this tile
  // the folded DFT bases (173 KB) stream through shared memory in 13 stages of 8 rows x 4
  // transforms (one 13 KB bulk copy each, double-buffered): read from L2 once per CTA with the
  // latency off the multiply-add loop
  auto issue = [&](int st) {
    const uint32_t bar = smem_u32(&bbar[st & 1]);
    mbar_expect_tx(bar, FE_STAGE * 4u);
    bulk_load_1d(smem_u32(BS + (st & 1) * FE_STAGE), tab + FE_OFF_BAS + (size_t)st * FE_STAGE, FE_STAGE * 4u, bar);
  };
  if (tid == 0) {
    mbar_init(smem_u32(&bbar[0]), 1);
    mbar_init(smem_u32(&bbar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    issue(0);
    issue(1);
  }
  const float* w_b = wav + (int64_t)b * ldw;
  const int s0 = t0 * FE_HOP;
  {
    constexpr int NL = (FE_XS + FE_THREADS - 1) / FE_THREADS;     // 24 loads per thread, all in flight at once
    float tmp[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int i = tid + j * FE_THREADS;
      tmp[j] = (i < FE_XS && s0 + i < S) ? __ldg(w_b + s0 + i) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int i = tid + j * FE_THREADS;
      if (i < FE_XS) xs[i] = tmp[j];
    }
  }
  __syncthreads();
  // ---- window + two folds.  A warp takes 32 consecutive n of one frame: frames start 160
  // floats apart (a multiple of the 32 banks), so lanes must differ in n, not in the frame --------
  const float* win = tab + FE_OFF_WIN;
  if (tid < 2 * FE_NB) {
    const int n = tid % FE_NB, fh = tid / FE_NB;    // thread = one n, 16 frames; its 4 window values stay in registers
    // v0..v3 = c0*a(n) + c1*a(400-n) + c2*a(200-n) + c3*a(200+n) patterns; n = 0 and n >= 100 are the edge rows
    const bool mid = n >= 1 && n < 100;
    const int i0 = n <= 100 ? n : 0, i1 = mid ? 400 - n : (n == 100 ? 300 : 200), i2 = mid ? 200 - n : 0, i3 = mid ? 200 + n : 0;
    const float w0 = n <= 100 ? __ldg(win + i0) : 0.f;
    const float w1 = n <= 100 ? __ldg(win + i1) : 0.f;
    const float w2 = mid ? __ldg(win + i2) : 0.f, w3 = mid ? __ldg(win + i3) : 0.f;
    float* vp = V + n * FE_VP + fh * 16;
#pragma unroll 4
    for (int ff = 0; ff < 16; ++ff) {
      const float* x = xs + (fh * 16 + ff) * FE_HOP;
      const float a = x[i0] * w0, ar = x[i1] * w1, c = x[i2] * w2, cr = x[i3] * w3;
      float v0, v1, v2, v3;
      if (mid) {
        const float e1 = a + ar, e2 = c + cr, o1 = a - ar, o2 = c - cr;
        v0 = e1 + e2; v1 = e1 - e2; v2 = o1 - o2; v3 = o1 + o2;
      } else if (n == 0) {                          // a = a(0), ar = a(200)
        v0 = a + ar; v1 = a - ar; v2 = 0.f; v3 = 0.f;
      } else {                                      // n = 100: a(100), a(300); rows 101..103 are zero (w = 0)
        v0 = a + ar; v1 = 0.f; v2 = 0.f; v3 = a - ar;
      }
      vp[0 * FE_NB * FE_VP + ff] = v0;
      vp[1 * FE_NB * FE_VP + ff] = v1;
      vp[2 * FE_NB * FE_VP + ff] = v2;
      vp[3 * FE_NB * FE_VP + ff] = v3;
    }
  }
  __syncthreads();
  // ---- four 104x104 transforms: 52 threads each, thread = 4 bins x 16 frames.  Accumulators are
  // frame PAIRS so that every multiply-add is one packed FFMA2 (fma.rn.f32x2, sm_100): half the
  // FMA-pipe issue slots of scalar FFMA ----------------------------------------------------------
  const int q = tid / 52, r = tid - q * 52;         // tid < 208 active
  const int kg = r % 26, fg = r / 26;
  const bool active = tid < 208;
  float2 acc[8][4];                                 // [frame pair][bin]
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = make_float2(0.f, 0.f); }
  for (int st = 0; st < FE_NB / 8; ++st) {
    if (tid == 0 && st >= 1 && st + 1 < FE_NB / 8) issue(st + 1);   // its buffer was released by the barrier ending stage st-1
    mbar_wait(smem_u32(&bbar[st & 1]), (uint32_t)((st >> 1) & 1));
    if (active) {
      const float4* bs = reinterpret_cast<const float4*>(BS + (st & 1) * FE_STAGE + q * (8 * FE_NB)) + kg;
      const float4* vq = reinterpret_cast<const float4*>(V + ((size_t)q * FE_NB + st * 8) * FE_VP + fg * 16);
#pragma unroll 2
      for (int rr = 0; rr < 8; ++rr) {
        const float4 bv = bs[rr * (FE_NB / 4)];
        const float2 b0 = make_float2(bv.x, bv.x), b1 = make_float2(bv.y, bv.y);
        const float2 b2 = make_float2(bv.z, bv.z), b3 = make_float2(bv.w, bv.w);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t4 = vq[rr * (FE_VP / 4) + j];
          const float2 x0 = make_float2(t4.x, t4.y), x1 = make_float2(t4.z, t4.w);
          acc[2 * j][0] = __ffma2_rn(x0, b0, acc[2 * j][0]);
          acc[2 * j][1] = __ffma2_rn(x0, b1, acc[2 * j][1]);
          acc[2 * j][2] = __ffma2_rn(x0, b2, acc[2 * j][2]);
          acc[2 * j][3] = __ffma2_rn(x0, b3, acc[2 * j][3]);
          acc[2 * j + 1][0] = __ffma2_rn(x1, b0, acc[2 * j + 1][0]);
          acc[2 * j + 1][1] = __ffma2_rn(x1, b1, acc[2 * j + 1][1]);
          acc[2 * j + 1][2] = __ffma2_rn(x1, b2, acc[2 * j + 1][2]);
          acc[2 * j + 1][3] = __ffma2_rn(x1, b3, acc[2 * j + 1][3]);
        }
      }
    }
    __syncthreads();
  }
  // the basis ring is free: fetch the 80x80 DCT matrix into it while the power spectrum and the mel bands are formed
  if (tid == 0 && mode == 0) {
    const uint32_t bar = smem_u32(&bbar[1]);        // its 7th use: 13 stages -> slot 1 completed phases 0..5
    mbar_expect_tx(bar, FE_NMEL * FE_NMFCC * 4u);
    bulk_load_1d(smem_u32(BS), tab + FE_OFF_DCT, FE_NMEL * FE_NMFCC * 4u, bar);
  }
  // V is dead (every thread is past the last stage's barrier): its space becomes the power
  // spectrum, row = parity*104 + k/2, 16 frames of a thread stored as four float4
  float* P = V;
  if (active && q < 2) {                            // real parts: q0 -> even bins, q1 -> odd bins
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4* p4 = reinterpret_cast<float4*>(P + (q * FE_NB + 4 * kg + c) * FE_VP + fg * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) p4[j] = make_float4(acc[2 * j][c].x, acc[2 * j][c].y, acc[2 * j + 1][c].x, acc[2 * j + 1][c].y);
    }
  }
  __syncthreads();
  if (active && q >= 2) {                           // imaginary parts complete |X|^2
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float4* p4 = reinterpret_cast<float4*>(P + ((q - 2) * FE_NB + 4 * kg + c) * FE_VP + fg * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 re = p4[j];
        const float2 i0 = acc[2 * j][c], i1 = acc[2 * j + 1][c];
        p4[j] = make_float4(fmaf(re.x, re.x, i0.x * i0.x), fmaf(re.y, re.y, i0.y * i0.y),
                            fmaf(re.z, re.z, i1.x * i1.x), fmaf(re.w, re.w, i1.y * i1.y));
      }
    }
  }
  __syncthreads();
  // ---- mel bands + log: lane = frame (conflict-free), L[m][f] ---------------------------
  float* L = xs;                                    // [80][32]
  const float* melw = tab + FE_OFF_MELW;
  const int* mello = reinterpret_cast<const int*>(tab + FE_OFF_MELLO);
  const int* melcnt = reinterpret_cast<const int*>(tab + FE_OFF_MELCNT);
  float lmax = -INFINITY;
  for (int idx = tid; idx < FE_FT * FE_NMEL; idx += FE_THREADS) {
    const int m = idx >> 5, f = idx & (FE_FT - 1);
    const int lo = __ldg(mello + m), cnt = __ldg(melcnt + m);
    float acc_m = 0.f;
    const float* wm = melw + m * FE_MAXW;
    {                                               // bins of the band's first parity, then the other: consecutive rows each
      const float* pe = P + ((lo & 1) * FE_NB + (lo >> 1)) * FE_VP + f;
      for (int i = 0; i < cnt; i += 2, pe += FE_VP) acc_m = fmaf(*pe, __ldg(wm + i), acc_m);
      const int k1 = lo + 1;
      const float* po = P + ((k1 & 1) * FE_NB + (k1 >> 1)) * FE_VP + f;
      for (int i = 1; i < cnt; i += 2, po += FE_VP) acc_m = fmaf(*po, __ldg(wm + i), acc_m);
    }
    float v;
    if (mode == 0) v = logf(acc_m + 1e-6f);
    else { v = 10.f * log10f(fmaxf(acc_m, 1e-10f)); if (f < nf) lmax = fmaxf(lmax, v); }
    L[idx] = v;
  }
  __syncthreads();
  float* o_b = out + (int64_t)b * out_stride_b + (int64_t)t0 * FE_NMFCC;
  // optional frame mask (model.py:377 `feats * mask` folded in): masked frames are written as zeros
  const uint8_t* mk = row_mask ? row_mask + (int64_t)b * ldmask + t0 : nullptr;
  if (mode == 0) {
    // ---- DCT-II: thread = coefficient c and 16 frames (8 frame pairs, packed FMAs) -----
    const int c = tid % FE_NMFCC, f0 = (tid / FE_NMFCC) * 16;     // 160 threads busy
    if (tid < 2 * FE_NMFCC) {
      mbar_wait(smem_u32(&bbar[1]), 0u);            // slot 1: stages 1,3,..,11 were phases 0..5 -> this copy is phase 6 (parity 0)
      const float* dct = BS + c;
      float2 a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = make_float2(0.f, 0.f);
#pragma unroll 4
      for (int m = 0; m < FE_NMEL; ++m) {
        const float d = dct[m * FE_NMFCC];
        const float2 d2 = make_float2(d, d);
        const float4* l4 = reinterpret_cast<const float4*>(L + m * FE_FT + f0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 t4 = l4[j];
          a[2 * j] = __ffma2_rn(make_float2(t4.x, t4.y), d2, a[2 * j]);
          a[2 * j + 1] = __ffma2_rn(make_float2(t4.z, t4.w), d2, a[2 * j + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (f0 + 2 * j < nf) o_b[(f0 + 2 * j) * FE_NMFCC + c] = (mk && !mk[f0 + 2 * j]) ? 0.f : a[j].x;
        if (f0 + 2 * j + 1 < nf) o_b[(f0 + 2 * j + 1) * FE_NMFCC + c] = (mk && !mk[f0 + 2 * j + 1]) ? 0.f : a[j].y;
      }
    }
  } else {
    for (int idx = tid; idx < nf * FE_NMEL; idx += FE_THREADS) {
      const int f = idx / FE_NMEL, m = idx - f * FE_NMEL;
      o_b[idx] = L[m * FE_FT + f];
    }
    lmax = warp_max(lmax);
    if (lane == 0 && lmax > -INFINITY) atomicMax(gmax, enc_ordered(lmax));
  }
}

// AmplitudeToDB's top_db: x = max(x, batch_max - top_db)
__global__ void db_floor_kernel(float* __restrict__ x, int64_t n, const unsigned* __restrict__ gmax, float top_db) {
  const float floor_v = dec_ordered(*gmax) - top_db;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = fmaxf(x[i], floor_v);
}

// compute_frame_mask + in_lens: one block per stream, one warp per frame in turn
__global__ void __launch_bounds__(1024)
frame_mask_kernel(const uint8_t* __restrict__ mask, int64_t ldm, int S, int T, int sub, float subsample, int nfeat,
                  uint8_t* __restrict__ frame_mask, int64_t* __restrict__ in_lens) {
  __shared__ int total;
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const uint8_t* m = mask + (int64_t)b * ldm;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  int cnt = 0;
  for (int t = warp; t < T; t += nw) {
    int c = 0;
    for (int i = lane; i < sub; i += 32) c += m[(int64_t)t * sub + i] != 0;
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) { frame_mask[(int64_t)b * T + t] = c > 0; cnt += c; }
  }
  // samples past T*sub count for in_lens but belong to no frame
  for (int64_t i = (int64_t)T * sub + threadIdx.x; i < S; i += blockDim.x) cnt += m[i] != 0;
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  if (lane == 0 && cnt) atomicAdd(&total, cnt);
  __syncthreads();
  if (threadIdx.x == 0) {
    // (mask.sum(1) / subsample).clamp(max=nfeat).long() with the division in fp32 as torch does
    const float qv = fminf(__fdiv_rn((float)total, subsample), (float)nfeat);
    in_lens[b] = (int64_t)qv;
  }
}

}  // namespace sc

using namespace sc;

extern "C" int64_t sc_frontend_tables_len(void) { return FE_TABLE_LEN; }

// Host-side (no CUDA): window, folded DFT bases, sparse mel bands and the DCT matrix, computed in
// double and rounded to fp32.  The caller uploads the buffer once and passes it to sc_frontend.
extern "C" int sc_frontend_tables(float* out, int64_t n, int sample_rate) {
  SC_CHECK_ARG(out && n >= FE_TABLE_LEN && sample_rate >= 2, SC_E_BADARG);
  const double PI = 3.14159265358979323846;
  for (int64_t i = 0; i < FE_TABLE_LEN; ++i) out[i] = 0.f;
  for (int i = 0; i < FE_NFFT; ++i) out[FE_OFF_WIN + i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / FE_NFFT));
  // bases: angle 2*pi*k*n/400 reduced exactly in integers before the trig call
  auto ang = [&](int k, int nn) { return 2.0 * PI * (double)((k * nn) % FE_NFFT) / FE_NFFT; };
  for (int nn = 0; nn <= 100; ++nn) {
    for (int j = 0; j <= 100; ++j) {
      auto at = [&](int qq) { return out + FE_OFF_BAS + (((nn / 8) * 4 + qq) * 8 + (nn % 8)) * FE_NB + j; };
      float *b0 = at(0), *b1 = at(1), *b2 = at(2), *b3 = at(3);
      *b0 = (float)cos(ang(2 * j, nn));                                   // Re, even bins k = 2j, n = 0..100
      if (nn < 100 && j < 100) *b1 = (float)cos(ang(2 * j + 1, nn));      // Re, odd bins, n = 0..99
      if (nn >= 1 && nn < 100 && j >= 1 && j < 100) *b2 = (float)(-sin(ang(2 * j, nn)));   // Im, even bins, n = 1..99
      if (nn >= 1 && j < 100) *b3 = (float)(-sin(ang(2 * j + 1, nn)));    // Im, odd bins, n = 1..100
    }
  }
  // htk mel bands (torchaudio.functional.melscale_fbanks, norm=None), kept sparse
  const double f_max = (double)(sample_rate / 2);
  const double m_max = 2595.0 * log10(1.0 + f_max / 700.0);
  double f_pts[FE_NMEL + 2];
  for (int i = 0; i < FE_NMEL + 2; ++i) f_pts[i] = 700.0 * (pow(10.0, (m_max * i / (FE_NMEL + 1)) / 2595.0) - 1.0);
  int* lo_t = reinterpret_cast<int*>(out + FE_OFF_MELLO);
  int* cnt_t = reinterpret_cast<int*>(out + FE_OFF_MELCNT);
  for (int m = 0; m < FE_NMEL; ++m) {
    int lo = -1, hi = -1;
    for (int k = 0; k < FE_NFREQ; ++k) {
      const double fr = (double)(sample_rate / 2) * k / (FE_NFREQ - 1);
      const double down = (fr - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      const double up = (f_pts[m + 2] - fr) / (f_pts[m + 2] - f_pts[m + 1]);
      const double wv = fmax(0.0, fmin(down, up));
      if (wv > 0.0) {
        if (lo < 0) lo = k;
        hi = k;
      }
    }
    const int cnt = lo < 0 ? 0 : hi - lo + 1;
    if (cnt > FE_MAXW) return SC_E_SHAPE;
    lo_t[m] = lo < 0 ? 0 : lo;
    cnt_t[m] = cnt;
    for (int i = 0; i < cnt; ++i) {
      const double fr = (double)(sample_rate / 2) * (lo + i) / (FE_NFREQ - 1);
      const double down = (fr - f_pts[m]) / (f_pts[m + 1] - f_pts[m]);
      const double up = (f_pts[m + 2] - fr) / (f_pts[m + 2] - f_pts[m + 1]);
      out[FE_OFF_MELW + m * FE_MAXW + i] = (float)fmax(0.0, fmin(down, up));
    }
  }
  // DCT-II, ortho (torchaudio.functional.create_dct): D[m][c]
  for (int m = 0; m < FE_NMEL; ++m)
    for (int c = 0; c < FE_NMFCC; ++c) {
      double d = cos(PI / FE_NMEL * (m + 0.5) * c) * sqrt(2.0 / FE_NMEL);
      if (c == 0) d *= 1.0 / sqrt(2.0);
      out[FE_OFF_DCT + m * FE_NMFCC + c] = (float)d;
    }
  return 0;
}

extern "C" int sc_frontend(const float* wav, int64_t ldw, int64_t B, int64_t S, const float* tables, int mode,
                           float top_db, const uint8_t* frame_mask, int64_t ldmask,
                           float* out, int64_t out_stride_b, unsigned int* gmax, void* stream) {
  SC_CHECK_ARG(B > 0 && S >= 0 && (mode == 0 || mode == 1), SC_E_BADARG);
  SC_CHECK_ARG(S < ((int64_t)1 << 31) && B < 65536, SC_E_SHAPE);
  if (S < FE_NFFT) return 0;                                     // no frame fits
  SC_CHECK_ARG(wav && tables && out && (mode == 0 || gmax), SC_E_BADARG);
  const int T = 1 + (int)((S - FE_NFFT) / FE_HOP);
  SC_CHECK_ARG(out_stride_b >= (int64_t)T * FE_NMFCC && ldw >= S, SC_E_BADARG);
  SC_CHECK_ARG(!frame_mask || (mode == 0 && ldmask >= T), SC_E_BADARG);   // the dB floor of the mel path must see unmasked values
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (FE_XS + 4 * FE_NB * FE_VP + 2 * FE_STAGE) * sizeof(float);
  static_assert(FE_XS % 4 == 0 && FE_NB % 8 == 0 && FE_VP % 4 == 0, "16-byte alignment of the shared-memory regions");
  static_assert(FE_XS >= FE_FT * FE_NMEL, "log-mel rows must fit in the sample space");
  static_assert(2 * FE_STAGE >= FE_NMEL * FE_NMFCC && ((FE_NB / 8) % 2) == 1, "DCT matrix reuses the basis ring; slot 1 must have completed an even number of phases");
  cudaError_t e = cudaFuncSetAttribute(frontend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (mode == 1) {
    e = cudaMemsetAsync(gmax, 0, sizeof(unsigned), st);           // ordered encoding: 0 is below every float
    if (e != cudaSuccess) return (int)e;
  }
  frontend_kernel<<<dim3((unsigned)cdiv(T, FE_FT), (unsigned)B), FE_THREADS, smem, st>>>(wav, ldw, (int)S, T, tables, mode, frame_mask, ldmask,
                                                                                          out, out_stride_b, gmax);
  if (mode == 1 && top_db >= 0.f) {
    // out rows are dense per stream only when out_stride_b == T*80; floor stream by stream otherwise
    if (out_stride_b == (int64_t)T * FE_NMFCC) {
      db_floor_kernel<<<592, 256, 0, st>>>(out, B * (int64_t)T * FE_NMFCC, gmax, top_db);
    } else {
      for (int64_t b = 0; b < B; ++b)
        db_floor_kernel<<<32, 256, 0, st>>>(out + b * out_stride_b, (int64_t)T * FE_NMFCC, gmax, top_db);
    }
  }
  SC_LAUNCH_RET();
}

extern "C" int sc_frame_mask(const uint8_t* sample_mask, int64_t ldm, int64_t B, int64_t S, int64_t T, int64_t sub,
                             float subsample, int64_t nfeat, uint8_t* frame_mask, int64_t* in_lens, void* stream) {
  SC_CHECK_ARG(B > 0 && S > 0 && T > 0 && sub > 0 && nfeat >= 0 && subsample > 0.f, SC_E_BADARG);
  SC_CHECK_ARG(sample_mask && frame_mask && in_lens && ldm >= S, SC_E_BADARG);
  SC_CHECK_ARG(S < ((int64_t)1 << 31) && T * sub <= S, SC_E_SHAPE);
  frame_mask_kernel<<<(unsigned)B, 1024, 0, (cudaStream_t)stream>>>(sample_mask, ldm, (int)S, (int)T, (int)sub, subsample,
                                                                    (int)nfeat, frame_mask, in_lens);
  SC_LAUNCH_RET();
}
