// K2: the LucyRNN recurrent scan, forward and reverse-time backward (sm_100a).
//
// What it replaces (reference, pure PyTorch/Triton): the decay scan loop lucyrnn.py:153-158
// (Triton twin lucyrnn_triton.py:158-177), the per-timestep cell loop lucyrnn.py:160-166 ->
// LucyRNNCell.forward lucyrnn.py:44-70, and the step path lucyrnn.py:172-184.
//
// Design (B200-first, not a port): the recurrence is linear and diagonal, so one thread owns
// VEC adjacent channels of one stream for the whole segment and keeps S and h in registers;
// a warp therefore touches one contiguous 128-byte line per gate per timestep (fully
// coalesced), gate rows are streamed with L1-bypassing loads software-pipelined U timesteps
// ahead, and nothing but the gates is read and nothing but h_t is written: 6H elements per
// frame forward, 12H backward — exactly the algorithmic traffic of SURVEY.md 8(d).
// The backward recomputes S inside SC_SCAN_CKPT-step intervals from checkpoints the forward
// leaves behind instead of storing S for every timestep.
#include "sc_common.cuh"
#include <stdlib.h>

namespace sc {

constexpr int CK = SC_SCAN_CKPT;

// ------------------------------------------------------------------ fused forward ----
template <typename T, int VEC, int U, bool TRAIN, bool PRECISE>
__global__ void __launch_bounds__(128)
lucy_scan_fwd_kernel(const T* __restrict__ G, int64_t ldg, const float* __restrict__ h0,
                     const float* __restrict__ s0, T* __restrict__ Hout, int64_t ldh,
                     float* __restrict__ hT, float* __restrict__ sT, float* __restrict__ Sckpt,
                     int B, int Tn, int H) {
  static_assert(CK % U == 0, "checkpoint interval must be a multiple of the unroll");
  const int groups = H / VEC;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)B * groups) return;
  const int b = (int)(gid / groups);
  const int ch = (int)(gid % groups) * VEC;
  const T* g = G + (int64_t)b * Tn * ldg + ch;
  T* ho = Hout + (int64_t)b * Tn * ldh + ch;
  const int nck = (Tn + CK - 1) / CK;

  float S[VEC], h[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    h[i] = h0[(int64_t)b * H + ch + i];
    S[i] = TRAIN ? 0.0f : s0[(int64_t)b * H + ch + i];
  }

  Vec<T, VEC> cur[U][5], nxt[U][5];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < Tn) {
#pragma unroll
      for (int gt = 0; gt < 5; ++gt) cur[u][gt] = vload<T, VEC>(g + (int64_t)u * ldg + (int64_t)gt * H);
    }

  for (int t0 = 0; t0 < Tn; t0 += U) {
    // prefetch the next U timesteps while this block is being consumed
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + U + u;
      if (t < Tn) {
#pragma unroll
        for (int gt = 0; gt < 5; ++gt) nxt[u][gt] = vload<T, VEC>(g + (int64_t)t * ldg + (int64_t)gt * H);
      }
    }
    if (Sckpt != nullptr && (t0 % CK) == 0) {
      float* ck = Sckpt + ((int64_t)b * nck + t0 / CK) * H + ch;
#pragma unroll
      for (int i = 0; i < VEC; ++i) ck[i] = S[i];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int t = t0 + u;
      if (t < Tn) {
        float z[VEC], k[VEC], v[VEC], p[VEC], q[VEC], out[VEC];
        unpack(cur[u][SC_GATE_Z], z); unpack(cur[u][SC_GATE_K], k); unpack(cur[u][SC_GATE_V], v);
        unpack(cur[u][SC_GATE_P], p); unpack(cur[u][SC_GATE_Q], q);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float d = sigmoidf_<PRECISE>(q[i]);
          const float kv = k[i] * v[i];
          S[i] = fmaf(d, S[i], kv);
          const float sp = TRAIN ? fmaf(d, S[i], kv) : S[i];
          const float c = tanhf_<PRECISE>(p[i] + sp);
          const float zh = sigmoidf_<PRECISE>(z[i]);
          h[i] = fmaf(zh, h[i] - c, c);            // (1-zh)*c + zh*h
          out[i] = h[i];
        }
        vstore<T, VEC>(ho + (int64_t)t * ldh, pack(out, (T*)nullptr));
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int gt = 0; gt < 5; ++gt) cur[u][gt] = nxt[u][gt];
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    hT[(int64_t)b * H + ch + i] = h[i];
    if (!TRAIN && sT != nullptr) sT[(int64_t)b * H + ch + i] = S[i];
  }
}

// ------------------------------------------------------------------ fused backward ---
// Reverse scans (App. A.3):
//   gamma_t = g_t + zh_{t+1} gamma_{t+1};    sigma_t = (d_t|1) da_t + d_{t+1} sigma_{t+1}
template <typename T, int VEC, bool TRAIN, bool PRECISE>
__global__ void __launch_bounds__(128)
lucy_scan_bwd_kernel(const T* __restrict__ G, int64_t ldg, const T* __restrict__ Hout, int64_t ldh,
                     const float* __restrict__ h0, const float* __restrict__ s0,
                     const float* __restrict__ Sckpt, const T* __restrict__ dHout, int64_t lddh,
                     T* __restrict__ dG, int64_t lddg, float* __restrict__ dbias,
                     int B, int Tn, int H) {
  const int groups = H / VEC;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)B * groups) return;
  const int b = (int)(gid / groups);
  const int ch = (int)(gid % groups) * VEC;
  const T* g = G + (int64_t)b * Tn * ldg + ch;
  const T* ho = Hout + (int64_t)b * Tn * ldh + ch;
  const T* dh = dHout + (int64_t)b * Tn * lddh + ch;
  T* dg = dG + (int64_t)b * Tn * lddg + ch;
  const int nck = (Tn + CK - 1) / CK;

  float gz[VEC], ds[VEC], acc[5][VEC], hfirst[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    gz[i] = 0.f; ds[i] = 0.f;
    hfirst[i] = h0[(int64_t)b * H + ch + i];
#pragma unroll
    for (int gt = 0; gt < 5; ++gt) acc[gt][i] = 0.f;
  }

  for (int j = nck - 1; j >= 0; --j) {
    const int t0 = j * CK;
    float Sin[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      // S before step t0.  The forward saved it; for j==0 it is 0 (train) / s0 (step).
      Sin[i] = Sckpt[((int64_t)b * nck + j) * H + ch + i];
    }
    // pass 1: recompute S_t over the interval (k,v,q stay in registers for pass 2)
    Vec<T, VEC> rk[CK], rv[CK], rq[CK];
    float Sl[CK][VEC];
#pragma unroll
    for (int u = 0; u < CK; ++u) {
      if (t0 + u < Tn) {
        rk[u] = vload<T, VEC>(g + (int64_t)(t0 + u) * ldg + (int64_t)SC_GATE_K * H);
        rv[u] = vload<T, VEC>(g + (int64_t)(t0 + u) * ldg + (int64_t)SC_GATE_V * H);
        rq[u] = vload<T, VEC>(g + (int64_t)(t0 + u) * ldg + (int64_t)SC_GATE_Q * H);
      }
    }
    {
      float S[VEC];
#pragma unroll
      for (int i = 0; i < VEC; ++i) S[i] = Sin[i];
#pragma unroll
      for (int u = 0; u < CK; ++u) {
        if (t0 + u < Tn) {
          float k[VEC], v[VEC], q[VEC];
          unpack(rk[u], k); unpack(rv[u], v); unpack(rq[u], q);
#pragma unroll
          for (int i = 0; i < VEC; ++i) {
            S[i] = fmaf(sigmoidf_<PRECISE>(q[i]), S[i], k[i] * v[i]);
            Sl[u][i] = S[i];
          }
        }
      }
    }
    // pass 2: reverse time
#pragma unroll
    for (int u = CK - 1; u >= 0; --u) {
      const int t = t0 + u;
      if (t < Tn) {
        float z[VEC], k[VEC], v[VEC], p[VEC], q[VEC], go[VEC], hp[VEC];
        unpack(vload<T, VEC>(g + (int64_t)t * ldg + (int64_t)SC_GATE_Z * H), z);
        unpack(vload<T, VEC>(g + (int64_t)t * ldg + (int64_t)SC_GATE_P * H), p);
        unpack(vload<T, VEC>(dh + (int64_t)t * lddh), go);
        if (t > 0) {
          unpack(vload<T, VEC>(ho + (int64_t)(t - 1) * ldh), hp);
        } else {
#pragma unroll
          for (int i = 0; i < VEC; ++i) hp[i] = hfirst[i];
        }
        unpack(rk[u], k); unpack(rv[u], v); unpack(rq[u], q);
        float dz[VEC], dk[VEC], dv[VEC], dp[VEC], dq[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          const float d = sigmoidf_<PRECISE>(q[i]);
          const float kv = k[i] * v[i];
          const float St = Sl[u][i];
          const float Sp = (u > 0) ? Sl[u - 1][i] : Sin[i];
          const float sp = TRAIN ? fmaf(d, St, kv) : St;
          const float c = tanhf_<PRECISE>(p[i] + sp);
          const float zh = sigmoidf_<PRECISE>(z[i]);
          const float gam = go[i] + gz[i];
          const float dzh = gam * (hp[i] - c);
          const float da = gam * (1.f - zh) * (1.f - c * c);
          gz[i] = zh * gam;
          dz[i] = dzh * zh * (1.f - zh);
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
        }
        T* row = dg + (int64_t)t * lddg;
        const Vec<T, VEC> oz = pack(dz, (T*)nullptr), ok = pack(dk, (T*)nullptr),
                          ov = pack(dv, (T*)nullptr), op = pack(dp, (T*)nullptr),
                          oq = pack(dq, (T*)nullptr);
        vstore<T, VEC>(row + (int64_t)SC_GATE_Z * H, oz);
        vstore<T, VEC>(row + (int64_t)SC_GATE_K * H, ok);
        vstore<T, VEC>(row + (int64_t)SC_GATE_V * H, ov);
        vstore<T, VEC>(row + (int64_t)SC_GATE_P * H, op);
        vstore<T, VEC>(row + (int64_t)SC_GATE_Q * H, oq);
        // bias gradient = column sum of what was actually written (rounded values)
        float rz[VEC], rkk[VEC], rvv[VEC], rp[VEC], rqq[VEC];
        unpack(oz, rz); unpack(ok, rkk); unpack(ov, rvv); unpack(op, rp); unpack(oq, rqq);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
          acc[SC_GATE_Z][i] += rz[i]; acc[SC_GATE_K][i] += rkk[i]; acc[SC_GATE_V][i] += rvv[i];
          acc[SC_GATE_P][i] += rp[i]; acc[SC_GATE_Q][i] += rqq[i];
        }
      }
    }
  }
  if (dbias != nullptr) {
#pragma unroll
    for (int gt = 0; gt < 5; ++gt)
#pragma unroll
      for (int i = 0; i < VEC; ++i) atomicAdd(dbias + (int64_t)gt * H + ch + i, acc[gt][i]);
  }
}

// ------------------------------------------------------------------ split scans ------
// General path (LayerNorm / unfused W_h / prefix_sum between the two scans).  One thread per
// channel, coalesced across channels, S saved for the backward.  Simpler than the fused
// kernels: these serve the non-default flag combinations.
template <typename T, bool PRECISE>
__global__ void __launch_bounds__(128)
sscan_fwd_kernel(const T* __restrict__ kk, const T* __restrict__ vv, const T* __restrict__ qq, int64_t ldg,
                 const T* __restrict__ addend, int64_t ldadd, const float* __restrict__ s0,
                 T* __restrict__ A, int64_t lda, float* __restrict__ S_all, float* __restrict__ sT,
                 int B, int Tn, int H, int train, int decay_mode, float lam) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)B * H) return;
  const int b = (int)(gid / H), ch = (int)(gid % H);
  const int64_t r0 = (int64_t)b * Tn;
  float S = train ? 0.f : s0[(int64_t)b * H + ch];
  // prefix_sum mode state (lucyrnn.py:126-142): logw = cumsum(log(exp(-lam t)+1e-7)),
  // num = cumsum(kv*exp(logw)), S_t = num/(exp(logw)+1e-7)
  float logw = 0.f, num = 0.f;
  for (int t = 0; t < Tn; ++t) {
    const int64_t r = r0 + t;
    const float k = ld_f(kk + r * ldg + ch), v = ld_f(vv + r * ldg + ch);
    const float d = sigmoidf_<PRECISE>(ld_f(qq + r * ldg + ch));
    const float kv = k * v;
    // learned decay: S entering every SC_SCAN_CKPT-step interval (the backward recomputes inside an interval)
    if (decay_mode == 0 && (t % SC_SCAN_CKPT) == 0) S_all[((int64_t)b * ((Tn + SC_SCAN_CKPT - 1) / SC_SCAN_CKPT) + t / SC_SCAN_CKPT) * H + ch] = S;
    if (decay_mode == 1) {
      logw += logf(expf(-lam * (float)t) + 1e-7f);
      const float w = expf(logw);
      num = fmaf(kv, w, num);
      S = num / (w + 1e-7f);
    } else {
      S = fmaf(d, S, kv);
    }
    if (decay_mode == 1) S_all[r * H + ch] = S;      // prefix_sum: every S_t
    const float sp = train ? fmaf(d, S, kv) : S;
    st_f(A + r * lda + ch, ld_f(addend + r * ldadd + ch) + sp);
  }
  if (!train && sT != nullptr) sT[(int64_t)b * H + ch] = S;
}

template <typename T, bool PRECISE>
__global__ void __launch_bounds__(128)
sscan_bwd_kernel(const T* __restrict__ kk, const T* __restrict__ vv, const T* __restrict__ qq, int64_t ldg,
                 const float* __restrict__ S_all, const float* __restrict__ s0,
                 const T* __restrict__ dA, int64_t ldda, T* __restrict__ dk, T* __restrict__ dv,
                 T* __restrict__ dq, int64_t lddg, float* __restrict__ dsum, int B, int Tn, int H, int train,
                 int decay_mode, float lam) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)B * H) return;
  const int b = (int)(gid / H), ch = (int)(gid % H);
  const int64_t r0 = (int64_t)b * Tn;
  float ak = 0.f, av = 0.f, aq = 0.f;   // column sums (dsum: [3][H], may be null)
  float ds = 0.f;      // d_{t+1} * sigma_{t+1}
  const int nck = (Tn + SC_SCAN_CKPT - 1) / SC_SCAN_CKPT;
  for (int c = nck - 1; c >= 0; --c) {
    const int t0 = c * SC_SCAN_CKPT;
    const int n = (Tn - t0 < SC_SCAN_CKPT) ? Tn - t0 : SC_SCAN_CKPT;
    const float Sin = S_all[((int64_t)b * nck + c) * H + ch];   // S entering the interval (checkpoint of the forward)
    float Sl[SC_SCAN_CKPT];
    {
      float S = Sin;
#pragma unroll
      for (int u = 0; u < SC_SCAN_CKPT; ++u) {
        if (u < n) {
          const int64_t r = r0 + t0 + u;
          const float d = sigmoidf_<PRECISE>(ld_f(qq + r * ldg + ch));
          S = fmaf(d, S, ld_f(kk + r * ldg + ch) * ld_f(vv + r * ldg + ch));
        }
        Sl[u] = S;
      }
    }
#pragma unroll
    for (int u = SC_SCAN_CKPT - 1; u >= 0; --u) {
      if (u >= n) continue;
      const int64_t r = r0 + t0 + u;
      const float k = ld_f(kk + r * ldg + ch), v = ld_f(vv + r * ldg + ch);
      const float d = sigmoidf_<PRECISE>(ld_f(qq + r * ldg + ch));
      const float da = ld_f(dA + r * ldda + ch);
      const float St = Sl[u];
      const float Sp = (u > 0) ? Sl[u - 1] : Sin;    // Sin of interval 0 is the initial state (0 on the training path)
      float dkv, dd;
      if (train) {
        const float sig = fmaf(d, da, ds);
        dkv = da + sig;
        dd = fmaf(St, da, Sp * sig);
        ds = d * sig;
      } else {
        const float sig = da + ds;
        dkv = sig;
        dd = Sp * sig;
        ds = d * sig;
      }
      const float o1 = dkv * v, o2 = dkv * k, o3 = dd * d * (1.f - d);
      st_f(dk + r * lddg + ch, o1);
      st_f(dv + r * lddg + ch, o2);
      st_f(dq + r * lddg + ch, o3);
      ak += o1; av += o2; aq += o3;
    }
  }
  if (dsum != nullptr) { atomicAdd(dsum + ch, ak); atomicAdd(dsum + H + ch, av); atomicAdd(dsum + 2 * (int64_t)H + ch, aq); }
}

// prefix_sum backward: w_t depends only on t, so a first forward pass rebuilds logw_t into
// registers-free form by recomputation: logw_t is accumulated forwards once to its final
// value and then walked back down by subtracting the same terms.
template <typename T, bool PRECISE>
__global__ void __launch_bounds__(128)
sscan_bwd_prefix_kernel(const T* __restrict__ kk, const T* __restrict__ vv, const T* __restrict__ qq,
                        int64_t ldg, const float* __restrict__ S_all, const T* __restrict__ dA,
                        int64_t ldda, T* __restrict__ dk, T* __restrict__ dv, T* __restrict__ dq,
                        int64_t lddg, float* __restrict__ dsum, int B, int Tn, int H, float lam) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)B * H) return;
  const int b = (int)(gid / H), ch = (int)(gid % H);
  const int64_t r0 = (int64_t)b * Tn;
  float ak = 0.f, av = 0.f, aq = 0.f;
  float logw = 0.f;
  for (int t = 0; t < Tn; ++t) logw += logf(expf(-lam * (float)t) + 1e-7f);
  float racc = 0.f;
  for (int t = Tn - 1; t >= 0; --t) {
    const int64_t r = r0 + t;
    const float k = ld_f(kk + r * ldg + ch), v = ld_f(vv + r * ldg + ch);
    const float d = sigmoidf_<PRECISE>(ld_f(qq + r * ldg + ch));
    const float da = ld_f(dA + r * ldda + ch);
    const float St = S_all[r * H + ch];
    const float w = expf(logw);
    racc += d * da / (w + 1e-7f);              // dL/dnum_t accumulated over tau >= t
    const float dkv = da + w * racc;
    const float dd = St * da;
    const float o1 = dkv * v, o2 = dkv * k, o3 = dd * d * (1.f - d);
    st_f(dk + r * lddg + ch, o1);
    st_f(dv + r * lddg + ch, o2);
    st_f(dq + r * lddg + ch, o3);
    ak += o1; av += o2; aq += o3;
    logw -= logf(expf(-lam * (float)t) + 1e-7f);
  }
  if (dsum != nullptr) { atomicAdd(dsum + ch, ak); atomicAdd(dsum + H + ch, av); atomicAdd(dsum + 2 * (int64_t)H + ch, aq); }
}

template <typename T, bool PRECISE>
__global__ void __launch_bounds__(128)
hscan_fwd_kernel(const T* __restrict__ An, int64_t ldan, const T* __restrict__ Zn, int64_t ldzn,
                 const float* __restrict__ h0, T* __restrict__ Hout, int64_t ldh,
                 float* __restrict__ hT, int B, int Tn, int H) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)B * H) return;
  const int b = (int)(gid / H), ch = (int)(gid % H);
  const int64_t r0 = (int64_t)b * Tn;
  float h = h0[(int64_t)b * H + ch];
  for (int t = 0; t < Tn; ++t) {
    const int64_t r = r0 + t;
    const float c = tanhf_<PRECISE>(ld_f(An + r * ldan + ch));
    const float zh = sigmoidf_<PRECISE>(ld_f(Zn + r * ldzn + ch));
    h = fmaf(zh, h - c, c);
    st_f(Hout + r * ldh + ch, h);
  }
  hT[(int64_t)b * H + ch] = h;
}

template <typename T, bool PRECISE>
__global__ void __launch_bounds__(128)
hscan_bwd_kernel(const T* __restrict__ An, int64_t ldan, const T* __restrict__ Zn, int64_t ldzn,
                 const T* __restrict__ Hout, int64_t ldh, const float* __restrict__ h0,
                 const T* __restrict__ dHout, int64_t lddh, T* __restrict__ dAn, int64_t lddan,
                 T* __restrict__ dZn, int64_t lddzn, int B, int Tn, int H) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (int64_t)B * H) return;
  const int b = (int)(gid / H), ch = (int)(gid % H);
  const int64_t r0 = (int64_t)b * Tn;
  float gz = 0.f;
  for (int t = Tn - 1; t >= 0; --t) {
    const int64_t r = r0 + t;
    const float c = tanhf_<PRECISE>(ld_f(An + r * ldan + ch));
    const float zh = sigmoidf_<PRECISE>(ld_f(Zn + r * ldzn + ch));
    const float hp = (t > 0) ? ld_f(Hout + (r - 1) * ldh + ch) : h0[(int64_t)b * H + ch];
    const float gam = ld_f(dHout + r * lddh + ch) + gz;
    gz = zh * gam;
    st_f(dAn + r * lddan + ch, gam * (1.f - zh) * (1.f - c * c));
    st_f(dZn + r * lddzn + ch, gam * (hp - c) * zh * (1.f - zh));
  }
}

// ------------------------------------------------------------------ dispatch ---------
template <typename T, int VEC, bool PRECISE>
static int launch_scan_fwd(const void* G, int64_t ldg, const float* h0, const float* s0, void* Hout,
                           int64_t ldh, float* hT, float* sT, float* Sckpt, int64_t B, int64_t Tn,
                           int64_t H, int train, cudaStream_t st) {
  const int64_t n = B * (H / VEC);
  const int threads = 128;
  const unsigned blocks = (unsigned)cdiv(n, threads);
  constexpr int U = 8;
  if (train)
    lucy_scan_fwd_kernel<T, VEC, U, true, PRECISE><<<blocks, threads, 0, st>>>(
        (const T*)G, ldg, h0, s0, (T*)Hout, ldh, hT, sT, Sckpt, (int)B, (int)Tn, (int)H);
  else
    lucy_scan_fwd_kernel<T, VEC, U, false, PRECISE><<<blocks, threads, 0, st>>>(
        (const T*)G, ldg, h0, s0, (T*)Hout, ldh, hT, sT, Sckpt, (int)B, (int)Tn, (int)H);
  SC_LAUNCH_RET();
}

template <typename T, int VEC, bool PRECISE>
static int launch_scan_bwd(const void* G, int64_t ldg, const void* Hout, int64_t ldh, const float* h0,
                           const float* s0, const float* Sckpt, const void* dHout, int64_t lddh,
                           void* dG, int64_t lddg, float* dbias, int64_t B, int64_t Tn, int64_t H,
                           int train, cudaStream_t st) {
  const int64_t n = B * (H / VEC);
  const int threads = 128;
  const unsigned blocks = (unsigned)cdiv(n, threads);
  if (train)
    lucy_scan_bwd_kernel<T, VEC, true, PRECISE><<<blocks, threads, 0, st>>>(
        (const T*)G, ldg, (const T*)Hout, ldh, h0, s0, Sckpt, (const T*)dHout, lddh, (T*)dG, lddg,
        dbias, (int)B, (int)Tn, (int)H);
  else
    lucy_scan_bwd_kernel<T, VEC, false, PRECISE><<<blocks, threads, 0, st>>>(
        (const T*)G, ldg, (const T*)Hout, ldh, h0, s0, Sckpt, (const T*)dHout, lddh, (T*)dG, lddg,
        dbias, (int)B, (int)Tn, (int)H);
  SC_LAUNCH_RET();
}

}  // namespace sc

namespace sc {
int scan_fwd_tma_dispatch(const void*, int64_t, const float*, const float*, void*, int64_t, float*, float*, float*,
                          int64_t, int64_t, int64_t, int, int, cudaStream_t);
int scan_bwd_tma_dispatch(const void*, int64_t, const void*, int64_t, const float*, const float*, const void*, int64_t,
                          void*, int64_t, float*, int64_t, int64_t, int64_t, int, int, cudaStream_t);
int sscan_fwd_tma_dispatch(const void*, const void*, const void*, int64_t, const void*, int64_t, const float*, void*, int64_t,
                           float*, float*, int64_t, int64_t, int64_t, int, int, cudaStream_t);
int sscan_bwd_tma_dispatch(const void*, const void*, const void*, int64_t, const float*, const float*, const void*, int64_t,
                           void*, void*, void*, int64_t, float*, int64_t, int64_t, int64_t, int, int, cudaStream_t);
int hscan_fwd_tma_dispatch(const void*, int64_t, const void*, int64_t, const float*, void*, int64_t, float*, int64_t, int64_t,
                           int64_t, int, cudaStream_t);
int hscan_bwd_tma_dispatch(const void*, int64_t, const void*, int64_t, const void*, int64_t, const float*, const void*, int64_t,
                           void*, int64_t, void*, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);
}
using namespace sc;

// SC_SCAN_GENERIC=1 in the environment forces the register-prefetch kernels (A/B measurements)
static bool scan_force_generic() {
  static const bool v = [] { const char* e = getenv("SC_SCAN_GENERIC"); return e && e[0] == '1'; }();
  return v;
}

static bool scan_args_ok(int64_t B, int64_t T, int64_t H) {
  return B > 0 && T >= 0 && H > 0 && B * T < (int64_t)1 << 31 && H < (1 << 24);
}

extern "C" int sc_lucy_scan_fwd(const void* G, int64_t ldg, const float* h0, const float* s0,
                                void* Hout, int64_t ldh, float* hT, float* sT, float* Sckpt,
                                int64_t B, int64_t T, int64_t H, int dtype, int train_mode,
                                void* stream) {
  SC_CHECK_ARG(h0 && hT && (train_mode || s0), SC_E_BADARG);
  SC_CHECK_ARG(scan_args_ok(B, T, H), SC_E_SHAPE);
  SC_CHECK_ARG(T == 0 || (G && Hout), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (!scan_force_generic() && Sckpt != nullptr) {
    const int rc = scan_fwd_tma_dispatch(G, ldg, h0, s0, Hout, ldh, hT, sT, Sckpt, B, T, H, dtype, train_mode, st);
    if (rc != SC_E_UNSUP) return rc;
  }
  if (dtype == SC_BF16) {
    SC_CHECK_ARG(H % 2 == 0 && ldg % 2 == 0 && ldh % 2 == 0, SC_E_ALIGN);
    SC_CHECK_ARG(((uintptr_t)G & 3) == 0 && ((uintptr_t)Hout & 3) == 0, SC_E_ALIGN);
    return launch_scan_fwd<bf16, 2, false>(G, ldg, h0, s0, Hout, ldh, hT, sT, Sckpt, B, T, H, train_mode, st);
  } else if (dtype == SC_F32) {
    return launch_scan_fwd<float, 1, true>(G, ldg, h0, s0, Hout, ldh, hT, sT, Sckpt, B, T, H, train_mode, st);
  }
  return SC_E_DTYPE;
}

extern "C" int sc_lucy_scan_bwd(const void* G, int64_t ldg, const void* Hout, int64_t ldh,
                                const float* h0, const float* s0, const float* Sckpt,
                                const void* dHout, int64_t lddh, void* dG, int64_t lddg, float* dbias,
                                int64_t B, int64_t T, int64_t H, int dtype, int train_mode,
                                void* stream) {
  SC_CHECK_ARG(h0 && Sckpt && (train_mode || s0), SC_E_BADARG);
  SC_CHECK_ARG(scan_args_ok(B, T, H), SC_E_SHAPE);
  if (T == 0) return 0;
  SC_CHECK_ARG(G && Hout && dHout && dG, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (!scan_force_generic()) {
    const int rc = scan_bwd_tma_dispatch(G, ldg, Hout, ldh, h0, Sckpt, dHout, lddh, dG, lddg, dbias, B, T, H, dtype, train_mode, st);
    if (rc != SC_E_UNSUP) return rc;
  }
  if (dtype == SC_BF16) {
    SC_CHECK_ARG(H % 2 == 0 && ldg % 2 == 0 && ldh % 2 == 0 && lddh % 2 == 0 && lddg % 2 == 0, SC_E_ALIGN);
    SC_CHECK_ARG((((uintptr_t)G | (uintptr_t)Hout | (uintptr_t)dHout | (uintptr_t)dG) & 3) == 0, SC_E_ALIGN);
    return launch_scan_bwd<bf16, 2, false>(G, ldg, Hout, ldh, h0, s0, Sckpt, dHout, lddh, dG, lddg, dbias, B, T, H, train_mode, st);
  } else if (dtype == SC_F32) {
    return launch_scan_bwd<float, 1, true>(G, ldg, Hout, ldh, h0, s0, Sckpt, dHout, lddh, dG, lddg, dbias, B, T, H, train_mode, st);
  }
  return SC_E_DTYPE;
}

extern "C" int sc_lucy_sscan_fwd(const void* k, const void* v, const void* q, int64_t ldg,
                                 const void* addend, int64_t ldadd, const float* s0,
                                 void* A, int64_t lda, float* S_all, float* sT,
                                 int64_t B, int64_t T, int64_t H, int dtype, int train_mode,
                                 int decay_mode, float lambda_decay, void* stream) {
  SC_CHECK_ARG(scan_args_ok(B, T, H), SC_E_SHAPE);
  SC_CHECK_ARG(decay_mode == 0 || (decay_mode == 1 && train_mode), SC_E_BADARG);
  SC_CHECK_ARG(train_mode || s0, SC_E_BADARG);
  if (T == 0) {
    if (!train_mode && sT && s0) return (int)cudaMemcpyAsync(sT, s0, sizeof(float) * B * H, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    return 0;
  }
  SC_CHECK_ARG(k && v && q && addend && A && S_all, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (!scan_force_generic() && decay_mode == 0) {
    const int rc = sscan_fwd_tma_dispatch(k, v, q, ldg, addend, ldadd, s0, A, lda, S_all, sT, B, T, H, dtype, train_mode, st);
    if (rc != SC_E_UNSUP) return rc;
  }
  const unsigned blocks = (unsigned)cdiv(B * H, 128);
  if (dtype == SC_BF16)
    sscan_fwd_kernel<bf16, false><<<blocks, 128, 0, st>>>((const bf16*)k, (const bf16*)v, (const bf16*)q, ldg,
        (const bf16*)addend, ldadd, s0, (bf16*)A, lda, S_all, sT, (int)B, (int)T, (int)H, train_mode, decay_mode, lambda_decay);
  else if (dtype == SC_F32)
    sscan_fwd_kernel<float, true><<<blocks, 128, 0, st>>>((const float*)k, (const float*)v, (const float*)q, ldg,
        (const float*)addend, ldadd, s0, (float*)A, lda, S_all, sT, (int)B, (int)T, (int)H, train_mode, decay_mode, lambda_decay);
  else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}

extern "C" int sc_lucy_sscan_bwd(const void* k, const void* v, const void* q, int64_t ldg,
                                 const float* S_all, const float* s0, const void* dA, int64_t ldda,
                                 void* dk, void* dv, void* dq, int64_t lddg, float* dsum,
                                 int64_t B, int64_t T, int64_t H, int dtype, int train_mode,
                                 int decay_mode, float lambda_decay, void* stream) {
  SC_CHECK_ARG(scan_args_ok(B, T, H), SC_E_SHAPE);
  SC_CHECK_ARG(decay_mode == 0 || (decay_mode == 1 && train_mode), SC_E_BADARG);
  SC_CHECK_ARG(train_mode || s0, SC_E_BADARG);
  if (T == 0) return 0;
  SC_CHECK_ARG(k && v && q && S_all && dA && dk && dv && dq, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (!scan_force_generic() && decay_mode == 0) {
    const int rc = sscan_bwd_tma_dispatch(k, v, q, ldg, S_all, s0, dA, ldda, dk, dv, dq, lddg, dsum, B, T, H, dtype, train_mode, st);
    if (rc != SC_E_UNSUP) return rc;
  }
  const unsigned blocks = (unsigned)cdiv(B * H, 128);
  if (dtype == SC_BF16) {
    if (decay_mode == 1)
      sscan_bwd_prefix_kernel<bf16, false><<<blocks, 128, 0, st>>>((const bf16*)k, (const bf16*)v, (const bf16*)q, ldg, S_all,
          (const bf16*)dA, ldda, (bf16*)dk, (bf16*)dv, (bf16*)dq, lddg, dsum, (int)B, (int)T, (int)H, lambda_decay);
    else
      sscan_bwd_kernel<bf16, false><<<blocks, 128, 0, st>>>((const bf16*)k, (const bf16*)v, (const bf16*)q, ldg, S_all, s0,
          (const bf16*)dA, ldda, (bf16*)dk, (bf16*)dv, (bf16*)dq, lddg, dsum, (int)B, (int)T, (int)H, train_mode, 0, lambda_decay);
  } else if (dtype == SC_F32) {
    if (decay_mode == 1)
      sscan_bwd_prefix_kernel<float, true><<<blocks, 128, 0, st>>>((const float*)k, (const float*)v, (const float*)q, ldg, S_all,
          (const float*)dA, ldda, (float*)dk, (float*)dv, (float*)dq, lddg, dsum, (int)B, (int)T, (int)H, lambda_decay);
    else
      sscan_bwd_kernel<float, true><<<blocks, 128, 0, st>>>((const float*)k, (const float*)v, (const float*)q, ldg, S_all, s0,
          (const float*)dA, ldda, (float*)dk, (float*)dv, (float*)dq, lddg, dsum, (int)B, (int)T, (int)H, train_mode, 0, lambda_decay);
  } else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}

extern "C" int sc_lucy_hscan_fwd(const void* An, int64_t ldan, const void* Zn, int64_t ldzn,
                                 const float* h0, void* Hout, int64_t ldh, float* hT,
                                 int64_t B, int64_t T, int64_t H, int dtype, void* stream) {
  SC_CHECK_ARG(scan_args_ok(B, T, H), SC_E_SHAPE);
  SC_CHECK_ARG(h0 && hT, SC_E_BADARG);
  SC_CHECK_ARG(T == 0 || (An && Zn && Hout), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (!scan_force_generic() && T > 0) {
    const int rc = hscan_fwd_tma_dispatch(An, ldan, Zn, ldzn, h0, Hout, ldh, hT, B, T, H, dtype, st);
    if (rc != SC_E_UNSUP) return rc;
  }
  const unsigned blocks = (unsigned)cdiv(B * H, 128);
  if (dtype == SC_BF16)
    hscan_fwd_kernel<bf16, false><<<blocks, 128, 0, st>>>((const bf16*)An, ldan, (const bf16*)Zn, ldzn, h0, (bf16*)Hout, ldh, hT, (int)B, (int)T, (int)H);
  else if (dtype == SC_F32)
    hscan_fwd_kernel<float, true><<<blocks, 128, 0, st>>>((const float*)An, ldan, (const float*)Zn, ldzn, h0, (float*)Hout, ldh, hT, (int)B, (int)T, (int)H);
  else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}

extern "C" int sc_lucy_hscan_bwd(const void* An, int64_t ldan, const void* Zn, int64_t ldzn,
                                 const void* Hout, int64_t ldh, const float* h0,
                                 const void* dHout, int64_t lddh, void* dAn, int64_t lddan,
                                 void* dZn, int64_t lddzn,
                                 int64_t B, int64_t T, int64_t H, int dtype, void* stream) {
  SC_CHECK_ARG(scan_args_ok(B, T, H), SC_E_SHAPE);
  if (T == 0) return 0;
  SC_CHECK_ARG(An && Zn && Hout && h0 && dHout && dAn && dZn, SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  if (!scan_force_generic()) {
    const int rc = hscan_bwd_tma_dispatch(An, ldan, Zn, ldzn, Hout, ldh, h0, dHout, lddh, dAn, lddan, dZn, lddzn, B, T, H, dtype, st);
    if (rc != SC_E_UNSUP) return rc;
  }
  const unsigned blocks = (unsigned)cdiv(B * H, 128);
  if (dtype == SC_BF16)
    hscan_bwd_kernel<bf16, false><<<blocks, 128, 0, st>>>((const bf16*)An, ldan, (const bf16*)Zn, ldzn, (const bf16*)Hout, ldh, h0,
        (const bf16*)dHout, lddh, (bf16*)dAn, lddan, (bf16*)dZn, lddzn, (int)B, (int)T, (int)H);
  else if (dtype == SC_F32)
    hscan_bwd_kernel<float, true><<<blocks, 128, 0, st>>>((const float*)An, ldan, (const float*)Zn, ldzn, (const float*)Hout, ldh, h0,
        (const float*)dHout, lddh, (float*)dAn, lddan, (float*)dZn, lddzn, (int)B, (int)T, (int)H);
  else return SC_E_DTYPE;
  SC_LAUNCH_RET();
}
