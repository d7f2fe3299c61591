// Arithmetic of the linear-domain CTC recursion with a per-node exponent (alpha = m * 2^e, m in [0.5, 1) fp32,
// e int) — the part of ctc_alpha_beta_lin_kernel (sc_ctc.cu, SC_CTC_WAVE=3) that is new.  Kept in a header that
// also compiles as plain C++ so that the bit manipulation is exercised on the CPU
// (tests/test_ctc_linear_algorithm_cpu.py builds tests/ctc_lin_host.cpp against it) before the kernel ever runs.
#pragma once
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SC_LIN_HD __host__ __device__ __forceinline__
#else
#define SC_LIN_HD static inline
#endif

constexpr int CTC_E_DEAD = -(1 << 28);          // exponent of a node with probability zero (mantissa 0)

SC_LIN_HD float ctc_lin_from_bits(int b) {
#ifdef __CUDA_ARCH__
  return __int_as_float(b);
#else
  float f; memcpy(&f, &b, sizeof f); return f;
#endif
}
SC_LIN_HD int ctc_lin_to_bits(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_int(f);
#else
  int b; memcpy(&b, &f, sizeof b); return b;
#endif
}

// m * 2^d for d <= 0: the power of two is built in the exponent field; a term more than 60 binades below the
// largest one cannot reach the last bit of the sum.
SC_LIN_HD float ctc_lin_scale_pow2(float m, int d) {
  return d < -60 ? 0.f : m * ctc_lin_from_bits((127 + d) << 23);
}

// v > 0 (normal) -> mantissa in [0.5, 1) and the exponent that goes with it on top of ebase; v == 0 -> dead node
SC_LIN_HD void ctc_lin_renorm(float v, int ebase, float& m, int& e) {
  if (v > 0.f) {
    const int bits = ctc_lin_to_bits(v);
    m = ctc_lin_from_bits((bits & 0x007fffff) | (126 << 23));
    e = ebase + ((bits >> 23) & 0xff) - 126;
  } else {
    m = 0.f;
    e = CTC_E_DEAD;
  }
}

// One node, one timestep: the three predecessors (stay, previous node, skip; dead ones carry m = 0, e = CTC_E_DEAD)
// and the node's emission as 2^ei * pf with pf in [1, 2) (pf = 0: masked).  Returns the new (mantissa, exponent)
// and, for the beta rows that are stored without their frame's emission, the aligned sum and its exponent.
SC_LIN_HD void ctc_lin_step(float ma, int ea, float mb, int eb, float mc, int ec, float pf, int ei,
                            float& mn, int& en, float& sum, int& emax) {
  emax = ea > eb ? ea : eb;
  emax = emax > ec ? emax : ec;
  sum = ctc_lin_scale_pow2(ma, ea - emax) + ctc_lin_scale_pow2(mb, eb - emax) + ctc_lin_scale_pow2(mc, ec - emax);
  ctc_lin_renorm(sum * pf, emax + ei, mn, en);
}

// Two predecessors only (a blank node: itself and the node before it).
SC_LIN_HD void ctc_lin_step2(float ma, int ea, float mb, int eb, float pf, int ei,
                             float& mn, int& en, float& sum, int& emax) {
  emax = ea > eb ? ea : eb;
  sum = ctc_lin_scale_pow2(ma, ea - emax) + ctc_lin_scale_pow2(mb, eb - emax);
  ctc_lin_renorm(sum * pf, emax + ei, mn, en);
}
