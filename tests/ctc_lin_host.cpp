// Host harness for statecatcher_b200/csrc/sc_ctc_lin_math.h (test infrastructure): runs the alpha recursion of
// ctc_alpha_beta_lin_kernel's per-node arithmetic — the SAME header the kernel compiles — over one utterance and
// writes log2(alpha) as doubles.  Input (binary, native endianness): int32 T, int32 S, S bytes of skip flags,
// T*S floats of lattice emissions in log2 units (<= 0 after the per-frame shift; anything below -1e6 is masked).
// Built and driven by tests/test_ctc_linear_algorithm_cpu.py.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "../statecatcher_b200/csrc/sc_ctc_lin_math.h"

static void split_emission(float e, float& pf, int& ei) {      // mirrors the kernel's lambda (ex2.approx -> exp2f)
  if (!(e > -1.0e6f)) { pf = 0.f; ei = 0; return; }
  const float fl = floorf(e);
  pf = exp2f(e - fl);
  ei = (int)fl;
}

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 3;
  int T = 0, S = 0;
  if (fread(&T, 4, 1, f) != 1 || fread(&S, 4, 1, f) != 1) return 4;
  std::vector<unsigned char> skip(S);
  std::vector<float> em((size_t)T * S);
  if (fread(skip.data(), 1, S, f) != (size_t)S || fread(em.data(), 4, em.size(), f) != em.size()) return 5;
  fclose(f);
  std::vector<float> pm(S + 2, 0.f), cm(S + 2, 0.f);           // two dead pad cells in front (s-1, s-2)
  std::vector<int> pe(S + 2, CTC_E_DEAD), ce(S + 2, CTC_E_DEAD);
  std::vector<double> out((size_t)T * S);
  for (int s = 0; s < S; ++s) {
    float m0 = 0.f; int e0 = CTC_E_DEAD;
    if (s < 2) { float pf; int ei; split_emission(em[s], pf, ei); ctc_lin_renorm(pf, ei, m0, e0); }
    pm[s + 2] = m0; pe[s + 2] = e0;
    out[s] = m0 > 0.f ? log2((double)m0) + e0 : -INFINITY;
  }
  for (int t = 1; t < T; ++t) {
    for (int s = 0; s < S; ++s) {
      float pf; int ei;
      split_emission(em[(size_t)t * S + s], pf, ei);
      float mc = 0.f; int ec = CTC_E_DEAD;
      if (skip[s]) { mc = pm[s]; ec = pe[s]; }
      float mn, sum; int en, emax;
      if (s & 1) ctc_lin_step(pm[s + 2], pe[s + 2], pm[s + 1], pe[s + 1], mc, ec, pf, ei, mn, en, sum, emax);
      else       ctc_lin_step2(pm[s + 2], pe[s + 2], pm[s + 1], pe[s + 1], pf, ei, mn, en, sum, emax);   // blank: two predecessors
      cm[s + 2] = mn; ce[s + 2] = en;
      out[(size_t)t * S + s] = mn > 0.f ? log2((double)mn) + en : -INFINITY;
    }
    pm.swap(cm); pe.swap(ce);
  }
  f = fopen(argv[2], "wb");
  if (!f) return 6;
  fwrite(out.data(), 8, out.size(), f);
  fclose(f);
  return 0;
}
