// K4: RNN-T wavefront loss — placeholder until the kernel lands.
#include "sc_common.cuh"
extern "C" int sc_rnnt_fwd_bwd(const float*, const int64_t*, int64_t, const int64_t*, const int64_t*,
                               int64_t, int64_t, int64_t, int64_t, int64_t, float*, float*, float*, float*,
                               float*, const float*, float*, void*) { return SC_E_UNSUP; }
