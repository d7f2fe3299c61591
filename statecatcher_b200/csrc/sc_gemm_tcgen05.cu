// K1 (bf16 training path): tcgen05 + TMA GEMM — placeholder until the kernel lands.
#include "sc_common.cuh"
namespace sc {
bool tc_gemm_fwd_ok(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, const void*, const void*, const void*) { return false; }
int tc_gemm_fwd(const void*, int64_t, const void*, int64_t, const float*, void*, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t) { return SC_E_UNSUP; }
bool tc_gemm_dgrad_ok(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, int, const void*, const void*, const void*) { return false; }
int tc_gemm_dgrad(const void*, int64_t, const void*, int64_t, void*, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t) { return SC_E_UNSUP; }
bool tc_gemm_wgrad_ok(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int, const void*, const void*, const void*) { return false; }
int tc_gemm_wgrad(const void*, int64_t, const void*, int64_t, float*, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t) { return SC_E_UNSUP; }
}
