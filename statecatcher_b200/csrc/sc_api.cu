// C-ABI glue: version / error strings and the projection (K1) dispatcher.
#include "sc_common.cuh"
#include <stdio.h>
#include <string.h>

namespace sc {
int simt_gemm_fwd(const void*, int64_t, const void*, int64_t, const float*, void*, int64_t, int64_t, int64_t, int64_t, int, int, cudaStream_t);
int simt_gemm_dgrad(const void*, int64_t, const void*, int64_t, void*, int64_t, int64_t, int64_t, int64_t, int, int, cudaStream_t);
int simt_gemm_wgrad(const void*, int64_t, const void*, int64_t, float*, int64_t, int64_t, int64_t, int64_t, int, int, cudaStream_t);
// tcgen05/TMA path (sc_gemm_tcgen05.cu).  *_ok say whether the shape/alignment is tiled by it.
bool tc_gemm_fwd_ok(int64_t lda, int64_t ldw, int64_t ldy, int64_t M, int64_t N, int64_t K, int in_dtype, int out_dtype,
                    const void* A, const void* W, const void* Y, bool forced);
int tc_gemm_fwd(const void*, int64_t, const void*, int64_t, const float*, void*, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);
bool tc_gemm_dgrad_ok(int64_t lddy, int64_t ldw, int64_t ldda, int64_t M, int64_t N, int64_t K, int in_dtype, int out_dtype,
                      const void* dY, const void* W, const void* dA, bool forced);
int tc_gemm_dgrad(const void*, int64_t, const void*, int64_t, void*, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);
bool tc_gemm_wgrad_ok(int64_t lddy, int64_t lda, int64_t lddw, int64_t M, int64_t N, int64_t K, int in_dtype,
                      const void* dY, const void* A, const void* dW);
int tc_gemm_wgrad(const void*, int64_t, const void*, int64_t, float*, int64_t, int64_t, int64_t, int64_t, int, cudaStream_t);
}  // namespace sc

using namespace sc;

extern "C" int sc_version(void) { return 8; }

extern "C" const char* sc_error_string(int code) {
  switch (code) {
    case 0: return "ok";
    case SC_E_BADARG: return "SC_E_BADARG: null pointer, negative size or unsupported flag";
    case SC_E_ALIGN: return "SC_E_ALIGN: pointer or stride violates kernel alignment";
    case SC_E_DTYPE: return "SC_E_DTYPE: unknown or unsupported dtype combination";
    case SC_E_SHAPE: return "SC_E_SHAPE: shape outside supported range";
    case SC_E_UNSUP: return "SC_E_UNSUP: request not supported by this build";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown statecatcher_b200 error";
}

extern "C" int sc_build_info(char* buf, int64_t n) {
  if (!buf || n <= 0) return SC_E_BADARG;
  snprintf(buf, (size_t)n, "statecatcher_b200 abi=%d arch=sm_100a cuda=%d.%d", sc_version(),
           CUDART_VERSION / 1000, (CUDART_VERSION % 1000) / 10);
  return 0;
}

extern "C" int64_t sc_gemm_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }

static bool gemm_sizes_ok(int64_t M, int64_t N, int64_t K) { return M >= 0 && N >= 0 && K >= 0; }

extern "C" int sc_gemm_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias,
                           void* Y, int64_t ldy, int64_t M, int64_t N, int64_t K,
                           int in_dtype, int out_dtype, int impl, void* stream) {
  SC_CHECK_ARG(gemm_sizes_ok(M, N, K) && impl >= 0 && impl <= 2, SC_E_BADARG);
  if (M == 0 || N == 0) return 0;
  SC_CHECK_ARG(Y && (K == 0 || (A && W)), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = impl != 1 && tc_gemm_fwd_ok(lda, ldw, ldy, M, N, K, in_dtype, out_dtype, A, W, Y, impl == 2);
  if (impl == 2 && !tc) return SC_E_UNSUP;
  if (tc) return tc_gemm_fwd(A, lda, W, ldw, bias, Y, ldy, M, N, K, out_dtype, st);
  return simt_gemm_fwd(A, lda, W, ldw, bias, Y, ldy, M, N, K, in_dtype, out_dtype, st);
}

extern "C" int sc_gemm_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw,
                             void* dA, int64_t ldda, int64_t M, int64_t N, int64_t K,
                             int in_dtype, int out_dtype, int impl, void* stream) {
  SC_CHECK_ARG(gemm_sizes_ok(M, N, K) && impl >= 0 && impl <= 2, SC_E_BADARG);
  if (M == 0 || K == 0) return 0;
  SC_CHECK_ARG(dA && (N == 0 || (dY && W)), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = impl != 1 && tc_gemm_dgrad_ok(lddy, ldw, ldda, M, N, K, in_dtype, out_dtype, dY, W, dA, impl == 2);
  if (impl == 2 && !tc) return SC_E_UNSUP;
  if (tc) return tc_gemm_dgrad(dY, lddy, W, ldw, dA, ldda, M, N, K, out_dtype, st);
  return simt_gemm_dgrad(dY, lddy, W, ldw, dA, ldda, M, N, K, in_dtype, out_dtype, st);
}

extern "C" int sc_gemm_wgrad(const void* dY, int64_t lddy, const void* A, int64_t lda,
                             float* dW, int64_t lddw, int64_t M, int64_t N, int64_t K,
                             int in_dtype, int accumulate, int impl, void* stream) {
  SC_CHECK_ARG(gemm_sizes_ok(M, N, K) && impl >= 0 && impl <= 2, SC_E_BADARG);
  if (N == 0 || K == 0) return 0;
  SC_CHECK_ARG(dW && (M == 0 || (dY && A)), SC_E_BADARG);
  cudaStream_t st = (cudaStream_t)stream;
  const bool tc = impl != 1 && tc_gemm_wgrad_ok(lddy, lda, lddw, M, N, K, in_dtype, dY, A, dW);
  if (impl == 2 && !tc) return SC_E_UNSUP;
  if (tc) return tc_gemm_wgrad(dY, lddy, A, lda, dW, lddw, M, N, K, accumulate, st);
  return simt_gemm_wgrad(dY, lddy, A, lda, dW, lddw, M, N, K, in_dtype, accumulate, st);
}
