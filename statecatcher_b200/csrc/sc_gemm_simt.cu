// K1 (fp32 parity path): SIMT fp32-FMA GEMM family for the projections.
//
// Replaces nn.Linear forward/backward of input_proj / W_fused / output_proj
// (lucyrnn.py:15, 23, 85; 113, 116, 186).  This kernel exists for the fp32 numerical
// contract (rtol 1e-4 against the fp32 reference): tensor cores have no fp32-exact mode, so
// fp32 activations take true fp32 FMAs here.  The bf16 training path (the benchmarked one)
// uses the tcgen05/TMA kernel in sc_gemm_tcgen05.cu; this kernel also accepts bf16 operands
// (fp32 accumulate) so that shapes the tcgen05 kernel does not tile still have a native path.
//
// One generic kernel: C[i,j] (+)= sum_r A(i,r) * B(j,r) (+ bias[j]) with arbitrary element
// strides for (i,r) on both operands, which covers
//   fwd   (i=m, j=n, r=k): A=X[m,k],  B=W[n,k]
//   dgrad (i=m, j=k, r=n): A=dY[m,n], B=W[n,k]   (B indexed [r,j])
//   wgrad (i=n, j=k, r=m): A=dY[m,n], B=X[m,k]   (both indexed [r,*]) with split-R + atomics.
#include "sc_common.cuh"

namespace sc {

constexpr int BM = 128, BN = 128, BK = 16, GT = 256, PAD = 4;

template <typename TA, typename TB, typename TC, bool A_RCONTIG, bool B_RCONTIG, bool ATOMIC>
__global__ void __launch_bounds__(GT)
gemm_simt_kernel(const TA* __restrict__ A, int64_t a_si, int64_t a_sr,
                 const TB* __restrict__ Bm, int64_t b_sj, int64_t b_sr,
                 const float* __restrict__ bias, TC* __restrict__ C, int64_t ldc,
                 int64_t I, int64_t J, int64_t R, int64_t r_chunk, int accumulate) {
  __shared__ float As[BK][BM + PAD];
  __shared__ float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t i0 = (int64_t)blockIdx.y * BM, j0 = (int64_t)blockIdx.x * BN;
  const int64_t r_begin = (int64_t)blockIdx.z * r_chunk;
  const int64_t r_end = (r_begin + r_chunk < R) ? r_begin + r_chunk : R;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
#pragma unroll
    for (int e0 = 0; e0 < BM * BK; e0 += GT) {
      const int e = e0 + tid;
      const int ii = A_RCONTIG ? e / BK : e % BM;
      const int rr = A_RCONTIG ? e % BK : e / BM;
      const int64_t gi = i0 + ii, gr = r0 + rr;
      float v = 0.f;
      if (gi < I && gr < r_end) v = ld_f(A + gi * a_si + gr * a_sr);
      As[rr][ii] = v;
    }
#pragma unroll
    for (int e0 = 0; e0 < BN * BK; e0 += GT) {
      const int e = e0 + tid;
      const int jj = B_RCONTIG ? e / BK : e % BN;
      const int rr = B_RCONTIG ? e % BK : e / BN;
      const int64_t gj = j0 + jj, gr = r0 + rr;
      float v = 0.f;
      if (gj < J && gr < r_end) v = ld_f(Bm + gj * b_sj + gr * b_sr);
      Bs[rr][jj] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      float a[8], b[8];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[r][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[r][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[r][tx * 8]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[r][tx * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gi = i0 + ty * 8 + i;
    if (gi >= I) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t gj = j0 + tx * 8 + j;
      if (gj >= J) continue;
      float v = acc[i][j];
      if (bias != nullptr && blockIdx.z == 0) v += bias[gj];
      TC* c = C + gi * ldc + gj;
      if (ATOMIC) {
        atomicAdd(reinterpret_cast<float*>(c), v);
      } else {
        if (accumulate) v += ld_f(c);
        st_f(c, v);
      }
    }
  }
}

__global__ void zero_rows_kernel(float* __restrict__ p, int64_t ld, int64_t rows, int64_t cols) {
  const int64_t n = rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[(i / cols) * ld + (i % cols)] = 0.f;
}

template <typename TA, typename TB, typename TC, bool AR, bool BR>
static int launch_simt(const void* A, int64_t a_si, int64_t a_sr, const void* B, int64_t b_sj, int64_t b_sr,
                       const float* bias, void* C, int64_t ldc, int64_t I, int64_t J, int64_t R,
                       int accumulate, bool allow_split, cudaStream_t st) {
  if (I == 0 || J == 0) return 0;
  dim3 grid((unsigned)cdiv(J, BN), (unsigned)cdiv(I, BM), 1);
  int64_t split = 1;
  if (allow_split && sizeof(TC) == 4) {
    const int64_t tiles = (int64_t)grid.x * grid.y;
    split = cdiv(2 * 148, tiles);
    const int64_t maxsplit = R / 512 > 0 ? R / 512 : 1;
    if (split > maxsplit) split = maxsplit;
    if (split > 64) split = 64;
  }
  if (split > 1) {
    int64_t chunk = cdiv(cdiv(R, split), BK) * BK;
    grid.z = (unsigned)cdiv(R, chunk);
    if (!accumulate) {
      zero_rows_kernel<<<(unsigned)min((int64_t)1024, cdiv(I * J, 256)), 256, 0, st>>>((float*)C, ldc, I, J);
    }
    gemm_simt_kernel<TA, TB, TC, AR, BR, true><<<grid, GT, 0, st>>>((const TA*)A, a_si, a_sr, (const TB*)B, b_sj, b_sr,
        bias, (TC*)C, ldc, I, J, R, chunk, 1);
  } else {
    gemm_simt_kernel<TA, TB, TC, AR, BR, false><<<grid, GT, 0, st>>>((const TA*)A, a_si, a_sr, (const TB*)B, b_sj, b_sr,
        bias, (TC*)C, ldc, I, J, R, R > 0 ? R : 1, accumulate);
  }
  SC_LAUNCH_RET();
}

int simt_gemm_fwd(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* Y,
                  int64_t ldy, int64_t M, int64_t N, int64_t K, int in_dtype, int out_dtype, cudaStream_t st) {
  if (in_dtype == SC_F32 && out_dtype == SC_F32)
    return launch_simt<float, float, float, true, true>(A, lda, 1, W, ldw, 1, bias, Y, ldy, M, N, K, 0, false, st);
  if (in_dtype == SC_BF16 && out_dtype == SC_BF16)
    return launch_simt<bf16, bf16, bf16, true, true>(A, lda, 1, W, ldw, 1, bias, Y, ldy, M, N, K, 0, false, st);
  if (in_dtype == SC_BF16 && out_dtype == SC_F32)
    return launch_simt<bf16, bf16, float, true, true>(A, lda, 1, W, ldw, 1, bias, Y, ldy, M, N, K, 0, false, st);
  return SC_E_DTYPE;
}
int simt_gemm_dgrad(const void* dY, int64_t lddy, const void* W, int64_t ldw, void* dA, int64_t ldda,
                    int64_t M, int64_t N, int64_t K, int in_dtype, int out_dtype, cudaStream_t st) {
  // out[m,k] = sum_n dY[m,n] W[n,k]
  if (in_dtype == SC_F32 && out_dtype == SC_F32)
    return launch_simt<float, float, float, true, false>(dY, lddy, 1, W, 1, ldw, nullptr, dA, ldda, M, K, N, 0, false, st);
  if (in_dtype == SC_BF16 && out_dtype == SC_BF16)
    return launch_simt<bf16, bf16, bf16, true, false>(dY, lddy, 1, W, 1, ldw, nullptr, dA, ldda, M, K, N, 0, false, st);
  if (in_dtype == SC_BF16 && out_dtype == SC_F32)
    return launch_simt<bf16, bf16, float, true, false>(dY, lddy, 1, W, 1, ldw, nullptr, dA, ldda, M, K, N, 0, false, st);
  return SC_E_DTYPE;
}
int simt_gemm_wgrad(const void* dY, int64_t lddy, const void* A, int64_t lda, float* dW, int64_t lddw,
                    int64_t M, int64_t N, int64_t K, int in_dtype, int accumulate, cudaStream_t st) {
  // out[n,k] = sum_m dY[m,n] A[m,k]
  if (in_dtype == SC_F32)
    return launch_simt<float, float, float, false, false>(dY, 1, lddy, A, 1, lda, nullptr, dW, lddw, N, K, M, accumulate, true, st);
  if (in_dtype == SC_BF16)
    return launch_simt<bf16, bf16, float, false, false>(dY, 1, lddy, A, 1, lda, nullptr, dW, lddw, N, K, M, accumulate, true, st);
  return SC_E_DTYPE;
}

}  // namespace sc
