// Generic strided GEMM on the FFMA pipe with fp32 accumulation and the fused epilogues of the MAGIC
// linear layers.  This is the fp32 "parity mode" GEMM (fp32 mode, 1e-4 relative) and the
// fallback shape handler for operands the tcgen05 kernel (gemm_tc.cu) does not accept.
//
//   C[m,n] = epi( alpha * sum_k A(m,k) * B(k,n) + bias[n] ) (+ beta * C[m,n])
//   A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn]   (any of the strides may be 1)
//
// epilogue: optional bias, optional pre-activation copy-out, act in {none, gelu, relu}, optional
// multiply by act'(pre) (backward fusion), optional dropout after the activation.
#include "common.cuh"
#include "gemm_epi.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

template <typename TA, typename TB, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, const TB* __restrict__ B,
                                                        TC* __restrict__ C, int M, int N, int K, long sam,
                                                        long sak, long sbk, long sbn, long ldc, GemmEpi epi) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const bool a_kmajor = (sak == 1), b_kmajor = (sbk == 1);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int idx = tid + i * 256;
      int m, k;
      if (a_kmajor) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? ldf(A, (size_t)gm * sam + (size_t)gk * sak) : 0.f;
      int n, kb;
      if (b_kmajor) { kb = idx % BK; n = idx / BK; } else { n = idx % BN; kb = idx / BN; }
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < N && gkb < K) ? ldf(B, (size_t)gkb * sbk + (size_t)gn * sbn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const Dropout dr = make_dropout(epi.drop_p, epi.seed_ptr, epi.salt);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      epi_store<TC>(epi, dr, C, acc[i][j], gm, gn, ldc, epi.bias ? epi.bias[gn] : 0.f);
    }
  }
}

template <typename TA, typename TB, typename TC>
int launch(const void* A, const void* B, void* C, int M, int N, int K, long sam, long sak, long sbk, long sbn,
           long ldc, const GemmEpi& epi, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  gemm_simt_kernel<TA, TB, TC><<<grid, 256, 0, st>>>((const TA*)A, (const TB*)B, (TC*)C, M, N, K, sam, sak, sbk,
                                                      sbn, ldc, epi);
  MAGIC_CHECK_LAUNCH("magic_gemm(simt)");
  return MAGIC_OK;
}

}  // namespace

int gemm_simt_dispatch(const void* A, int a_dt, const void* B, int b_dt, void* C, int c_dt, int M, int N, int K,
                       long sam, long sak, long sbk, long sbn, long ldc, const GemmEpi& epi, cudaStream_t st) {
  typedef __nv_bfloat16 bf;
  const int key = a_dt * 4 + b_dt * 2 + c_dt;
  switch (key) {
    case 0: return launch<float, float, float>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
    case 1: return launch<float, float, bf>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
    case 2: return launch<float, bf, float>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
    case 3: return launch<float, bf, bf>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
    case 4: return launch<bf, float, float>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
    case 5: return launch<bf, float, bf>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
    case 6: return launch<bf, bf, float>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
    case 7: return launch<bf, bf, bf>(A, B, C, M, N, K, sam, sak, sbk, sbn, ldc, epi, st);
  }
  magic_set_error("magic_gemm: bad dtype combination");
  return MAGIC_ERR_ARG;
}
