// fp32 CUDA-core GEMM with arbitrary operand strides: the "fp32 mode" contraction used for the 1e-4 parity
// gate (tcgen05 kind::tf32 would be ~1e-3). Same epilogue contract as the tensor-core kernel.
//   C[m,n] = sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk]
#include "common.cuh"
#include "kernels.h"
#include "epilogue.cuh"

namespace fervit {

namespace simt {

constexpr int TM = 64, TN = 64, TK = 16, PAD = 4;

__device__ __forceinline__ void load_tile(const float* __restrict__ P, long long s_mn, long long s_k, int mn0,
                                          int k0, int MN, int K, float (*S)[TM + PAD], int t) {
  if (s_k == 1) {
    // k contiguous: thread -> (row = t/4, 4 consecutive k)
    const int r = t >> 2, kk = (t & 3) * 4;
    const int gr = mn0 + r, gk = k0 + kk;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gr < MN) {
      const float* src = P + (long long)gr * s_mn + gk;
      if (gk + 3 < K && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const float4 q = *reinterpret_cast<const float4*>(src);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (gk + i < K) v[i] = src[i];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) S[kk + i][r] = v[i];
  } else {
    // mn contiguous (or generic): thread -> (k = t/16, 4 consecutive rows)
    const int kk = t >> 4, r = (t & 15) * 4;
    const int gk = k0 + kk, gr = mn0 + r;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gk < K) {
      const float* src = P + (long long)gk * s_k + (long long)gr * s_mn;
      if (s_mn == 1 && gr + 3 < MN && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
        const float4 q = *reinterpret_cast<const float4*>(src);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (gr + i < MN) v[i] = src[(long long)i * s_mn];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) S[kk][r + i] = v[i];
  }
}

__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, long long sam, long long sak, const float* __restrict__ B,
                 long long sbn, long long sbk, int M, int N, int K, int k_per_split, Epilogue epi) {
  __shared__ float As[TK][TM + PAD];
  __shared__ float Bs[TK][TN + PAD];
  const int t = threadIdx.x;
  const int ty = t >> 4, tx = t & 15;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const int kbeg = blockIdx.z * k_per_split;
  const int kend = min(K, kbeg + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += TK) {
    load_tile(A, sam, sak, m0, k0, M, kend, As, t);
    load_tile(B, sbn, sbk, n0, k0, N, kend, Bs, t);
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  if (gridDim.z > 1) epi.out_f32 += (size_t)blockIdx.z * (size_t)M * (size_t)N;
  float alpha = epi.alpha;
  if (epi.alpha_ptr) alpha *= __ldg(epi.alpha_ptr);
  const int col = n0 + tx * 4;
  if (col >= N) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row < M) epilogue_apply<float, 4, EPK_GENERIC>(epi, alpha, row, col, N, acc[i]);
  }
}

}  // namespace simt

int gemm_f32_simt(const float* A, long long sam, long long sak, const float* B, long long sbn, long long sbk, int M,
                  int N, int K, int splits, const Epilogue& epi, cudaStream_t stream) {
  FV_CHECK(M > 0 && N > 0 && K > 0, "gemm_f32_simt: empty problem M=%d N=%d K=%d", M, N, K);
  FV_CHECK(N % 4 == 0, "gemm_f32_simt: N must be a multiple of 4 (got %d)", N);
  FV_CHECK(epi.ldo % 4 == 0, "gemm_f32_simt: ldo must be a multiple of 4 (got %d)", epi.ldo);
  if (splits < 1) splits = 1;
  int k_per_split = ceil_div(ceil_div(K, splits), simt::TK) * simt::TK;
  splits = ceil_div(K, k_per_split);
  if (splits > 1)
    FV_CHECK(epi.out_f32 != nullptr && epi.out == nullptr && epi.bias == nullptr && epi.residual == nullptr &&
                 epi.act == 0 && epi.act_bwd == 0 && epi.out_pre == nullptr && epi.remap_L == 0 && epi.ldo == N,
             "gemm_f32_simt: split-K supports only a plain fp32 partial output");
  dim3 grid(ceil_div(N, simt::TN), ceil_div(M, simt::TM), splits);
  ProfScope prof(3, 2.0 * M * (double)N * K, stream);
  simt::gemm_simt_kernel<<<grid, 256, 0, stream>>>(A, sam, sak, B, sbn, sbk, M, N, K, k_per_split, epi);
  FV_COUNT_LAUNCH();
  FV_LAUNCH_CHECK();
  return 0;
}

int gemm_f32_simt_effective_splits(int K, int splits) {
  if (splits < 1) splits = 1;
  const int k_per_split = ceil_div(ceil_div(K, splits), simt::TK) * simt::TK;
  return ceil_div(K, k_per_split);
}

}  // namespace fervit
