// Small fp32 matrix products for the classifier head (and EfficientNet's squeeze-excitation layers): logits = feat W^T + b and
// its two gradients (reference nbdt/models/resnet.py:142-148: avg_pool -> view -> self.linear).  These are a few MFLOP..GFLOP
// on fp32 features -- far below anything that matters for the step time -- but they were the last library (cuBLAS) launches
// inside the timed train step; now every launch of the step is one of this repository's kernels.  64 x 64 output tile per CTA,
// 4 x 4 per thread, 16-wide K slabs through shared memory, arbitrary strides (so the transposed products need no copies).
#include <cuda_runtime.h>

#include "common.h"

namespace {

constexpr int kTM = 64, kTN = 64, kTK = 16;

// C[m, n] = sum_k A(m, k) * B(k, n) (+ bias[n]) (+ C[m, n] if accumulate);  A(m, k) = A[m * am + k * ak], B(k, n) = B[k * bk + n * bn]
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, long long am, long long ak,
                                                    const float* __restrict__ B, long long bk, long long bn, float* __restrict__ C, long long ldc,
                                                    const float* __restrict__ bias, int accumulate) {
  __shared__ float As[kTK][kTM + 4];
  __shared__ float Bs[kTK][kTN + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kTK) {
    for (int e = threadIdx.x; e < kTM * kTK; e += 256) {
      // consecutive threads walk the unit-stride direction of each operand
      int mm, kk;
      if (ak == 1) { kk = e % kTK; mm = e / kTK; } else { mm = e % kTM; kk = e / kTM; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < K) ? A[m * am + k * ak] : 0.f;
    }
    for (int e = threadIdx.x; e < kTN * kTK; e += 256) {
      int nn, kk;
      if (bk == 1) { kk = e % kTK; nn = e / kTK; } else { nn = e % kTN; kk = e / kTN; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < K) ? B[k * bk + n * bn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j] + (bias ? bias[n] : 0.f);
      if (accumulate) v += C[m * ldc + n];
      C[m * ldc + n] = v;
    }
  }
}

// out[n] = sum_m X[m, n]
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ X, int M, int N, long long ld, float* __restrict__ out) {
  const int n = blockIdx.x * 32 + (threadIdx.x & 31);
  float s = 0.f;
  if (n < N)
    for (int m = threadIdx.x >> 5; m < M; m += 8) s += X[m * ld + n];
  __shared__ float red[8][33];
  red[threadIdx.x >> 5][threadIdx.x & 31] = s;
  __syncthreads();
  if (threadIdx.x < 32 && n < N) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    out[n] = t;
  }
}

}  // namespace

extern "C" {

int nbdt_sgemm(int M, int N, int K, const float* A, long long a_row_stride, long long a_col_stride, const float* B, long long b_row_stride,
               long long b_col_stride, float* C, long long ldc, const float* bias_n, int accumulate, void* stream) {
  NBDT_REQUIRE(M >= 0 && N >= 0 && K >= 0, "negative dimension");
  if (M == 0 || N == 0) return NBDT_OK;
  NBDT_REQUIRE(A && B && C, "NULL matrix");
  dim3 grid((N + kTN - 1) / kTN, (M + kTM - 1) / kTM);
  sgemm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(M, N, K, A, a_row_stride, a_col_stride, B, b_row_stride, b_col_stride, C, ldc, bias_n, accumulate);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

int nbdt_colsum(const float* X, int M, int N, long long ld, float* out, void* stream) {
  NBDT_REQUIRE(M >= 0 && N >= 0, "negative dimension");
  if (N == 0) return NBDT_OK;
  NBDT_REQUIRE(X && out, "NULL matrix");
  colsum_kernel<<<(N + 31) / 32, 256, 0, (cudaStream_t)stream>>>(X, M, N, ld, out);
  NBDT_CUDA_OK(cudaGetLastError());
  return NBDT_OK;
}

}  // extern "C"
