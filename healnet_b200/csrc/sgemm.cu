// sgemm.cu — strided, batched fp32 SIMT contraction used by the backward pass for everything that is not a large
// regular GEMM: per-head products with C = 13..63 wide contexts, transposed operands (weight gradients contract over
// the row axis), operands stored as split fp16 [hi | lo] pairs by the forward pass, accumulation into gradient
// buffers (tied layers). The large regular products of the backward pass go through the tcgen05 GEMM (gemm.cu,
// bf16 hi/lo operands); this kernel trades speed for generality: any strides, any sizes, exact fp32 arithmetic.
//
//   C[b1][b2][m][n] (+)= alpha * sum_k A[b1][b2](m, k) * B[b1][b2](k, n)
#include "common.cuh"

namespace hn {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float ld_elem(const SgOperand& o, long idx) {
  if (o.type == 0) return __ldg(static_cast<const float*>(o.p) + idx);
  const __half* h = static_cast<const __half*>(o.p) + idx;
  return __half2float(h[0]) + __half2float(h[o.lo_off]);
}

__global__ void __launch_bounds__(256) sgemm_kernel(SgArgs a) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int b1 = blockIdx.z / a.nb2, b2 = blockIdx.z % a.nb2;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  const long a_base = b1 * a.A.s_b1 + b2 * a.A.s_b2, b_base = b1 * a.B.s_b1 + b2 * a.B.s_b2;
  const bool a_kfast = a.A.s_col == 1;   // k contiguous in memory
  const bool b_nfast = a.B.s_col == 1;   // n contiguous in memory
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += TK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int k = a_kfast ? (idx & 15) : (idx >> 6), m = a_kfast ? (idx >> 4) : (idx & 63);
      float v = 0.f;
      if (m0 + m < a.M && k0 + k < a.K) v = ld_elem(a.A, a_base + static_cast<long>(m0 + m) * a.A.s_row + static_cast<long>(k0 + k) * a.A.s_col);
      As[k][m] = v;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int k = b_nfast ? (idx >> 6) : (idx & 15), n = b_nfast ? (idx & 63) : (idx >> 4);
      float v = 0.f;
      if (n0 + n < a.N && k0 + k < a.K) v = ld_elem(a.B, b_base + static_cast<long>(k0 + k) * a.B.s_row + static_cast<long>(n0 + n) * a.B.s_col);
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* c = a.C + b1 * a.c_b1 + b2 * a.c_b2;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= a.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= a.N) continue;
      float* dst = c + static_cast<long>(m) * a.c_row + static_cast<long>(n) * a.c_col;
      const float v = a.alpha * acc[i][j];
      *dst = a.accumulate ? *dst + v : v;
    }
  }
}

}  // namespace

int launch_sgemm(const SgArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.nb1 <= 0 || a.nb2 <= 0) return 0;
  HN_REQUIRE(a.K >= 0 && a.A.p != nullptr && a.B.p != nullptr && a.C != nullptr, "sgemm: bad argument");
  const long nz = static_cast<long>(a.nb1) * a.nb2;
  HN_REQUIRE(nz <= 65535, "sgemm: too many batches");
  const dim3 grid((a.N + TN - 1) / TN, (a.M + TM - 1) / TM, static_cast<unsigned>(nz));
  HN_REQUIRE(grid.y <= 65535, "sgemm: M too large for this kernel's grid");
  HN_CHECK_CUDA(launch_k(sgemm_kernel, grid, dim3(256), 0, stream, a));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hn
