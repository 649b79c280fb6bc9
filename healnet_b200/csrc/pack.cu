// pack.cu — one-time (per weight version) repacking of reference-format fp32 parameters into the fp16
// operand layouts the tensor-core kernels consume. Runs on the device; nothing here is on the per-forward
// path. Layout conventions:
//   * heads are padded to a pitch hp of 64 (dim_head <= 64) or 128 columns ("head-padded"): column h*hp+d holds
//     head h, dim d (zero for d >= dh)
//   * softmax scale 2/sqrt(dh) (healnet.py:375,409,419) times log2(e) is folded into the Q projection
//   * context LayerNorm affine (gamma, beta; healnet.py:310-311,318) is folded into the K/V projection
//   * FeedForward first Linear rows are interleaved (a_j, g_j) so the gate fuses into the GEMM epilogue
#include "pack.cuh"

namespace hn {
namespace {

// hi = fp16(v) at dst[k]; lo = fp16(v - hi) at dst[lo_off + k] when lo_off > 0
__device__ __forceinline__ void put_split(__half* row, int k, int lo_off, float v) {
  const __half hi = __float2half_rn(v);
  row[k] = hi;
  if (lo_off > 0) row[lo_off + k] = __float2half_rn(v - __half2float(hi));
}

// dst[(dst_row0 + h*hp + d)][k] = scale * src[(src_row0 + h*dh + d)][k] * (colscale ? colscale[k] : 1)
__global__ void pack_headpad_rows_kernel(__half* __restrict__ dst, int ld_dst, int dst_row0,
                                         const float* __restrict__ src, int ld_src, int src_row0, int n_heads,
                                         int dh, int K, float scale, const float* __restrict__ colscale, int seg,
                                         int lo_off, int hp) {
  const int r = blockIdx.x;  // 0 .. n_heads*hp
  const int h = r / hp, d = r % hp;
  __half* out = dst + static_cast<size_t>(dst_row0 + r) * ld_dst;
  const float* in = src + static_cast<size_t>(src_row0 + h * dh + d) * ld_src;
  for (int k = threadIdx.x; k < seg; k += blockDim.x) {
    float v = 0.f;
    if (d < dh && k < K) v = scale * in[k] * (colscale ? colscale[k] : 1.f);
    put_split(out, k, lo_off, v);
  }
}

// dst[r][h*hp + d] = src[r][h*dh + d]
__global__ void pack_headpad_cols_kernel(__half* __restrict__ dst, int ld_dst, const float* __restrict__ src,
                                         int ld_src, int n_heads, int dh, int seg, int lo_off, int hp) {
  const int r = blockIdx.x;
  for (int c = threadIdx.x; c < seg; c += blockDim.x) {
    const int h = c / hp, d = c % hp;
    float v = 0.f;
    if (h < n_heads && d < dh) v = src[static_cast<size_t>(r) * ld_src + h * dh + d];
    put_split(dst + static_cast<size_t>(r) * ld_dst, c, lo_off, v);
  }
}

// FeedForward Linear(D, 8D): rows [0,4D) are "a", rows [4D,8D) are gates "g" (chunk(2), healnet.py:330).
// dst row 2j = a_j, row 2j+1 = g_j; bias likewise.
__global__ void pack_ff1_kernel(__half* __restrict__ dst, int ld_dst, float* __restrict__ bias_dst,
                                const float* __restrict__ W, const float* __restrict__ bias, int D, int hidden,
                                int seg, int lo_off) {
  const int r = blockIdx.x;  // 0 .. 2*hidden
  const int j = r >> 1;
  const int srow = (r & 1) ? hidden + j : j;
  for (int k = threadIdx.x; k < seg; k += blockDim.x)
    put_split(dst + static_cast<size_t>(r) * ld_dst, k, lo_off, k < D ? W[static_cast<size_t>(srow) * D + k] : 0.f);
  if (threadIdx.x == 0) bias_dst[r] = bias[srow];
}

__global__ void pack_plain_kernel(__half* __restrict__ dst, int ld_dst, const float* __restrict__ src, int ld_src,
                                  int K, int seg, int lo_off) {
  const int r = blockIdx.x;
  for (int k = threadIdx.x; k < seg; k += blockDim.x)
    put_split(dst + static_cast<size_t>(r) * ld_dst, k, lo_off, k < K ? src[static_cast<size_t>(r) * ld_src + k] : 0.f);
}

// bias_dst[dst_row0 + h*hp + d] = sum_c W[(src_row0 + h*dh + d)][c] * beta[c]
__global__ void fold_beta_headpad_kernel(float* __restrict__ bias_dst, int dst_row0, const float* __restrict__ W,
                                         int ld, int src_row0, int n_heads, int dh, int C,
                                         const float* __restrict__ beta, int hp) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_heads * hp) return;
  const int h = r / hp, d = r % hp;
  float acc = 0.f;
  if (d < dh) {
    const float* w = W + static_cast<size_t>(src_row0 + h * dh + d) * ld;
    for (int c = lane; c < C; c += 32) acc += w[c] * beta[c];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) bias_dst[dst_row0 + r] = acc;
}

// small-C cross-attention: Aq[(h*zw + c)][k] = scale * gamma[c] * sum_d Wk[h*dh+d][c] * Wq[h*dh+d][k]
__global__ void pack_smallc_q_kernel(__half* __restrict__ Aq, int ld_dst, const float* __restrict__ Wq,
                                     const float* __restrict__ Wkv, const float* __restrict__ gamma, int D, int C,
                                     int dh, float scale, int zw, int seg, int lo_off) {
  const int r = blockIdx.x;  // h*zw + c
  const int h = r / zw, c = r % zw;
  for (int k = threadIdx.x; k < seg; k += blockDim.x) {
    float acc = 0.f;
    if (c < C && k < D) {
      for (int d = 0; d < dh; ++d)
        acc += Wkv[static_cast<size_t>(h * dh + d) * C + c] * Wq[static_cast<size_t>(h * dh + d) * D + k];
      acc *= scale * gamma[c];
    }
    put_split(Aq + static_cast<size_t>(r) * ld_dst, k, lo_off, acc);
  }
}

// small-C V side: Wv'[i][c] = Wv[i][c] * gamma[c] (zw-wide rows, zero pad), bv[i] = sum_c Wv[i][c] * beta[c]
__global__ void pack_smallc_v_kernel(float* __restrict__ Wv_dst, float* __restrict__ bv_dst,
                                     const float* __restrict__ Wkv, const float* __restrict__ gamma,
                                     const float* __restrict__ beta, int inner, int C, int zw) {
  const int i = blockIdx.x;
  const int lane = threadIdx.x;  // 32 threads
  float acc = 0.f;
  for (int c = lane; c < zw; c += 32) {
    const float w = c < C ? Wkv[static_cast<size_t>(inner + i) * C + c] : 0.f;
    Wv_dst[static_cast<size_t>(i) * zw + c] = c < C ? w * gamma[c] : 0.f;
    acc += c < C ? w * beta[c] : 0.f;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) bv_dst[i] = acc;
}

// small-C output side: the V projection and the output projection are both linear in the merged accumulator u, so
// they fold into ONE weight:  WoS[n][h*zw + c] = sum_d Wo[n][h*dh + d] * Wv'[h*dh + d][c]  (split hi | lo rows),
// boS[n] = bo[n] + sum_i Wo[n][i] * bv'[i]. The out-projection GEMM then contracts over H*zw (256) instead of
// H*64 (512) columns and the combine kernel only has to merge and normalise.
__global__ void pack_smallc_out_kernel(__half* __restrict__ WoS, int ld_dst, float* __restrict__ boS,
                                       const float* __restrict__ Wo, const float* __restrict__ bo,
                                       const float* __restrict__ Wv, const float* __restrict__ bv, int inner, int H,
                                       int dh, int zw, int seg, int lo_off) {
  const int n = blockIdx.x;
  const float* wo = Wo + static_cast<size_t>(n) * inner;
  __half* dst = WoS + static_cast<size_t>(n) * ld_dst;
  for (int j = threadIdx.x; j < seg; j += blockDim.x) {
    float acc = 0.f;
    if (j < H * zw) {
      const int h = j / zw, c = j % zw;
      for (int d = 0; d < dh; ++d) acc += wo[h * dh + d] * Wv[static_cast<size_t>(h * dh + d) * zw + c];
    }
    const __half hi = __float2half_rn(acc);
    dst[j] = hi;
    dst[lo_off + j] = __float2half_rn(acc - __half2float(hi));
  }
  __shared__ float red[32];
  float b = 0.f;
  for (int i = threadIdx.x; i < inner; i += blockDim.x) b += wo[i] * bv[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = b;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = bo[n];
    for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += red[w];
    boS[n] = t;
  }
}
}  // namespace

#define PACK_LAUNCH_CHECK()                    \
  do {                                         \
    HN_CHECK_CUDA(cudaGetLastError());         \
  } while (0)

int pack_headpad_rows(__half* dst, int ld_dst, int dst_row0, const float* src, int ld_src, int src_row0,
                      int n_heads, int dh, int K, float scale, const float* colscale, int seg, int lo_off, int hp,
                      cudaStream_t st) {
  pack_headpad_rows_kernel<<<n_heads * hp, 128, 0, st>>>(dst, ld_dst, dst_row0, src, ld_src, src_row0, n_heads, dh, K,
                                                        scale, colscale, seg, lo_off, hp);
  PACK_LAUNCH_CHECK();
  return 0;
}
int pack_headpad_cols(__half* dst, int ld_dst, const float* src, int ld_src, int rows, int n_heads, int dh, int seg,
                      int lo_off, int hp, cudaStream_t st) {
  pack_headpad_cols_kernel<<<rows, 128, 0, st>>>(dst, ld_dst, src, ld_src, n_heads, dh, seg, lo_off, hp);
  PACK_LAUNCH_CHECK();
  return 0;
}
int pack_ff1(__half* dst, int ld_dst, float* bias_dst, const float* W, const float* bias, int D, int hidden, int seg,
             int lo_off, cudaStream_t st) {
  pack_ff1_kernel<<<2 * hidden, 128, 0, st>>>(dst, ld_dst, bias_dst, W, bias, D, hidden, seg, lo_off);
  PACK_LAUNCH_CHECK();
  return 0;
}
int pack_plain(__half* dst, int ld_dst, const float* src, int ld_src, int rows, int K, int seg, int lo_off,
               cudaStream_t st) {
  pack_plain_kernel<<<rows, 128, 0, st>>>(dst, ld_dst, src, ld_src, K, seg, lo_off);
  PACK_LAUNCH_CHECK();
  return 0;
}
int fold_beta_headpad(float* bias_dst, int dst_row0, const float* W, int ld, int src_row0, int n_heads, int dh,
                      int C, const float* beta, int hp, cudaStream_t st) {
  const int rows = n_heads * hp;
  fold_beta_headpad_kernel<<<(rows + 3) / 4, 128, 0, st>>>(bias_dst, dst_row0, W, ld, src_row0, n_heads, dh, C, beta,
                                                          hp);
  PACK_LAUNCH_CHECK();
  return 0;
}
int pack_smallc_q(__half* Aq, int ld_dst, const float* Wq, const float* Wkv, const float* gamma, int H, int D, int C,
                  int dh, float scale, int zw, int seg, int lo_off, cudaStream_t st) {
  pack_smallc_q_kernel<<<H * zw, 128, 0, st>>>(Aq, ld_dst, Wq, Wkv, gamma, D, C, dh, scale, zw, seg, lo_off);
  PACK_LAUNCH_CHECK();
  return 0;
}
int pack_smallc_v(float* Wv_dst, float* bv_dst, const float* Wkv, const float* gamma, const float* beta, int inner,
                  int C, int zw, cudaStream_t st) {
  pack_smallc_v_kernel<<<inner, 32, 0, st>>>(Wv_dst, bv_dst, Wkv, gamma, beta, inner, C, zw);
  PACK_LAUNCH_CHECK();
  return 0;
}

int pack_smallc_out(__half* WoS, int ld_dst, float* boS, const float* Wo, const float* bo, const float* Wv,
                    const float* bv, int D, int inner, int H, int dh, int zw, int seg, int lo_off, cudaStream_t st) {
  pack_smallc_out_kernel<<<D, 256, 0, st>>>(WoS, ld_dst, boS, Wo, bo, Wv, bv, inner, H, dh, zw, seg, lo_off);
  PACK_LAUNCH_CHECK();
  return 0;
}

}  // namespace hn
