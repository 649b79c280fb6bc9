// pack.cuh — weight repacking launchers (pack.cu). Every fp16 destination row is `seg` columns wide (zero
// padded beyond the source width); lo_off > 0 additionally stores lo = fp16(v - fp16(v)) at column lo_off + k
// (split-precision operands, see gemm.cu). hp = head pitch: 64 (dim_head <= 64) or 128 columns per head.
#pragma once
#include "common.cuh"

namespace hn {
int pack_headpad_rows(__half* dst, int ld_dst, int dst_row0, const float* src, int ld_src, int src_row0,
                      int n_heads, int dh, int K, float scale, const float* colscale, int seg, int lo_off, int hp,
                      cudaStream_t st);
int pack_headpad_cols(__half* dst, int ld_dst, const float* src, int ld_src, int rows, int n_heads, int dh, int seg,
                      int lo_off, int hp, cudaStream_t st);
int pack_ff1(__half* dst, int ld_dst, float* bias_dst, const float* W, const float* bias, int D, int hidden, int seg,
             int lo_off, cudaStream_t st);
int pack_plain(__half* dst, int ld_dst, const float* src, int ld_src, int rows, int K, int seg, int lo_off,
               cudaStream_t st);
int fold_beta_headpad(float* bias_dst, int dst_row0, const float* W, int ld, int src_row0, int n_heads, int dh,
                      int C, const float* beta, int hp, cudaStream_t st);
int pack_smallc_q(__half* Aq, int ld_dst, const float* Wq, const float* Wkv, const float* gamma, int H, int D, int C,
                  int dh, float scale, int zw, int seg, int lo_off, cudaStream_t st);
int pack_smallc_v(float* Wv_dst, float* bv_dst, const float* Wkv, const float* gamma, const float* beta, int inner,
                  int C, int zw, cudaStream_t st);
// small-C path: V projection folded into the output projection (see pack.cu); Wv / bv are pack_smallc_v's outputs
int pack_smallc_out(__half* WoS, int ld_dst, float* boS, const float* Wo, const float* bo, const float* Wv,
                    const float* bv, int D, int inner, int H, int dh, int zw, int seg, int lo_off, cudaStream_t st);
}  // namespace hn
