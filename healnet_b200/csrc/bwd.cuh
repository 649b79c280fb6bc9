// bwd.cuh — internal declarations of the backward-pass kernels (bwdops.cu) used by backward.cu.
#pragma once
#include "common.cuh"

namespace hn {

// column reductions over R rows: mode 0: sum a ; 1: sum a*b ; 2: sum a * (b - mu_r) * rstd_r (stats = (mu, rstd) per row).
// out[c] = (accumulate ? out[c] : 0) + scale * sum. partial: scratch of colsum_slabs(R) * C floats.
int colsum_slabs(long R);
int launch_colsum(int mode, const float* a, long lda, const float* b, long ldb, const float* stats, long R, int C,
                  float scale, float* out, int accumulate, float* partial, cudaStream_t st);
// LayerNorm backward over rows: dx (+)= d LN(x) / dx applied to dy ; stats[r] = (mu, rstd)
int launch_ln_bwd_rows(const float* x, const float* dy, const float* gamma, float* dx, float* stats, long rows, int D,
                       int accumulate, cudaStream_t st);
int launch_leaky_bwd(const float* dx, const float* x_out, const float* x_in, float* dy, long n, cudaStream_t st);
// h [rows][2F] = [a | g] without bias -> [d a | d g] in place
int launch_gate_bwd(float* h, const float* b1, const float* dhid, long rows, int F, int snn, cudaStream_t st);
int launch_batch_sum(const float* in, float* out, long n, int batch, cudaStream_t st);
int launch_head_bwd(const float* x, int batch, int L, int D, const float* ln_w, const float* ln_b, const float* W,
                    int out_dims, const float* dlogits, float* g_ln_w, float* g_ln_b, float* g_W, float* g_bias,
                    float* dpooled, float* dx, cudaStream_t st);
// merged (M, den) per (sample, head, latent row) from the split partials of a forward attention launch
int launch_row_stats(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L, int w,
                     int den_col, float den_scale, float* stats, cudaStream_t st);
int launch_softmax_recompute(float* S, const float* stats, int H, int L, long N, const uint64_t* mask_bits, long total,
                             cudaStream_t st);
// dP rows -> dt = P * (dP - sum_n P dP), in place
int launch_softmax_bwd(float* dP, const float* P, long N, long n_rows, cudaStream_t st);
int launch_small_pre(const __half* u, int ld, int lo_off, int zw, int H, int C, long rows, const float* gamma,
                     const float* beta, float* u32, float* cnu, cudaStream_t st);
int launch_small_du(const float* g, const float* u32, const float* gamma, int C, long R, float* du, float* delta,
                    cudaStream_t st);
int launch_scale_cols(const float* in, const float* gamma, float scale, int C, long n, float* out, cudaStream_t st);
int launch_kv_fold_bwd(const float* dWp, const float* W, const float* gamma, const float* beta, const float* sv,
                       int rows2I, int I, int C, float* gW, float* ggamma, float* gbeta, cudaStream_t st);

// bf16 hi/lo operand rows for the tensor-core GEMMs of the backward (hi = bf16(v), lo = bf16(v - hi): 16 significant
// bits with fp32's exponent range)
// general strided / batched form, both operands of one product in ONE launch: dst[(z R + r) 2 seg + c] = hi | lo of
// p[b1 s_b1 + b2 s_b2 + r s_row + c s_col] (z = b1 nb2 + b2; type 0 = fp32, 1 = split fp16 with the lo part at
// + lo_off), c < K, pad columns zero. The operands are given as (rows, contraction index) views.
int launch_pack_bf16_pair(const SgOperand& A, long RA, const SgOperand& B_as_rows, long RB, int K, int nb1, int nb2,
                          void* dstA, void* dstB, int seg, cudaStream_t st);
// gate backward on interleaved (a_j, g_j) pre-activations (bias included) -> dh = [d a | d g]
int launch_gate_bwd_il(const float* h_il, const float* dhid, float* dh, long rows, int F, int snn, cudaStream_t st);

// streaming backward of the small-context cross-attention
struct SmallBwdArgs {
  const float* r;       // [R][C]  log2-unit score vectors: s_lt = r_l . z_t
  const float* du;      // [R][C]
  const float* delta;   // [R]     du_l . u_l
  const float* stats;   // [(b*H + h)*L + l][2] = (M, den)
  const __half* z;      // [b][N][z_ld] split rows: hi at column c, lo at z_lo + c
  int z_ld, z_lo;
  const uint64_t* mask_bits;
  float* dr_part;       // [nsplit][R][C] scratch
  int batch, H, L, C, nsplit;
  long N, R_total;
};
int small_attn_bwd_nsplit(int batch, int H, int L, long N);

// tensor-core version of the same pass (xattn_small.cu: attn_small_bwd_kernel). R / DU: split fp16 rows
// [b][L][rq_ld], head h at columns h*kd (hi) and lo_off + h*kd (lo); R carries P_SHIFT - M in column C of its hi part;
// DU rows scaled by a per-row power of two; part: [b][nsplit][H][L][kd] fp32 partial accumulators of dR.
struct SmallBwdTcArgs {
  const __half* rq;
  const __half* duq;
  int rq_ld, lo_off;
  const __half* z;      // [b][N][2 kd] = [hi | lo]
  int kd;
  const float* row_a;   // [(b*L + l)*H + h]
  const float* row_d;
  const uint64_t* mask_bits;
  float* part;
  int batch, H, L, nsplit;
  long N;
  int C = 63;           // context width: products only run over the 16-column steps that hold context columns
  int merged_tail = 0;  // kd 32, 17 <= C <= 23: R_lo rows and z lo rows carry the merged tail (launch_small_bwd_prep)
};
int launch_small_attention_bwd(const SmallBwdTcArgs& a, cudaStream_t stream);
// r / du / delta / stats -> the kernel's operands (R, DU split rows, per-row factors); scale[R] = the power of two DU was
// multiplied with. Buffers rq / duq must hold rows * rq_ld halves.
// kd 32 and 17 <= C <= 23: the lo half of an R row is written with the merged tail of the score product
// ([R_lo 0..15 | R_hi 16..C-1, 0, R_lo 16..C-1, 0..], xattn_small.cu) — set SmallBwdTcArgs::merged_tail accordingly.
int launch_small_bwd_prep(const float* r, const float* du, const float* delta, const float* stats, int batch, int H,
                          int L, int C, int kd, __half* rq, __half* duq, int rq_ld, int lo_off, float* row_a,
                          float* row_d, float* scale, cudaStream_t st);
// dr[R][C] = (sum over splits of part) / scale
int launch_small_bwd_finish(const float* part, const float* scale, int batch, int nsplit, int H, int L, int C, int kd,
                            float* dr, cudaStream_t st);
int launch_small_attn_bwd(const SmallBwdArgs& a, float* dr, cudaStream_t st);

}  // namespace hn
