// common.cuh — shared declarations of the HEALNet-B200 kernel library (internal; the public C ABI is
// include/healnet_b200.h).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace hn {

// thread-local error string behind hn_last_error()
void set_error(const std::string& msg);
const char* get_error();

#define HN_CHECK_CUDA(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      ::hn::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
      return -10;                                                                                    \
    }                                                                                                \
  } while (0)

#define HN_REQUIRE(cond, msg)                                                                        \
  do {                                                                                               \
    if (!(cond)) {                                                                                   \
      ::hn::set_error(std::string(msg) + " [" #cond "]");                                            \
      return -1;                                                                                     \
    }                                                                                                \
  } while (0)

// ------------------------------------------------------------------ programmatic dependent launch
// Every kernel of the forward is launched with the programmatic-stream-serialization attribute and calls
// HN_PDL_LAUNCH() first thing and HN_PDL_WAIT() before it touches global memory: the next kernel's launch latency and
// prologue (barrier init, TMEM allocation, descriptor prefetch) then overlap the tail of the current one. A forward is
// ~150 short dependent kernels, so the ~2-3 us saved per launch add up. HN_PDL=0 turns the attribute off (the device
// side instructions are no-ops then).
#define HN_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define HN_PDL_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kc(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                             unsigned cluster_x, bool cooperative, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int n = 0;
  if (cooperative) {  // gang-scheduled grid: CTAs that wait for each other can never be starved by another stream
    attr[n].id = cudaLaunchAttributeCooperative;
    attr[n].val.cooperative = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                            Args... args) {
  return launch_kc(kernel, grid, block, smem, stream, 1u, false, args...);
}

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline long round_up_l(long x, long m) { return (x + m - 1) / m * m; }

// ------------------------------------------------------------------ GEMM (gemm.cu)
// C[M,N] = A[M,K] * B[N,K]^T, fp16 operands (row-major, K contiguous), fp32 accumulation in TMEM.
enum GemmEpilogue : int {
  EPI_F16 = 0,        // out_h[m][n] = half(acc + bias[n])                  (bias optional)
  EPI_GATE_F16 = 1,   // out_h[m][n/2] = (acc[n]+b[n]) * act(acc[n+1]+b[n+1]) for even n   (FeedForward gate)
  EPI_RES = 2,        // x[m][n] += acc + bias[n]                           (fp32 residual stream, in place)
  EPI_RES_LEAKY = 3,  // x[m][n] += leaky_relu_0.01(acc + bias[n])
  EPI_F32 = 4,        // out_f[m][n] = acc + bias[n]
  EPI_LEAKY_F32 = 5,  // out_f[m][n] = leaky_relu_0.01(acc + bias[n])
};
enum GateAct : int { ACT_SELU = 0, ACT_GELU = 1 };

struct GemmArgs {
  const __half* A;  // [M][lda]
  const __half* B;  // [N][ldb]
  int M, N, K;
  int lda, ldb;     // in elements, multiples of 8
  int epi;
  int act;
  const float* bias;  // [N] or null
  void* out;          // half* or float* depending on epi (x for the residual modes)
  int ldo;            // in elements of the output type
  // split-precision operands (gemm.cu header): terms = 1 plain; 2 = A.(B_hi + B_lo); 3 = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo
  int terms = 1;
  int a_seg = 0, b_seg = 0;  // column offset of the lo segment inside A / B rows (multiple of 64, >= K)
  int out_seg = 0;           // fp16 epilogues: > 0 -> lo part of the output stored at column + out_seg
  // Residual epilogues (EPI_RES / EPI_RES_LEAKY) whose output rows are complete rows of x (ldo == N <= 1024): also
  // emit LayerNorm(x_new) * gamma + beta — the PreNorm of the NEXT module — as split fp16 (hi | lo at + ln_seg) into
  // ln_out, so that module's separate LayerNorm launch disappears. Needs one tile per CTA (gemm_can_fuse_ln): every
  // CTA publishes its tile on the row block's counter, waits for the block to be complete and normalises its share
  // of the block's rows with the same two-pass arithmetic as layernorm_f16_kernel.
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  __half* ln_out = nullptr;
  int ln_ld = 0, ln_seg = 0;
  unsigned* ln_counters = nullptr;  // ceil(M / 128) words, zero at the start of the forward
  int ln_epoch = 0;                 // 1-based index of this launch among the forward's fused launches
  // operands are bf16 instead of fp16 (same split layout): the backward pass, whose gradient operands need fp32's
  // exponent range; fp32-output epilogues only
  int bf16 = 0;
  // batched form (backward pass): nbatch independent products whose operand rows are stacked — batch z reads A rows
  // [z * a_brows, z * a_brows + M) and B rows [z * b_brows, z * b_brows + N) — and whose outputs start at
  // out + (z / nb2) * out_b1 + (z % nb2) * out_b2 (elements). fp32 epilogues, no clusters, no fused LayerNorm.
  int nbatch = 1, nb2 = 1;
  long a_brows = 0, b_brows = 0, out_b1 = 0, out_b2 = 0;
  float alpha = 1.f;  // out (+)= alpha * A B^T (+ bias)
};
int launch_gemm(const GemmArgs& a, cudaStream_t stream);
// can this residual GEMM also emit the next LayerNorm (GemmArgs::ln_*)? (shape / alignment rules, HN_GEMM_LN switch)
bool gemm_can_fuse_ln(const GemmArgs& a);

// ------------------------------------------------------------------ strided batched fp32 contraction (sgemm.cu)
// C[b1][b2][m][n] (+)= alpha * sum_k A(m, k) * B(k, n); operands fp32 (type 0) or split fp16 pairs as the forward
// stores them (type 1: value = p[i] + p[i + lo_off]); all strides in elements of the operand's storage type.
struct SgOperand {
  const void* p = nullptr;
  int type = 0;
  int lo_off = 0;
  long s_row = 0, s_col = 0;   // A: (m, k) strides; B: (k, n) strides
  long s_b1 = 0, s_b2 = 0;
};
struct SgArgs {
  int M = 0, N = 0, K = 0;
  SgOperand A, B;
  float* C = nullptr;
  long c_row = 0, c_col = 1, c_b1 = 0, c_b2 = 0;
  int nb1 = 1, nb2 = 1;
  float alpha = 1.f;
  int accumulate = 0;
};
int launch_sgemm(const SgArgs& a, cudaStream_t stream);

// ------------------------------------------------------------------ row ops (rowops.cu)
// y[r] = [hi | lo] split fp16 of LN(x[r]) * gamma + beta: hi in columns [0, seg), lo in [lo_seg, lo_seg + seg)
// (lo_seg == 0: hi only); pad columns [D, seg) are zero-filled in both segments.
int launch_layernorm_f16(const float* x, int ldx, const float* gamma, const float* beta, __half* y, int ldy, int seg,
                         int lo_seg, long rows, int D, cudaStream_t stream);
// per-axis Fourier tables: tab[axis_off[a] + j][2B+1] (fp32)
int launch_axis_tables(float* tab, const int* axis_sizes, int n_axes, int n_bands, float max_freq,
                       cudaStream_t stream);
// standardised context rows, small-C layout: z[b][N][zw] fp16 = [(v-mean)*rstd (C values), 1, 0...], zw = 32 | 64;
// split != 0: z[b][N][2 zw] = [hi (zw) | lo (zw)]
// (tok0: index of the first token of this buffer on the modality's full token axis; N = tokens in the buffer)
// in_dtype: element type of the caller's raw modality buffer: 0 fp32, 1 bf16, 2 fp16
int launch_build_z_small(const void* raw, __half* z, int zw, int batch, long N, int c_raw, int n_axes,
                         const int* axis_sizes /*host*/, int n_bands, const float* tab, int fourier,
                         cudaStream_t stream, long tok0 = 0, int split = 0, int in_dtype = 0);
// standardised context rows, generic layout: z[b*N][ldz] fp16 (pad cols zero); lo_seg > 0: split [hi | lo]
int launch_build_z_large(const void* raw, __half* z, int ldz, int lo_seg, int batch, long N, int c_raw, int n_axes,
                         const int* axis_sizes /*host*/, int n_bands, const float* tab, int fourier,
                         cudaStream_t stream, long tok0 = 0, int in_dtype = 0);
// pooled head: logits[b][o] = LN(mean_L x[b]) . W[o] + bias[o]
// (pooled: batch * D floats of scratch)
int launch_head(const float* x, int batch, int L, int D, const float* ln_w, const float* ln_b, const float* W,
                const float* bias, int out_dims, float* pooled, float* logits, cudaStream_t stream);
// user mask (b, N) bytes -> per-64-token-tile bit words (bit j = keep token tile*64+j)
int launch_pack_mask(const uint8_t* mask, uint64_t* bits, int batch, long N, cudaStream_t stream);

// ------------------------------------------------------------------ attention (xattn.cu)
struct AttnArgs {
  // Q: [batch][L][q_ld] fp16, head h at columns h*KD .. (KD = 32 small-C / 64 generic); pre-scaled by 2/sqrt(dh)*log2(e)
  const __half* Q;
  int q_ld;
  // K/V: [batch][N][kv_ld] fp16. generic: head h K at k_col0 + h*64, V at v_col0 + h*64.
  // small-C ("shared"): one 32-wide standardised-context tile serves as K and V for every head.
  const __half* KV;
  long kv_ld;
  int k_col0, v_col0;
  // precise mode: Q, K, V rows also carry lo = fp16(x - hi) parts at these column offsets;
  // S = Qh.Kh + Ql.Kh + Qh.Kl, U += P.Vh + P.Vl. Small-C path: Q' lo at q_lo_off, z rows [hi (kd) | lo (kd)] (kv_ld =
  // 2 kd), S = Q'h.zh + Q'l.zh + Q'h.zl, U += P.zh. Single fp16 score operands are off by |s| 2^-11, which a peaked
  // softmax does not average away (tests/test_gpu_fullsize.py), so the forward always runs precise.
  int precise = 0;
  int q_lo_off = 0, kv_lo_off = 0;
  // precise generic path on a LONG token axis: contract P with the hi parts of V only (a convex combination of fp16
  // values: rounding stays below 2^-11 relative and averages over the tokens — what the small-context path does with
  // z); short axes (N <= 2048: few tokens, no averaging) keep P.Vh + P.Vl
  int v_hi_only = 0;
  int c_ones = 0;        // shared_kv only: index of the 1.0 column of z (= context width C)
  // shared_kv + precise, kd 32, 17 <= C <= 23: the lo half of every z row also carries the HI parts of columns 16..C-1
  // in its columns C+1.. (written by launch_build_z_small): the kernel then folds the second 16-column step of both
  // lo-order score products into ONE UMMA (five score UMMAs per tile instead of six; xattn_small.cu)
  int z_tail_merged = 0;
  int shared_kv;  // 1 = small-C path
  int kd;         // operand width per head: 64 generic; 32 or 64 (= z row width) on the small-C path
  int hp = 64;    // generic path: head pitch in Q / K / V columns and accumulator width, 64 or 128 (dim_head > 64)
  int batch, L, H;
  long N;
  int nsplit;
  const uint64_t* mask_bits;  // [batch][ceil(N/64)] or null
  float* part_acc;            // [batch][nsplit][H][L][VD] fp32 un-normalised accumulators
  float* part_ml;             // [batch][nsplit][H][L][2] (running max in log2 units, row sum (generic only))
  // generic path, nsplit == 1 only: write the normalised rows straight into O[(b*L + l)][h*hp + c] (fp16, lo part at
  // + out_lo_seg when > 0) instead of the accumulator partials — saves the combine launch (latent self-attention,
  // tabular row). part_ml is still written (the attention-weight export reads it).
  __half* out = nullptr;
  int out_ld = 0, out_lo_seg = 0;
};
int launch_attention(const AttnArgs& a, cudaStream_t stream);
int attention_pick_nsplit(int batch, int L, int H, long N);
// small-context streaming kernel (xattn_small.cu): shared_kv path, one CTA per SM, several row blocks per CTA
int launch_small_attention(const AttnArgs& a, cudaStream_t stream);
int small_attention_pick_nsplit(int batch, int L, int H, long N, int kd);

// Token-axis sharding across GPUs: every rank reduces its own splits to ONE partial per
// (sample, head, latent row) in its exchange buffer (peer-mapped through CUDA IPC) and raises a flag in every peer's
// buffer; the combine kernels then read all ranks' partials straight over NVLink, in rank order on every rank, so all
// ranks hold bit-identical latents afterwards. Exchange buffer: [XchgHeader | slot 0 | slot 1], slot = acc | ml.
constexpr int HN_MAX_PEERS = 8;
struct XchgHeader {
  unsigned long long flags[HN_MAX_PEERS];  // flags[r] = sequence number of the last exchange rank r has published
  unsigned int blocks_done;                // merge kernel: last-block-signals counter
  int error;                               // set when a wait timed out
  unsigned int pad[46];
};
static_assert(sizeof(XchgHeader) == 256, "exchange header is one 256-byte line");
struct PeerParts {
  int world = 0;  // 0: plain local partials
  int rank = 0;
  const float* acc[HN_MAX_PEERS];  // slot of every rank for this exchange (own rank included)
  const float* ml[HN_MAX_PEERS];
  XchgHeader* hdr[HN_MAX_PEERS];   // header of every rank's buffer (own: hdr[rank])
  unsigned long long seq = 0;
  long long timeout_clk = 60000000000LL;  // SM clocks a wait for a peer may last (~30 s) before it is given up
};
// overwrites out[0..n) with NaN when this rank's exchange header carries the time-out flag
int launch_poison_on_error(float* out, long n, const XchgHeader* hdr, cudaStream_t stream);
// local splits -> one partial in this rank's slot, then flag every peer (last block signals)
int launch_merge_signal(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L, int w,
                        float* slot_acc, float* slot_ml, const PeerParts& peers, cudaStream_t stream);

// combine split partials. generic: O[b*L][h*hp+d] = sum_s w_s acc_s[d] / sum_s w_s l_s   (fp16, ld = o_ld)
// (lo_seg > 0: O rows are split [hi | lo], lo at column + lo_seg)
int launch_combine_generic(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L,
                           __half* O, int o_ld, int lo_seg, int hp, cudaStream_t stream,
                           const PeerParts* peers = nullptr, int den_col = -1);
// (den_col >= 0: small-context partials, hp = zw = 32 | 64: O = acc[0..den_col) / acc[den_col], zeros beyond; the V and
// output projections are folded into one weight, pack_smallc_out)
// small-C: u = sum_s w_s acc_s[0..C) / sum_s w_s acc_s[C];  O[b*L][h*64+d] = u . Wv[h*dh+d][:] + bv[h*dh+d]
int launch_combine_vproj(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L, int C,
                         int zw, int dh, const float* Wv /*[H*dh][zw]*/, const float* bv /*[H*dh]*/, __half* O,
                         int o_ld, int lo_seg, int hp, cudaStream_t stream, const PeerParts* peers = nullptr);
// opt-in export of attn = softmax(...) of one attention call, out[(b*H + h)][l][n] fp32; call after the attention
// kernel (needs its partials' row statistics), before the buffers are reused
int launch_attn_export(const AttnArgs& a, float* out, cudaStream_t stream);
// x[b][i] = src[i]  (latent broadcast, healnet.py:225)
int launch_broadcast_rows(const float* src, float* dst, long n, int batch, cudaStream_t stream);

}  // namespace hn
