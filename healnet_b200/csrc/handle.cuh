// handle.cuh — internal state shared by the translation units that orchestrate whole passes (api.cu: forward,
// backward.cu: training-mode tape + backward): the handle behind the C ABI, the packed-weight records and the
// workspace plan. Not part of the public interface (include/healnet_b200.h).
#pragma once
#include <map>
#include <vector>

#include "../../include/healnet_b200.h"
#include "common.cuh"

namespace hn {

// head pitch: every head occupies 64 (dim_head <= 64) or 128 columns of Q / K / V / O, zero padded
constexpr int MAX_DIM_HEAD = 128;
inline int head_pitch(int dim_head) { return dim_head <= 64 ? 64 : 128; }
constexpr float LOG2E = 1.4426950408889634074f;

// bump allocator over a caller-provided (or handle-owned) device buffer; 256-byte aligned pieces
struct Arena {
  char* base = nullptr;
  size_t off = 0;
  template <typename T>
  T* take(size_t count) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

struct AttnPacked {     // one PreNorm(Attention) module
  bool small = false;   // reassociated small-context path (cross-attention with C <= 63)
  int zw = 0;           // small path: width of a z row (32 or 64)
  int C = 0;            // context width (cross) or D (self)
  // generic / self
  // all fp16 weight rows are split [hi (seg cols) | lo (seg cols)], seg = round_up(K, 64)
  __half* Wq = nullptr;     // cross generic: [H*64][2 segD] scaled;  self: [3*lH*64][2 segD] (Q scaled | K | V)
  __half* Wkv = nullptr;    // cross generic: [2*H*64][2 segC], context-LN gamma folded
  float* bkv = nullptr;     // cross generic: [2*H*64], context-LN beta folded (K part zero: it cancels in softmax)
  // small (used when the token axis is long; short axes take the generic precise path whatever C is)
  __half* WqS = nullptr;    // [H*zw][2 segD]  Wk^T Wq reassociated, gamma and scale folded
  float* Wv = nullptr;      // [I][zw]  gamma folded
  float* bv = nullptr;      // [I]      beta folded
  __half* WoS = nullptr;    // [D][2 seg(H*zw)]  Wo . Wv' folded (split hi | lo): the small path's out-projection weight
  float* boS = nullptr;     // [D]      bo + Wo . bv
  __half* Wo = nullptr;     // [D][2 * H*64] head-padded columns
};
struct FFPacked {
  __half* W1 = nullptr;  // [8D][2 segD] rows interleaved (a_j, g_j)
  float* b1 = nullptr;   // [8D] interleaved
  __half* W2 = nullptr;  // [D][2 seg4D]
};
// One PreNorm(module) application of a training-mode forward, in execution order, with the byte offsets of what it
// left on the tape for hn_backward (backward.cu).
struct BlockRec {
  int kind = 0;      // 0 cross-attention (small-context path), 1 cross-attention (generic path), 2 latent self-attention,
                     // 3 feed-forward
  int layer = 0;
  int m = 0;         // modality; n_modalities for the latent block
  size_t x_in = 0;   // fp32 [rows][D]: residual stream entering the block
  size_t xn = 0;     // split fp16 [rows][2 segD]: LayerNorm(x_in)
  size_t o = 0;      // attention: split fp16 normalised attention output (u rows on the small path); FF: split hidden rows
  size_t stats = 0;  // attention: fp32 [(b*H + h)*L + l][2] = (row max in log2 units, denominator)
};
struct TrainState {
  bool valid = false;
  int batch = 0;
  int axis_sizes[HN_MAX_MODALITIES * HN_MAX_AXES] = {};
  bool present[HN_MAX_MODALITIES] = {};
  int skip[HN_MAX_MODALITIES] = {};
  long mask_tokens = 0;
  std::vector<BlockRec> blocks;
  size_t x_final = 0;   // fp32 [rows][D]: residual stream after the last block
  size_t bytes = 0;
};

struct SlotKey {
  std::vector<const void*> ptrs;
  bool operator<(const SlotKey& o) const { return ptrs < o.ptrs; }
};


}  // namespace hn

struct hn_handle {
  hn_desc d;
  int M = 0, I = 0, lI = 0;
  int hpx = 64, hpl = 64;         // head pitch of the cross / latent attention (64 or 128 columns per head)
  int segD = 0, seg4D = 0;        // hi/lo segment widths of D-wide / 4D-wide split operands (multiples of 64)
  int C[HN_MAX_MODALITIES];       // context width per modality
  // registered fp32 parameters: index (layer + 1) * slots_per_layer + slot
  int slots_per_layer = 0;
  std::vector<std::vector<const float*>> w;
  // packed store
  void* packed = nullptr;
  size_t packed_bytes = 0;
  bool packed_valid = false;
  std::vector<hn::AttnPacked> attn;   // [layer][M + 1] (index M = latent self-attention)
  std::vector<hn::FFPacked> ff;       // [layer][M + 1]
  int launches = 0;
  int io_dtype = 0;               // element type of the modality input buffers: 0 fp32, 1 bf16, 2 fp16 (hn_set_io_dtype)
  // optional per-launch timing of the cross-attention kernels (bench.py roofline): CUDA event pairs on the
  // forward's own stream, one pair per (layer, modality), read back after the caller synchronises
  // opt-in attention-weight export buffers, index layer * (M + 1) + module (M = latent self-attention); null = off
  std::vector<float*> export_ptrs;
  bool profile = false;
  std::vector<cudaEvent_t> ev;          // 2 per slot
  std::vector<int> ev_mod;              // modality of each recorded slot in the last forward
  std::vector<int> ev_kind;             // 0 cross-attention kernel, 1 K/V projection GEMM, 2 context-row build
  std::vector<double> ev_flops;         // tensor-core FLOPs the launch executed (padded tiles included)
  std::vector<double> ev_useful;        // unpadded algorithmic FLOPs of the same launch
  std::vector<double> ev_exps;          // softmax exponentials the launch evaluated
  // token-axis sharding across GPUs (hn_set_exchange): peer-mapped exchange buffers, own rank included
  int x_rank = 0, x_world = 0;
  char* x_bufs[hn::HN_MAX_PEERS] = {};
  size_t x_bytes = 0;
  unsigned long long x_seq = 0;         // exchanges published so far (all ranks advance in lock step)
  long long x_timeout_clk = 60000000000LL;  // peer-wait bound in SM clocks (hn_set_exchange_timeout)
  // training: gradient buffers registered per slot (same indexing as w), the record of the last training-mode forward
  std::vector<std::vector<float*>> g;
  hn::TrainState train;
  char* tape = nullptr;   // non-null only while a training-mode forward is recording
  int bwd_variant = 0;    // 0: tensor-core kernels for the heavy backward contractions; 1: exact fp32 SIMT checkers
};


namespace hn {

inline int slot_index(const hn_handle* h, int layer, int slot) { return (layer + 1) * h->slots_per_layer + slot; }

inline int ctx_ld(int C) { return round_up(C, 8); }
inline int seg_of(int K) { return round_up(K, 64); }
// Token axes up to this length are never streamed by the small-context kernel nor sharded across GPUs. (Round 1 also
// used it as the limit of the precise — split hi/lo — attention; the full-size peaked-softmax parity cases showed that
// single fp16 score operands are not enough on long axes either, so every attention now runs precise.)
constexpr long PRECISE_MAX_TOKENS = 2048;

struct ModPlan {
  bool present = false;
  bool small = false;
  int zw = 0;       // small: z row width
  int ldz = 0;      // generic: z row pitch
  bool precise = false;  // generic: short token axis -> split z / K / V / Q and the precise attention kernel
  int segC = 0;
  int C = 0, c_raw = 0, n_axes = 0;
  int axes[HN_MAX_AXES];
  long N = 0;       // tokens of the modality (decides the path, so every rank of a token-sharded run agrees)
  long Nl = 0;      // tokens held by this rank (== N unless the token axis is sharded across GPUs)
  long tok0 = 0;    // first local token on the full axis
  bool sharded = false;
  int nsplit = 1;
  bool masked = false;
  float* tab = nullptr;
  __half* z = nullptr;
};

struct Workspace {
  ModPlan mod[HN_MAX_MODALITIES];
  float* x = nullptr;        // [b*L][D] fp32 residual stream
  __half* xn = nullptr;      // [b*L][2 segD]       split
  __half* q = nullptr;       // [b*L][2 qw]         split (small-C Q': hi only)
  __half* o = nullptr;       // [b*L][2 ow]         split
  __half* hid = nullptr;     // [b*L][2 seg4D]      split
  __half* kv = nullptr;      // [b*Nmax][2*H*64] (x2 when precise)   (generic cross-attention only)
  bool self_precise = false;
  float* part_acc = nullptr;
  float* part_ml = nullptr;
  uint64_t* mask_bits = nullptr;
  float* pooled = nullptr;   // [b][D] mean over latents (head)
  unsigned* ln_counters = nullptr;  // per 128-row block arrival counters of the fused LayerNorm (gemm.cu)
  int self_nsplit = 1;
  size_t bytes = 0;
};


// Lays the tape of a training-mode forward out (execution order of forward_impl): fills `blocks`, returns the bytes.
size_t plan_tape(const hn_handle* h, int batch, const Workspace& ws, const int* skip_latent_block,
                 std::vector<BlockRec>& blocks, size_t& x_final);
// the forward, optionally recording onto h->tape (api.cu)
int forward_impl(hn_handle* h, int batch, const void* const* modality_ptrs, void* const* modality_ready_events,
                 const int* axis_sizes, const long* tok_begin, const long* tok_count, const int* skip_latent_block,
                 const uint8_t* mask, long mask_tokens, float* latents_out, float* logits_out, void* workspace,
                 size_t workspace_bytes, void* cuda_stream);

// Lays the forward workspace out over `base` (null: sizing pass); defined in api.cu
int plan_workspace(const hn_handle* h, int batch, const int* axis_sizes, const bool* present, long mask_tokens,
                   char* base, Workspace& ws, const long* tok_begin = nullptr, const long* tok_count = nullptr);

}  // namespace hn
