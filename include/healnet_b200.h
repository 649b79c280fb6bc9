/* healnet_b200.h — C ABI of the B200-native HEALNet fusion-forward library (libhealnet_b200.so).
 *
 * The reference (konst-int-i/healnet) has no FFI / plugin interface: its boundary for this path is the
 * Python nn.Module surface of `HealNet` / `Attention` (healnet/models/healnet.py:14-262, 369-426).
 * This header is the native boundary underneath the drop-in module `healnet_b200.HealNet`; every entry
 * point names the reference code it replaces. Conventions: plain pointers and sizes only, all buffers
 * caller-owned device memory (fp32 unless stated), every call is stream-ordered on the given CUDA stream
 * and never synchronises the host, returns 0 on success or a negative code (message via
 * hn_last_error()), never throws or exits. One handle may be used from one stream at a time.
 */
#ifndef HEALNET_B200_H_
#define HEALNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HN_API __attribute__((visibility("default")))
#else
#define HN_API
#endif

#define HN_MAX_MODALITIES 16
#define HN_MAX_AXES 4

typedef struct hn_handle hn_handle;

/* Constructor hyper-parameters; field names follow HealNet.__init__ (healnet/models/healnet.py:15-38). */
typedef struct hn_desc {
  int n_modalities;
  int depth;
  int l_c;              /* number of latents (rows of the latent array) */
  int l_d;              /* latent width */
  int x_heads;
  int cross_dim_head;
  int l_heads;
  int latent_dim_head;
  int num_freq_bands;
  int out_dims;
  int self_per_cross_attn;   /* 0 or 1 (the reference breaks for >= 2, healnet.py:242) */
  /* cross_dim_head and latent_dim_head: 1..128 (latent_dim_head is unused when self_per_cross_attn == 0) */
  int snn;                   /* 1: a*selu(g) gate, 0: a*gelu(g) (healnet.py:323-331,342) */
  int final_classifier_head;
  int fourier_encode_data;
  float max_freq;
  int channel_dims[HN_MAX_MODALITIES];
  int num_spatial_axes[HN_MAX_MODALITIES];
} hn_desc;

/* Replaces HealNet.__init__ (healnet.py:121-185): validates the configuration, sizes the packed-weight
 * store. No parameters are created here — they stay torch nn.Parameters owned by the caller. */
HN_API int hn_create(const hn_desc* desc, hn_handle** out);
HN_API int hn_destroy(hn_handle* h);

/* Thread-local description of the last failure on the calling thread. */
HN_API const char* hn_last_error(void);

/* Registers the (borrowed, fp32, contiguous, device) parameter tensors of one module, in state_dict
 * naming:
 *   layer >= 0, slot 2m     cross-attention of modality m (healnet.py:146-149):
 *        {norm.weight, norm.bias, norm_context.weight, norm_context.bias,
 *         fn.to_q.weight, fn.to_kv.weight, fn.to_out.0.weight, fn.to_out.0.bias}            n = 8
 *   layer >= 0, slot 2m+1   cross feed-forward (healnet.py:153):
 *        {norm.weight, norm.bias, fn.net.0.weight, fn.net.0.bias, fn.net.2.weight, fn.net.2.bias}  n = 6
 *   layer >= 0, slot 2M     latent self-attention (healnet.py:151):
 *        {norm.weight, norm.bias, fn.to_q.weight, fn.to_kv.weight, fn.to_out.0.weight, fn.to_out.0.bias}  n = 6
 *   layer >= 0, slot 2M+1   latent feed-forward (healnet.py:154): as the cross feed-forward      n = 6
 *   layer == -1, slot 0     {latents}                                                        n = 1
 *   layer == -1, slot 1     {to_logits.1.weight, to_logits.1.bias, to_logits.2.weight, to_logits.2.bias} n = 4
 * Tied layers (weight_tie_layers, healnet.py:161, 278-290) simply register the same pointers again. */
HN_API int hn_set_weights(hn_handle* h, int layer, int slot, const void* const* dev_ptrs, int n);

/* Repacks the registered fp32 parameters into the fp16 tensor-core operand layouts (scale / LayerNorm
 * affine folding, head padding, gate interleave). Call once after hn_set_weights and again whenever
 * parameter values change. Allocates the packed store on first use. */
HN_API int hn_pack_weights(hn_handle* h, void* cuda_stream);

/* Scratch requirement of hn_forward for `batch` samples whose modality m has spatial extents
 * axis_sizes[m*HN_MAX_AXES + a] (a < num_spatial_axes[m]). Returns 0 on error. */
HN_API size_t hn_workspace_bytes(const hn_handle* h, int batch, const int* axis_sizes);

/* Replaces HealNet.forward (healnet.py:190-250), default verbose=False semantics.
 *   modality_ptrs[m] : (batch, *axes_m, channel_dims[m]) row-major channel-last, fp32 unless hn_set_io_dtype says
 *                      otherwise, or NULL when the
 *                      modality is missing (its cross-attention + cross-FF are skipped, the latent
 *                      self-attention block still runs — healnet.py:229-245).
 *   skip_latent_block: optional n_modalities flags; nonzero also skips the latent self-attention + FF that
 *                      follows modality m (the reference's `verbose=True` + missing-modality path, healnet.py:229-232).
 *   mask             : optional uint8 (batch, mask_tokens), nonzero = keep (healnet.py:411-415); applied to
 *                      every cross-attention whose token count equals mask_tokens.
 *   latents_out      : optional fp32 (batch, l_c, l_d) — the `return_embeddings=True` result.
 *   logits_out       : optional fp32 (batch, out_dims)  — to_logits(x) (healnet.py:181-185,250).
 * The caller's input buffers are never written (the reference mutates its input list, healnet.py:222). */
HN_API int hn_forward(hn_handle* h, int batch, const void* const* modality_ptrs, const int* axis_sizes,
                      const int* skip_latent_block, const uint8_t* mask, long mask_tokens, float* latents_out,
                      float* logits_out, void* workspace, size_t workspace_bytes, void* cuda_stream);

/* Element type of the modality input buffers handed to hn_forward* from now on: 0 = fp32 (default), 1 = bf16,
 * 2 = fp16 — the reference's dtype follows its inputs (healnet.py:212; BASELINE config 3 runs in bf16). The rows are
 * widened on the fly while they are standardised (no fp32 copy of the inputs exists anywhere); parameters are
 * registered in fp32, outputs are fp32. */
HN_API int hn_set_io_dtype(hn_handle* h, int dtype);

/* hn_forward with per-modality "input ready" events: modality_ready_events[m] (a cudaEvent_t, or NULL) is waited on
 * by the forward's stream right before modality m's buffer is first read (just ahead of its layer-0 cross-attention),
 * so a caller can copy a large late modality host-to-device on another stream while the earlier modalities are
 * already being processed (healnet/main.py:415 copies all features up front). modality_ready_events may be NULL. */
HN_API int hn_forward_ex(hn_handle* h, int batch, const void* const* modality_ptrs, void* const* modality_ready_events,
                         const int* axis_sizes, const int* skip_latent_block, const uint8_t* mask, long mask_tokens,
                         float* latents_out, float* logits_out, void* workspace, size_t workspace_bytes,
                         void* cuda_stream);

/* Token-axis ("split-N") sharding of one forward across the GPUs of a node — SURVEY.md section 8 row f4; the
 * reference has no distributed code (its only multi-sample axis is the batch, healnet/main.py:415-432). Every rank
 * holds the whole (small) batch and, for each long modality, only tokens [tok_begin[m], tok_begin[m] + tok_count[m])
 * of its row-major flattened token axis: modality_ptrs[m] is then fp32 (batch, tok_count[m], channel_dims[m]) while
 * axis_sizes still describes the FULL modality (positions are global). tok_count[m] <= 0 or == the full count means
 * "replicated" (short axes, <= 2048 tokens, must be). Each rank runs the streaming attention over its tokens; the
 * per-row partials (running max, un-normalised accumulator, denominator) are exchanged through peer-mapped buffers
 * (CUDA IPC over NVLink) and merged by the combine kernel in rank order, so every rank ends up with bit-identical
 * latents / logits. mask (if any) covers the local tokens: (batch, mask_tokens == tok_count[m]).
 * Set-up, once per process group:  bytes = hn_exchange_bytes(h, max_batch); hn_exchange_alloc(bytes, &mine, handle);
 * all-gather the 64-byte handles; hn_exchange_open() the peers'; hn_set_exchange(h, rank, world, bufs, bytes).
 * All ranks must issue the same sequence of hn_forward_split calls. A peer that never shows up makes the waiting
 * kernels give up after the exchange time-out (~30 s, hn_set_exchange_timeout) instead of hanging the GPU: the outputs
 * of that forward are NaN and hn_exchange_error* report it. */
HN_API size_t hn_exchange_bytes(const hn_handle* h, int batch);
HN_API int hn_exchange_alloc(size_t bytes, void** dev_ptr, unsigned char* ipc_handle_out /* 64 bytes */);
HN_API int hn_exchange_open(const unsigned char* ipc_handle /* 64 bytes */, void** peer_ptr);
HN_API int hn_exchange_close(void* peer_ptr);
HN_API int hn_exchange_free(void* dev_ptr);
HN_API int hn_set_exchange(hn_handle* h, int rank, int world, void* const* bufs /* [world], own buffer at [rank] */,
                           size_t bytes);
HN_API int hn_exchange_error(const hn_handle* h, int* error_out);  /* synchronous read of the time-out flag */
/* stream-ordered read of the flag into PINNED host memory (check it once a later event on the stream has completed) */
HN_API int hn_exchange_error_async(const hn_handle* h, int* pinned_host_out, void* cuda_stream);
/* how long a combine kernel waits for a peer's partials before giving up (default ~30 s). A forward whose wait timed
 * out overwrites its outputs with NaN and leaves the flag set until hn_set_exchange is called again. */
HN_API int hn_set_exchange_timeout(hn_handle* h, double seconds);
HN_API size_t hn_workspace_bytes_split(const hn_handle* h, int batch, const int* axis_sizes, const long* tok_count);
HN_API int hn_forward_split(hn_handle* h, int batch, const void* const* modality_ptrs,
                            void* const* modality_ready_events, const int* axis_sizes, const long* tok_begin,
                            const long* tok_count, const int* skip_latent_block, const uint8_t* mask,
                            long mask_tokens, float* latents_out, float* logits_out, void* workspace,
                            size_t workspace_bytes, void* cuda_stream);

/* ---- training step ---------------------------------------------------------------
 * Replaces what torch.autograd does for the reference's training loop (healnet/main.py:426-467: logits = model(x);
 * loss.backward()): the gradient of the forward with respect to every parameter. The loss, the L1 regulariser
 * (healnet/utils/train_utils.py:5-14) and the optimizer stay with the caller.
 *   hn_forward_train : hn_forward_ex that also records, per PreNorm(module) block, what the backward needs (the fp32
 *                      residual stream, its LayerNorm, the normalised attention output / gated hidden rows, softmax row
 *                      statistics) on a caller-owned tape of hn_tape_bytes() bytes. Nothing of size tokens x latents
 *                      is stored: attention probabilities are recomputed by hn_backward. One recording per handle.
 *   hn_set_grads     : registers fp32 gradient buffers, same (layer, slot) layout and tensor order as hn_set_weights
 *                      (register the weights first). hn_backward ACCUMULATES into them (+=): zero them per step; tied
 *                      layers register the same buffers again and the contributions add up.
 *   hn_backward      : given d loss / d logits (batch, out_dims) OR d loss / d latents (batch, l_c, l_d) of the recorded
 *                      forward, the SAME workspace that forward used (it still holds the standardised context rows),
 *                      its tape and hn_backward_scratch_bytes() bytes of scratch: accumulates every parameter gradient.
 *                      No gradient is produced for the inputs (the reference's data tensors never require grad).
 * Token-sharded forwards and dropout (reference default 0) are not supported in training mode. */
HN_API size_t hn_tape_bytes(const hn_handle* h, int batch, const int* axis_sizes);
HN_API size_t hn_backward_scratch_bytes(const hn_handle* h, int batch, const int* axis_sizes);
HN_API int hn_forward_train(hn_handle* h, int batch, const void* const* modality_ptrs, void* const* modality_ready_events,
                            const int* axis_sizes, const int* skip_latent_block, const uint8_t* mask, long mask_tokens,
                            float* latents_out, float* logits_out, void* workspace, size_t workspace_bytes, void* tape,
                            size_t tape_bytes, void* cuda_stream);
HN_API int hn_set_grads(hn_handle* h, int layer, int slot, void* const* dev_ptrs, int n);
/* 0 (default): the heavy contractions of hn_backward run on tensor cores (streaming small-context pass on tcgen05 with
 * fp16 hi/lo operands, large regular GEMMs with bf16 hi/lo operands); 1: the exact fp32 SIMT kernels they are checked
 * against (tests/test_gpu_backward.py compares both with the reference's gradients). */
HN_API int hn_set_backward_variant(hn_handle* h, int variant);
HN_API int hn_backward(hn_handle* h, const float* grad_latents, const float* grad_logits, void* workspace,
                       size_t workspace_bytes, const void* tape, size_t tape_bytes, void* scratch, size_t scratch_bytes,
                       void* cuda_stream);

/* Number of kernels hn_forward enqueued on its last call for this handle (for bench accounting). */
HN_API int hn_last_launch_count(const hn_handle* h);

/* Opt-in replacement for `Attention.attn_weights` (healnet.py:420) / HealNet.get_attention_weights() (:252-262).
 * Registers (dev_out != NULL) or clears a caller-owned fp32 device buffer of shape (batch * heads, l_c, N) for the
 * attention call of `module` in `layer`: module m < n_modalities = cross-attention of modality m (N = its token
 * count, heads = x_heads); module == n_modalities = latent self-attention (N = l_c, heads = l_heads; it runs once
 * per modality and the buffer keeps the last call, as the reference's attribute does). Subsequent hn_forward calls
 * fill it with the softmax matrix actually applied. The streaming kernels never materialise this matrix (9.87 GB per
 * sample and layer at the README shapes), so export is off by default and the caller sizes the buffers. */
HN_API int hn_set_attention_export(hn_handle* h, int layer, int module, float* dev_out);

/* Measurement hook (bench.py roofline): when enabled, hn_forward brackets the heavy launches with CUDA event pairs on
 * its own stream. kind 0 = the streaming cross-attention kernel of a modality, 1 = the K/V projection GEMM of the
 * generic (wide-context) path, 2 = the context-row build (Fourier tables + standardisation). After the caller has
 * synchronised the stream, hn_profile_read sums, for one (kind, modality), the device time of those launches in the
 * LAST forward, their count, the tensor-core FLOPs they executed (padded tiles included), the unpadded algorithmic
 * FLOPs of the same contractions (SURVEY.md section 8d accounting) and the softmax exponentials evaluated. */
HN_API int hn_profile_enable(hn_handle* h, int on);
HN_API int hn_profile_read(hn_handle* h, int kind, int modality, float* ms, int* launches, double* flops_executed,
                           double* flops_useful, double* exps);

/* Replaces Attention.forward (healnet.py:400-426) for the stand-alone `Attention` module:
 *   out = LeakyReLU_0.01( softmax(2 q k^T / sqrt(dim_head)) v  Wo^T + bo ),  q = x Wq^T, [k,v] = ctx Wkv^T.
 * x (batch, n_q, query_dim), context (batch, n_ctx, context_dim) (NULL: self-attention on x),
 * weights in reference layout (to_q.weight, to_kv.weight, to_out.0.weight, to_out.0.bias), all fp32.
 * workspace: hn_attention_workspace_bytes(...) bytes of device scratch. */
HN_API size_t hn_attention_workspace_bytes(int batch, int n_q, long n_ctx, int query_dim, int context_dim, int heads,
                                    int dim_head);
HN_API int hn_attention_forward(int batch, int n_q, long n_ctx, int query_dim, int context_dim, int heads, int dim_head,
                         const float* x, const float* context, const float* w_q, const float* w_kv,
                         const float* w_out, const float* b_out, const uint8_t* mask, float* out, void* workspace,
                         size_t workspace_bytes, void* cuda_stream);
/* Same; weights_packed != 0 promises that `workspace` still holds the packed weights of an earlier call with the SAME
 * shapes and the same (unchanged) weight tensors, so the four weight-packing launches are skipped (the module-level
 * Attention keeps one workspace per module and tracks its parameters' versions). */
HN_API int hn_attention_forward_cached(int batch, int n_q, long n_ctx, int query_dim, int context_dim, int heads,
                                       int dim_head, const float* x, const float* context, const float* w_q,
                                       const float* w_kv, const float* w_out, const float* b_out, const uint8_t* mask,
                                       float* out, void* workspace, size_t workspace_bytes, int weights_packed,
                                       void* cuda_stream);

/* ---- kernel-level entry points (used by the parity tests and profiling scripts) ------------------ */
/* C[M,N] = A[M,K] B[N,K]^T on tcgen05; A, B fp16 row-major; epi: 0 f16 out, 1 gated f16 out (N/2 cols),
 * 2 fp32 residual +=, 3 fp32 residual += leaky, 4 fp32 out, 5 fp32 leaky out; act: 0 selu, 1 gelu.
 * Split precision: operand rows may hold [hi | lo] fp16 pairs (lo = fp16(x - hi)) with the lo part a_seg / b_seg
 * columns (multiple of 64, >= K) after the hi part; terms 1: A.B, 2: A.(B_hi + B_lo),
 * 3: A_hi.B_hi + A_lo.B_hi + A_hi.B_lo. out_seg > 0 (fp16 epilogues): lo part of the result at column + out_seg. */
HN_API int hn_op_gemm(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int epi, int act,
                      const float* bias, void* out, int ldo, int terms, int a_seg, int b_seg, int out_seg,
                      void* cuda_stream);
/* y rows = [hi (seg cols) | lo at column lo_seg (0: none)] of LayerNorm(x) * gamma + beta, zero padded to seg. */
HN_API int hn_op_layernorm_f16(const float* x, int ldx, const float* gamma, const float* beta, void* y, int ldy,
                               int seg, int lo_seg, long rows, int D, void* cuda_stream);
/* Fourier tables + standardised context rows z (fp16). small == 1: dense (batch, N, ldz) rows, ldz = 32 or 64,
 * with the ones column at index C < ldz; small == 2: the same as split rows (batch, N, 2 ldz) = [hi | lo],
 * lo = fp16(value - hi) (what the forward streams); small == 0: (batch*N, ldz).
 * tab: scratch of sum(axis sizes)*(2*bands+1) floats. */
HN_API int hn_op_build_context(const float* raw, void* z, int ldz, int small, int batch, int c_raw, int n_axes,
                        const int* axis_sizes, int n_bands, float max_freq, int fourier, float* tab,
                        void* cuda_stream);
/* token-axis split count the library would choose; small_kd = 32 | 64 for the small-context kernel, 0 generic */
HN_API int hn_op_attention_nsplit(int batch, int L, int H, long N, int small_kd);
/* Streaming attention partials + combine. shared_kv == 1: small-C path (Q rows kd = kv_ld = 32 | 64 wide per head,
 * KV = z rows whose column c_ones is 1.0; Q column c_ones must be 0); shared_kv == 3: the same on split operands —
 * Q' rows [hi (H kd) | lo at column q_ld / 2], z rows [hi (kd) | lo (kd)] (kv_ld = 2 kd), scores from three fp16
 * products; shared_kv == 4: as 3, on z rows whose lo half carries the merged tail the context-row builder writes for
 * kd 32 and 17 <= C <= 23 (lo columns C+1 .. 2C-16 = hi columns 16 .. C-1; the forward's mode: five score products per
 * tile instead of six); part_acc rows are kd (small-C) or 64 (generic: head_pitch = 64 | 128) floats wide;
 * head_pitch is the column pitch of one head in Q / K / V / O. */
HN_API int hn_op_attention(const void* Q, int q_ld, const void* KV, long kv_ld, int k_col0, int v_col0, int shared_kv,
                           int c_ones, int head_pitch, int batch, int L, int H, long N, int nsplit, const uint8_t* mask,
                           void* mask_bits_scratch, float* part_acc, float* part_ml, void* cuda_stream);
HN_API int hn_op_combine(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L, int small_C,
                         int zw, int dh, int head_pitch, const float* Wv, const float* bv, void* O, int o_ld,
                         void* cuda_stream);
#ifdef __cplusplus
}
#endif
#endif /* HEALNET_B200_H_ */
