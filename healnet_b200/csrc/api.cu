// api.cu — the C ABI of libhealnet_b200.so (include/healnet_b200.h): handle management, weight
// registration / repacking, and the stream-ordered orchestration of one HealNet.forward
// (reference healnet/models/healnet.py:190-250) as a fixed sequence of the sm_100a kernels in
// rowops.cu / gemm.cu / xattn.cu. No host synchronisation, no allocation inside hn_forward.
#include <cmath>
#include <cstring>
#include <map>
#include <new>
#include <vector>

#include "../../include/healnet_b200.h"
#include "common.cuh"
#include "bwd.cuh"
#include "handle.cuh"
#include "pack.cuh"

namespace hn {
namespace {
thread_local std::string g_error;
}
void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

}  // namespace hn

using namespace hn;

namespace {

// sizes (or carves, when arena.base != null) the packed store; dedupes tied layers by pointer identity
int plan_packed(hn_handle* h, Arena& ar) {
  const hn_desc& d = h->d;
  const int M = h->M, D = d.l_d;
  std::map<SlotKey, AttnPacked> attn_seen;
  std::map<SlotKey, FFPacked> ff_seen;
  h->attn.assign(static_cast<size_t>(d.depth) * (M + 1), AttnPacked());
  h->ff.assign(static_cast<size_t>(d.depth) * (M + 1), FFPacked());
  for (int l = 0; l < d.depth; ++l) {
    for (int m = 0; m <= M; ++m) {
      const bool self = (m == M);
      if (self && d.self_per_cross_attn == 0) continue;
      const std::vector<const float*>& wa = h->w[slot_index(h, l, 2 * m)];
      const std::vector<const float*>& wf = h->w[slot_index(h, l, 2 * m + 1)];
      HN_REQUIRE(!wa.empty() && !wf.empty(), "hn_pack_weights: a layer slot has no registered weights");
      SlotKey ka{std::vector<const void*>(wa.begin(), wa.end())};
      ka.ptrs.push_back(reinterpret_cast<const void*>(static_cast<intptr_t>(self ? -1 : h->C[m])));
      auto ita = attn_seen.find(ka);
      if (ita != attn_seen.end()) {
        h->attn[l * (M + 1) + m] = ita->second;
      } else {
        AttnPacked p;
        if (self) {
          p.C = D;
          p.Wq = ar.take<__half>(static_cast<size_t>(3) * d.l_heads * h->hpl * 2 * h->segD);
          p.Wo = ar.take<__half>(static_cast<size_t>(D) * 2 * d.l_heads * h->hpl);
        } else {
          p.C = h->C[m];
          p.small = p.C <= 63;
          if (p.small) {
            p.zw = p.C <= 31 ? 32 : 64;
            p.WqS = ar.take<__half>(static_cast<size_t>(d.x_heads) * p.zw * 2 * h->segD);
            p.Wv = ar.take<float>(static_cast<size_t>(h->I) * p.zw);
            p.bv = ar.take<float>(h->I);
            p.WoS = ar.take<__half>(static_cast<size_t>(D) * 2 * seg_of(d.x_heads * p.zw));
            p.boS = ar.take<float>(D);
          }
          p.Wq = ar.take<__half>(static_cast<size_t>(d.x_heads) * h->hpx * 2 * h->segD);
          p.Wkv = ar.take<__half>(static_cast<size_t>(2) * d.x_heads * h->hpx * 2 * seg_of(p.C));
          p.bkv = ar.take<float>(static_cast<size_t>(2) * d.x_heads * h->hpx);
          p.Wo = ar.take<__half>(static_cast<size_t>(D) * 2 * d.x_heads * h->hpx);
        }
        attn_seen[ka] = p;
        h->attn[l * (M + 1) + m] = p;
      }
      SlotKey kf{std::vector<const void*>(wf.begin(), wf.end())};
      auto itf = ff_seen.find(kf);
      if (itf != ff_seen.end()) {
        h->ff[l * (M + 1) + m] = itf->second;
      } else {
        FFPacked p;
        p.W1 = ar.take<__half>(static_cast<size_t>(8) * D * 2 * h->segD);
        p.b1 = ar.take<float>(static_cast<size_t>(8) * D);
        p.W2 = ar.take<__half>(static_cast<size_t>(D) * 2 * h->seg4D);
        ff_seen[kf] = p;
        h->ff[l * (M + 1) + m] = p;
      }
    }
  }
  return 0;
}

}  // namespace

// Lays the forward workspace out over `base` (null: sizing pass). present[m] tells which modalities are given.
int hn::plan_workspace(const hn_handle* h, int batch, const int* axis_sizes, const bool* present, long mask_tokens,
                       char* base, Workspace& ws, const long* tok_begin, const long* tok_count) {
  const hn_desc& d = h->d;
  const int M = h->M, L = d.l_c, D = d.l_d;
  Arena ar;
  ar.base = base;
  const long rows = static_cast<long>(batch) * L;
  ws.x = ar.take<float>(rows * D);
  ws.xn = ar.take<__half>(rows * 2 * h->segD);
  int qw = d.self_per_cross_attn ? 3 * d.l_heads * h->hpl : 0;
  int ow = d.self_per_cross_attn ? d.l_heads * h->hpl : 0;
  size_t part_acc_elems = 0, part_ml_elems = 0, kv_elems = 0, mask_words = 0;
  const int n_ltiles = (L + 127) / 128;
  ws.self_precise = true;
  if (d.self_per_cross_attn) {
    ws.self_nsplit = attention_pick_nsplit(batch, L, d.l_heads, L);
    part_acc_elems = static_cast<size_t>(batch) * ws.self_nsplit * d.l_heads * n_ltiles * 128 * h->hpl;
    part_ml_elems = static_cast<size_t>(batch) * ws.self_nsplit * d.l_heads * n_ltiles * 128 * 2;
  }
  for (int m = 0; m < M; ++m) {
    ModPlan& mp = ws.mod[m];
    mp = ModPlan();
    mp.present = present ? present[m] : true;
    if (!mp.present) continue;
    mp.n_axes = d.num_spatial_axes[m];
    mp.c_raw = d.channel_dims[m];
    mp.C = h->C[m];
    mp.N = 1;
    for (int a = 0; a < mp.n_axes; ++a) {
      mp.axes[a] = axis_sizes[m * HN_MAX_AXES + a];
      HN_REQUIRE(mp.axes[a] >= 1, "hn_forward: axis sizes must be >= 1");
      mp.N *= mp.axes[a];
    }
    HN_REQUIRE(mp.N < (1L << 31), "hn_forward: token axis too long");
    mp.Nl = mp.N;
    if (tok_count != nullptr && tok_count[m] > 0 && tok_count[m] < mp.N) {
      mp.sharded = true;
      mp.Nl = tok_count[m];
      mp.tok0 = tok_begin != nullptr ? tok_begin[m] : 0;
      HN_REQUIRE(mp.tok0 >= 0 && mp.tok0 + mp.Nl <= mp.N, "hn_forward_split: token range outside the modality");
      HN_REQUIRE(mp.N > PRECISE_MAX_TOKENS, "hn_forward_split: short token axes are replicated, not sharded");
    }
    mp.small = mp.C <= 63 && mp.N > PRECISE_MAX_TOKENS;
    mp.masked = (mask_tokens > 0 && mask_tokens == mp.Nl);
    int axsum = 0;
    for (int a = 0; a < mp.n_axes; ++a) axsum += mp.axes[a];
    mp.tab = ar.take<float>(static_cast<size_t>(axsum) * (2 * d.num_freq_bands + 1));
    const int vd = mp.small ? (mp.C <= 31 ? 32 : 64) : h->hpx;
    if (mp.small) {
      mp.zw = vd;
      mp.z = ar.take<__half>(static_cast<size_t>(batch) * mp.Nl * 2 * mp.zw);  // rows [hi | lo]
      qw = qw > d.x_heads * mp.zw ? qw : d.x_heads * mp.zw;
    } else {
      mp.precise = true;
      mp.segC = seg_of(mp.C);
      mp.ldz = mp.precise ? 2 * mp.segC : ctx_ld(mp.C);
      mp.z = ar.take<__half>(static_cast<size_t>(batch) * mp.Nl * mp.ldz);
      const size_t kv = static_cast<size_t>(batch) * mp.Nl * 2 * d.x_heads * h->hpx * (mp.precise ? 2 : 1);
      kv_elems = kv > kv_elems ? kv : kv_elems;
      qw = qw > d.x_heads * h->hpx ? qw : d.x_heads * h->hpx;
    }
    ow = ow > d.x_heads * h->hpx ? ow : d.x_heads * h->hpx;
    mp.nsplit = mp.small ? small_attention_pick_nsplit(batch, L, d.x_heads, mp.Nl, vd)
                         : attention_pick_nsplit(batch, L, d.x_heads, mp.Nl);
    const size_t pa = static_cast<size_t>(batch) * mp.nsplit * d.x_heads * n_ltiles * 128 * vd;
    const size_t pm = static_cast<size_t>(batch) * mp.nsplit * d.x_heads * n_ltiles * 128 * 2;
    part_acc_elems = pa > part_acc_elems ? pa : part_acc_elems;
    part_ml_elems = pm > part_ml_elems ? pm : part_ml_elems;
    if (mp.masked) mask_words = static_cast<size_t>(batch) * ((mp.Nl + 63) / 64);
  }
  ws.q = ar.take<__half>(rows * 2 * (qw > 8 ? qw : 8));
  ws.o = ar.take<__half>(rows * 2 * (ow > 8 ? ow : 8));
  ws.hid = ar.take<__half>(rows * 2 * h->seg4D);
  ws.kv = ar.take<__half>(kv_elems);
  ws.part_acc = ar.take<float>(part_acc_elems);
  ws.part_ml = ar.take<float>(part_ml_elems);
  ws.mask_bits = ar.take<uint64_t>(mask_words);
  ws.pooled = ar.take<float>(static_cast<size_t>(batch) * D);
  ws.ln_counters = ar.take<unsigned>(static_cast<size_t>((rows + 127) / 128));
  ws.bytes = ar.off + 256;
  return 0;
}

namespace {

#define HN_TRY(expr)          \
  do {                        \
    int _rc = (expr);         \
    if (_rc != 0) return _rc; \
    ++h->launches;            \
  } while (0)

// measurement hook: one CUDA event pair per bracketed launch; kind 0 = streaming cross-attention kernel, 1 = K/V
// projection GEMM of the generic path, 2 = context-row build (Fourier tables + standardisation)
int profile_begin_raw(hn_handle* h, int kind, int modality, double flops_exec, double flops_useful, double exps,
                      cudaStream_t st) {
  if (!h->profile) return 0;
  const size_t slot = h->ev_mod.size();
  while (h->ev.size() < 2 * (slot + 1)) {
    cudaEvent_t e;
    HN_CHECK_CUDA(cudaEventCreate(&e));
    h->ev.push_back(e);
  }
  h->ev_mod.push_back(modality);
  h->ev_kind.push_back(kind);
  h->ev_flops.push_back(flops_exec);
  h->ev_useful.push_back(flops_useful);
  h->ev_exps.push_back(exps);
  HN_CHECK_CUDA(cudaEventRecord(h->ev[2 * slot], st));
  return 0;
}
// cross-attention launch: executed FLOPs count the padded tiles (rows to 128, tokens to 64, operand width kd);
// useful FLOPs are the unpadded contraction (context width C on the reassociated small-context path, dim_head on the
// generic one), SURVEY.md section 8d
int profile_begin(hn_handle* h, int modality, const AttnArgs& a, int useful_width, cudaStream_t st) {
  if (!h->profile) return 0;
  const double rows = static_cast<double>((a.L + 127) / 128) * 128.0, toks = static_cast<double>((a.N + 63) / 64) * 64.0;
  const double kd = a.shared_kv ? a.kd : a.hp;
  // products per tile: S + PV; precise generic: 3 + 2, split small-context: 3 + 1
  const double passes = a.precise ? (a.shared_kv ? 4.0 : 5.0) : 2.0;
  return profile_begin_raw(h, 0, modality, static_cast<double>(a.batch) * a.H * rows * toks * 2.0 * kd * passes,
                           static_cast<double>(a.batch) * a.H * a.L * static_cast<double>(a.N) * 4.0 * useful_width,
                           static_cast<double>(a.batch) * a.H * rows * toks, st);
}
void profile_end(hn_handle* h, cudaStream_t st) {
  if (!h->profile) return;
  cudaEventRecord(h->ev[2 * (h->ev_mod.size() - 1) + 1], st);
}

// x += FeedForward(LN(x))   (healnet.py:237/245, 339-351)
// The PreNorm LayerNorms of a forward in execution order. Every one but the first follows a residual GEMM, which can
// emit it from its epilogue (GemmArgs::ln_*): `valid` says ws.xn already holds the LayerNorm that is due next.
struct LnPlan {
  std::vector<std::pair<const float*, const float*>> seq;  // (gamma, beta)
  size_t next = 0;
  bool valid = false;
  int fused = 0;  // fused launches so far (the row-block counters grow by tiles_n with each)
};
// LayerNorm due now -> ws.xn (skipped when the previous residual GEMM produced it)
int ln_step(hn_handle* h, LnPlan& lp, Workspace& ws, long rows, cudaStream_t st) {
  const int D = h->d.l_d, sD = h->segD;
  HN_REQUIRE(lp.next < lp.seq.size(), "hn_forward: LayerNorm plan out of sync");
  if (!lp.valid)
    HN_TRY(launch_layernorm_f16(ws.x, D, lp.seq[lp.next].first, lp.seq[lp.next].second, ws.xn, 2 * sD, sD, sD, rows, D, st));
  lp.valid = false;
  ++lp.next;
  return 0;
}
// residual GEMM that also emits the next LayerNorm when the kernel can
int residual_gemm(hn_handle* h, GemmArgs g, LnPlan& lp, Workspace& ws, cudaStream_t st) {
  bool fuse = lp.next < lp.seq.size() && h->segD == h->d.l_d && gemm_can_fuse_ln(g);
  if (fuse) {
    g.ln_gamma = lp.seq[lp.next].first;
    g.ln_beta = lp.seq[lp.next].second;
    g.ln_out = ws.xn;
    g.ln_ld = 2 * h->segD;
    g.ln_seg = h->segD;
    g.ln_counters = ws.ln_counters;
    g.ln_epoch = lp.fused + 1;
    fuse = (reinterpret_cast<uintptr_t>(g.ln_gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.ln_beta) & 15) == 0;
    if (!fuse) g.ln_out = nullptr;
  }
  HN_TRY(launch_gemm(g, st));
  lp.valid = fuse;
  if (fuse) ++lp.fused;
  return 0;
}

// training-mode forward: copy a buffer onto the tape (stream-ordered device-to-device copy)
int tape_put(hn_handle* h, size_t off, const void* src, size_t bytes, cudaStream_t st) {
  if (h->tape == nullptr) return 0;
  HN_CHECK_CUDA(cudaMemcpyAsync(h->tape + off, src, bytes, cudaMemcpyDeviceToDevice, st));
  return 0;
}

int run_ff(hn_handle* h, const std::vector<const float*>& wf, const FFPacked& fp, Workspace& ws, long rows,
           LnPlan& lp, const BlockRec* rec, cudaStream_t st) {
  const hn_desc& d = h->d;
  const int D = d.l_d;
  const int sD = h->segD, s4 = h->seg4D;
  int rc = 0;
  if (rec != nullptr) rc = tape_put(h, rec->x_in, ws.x, sizeof(float) * rows * D, st);
  if (rc != 0) return rc;
  rc = ln_step(h, lp, ws, rows, st);
  if (rc != 0) return rc;
  if (rec != nullptr) rc = tape_put(h, rec->xn, ws.xn, sizeof(__half) * rows * 2 * sD, st);
  if (rc != 0) return rc;
  GemmArgs g1{ws.xn, fp.W1, static_cast<int>(rows), 8 * D, D, 2 * sD, 2 * sD, EPI_GATE_F16,
              d.snn ? ACT_SELU : ACT_GELU, fp.b1, ws.hid, 2 * s4, 3, sD, sD, s4};
  HN_TRY(launch_gemm(g1, st));
  if (rec != nullptr) rc = tape_put(h, rec->o, ws.hid, sizeof(__half) * rows * 2 * s4, st);
  if (rc != 0) return rc;
  GemmArgs g2{ws.hid, fp.W2, static_cast<int>(rows), D, 4 * D, 2 * s4, 2 * s4, EPI_RES, 0, wf[5], ws.x, D, 3, s4, s4, 0};
  return residual_gemm(h, g2, lp, ws, st);
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* hn_last_error(void) { return hn::get_error(); }

int hn_create(const hn_desc* desc, hn_handle** out) {
  HN_REQUIRE(desc != nullptr && out != nullptr, "hn_create: null argument");
  const hn_desc& d = *desc;
  HN_REQUIRE(d.n_modalities >= 1 && d.n_modalities <= HN_MAX_MODALITIES, "hn_create: 1..16 modalities supported");
  HN_REQUIRE(d.depth >= 1, "hn_create: depth must be >= 1");
  HN_REQUIRE(d.l_c >= 1 && d.l_d >= 1, "hn_create: latent array must be non-empty");
  HN_REQUIRE(d.x_heads >= 1 && d.l_heads >= 1, "hn_create: head counts must be >= 1");
  HN_REQUIRE(d.cross_dim_head >= 1 && d.cross_dim_head <= MAX_DIM_HEAD, "hn_create: cross_dim_head must be in 1..128");
  HN_REQUIRE(d.self_per_cross_attn == 0 || d.self_per_cross_attn == 1,
             "hn_create: self_per_cross_attn must be 0 or 1 (the reference fails for >= 2, healnet.py:242)");
  // with self_per_cross_attn == 0 no latent attention module exists (healnet.py:166-168): its head size is unused
  HN_REQUIRE(d.latent_dim_head >= 1 && (d.self_per_cross_attn == 0 || d.latent_dim_head <= MAX_DIM_HEAD),
             "hn_create: latent_dim_head must be in 1..128");
  HN_REQUIRE(d.num_freq_bands >= 1 || !d.fourier_encode_data, "hn_create: num_freq_bands must be >= 1");
  HN_REQUIRE(!d.final_classifier_head || d.out_dims >= 1, "hn_create: out_dims must be >= 1");
  hn_handle* h = new (std::nothrow) hn_handle();
  HN_REQUIRE(h != nullptr, "hn_create: out of host memory");
  h->d = d;
  h->M = d.n_modalities;
  h->I = d.x_heads * d.cross_dim_head;
  h->lI = d.l_heads * d.latent_dim_head;
  h->hpx = head_pitch(d.cross_dim_head);
  h->hpl = head_pitch(d.latent_dim_head);
  h->segD = round_up(d.l_d, 64);
  h->seg4D = round_up(4 * d.l_d, 64);
  for (int m = 0; m < h->M; ++m) {
    if (!(d.num_spatial_axes[m] >= 1 && d.num_spatial_axes[m] <= HN_MAX_AXES && d.channel_dims[m] >= 1)) {
      delete h;
      set_error("hn_create: each modality needs 1..4 spatial axes and >= 1 channel");
      return -1;
    }
    h->C[m] = d.channel_dims[m] + (d.fourier_encode_data ? d.num_spatial_axes[m] * (2 * d.num_freq_bands + 1) : 0);
  }
  h->slots_per_layer = 2 * h->M + 2;
  h->w.assign(static_cast<size_t>(d.depth + 1) * h->slots_per_layer, std::vector<const float*>());
  h->export_ptrs.assign(static_cast<size_t>(d.depth) * (h->M + 1), nullptr);
  *out = h;
  return 0;
}

int hn_destroy(hn_handle* h) {
  if (h == nullptr) return 0;
  if (h->packed != nullptr) cudaFree(h->packed);
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  delete h;
  return 0;
}

int hn_set_weights(hn_handle* h, int layer, int slot, const void* const* dev_ptrs, int n) {
  HN_REQUIRE(h != nullptr && dev_ptrs != nullptr, "hn_set_weights: null argument");
  HN_REQUIRE(layer >= -1 && layer < h->d.depth, "hn_set_weights: layer out of range");
  int expect;
  if (layer == -1) {
    HN_REQUIRE(slot == 0 || slot == 1, "hn_set_weights: layer -1 has slots 0 (latents) and 1 (to_logits)");
    expect = slot == 0 ? 1 : 4;
  } else {
    HN_REQUIRE(slot >= 0 && slot < 2 * h->M + 2, "hn_set_weights: slot out of range");
    expect = (slot % 2 == 1) ? 6 : (slot < 2 * h->M ? 8 : 6);
  }
  HN_REQUIRE(n == expect, "hn_set_weights: wrong number of tensors for this slot");
  std::vector<const float*> v(n);
  for (int i = 0; i < n; ++i) {
    HN_REQUIRE(dev_ptrs[i] != nullptr, "hn_set_weights: null tensor pointer");
    v[i] = static_cast<const float*>(dev_ptrs[i]);
  }
  std::vector<const float*>& dst = h->w[slot_index(h, layer, slot)];
  if (dst != v) {
    dst = v;
    h->packed_valid = false;
  }
  return 0;
}

int hn_pack_weights(hn_handle* h, void* cuda_stream) {
  HN_REQUIRE(h != nullptr, "hn_pack_weights: null handle");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const hn_desc& d = h->d;
  const int M = h->M, D = d.l_d, sD = h->segD;
  Arena sizing;
  int rc = plan_packed(h, sizing);
  if (rc != 0) return rc;
  const size_t need = sizing.off + 256;
  if (h->packed == nullptr || h->packed_bytes < need) {
    if (h->packed != nullptr) {
      HN_CHECK_CUDA(cudaStreamSynchronize(st));
      cudaFree(h->packed);
      h->packed = nullptr;
    }
    HN_CHECK_CUDA(cudaMalloc(&h->packed, need));
    h->packed_bytes = need;
  }
  Arena ar;
  ar.base = static_cast<char*>(h->packed);
  rc = plan_packed(h, ar);
  if (rc != 0) return rc;
  std::map<const void*, bool> done;
  for (int l = 0; l < d.depth; ++l) {
    for (int m = 0; m <= M; ++m) {
      const bool self = (m == M);
      if (self && d.self_per_cross_attn == 0) continue;
      const std::vector<const float*>& wa = h->w[slot_index(h, l, 2 * m)];
      const std::vector<const float*>& wf = h->w[slot_index(h, l, 2 * m + 1)];
      const AttnPacked& ap = h->attn[l * (M + 1) + m];
      const FFPacked& fp = h->ff[l * (M + 1) + m];
      if (!done[ap.Wq]) {
        done[ap.Wq] = true;
        if (self) {
          // {norm.w, norm.b, to_q.w [lI][D], to_kv.w [2lI][D], to_out.w [D][lI], to_out.b}
          const int lh = d.l_heads, ldh = d.latent_dim_head;
          const float scale = 2.f / std::sqrt(static_cast<float>(ldh)) * LOG2E;
          const int hp = h->hpl;
          rc = pack_headpad_rows(ap.Wq, 2 * sD, 0, wa[2], D, 0, lh, ldh, D, scale, nullptr, sD, sD, hp, st);
          if (rc == 0)
            rc = pack_headpad_rows(ap.Wq, 2 * sD, lh * hp, wa[3], D, 0, lh, ldh, D, 1.f, nullptr, sD, sD, hp, st);
          if (rc == 0)
            rc = pack_headpad_rows(ap.Wq, 2 * sD, 2 * lh * hp, wa[3], D, h->lI, lh, ldh, D, 1.f, nullptr, sD, sD, hp, st);
          if (rc == 0) rc = pack_headpad_cols(ap.Wo, 2 * lh * hp, wa[4], h->lI, D, lh, ldh, lh * hp, lh * hp, hp, st);
        } else {
          // {norm.w, norm.b, norm_context.w, norm_context.b, to_q.w [I][D], to_kv.w [2I][C], to_out.w [D][I], to_out.b}
          const int H = d.x_heads, dh = d.cross_dim_head, C = ap.C;
          const float scale = 2.f / std::sqrt(static_cast<float>(dh)) * LOG2E;
          if (ap.small) {
            rc = pack_smallc_q(ap.WqS, 2 * sD, wa[4], wa[5], wa[2], H, D, C, dh, scale, ap.zw, sD, sD, st);
            if (rc == 0) rc = pack_smallc_v(ap.Wv, ap.bv, wa[5], wa[2], wa[3], h->I, C, ap.zw, st);
            if (rc == 0) {
              const int sHZ = seg_of(H * ap.zw);
              rc = pack_smallc_out(ap.WoS, 2 * sHZ, ap.boS, wa[6], wa[7], ap.Wv, ap.bv, D, h->I, H, dh, ap.zw, sHZ, sHZ, st);
            }
          }
          if (rc == 0) {
            const int sC = seg_of(C);
            const int hp = h->hpx;
            rc = pack_headpad_rows(ap.Wq, 2 * sD, 0, wa[4], D, 0, H, dh, D, scale, nullptr, sD, sD, hp, st);
            if (rc == 0) rc = pack_headpad_rows(ap.Wkv, 2 * sC, 0, wa[5], C, 0, H, dh, C, 1.f, wa[2], sC, sC, hp, st);
            if (rc == 0)
              rc = pack_headpad_rows(ap.Wkv, 2 * sC, H * hp, wa[5], C, h->I, H, dh, C, 1.f, wa[2], sC, sC, hp, st);
            if (rc == 0) HN_CHECK_CUDA(cudaMemsetAsync(ap.bkv, 0, sizeof(float) * H * hp, st));
            if (rc == 0) rc = fold_beta_headpad(ap.bkv, H * hp, wa[5], C, h->I, H, dh, C, wa[3], hp, st);
          }
          if (rc == 0)
            rc = pack_headpad_cols(ap.Wo, 2 * H * h->hpx, wa[6], h->I, D, H, dh, H * h->hpx, H * h->hpx, h->hpx, st);
        }
        if (rc != 0) return rc;
      }
      if (!done[fp.W1]) {
        done[fp.W1] = true;
        // {norm.w, norm.b, net.0.w [8D][D], net.0.b [8D], net.2.w [D][4D], net.2.b [D]}
        rc = pack_ff1(fp.W1, 2 * sD, fp.b1, wf[2], wf[3], D, 4 * D, sD, sD, st);
        if (rc == 0) rc = pack_plain(fp.W2, 2 * h->seg4D, wf[4], 4 * D, D, 4 * D, h->seg4D, h->seg4D, st);
        if (rc != 0) return rc;
      }
    }
  }
  h->packed_valid = true;
  return 0;
}

size_t hn_workspace_bytes(const hn_handle* h, int batch, const int* axis_sizes) {
  if (h == nullptr || batch < 1 || axis_sizes == nullptr) {
    set_error("hn_workspace_bytes: bad argument");
    return 0;
  }
  Workspace ws;
  // size for the worst case: every modality present and masked
  long mask_tokens = 0;
  for (int m = 0; m < h->M; ++m) {
    long n = 1;
    for (int a = 0; a < h->d.num_spatial_axes[m]; ++a) n *= axis_sizes[m * HN_MAX_AXES + a];
    mask_tokens = n > mask_tokens ? n : mask_tokens;
  }
  if (plan_workspace(h, batch, axis_sizes, nullptr, 0, nullptr, ws) != 0) return 0;
  return ws.bytes + static_cast<size_t>(batch) * ((mask_tokens + 63) / 64) * 8 + 256;
}

size_t hn_workspace_bytes_split(const hn_handle* h, int batch, const int* axis_sizes, const long* tok_count) {
  if (h == nullptr || batch < 1 || axis_sizes == nullptr || tok_count == nullptr) {
    set_error("hn_workspace_bytes_split: bad argument");
    return 0;
  }
  Workspace ws;
  long mask_tokens = 0;
  for (int m = 0; m < h->M; ++m) {
    long n = 1;
    for (int a = 0; a < h->d.num_spatial_axes[m]; ++a) n *= axis_sizes[m * HN_MAX_AXES + a];
    if (tok_count[m] > 0 && tok_count[m] < n) n = tok_count[m];
    mask_tokens = n > mask_tokens ? n : mask_tokens;
  }
  if (plan_workspace(h, batch, axis_sizes, nullptr, 0, nullptr, ws, nullptr, tok_count) != 0) return 0;
  return ws.bytes + static_cast<size_t>(batch) * ((mask_tokens + 63) / 64) * 8 + 256;
}

int hn_last_launch_count(const hn_handle* h) { return h ? h->launches : 0; }

int hn_set_attention_export(hn_handle* h, int layer, int module, float* dev_out) {
  HN_REQUIRE(h != nullptr, "hn_set_attention_export: null handle");
  HN_REQUIRE(layer >= 0 && layer < h->d.depth && module >= 0 && module <= h->M, "hn_set_attention_export: bad index");
  HN_REQUIRE(module < h->M || h->d.self_per_cross_attn, "hn_set_attention_export: model has no latent self-attention");
  h->export_ptrs[layer * (h->M + 1) + module] = dev_out;
  return 0;
}

int hn_set_io_dtype(hn_handle* h, int dtype) {
  HN_REQUIRE(h != nullptr && dtype >= 0 && dtype <= 2, "hn_set_io_dtype: dtype must be 0 (fp32), 1 (bf16) or 2 (fp16)");
  h->io_dtype = dtype;
  return 0;
}

int hn_profile_enable(hn_handle* h, int on) {
  HN_REQUIRE(h != nullptr, "hn_profile_enable: null handle");
  h->profile = on != 0;
  return 0;
}

int hn_profile_read(hn_handle* h, int kind, int modality, float* ms, int* launches, double* flops,
                    double* flops_useful, double* exps) {
  HN_REQUIRE(h != nullptr && ms && launches && flops && flops_useful && exps, "hn_profile_read: null argument");
  *ms = 0.f;
  *launches = 0;
  *flops = 0.0;
  *flops_useful = 0.0;
  *exps = 0.0;
  for (size_t i = 0; i < h->ev_mod.size(); ++i) {
    if (h->ev_mod[i] != modality || h->ev_kind[i] != kind) continue;
    float t = 0.f;
    HN_CHECK_CUDA(cudaEventElapsedTime(&t, h->ev[2 * i], h->ev[2 * i + 1]));
    *ms += t;
    *launches += 1;
    *flops += h->ev_flops[i];
    *flops_useful += h->ev_useful[i];
    *exps += h->ev_exps[i];
  }
  return 0;
}

int hn_forward(hn_handle* h, int batch, const void* const* modality_ptrs, const int* axis_sizes,
               const int* skip_latent_block, const uint8_t* mask, long mask_tokens, float* latents_out,
               float* logits_out, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  return hn_forward_ex(h, batch, modality_ptrs, nullptr, axis_sizes, skip_latent_block, mask, mask_tokens, latents_out,
                       logits_out, workspace, workspace_bytes, cuda_stream);
}

int hn_forward_ex(hn_handle* h, int batch, const void* const* modality_ptrs, void* const* modality_ready_events,
                  const int* axis_sizes, const int* skip_latent_block, const uint8_t* mask, long mask_tokens,
                  float* latents_out, float* logits_out, void* workspace, size_t workspace_bytes,
                  void* cuda_stream) {
  return forward_impl(h, batch, modality_ptrs, modality_ready_events, axis_sizes, nullptr, nullptr, skip_latent_block,
                      mask, mask_tokens, latents_out, logits_out, workspace, workspace_bytes, cuda_stream);
}

int hn_forward_split(hn_handle* h, int batch, const void* const* modality_ptrs, void* const* modality_ready_events,
                     const int* axis_sizes, const long* tok_begin, const long* tok_count,
                     const int* skip_latent_block, const uint8_t* mask, long mask_tokens, float* latents_out,
                     float* logits_out, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  HN_REQUIRE(tok_begin != nullptr && tok_count != nullptr, "hn_forward_split: token ranges required");
  return forward_impl(h, batch, modality_ptrs, modality_ready_events, axis_sizes, tok_begin, tok_count,
                      skip_latent_block, mask, mask_tokens, latents_out, logits_out, workspace, workspace_bytes,
                      cuda_stream);
}

// bytes of one exchange slot (un-normalised accumulators + (max, sum) per (sample, head, latent row)), 256-aligned
static size_t xchg_slot_bytes(const hn_handle* h, int batch) {
  const size_t rows = static_cast<size_t>(batch) * h->d.x_heads * h->d.l_c;
  const size_t w = h->hpx > 64 ? 128 : 64;  // widest accumulator row of any path (small-context rows are <= 64)
  return ((rows * (w + 2) * sizeof(float)) + 255) & ~size_t(255);
}

size_t hn_exchange_bytes(const hn_handle* h, int batch) {
  if (h == nullptr || batch < 1) return 0;
  return sizeof(XchgHeader) + 2 * xchg_slot_bytes(h, batch);
}

int hn_exchange_alloc(size_t bytes, void** dev_ptr, unsigned char* ipc_handle_out) {
  HN_REQUIRE(dev_ptr != nullptr && ipc_handle_out != nullptr && bytes >= sizeof(XchgHeader), "hn_exchange_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  void* p = nullptr;
  HN_CHECK_CUDA(cudaMalloc(&p, bytes));
  HN_CHECK_CUDA(cudaMemset(p, 0, bytes));
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
    return -10;
  }
  memcpy(ipc_handle_out, &hd, 64);
  *dev_ptr = p;
  return 0;
}

int hn_exchange_open(const unsigned char* ipc_handle, void** peer_ptr) {
  HN_REQUIRE(ipc_handle != nullptr && peer_ptr != nullptr, "hn_exchange_open: null argument");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, ipc_handle, 64);
  HN_CHECK_CUDA(cudaIpcOpenMemHandle(peer_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int hn_exchange_close(void* peer_ptr) {
  if (peer_ptr != nullptr) HN_CHECK_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return 0;
}

int hn_exchange_free(void* dev_ptr) {
  if (dev_ptr != nullptr) HN_CHECK_CUDA(cudaFree(dev_ptr));
  return 0;
}

int hn_set_exchange(hn_handle* h, int rank, int world, void* const* bufs, size_t bytes) {
  HN_REQUIRE(h != nullptr, "hn_set_exchange: null handle");
  if (world <= 1 || bufs == nullptr) {
    h->x_world = 0;
    return 0;
  }
  HN_REQUIRE(world <= HN_MAX_PEERS && rank >= 0 && rank < world, "hn_set_exchange: at most 8 ranks");
  for (int r = 0; r < world; ++r) {
    HN_REQUIRE(bufs[r] != nullptr && (reinterpret_cast<uintptr_t>(bufs[r]) & 255) == 0,
               "hn_set_exchange: exchange buffers must be 256-byte aligned device pointers");
    h->x_bufs[r] = static_cast<char*>(bufs[r]);
  }
  h->x_rank = rank;
  h->x_world = world;
  h->x_bytes = bytes;
  h->x_seq = 0;
  // sequence numbers restart at 0: stale flags / a stale error word in a re-registered buffer must not satisfy (or
  // fail) the first waits. Every rank clears its OWN header; the caller's barrier orders this before any publish.
  HN_CHECK_CUDA(cudaMemset(h->x_bufs[rank], 0, sizeof(XchgHeader)));
  HN_CHECK_CUDA(cudaDeviceSynchronize());
  return 0;
}

int hn_set_exchange_timeout(hn_handle* h, double seconds) {
  HN_REQUIRE(h != nullptr && seconds > 0.0, "hn_set_exchange_timeout: bad argument");
  h->x_timeout_clk = static_cast<long long>(seconds * 1.9e9);
  return 0;
}

int hn_exchange_error_async(const hn_handle* h, int* pinned_host_out, void* cuda_stream) {
  HN_REQUIRE(h != nullptr && pinned_host_out != nullptr && h->x_world > 1, "hn_exchange_error_async: no exchange registered");
  const XchgHeader* hdr = reinterpret_cast<const XchgHeader*>(h->x_bufs[h->x_rank]);
  HN_CHECK_CUDA(cudaMemcpyAsync(pinned_host_out, &hdr->error, sizeof(int), cudaMemcpyDeviceToHost,
                                static_cast<cudaStream_t>(cuda_stream)));
  return 0;
}

int hn_exchange_error(const hn_handle* h, int* error_out) {
  HN_REQUIRE(h != nullptr && error_out != nullptr && h->x_world > 1, "hn_exchange_error: no exchange registered");
  const XchgHeader* hdr = reinterpret_cast<const XchgHeader*>(h->x_bufs[h->x_rank]);
  HN_CHECK_CUDA(cudaMemcpy(error_out, &hdr->error, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

// Publishes this rank's partials of one cross-attention pass and fills the peer table the combine kernel reads.
static int exchange_partials(hn_handle* h, const Workspace& ws, int batch, int nsplit, int H, int L, int w,
                             PeerParts& pp, cudaStream_t st) {
  HN_REQUIRE(h->x_world > 1, "hn_forward_split: hn_set_exchange has not been called");
  const size_t slot = xchg_slot_bytes(h, batch);
  HN_REQUIRE(sizeof(XchgHeader) + 2 * slot <= h->x_bytes, "hn_forward_split: exchange buffer too small (hn_exchange_bytes)");
  const unsigned long long seq = ++h->x_seq;
  const size_t off = sizeof(XchgHeader) + (seq & 1) * slot;
  const size_t rows = static_cast<size_t>(batch) * H * L;
  pp.world = h->x_world;
  pp.rank = h->x_rank;
  pp.seq = seq;
  pp.timeout_clk = h->x_timeout_clk;
  for (int r = 0; r < h->x_world; ++r) {
    pp.hdr[r] = reinterpret_cast<XchgHeader*>(h->x_bufs[r]);
    pp.acc[r] = reinterpret_cast<const float*>(h->x_bufs[r] + off);
    pp.ml[r] = pp.acc[r] + rows * w;
  }
  float* my_acc = reinterpret_cast<float*>(h->x_bufs[h->x_rank] + off);
  return launch_merge_signal(ws.part_acc, ws.part_ml, batch, nsplit, H, L, w, my_acc, my_acc + rows * w, pp, st);
}

}  // extern "C"

int hn::forward_impl(hn_handle* h, int batch, const void* const* modality_ptrs, void* const* modality_ready_events,
                     const int* axis_sizes, const long* tok_begin, const long* tok_count,
                     const int* skip_latent_block, const uint8_t* mask, long mask_tokens, float* latents_out,
                     float* logits_out, void* workspace, size_t workspace_bytes, void* cuda_stream) {
  HN_REQUIRE(h != nullptr && modality_ptrs != nullptr && axis_sizes != nullptr, "hn_forward: null argument");
  HN_REQUIRE(batch >= 1, "hn_forward: batch must be >= 1");
  HN_REQUIRE(h->packed_valid, "hn_forward: call hn_pack_weights after registering / changing weights");
  HN_REQUIRE(workspace != nullptr && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
             "hn_forward: workspace must be a 256-byte aligned device buffer");
  HN_REQUIRE(latents_out != nullptr || logits_out != nullptr, "hn_forward: no output requested");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const hn_desc& d = h->d;
  const int M = h->M, L = d.l_c, D = d.l_d;
  const long rows = static_cast<long>(batch) * L;
  HN_REQUIRE(rows < (1L << 31), "hn_forward: batch * l_c too large");
  if (logits_out != nullptr) HN_REQUIRE(d.final_classifier_head, "hn_forward: model has no classifier head");
  const std::vector<const float*>& wl = h->w[slot_index(h, -1, 0)];
  HN_REQUIRE(wl.size() == 1, "hn_forward: latents not registered");
  if (logits_out != nullptr) HN_REQUIRE(h->w[slot_index(h, -1, 1)].size() == 4, "hn_forward: to_logits not registered");

  bool present[HN_MAX_MODALITIES];
  for (int m = 0; m < M; ++m) present[m] = modality_ptrs[m] != nullptr;
  if (mask == nullptr) mask_tokens = 0;
  Workspace ws;
  int rc = plan_workspace(h, batch, axis_sizes, present, mask_tokens, static_cast<char*>(workspace), ws, tok_begin,
                          tok_count);
  if (rc != 0) return rc;
  HN_REQUIRE(ws.bytes <= workspace_bytes, "hn_forward: workspace too small (see hn_workspace_bytes)");
  h->launches = 0;
  h->ev_mod.clear();
  h->ev_kind.clear();
  h->ev_flops.clear();
  h->ev_useful.clear();
  h->ev_exps.clear();

  // ---- once per forward: positional tables + standardised context rows (shared by all layers)
  const int sD = h->segD;
  if (h->seg4D != 4 * D)  // pad columns of the split hidden rows are never written by the gate epilogue
    HN_CHECK_CUDA(cudaMemsetAsync(ws.hid, 0, sizeof(__half) * rows * 2 * h->seg4D, st));
  // The standardised context rows of a modality are built right before its first cross-attention (layer 0), after
  // waiting for the caller's "input ready" event if one was given — so the host-to-device copy of a large late
  // modality (the volume) overlaps the layer-0 work on the earlier ones.
  bool mask_packed = false;
  auto build_context = [&](int m) -> int {
    ModPlan& mp = ws.mod[m];
    const void* raw = modality_ptrs[m];
    if (modality_ready_events != nullptr && modality_ready_events[m] != nullptr)
      HN_CHECK_CUDA(cudaStreamWaitEvent(st, static_cast<cudaEvent_t>(modality_ready_events[m]), 0));
    int prc = profile_begin_raw(h, 2, m, 0.0, 0.0, 0.0, st);
    if (prc != 0) return prc;
    if (d.fourier_encode_data)
      HN_TRY(launch_axis_tables(mp.tab, mp.axes, mp.n_axes, d.num_freq_bands, d.max_freq, st));
    if (mp.small)
      HN_TRY(launch_build_z_small(raw, mp.z, mp.zw, batch, mp.Nl, mp.c_raw, mp.n_axes, mp.axes, d.num_freq_bands,
                                  mp.tab, d.fourier_encode_data, st, mp.tok0, 1, h->io_dtype));
    else
      HN_TRY(launch_build_z_large(raw, mp.z, mp.ldz, mp.precise ? mp.segC : 0, batch, mp.Nl, mp.c_raw, mp.n_axes,
                                  mp.axes, d.num_freq_bands, mp.tab, d.fourier_encode_data, st, mp.tok0, h->io_dtype));
    profile_end(h, st);
    if (mp.masked && !mask_packed) {
      HN_TRY(launch_pack_mask(mask, ws.mask_bits, batch, mp.Nl, st));
      mask_packed = true;
    }
    return 0;
  };
  // ---- x = repeat(latents, 'n d -> b n d')   (healnet.py:225)
  HN_TRY(launch_broadcast_rows(wl[0], ws.x, static_cast<long>(L) * D, batch, st));

  // LayerNorm plan: same control flow as the loop below
  LnPlan lp;
  for (int l = 0; l < d.depth; ++l) {
    for (int m = 0; m < M; ++m) {
      if (ws.mod[m].present) {
        const std::vector<const float*>& wa = h->w[slot_index(h, l, 2 * m)];
        const std::vector<const float*>& wf = h->w[slot_index(h, l, 2 * m + 1)];
        lp.seq.emplace_back(wa[0], wa[1]);
        lp.seq.emplace_back(wf[0], wf[1]);
      }
      if (d.self_per_cross_attn && !(skip_latent_block != nullptr && skip_latent_block[m] != 0)) {
        const std::vector<const float*>& wa = h->w[slot_index(h, l, 2 * M)];
        const std::vector<const float*>& wf = h->w[slot_index(h, l, 2 * M + 1)];
        lp.seq.emplace_back(wa[0], wa[1]);
        lp.seq.emplace_back(wf[0], wf[1]);
      }
    }
  }
  HN_CHECK_CUDA(cudaMemsetAsync(ws.ln_counters, 0, sizeof(unsigned) * static_cast<size_t>((rows + 127) / 128), st));

  // training-mode forward (hn_forward_train): every block leaves what hn_backward needs on the tape
  const bool taping = h->tape != nullptr;
  size_t bi = 0;  // index of the next block record
  auto next_rec = [&]() -> const BlockRec* { return taping ? &h->train.blocks[bi++] : nullptr; };

  for (int l = 0; l < d.depth; ++l) {
    for (int m = 0; m < M; ++m) {
      ModPlan& mp = ws.mod[m];
      if (mp.present) {
        if (l == 0) {
          rc = build_context(m);
          if (rc != 0) return rc;
        }
        const BlockRec* rec = next_rec();
        if (rec != nullptr) {
          HN_REQUIRE(!mp.sharded, "hn_forward_train: token-sharded training is not supported");
          rc = tape_put(h, rec->x_in, ws.x, sizeof(float) * rows * D, st);
          if (rc != 0) return rc;
        }
        const std::vector<const float*>& wa = h->w[slot_index(h, l, 2 * m)];
        const std::vector<const float*>& wf = h->w[slot_index(h, l, 2 * m + 1)];
        const AttnPacked& ap = h->attn[l * (M + 1) + m];
        const FFPacked& fp = h->ff[l * (M + 1) + m];
        const int H = d.x_heads, HPx = h->hpx, ow = H * HPx;
        // PreNorm + to_q (split operands in, split Q / Q' out)
        rc = ln_step(h, lp, ws, rows, st);
        if (rc != 0) return rc;
        if (rec != nullptr) {
          rc = tape_put(h, rec->xn, ws.xn, sizeof(__half) * rows * 2 * sD, st);
          if (rc != 0) return rc;
        }
        const int qw = mp.small ? H * mp.zw : H * HPx;
        const bool q_split = true;
        GemmArgs gq{ws.xn, mp.small ? ap.WqS : ap.Wq, static_cast<int>(rows), qw, D, 2 * sD, 2 * sD, EPI_F16, 0,
                    nullptr, ws.q, q_split ? 2 * qw : qw, 3, sD, sD, q_split ? qw : 0};
        HN_TRY(launch_gemm(gq, st));
        AttnArgs aa{};
        aa.Q = ws.q;
        aa.q_ld = q_split ? 2 * qw : qw;
        aa.batch = batch;
        aa.L = L;
        aa.H = H;
        aa.N = mp.Nl;
        PeerParts pp;
        if (mp.sharded)
          HN_REQUIRE(h->export_ptrs[l * (M + 1) + m] == nullptr,
                     "hn_forward_split: attention-weight export is not available on a sharded token axis");
        aa.nsplit = mp.nsplit;
        aa.mask_bits = mp.masked ? ws.mask_bits : nullptr;
        aa.part_acc = ws.part_acc;
        aa.part_ml = ws.part_ml;
        if (mp.small) {
          aa.KV = mp.z;
          aa.kv_ld = 2 * mp.zw;
          aa.shared_kv = 1;
          aa.kd = mp.zw;
          aa.c_ones = mp.C;
          aa.precise = 1;
          aa.z_tail_merged = (mp.zw == 32 && mp.C >= 17 && mp.C <= 23) ? 1 : 0;   // as launch_build_z_small wrote them
#ifdef HN_DEBUG
          {  // timing experiment only: single-term scores (shows what the two extra score products cost)
            static const bool nosplit = getenv("HN_SMALL_NOSPLIT") != nullptr;
            if (nosplit) aa.precise = 0;
            static const bool nomerge = getenv("HN_SMALL_NOMERGE") != nullptr;   // A/B of the merged score tail
            if (nomerge) aa.z_tail_merged = 0;
          }
#endif
          aa.q_lo_off = qw;
          rc = profile_begin(h, m, aa, mp.C, st);
          if (rc != 0) return rc;
          HN_TRY(launch_attention(aa, st));
          profile_end(h, st);
          if (h->export_ptrs[l * (M + 1) + m] != nullptr) HN_TRY(launch_attn_export(aa, h->export_ptrs[l * (M + 1) + m], st));
          if (mp.sharded) HN_TRY(exchange_partials(h, ws, batch, mp.nsplit, H, L, mp.zw, pp, st));
          // merge + normalise only: u[b*L][h*zw + c] (split hi | lo); the V projection is folded into WoS
          const int sHZ = seg_of(H * mp.zw);
          if (sHZ != H * mp.zw)
            HN_CHECK_CUDA(cudaMemsetAsync(ws.o, 0, sizeof(__half) * rows * 2 * sHZ, st));
          if (rec != nullptr)  // merged row statistics: P'(t) = 2^(s_t - M + P_SHIFT), denominator column C
            HN_TRY(launch_row_stats(ws.part_acc, ws.part_ml, batch, mp.nsplit, H, L, mp.zw, mp.C, 0.0009765625f,
                                    reinterpret_cast<float*>(h->tape + rec->stats), st));
          HN_TRY(launch_combine_generic(ws.part_acc, ws.part_ml, batch, mp.nsplit, H, L, ws.o, 2 * sHZ, sHZ, mp.zw, st,
                                        mp.sharded ? &pp : nullptr, mp.C));
          if (rec != nullptr) {
            rc = tape_put(h, rec->o, ws.o, sizeof(__half) * rows * 2 * sHZ, st);
            if (rc != 0) return rc;
          }
        } else {
          // K/V projection of the standardised context (context LayerNorm affine folded into the weights).
          // Weights are always split (their rounding would not average out over tokens); z, K and V are split
          // only on short token axes.
          const long tok = static_cast<long>(batch) * mp.Nl;
          HN_REQUIRE(tok < (1L << 31), "hn_forward: batch * tokens too large for the K/V projection");
          const int kvw = 2 * H * HPx;
          GemmArgs gkv{mp.z, ap.Wkv, static_cast<int>(tok), kvw, mp.C, mp.ldz, 2 * mp.segC, EPI_F16, 0, ap.bkv,
                       ws.kv, mp.precise ? 2 * kvw : kvw, mp.precise ? 3 : 2, mp.segC, mp.segC,
                       mp.precise ? kvw : 0};
          // (Long token axis: the attention contracts P with the hi parts of V only, v_hi_only below. Forming the V half of
          // this projection from single fp16 operands as well — one product instead of three, no lo store — was built and
          // measured at +15 % on cfg 5, +7 % on cfg 2 / 4, and takes the latent error of the full-size PEAKED cfg 5 case
          // from 4.5e-4 to 7.4e-4, past the 5e-4 tests/test_gpu_fullsize.py allows: with one dominant token nothing
          // averages the 2^-12 operand rounding out. Parity first: removed again)
          rc = profile_begin_raw(h, 1, m, 2.0 * tok * kvw * round_up(mp.C, 64) * gkv.terms,
                                 2.0 * tok * 2.0 * h->I * mp.C, 0.0, st);
          if (rc != 0) return rc;
          HN_TRY(launch_gemm(gkv, st));
          profile_end(h, st);
          aa.KV = ws.kv;
          aa.kv_ld = mp.precise ? 2 * kvw : kvw;
          aa.k_col0 = 0;
          aa.v_col0 = H * HPx;
          aa.shared_kv = 0;
          aa.kd = 64;
          aa.hp = HPx;
          aa.precise = mp.precise ? 1 : 0;
          aa.q_lo_off = qw;
          aa.kv_lo_off = kvw;
          aa.v_hi_only = mp.N > 2048 ? 1 : 0;   // long axes: P.Vh only (common.cuh); decided on the FULL axis length
          const bool direct = mp.nsplit == 1 && !mp.sharded;  // one split: the kernel normalises and stores O itself
          if (direct) {
            aa.out = ws.o;
            aa.out_ld = 2 * ow;
            aa.out_lo_seg = ow;
          }
          rc = profile_begin(h, m, aa, d.cross_dim_head, st);
          if (rc != 0) return rc;
          HN_TRY(launch_attention(aa, st));
          profile_end(h, st);
          if (h->export_ptrs[l * (M + 1) + m] != nullptr) HN_TRY(launch_attn_export(aa, h->export_ptrs[l * (M + 1) + m], st));
          if (mp.sharded) HN_TRY(exchange_partials(h, ws, batch, mp.nsplit, H, L, HPx, pp, st));
          if (rec != nullptr)
            HN_TRY(launch_row_stats(ws.part_acc, ws.part_ml, batch, mp.nsplit, H, L, HPx, -1, 1.f,
                                    reinterpret_cast<float*>(h->tape + rec->stats), st));
          if (!direct)
            HN_TRY(launch_combine_generic(ws.part_acc, ws.part_ml, batch, mp.nsplit, H, L, ws.o, 2 * ow, ow, HPx, st,
                                          mp.sharded ? &pp : nullptr));
          if (rec != nullptr) {
            rc = tape_put(h, rec->o, ws.o, sizeof(__half) * rows * 2 * ow, st);
            if (rc != 0) return rc;
          }
        }
        // x = LeakyReLU(O Wo^T + bo) + x   (healnet.py:383-386, 426, 236)
        if (mp.small) {
          const int kz = H * mp.zw, sHZ = seg_of(kz);
          GemmArgs go{ws.o, ap.WoS, static_cast<int>(rows), D, kz, 2 * sHZ, 2 * sHZ, EPI_RES_LEAKY, 0, ap.boS, ws.x,
                      D, 3, sHZ, sHZ, 0};
          rc = residual_gemm(h, go, lp, ws, st);
        } else {
          GemmArgs go{ws.o, ap.Wo, static_cast<int>(rows), D, ow, 2 * ow, 2 * ow, EPI_RES_LEAKY, 0, wa[7], ws.x, D,
                      3, ow, ow, 0};
          rc = residual_gemm(h, go, lp, ws, st);
        }
        if (rc != 0) return rc;
        rc = run_ff(h, wf, fp, ws, rows, lp, next_rec(), st);
        if (rc != 0) return rc;
      }
      if (d.self_per_cross_attn && !(skip_latent_block != nullptr && skip_latent_block[m] != 0)) {
        // inside the modality loop (healnet.py:241-245)
        const std::vector<const float*>& wa = h->w[slot_index(h, l, 2 * M)];
        const std::vector<const float*>& wf = h->w[slot_index(h, l, 2 * M + 1)];
        const AttnPacked& ap = h->attn[l * (M + 1) + M];
        const FFPacked& fp = h->ff[l * (M + 1) + M];
        const int lh = d.l_heads, HPl = h->hpl, ow = lh * HPl, qw = 3 * lh * HPl;
        const bool prec = ws.self_precise;
        const BlockRec* rec = next_rec();
        if (rec != nullptr) {
          rc = tape_put(h, rec->x_in, ws.x, sizeof(float) * rows * D, st);
          if (rc != 0) return rc;
        }
        rc = ln_step(h, lp, ws, rows, st);
        if (rc != 0) return rc;
        if (rec != nullptr) {
          rc = tape_put(h, rec->xn, ws.xn, sizeof(__half) * rows * 2 * sD, st);
          if (rc != 0) return rc;
        }
        GemmArgs gq{ws.xn, ap.Wq, static_cast<int>(rows), qw, D, 2 * sD, 2 * sD, EPI_F16, 0, nullptr, ws.q,
                    prec ? 2 * qw : qw, 3, sD, sD, prec ? qw : 0};
        HN_TRY(launch_gemm(gq, st));
        AttnArgs aa{};
        aa.Q = ws.q;
        aa.q_ld = prec ? 2 * qw : qw;
        aa.KV = ws.q;
        aa.kv_ld = aa.q_ld;
        aa.k_col0 = lh * HPl;
        aa.v_col0 = 2 * lh * HPl;
        aa.shared_kv = 0;
        aa.kd = 64;
        aa.hp = HPl;
        aa.precise = prec ? 1 : 0;
        aa.q_lo_off = qw;
        aa.kv_lo_off = qw;
        aa.batch = batch;
        aa.L = L;
        aa.H = lh;
        aa.N = L;
        aa.nsplit = ws.self_nsplit;
        aa.mask_bits = nullptr;
        aa.part_acc = ws.part_acc;
        aa.part_ml = ws.part_ml;
        const bool direct = ws.self_nsplit == 1;
        if (direct) {
          aa.out = ws.o;
          aa.out_ld = 2 * ow;
          aa.out_lo_seg = ow;
        }
        HN_TRY(launch_attention(aa, st));
        if (h->export_ptrs[l * (M + 1) + M] != nullptr) HN_TRY(launch_attn_export(aa, h->export_ptrs[l * (M + 1) + M], st));
        if (rec != nullptr)
          HN_TRY(launch_row_stats(ws.part_acc, ws.part_ml, batch, ws.self_nsplit, lh, L, HPl, -1, 1.f,
                                  reinterpret_cast<float*>(h->tape + rec->stats), st));
        if (!direct)
          HN_TRY(launch_combine_generic(ws.part_acc, ws.part_ml, batch, ws.self_nsplit, lh, L, ws.o, 2 * ow, ow, HPl, st));
        if (rec != nullptr) {
          rc = tape_put(h, rec->o, ws.o, sizeof(__half) * rows * 2 * ow, st);
          if (rc != 0) return rc;
        }
        GemmArgs go{ws.o, ap.Wo, static_cast<int>(rows), D, ow, 2 * ow, 2 * ow, EPI_RES_LEAKY, 0, wa[5], ws.x, D,
                    3, ow, ow, 0};
        rc = residual_gemm(h, go, lp, ws, st);
        if (rc != 0) return rc;
        rc = run_ff(h, wf, fp, ws, rows, lp, next_rec(), st);
        if (rc != 0) return rc;
      }
    }
  }
  if (taping) {
    HN_REQUIRE(bi == h->train.blocks.size(), "hn_forward_train: tape plan out of sync with the forward");
    rc = tape_put(h, h->train.x_final, ws.x, sizeof(float) * rows * D, st);
    if (rc != 0) return rc;
  }
  bool any_sharded = false;
  for (int m = 0; m < M; ++m) any_sharded = any_sharded || (ws.mod[m].present && ws.mod[m].sharded);
  const XchgHeader* my_hdr = any_sharded ? reinterpret_cast<const XchgHeader*>(h->x_bufs[h->x_rank]) : nullptr;
  if (latents_out != nullptr) {
    HN_CHECK_CUDA(cudaMemcpyAsync(latents_out, ws.x, sizeof(float) * rows * D, cudaMemcpyDeviceToDevice, st));
    if (my_hdr != nullptr) HN_TRY(launch_poison_on_error(latents_out, rows * D, my_hdr, st));
  }
  if (logits_out != nullptr) {
    const std::vector<const float*>& wh = h->w[slot_index(h, -1, 1)];
    HN_TRY(launch_head(ws.x, batch, L, D, wh[0], wh[1], wh[2], wh[3], d.out_dims, ws.pooled, logits_out, st));
    ++h->launches;
    if (my_hdr != nullptr) HN_TRY(launch_poison_on_error(logits_out, static_cast<long>(batch) * d.out_dims, my_hdr, st));
  }
  return 0;
}

extern "C" {

// ------------------------------------------------------------------------------------ stand-alone Attention
}  // extern "C"
namespace {
struct AttnWs {
  __half *xh, *ch, *wq, *wkv, *wo, *q, *kv, *o;
  float *part_acc, *part_ml;
  uint64_t* mask_bits;
  int nsplit, sq, sc;
  bool prec;
  size_t bytes;
};
void plan_attn_ws(int batch, int n_q, long n_ctx, int qd, int cd, int heads, int dim_head, bool self, char* base,
                  AttnWs& w) {
  Arena ar;
  ar.base = base;
  w.sq = seg_of(qd);
  w.sc = seg_of(cd);
  w.prec = true;
  const long rows = static_cast<long>(batch) * n_q, toks = static_cast<long>(batch) * n_ctx;
  const int hw = heads * head_pitch(dim_head);
  w.xh = ar.take<__half>(rows * 2 * w.sq);
  w.ch = self ? w.xh : ar.take<__half>(toks * (w.prec ? 2 * w.sc : ctx_ld(cd)));
  w.wq = ar.take<__half>(static_cast<size_t>(hw) * 2 * w.sq);
  w.wkv = ar.take<__half>(static_cast<size_t>(2) * hw * 2 * w.sc);
  w.wo = ar.take<__half>(static_cast<size_t>(qd) * 2 * hw);
  w.q = ar.take<__half>(rows * 2 * hw);
  w.kv = ar.take<__half>(toks * 2 * hw * (w.prec ? 2 : 1));
  w.o = ar.take<__half>(rows * 2 * hw);
  w.nsplit = attention_pick_nsplit(batch, n_q, heads, n_ctx);
  const int n_ltiles = (n_q + 127) / 128;
  w.part_acc = ar.take<float>(static_cast<size_t>(batch) * w.nsplit * heads * n_ltiles * 128 * head_pitch(dim_head));
  w.part_ml = ar.take<float>(static_cast<size_t>(batch) * w.nsplit * heads * n_ltiles * 128 * 2);
  w.mask_bits = ar.take<uint64_t>(static_cast<size_t>(batch) * ((n_ctx + 63) / 64));
  w.bytes = ar.off + 256;
}
}  // namespace

extern "C" {

size_t hn_attention_workspace_bytes(int batch, int n_q, long n_ctx, int query_dim, int context_dim, int heads,
                                    int dim_head) {
  if (batch < 1 || n_q < 1 || n_ctx < 1 || query_dim < 1 || context_dim < 1 || heads < 1 || dim_head < 1) return 0;
  AttnWs w;
  plan_attn_ws(batch, n_q, n_ctx, query_dim, context_dim, heads, dim_head, false, nullptr, w);
  return w.bytes;
}

int hn_attention_forward(int batch, int n_q, long n_ctx, int query_dim, int context_dim, int heads, int dim_head,
                         const float* x, const float* context, const float* w_q, const float* w_kv,
                         const float* w_out, const float* b_out, const uint8_t* mask, float* out, void* workspace,
                         size_t workspace_bytes, void* cuda_stream) {
  return hn_attention_forward_cached(batch, n_q, n_ctx, query_dim, context_dim, heads, dim_head, x, context, w_q, w_kv,
                                     w_out, b_out, mask, out, workspace, workspace_bytes, 0, cuda_stream);
}

int hn_attention_forward_cached(int batch, int n_q, long n_ctx, int query_dim, int context_dim, int heads, int dim_head,
                                const float* x, const float* context, const float* w_q, const float* w_kv,
                                const float* w_out, const float* b_out, const uint8_t* mask, float* out,
                                void* workspace, size_t workspace_bytes, int weights_packed, void* cuda_stream) {
  HN_REQUIRE(batch >= 1 && n_q >= 1 && query_dim >= 1 && heads >= 1, "hn_attention_forward: bad shape");
  HN_REQUIRE(dim_head >= 1 && dim_head <= MAX_DIM_HEAD, "hn_attention_forward: dim_head must be in 1..128");
  HN_REQUIRE(x && w_q && w_kv && w_out && b_out && out && workspace, "hn_attention_forward: null argument");
  const bool self = (context == nullptr);
  if (self) {
    n_ctx = n_q;
    context_dim = query_dim;
  }
  HN_REQUIRE(n_ctx >= 1 && context_dim >= 1, "hn_attention_forward: bad context shape");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "hn_attention_forward: workspace must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  AttnWs w;
  plan_attn_ws(batch, n_q, n_ctx, query_dim, context_dim, heads, dim_head, self, static_cast<char*>(workspace), w);
  HN_REQUIRE(w.bytes <= workspace_bytes, "hn_attention_forward: workspace too small");
  const int hp = head_pitch(dim_head);
  const int inner = heads * dim_head, hw = heads * hp, sq = w.sq, sc = w.sc;
  const long rows = static_cast<long>(batch) * n_q, toks = static_cast<long>(batch) * n_ctx;
  HN_REQUIRE(rows < (1L << 31) && toks < (1L << 31), "hn_attention_forward: problem too large");
  const float scale = 2.f / std::sqrt(static_cast<float>(dim_head)) * LOG2E;
  const int ldc = self ? 2 * sq : (w.prec ? 2 * sc : ctx_ld(context_dim));
  const bool c_split = self || w.prec;
  int rc;
#define HN_TRY2(expr)         \
  do {                        \
    rc = (expr);              \
    if (rc != 0) return rc;   \
  } while (0)
  HN_TRY2(pack_plain(w.xh, 2 * sq, x, query_dim, static_cast<int>(rows), query_dim, sq, sq, st));
  if (!self)
    HN_TRY2(pack_plain(w.ch, ldc, context, context_dim, static_cast<int>(toks), context_dim, w.prec ? sc : ldc,
                       w.prec ? sc : 0, st));
  if (!weights_packed) {  // (the packed weights of an earlier call with the same shapes still sit in this workspace)
    HN_TRY2(pack_headpad_rows(w.wq, 2 * sq, 0, w_q, query_dim, 0, heads, dim_head, query_dim, scale, nullptr, sq, sq, hp,
                              st));
    HN_TRY2(pack_headpad_rows(w.wkv, 2 * sc, 0, w_kv, context_dim, 0, heads, dim_head, context_dim, 1.f, nullptr, sc, sc,
                              hp, st));
    HN_TRY2(pack_headpad_rows(w.wkv, 2 * sc, hw, w_kv, context_dim, inner, heads, dim_head, context_dim, 1.f, nullptr,
                              sc, sc, hp, st));
    HN_TRY2(pack_headpad_cols(w.wo, 2 * hw, w_out, inner, query_dim, heads, dim_head, hw, hw, hp, st));
  }
  GemmArgs gq{w.xh, w.wq, static_cast<int>(rows), hw, query_dim, 2 * sq, 2 * sq, EPI_F16, 0, nullptr, w.q,
              w.prec ? 2 * hw : hw, 3, sq, sq, w.prec ? hw : 0};
  HN_TRY2(launch_gemm(gq, st));
  GemmArgs gkv{w.ch, w.wkv, static_cast<int>(toks), 2 * hw, context_dim, ldc, 2 * sc, EPI_F16, 0, nullptr, w.kv,
               w.prec ? 4 * hw : 2 * hw, (c_split && w.prec) ? 3 : 2, sc, sc, w.prec ? 2 * hw : 0};
  HN_TRY2(launch_gemm(gkv, st));
  if (mask != nullptr) HN_TRY2(launch_pack_mask(mask, w.mask_bits, batch, n_ctx, st));
  AttnArgs aa{};
  aa.Q = w.q;
  aa.q_ld = w.prec ? 2 * hw : hw;
  aa.KV = w.kv;
  aa.kv_ld = w.prec ? 4 * hw : 2 * hw;
  aa.k_col0 = 0;
  aa.v_col0 = hw;
  aa.shared_kv = 0;
  aa.kd = 64;
  aa.precise = w.prec ? 1 : 0;
  aa.q_lo_off = hw;
  aa.kv_lo_off = 2 * hw;
  aa.hp = hp;
  aa.batch = batch;
  aa.L = n_q;
  aa.H = heads;
  aa.N = n_ctx;
  aa.nsplit = w.nsplit;
  aa.mask_bits = mask ? w.mask_bits : nullptr;
  aa.part_acc = w.part_acc;
  aa.part_ml = w.part_ml;
  HN_TRY2(launch_attention(aa, st));
  HN_TRY2(launch_combine_generic(w.part_acc, w.part_ml, batch, w.nsplit, heads, n_q, w.o, 2 * hw, hw, hp, st));
  GemmArgs go{w.o, w.wo, static_cast<int>(rows), query_dim, hw, 2 * hw, 2 * hw, EPI_LEAKY_F32, 0, b_out, out,
              query_dim, 3, hw, hw, 0};
  HN_TRY2(launch_gemm(go, st));
#undef HN_TRY2
  return 0;
}

// ------------------------------------------------------------------------------------ kernel-level entry points
int hn_op_gemm(const void* A, const void* B, int M, int N, int K, int lda, int ldb, int epi, int act,
               const float* bias, void* out, int ldo, int terms, int a_seg, int b_seg, int out_seg,
               void* cuda_stream) {
  GemmArgs g{static_cast<const __half*>(A), static_cast<const __half*>(B), M, N, K, lda, ldb, epi, act, bias, out,
             ldo, terms, a_seg, b_seg, out_seg};
  return launch_gemm(g, static_cast<cudaStream_t>(cuda_stream));
}

int hn_op_layernorm_f16(const float* x, int ldx, const float* gamma, const float* beta, void* y, int ldy, int seg,
                        int lo_seg, long rows, int D, void* cuda_stream) {
  return launch_layernorm_f16(x, ldx, gamma, beta, static_cast<__half*>(y), ldy, seg, lo_seg, rows, D,
                              static_cast<cudaStream_t>(cuda_stream));
}

int hn_op_build_context(const float* raw, void* z, int ldz, int small, int batch, int c_raw, int n_axes,
                        const int* axis_sizes, int n_bands, float max_freq, int fourier, float* tab,
                        void* cuda_stream) {
  HN_REQUIRE(raw && z && axis_sizes && n_axes >= 1 && n_axes <= HN_MAX_AXES, "hn_op_build_context: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  long N = 1;
  for (int a = 0; a < n_axes; ++a) N *= axis_sizes[a];
  if (fourier) {
    HN_REQUIRE(tab != nullptr, "hn_op_build_context: table scratch required");
    int rc = launch_axis_tables(tab, axis_sizes, n_axes, n_bands, max_freq, st);
    if (rc != 0) return rc;
  }
  if (small)
    return launch_build_z_small(raw, static_cast<__half*>(z), ldz, batch, N, c_raw, n_axes, axis_sizes, n_bands, tab,
                                fourier, st, 0, small == 2 ? 1 : 0);
  return launch_build_z_large(raw, static_cast<__half*>(z), ldz, 0, batch, N, c_raw, n_axes, axis_sizes, n_bands,
                              tab, fourier, st);
}

int hn_op_attention_nsplit(int batch, int L, int H, long N, int small_kd) {
  if (small_kd > 0) return small_attention_pick_nsplit(batch, L, H, N, small_kd);
  return attention_pick_nsplit(batch, L, H, N);
}

int hn_op_attention(const void* Q, int q_ld, const void* KV, long kv_ld, int k_col0, int v_col0, int shared_kv,
                    int c_ones, int head_pitch_cols, int batch, int L, int H, long N, int nsplit, const uint8_t* mask,
                    void* mask_bits_scratch, float* part_acc, float* part_ml, void* cuda_stream) {
  // shared_kv: 0 generic, 1 small-context kernel (xattn_small.cu) on single fp16 operands, 3 the same on split
  // operands (Q' rows [hi | lo at q_ld / 2], z rows [hi (kd) | lo (kd)], kv_ld = 2 kd); 4 = 3 with z rows whose lo half
  // carries the merged tail written by the context-row builder (kd 32, 17 <= C <= 23) — the mode the forward uses
  HN_REQUIRE(shared_kv == 0 || shared_kv == 1 || shared_kv == 3 || shared_kv == 4,
             "hn_op_attention: shared_kv must be 0, 1, 3 or 4");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  AttnArgs aa{};
  aa.Q = static_cast<const __half*>(Q);
  aa.q_ld = q_ld;
  aa.KV = static_cast<const __half*>(KV);
  aa.kv_ld = kv_ld;
  aa.k_col0 = k_col0;
  aa.v_col0 = v_col0;
  aa.shared_kv = shared_kv ? 1 : 0;
  aa.c_ones = c_ones;
  aa.hp = head_pitch_cols > 0 ? head_pitch_cols : 64;
  aa.kd = shared_kv >= 3 ? static_cast<int>(kv_ld / 2) : (shared_kv ? static_cast<int>(kv_ld) : 64);
  if (shared_kv >= 3) {
    aa.precise = 1;
    aa.q_lo_off = q_ld / 2;
    aa.z_tail_merged = shared_kv == 4;  // the z rows carry the merged tail (see hn_op_build_context, split rows)
  }
  aa.batch = batch;
  aa.L = L;
  aa.H = H;
  aa.N = N;
  aa.nsplit = nsplit;
  aa.part_acc = part_acc;
  aa.part_ml = part_ml;
  if (mask != nullptr) {
    HN_REQUIRE(mask_bits_scratch != nullptr, "hn_op_attention: mask scratch required");
    int rc = launch_pack_mask(mask, static_cast<uint64_t*>(mask_bits_scratch), batch, N, st);
    if (rc != 0) return rc;
    aa.mask_bits = static_cast<const uint64_t*>(mask_bits_scratch);
  }
  return launch_attention(aa, st);
}

int hn_op_combine(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L, int small_C,
                  int zw, int dh, int head_pitch_cols, const float* Wv, const float* bv, void* O, int o_ld,
                  void* cuda_stream) {
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  if (small_C > 0)
    return launch_combine_vproj(part_acc, part_ml, batch, nsplit, H, L, small_C, zw, dh, Wv, bv,
                                static_cast<__half*>(O), o_ld, 0, head_pitch_cols, st);
  return launch_combine_generic(part_acc, part_ml, batch, nsplit, H, L, static_cast<__half*>(O), o_ld, 0,
                                head_pitch_cols, st);
}

}  // extern "C"
