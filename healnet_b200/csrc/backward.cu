// backward.cu — training step support: the gradient of HealNet.forward
// (reference healnet/models/healnet.py:190-250) with respect to every parameter, as the reference's training loop
// obtains it from autograd (healnet/main.py:426-467: loss.backward() on the logits; L1 regularisation and the optimizer
// stay in PyTorch, healnet/utils/train_utils.py:5-14).
//
// hn_forward_train runs the ordinary forward and leaves, per PreNorm(module) block, on a caller-owned TAPE: the fp32
// residual stream entering the block, its LayerNorm (split fp16, 22 significant bits), the normalised attention output
// (or the gated hidden rows of a feed-forward) and the merged softmax row statistics (M, den). The standardised
// context rows z stay in the forward workspace, which the caller keeps alive until hn_backward.
//
// hn_backward walks the blocks in reverse. Nothing of size N (tokens) is ever stored per latent row: attention
// probabilities are RECOMPUTED from scores and the saved row statistics —
//   * small-context streaming path (image / volume): no gradient flows into the context, so the whole attention
//     backward is one more pass over the token axis that accumulates dr_l = sum_t p_lt (du_l.z_t - du_l.u_l) z_t
//     (bwdops.cu: small_attn_bwd_kernel); everything else happens on (rows x C) arrays — the reassociation of the
//     forward (xattn_small.cu) is undone analytically: r = c gamma * (Wk^T q), o = Wv (gamma * u + beta);
//   * generic path (wide contexts, tabular row, latent self-attention): S, P and dP of one attention call are
//     materialised in scratch ((b h) x L x N fp32) between strided batched contractions;
//   * latent side: dgrad / wgrad contractions, LayerNorm / gate / LeakyReLU backward as row kernels.
// All arithmetic of this first generation is exact fp32 (sgemm.cu) — gradients match the reference's autograd to ~1e-5
// (tests/test_gpu_backward.py, golden gradients produced by the unmodified reference). Parameter gradients are
// ACCUMULATED into the buffers registered with hn_set_grads (tied layers register the same buffer several times).
#include <cmath>
#include <cstring>

#include "../../include/healnet_b200.h"
#include "bwd.cuh"
#include "common.cuh"
#include "handle.cuh"

using namespace hn;

namespace hn {

size_t plan_tape(const hn_handle* h, int batch, const Workspace& ws, const int* skip, std::vector<BlockRec>& blocks,
                 size_t& x_final) {
  const hn_desc& d = h->d;
  const int M = h->M, L = d.l_c, D = d.l_d;
  const size_t rows = static_cast<size_t>(batch) * L;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    off = (off + 255) & ~size_t(255);
    const size_t o = off;
    off += bytes;
    return o;
  };
  blocks.clear();
  auto push_attn = [&](int kind, int l, int m, size_t o_cols, int heads) {
    BlockRec r;
    r.kind = kind;
    r.layer = l;
    r.m = m;
    r.x_in = take(rows * D * sizeof(float));
    r.xn = take(rows * 2 * h->segD * sizeof(__half));
    r.stats = take(static_cast<size_t>(batch) * heads * L * 2 * sizeof(float));
    r.o = take(rows * 2 * o_cols * sizeof(__half));
    blocks.push_back(r);
  };
  auto push_ff = [&](int l, int m) {
    BlockRec r;
    r.kind = 3;
    r.layer = l;
    r.m = m;
    r.x_in = take(rows * D * sizeof(float));
    r.xn = take(rows * 2 * h->segD * sizeof(__half));
    r.o = take(rows * 2 * h->seg4D * sizeof(__half));
    blocks.push_back(r);
  };
  for (int l = 0; l < d.depth; ++l) {
    for (int m = 0; m < M; ++m) {
      const ModPlan& mp = ws.mod[m];
      if (mp.present) {
        if (mp.small)
          push_attn(0, l, m, round_up(d.x_heads * mp.zw, 64), d.x_heads);
        else
          push_attn(1, l, m, d.x_heads * h->hpx, d.x_heads);
        push_ff(l, m);
      }
      if (d.self_per_cross_attn && !(skip != nullptr && skip[m] != 0)) {
        push_attn(2, l, M, d.l_heads * h->hpl, d.l_heads);
        push_ff(l, M);
      }
    }
  }
  x_final = take(rows * D * sizeof(float));
  return off + 256;
}

}  // namespace hn

namespace {

inline SgOperand F32(const float* p, long s_row, long s_col, long s_b1 = 0, long s_b2 = 0) {
  SgOperand o;
  o.p = p;
  o.type = 0;
  o.s_row = s_row;
  o.s_col = s_col;
  o.s_b1 = s_b1;
  o.s_b2 = s_b2;
  return o;
}
inline SgOperand H16(const __half* p, int lo_off, long s_row, long s_col, long s_b1 = 0, long s_b2 = 0) {
  SgOperand o;
  o.p = p;
  o.type = 1;
  o.lo_off = lo_off;
  o.s_row = s_row;
  o.s_col = s_col;
  o.s_b1 = s_b1;
  o.s_b2 = s_b2;
  return o;
}
inline int sg(cudaStream_t st, int M, int N, int K, const SgOperand& A, const SgOperand& B, float* C, long c_row,
              float alpha = 1.f, int accumulate = 0, int nb1 = 1, long c_b1 = 0, int nb2 = 1, long c_b2 = 0) {
  SgArgs a;
  a.M = M;
  a.N = N;
  a.K = K;
  a.A = A;
  a.B = B;
  a.C = C;
  a.c_row = c_row;
  a.c_col = 1;
  a.c_b1 = c_b1;
  a.c_b2 = c_b2;
  a.nb1 = nb1;
  a.nb2 = nb2;
  a.alpha = alpha;
  a.accumulate = accumulate;
  return launch_sgemm(a, st);
}

// scratch of hn_backward, carved from one caller-owned buffer
struct BwdScratch {
  float *dx, *dy, *dxn, *lnstats, *colpart;
  float *hbuf, *dhid;
  float *dO, *dq, *ofull, *qf;
  float *S, *dP, *dKV, *dWp, *sv;
  float *u32, *cnu, *g, *du, *w, *rl, *dr, *dw, *delta, *dr_part;
  void *pA, *pB;               // bf16 hi/lo operand rows of the tensor-core GEMMs
  size_t capA, capB;           // their capacities in elements
  float* dh2;                  // [rows][8D] gate gradients (tensor-core feed-forward path)
  __half *rq, *duq;            // tensor-core streaming backward: split operand rows
  float *row_a, *row_d, *row_s;
  float* dpooled;
  size_t bytes;
};

void plan_scratch(const hn_handle* h, int batch, const Workspace& ws, char* base, BwdScratch& s) {
  const hn_desc& d = h->d;
  const int M = h->M, L = d.l_c, D = d.l_d;
  const size_t rows = static_cast<size_t>(batch) * L;
  Arena ar;
  ar.base = base;
  const int I = h->I, lI = d.self_per_cross_attn ? h->lI : 0;
  const size_t Imax = static_cast<size_t>(I > lI ? I : lI);
  const size_t Hmax = static_cast<size_t>(d.x_heads > d.l_heads ? d.x_heads : d.l_heads);
  size_t n_gen = d.self_per_cross_attn ? L : 0, c_gen = d.self_per_cross_attn ? D : 0, c_small = 0;
  size_t sp_part = 0, zw_max = 0;
  for (int m = 0; m < M; ++m) {
    const ModPlan& mp = ws.mod[m];
    if (!mp.present) continue;
    if (mp.small) {
      c_small = static_cast<size_t>(mp.C) > c_small ? mp.C : c_small;
      size_t part = static_cast<size_t>(small_attn_bwd_nsplit(batch, d.x_heads, L, mp.Nl)) * rows * d.x_heads * mp.C;
      // tensor-core variant: one [128-row block][zw] partial per (sample, split, head, latent tile)
      const size_t part_tc = static_cast<size_t>(batch) * small_attention_pick_nsplit(batch, L, d.x_heads, mp.Nl, mp.zw) *
                             d.x_heads * ((L + 127) / 128) * 128 * mp.zw;
      part = part_tc > part ? part_tc : part;
      sp_part = part > sp_part ? part : sp_part;
      zw_max = static_cast<size_t>(mp.zw) > zw_max ? mp.zw : zw_max;
    } else {
      n_gen = static_cast<size_t>(mp.Nl) > n_gen ? mp.Nl : n_gen;
      c_gen = static_cast<size_t>(mp.C) > c_gen ? mp.C : c_gen;
    }
  }
  size_t colw = 8 * static_cast<size_t>(D);
  colw = 2 * Imax > colw ? 2 * Imax : colw;
  colw = c_gen > colw ? c_gen : colw;
  s.dx = ar.take<float>(rows * D);
  s.dy = ar.take<float>(rows * D);
  s.dxn = ar.take<float>(rows * D);
  s.lnstats = ar.take<float>(rows * 2);
  s.colpart = ar.take<float>(64 * colw);
  s.hbuf = ar.take<float>(rows * 8 * D);
  s.dhid = ar.take<float>(rows * 4 * D);
  s.dO = ar.take<float>(rows * Imax);
  s.dq = ar.take<float>(rows * Imax);
  s.ofull = ar.take<float>(rows * Imax);
  s.qf = ar.take<float>(rows * Imax);
  s.S = ar.take<float>(static_cast<size_t>(batch) * Hmax * L * n_gen);
  s.dP = ar.take<float>(static_cast<size_t>(batch) * Hmax * L * n_gen);
  s.dKV = ar.take<float>(static_cast<size_t>(batch) * n_gen * 2 * Imax);
  s.dWp = ar.take<float>(2 * Imax * (c_gen > c_small ? c_gen : c_small));
  s.sv = ar.take<float>(2 * Imax);
  const size_t rc = rows * d.x_heads * c_small;
  s.u32 = ar.take<float>(rc);
  s.cnu = ar.take<float>(rc);
  s.g = ar.take<float>(rc);
  s.du = ar.take<float>(rc);
  s.w = ar.take<float>(rc);
  s.rl = ar.take<float>(rc);
  s.dr = ar.take<float>(rc);
  s.dw = ar.take<float>(rc);
  s.delta = ar.take<float>(rows * d.x_heads);
  s.dr_part = ar.take<float>(sp_part);
  {
    const size_t sR = round_up_l(static_cast<long>(rows), 64), sDd = h->segD, s8 = round_up(8 * D, 64), F = 4 * static_cast<size_t>(D);
    size_t a = 2 * F * 2 * sR;                       // (dh)^T
    a = rows * 2 * s8 > a ? rows * 2 * s8 : a;       // dh
    size_t bb = F * 2 * sR;                          // hid^T
    bb = static_cast<size_t>(D) * 2 * s8 > bb ? static_cast<size_t>(D) * 2 * s8 : bb;   // W1^T
    bb = F * 2 * sDd > bb ? F * 2 * sDd : bb;        // W2^T
    // attention blocks (mm() below): the latent-side products and the per-head products of the materialised core.
    // Operand rows are [hi | lo] of round_up(K, 64) columns each.
    auto seg2 = [](size_t k) { return 2 * static_cast<size_t>(round_up_l(static_cast<long>(k), 64)); };
    auto up = [](size_t& x, size_t v) { if (v > x) x = v; };
    const size_t I2 = 2 * Imax;
    up(a, rows * seg2(D));  up(bb, Imax * seg2(D));            // dO = dy Wo, q = xn Wq^T
    up(a, D * seg2(rows));  up(bb, Imax * seg2(rows));         // dWo += dy^T o
    up(a, I2 * seg2(rows)); up(bb, D * seg2(rows));            // dWq / dWkv
    up(a, rows * seg2(I2)); up(bb, D * seg2(I2));              // dxn = dq Wq, dKV Wkv
    if (n_gen > 0) {
      // core of the generic / latent attention with n_gen tokens per sample, capped: larger token axes keep the fp32 path
      const size_t nz = static_cast<size_t>(batch) * Hmax, cap = static_cast<size_t>(96) << 20;
      size_t ca = nz * (n_gen > static_cast<size_t>(L) ? n_gen : L) * seg2(n_gen > static_cast<size_t>(L) ? n_gen : L);
      size_t cb = nz * n_gen * seg2(L > 128 ? L : 128);
      up(a, ca < cap ? ca : cap);
      up(bb, cb < cap ? cb : cap);
    }
    s.capA = a;
    s.capB = bb;
    s.pA = ar.take<__half>(a);
    s.pB = ar.take<__half>(bb);
    s.dh2 = ar.take<float>(rows * 8 * D);
  }
  s.rq = ar.take<__half>(rows * 2 * d.x_heads * zw_max);
  s.duq = ar.take<__half>(rows * 2 * d.x_heads * zw_max);
  s.row_a = ar.take<float>(rows * d.x_heads);
  s.row_d = ar.take<float>(rows * d.x_heads);
  s.row_s = ar.take<float>(rows * d.x_heads);
  s.dpooled = ar.take<float>(static_cast<size_t>(batch) * D);
  s.bytes = ar.off + 256;
}

#define BW(expr)              \
  do {                        \
    int _rc = (expr);         \
    if (_rc != 0) return _rc; \
  } while (0)

// dx += LayerNorm-backward(dxn) through PreNorm.norm of the block; accumulates the affine gradients
int ln_backward(const float* x_in, const float* dxn, const float* gamma, float* g_gamma, float* g_beta, long rows,
                int D, BwdScratch& s, cudaStream_t st) {
  BW(launch_ln_bwd_rows(x_in, dxn, gamma, s.dx, s.lnstats, rows, D, 1, st));
  BW(launch_colsum(2, dxn, D, x_in, D, s.lnstats, rows, D, 1.f, g_gamma, 1, s.colpart, st));
  BW(launch_colsum(0, dxn, D, nullptr, 0, nullptr, rows, D, 1.f, g_beta, 1, s.colpart, st));
  return 0;
}

// C[b1][b2][m][n] (+)= alpha * sum_k A(m, k) B(k, n) — the contraction of sg() above — on tcgen05: both operands are
// packed to bf16 hi/lo rows with the contraction index contiguous (three-term products: 16 significant bits, fp32
// range), batches stacked along the rows, one batched GEMM launch. Small products, operands that do not fit the
// packing buffers, and the fp32 checker variant (hn_set_backward_variant(1)) keep the exact fp32 SIMT kernel.
int mm(const hn_handle* h, BwdScratch& s, cudaStream_t st, int M, int N, int K, const SgOperand& A, const SgOperand& B,
       float* C, long c_row, float alpha = 1.f, int accumulate = 0, int nb1 = 1, long c_b1 = 0, int nb2 = 1,
       long c_b2 = 0) {
  const long nz = static_cast<long>(nb1) * nb2;
  const size_t segK = static_cast<size_t>(round_up_l(K, 64));
  const size_t needA = static_cast<size_t>(nz) * M * 2 * segK, needB = static_cast<size_t>(nz) * N * 2 * segK;
  const bool small = static_cast<double>(M) * N * K * nz < 4.0e6 || K < 16 || M < 16 || N < 8;
  if (h->bwd_variant != 0 || small || needA > s.capA || needB > s.capB || c_row >= 2147483647L || nz > 65535 ||
      static_cast<long>(M) * nz >= 2147483647L || static_cast<long>(N) * nz >= 2147483647L)
    return sg(st, M, N, K, A, B, C, c_row, alpha, accumulate, nb1, c_b1, nb2, c_b2);
  const int seg = static_cast<int>(segK);
  SgOperand Bt = B;   // B(k, n) viewed as rows n x contraction index k
  Bt.s_row = B.s_col;
  Bt.s_col = B.s_row;
  BW(launch_pack_bf16_pair(A, M, Bt, N, K, nb1, nb2, s.pA, s.pB, seg, st));
  GemmArgs g{static_cast<const __half*>(s.pA), static_cast<const __half*>(s.pB), M, N, K, 2 * seg, 2 * seg,
             accumulate ? EPI_RES : EPI_F32, 0, nullptr, C, static_cast<int>(c_row), 3, seg, seg, 0};
  g.bf16 = 1;
  g.nbatch = static_cast<int>(nz);
  g.nb2 = nb2;
  g.a_brows = M;
  g.b_brows = N;
  g.out_b1 = c_b1;
  g.out_b2 = c_b2;
  g.alpha = alpha;
  return launch_gemm(g, st);
}

// Feed-forward backward with every contraction on tensor cores (mm(): bf16 hi/lo operands, three-term products), the
// recomputed pre-activations through the forward's own fp16-split GEMM.
int ff_backward_tc(hn_handle* h, const BlockRec& rec, const std::vector<const float*>& wf, const std::vector<float*>& gf,
                   const FFPacked& fp, const char* tape, long rows, BwdScratch& s, cudaStream_t st) {
  const hn_desc& d = h->d;
  const int D = d.l_d, F = 4 * D, sD = h->segD, s4 = h->seg4D;
  const int R = static_cast<int>(rows);
  const float* x_in = reinterpret_cast<const float*>(tape + rec.x_in);
  const __half* xn = reinterpret_cast<const __half*>(tape + rec.xn);
  const __half* hid = reinterpret_cast<const __half*>(tape + rec.o);
  BW(launch_colsum(0, s.dx, D, nullptr, 0, nullptr, rows, D, 1.f, gf[5], 1, s.colpart, st));
  BW(mm(h, s, st, D, F, R, F32(s.dx, 1, D), H16(hid, s4, 2 * s4, 1), gf[4], F, 1.f, 1));            // dW2 += dx^T hid
  BW(mm(h, s, st, R, F, D, F32(s.dx, D, 1), F32(wf[4], F, 1), s.dhid, F));                           // dhid = dx W2
  // [a | g] (interleaved, bias included) = LN(x) W1^T + b1 through the forward's GEMM
  GemmArgs g1{xn, fp.W1, R, 2 * F, D, 2 * sD, 2 * sD, EPI_F32, 0, fp.b1, s.hbuf, 2 * F, 3, sD, sD, 0};
  BW(launch_gemm(g1, st));
  BW(launch_gate_bwd_il(s.hbuf, s.dhid, s.dh2, rows, F, d.snn, st));
  BW(launch_colsum(0, s.dh2, 2 * F, nullptr, 0, nullptr, rows, 2 * F, 1.f, gf[3], 1, s.colpart, st));
  BW(mm(h, s, st, 2 * F, D, R, F32(s.dh2, 1, 2 * F), H16(xn, sD, 2 * sD, 1), gf[2], D, 1.f, 1));    // dW1 += dh^T xn
  BW(mm(h, s, st, R, D, 2 * F, F32(s.dh2, 2 * F, 1), F32(wf[2], D, 1), s.dxn, D));                   // dxn = dh W1
  return ln_backward(x_in, s.dxn, wf[0], gf[0], gf[1], rows, D, s, st);
}

// x_out = x_in + W2 (a * act(g)) + b2, [a | g] = W1 LN(x_in) + b1   (healnet.py:339-351, 237/245)
int ff_backward(hn_handle* h, const BlockRec& rec, const std::vector<const float*>& wf, const std::vector<float*>& gf,
                const char* tape, long rows, BwdScratch& s, cudaStream_t st) {
  const hn_desc& d = h->d;
  const int D = d.l_d, F = 4 * D, sD = h->segD, s4 = h->seg4D;
  const float* x_in = reinterpret_cast<const float*>(tape + rec.x_in);
  const __half* xn = reinterpret_cast<const __half*>(tape + rec.xn);
  const __half* hid = reinterpret_cast<const __half*>(tape + rec.o);
  const int R = static_cast<int>(rows);
  BW(launch_colsum(0, s.dx, D, nullptr, 0, nullptr, rows, D, 1.f, gf[5], 1, s.colpart, st));
  BW(sg(st, D, F, R, F32(s.dx, 1, D), H16(hid, s4, 2 * s4, 1), gf[4], F, 1.f, 1));                   // dW2 += dx^T hid
  BW(sg(st, R, F, D, F32(s.dx, D, 1), F32(wf[4], F, 1), s.dhid, F));                                  // dhid = dx W2
  BW(sg(st, R, 2 * F, D, H16(xn, sD, 2 * sD, 1), F32(wf[2], 1, D), s.hbuf, 2 * F));                   // [a | g] - b1
  BW(launch_gate_bwd(s.hbuf, wf[3], s.dhid, rows, F, d.snn, st));
  BW(launch_colsum(0, s.hbuf, 2 * F, nullptr, 0, nullptr, rows, 2 * F, 1.f, gf[3], 1, s.colpart, st));
  BW(sg(st, 2 * F, D, R, F32(s.hbuf, 1, 2 * F), H16(xn, sD, 2 * sD, 1), gf[2], D, 1.f, 1));           // dW1 += dh^T xn
  BW(sg(st, R, D, 2 * F, F32(s.hbuf, 2 * F, 1), F32(wf[2], D, 1), s.dxn, D));                         // dxn = dh W1
  return ln_backward(x_in, s.dxn, wf[0], gf[0], gf[1], rows, D, s, st);
}

// The materialised attention core shared by the generic cross-attention and the latent self-attention:
// in: dO [rows][I], Q (split, log2-scaled, head pitch hp) and K / V (split, head pitch hp) of b samples of N
// tokens; out: dq [rows][I] (w.r.t. the unscaled q), dKV [b*N][2I] (K gradient in columns [0, I), V in [I, 2I)).
int attention_core_backward(const hn_handle* h, int batch, int H, int L, long N, int dh, int hp, float c_nat, const __half* Q, int q_ld,
                            int q_lo, const __half* KV, long kv_ld, int kv_lo, int k_col0, int v_col0,
                            const float* stats, const uint64_t* mask_bits, BwdScratch& s, cudaStream_t st) {
  const int I = H * dh;
  const long LN_ = static_cast<long>(L) * N, HLN = static_cast<long>(H) * LN_;
  const int n = static_cast<int>(N);
  // S = Q K^T (log2 units)
  BW(mm(h, s, st, L, n, hp, H16(Q, q_lo, q_ld, 1, static_cast<long>(L) * q_ld, hp),
        H16(KV + k_col0, kv_lo, 1, kv_ld, N * kv_ld, hp), s.S, N, 1.f, 0, batch, HLN, H, LN_));
  BW(launch_softmax_recompute(s.S, stats, H, L, N, mask_bits, static_cast<long>(batch) * H * L, st));
  // dV[n][h, d] = sum_l P[l][n] dO[l][h, d]
  BW(mm(h, s, st, n, dh, L, F32(s.S, 1, N, HLN, LN_), F32(s.dO, I, 1, static_cast<long>(L) * I, dh), s.dKV + I, 2 * I, 1.f, 0,
        batch, N * 2 * I, H, dh));
  // dP[l][n] = dO[l] . V[n]
  BW(mm(h, s, st, L, n, dh, F32(s.dO, I, 1, static_cast<long>(L) * I, dh), H16(KV + v_col0, kv_lo, 1, kv_ld, N * kv_ld, hp), s.dP,
        N, 1.f, 0, batch, HLN, H, LN_));
  BW(launch_softmax_bwd(s.dP, s.S, N, static_cast<long>(batch) * H * L, st));  // dt = P (dP - sum P dP), natural-log units
  // dq = c dt K ;  dK = c dt^T q = ln2 dt^T Q  (Q carries c log2(e))
  BW(mm(h, s, st, L, dh, n, F32(s.dP, N, 1, HLN, LN_), H16(KV + k_col0, kv_lo, kv_ld, 1, N * kv_ld, hp), s.dq, I, c_nat, 0, batch,
        static_cast<long>(L) * I, H, dh));
  BW(mm(h, s, st, n, dh, L, F32(s.dP, 1, N, HLN, LN_), H16(Q, q_lo, q_ld, 1, static_cast<long>(L) * q_ld, hp), s.dKV, 2 * I,
        0.69314718055994530942f, 0, batch, N * 2 * I, H, dh));
  return 0;
}

}  // namespace

extern "C" {

int hn_set_grads(hn_handle* h, int layer, int slot, void* const* dev_ptrs, int n) {
  HN_REQUIRE(h != nullptr && dev_ptrs != nullptr, "hn_set_grads: null argument");
  HN_REQUIRE(layer >= -1 && layer < h->d.depth, "hn_set_grads: layer out of range");
  const int idx = slot_index(h, layer, slot);
  HN_REQUIRE(slot >= 0 && idx < static_cast<int>(h->w.size()), "hn_set_grads: slot out of range");
  HN_REQUIRE(static_cast<size_t>(n) == h->w[idx].size(), "hn_set_grads: register the weights of the slot first (same count)");
  if (h->g.size() != h->w.size()) h->g.assign(h->w.size(), std::vector<float*>());
  std::vector<float*> v(n);
  for (int i = 0; i < n; ++i) {
    HN_REQUIRE(dev_ptrs[i] != nullptr, "hn_set_grads: null gradient pointer");
    v[i] = static_cast<float*>(dev_ptrs[i]);
  }
  h->g[idx] = v;
  return 0;
}

static int fill_present(const hn_handle* h, const int* axis_sizes, bool* present_all) {
  (void)axis_sizes;
  for (int m = 0; m < h->M; ++m) present_all[m] = true;
  return 0;
}

size_t hn_tape_bytes(const hn_handle* h, int batch, const int* axis_sizes) {
  if (h == nullptr || batch < 1 || axis_sizes == nullptr) {
    set_error("hn_tape_bytes: bad argument");
    return 0;
  }
  Workspace ws;
  bool present[HN_MAX_MODALITIES];
  fill_present(h, axis_sizes, present);
  if (plan_workspace(h, batch, axis_sizes, present, 0, nullptr, ws) != 0) return 0;
  std::vector<BlockRec> blocks;
  size_t xf = 0;
  return plan_tape(h, batch, ws, nullptr, blocks, xf);
}

size_t hn_backward_scratch_bytes(const hn_handle* h, int batch, const int* axis_sizes) {
  if (h == nullptr || batch < 1 || axis_sizes == nullptr) {
    set_error("hn_backward_scratch_bytes: bad argument");
    return 0;
  }
  Workspace ws;
  bool present[HN_MAX_MODALITIES];
  fill_present(h, axis_sizes, present);
  if (plan_workspace(h, batch, axis_sizes, present, 0, nullptr, ws) != 0) return 0;
  BwdScratch s;
  plan_scratch(h, batch, ws, nullptr, s);
  return s.bytes;
}

int hn_set_backward_variant(hn_handle* h, int variant) {
  HN_REQUIRE(h != nullptr && (variant == 0 || variant == 1), "hn_set_backward_variant: 0 (tensor cores) or 1 (fp32 checker)");
  h->bwd_variant = variant;
  return 0;
}

int hn_forward_train(hn_handle* h, int batch, const void* const* modality_ptrs, void* const* modality_ready_events,
                     const int* axis_sizes, const int* skip_latent_block, const uint8_t* mask, long mask_tokens,
                     float* latents_out, float* logits_out, void* workspace, size_t workspace_bytes, void* tape,
                     size_t tape_bytes, void* cuda_stream) {
  HN_REQUIRE(h != nullptr && modality_ptrs != nullptr && axis_sizes != nullptr && tape != nullptr,
             "hn_forward_train: null argument");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(tape) & 255) == 0, "hn_forward_train: tape must be 256-byte aligned");
  TrainState& t = h->train;
  t.valid = false;
  t.batch = batch;
  memcpy(t.axis_sizes, axis_sizes, sizeof(t.axis_sizes));
  for (int m = 0; m < h->M; ++m) {
    t.present[m] = modality_ptrs[m] != nullptr;
    t.skip[m] = skip_latent_block != nullptr ? skip_latent_block[m] : 0;
  }
  t.mask_tokens = mask != nullptr ? mask_tokens : 0;
  Workspace ws;
  int rc = plan_workspace(h, batch, axis_sizes, t.present, t.mask_tokens, nullptr, ws);
  if (rc != 0) return rc;
  t.bytes = plan_tape(h, batch, ws, t.skip, t.blocks, t.x_final);
  HN_REQUIRE(t.bytes <= tape_bytes, "hn_forward_train: tape too small (see hn_tape_bytes)");
  h->tape = static_cast<char*>(tape);
  rc = forward_impl(h, batch, modality_ptrs, modality_ready_events, axis_sizes, nullptr, nullptr, skip_latent_block, mask,
                    mask_tokens, latents_out, logits_out, workspace, workspace_bytes, cuda_stream);
  h->tape = nullptr;
  t.valid = rc == 0;
  return rc;
}

int hn_backward(hn_handle* h, const float* grad_latents, const float* grad_logits, void* workspace,
                size_t workspace_bytes, const void* tape_v, size_t tape_bytes, void* scratch, size_t scratch_bytes,
                void* cuda_stream) {
  HN_REQUIRE(h != nullptr && workspace != nullptr && tape_v != nullptr && scratch != nullptr, "hn_backward: null argument");
  HN_REQUIRE((grad_latents != nullptr) != (grad_logits != nullptr),
             "hn_backward: pass the gradient of exactly one output (latents or logits)");
  const TrainState& t = h->train;
  HN_REQUIRE(t.valid, "hn_backward: no training-mode forward recorded on this handle (hn_forward_train)");
  HN_REQUIRE(t.bytes <= tape_bytes, "hn_backward: tape too small");
  HN_REQUIRE(h->g.size() == h->w.size(), "hn_backward: register gradient buffers first (hn_set_grads)");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "hn_backward: scratch must be 256-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
  const char* tape = static_cast<const char*>(tape_v);
  const hn_desc& d = h->d;
  const int M = h->M, L = d.l_c, D = d.l_d, sD = h->segD, batch = t.batch;
  const long rows = static_cast<long>(batch) * L;
  const int R = static_cast<int>(rows);
  Workspace ws;
  int rc = plan_workspace(h, batch, t.axis_sizes, t.present, t.mask_tokens, static_cast<char*>(workspace), ws);
  if (rc != 0) return rc;
  HN_REQUIRE(ws.bytes <= workspace_bytes, "hn_backward: workspace too small");
  BwdScratch s;
  plan_scratch(h, batch, ws, static_cast<char*>(scratch), s);
  HN_REQUIRE(s.bytes <= scratch_bytes, "hn_backward: scratch too small (see hn_backward_scratch_bytes)");
  auto grads = [&](int layer, int slot) -> const std::vector<float*>& { return h->g[slot_index(h, layer, slot)]; };
  auto weights = [&](int layer, int slot) -> const std::vector<const float*>& { return h->w[slot_index(h, layer, slot)]; };
  for (const BlockRec& r : t.blocks) {
    const int slot = r.kind == 3 ? 2 * r.m + 1 : 2 * r.m;
    HN_REQUIRE(grads(r.layer, slot).size() == weights(r.layer, slot).size() && !grads(r.layer, slot).empty(),
               "hn_backward: a used layer slot has no registered gradient buffers");
  }

  // ---- head (or the latent array itself)
  const float* x_final = reinterpret_cast<const float*>(tape + t.x_final);
  if (grad_logits != nullptr) {
    HN_REQUIRE(d.final_classifier_head, "hn_backward: model has no classifier head");
    const std::vector<const float*>& wh = weights(-1, 1);
    const std::vector<float*>& gh = grads(-1, 1);
    HN_REQUIRE(wh.size() == 4 && gh.size() == 4, "hn_backward: to_logits weights / gradients not registered");
    BW(launch_head_bwd(x_final, batch, L, D, wh[0], wh[1], wh[2], d.out_dims, grad_logits, gh[0], gh[1], gh[2], gh[3],
                       s.dpooled, s.dx, st));
  } else {
    HN_CHECK_CUDA(cudaMemcpyAsync(s.dx, grad_latents, sizeof(float) * rows * D, cudaMemcpyDeviceToDevice, st));
  }

  for (int bi = static_cast<int>(t.blocks.size()) - 1; bi >= 0; --bi) {
    const BlockRec& rec = t.blocks[bi];
    const float* x_in = reinterpret_cast<const float*>(tape + rec.x_in);
    const float* x_out = bi + 1 < static_cast<int>(t.blocks.size())
                             ? reinterpret_cast<const float*>(tape + t.blocks[bi + 1].x_in)
                             : x_final;
    if (rec.kind == 3) {
      if (h->bwd_variant == 0)
        BW(ff_backward_tc(h, rec, weights(rec.layer, 2 * rec.m + 1), grads(rec.layer, 2 * rec.m + 1),
                          h->ff[rec.layer * (M + 1) + rec.m], tape, rows, s, st));
      else
        BW(ff_backward(h, rec, weights(rec.layer, 2 * rec.m + 1), grads(rec.layer, 2 * rec.m + 1), tape, rows, s, st));
      continue;
    }
    const std::vector<const float*>& wa = weights(rec.layer, 2 * rec.m);
    const std::vector<float*>& ga = grads(rec.layer, 2 * rec.m);
    const __half* xn = reinterpret_cast<const __half*>(tape + rec.xn);
    const __half* o = reinterpret_cast<const __half*>(tape + rec.o);
    const float* stats = reinterpret_cast<const float*>(tape + rec.stats);
    const bool self = rec.kind == 2;
    // x_out = x_in + LeakyReLU(o Wo^T + bo)   (healnet.py:383-386, 426, 236/244)
    const int iWq = self ? 2 : 4, iWkv = self ? 3 : 5, iWo = self ? 4 : 6, ibo = self ? 5 : 7;
    BW(launch_leaky_bwd(s.dx, x_out, x_in, s.dy, rows * D, st));
    BW(launch_colsum(0, s.dy, D, nullptr, 0, nullptr, rows, D, 1.f, ga[ibo], 1, s.colpart, st));
    const int H = self ? d.l_heads : d.x_heads, dh = self ? d.latent_dim_head : d.cross_dim_head;
    const int hp = self ? h->hpl : h->hpx, I = H * dh;
    const float c_nat = 2.f / std::sqrt(static_cast<float>(dh));
    BW(mm(h, s, st, R, I, D, F32(s.dy, D, 1), F32(wa[iWo], I, 1), s.dO, I));  // dO = dy Wo  (unpadded head layout)

    if (rec.kind == 0) {
      // ---------------------------------------------------------------- small-context cross-attention
      const ModPlan& mp = ws.mod[rec.m];
      const int C = mp.C, zw = mp.zw, sHZ = round_up(H * zw, 64);
      const long RH = rows * H;
      const float* gamma = wa[2];
      const float* beta = wa[3];
      const float* Wkv = wa[5];
      BW(launch_small_pre(o, 2 * sHZ, sHZ, zw, H, C, rows, gamma, beta, s.u32, s.cnu, st));
      // o_h = Wv_h (gamma * u + beta): rebuilt for the out-projection weight gradient
      BW(mm(h, s, st, R, dh, C, F32(s.cnu, static_cast<long>(H) * C, 1, C), F32(Wkv + static_cast<long>(I) * C, 1, C, static_cast<long>(dh) * C),
            s.ofull, I, 1.f, 0, H, dh));
      BW(mm(h, s, st, D, I, R, F32(s.dy, 1, D), F32(s.ofull, I, 1), ga[iWo], I, 1.f, 1));              // dWo += dy^T o
      // g = Wv_h^T dO_h ; dWv_h += dO_h^T (gamma * u + beta)
      BW(mm(h, s, st, R, C, dh, F32(s.dO, I, 1, dh), F32(Wkv + static_cast<long>(I) * C, C, 1, static_cast<long>(dh) * C), s.g,
            static_cast<long>(H) * C, 1.f, 0, H, C));
      BW(mm(h, s, st, dh, C, R, F32(s.dO, 1, I, dh), F32(s.cnu, static_cast<long>(H) * C, 1, C), ga[5] + static_cast<long>(I) * C, C,
            1.f, 1, H, static_cast<long>(dh) * C));
      BW(launch_colsum(1, s.g, C, s.u32, C, nullptr, RH, C, 1.f, ga[2], 1, s.colpart, st));     // dgamma += sum g * u
      BW(launch_colsum(0, s.g, C, nullptr, 0, nullptr, RH, C, 1.f, ga[3], 1, s.colpart, st));   // dbeta  += sum g
      BW(launch_small_du(s.g, s.u32, gamma, C, RH, s.du, s.delta, st));
      // scores: s_lt = r_l . z_t with r = c gamma * (Wk_h^T q_l)
      BW(mm(h, s, st, R, I, D, H16(xn, sD, 2 * sD, 1), F32(wa[iWq], 1, D), s.qf, I));                  // q = xn Wq^T
      BW(mm(h, s, st, R, C, dh, F32(s.qf, I, 1, dh), F32(Wkv, C, 1, static_cast<long>(dh) * C), s.w, static_cast<long>(H) * C, 1.f,
            0, H, C));
      BW(launch_scale_cols(s.w, gamma, c_nat * LOG2E, C, RH * C, s.rl, st));
      if (h->bwd_variant == 0) {
        // tensor-core streaming pass (xattn_small.cu: attn_small_bwd_kernel)
        const int ld = 2 * H * zw;
        BW(launch_small_bwd_prep(s.rl, s.du, s.delta, stats, batch, H, L, C, zw, s.rq, s.duq, ld, H * zw, s.row_a, s.row_d,
                                 s.row_s, st));
        SmallBwdTcArgs ta;
        ta.rq = s.rq;
        ta.duq = s.duq;
        ta.rq_ld = ld;
        ta.lo_off = H * zw;
        ta.z = mp.z;
        ta.kd = zw;
        ta.row_a = s.row_a;
        ta.row_d = s.row_d;
        ta.mask_bits = mp.masked ? ws.mask_bits : nullptr;
        ta.part = s.dr_part;
        ta.batch = batch;
        ta.H = H;
        ta.L = L;
        ta.nsplit = small_attention_pick_nsplit(batch, L, H, mp.Nl, zw);
        ta.N = mp.Nl;
        ta.C = C;
        ta.merged_tail = (zw == 32 && C >= 17 && C <= 23) ? 1 : 0;   // as launch_small_bwd_prep and the context-row builder wrote them
        BW(launch_small_attention_bwd(ta, st));
        BW(launch_small_bwd_finish(s.dr_part, s.row_s, batch, ta.nsplit, H, L, C, zw, s.dr, st));
      } else {
        // exact fp32 SIMT pass (bwdops.cu), kept as the checker of the tensor-core kernel
        SmallBwdArgs sa;
        sa.r = s.rl;
        sa.du = s.du;
        sa.delta = s.delta;
        sa.stats = stats;
        sa.z = mp.z;
        sa.z_ld = 2 * zw;
        sa.z_lo = zw;
        sa.mask_bits = mp.masked ? ws.mask_bits : nullptr;
        sa.dr_part = s.dr_part;
        sa.batch = batch;
        sa.H = H;
        sa.L = L;
        sa.C = C;
        sa.nsplit = small_attn_bwd_nsplit(batch, H, L, mp.Nl);
        sa.N = mp.Nl;
        sa.R_total = RH;
        BW(launch_small_attn_bwd(sa, s.dr, st));
      }
      BW(launch_colsum(1, s.w, C, s.dr, C, nullptr, RH, C, c_nat, ga[2], 1, s.colpart, st));    // dgamma += c sum w * dr
      BW(launch_scale_cols(s.dr, gamma, c_nat, C, RH * C, s.dw, st));                            // dw = c gamma * dr
      BW(mm(h, s, st, R, dh, C, F32(s.dw, static_cast<long>(H) * C, 1, C), F32(Wkv, 1, C, static_cast<long>(dh) * C), s.dq, I, 1.f, 0,
            H, dh));                                                                              // dq_h = Wk_h dw
      BW(mm(h, s, st, dh, C, R, F32(s.qf, 1, I, dh), F32(s.dw, static_cast<long>(H) * C, 1, C), ga[5], C, 1.f, 1, H,
            static_cast<long>(dh) * C));                                                          // dWk_h += q_h^T dw
    } else {
      // ---------------------------------------------------------------- generic cross-attention / latent self-attention
      const int ow = H * hp;
      BW(mm(h, s, st, D, dh, R, F32(s.dy, 1, D), H16(o, ow, 2 * ow, 1, hp), ga[iWo], I, 1.f, 1, H, dh));  // dWo += dy^T o
      const AttnPacked& ap = h->attn[rec.layer * (M + 1) + rec.m];
      if (self) {
        const int qw = 3 * ow;
        GemmArgs gq{xn, ap.Wq, R, qw, D, 2 * sD, 2 * sD, EPI_F16, 0, nullptr, ws.q, 2 * qw, 3, sD, sD, qw};
        BW(launch_gemm(gq, st));                                                                  // [Q | K | V] as in the forward
        BW(attention_core_backward(h, batch, H, L, L, dh, hp, c_nat, ws.q, 2 * qw, qw, ws.q, 2 * qw, qw, ow, 2 * ow, stats,
                                   nullptr, s, st));
        BW(mm(h, s, st, 2 * I, D, R, F32(s.dKV, 1, 2 * I), H16(xn, sD, 2 * sD, 1), ga[iWkv], D, 1.f, 1));  // dWkv += dKV^T xn
        BW(mm(h, s, st, R, D, 2 * I, F32(s.dKV, 2 * I, 1), F32(wa[iWkv], D, 1), s.dxn, D));               // dxn  = dKV Wkv
      } else {
        const ModPlan& mp = ws.mod[rec.m];
        const int C = mp.C, qw = ow, kvw = 2 * ow;
        const long tok = static_cast<long>(batch) * mp.Nl;
        GemmArgs gq{xn, ap.Wq, R, qw, D, 2 * sD, 2 * sD, EPI_F16, 0, nullptr, ws.q, 2 * qw, 3, sD, sD, qw};
        BW(launch_gemm(gq, st));
        GemmArgs gkv{mp.z, ap.Wkv, static_cast<int>(tok), kvw, C, mp.ldz, 2 * mp.segC, EPI_F16, 0, ap.bkv, ws.kv, 2 * kvw, 3,
                     mp.segC, mp.segC, kvw};
        BW(launch_gemm(gkv, st));
        BW(attention_core_backward(h, batch, H, L, mp.Nl, dh, hp, c_nat, ws.q, 2 * qw, qw, ws.kv, 2 * kvw, kvw, 0, ow, stats,
                                   mp.masked ? ws.mask_bits : nullptr, s, st));
        // K = Wk (gamma * z) (+ const), V = Wv (gamma * z + beta): weight gradients through the folded context LayerNorm
        BW(mm(h, s, st, 2 * I, C, static_cast<int>(tok), F32(s.dKV, 1, 2 * I), H16(mp.z, mp.segC, mp.ldz, 1), s.dWp, C));
        BW(launch_colsum(0, s.dKV, 2 * I, nullptr, 0, nullptr, tok, 2 * I, 1.f, s.sv, 0, s.colpart, st));
        BW(launch_kv_fold_bwd(s.dWp, wa[5], wa[2], wa[3], s.sv, 2 * I, I, C, ga[5], ga[2], ga[3], st));
      }
    }
    // q = Wq LN(x): dWq += dq^T xn ; dxn (+)= dq Wq
    BW(mm(h, s, st, I, D, R, F32(s.dq, 1, I), H16(xn, sD, 2 * sD, 1), ga[iWq], D, 1.f, 1));
    BW(mm(h, s, st, R, D, I, F32(s.dq, I, 1), F32(wa[iWq], D, 1), s.dxn, D, 1.f, self ? 1 : 0));
    BW(ln_backward(x_in, s.dxn, wa[0], ga[0], ga[1], rows, D, s, st));
  }
  // x0 = repeat(latents, 'n d -> b n d')   (healnet.py:225)
  const std::vector<float*>& gl = grads(-1, 0);
  HN_REQUIRE(gl.size() == 1, "hn_backward: gradient buffer of `latents` not registered");
  BW(launch_batch_sum(s.dx, gl[0], static_cast<long>(L) * D, batch, st));
  return 0;
}

}  // extern "C"
