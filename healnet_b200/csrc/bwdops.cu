// bwdops.cu — row / column / element-wise kernels of the backward pass (training step of the reference,
// healnet/main.py:426-467: loss.backward() through HealNet.forward, healnet/models/healnet.py:190-250) and the first
// generation of the streaming small-context attention backward. Everything here is exact fp32 SIMT arithmetic; the
// heavy contractions are in sgemm.cu (strided fp32) and gemm.cu (tcgen05).
//
// Conventions: "rows" = batch * l_c latent rows; R = rows * heads for per-head arrays [R][C]; gradients of parameters
// are ACCUMULATED (+=) into caller-zeroed buffers so that tied layers and repeated modules sum up naturally.
#include <cuda_bf16.h>

#include "common.cuh"
#include "bwd.cuh"

namespace hn {
namespace {

constexpr float LN_EPS = 1e-5f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float ld_split(const __half* p, long i, int lo_off) {
  return __half2float(p[i]) + (lo_off > 0 ? __half2float(p[i + lo_off]) : 0.f);
}

// ------------------------------------------------------------------ column reductions
// partial[s][c] = sum over the rows of slab s of f(r, c); MODE 0: a, 1: a * b, 2: a * (x - mu_r) * rstd_r (LayerNorm gamma)
template <int MODE>
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ a, long lda,
                                                             const float* __restrict__ b, long ldb,
                                                             const float* __restrict__ stats, long R, int C,
                                                             long rows_per_slab, float* __restrict__ partial) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const long r0 = static_cast<long>(blockIdx.y) * rows_per_slab;
  const long r1 = r0 + rows_per_slab < R ? r0 + rows_per_slab : R;
  float s = 0.f;
  if (c < C) {
    for (long r = r0 + w; r < r1; r += 8) {
      float v = a[r * lda + c];
      if (MODE == 1) v *= b[r * ldb + c];
      if (MODE == 2) v *= (b[r * ldb + c] - stats[2 * r]) * stats[2 * r + 1];
      s += v;
    }
  }
  red[w][lane] = s;
  __syncthreads();
  if (w == 0 && c < C) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][lane];
    partial[static_cast<long>(blockIdx.y) * C + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ partial, int slabs, int C, float scale,
                                    float* __restrict__ out, int accumulate) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t = 0.f;
  for (int s = 0; s < slabs; ++s) t += partial[static_cast<long>(s) * C + c];
  out[c] = (accumulate ? out[c] : 0.f) + scale * t;
}

// ------------------------------------------------------------------ LayerNorm backward (rows)
// y = (x - mu) * rstd * gamma + beta over the last dim (biased variance, eps 1e-5; healnet.py:310-311).
// dx[r] += rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma; also stores (mu, rstd) for the column pass.
__global__ void __launch_bounds__(256) ln_bwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          const float* __restrict__ gamma, float* __restrict__ dx,
                                                          float* __restrict__ stats, long rows, int D, int accumulate) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long r = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  const float* xr = x + r * D;
  const float* gr = dy + r * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += xr[c];
  const float mu = warp_sum(s) / D;
  float q = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float d = xr[c] - mu;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / D + LN_EPS);
  float m1 = 0.f, m2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float g = gr[c] * gamma[c];
    m1 += g;
    m2 += g * (xr[c] - mu) * rstd;
  }
  m1 = warp_sum(m1) / D;
  m2 = warp_sum(m2) / D;
  float* o = dx + r * D;
  for (int c = lane; c < D; c += 32) {
    const float g = gr[c] * gamma[c];
    const float v = rstd * (g - m1 - (xr[c] - mu) * rstd * m2);
    o[c] = accumulate ? o[c] + v : v;
  }
  if (lane == 0) {
    stats[2 * r] = mu;
    stats[2 * r + 1] = rstd;
  }
}

// ------------------------------------------------------------------ element-wise
// dy = dx * leaky_relu'(y) with sign(y) recovered from the residual update: x_out - x_in = leaky_relu(y)
__global__ void leaky_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ x_out,
                                 const float* __restrict__ x_in, float* __restrict__ dy, long n) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    dy[i] = (x_out[i] - x_in[i]) > 0.f ? dx[i] : 0.01f * dx[i];
}

// FeedForward gate backward (healnet.py:323-331, 344-346): h = [a | g] (bias not yet added), hid = a * act(g);
// in place: h <- [dhid * act(g) | dhid * a * act'(g)]
__global__ void gate_bwd_kernel(float* __restrict__ h, const float* __restrict__ b1, const float* __restrict__ dhid,
                                long rows, int F, int snn) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const long n = rows * F;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / F;
    const int j = static_cast<int>(i - r * F);
    float* hr = h + r * 2 * F;
    const float a = hr[j] + b1[j], g = hr[F + j] + b1[F + j], d = dhid[i];
    float act, dact;
    if (snn) {
      const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
      const float e = expf(g);
      act = scale * (g > 0.f ? g : alpha * (e - 1.f));
      dact = scale * (g > 0.f ? 1.f : alpha * e);
    } else {
      const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752440f));
      act = g * cdf;
      dact = cdf + g * 0.3989422804014327f * expf(-0.5f * g * g);
    }
    hr[j] = d * act;
    hr[F + j] = d * a * dact;
  }
}

// out[i] (+)= sum_b in[b * n + i]   (gradient of the broadcast latent array, healnet.py:225)
__global__ void batch_sum_kernel(const float* __restrict__ in, float* __restrict__ out, long n, int batch) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < batch; ++b) s += in[static_cast<long>(b) * n + i];
    out[i] += s;
  }
}

// ------------------------------------------------------------------ classifier head backward
// logits = LN(mean_l x) W^T + bias (healnet.py:181-185). One block walks the batch (deterministic accumulation).
// dpooled[b][d] = gradient w.r.t. mean_l x; the caller spreads it as dx[b][l][d] = dpooled[b][d] / L.
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ x, int batch, int L, int D,
                                                       const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                       const float* __restrict__ W, int out_dims,
                                                       const float* __restrict__ dlogits, float* __restrict__ g_ln_w,
                                                       float* __restrict__ g_ln_b, float* __restrict__ g_W,
                                                       float* __restrict__ g_bias, float* __restrict__ dpooled) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  extern __shared__ float sm[];
  float* pooled = sm;            // [D]
  float* dpn = sm + D;           // [D]
  __shared__ float red[32];
  __shared__ float bc[4];
  const int tid = threadIdx.x;
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float t = 0.f;
    if (tid < 32) {
      t = tid < (blockDim.x >> 5) ? red[tid] : 0.f;
      t = warp_sum(t);
      if (tid == 0) bc[0] = t;
    }
    __syncthreads();
    t = bc[0];
    __syncthreads();
    return t;
  };
  for (int b = 0; b < batch; ++b) {
    for (int d = tid; d < D; d += blockDim.x) {
      float s = 0.f;
      const float* xb = x + static_cast<long>(b) * L * D + d;
      for (int l = 0; l < L; ++l) s += xb[static_cast<long>(l) * D];
      pooled[d] = s / L;
    }
    __syncthreads();
    float s = 0.f;
    for (int d = tid; d < D; d += blockDim.x) s += pooled[d];
    const float mu = block_sum(s) / D;
    float q = 0.f;
    for (int d = tid; d < D; d += blockDim.x) q += (pooled[d] - mu) * (pooled[d] - mu);
    const float rstd = rsqrtf(block_sum(q) / D + LN_EPS);
    // dpn = dlogits W ; parameter gradients of the Linear and the LayerNorm affine
    float m1 = 0.f, m2 = 0.f;
    for (int d = tid; d < D; d += blockDim.x) {
      const float xhat = (pooled[d] - mu) * rstd;
      const float pn = xhat * ln_w[d] + ln_b[d];
      float g = 0.f;
      for (int o = 0; o < out_dims; ++o) {
        const float dl = dlogits[b * out_dims + o];
        g += dl * W[static_cast<long>(o) * D + d];
        g_W[static_cast<long>(o) * D + d] += dl * pn;
      }
      g_ln_w[d] += g * xhat;
      g_ln_b[d] += g;
      const float gh = g * ln_w[d];
      dpn[d] = gh;
      m1 += gh;
      m2 += gh * xhat;
    }
    if (tid < out_dims) g_bias[tid] += dlogits[b * out_dims + tid];
    m1 = block_sum(m1) / D;
    m2 = block_sum(m2) / D;
    for (int d = tid; d < D; d += blockDim.x) {
      const float xhat = (pooled[d] - mu) * rstd;
      dpooled[static_cast<long>(b) * D + d] = rstd * (dpn[d] - m1 - xhat * m2);
    }
    __syncthreads();
  }
}
__global__ void spread_pooled_kernel(const float* __restrict__ dpooled, float* __restrict__ dx, int batch, int L, int D) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const long n = static_cast<long>(batch) * L * D;
  const float inv = 1.f / L;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long b = i / (static_cast<long>(L) * D);
    dx[i] = dpooled[b * D + i % D] * inv;
  }
}

// ------------------------------------------------------------------ attention row statistics (training forward)
// merged over the token splits: stats[(b*H + h)*L + l] = (M, den): softmax weight of token t = 2^(s_t - M) / den
__global__ void row_stats_kernel(const float* __restrict__ part_acc, const float* __restrict__ part_ml, int batch,
                                 int nsplit, int H, int L, int w, int den_col, float den_scale,
                                 float* __restrict__ stats) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const long total = static_cast<long>(batch) * H * L;
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int l = static_cast<int>(i % L), h = static_cast<int>((i / L) % H), b = static_cast<int>(i / (static_cast<long>(L) * H));
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, part_ml[((((static_cast<long>(b) * nsplit + s) * H + h) * L) + l) * 2]);
  float den = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const long base = (((static_cast<long>(b) * nsplit + s) * H + h) * L) + l;
    const float m = part_ml[base * 2];
    const float wgt = (m == -INFINITY) ? 0.f : exp2f(m - M);
    den += wgt * (den_col >= 0 ? part_acc[base * w + den_col] : part_ml[base * 2 + 1]);
  }
  stats[2 * i] = M;
  stats[2 * i + 1] = den * den_scale;
}

// ------------------------------------------------------------------ generic attention backward (materialised)
// S [bh][L][N] (log2 units) -> P = 2^(S - M) / sum_n 2^(S - M) in place (one warp per row); masked tokens -> 0.
// M (the forward's reference max) only keeps the exponentials in range; the row is normalised by ITS OWN sum rather
// than the forward's denominator, so that the recomputed probabilities sum to one exactly (a one-token axis gives
// P = 1 and, downstream, dt = 0 exactly — see softmax_bwd_kernel).
__global__ void __launch_bounds__(256) softmax_recompute_kernel(float* __restrict__ S, const float* __restrict__ stats,
                                                                int H, int L, long N,
                                                                const uint64_t* __restrict__ mask_bits, long n_rows) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);  // (b*H + h)*L + l
  if (row >= n_rows) return;
  const long words = (N + 63) / 64;
  const uint64_t* mb = mask_bits != nullptr ? mask_bits + (row / (static_cast<long>(H) * L)) * words : nullptr;
  float* s = S + row * N;
  const float M = stats[2 * row];
  float sum = 0.f;
  for (long n = lane; n < N; n += 32) {
    const bool keep = mb == nullptr || ((mb[n / 64] >> (n % 64)) & 1ull);
    const float e = keep ? exp2f(s[n] - M) : 0.f;
    s[n] = e;
    sum += e;
  }
  const float inv = 1.f / warp_sum(sum);
  for (long n = lane; n < N; n += 32) s[n] *= inv;
}
// dP [bh][L][N] -> dt = P * (dP - D) in place, D = sum_n P dP of the row (one warp per row). D is formed from the
// very P and dP it is subtracted from (not as dO . O): a single-token axis (the tabular modality) then gives dt = 0
// EXACTLY, as autograd does for a softmax over one element — Adam would turn a 1e-9 residue into full-size steps.
__global__ void __launch_bounds__(256) softmax_bwd_kernel(float* __restrict__ dP, const float* __restrict__ P, long N,
                                                          long n_rows) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const float* p = P + row * N;
  float* d = dP + row * N;
  float s = 0.f;
  for (long n = lane; n < N; n += 32) s += p[n] * d[n];
  s = warp_sum(s);
  for (long n = lane; n < N; n += 32) d[n] = p[n] * (d[n] - s);
}

// ------------------------------------------------------------------ small-context path helpers on [R = rows*H][C]
// u (split fp16 [rows][ld], head pitch zw) -> u32 [R][C], cnu = gamma * u + beta
__global__ void small_pre_kernel(const __half* __restrict__ u, int ld, int lo_off, int zw, int H, int C, long rows,
                                 const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float* __restrict__ u32, float* __restrict__ cnu) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const long n = rows * H * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long rh = i / C;
    const long r = rh / H;
    const int h = static_cast<int>(rh % H);
    const float v = ld_split(u, r * ld + h * zw + c, lo_off);
    u32[i] = v;
    cnu[i] = gamma[c] * v + beta[c];
  }
}
// du = gamma * g ; delta[R] = sum_c du * u   (one warp per R)
__global__ void small_du_kernel(const float* __restrict__ g, const float* __restrict__ u32,
                                const float* __restrict__ gamma, int C, long R, float* __restrict__ du,
                                float* __restrict__ delta) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long wid = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= R) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = gamma[c] * g[wid * C + c];
    du[wid * C + c] = d;
    s += d * u32[wid * C + c];
  }
  s = warp_sum(s);
  if (lane == 0) delta[wid] = s;
}
// out = scale * gamma[c] * in   (r_log2 = log2e * c * gamma * w ;  dw = c * gamma * dr)
__global__ void scale_cols_kernel(const float* __restrict__ in, const float* __restrict__ gamma, float scale, int C,
                                  long n, float* __restrict__ out) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = scale * gamma[i % C] * in[i];
}

// K/V weight gradient through the folded context LayerNorm: W' = W * gamma (per column), bias_v = Wv beta.
//   gW[i][c] += dWp[i][c] * gamma[c] (+ sv[i] * beta[c] on the V rows) ; ggamma[c] += sum_i dWp[i][c] * W[i][c] ;
//   gbeta[c] += sum_{i in V} sv[i] * W[i][c].          One block per 32 columns, deterministic.
__global__ void __launch_bounds__(256) kv_fold_bwd_kernel(const float* __restrict__ dWp, const float* __restrict__ W,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          const float* __restrict__ sv, int rows2I, int I, int C,
                                                          float* __restrict__ gW, float* __restrict__ ggamma,
                                                          float* __restrict__ gbeta) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  __shared__ float red[2][8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float sg = 0.f, sb = 0.f;
  if (c < C) {
    for (int i = w; i < rows2I; i += 8) {
      const long idx = static_cast<long>(i) * C + c;
      const float d = dWp[idx], wv = W[idx];
      float g = d * gamma[c];
      sg += d * wv;
      if (sv != nullptr && i >= I) {
        g += sv[i] * beta[c];
        sb += sv[i] * wv;
      }
      gW[idx] += g;
    }
  }
  red[0][w][lane] = sg;
  red[1][w][lane] = sb;
  __syncthreads();
  if (w == 0 && c < C) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a += red[0][i][lane];
      b += red[1][i][lane];
    }
    ggamma[c] += a;
    if (sv != nullptr) gbeta[c] += b;
  }
}

// ------------------------------------------------------------------ small-context streaming attention backward (SIMT)
// For one (sample, head), 64 latent rows and a range of tokens: recompute p = 2^(r.z - M) / den, form
// dt = p * (du.z - delta) and accumulate dr[l][c] += dt * z[c]  (no gradient flows into the context rows z, so this is
// the whole backward of the streaming pass; healnet.py:409-424 in the reassociated form of xattn_small.cu).
// Thread = (row, token quarter); r / du live in shared memory ([c][row]: conflict-free), the z tile as [c][token] so a
// thread reads four tokens of one column with one broadcast float4; dr accumulates in registers.
template <int CW>
__global__ void __launch_bounds__(256) small_attn_bwd_kernel(SmallBwdArgs p) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  constexpr int TT = 64;
  extern __shared__ __align__(16) float sm_dyn[];
  float (*rs)[64] = reinterpret_cast<float (*)[64]>(sm_dyn);
  float (*dus)[64] = reinterpret_cast<float (*)[64]>(sm_dyn + CW * 64);
  float (*zs)[TT] = reinterpret_cast<float (*)[TT]>(sm_dyn + 2 * CW * 64);
  __shared__ uint64_t mbits;
  const int tid = threadIdx.x, row = tid & 63, tq = tid >> 6;
  const int bh = blockIdx.z, b = bh / p.H, h = bh % p.H;
  const int l0 = blockIdx.y * 64;
  const int split = blockIdx.x;
  const long tiles = (p.N + TT - 1) / TT;
  const long t_begin = tiles * split / p.nsplit, t_end = tiles * (split + 1) / p.nsplit;
  const int l = l0 + row;
  const bool row_ok = l < p.L;
  const long R = (static_cast<long>(b) * p.L + (row_ok ? l : 0)) * p.H + h;   // index into [rows][H] arrays
  for (int i = tid; i < CW * 64; i += 256) {
    const int c = i >> 6, rr = i & 63;
    const long Rr = (static_cast<long>(b) * p.L + l0 + rr) * p.H + h;
    const bool ok = c < p.C && l0 + rr < p.L;
    rs[c][rr] = ok ? p.r[Rr * p.C + c] : 0.f;
    dus[c][rr] = ok ? p.du[Rr * p.C + c] : 0.f;
  }
  const float M = row_ok ? p.stats[2 * ((static_cast<long>(b) * p.H + h) * p.L + l)] : 0.f;
  const float inv_den = row_ok ? 1.f / p.stats[2 * ((static_cast<long>(b) * p.H + h) * p.L + l) + 1] : 0.f;
  const float delta = row_ok ? p.delta[R] : 0.f;
  float acc[CW];
#pragma unroll
  for (int c = 0; c < CW; ++c) acc[c] = 0.f;
  for (long t = t_begin; t < t_end; ++t) {
    __syncthreads();
    // z tile: 64 tokens x C columns, fp32 = hi + lo, transposed into [c][token]
    for (int i = tid; i < TT * CW; i += 256) {
      const int c = i % CW, tk = i / CW;
      const long n = t * TT + tk;
      float v = 0.f;
      if (n < p.N && c < p.C) {
        const __half* zr = p.z + (static_cast<long>(b) * p.N + n) * p.z_ld;
        v = __half2float(zr[c]) + __half2float(zr[p.z_lo + c]);
      }
      zs[c][tk] = v;
    }
    if (tid == 0) {
      uint64_t bits = ~0ull;
      if (p.mask_bits != nullptr) bits = p.mask_bits[static_cast<long>(b) * tiles + t];
      const long rem = p.N - t * TT;
      if (rem < TT) bits &= (1ull << rem) - 1ull;
      mbits = bits;
    }
    __syncthreads();
    const uint64_t bits = mbits;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
      const int tk = tq * 16 + q4 * 4;
      float s[4] = {0.f, 0.f, 0.f, 0.f}, g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const float4 zv = *reinterpret_cast<const float4*>(&zs[c][tk]);
        const float rv = rs[c][row], dv = dus[c][row];
        s[0] = fmaf(rv, zv.x, s[0]); s[1] = fmaf(rv, zv.y, s[1]); s[2] = fmaf(rv, zv.z, s[2]); s[3] = fmaf(rv, zv.w, s[3]);
        g[0] = fmaf(dv, zv.x, g[0]); g[1] = fmaf(dv, zv.y, g[1]); g[2] = fmaf(dv, zv.z, g[2]); g[3] = fmaf(dv, zv.w, g[3]);
      }
      float dt[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool keep = (bits >> (tk + j)) & 1ull;
        dt[j] = keep ? exp2f(s[j] - M) * inv_den * (g[j] - delta) : 0.f;
      }
#pragma unroll
      for (int c = 0; c < CW; ++c) {
        const float4 zv = *reinterpret_cast<const float4*>(&zs[c][tk]);
        acc[c] = fmaf(dt[0], zv.x, fmaf(dt[1], zv.y, fmaf(dt[2], zv.z, fmaf(dt[3], zv.w, acc[c]))));
      }
    }
  }
  // reduce the four token quarters of every row through shared memory (fixed order), one partial per split
  __syncthreads();
  float* red = &rs[0][0];  // CW*64 floats, reused
  for (int q = 1; q < 4; ++q) {
    if (tq == q) {
#pragma unroll
      for (int c = 0; c < CW; ++c) red[c * 64 + row] = acc[c];
    }
    __syncthreads();
    if (tq == 0) {
#pragma unroll
      for (int c = 0; c < CW; ++c) acc[c] += red[c * 64 + row];
    }
    __syncthreads();
  }
  if (tq == 0 && row_ok) {
    float* o = p.dr_part + (static_cast<long>(split) * p.R_total + R) * p.C;
    for (int c = 0; c < p.C; ++c) o[c] = acc[c];
  }
}
__global__ void sum_splits_kernel(const float* __restrict__ part, int nsplit, long n, float* __restrict__ out) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[static_cast<long>(k) * n + i];
    out[i] = s;
  }
}

// ------------------------------------------------------------------ bf16 hi/lo operand packing for the backward GEMMs
// dst rows = [hi (seg cols) | lo (seg cols)], hi = bf16(v), lo = bf16(v - hi) (16 significant bits, fp32 range), pad
// columns zero. src: fp32 (type 0) or the forward's split fp16 rows (type 1: value = hi + lo at + lo_off).
__device__ __forceinline__ float ld_src(const void* src, int src_half, long idx, int lo_off) {
  if (!src_half) return static_cast<const float*>(src)[idx];
  const __half* hp = static_cast<const __half*>(src);
  return __half2float(hp[idx]) + __half2float(hp[idx + lo_off]);
}
__device__ __forceinline__ void st_bf16_split(__nv_bfloat16* dst, long idx, int seg, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  dst[idx] = hi;
  dst[idx + seg] = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// Strided, batched operands of every backward product: dst[(z R + r) 2 seg + c] = hi, [+ seg] = lo
// of src(z, r, c) = p[b1 s_b1 + b2 s_b2 + r s_r + c s_c], z = b1 nb2 + b2; columns [C, seg) zero. 32 x 32 tiles:
// read along whichever of (r, c) is contiguous in the source, always write along c.
struct PackSrc {
  const void* p;
  int type, lo_off;
  long s_r, s_c, s_b1, s_b2;
};
// One launch packs BOTH operands of a product: blocks [0, blocks_a) work on operand a, the rest on operand b.
struct PackJob {
  PackSrc s;
  long R;
  int C, seg;
  __nv_bfloat16* dst;
  int tiles_c, tiles_r;   // 32 x 32 tiles per batch entry
};
__global__ void __launch_bounds__(256) pack_bf16_strided_kernel(PackJob ja, PackJob jb, long blocks_a, int nb2) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  __shared__ float tile[32][33];
  const bool second = static_cast<long>(blockIdx.x) >= blocks_a;
  const PackJob& j = second ? jb : ja;
  long blk = static_cast<long>(blockIdx.x) - (second ? blocks_a : 0);
  const PackSrc& s = j.s;
  const long R = j.R;
  const int C = j.C, seg = j.seg;
  __nv_bfloat16* __restrict__ dst = j.dst;
  const int tc = static_cast<int>(blk % j.tiles_c);
  blk /= j.tiles_c;
  const int tr = static_cast<int>(blk % j.tiles_r);
  const int z = static_cast<int>(blk / j.tiles_r);
  const long base = static_cast<long>(z / nb2) * s.s_b1 + static_cast<long>(z % nb2) * s.s_b2;
  const long r0 = static_cast<long>(tr) * 32;
  const int c0 = tc * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const bool r_fast = s.s_r == 1 && s.s_c != 1;
  for (int j = ty; j < 32; j += 8) {
    const long r = r_fast ? r0 + tx : r0 + j;
    const int c = r_fast ? c0 + j : c0 + tx;
    float v = 0.f;
    if (r < R && c < C) v = ld_src(s.p, s.type, base + r * s.s_r + c * s.s_c, s.lo_off);
    if (r_fast) tile[tx][j] = v; else tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const long r = r0 + j;
    const int c = c0 + tx;
    if (r < R && c < seg) st_bf16_split(dst, (static_cast<long>(z) * R + r) * 2 * seg + c, seg, tile[j][tx]);
  }
}

// FeedForward gate backward on the recomputed pre-activations in the forward GEMM's interleaved layout:
// h_il[r][2j] = a_j, h_il[r][2j + 1] = g_j (bias included) -> dh[r][j] = d a_j, dh[r][F + j] = d g_j
__global__ void gate_bwd_il_kernel(const float* __restrict__ h_il, const float* __restrict__ dhid,
                                   float* __restrict__ dh, long rows, int F, int snn) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const long n = rows * F;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long r = i / F;
    const int j = static_cast<int>(i - r * F);
    const float2 ag = *reinterpret_cast<const float2*>(h_il + r * 2 * F + 2 * j);
    const float a = ag.x, g = ag.y, d = dhid[i];
    float act, dact;
    if (snn) {
      const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
      const float e = expf(g);
      act = scale * (g > 0.f ? g : alpha * (e - 1.f));
      dact = scale * (g > 0.f ? 1.f : alpha * e);
    } else {
      const float cdf = 0.5f * (1.f + erff(g * 0.70710678118654752440f));
      act = g * cdf;
      dact = cdf + g * 0.3989422804014327f * expf(-0.5f * g * g);
    }
    dh[r * 2 * F + j] = d * act;
    dh[r * 2 * F + F + j] = d * a * dact;
  }
}

// operands of the tensor-core streaming backward (xattn_small.cu): one warp per (row, head)
__global__ void __launch_bounds__(256) small_bwd_prep_kernel(const float* __restrict__ r, const float* __restrict__ du,
                                                             const float* __restrict__ delta,
                                                             const float* __restrict__ stats, int batch, int H, int L,
                                                             int C, int kd, __half* __restrict__ rq,
                                                             __half* __restrict__ duq, int rq_ld, int lo_off,
                                                             float* __restrict__ row_a, float* __restrict__ row_d,
                                                             float* __restrict__ scale) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long R = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);   // (b*L + l)*H + h
  if (R >= static_cast<long>(batch) * L * H) return;
  const int h = static_cast<int>(R % H);
  const long bl = R / H;
  const int l = static_cast<int>(bl % L), b = static_cast<int>(bl / L);
  float mx = 0.f;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, fabsf(du[R * C + c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float s = (mx > 0.f && mx < INFINITY) ? exp2f(-static_cast<float>(ilogbf(mx))) : 1.f;
  const long st = (static_cast<long>(b) * H + h) * L + l;
  const float M = stats[2 * st], den = stats[2 * st + 1];
  const bool dead = !(den > 0.f) || M == -INFINITY;   // fully masked row
  __half* rr = rq + bl * rq_ld + h * kd;
  __half* dr = duq + bl * rq_ld + h * kd;
  for (int c = lane; c < kd; c += 32) {
    float rv = 0.f, dv = 0.f;
    if (c < C) {
      rv = r[R * C + c];
      dv = du[R * C + c] * s;
    } else if (c == C) {
      rv = dead ? 0.f : 10.f - M;   // P_SHIFT - M: exactly the fp16 value the forward folded into Q'
    }
    const __half rh = __float2half_rn(rv), dh = __float2half_rn(dv);
    const __half rl = __float2half_rn(rv - __half2float(rh));
    rr[c] = rh;
    dr[c] = dh;
    dr[lo_off + c] = __float2half_rn(dv - __half2float(dh));
    if (kd == 32 && C >= 17 && C <= 23) {
      // merged tail of the score product (xattn_small.cu): lo half = [R_lo 0..15 | R_hi 16..C-1 | 0 | R_lo 16..C-1 | 0..]
      if (c < 16) {
        rr[lo_off + c] = rl;
      } else if (c < C) {
        rr[lo_off + c] = rh;
        rr[lo_off + C + 1 + (c - 16)] = rl;
      } else if (c == C || c > 2 * C - 16) {
        rr[lo_off + c] = __float2half_rn(0.f);
      }
    } else {
      rr[lo_off + c] = rl;
    }
  }
  if (lane == 0) {
    row_a[R] = dead ? 0.f : 0.0009765625f / den;
    row_d[R] = delta[R] * s;
    scale[R] = s;
  }
}
__global__ void small_bwd_finish_kernel(const float* __restrict__ part, const float* __restrict__ scale, int batch,
                                        int nsplit, int H, int L, int C, int kd, float* __restrict__ dr) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const long n = static_cast<long>(batch) * L * H * C;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long R = i / C;
    const int h = static_cast<int>(R % H);
    const long bl = R / H;
    const int l = static_cast<int>(bl % L), b = static_cast<int>(bl / L);
    float s = 0.f;
    for (int k = 0; k < nsplit; ++k) s += part[((((static_cast<long>(b) * nsplit + k) * H + h) * L) + l) * kd + c];
    dr[i] = s / scale[R];
  }
}

inline unsigned ew_grid(long n) {
  const long g = (n + 255) / 256;
  return static_cast<unsigned>(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace

// ================================================================== launchers
int colsum_slabs(long R) {
  long s = (R + 255) / 256;
  return static_cast<int>(s < 1 ? 1 : (s > 64 ? 64 : s));
}
int launch_colsum(int mode, const float* a, long lda, const float* b, long ldb, const float* stats, long R, int C,
                  float scale, float* out, int accumulate, float* partial, cudaStream_t st) {
  if (C <= 0) return 0;
  const int slabs = colsum_slabs(R);
  const long per = (R + slabs - 1) / slabs;
  const dim3 grid((C + 31) / 32, slabs);
  if (mode == 0)
    HN_CHECK_CUDA(launch_k(colsum_partial_kernel<0>, grid, dim3(256), 0, st, a, lda, b, ldb, stats, R, C, per, partial));
  else if (mode == 1)
    HN_CHECK_CUDA(launch_k(colsum_partial_kernel<1>, grid, dim3(256), 0, st, a, lda, b, ldb, stats, R, C, per, partial));
  else
    HN_CHECK_CUDA(launch_k(colsum_partial_kernel<2>, grid, dim3(256), 0, st, a, lda, b, ldb, stats, R, C, per, partial));
  HN_CHECK_CUDA(launch_k(colsum_final_kernel, dim3((C + 127) / 128), dim3(128), 0, st, partial, slabs, C, scale, out, accumulate));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_ln_bwd_rows(const float* x, const float* dy, const float* gamma, float* dx, float* stats, long rows, int D,
                       int accumulate, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(ln_bwd_rows_kernel, dim3(static_cast<unsigned>((rows + 7) / 8)), dim3(256), 0, st, x, dy, gamma, dx,
                         stats, rows, D, accumulate));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_leaky_bwd(const float* dx, const float* x_out, const float* x_in, float* dy, long n, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(leaky_bwd_kernel, dim3(ew_grid(n)), dim3(256), 0, st, dx, x_out, x_in, dy, n));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_gate_bwd(float* h, const float* b1, const float* dhid, long rows, int F, int snn, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(gate_bwd_kernel, dim3(ew_grid(rows * F)), dim3(256), 0, st, h, b1, dhid, rows, F, snn));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_batch_sum(const float* in, float* out, long n, int batch, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(batch_sum_kernel, dim3(ew_grid(n)), dim3(256), 0, st, in, out, n, batch));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_head_bwd(const float* x, int batch, int L, int D, const float* ln_w, const float* ln_b, const float* W,
                    int out_dims, const float* dlogits, float* g_ln_w, float* g_ln_b, float* g_W, float* g_bias,
                    float* dpooled, float* dx, cudaStream_t st) {
  HN_REQUIRE(out_dims <= 256, "head backward: at most 256 outputs");
  HN_CHECK_CUDA(launch_k(head_bwd_kernel, dim3(1), dim3(256), 2 * D * sizeof(float), st, x, batch, L, D, ln_w, ln_b, W,
                         out_dims, dlogits, g_ln_w, g_ln_b, g_W, g_bias, dpooled));
  HN_CHECK_CUDA(launch_k(spread_pooled_kernel, dim3(ew_grid(static_cast<long>(batch) * L * D)), dim3(256), 0, st, dpooled,
                         dx, batch, L, D));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_row_stats(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L, int w,
                     int den_col, float den_scale, float* stats, cudaStream_t st) {
  const long total = static_cast<long>(batch) * H * L;
  HN_CHECK_CUDA(launch_k(row_stats_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, st, part_acc,
                         part_ml, batch, nsplit, H, L, w, den_col, den_scale, stats));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_softmax_recompute(float* S, const float* stats, int H, int L, long N, const uint64_t* mask_bits, long total,
                             cudaStream_t st) {  // total = number of rows (b * H * L)
  HN_CHECK_CUDA(launch_k(softmax_recompute_kernel, dim3(static_cast<unsigned>((total + 7) / 8)), dim3(256), 0, st, S, stats, H, L,
                         N, mask_bits, total));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_softmax_bwd(float* dP, const float* P, long N, long n_rows, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(softmax_bwd_kernel, dim3(static_cast<unsigned>((n_rows + 7) / 8)), dim3(256), 0, st, dP, P, N, n_rows));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_small_pre(const __half* u, int ld, int lo_off, int zw, int H, int C, long rows, const float* gamma,
                     const float* beta, float* u32, float* cnu, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(small_pre_kernel, dim3(ew_grid(rows * H * C)), dim3(256), 0, st, u, ld, lo_off, zw, H, C, rows, gamma,
                         beta, u32, cnu));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_small_du(const float* g, const float* u32, const float* gamma, int C, long R, float* du, float* delta,
                    cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(small_du_kernel, dim3(static_cast<unsigned>((R + 7) / 8)), dim3(256), 0, st, g, u32, gamma, C, R, du, delta));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_scale_cols(const float* in, const float* gamma, float scale, int C, long n, float* out, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(scale_cols_kernel, dim3(ew_grid(n)), dim3(256), 0, st, in, gamma, scale, C, n, out));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_kv_fold_bwd(const float* dWp, const float* W, const float* gamma, const float* beta, const float* sv,
                       int rows2I, int I, int C, float* gW, float* ggamma, float* gbeta, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(kv_fold_bwd_kernel, dim3((C + 31) / 32), dim3(256), 0, st, dWp, W, gamma, beta, sv, rows2I, I, C, gW,
                         ggamma, gbeta));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_pack_bf16_pair(const SgOperand& A, long RA, const SgOperand& B_as_rows, long RB, int K, int nb1, int nb2,
                          void* dstA, void* dstB, int seg, cudaStream_t st) {
  HN_REQUIRE(seg >= K && nb1 >= 1 && nb2 >= 1, "pack_bf16_pair: bad shape");
  const long nz = static_cast<long>(nb1) * nb2;
  auto job = [&](const SgOperand& o, long R, void* dst) {
    PackJob j;
    j.s = PackSrc{o.p, o.type, o.lo_off, o.s_row, o.s_col, o.s_b1, o.s_b2};
    j.R = R;
    j.C = K;
    j.seg = seg;
    j.dst = static_cast<__nv_bfloat16*>(dst);
    j.tiles_c = (seg + 31) / 32;
    j.tiles_r = static_cast<int>((R + 31) / 32);
    return j;
  };
  const PackJob ja = job(A, RA, dstA), jb = job(B_as_rows, RB, dstB);
  const long ba = static_cast<long>(ja.tiles_c) * ja.tiles_r * nz, bb = static_cast<long>(jb.tiles_c) * jb.tiles_r * nz;
  HN_REQUIRE(ba + bb < 2147483647L, "pack_bf16_pair: grid too large");
  HN_CHECK_CUDA(launch_k(pack_bf16_strided_kernel, dim3(static_cast<unsigned>(ba + bb)), dim3(256), 0, st, ja, jb, ba, nb2));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_gate_bwd_il(const float* h_il, const float* dhid, float* dh, long rows, int F, int snn, cudaStream_t st) {
  HN_CHECK_CUDA(launch_k(gate_bwd_il_kernel, dim3(ew_grid(rows * F)), dim3(256), 0, st, h_il, dhid, dh, rows, F, snn));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_small_bwd_prep(const float* r, const float* du, const float* delta, const float* stats, int batch, int H,
                          int L, int C, int kd, __half* rq, __half* duq, int rq_ld, int lo_off, float* row_a,
                          float* row_d, float* scale, cudaStream_t st) {
  const long R = static_cast<long>(batch) * L * H;
  HN_CHECK_CUDA(launch_k(small_bwd_prep_kernel, dim3(static_cast<unsigned>((R + 7) / 8)), dim3(256), 0, st, r, du, delta, stats,
                         batch, H, L, C, kd, rq, duq, rq_ld, lo_off, row_a, row_d, scale));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int launch_small_bwd_finish(const float* part, const float* scale, int batch, int nsplit, int H, int L, int C, int kd,
                            float* dr, cudaStream_t st) {
  const long n = static_cast<long>(batch) * L * H * C;
  HN_CHECK_CUDA(launch_k(small_bwd_finish_kernel, dim3(ew_grid(n)), dim3(256), 0, st, part, scale, batch, nsplit, H, L, C, kd, dr));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int small_attn_bwd_nsplit(int batch, int H, int L, long N) {
  const long base = static_cast<long>(batch) * H * ((L + 63) / 64);
  const long tiles = (N + 63) / 64;
  long s = (148 * 4 + base - 1) / base;
  if (s > tiles / 4) s = tiles / 4;
  if (s < 1) s = 1;
  if (s > 512) s = 512;
  return static_cast<int>(s);
}
int launch_small_attn_bwd(const SmallBwdArgs& a, float* dr, cudaStream_t st) {
  HN_REQUIRE(a.C >= 1 && a.C <= 64 && a.nsplit >= 1, "small attention backward: context width must be <= 64");
  const dim3 grid(a.nsplit, (a.L + 63) / 64, a.batch * a.H);
  if (a.C <= 32) {
    HN_CHECK_CUDA(launch_k(small_attn_bwd_kernel<32>, grid, dim3(256), 3 * 32 * 64 * sizeof(float), st, a));
  } else {
    constexpr int SMEM = 3 * 64 * 64 * sizeof(float);
    HN_CHECK_CUDA(cudaFuncSetAttribute(small_attn_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    HN_CHECK_CUDA(launch_k(small_attn_bwd_kernel<64>, grid, dim3(256), SMEM, st, a));
  }
  const long n = a.R_total * a.C;
  HN_CHECK_CUDA(launch_k(sum_splits_kernel, dim3(ew_grid(n)), dim3(256), 0, st, a.dr_part, a.nsplit, n, dr));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hn
