// rowops.cu — the HBM-bound row kernels around the tensor-core work:
//   * LayerNorm of latent rows -> fp16 GEMM operands                    (PreNorm.norm, healnet.py:313-314)
//   * Fourier positional tables + standardised context rows "z"         (healnet.py:211-221, 292-302, 318)
//   * split-N partial combine (+ fused V projection on the small-C path)
//   * mean-pool -> LayerNorm -> Linear head                             (healnet.py:181-185)
// LayerNorm of the context is split as LN(c) = gamma * z + beta with z = (c - mean) * rstd: z depends only
// on the input, so it is built ONCE per forward and shared by all `depth` layers; gamma/beta are folded
// into the projection weights at pack time (pack.cu).
#include <cstdlib>

#include <cuda_bf16.h>

#include "common.cuh"

namespace hn {

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("HN_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

namespace {

constexpr float LN_EPS = 1e-5f;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------ LayerNorm rows -> split fp16
// y[r] = [hi (seg cols) | lo (seg cols)] with hi = fp16(v), lo = fp16(v - hi), v = LN(x[r]) * gamma + beta;
// pad columns [D, seg) of both segments are zero. lo_seg == 0: single fp16 row of `seg` columns.
__device__ __forceinline__ void store_split(__half* yr, int c, int seg, int lo_seg, float v) {
  const __half hi = __float2half_rn(v);
  yr[c] = hi;
  if (lo_seg > 0) yr[lo_seg + c] = __float2half_rn(v - __half2float(hi));
}
// one warp per row; row cached in registers (D <= 32*MAXV) else re-read from L2
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_f16_kernel(const float* __restrict__ x, int ldx,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, __half* __restrict__ y,
                                                            int ldy, int seg, int lo_seg, long rows, int D) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    v[i] = c < D ? xr[c] : 0.f;
    s += v[i];
  }
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    const float d = c < D ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / D + LN_EPS);
  __half* yr = y + row * ldy;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int c = lane + i * 32;
    if (c < seg) store_split(yr, c, seg, lo_seg, c < D ? (v[i] - mean) * rstd * gamma[c] + beta[c] : 0.f);
  }
}

__global__ void __launch_bounds__(256) layernorm_f16_big_kernel(const float* __restrict__ x, int ldx,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta,
                                                                __half* __restrict__ y, int ldy, int seg, int lo_seg,
                                                                long rows, int D) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long row = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* xr = x + row * ldx;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += xr[c];
  const float mean = warp_sum(s) / D;
  float q = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float d = xr[c] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / D + LN_EPS);
  __half* yr = y + row * ldy;
  for (int c = lane; c < seg; c += 32)
    store_split(yr, c, seg, lo_seg, c < D ? (xr[c] - mean) * rstd * gamma[c] + beta[c] : 0.f);
}

// ------------------------------------------------------------------ Fourier axis tables
// tab[(off_a + j) * (2B+1) + k]: k < B sin(pi p f_k), B <= k < 2B cos(pi p f_{k-B}), k == 2B: p
// p = linspace(-1, 1, size)[j] (size 1 -> -1), f_k = linspace(1, max_freq/2, B)[k]   (healnet.py:212, 292-302)
struct AxisInfo {
  int n_axes;
  int size[4];
  int off[4];
};
__global__ void axis_tables_kernel(float* __restrict__ tab, AxisInfo ax, int B, float max_freq) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int F = 2 * B + 1;
  int total = 0;
  for (int a = 0; a < ax.n_axes; ++a) total += ax.size[a];
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total * F; idx += gridDim.x * blockDim.x) {
    const int r = idx / F, k = idx % F;
    int a = 0;
    while (a + 1 < ax.n_axes && r >= ax.off[a + 1]) ++a;
    const int j = r - ax.off[a], n = ax.size[a];
    // torch.linspace fp32 semantics: step = (end-start)/(n-1); lower half start + j*step, upper half end - (n-1-j)*step
    float p;
    if (n == 1) {
      p = -1.f;
    } else {
      const float step = 2.f / static_cast<float>(n - 1);
      p = (j < n / 2) ? (-1.f + step * j) : (1.f - step * (n - 1 - j));
    }
    float out;
    if (k == 2 * B) {
      out = p;
    } else {
      const int kk = k < B ? k : k - B;
      float f;
      if (B == 1) {
        f = 1.f;
      } else {
        const float fstep = (max_freq * 0.5f - 1.f) / static_cast<float>(B - 1);
        f = (kk < B / 2) ? (1.f + fstep * kk) : (max_freq * 0.5f - fstep * (B - 1 - kk));
      }
      const float arg = p * f * 3.14159265358979323846f;  // fp32 product, as the reference computes it
      out = k < B ? static_cast<float>(sin(static_cast<double>(arg))) : static_cast<float>(cos(static_cast<double>(arg)));
    }
    tab[idx] = out;
  }
}

// raw modality elements as the caller holds them: fp32, bf16 or fp16 (hn_set_io_dtype) -> float
template <typename T>
__device__ __forceinline__ float in_f(const T* p) {
  return static_cast<float>(__ldg(p));
}
template <>
__device__ __forceinline__ float in_f<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(*p);
}
template <>
__device__ __forceinline__ float in_f<__half>(const __half* p) {
  return __half2float(*p);
}

// eight consecutive values -> one 16-byte store of their fp16 hi parts and (lo_dst != null) one of the lo parts
__device__ __forceinline__ void store_hi_lo8(uint4* hi_dst, uint4* lo_dst, const float* o) {
  uint4 w, wl;
  uint32_t* ph = reinterpret_cast<uint32_t*>(&w);
  uint32_t* pl = reinterpret_cast<uint32_t*>(&wl);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const __half2 h = __floats2half2_rn(o[2 * k], o[2 * k + 1]);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(o[2 * k] - hf.x, o[2 * k + 1] - hf.y);
    ph[k] = *reinterpret_cast<const uint32_t*>(&h);
    pl[k] = *reinterpret_cast<const uint32_t*>(&l);
  }
  *hi_dst = w;
  if (lo_dst != nullptr) *lo_dst = wl;
}

// ------------------------------------------------------------------ z rows, small-C layout (C <= ZW-1)
// one thread per token: raw channels + table features -> standardise -> ZW fp16 = one 64/128-byte row
// split != 0: rows are [hi (ZW) | lo (ZW)], lo = fp16(value - hi) (the streaming kernel's three-term score product)
template <int ZW, typename T>
__global__ void __launch_bounds__(256) build_z_small_kernel(const T* __restrict__ raw, __half* __restrict__ z,
                                                            long tokens_total, long N, int c_raw, AxisInfo ax,
                                                            int F, const float* __restrict__ tab, long tok0, int split) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const long t = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= tokens_total) return;
  const long n = t % N + tok0;
  float v[ZW];
  const T* r = raw + t * c_raw;
#pragma unroll
  for (int i = 0; i < ZW; ++i) v[i] = i < c_raw ? in_f(r + i) : 0.f;
  int C = c_raw;
  // row-major token index -> per-axis coordinates
  long rem = n;
  int coord[4];
  for (int a = ax.n_axes - 1; a >= 0; --a) {
    coord[a] = static_cast<int>(rem % ax.size[a]);
    rem /= ax.size[a];
  }
  if (F > 0) {
    for (int a = 0; a < ax.n_axes; ++a) {
      const float* tr = tab + static_cast<long>(ax.off[a] + coord[a]) * F;
      for (int k = 0; k < F; ++k) {
        const float tv = __ldg(tr + k);
#pragma unroll
        for (int i = 0; i < ZW; ++i)
          if (i == C + k) v[i] = tv;
      }
      C += F;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ZW; ++i)
    if (i < C) s += v[i];
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < ZW; ++i)
    if (i < C) {
      const float d = v[i] - mean;
      q += d * d;
    }
  const float rstd = rsqrtf(q / C + LN_EPS);
#pragma unroll
  for (int i = 0; i < ZW; ++i) {
    float val = 0.f;
    if (i < C) val = (v[i] - mean) * rstd;
    if (i == C) val = 1.f;  // ones column: the PV UMMA accumulates the softmax denominator for free
    v[i] = val;
  }
  uint4* dst = reinterpret_cast<uint4*>(z + t * (split ? 2 * ZW : ZW));
#pragma unroll
  for (int i = 0; i < ZW / 8; ++i) store_hi_lo8(dst + i, split ? dst + ZW / 8 + i : nullptr, v + 8 * i);
  if (ZW == 32 && split && C >= 17 && C <= 23) {
    // merged tail of the streaming kernel's score products (xattn_small.cu): the lo half also carries the hi parts of
    // columns 16..C-1, right behind its (zero) column C
    __half* lo = z + t * 2 * ZW + ZW;
#pragma unroll
    for (int i = 16; i < 24; ++i)
      if (i < C) lo[C + 1 + i - 16] = __float2half_rn(v[i]);
  }
}

// Specialisation for the shapes the path is run on (image / volume: 1-4 raw channels, 1-3 axes, the default 2
// frequency bands -> F = 5): every feature lands in a compile-time register slot, so the row costs ~150 instructions
// instead of the ~1000 of the generic kernel above (whose runtime column positions need a compare per slot).
template <int CRAW, int NAX, typename T>
__global__ void __launch_bounds__(256) build_z_small32_fast_kernel(const T* __restrict__ raw, __half* __restrict__ z,
                                                                   long tokens_total, long N, AxisInfo ax,
                                                                   const float* __restrict__ tab, long tok0, int split) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  constexpr int F = 5, ZW = 32, C = CRAW + NAX * F;
  static_assert(C < ZW, "context row must leave room for the ones column");
  const long t_raw = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long t = t_raw < tokens_total ? t_raw : tokens_total - 1;   // (tail threads recompute the last row: all reach the barrier)
  unsigned rem = (tokens_total < (1L << 31))
                     ? static_cast<unsigned>(t) % static_cast<unsigned>(N) + static_cast<unsigned>(tok0)
                     : static_cast<unsigned>(t % N + tok0);
  float v[C];
  const T* r = raw + t * CRAW;
#pragma unroll
  for (int i = 0; i < CRAW; ++i) v[i] = in_f(r + i);
#pragma unroll
  for (int a = NAX - 1; a >= 0; --a) {
    const unsigned sz = static_cast<unsigned>(ax.size[a]);
    const unsigned qd = rem / sz;
    const float* tr = tab + static_cast<long>(ax.off[a] + static_cast<int>(rem - qd * sz)) * F;
    rem = qd;
#pragma unroll
    for (int k = 0; k < F; ++k) v[CRAW + a * F + k] = __ldg(tr + k);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) s += v[i];
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(q / C + LN_EPS);
  // the row as packed fp16 words: hi half, and (split) lo half = fp16(value - hi), with the merged tail (17 <= C <= 23:
  // the hi parts of columns 16..C-1 again behind the lo half's zero column C, see xattn_small.cu)
  uint32_t hw[ZW / 2], lw[ZW / 2];
#pragma unroll
  for (int k = 0; k < ZW / 2; ++k) {
    float o[2], l[2];
    __half hh[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int i = 2 * k + e;
      o[e] = i < C ? (v[i < C ? i : 0] - mean) * rstd : (i == C ? 1.f : 0.f);
      hh[e] = __float2half_rn(o[e]);
      l[e] = o[e] - __half2float(hh[e]);
      if (C >= 17 && C <= 23 && i > C && i <= 2 * C - 16) l[e] = (v[(i - C - 1 + 16) < C ? (i - C - 1 + 16) : 0] - mean) * rstd;
    }
    const __half2 h2 = __halves2half2(hh[0], hh[1]);
    const __half2 l2 = __floats2half2_rn(l[0], l[1]);
    hw[k] = *reinterpret_cast<const uint32_t*>(&h2);
    lw[k] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  // Coalesced stores: a thread's row is 64 (128) contiguous bytes, so direct 16-byte stores touch 32 different lines per
  // warp instruction; the block's rows are contiguous in z, so they go through shared memory (16-byte chunks XOR-swizzled
  // by the row index: the minimum of four wavefronts per 512-byte warp access on both sides) and leave as 512-byte warp stores.
  constexpr int CH = ZW / 8;                       // 16-byte chunks per half
  const int nch = split ? 2 * CH : CH;              // chunks per row
  __shared__ uint4 stage[256 * 2 * CH];
  {
    const int r = threadIdx.x;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      stage[r * nch + (c ^ (r & (nch - 1)))] = make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]);
      if (split)
        stage[r * nch + ((CH + c) ^ (r & (nch - 1)))] = make_uint4(lw[4 * c], lw[4 * c + 1], lw[4 * c + 2], lw[4 * c + 3]);
    }
  }
  __syncthreads();
  const long row0 = static_cast<long>(blockIdx.x) * blockDim.x;
  const long rows_here = tokens_total - row0 < 256 ? tokens_total - row0 : 256;
  uint4* dst = reinterpret_cast<uint4*>(z) + row0 * nch;
  for (int idx = threadIdx.x; idx < rows_here * nch; idx += 256) {
    const int r = idx / nch, c = idx - r * nch;
    dst[idx] = stage[r * nch + (c ^ (r & (nch - 1)))];
  }
}

// ------------------------------------------------------------------ z rows, generic layout (any C)
// one warp per token row
template <typename T>
__global__ void __launch_bounds__(256) build_z_large_kernel(const T* __restrict__ raw, __half* __restrict__ z,
                                                            int ldz, int seg, int lo_seg, long tokens_total, long N,
                                                            int c_raw, AxisInfo ax, int F,
                                                            const float* __restrict__ tab, long tok0) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long t = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens_total) return;
  const long n = t % N + tok0;
  const T* r = raw + t * c_raw;
  const int n_feat = F * ax.n_axes;
  const int C = c_raw + n_feat;
  long rem = n;
  int coord[4] = {0, 0, 0, 0};
  for (int a = ax.n_axes - 1; a >= 0; --a) {
    coord[a] = static_cast<int>(rem % ax.size[a]);
    rem /= ax.size[a];
  }
  auto feat = [&](int c) -> float {
    if (c < c_raw) return in_f(r + c);
    const int k = c - c_raw;
    const int a = k / F;
    return __ldg(tab + static_cast<long>(ax.off[a] + coord[a]) * F + (k - a * F));
  };
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s += feat(c);
  const float mean = warp_sum(s) / C;
  float q = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float d = feat(c) - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) / C + LN_EPS);
  __half* zr = z + t * ldz;
  for (int c = lane; c < seg; c += 32) store_split(zr, c, seg, lo_seg, c < C ? (feat(c) - mean) * rstd : 0.f);
}

// Wide-context fast path (WSI patch features: c_raw a multiple of 4, up to 1024 raw channels, at most 32 positional
// features): the row is read ONCE with float4 loads and kept in registers for the two-pass statistics (the generic
// kernel above walks it three times through scalar loads), and leaves as 8-byte stores. NQ = float4 per lane.
template <int NQ>
__global__ void __launch_bounds__(256) build_z_large_fast_kernel(const float* __restrict__ raw, __half* __restrict__ z,
                                                                 int ldz, int seg, int lo_seg, long tokens_total, long N,
                                                                 int c_raw, AxisInfo ax, int F,
                                                                 const float* __restrict__ tab, long tok0) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long t = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tokens_total) return;
  const int n_feat = F * ax.n_axes;
  const int C = c_raw + n_feat;
  const float4* r4 = reinterpret_cast<const float4*>(raw + t * c_raw);
  float4 v[NQ];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = (q * 32 + lane) * 4;
    v[q] = c < c_raw ? __ldg(r4 + q * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
    s += v[q].x + v[q].y + v[q].z + v[q].w;
  }
  // positional features: lane k < n_feat owns feature k (axis k / F, entry k % F)
  float pf = 0.f;
  if (lane < n_feat) {
    unsigned rem = static_cast<unsigned>(t % N + tok0);
    const int a_own = lane / F;
    int row = 0;
    for (int a = ax.n_axes - 1; a >= 0; --a) {
      const unsigned sz = static_cast<unsigned>(ax.size[a]);
      const unsigned qd = rem / sz;
      if (a == a_own) row = ax.off[a] + static_cast<int>(rem - qd * sz);
      rem = qd;
    }
    pf = __ldg(tab + static_cast<long>(row) * F + (lane - a_own * F));
    s += pf;
  }
  const float mean = warp_sum(s) / C;
  float qq = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    if ((q * 32 + lane) * 4 < c_raw) {
      const float d0 = v[q].x - mean, d1 = v[q].y - mean, d2 = v[q].z - mean, d3 = v[q].w - mean;
      qq += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
  }
  if (lane < n_feat) qq += (pf - mean) * (pf - mean);
  const float rstd = rsqrtf(warp_sum(qq) / C + LN_EPS);
  __half* zr = z + t * ldz;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = (q * 32 + lane) * 4;
    if (c < c_raw) {
      uint2 hi, lo;
      const float o0 = (v[q].x - mean) * rstd, o1 = (v[q].y - mean) * rstd, o2 = (v[q].z - mean) * rstd,
                  o3 = (v[q].w - mean) * rstd;
      const __half2 h01 = __floats2half2_rn(o0, o1), h23 = __floats2half2_rn(o2, o3);
      hi.x = *reinterpret_cast<const uint32_t*>(&h01);
      hi.y = *reinterpret_cast<const uint32_t*>(&h23);
      *reinterpret_cast<uint2*>(zr + c) = hi;
      if (lo_seg > 0) {
        const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
        const __half2 l01 = __floats2half2_rn(o0 - f01.x, o1 - f01.y), l23 = __floats2half2_rn(o2 - f23.x, o3 - f23.y);
        lo.x = *reinterpret_cast<const uint32_t*>(&l01);
        lo.y = *reinterpret_cast<const uint32_t*>(&l23);
        *reinterpret_cast<uint2*>(zr + lo_seg + c) = lo;
      }
    }
  }
  // positional features and the zero pad [C, seg)
  for (int c = c_raw + lane; c < seg; c += 32)  // (feature k sits in lane k: only the first trip holds features)
    store_split(zr, c, seg, lo_seg, (c < C) ? (pf - mean) * rstd : 0.f);
}

// ------------------------------------------------------------------ head: mean_L -> LN -> Linear
// stage 1: pooled[b][d] = mean_l x[b][l][d]; block = (sample, 32 columns), 8 warps stride the rows
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ x, int L, int D,
                                                   float* __restrict__ pooled) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  __shared__ float part[8][33];
  const int b = blockIdx.y, d = blockIdx.x * 32 + (threadIdx.x & 31), w = threadIdx.x >> 5;
  const float* xb = x + static_cast<size_t>(b) * L * D;
  float s = 0.f;
  if (d < D)
    for (int l = w; l < L; l += 8) s += xb[static_cast<size_t>(l) * D + d];
  part[w][threadIdx.x & 31] = s;
  __syncthreads();
  if (w == 0 && d < D) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
    pooled[static_cast<size_t>(b) * D + d] = t / L;
  }
}
// stage 2: logits[b][o] = LN(pooled[b]) . W[o] + bias[o]; one block per sample
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ pooled_g, int D,
                                                   const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                   const float* __restrict__ W, const float* __restrict__ bias,
                                                   int out_dims, float* __restrict__ logits) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  extern __shared__ float pooled[];  // D floats
  __shared__ float red[32];
  const int b = blockIdx.x;
  for (int d = threadIdx.x; d < D; d += blockDim.x) pooled[d] = pooled_g[static_cast<size_t>(b) * D + d];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float s = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) s += pooled[d];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < nw; ++w) tot += red[w];
  const float mean = tot / D;
  __syncthreads();
  float q = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float dd = pooled[d] - mean;
    q += dd * dd;
  }
  q = warp_sum(q);
  if (lane == 0) red[warp] = q;
  __syncthreads();
  tot = 0.f;
  for (int w = 0; w < nw; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / D + LN_EPS);
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) pooled[d] = (pooled[d] - mean) * rstd * ln_w[d] + ln_b[d];
  __syncthreads();
  for (int o = warp; o < out_dims; o += nw) {
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) acc += pooled[d] * W[static_cast<size_t>(o) * D + d];
    acc = warp_sum(acc);
    if (lane == 0) logits[static_cast<size_t>(b) * out_dims + o] = acc + bias[o];
  }
}

// ------------------------------------------------------------------ mask bytes -> tile bit words
__global__ void pack_mask_kernel(const uint8_t* __restrict__ mask, uint64_t* __restrict__ bits, long N,
                                 long tiles_per_sample, long total_tiles) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const long w = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (w >= total_tiles) return;
  const long b = w / tiles_per_sample, tile = w % tiles_per_sample;
  uint64_t v = 0;
  for (int j = 0; j < 64; ++j) {
    const long n = tile * 64 + j;
    if (n < N && mask[b * N + n]) v |= (1ull << j);
  }
  bits[w] = v;
}

// ------------------------------------------------------------------ token-axis sharding across GPUs (peer memory)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Block-wide wait until every rank has published exchange `seq` (its flag in OUR header, written by the peer over
// NVLink). Bounded: after pp.timeout_clk clocks (default ~30 s, hn_set_exchange_timeout) the block gives up and records
// the error, so a dead peer can never hang the GPU; the forward then poisons its outputs with NaN
// (poison_on_error_kernel) and the host raises at its next check (hn_exchange_error*). Call with all threads of the block.
__device__ __forceinline__ void peers_wait(const PeerParts& pp) {
  if (pp.world == 0) return;
  if (threadIdx.x < pp.world) {
    const unsigned long long* f = &pp.hdr[pp.rank]->flags[threadIdx.x];
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < pp.seq) {
      if (clock64() - t0 > pp.timeout_clk) {
        pp.hdr[pp.rank]->error = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
}
// split s of row (b, h, l): local layout [b][nsplit][H][L], peer layout = one partial per rank
__device__ __forceinline__ long part_row(const PeerParts& pp, int b, int s, int nsplit, int H, int h, int L, int l) {
  return pp.world ? ((static_cast<long>(b) * H + h) * L + l) : ((((static_cast<long>(b) * nsplit + s) * H + h) * L) + l);
}

// one warp per (b, h, l): merges the local splits into this rank's slot (un-normalised accumulator at the merged
// max, merged max, merged row sum); the last block to finish publishes the exchange to every peer
__global__ void __launch_bounds__(256) merge_signal_kernel(const float* __restrict__ part_acc,
                                                           const float* __restrict__ part_ml, int batch, int nsplit,
                                                           int H, int L, int w, float* __restrict__ slot_acc,
                                                           float* __restrict__ slot_ml, PeerParts pp) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  const int lane = threadIdx.x & 31;
  const long wid = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long total = static_cast<long>(batch) * H * L;
  if (wid < total) {
    const int l = static_cast<int>(wid % L);
    const int h = static_cast<int>((wid / L) % H);
    const int b = static_cast<int>(wid / (static_cast<long>(L) * H));
    float M = -INFINITY;
    for (int s = lane; s < nsplit; s += 32)
      M = fmaxf(M, part_ml[((((static_cast<long>(b) * nsplit + s) * H + h) * L) + l) * 2]);
    M = warp_max(M);
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    float den = 0.f;
    const int nv = w / 32;
    for (int s = 0; s < nsplit; ++s) {
      const long base = (((static_cast<long>(b) * nsplit + s) * H + h) * L) + l;
      const float m = part_ml[base * 2], ls = part_ml[base * 2 + 1];
      const float wgt = (m == -INFINITY) ? 0.f : exp2f(m - M);
#pragma unroll
      for (int v = 0; v < 4; ++v)
        if (v < nv) a[v] += wgt * part_acc[base * w + v * 32 + lane];
      den += wgt * ls;
    }
    const long orow = (static_cast<long>(b) * H + h) * L + l;
#pragma unroll
    for (int v = 0; v < 4; ++v)
      if (v < nv) slot_acc[orow * w + v * 32 + lane] = a[v];
    if (lane == 0) {
      slot_ml[orow * 2] = M;
      slot_ml[orow * 2 + 1] = den;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    XchgHeader* mine = pp.hdr[pp.rank];
    __threadfence();
    const unsigned int prev = atomicAdd(&mine->blocks_done, 1u);
    if (prev == gridDim.x - 1) {
      mine->blocks_done = 0;
      __threadfence_system();
      for (int r = 0; r < pp.world; ++r) st_release_sys(&pp.hdr[r]->flags[pp.rank], pp.seq);
    }
  }
}

// a token-sharded forward whose peer wait timed out must not hand out plausible-looking numbers: NaN them
__global__ void poison_on_error_kernel(float* __restrict__ out, long n, const XchgHeader* __restrict__ hdr) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  if (hdr->error == 0) return;
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------ split combine
// part_acc [b][s][h][L][hp], part_ml [b][s][h][L][2]; one warp per (b, l, h); hp = 64 | 128 accumulator columns
__global__ void __launch_bounds__(256) combine_generic_kernel(const float* __restrict__ part_acc,
                                                              const float* __restrict__ part_ml, int batch,
                                                              int nsplit, int H, int L, __half* __restrict__ O,
                                                              int o_ld, int lo_seg, int hp, int den_col,
                                                              PeerParts pp) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  peers_wait(pp);
  const int lane = threadIdx.x & 31;
  const long wid = static_cast<long>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long total = static_cast<long>(batch) * L * H;
  if (wid >= total) return;
  const int h = static_cast<int>(wid % H);
  const int l = static_cast<int>((wid / H) % L);
  const int b = static_cast<int>(wid / (static_cast<long>(H) * L));
  float M = -INFINITY;
  for (int s = lane; s < nsplit; s += 32)
    M = fmaxf(M, (pp.world ? pp.ml[s] : part_ml)[part_row(pp, b, s, nsplit, H, h, L, l) * 2]);
  M = warp_max(M);
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  float den = 0.f;
  const int nv = hp / 32;
  // four splits per step with all their loads issued before the first use: the merge is a chain of L2 round trips
  // otherwise (one warp per row, nothing else to hide them)
  for (int s0 = 0; s0 < nsplit; s0 += 4) {
    float m4[4], l4[4], v4[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = s0 + u < nsplit ? s0 + u : nsplit - 1;
      const long base = part_row(pp, b, s, nsplit, H, h, L, l);
      const float* pml = pp.world ? pp.ml[s] : part_ml;
      const float* pac = pp.world ? pp.acc[s] : part_acc;
      m4[u] = pml[base * 2];
      l4[u] = pml[base * 2 + 1];
#pragma unroll
      for (int v = 0; v < 4; ++v) v4[u][v] = v < nv ? pac[base * hp + v * 32 + lane] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // fixed split order: the sum is the same whatever the unrolling
      if (s0 + u < nsplit) {
        const float w = (m4[u] == -INFINITY) ? 0.f : exp2f(m4[u] - M);
#pragma unroll
        for (int v = 0; v < 4; ++v) a[v] += w * v4[u][v];
        den += w * l4[u];
      }
    }
  }
  // small-context partials keep their denominator in accumulator column den_col (the column that met z's 1.0);
  // columns from there on are padding and leave as zeros
  if (den_col >= 0) den = __shfl_sync(0xffffffffu, den_col < 32 ? a[0] : a[1], den_col & 31);
  const float inv = 1.f / den;
  __half* o = O + (static_cast<long>(b) * L + l) * o_ld + h * hp;
#pragma unroll
  for (int v = 0; v < 4; ++v)
    if (v < nv) store_split(o, v * 32 + lane, 0, lo_seg, (den_col >= 0 && v * 32 + lane >= den_col) ? 0.f : a[v] * inv);
}

// small-C: acc rows are zw wide: [sum_t p z_c (c < C), sum_t p (col C), 0...]; then the V projection
// O[b*L + l][h*hp + d] = (u / den) . Wv'[h*dh + d][:] + bv[h*dh + d], hp = 64 | 128 output columns per head.
// Block = (32 latent rows, head, sample): warp w merges the splits of 4 rows (lane = column), the head's Wv' sits
// transposed in shared memory so the projection reads are conflict-free broadcasts.
__global__ void __launch_bounds__(256) combine_vproj_kernel(const float* __restrict__ part_acc,
                                                            const float* __restrict__ part_ml, int batch,
                                                            int nsplit, int H, int L, int C, int zw, int dh,
                                                            const float* __restrict__ Wv,
                                                            const float* __restrict__ bv, __half* __restrict__ O,
                                                            int o_ld, int lo_seg, int hp, PeerParts pp) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  __shared__ float wT[64][129];  // wT[c][d] = Wv'[h*dh + d][c]
  __shared__ float u_s[32][65];  // merged, normalised rows
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int l0 = blockIdx.x * 32, h = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < hp * 64; i += 256) {
    const int d = i / 64, c = i % 64;
    wT[c][d] = (d < dh && c < C) ? Wv[static_cast<long>(h * dh + d) * zw + c] : 0.f;
  }
  peers_wait(pp);  // (after the weight tile is on its way: the wait overlaps those loads)
  for (int r = w; r < 32; r += 8) {
    const int l = l0 + r;
    float acc0 = 0.f, acc1 = 0.f;
    if (l < L) {
      float M = -INFINITY;
      for (int s = lane; s < nsplit; s += 32)
        M = fmaxf(M, (pp.world ? pp.ml[s] : part_ml)[part_row(pp, b, s, nsplit, H, h, L, l) * 2]);
      M = warp_max(M);
      for (int s = 0; s < nsplit; ++s) {
        const long base = part_row(pp, b, s, nsplit, H, h, L, l);
        const float* pml = pp.world ? pp.ml[s] : part_ml;
        const float* pac = pp.world ? pp.acc[s] : part_acc;
        const float m = pml[base * 2];
        const float wgt = (m == -INFINITY) ? 0.f : exp2f(m - M);
        acc0 += wgt * pac[base * zw + lane];
        if (zw > 32) acc1 += wgt * pac[base * zw + 32 + lane];
      }
    }
    u_s[r][lane] = acc0;
    u_s[r][lane + 32] = acc1;
  }
  __syncthreads();
  // 32 rows x hp output columns over 256 threads
  for (int i = threadIdx.x; i < 32 * hp; i += 256) {
    const int r = i / hp, d = i % hp, l = l0 + r;
    if (l >= L) continue;
    float out = 0.f;
    if (d < dh) {
      for (int c = 0; c < C; ++c) out += u_s[r][c] * wT[c][d];
      out = out / u_s[r][C] + bv[h * dh + d];
    }
    __half* o = O + (static_cast<long>(b) * L + l) * o_ld + h * hp;
    store_split(o, d, 0, lo_seg, out);
  }
}

__global__ void broadcast_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, long n, int batch) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  for (long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const float v = src[i];
    for (int b = 0; b < batch; ++b) dst[static_cast<long>(b) * n + i] = v;
  }
}

// ------------------------------------------------------------------ attention-weight export (opt-in)
// Materialises attn = softmax(2 q k^T / sqrt(dh)) of one Attention call — what the reference keeps in
// `Attention.attn_weights` (healnet.py:420) and hands out through get_attention_weights() (:252-262) — from the
// same fp16 operands the streaming kernel used, normalised with the row statistics of the split partials, so
// the exported matrix is the one the fused forward actually applied. out[(b*H + h)][l][n], fp32.
// Block: 16 query rows x 256 tokens of one (sample, head); thread = token.
template <int KW>  // operand width per head in fp16 elements (32, 64 or 128)
__global__ void __launch_bounds__(256) attn_export_kernel(
    const __half* __restrict__ Q, int q_ld, int q_col0, int q_lo_off, const __half* __restrict__ K, long k_ld,
    int k_col0, int k_lo_off, int q_pitch, int k_pitch, int batch, int H, int L, long N, int nsplit,
    const float* __restrict__ part_acc, const float* __restrict__ part_ml, int acc_w, int den_col, float den_scale,
    const uint64_t* __restrict__ mask_bits, float* __restrict__ out) {
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();
  __shared__ float q_s[16][KW];
  __shared__ float m_s[16], inv_s[16];
  const int bh = blockIdx.z, b = bh / H, h = bh % H;
  const int l0 = blockIdx.y * 16;
  const long n = static_cast<long>(blockIdx.x) * 256 + threadIdx.x;
  for (int i = threadIdx.x; i < 16 * KW; i += 256) {
    const int r = i / KW, c = i % KW, l = l0 + r;
    float v = 0.f;
    if (l < L) {
      const __half* qr = Q + (static_cast<long>(b) * L + l) * q_ld + q_col0 + h * q_pitch + c;
      v = __half2float(qr[0]);
      if (q_lo_off > 0) v += __half2float(qr[q_lo_off]);
    }
    q_s[r][c] = v;
  }
  if (threadIdx.x < 16) {
    const int l = l0 + threadIdx.x;
    float M = -INFINITY, den = 0.f;
    if (l < L) {
      for (int s = 0; s < nsplit; ++s)
        M = fmaxf(M, part_ml[((((static_cast<long>(b) * nsplit + s) * H + h) * L) + l) * 2]);
      for (int s = 0; s < nsplit; ++s) {
        const long base = (((static_cast<long>(b) * nsplit + s) * H + h) * L) + l;
        const float m = part_ml[base * 2];
        const float w = (m == -INFINITY) ? 0.f : exp2f(m - M);
        den += w * (den_col >= 0 ? part_acc[base * acc_w + den_col] : part_ml[base * 2 + 1]);
      }
    }
    m_s[threadIdx.x] = M;
    inv_s[threadIdx.x] = 1.f / (den * den_scale);
  }
  __syncthreads();
  if (n >= N) return;
  float k[KW];
  const __half* kr = K + (static_cast<long>(b) * N + n) * k_ld + k_col0 + h * k_pitch;
#pragma unroll
  for (int c = 0; c < KW; c += 8) {
    const uint4 raw = *reinterpret_cast<const uint4*>(kr + c);
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      k[c + 2 * j] = f.x;
      k[c + 2 * j + 1] = f.y;
    }
    if (k_lo_off > 0) {
      const uint4 rawl = *reinterpret_cast<const uint4*>(kr + k_lo_off + c);
      const __half2* l2 = reinterpret_cast<const __half2*>(&rawl);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(l2[j]);
        k[c + 2 * j] += f.x;
        k[c + 2 * j + 1] += f.y;
      }
    }
  }
  bool keep = true;
  if (mask_bits != nullptr) keep = (mask_bits[static_cast<long>(b) * ((N + 63) / 64) + n / 64] >> (n % 64)) & 1ull;
  for (int r = 0; r < 16 && l0 + r < L; ++r) {
    float sdot = 0.f;
#pragma unroll
    for (int c = 0; c < KW; ++c) sdot += q_s[r][c] * k[c];
    const float p = keep ? exp2f(sdot - m_s[r]) * inv_s[r] : 0.f;
    out[(static_cast<long>(bh) * L + l0 + r) * N + n] = p;
  }
}

AxisInfo make_axis(const int* sizes, int n_axes) {
  AxisInfo ax{};
  ax.n_axes = n_axes;
  int off = 0;
  for (int a = 0; a < 4; ++a) {
    ax.size[a] = a < n_axes ? sizes[a] : 1;
    ax.off[a] = off;
    if (a < n_axes) off += sizes[a];
  }
  return ax;
}
}  // namespace

int launch_layernorm_f16(const float* x, int ldx, const float* gamma, const float* beta, __half* y, int ldy, int seg,
                         int lo_seg, long rows, int D, cudaStream_t stream) {
  if (rows <= 0) return 0;
  HN_REQUIRE(seg >= D && (lo_seg == 0 || lo_seg >= seg) && ldy >= lo_seg + seg, "layernorm: bad output layout");
  const int wpb = 8;
  const unsigned grid = static_cast<unsigned>((rows + wpb - 1) / wpb);
  if (seg <= 128)
    HN_CHECK_CUDA(launch_k(layernorm_f16_kernel<4>, dim3(grid), dim3(wpb * 32), 0, stream, x, ldx, gamma, beta, y, ldy,
                           seg, lo_seg, rows, D));
  else if (seg <= 512)
    HN_CHECK_CUDA(launch_k(layernorm_f16_kernel<16>, dim3(grid), dim3(wpb * 32), 0, stream, x, ldx, gamma, beta, y, ldy,
                           seg, lo_seg, rows, D));
  else if (seg <= 1024)
    HN_CHECK_CUDA(launch_k(layernorm_f16_kernel<32>, dim3(grid), dim3(wpb * 32), 0, stream, x, ldx, gamma, beta, y, ldy,
                           seg, lo_seg, rows, D));
  else
    HN_CHECK_CUDA(launch_k(layernorm_f16_big_kernel, dim3(grid), dim3(wpb * 32), 0, stream, x, ldx, gamma, beta, y, ldy,
                           seg, lo_seg, rows, D));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_axis_tables(float* tab, const int* axis_sizes, int n_axes, int n_bands, float max_freq,
                       cudaStream_t stream) {
  HN_REQUIRE(n_axes >= 1 && n_axes <= 4, "at most 4 spatial axes are supported");
  AxisInfo ax = make_axis(axis_sizes, n_axes);
  HN_CHECK_CUDA(launch_k(axis_tables_kernel, dim3(8), dim3(256), 0, stream, tab, ax, n_bands, max_freq));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace {
template <typename T>
int build_z_small_t(const T* raw, __half* z, int zw, int batch, long N, int c_raw, int n_axes, const int* axis_sizes,
                    int n_bands, const float* tab, int fourier, cudaStream_t stream, long tok0, int split) {
  const int F = fourier ? 2 * n_bands + 1 : 0;
  const int C = c_raw + F * n_axes;
  HN_REQUIRE(zw == 32 || zw == 64, "small-C context rows are 32 or 64 wide");
  HN_REQUIRE(C <= zw - 1, "small-C context path needs C <= row width - 1");
  AxisInfo ax = make_axis(axis_sizes, n_axes);
  const long total = static_cast<long>(batch) * N;
  const unsigned grid = static_cast<unsigned>((total + 255) / 256);
  if (zw == 32 && F == 5 && c_raw >= 1 && c_raw <= 4 && n_axes >= 1 && n_axes <= 3) {
#define HN_FAST(CR, NA)                                                                                         \
  if (c_raw == CR && n_axes == NA) {                                                                            \
    HN_CHECK_CUDA(launch_k(build_z_small32_fast_kernel<CR, NA, T>, dim3(grid), dim3(256), 0, stream, raw, z, total, N, \
                           ax, tab, tok0, split));                                                              \
    return 0;                                                                                                   \
  }
    HN_FAST(1, 1) HN_FAST(1, 2) HN_FAST(1, 3) HN_FAST(2, 1) HN_FAST(2, 2) HN_FAST(2, 3)
    HN_FAST(3, 1) HN_FAST(3, 2) HN_FAST(3, 3) HN_FAST(4, 1) HN_FAST(4, 2) HN_FAST(4, 3)
#undef HN_FAST
  }
  if (zw == 32)
    HN_CHECK_CUDA(launch_k(build_z_small_kernel<32, T>, dim3(grid), dim3(256), 0, stream, raw, z, total, N, c_raw, ax, F,
                           tab, tok0, split));
  else
    HN_CHECK_CUDA(launch_k(build_z_small_kernel<64, T>, dim3(grid), dim3(256), 0, stream, raw, z, total, N, c_raw, ax, F,
                           tab, tok0, split));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

// in_dtype: element type of `raw` — 0 fp32, 1 bf16, 2 fp16 (hn_set_io_dtype)
int launch_build_z_small(const void* raw, __half* z, int zw, int batch, long N, int c_raw, int n_axes,
                         const int* axis_sizes, int n_bands, const float* tab, int fourier, cudaStream_t stream,
                         long tok0, int split, int in_dtype) {
  if (in_dtype == 1)
    return build_z_small_t(static_cast<const __nv_bfloat16*>(raw), z, zw, batch, N, c_raw, n_axes, axis_sizes, n_bands, tab,
                           fourier, stream, tok0, split);
  if (in_dtype == 2)
    return build_z_small_t(static_cast<const __half*>(raw), z, zw, batch, N, c_raw, n_axes, axis_sizes, n_bands, tab,
                           fourier, stream, tok0, split);
  return build_z_small_t(static_cast<const float*>(raw), z, zw, batch, N, c_raw, n_axes, axis_sizes, n_bands, tab, fourier,
                         stream, tok0, split);
}

int launch_build_z_large(const void* raw_v, __half* z, int ldz, int lo_seg, int batch, long N, int c_raw, int n_axes,
                         const int* axis_sizes, int n_bands, const float* tab, int fourier, cudaStream_t stream,
                         long tok0, int in_dtype) {
  const int F = fourier ? 2 * n_bands + 1 : 0;
  const int seg = lo_seg > 0 ? lo_seg : ldz;
  HN_REQUIRE(seg >= c_raw + F * n_axes && ldz >= lo_seg + seg, "build_z_large: bad output layout");
  AxisInfo ax = make_axis(axis_sizes, n_axes);
  const long total = static_cast<long>(batch) * N;
  const unsigned grid = static_cast<unsigned>((total + 7) / 8);
  const int n_feat = F * n_axes;
  if (in_dtype == 1) {
    HN_CHECK_CUDA(launch_k(build_z_large_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, stream,
                           static_cast<const __nv_bfloat16*>(raw_v), z, ldz, seg, lo_seg, total, N, c_raw, ax, F, tab, tok0));
    HN_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  if (in_dtype == 2) {
    HN_CHECK_CUDA(launch_k(build_z_large_kernel<__half>, dim3(grid), dim3(256), 0, stream,
                           static_cast<const __half*>(raw_v), z, ldz, seg, lo_seg, total, N, c_raw, ax, F, tab, tok0));
    HN_CHECK_CUDA(cudaGetLastError());
    return 0;
  }
  const float* raw = static_cast<const float*>(raw_v);
  const bool fast = c_raw % 4 == 0 && c_raw >= 4 && c_raw <= 1024 && n_feat <= 32 && N < (1L << 31) &&
                    (reinterpret_cast<uintptr_t>(raw) & 15) == 0 && ldz % 4 == 0 && lo_seg % 4 == 0 &&
                    (reinterpret_cast<uintptr_t>(z) & 7) == 0;
  if (fast && c_raw <= 256)
    HN_CHECK_CUDA(launch_k(build_z_large_fast_kernel<2>, dim3(grid), dim3(256), 0, stream, raw, z, ldz, seg, lo_seg, total,
                           N, c_raw, ax, F, tab, tok0));
  else if (fast && c_raw <= 512)
    HN_CHECK_CUDA(launch_k(build_z_large_fast_kernel<4>, dim3(grid), dim3(256), 0, stream, raw, z, ldz, seg, lo_seg, total,
                           N, c_raw, ax, F, tab, tok0));
  else if (fast)
    HN_CHECK_CUDA(launch_k(build_z_large_fast_kernel<8>, dim3(grid), dim3(256), 0, stream, raw, z, ldz, seg, lo_seg, total,
                           N, c_raw, ax, F, tab, tok0));
  else
    HN_CHECK_CUDA(launch_k(build_z_large_kernel<float>, dim3(grid), dim3(256), 0, stream, raw, z, ldz, seg, lo_seg, total, N,
                           c_raw, ax, F, tab, tok0));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_head(const float* x, int batch, int L, int D, const float* ln_w, const float* ln_b, const float* W,
                const float* bias, int out_dims, float* pooled, float* logits, cudaStream_t stream) {
  HN_CHECK_CUDA(launch_k(pool_kernel, dim3((D + 31) / 32, batch), dim3(256), 0, stream, x, L, D, pooled));
  HN_CHECK_CUDA(cudaGetLastError());
  HN_CHECK_CUDA(launch_k(head_kernel, dim3(batch), dim3(256), (D + 2) * sizeof(float), stream, pooled, D, ln_w, ln_b, W,
                         bias, out_dims, logits));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_pack_mask(const uint8_t* mask, uint64_t* bits, int batch, long N, cudaStream_t stream) {
  const long tiles = (N + 63) / 64;
  const long total = tiles * batch;
  HN_CHECK_CUDA(launch_k(pack_mask_kernel, dim3(static_cast<unsigned>((total + 127) / 128)), dim3(128), 0, stream, mask,
                         bits, N, tiles, total));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_combine_generic(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L,
                           __half* O, int o_ld, int lo_seg, int hp, cudaStream_t stream, const PeerParts* peers,
                           int den_col) {
  HN_REQUIRE(hp == 32 || hp == 64 || hp == 128, "combine: accumulator rows are 32, 64 or 128 wide");
  HN_REQUIRE(den_col < hp && den_col < 64, "combine: denominator column outside the accumulator row");
  const long total = static_cast<long>(batch) * L * H;
  PeerParts pp;
  if (peers != nullptr) pp = *peers;
  HN_CHECK_CUDA(launch_k(combine_generic_kernel, dim3(static_cast<unsigned>((total + 7) / 8)), dim3(256), 0, stream,
                         part_acc, part_ml, batch, pp.world ? pp.world : nsplit, H, L, O, o_ld, lo_seg, hp, den_col, pp));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_combine_vproj(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L, int C,
                         int zw, int dh, const float* Wv, const float* bv, __half* O, int o_ld, int lo_seg, int hp,
                         cudaStream_t stream, const PeerParts* peers) {
  HN_REQUIRE((zw == 32 || zw == 64) && C <= zw - 1 && (hp == 64 || hp == 128) && dh <= hp,
             "combine_vproj: C < zw and dim_head <= head pitch (64 | 128) required");
  PeerParts pp;
  if (peers != nullptr) pp = *peers;
  HN_CHECK_CUDA(launch_k(combine_vproj_kernel, dim3((L + 31) / 32, H, batch), dim3(256), 0, stream, part_acc, part_ml,
                         batch, pp.world ? pp.world : nsplit, H, L, C, zw, dh, Wv, bv, O, o_ld, lo_seg, hp, pp));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_merge_signal(const float* part_acc, const float* part_ml, int batch, int nsplit, int H, int L, int w,
                        float* slot_acc, float* slot_ml, const PeerParts& peers, cudaStream_t stream) {
  HN_REQUIRE(peers.world >= 1 && peers.world <= HN_MAX_PEERS, "merge: bad peer table");
  HN_REQUIRE(w == 32 || w == 64 || w == 128, "merge: accumulator rows are 32, 64 or 128 wide");
  const long total = static_cast<long>(batch) * H * L;
  HN_CHECK_CUDA(launch_k(merge_signal_kernel, dim3(static_cast<unsigned>((total + 7) / 8)), dim3(256), 0, stream, part_acc,
                         part_ml, batch, nsplit, H, L, w, slot_acc, slot_ml, peers));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_poison_on_error(float* out, long n, const XchgHeader* hdr, cudaStream_t stream) {
  const unsigned grid = static_cast<unsigned>(n / 256 + 1 < 592 ? n / 256 + 1 : 592);
  HN_CHECK_CUDA(launch_k(poison_on_error_kernel, dim3(grid), dim3(256), 0, stream, out, n, hdr));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_attn_export(const AttnArgs& a, float* out, cudaStream_t stream) {
  HN_REQUIRE(out != nullptr, "attention export: null output");
  const dim3 grid(static_cast<unsigned>((a.N + 255) / 256), (a.L + 15) / 16, a.batch * a.H);
  // (split small-context rows are [z_hi (kd) | z_lo (kd)])
  const int q_lo = a.precise ? a.q_lo_off : 0, k_lo = a.precise ? (a.shared_kv ? a.kd : a.kv_lo_off) : 0;
  // small-context path: Q' heads are kd columns apart, every head reads the same z row, the denominator is the
  // accumulator column that met z's 1.0; generic path: heads hp apart in Q and K, denominator = tracked row sum
  const int kw = a.shared_kv ? a.kd : a.hp;
  const int q_pitch = kw, k_pitch = a.shared_kv ? 0 : a.hp;
  const int k_col0 = a.shared_kv ? 0 : a.k_col0;
  const int den_col = a.shared_kv ? a.c_ones : -1;
  // xattn_small.cu accumulates P' = 2^P_SHIFT * P (P_SHIFT = 10): its denominator column carries that factor
  const float den_scale = a.shared_kv ? 0.0009765625f : 1.f;
#define HN_EXPORT(KW)                                                                                              \
  attn_export_kernel<KW><<<grid, 256, 0, stream>>>(a.Q, a.q_ld, 0, q_lo, a.KV, a.kv_ld, k_col0, k_lo, q_pitch,      \
                                                   k_pitch, a.batch, a.H, a.L, a.N, a.nsplit, a.part_acc,          \
                                                   a.part_ml, kw, den_col, den_scale, a.mask_bits, out)
  if (kw == 32)
    HN_EXPORT(32);
  else if (kw == 64)
    HN_EXPORT(64);
  else
    HN_EXPORT(128);
#undef HN_EXPORT
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_broadcast_rows(const float* src, float* dst, long n, int batch, cudaStream_t stream) {
  const unsigned grid = static_cast<unsigned>(n / 256 + 1 < 1184 ? n / 256 + 1 : 1184);
  HN_CHECK_CUDA(launch_k(broadcast_rows_kernel, dim3(grid), dim3(256), 0, stream, src, dst, n, batch));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace hn
