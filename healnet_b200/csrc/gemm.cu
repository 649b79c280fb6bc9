// gemm.cu — latent-side dense contractions of the HEALNet fusion stack on tcgen05 tensor cores.
//
//   C[M,N] = A[M,K] · B[N,K]ᵀ   fp16 operands, fp32 accumulate in TMEM, fused epilogues:
//     to_q / to_kv / fused QKV projections        -> EPI_F16 (+bias)          (reference healnet.py:403-405)
//     to_out + bias + LeakyReLU(0.01) + residual   -> EPI_RES_LEAKY            (healnet.py:383-386,426,236)
//     FeedForward Linear(D,8D) + a*selu(g) gate    -> EPI_GATE_F16             (healnet.py:328-331,344-346)
//     FeedForward Linear(4D,D) + bias + residual   -> EPI_RES                  (healnet.py:347,237)
//
// Precision: fp16 operands carry 11 significant bits, which is not enough for the latent residual stream to
// track the fp32 reference to rtol 1e-3 / atol 1e-4 (measured: logits off by up to 4e-4). The latent-side
// GEMMs therefore run "split": every operand x is held as hi = fp16(x), lo = fp16(x - hi) in adjacent column
// segments, and the product is accumulated as A_hi.B_hi + A_lo.B_hi + A_hi.B_lo (terms = 3; the dropped
// lo.lo term is ~2^-22) — three passes over the K tiles into the same TMEM accumulator. terms = 2 keeps only
// A_hi.(B_hi + B_lo) (exact weights, fp16 activations) for the wide-context K/V projection, whose per-token
// activation rounding averages out under the softmax while weight rounding would not.
//
// Persistent kernel, one CTA per SM, static round-robin over the 128 x BN output tiles (m fastest, so CTAs that run
// side by side read the same weight tile): warp 0 = TMA producer, warp 1 = single-thread UMMA issuer (+ TMEM
// allocator), warps 2..9 = epilogue (each thread owns one accumulator row = one TMEM lane, two warps per quadrant). These GEMMs have few rows
// (M = batch * L = 2048 at cfg 1), so they are bound by the bytes an SM has to pull in from L2 per k-tile, not by the
// tensor pipe. Hence:
//   * every distinct operand tile is loaded ONCE per k-tile — a stage holds [A_hi | A_lo | B_hi | B_lo] and feeds the
//     three UMMA terms (the first version re-read A_hi and B_hi in separate passes: 6 tile loads instead of 4);
//   * the ring uses the whole shared memory of the SM (up to 8 stages) to cover L2 latency;
//   * two accumulators in TMEM (2 x BN columns): the epilogue of tile j overlaps the main loop of tile j+1, and the
//     per-CTA set-up (barriers, TMEM allocation, descriptor fetch) is paid once per SM instead of once per tile;
//   * BN = 256 whenever that still gives every SM a tile (halves the A bytes per FLOP).
//   * thread-block clusters with TMA multicast (CN = 2 | 4 CTAs): the CTAs of a cluster work on output tiles that share
//     one operand tile (the activations when the cluster runs along N, the weights when it runs along M); each CTA
//     fetches 1/CN of that tile and multicasts it to all of them, so the dominant L2 -> SM traffic drops by CN. The
//     stage's empty barrier then collects one tcgen05.commit from every CTA of the cluster (multicast arrive).
// M/N/K tails are handled by TMA out-of-bounds zero fill plus guards in the epilogue, so any shape whose row pitches
// are multiples of 8 elements is legal.
#include <cstdlib>

#include "common.cuh"
#include "tc05.cuh"

namespace hn {
namespace {
using namespace tc05;

constexpr int BM = 128;

struct GemmDev {
  int M, N, K;
  int epi, act;
  const float* bias;
  void* out;
  int ldo;
  int vec_ok;  // output rows are 16-byte aligned -> vector stores allowed
  int terms;   // 1, 2 or 3 UMMA terms per k-tile (see header comment)
  int a_seg, b_seg;  // column offset of the lo segment of A / B (elements)
  int out_seg;       // fp16 outputs: > 0 -> also store lo = fp16(v - hi) at column + out_seg
  int stages;        // depth of the operand ring (host: as many as fit the SM's shared memory, <= 8)
  int bias_vec;      // bias is 16-byte aligned -> float4 loads
  const float* ln_gamma;  // fused LayerNorm of completed row blocks (see GemmArgs), ln_out == null: off
  const float* ln_beta;
  __half* ln_out;
  int ln_ld, ln_seg;
  unsigned* ln_counters;
  int ln_epoch;  // fused launches so far in this forward, this one included
  int tiles_m, tiles_n;  // output tiles; with clusters: super-tiles of CN tiles along N (SHARE_A) or M (!SHARE_A)
  int bf16;              // operands are bf16 (kind::f16 with a_format = b_format = bf16)
  int nbatch, nb2;       // batched form: see GemmArgs
  int a_brows, b_brows;
  long out_b1, out_b2;
  float alpha;
};

__device__ __forceinline__ void split_half(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack_hi2(float a, float b, uint32_t& lo_bits) {
  __half ha, la, hb, lb;
  split_half(a, ha, la);
  split_half(b, hb, lb);
  __half2 h = __halves2half2(ha, hb), l = __halves2half2(la, lb);
  lo_bits = *reinterpret_cast<uint32_t*>(&l);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float selu_f(float x) {
  const float alpha = 1.6732632423543772848170429916717f;
  const float scale = 1.0507009873554804934193349852946f;
  return scale * (x > 0.f ? x : alpha * (__expf(x) - 1.f));
}
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float leaky_f(float x) { return x > 0.f ? x : 0.01f * x; }

constexpr int MAX_STAGES = 8;
constexpr int SMEM_BUDGET = 220 * 1024;

constexpr int EPI_WARPS = 8;  // two per TMEM lane quadrant: each takes every other 32-column chunk of the tile

// CN: CTAs per cluster (1 = no cluster). SHARE_A: the cluster's CTAs take CN consecutive n-tiles of the same m-tile and
// share (multicast) the A tile; otherwise CN consecutive m-tiles of the same n-tile sharing the B tile. The tensor
// map of the shared operand has a box of 1/CN of the tile rows.
template <int BN, int BK, int CN, bool SHARE_A>
__global__ void __launch_bounds__((2 + EPI_WARPS) * 32, 1) gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                      const __grid_constant__ CUtensorMap tmB, GemmDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr uint32_t LAYOUT = (BK == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * BK * 2;
  __shared__ uint64_t full_bar[MAX_STAGES], empty_bar[MAX_STAGES], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_a = p.terms == 3 ? 2 : 1, n_b = p.terms >= 2 ? 2 : 1;  // distinct A / B tiles per k-tile
  const int stage_bytes = n_a * A_BYTES + n_b * B_BYTES;
  const int stages = p.stages;
  // persistent loop over super-tiles (= tiles when CN == 1); every CTA of a cluster walks the same sequence
  const int crank = CN > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int sup_m = SHARE_A ? p.tiles_m : p.tiles_m / CN, sup_n = SHARE_A ? p.tiles_n / CN : p.tiles_n;
  const int tiles_per = sup_m * sup_n;          // per batch
  const int n_tiles = tiles_per * p.nbatch;
  const int first = static_cast<int>(blockIdx.x) / CN, stride = static_cast<int>(gridDim.x) / CN;
  const int k_tiles = (p.K + BK - 1) / BK;
  constexpr uint16_t CMASK = static_cast<uint16_t>((1u << CN) - 1u);
  // (row / column origin of a tile INSIDE its batch; the batch index is tile / tiles_per)
  auto tile_m0 = [&](int t) { t %= tiles_per; return ((t % sup_m) * (SHARE_A ? 1 : CN) + (SHARE_A ? 0 : crank)) * BM; };
  auto tile_n0 = [&](int t) { t %= tiles_per; return ((t / sup_m) * (SHARE_A ? CN : 1) + (SHARE_A ? crank : 0)) * BN; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], CN);  // one tcgen05.commit from every CTA of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&acc_full[a], 1);
      mbar_init(&acc_empty[a], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<2 * BN>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (CN > 1) cluster_sync_all();  // every CTA's barriers exist before any peer multicasts into it
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();  // everything above overlapped the previous kernel's tail; operands / x are valid from here on

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      int it = 0;
      for (int tile = first; tile < n_tiles; tile += stride) {
        const int z = tile / tiles_per;
        const int m0 = tile_m0(tile) + z * p.a_brows, n0 = tile_n0(tile) + z * p.b_brows;  // operand ROW coordinates
        for (int kt = 0; kt < k_tiles; ++kt, ++it) {
          const int s = it % stages;
          mbar_wait(&empty_bar[s], ((it / stages) & 1) ^ 1);  // (clusters: released by every CTA that reads it)
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          uint8_t* st = smem + s * stage_bytes;
          const int k0 = kt * BK;
          if (CN > 1 && SHARE_A) {  // my quarter (half) of the shared A tile, delivered to every CTA of the cluster
            constexpr int SL = A_BYTES / CN, ROWS = BM / CN;
            tma_load_2d_mc(st + crank * SL, &tmA, &full_bar[s], k0, m0 + crank * ROWS, CMASK);
            if (n_a == 2) tma_load_2d_mc(st + A_BYTES + crank * SL, &tmA, &full_bar[s], k0 + p.a_seg, m0 + crank * ROWS, CMASK);
          } else {
            tma_load_2d(st, &tmA, &full_bar[s], k0, m0);
            if (n_a == 2) tma_load_2d(st + A_BYTES, &tmA, &full_bar[s], k0 + p.a_seg, m0);
          }
          if (CN > 1 && !SHARE_A) {
            constexpr int SL = B_BYTES / CN, ROWS = BN / CN;
            tma_load_2d_mc(st + n_a * A_BYTES + crank * SL, &tmB, &full_bar[s], k0, n0 + crank * ROWS, CMASK);
            if (n_b == 2)
              tma_load_2d_mc(st + n_a * A_BYTES + B_BYTES + crank * SL, &tmB, &full_bar[s], k0 + p.b_seg, n0 + crank * ROWS, CMASK);
          } else {
            tma_load_2d(st + n_a * A_BYTES, &tmB, &full_bar[s], k0, n0);
            if (n_b == 2) tma_load_2d(st + n_a * A_BYTES + B_BYTES, &tmB, &full_bar[s], k0 + p.b_seg, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = idesc_f16(BM, BN, false, false) | (p.bf16 ? ((1u << 7) | (1u << 10)) : 0u);
      int it = 0, j = 0;
      for (int tile = first; tile < n_tiles; tile += stride, ++j) {
        const int acc = j & 1;
        mbar_wait(&acc_empty[acc], ((j >> 1) & 1) ^ 1);  // epilogue of tile j-2 has drained this accumulator
        fence_after_sync();
        const uint32_t tD = tmem + acc * BN;
        for (int kt = 0; kt < k_tiles; ++kt, ++it) {
          const int s = it % stages;
          mbar_wait(&full_bar[s], (it / stages) & 1);
          fence_after_sync();
          const uint32_t a_hi = smem_u32(smem + s * stage_bytes);
          const uint32_t a_lo = a_hi + A_BYTES;
          const uint32_t b_hi = a_hi + n_a * A_BYTES;
          const uint32_t b_lo = b_hi + B_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t dah = smem_desc(a_hi + k * 32, 16, SBO, LAYOUT), dbh = smem_desc(b_hi + k * 32, 16, SBO, LAYOUT);
            umma_ss(tD, dah, dbh, idesc, (kt | k) != 0);
            if (p.terms == 3) umma_ss(tD, smem_desc(a_lo + k * 32, 16, SBO, LAYOUT), dbh, idesc, true);
            if (p.terms >= 2) umma_ss(tD, dah, smem_desc(b_lo + k * 32, 16, SBO, LAYOUT), idesc, true);
          }
          if (CN > 1) umma_commit_mc(&empty_bar[s], CMASK); else umma_commit(&empty_bar[s]);
        }
        umma_commit(&acc_full[acc]);
      }
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32); the two warps of a quadrant interleave
    // the 32-column chunks (a single warp per scheduler advances at ~0.2 IPC on this dependent code: the FF1 gate
    // epilogue, not the tensor pipe, used to bound that GEMM)
    const uint32_t lane_base = (warp & 3) * 32;
    const int chunk0 = (warp - 2) >> 2;
    int j = 0;
    for (int tile = first; tile < n_tiles; tile += stride, ++j) {
    const int m0 = tile_m0(tile), n0 = tile_n0(tile);
    const int zb = tile / tiles_per;
    const size_t out_off = static_cast<size_t>(zb / p.nb2) * p.out_b1 + static_cast<size_t>(zb % p.nb2) * p.out_b2;
    const int acc = j & 1;
    const uint32_t tD = tmem + acc * BN;
    const int row = m0 + lane_base + lane;
    const bool row_ok = row < p.M;
    mbar_wait_sleepy(&acc_full[acc], (j >> 1) & 1, 2000);
    fence_after_sync();
#pragma unroll 1
    for (int c = chunk0; c < BN / 32; c += EPI_WARPS / 4) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tD, lane_base, c * 32), r);
      tmem_wait_ld();
      const int nb = n0 + c * 32;
      if (nb >= p.N) continue;
      const bool full = (nb + 32 <= p.N) && p.vec_ok;
      float v[32];
      if (p.bias_vec && nb + 32 <= p.N) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nb + j));
          v[j] = fmaf(__uint_as_float(r[j]), p.alpha, b4.x);
          v[j + 1] = fmaf(__uint_as_float(r[j + 1]), p.alpha, b4.y);
          v[j + 2] = fmaf(__uint_as_float(r[j + 2]), p.alpha, b4.z);
          v[j + 3] = fmaf(__uint_as_float(r[j + 3]), p.alpha, b4.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float b = 0.f;
          if (p.bias != nullptr && nb + j < p.N) b = __ldg(p.bias + nb + j);
          v[j] = fmaf(__uint_as_float(r[j]), p.alpha, b);
        }
      }
      if (row_ok) {
      if (p.epi == EPI_F16) {
        __half* o = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(row) * p.ldo + nb;
        const int out_seg = p.out_seg;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 q, ql;
            q.x = pack_hi2(v[j], v[j + 1], ql.x);
            q.y = pack_hi2(v[j + 2], v[j + 3], ql.y);
            q.z = pack_hi2(v[j + 4], v[j + 5], ql.z);
            q.w = pack_hi2(v[j + 6], v[j + 7], ql.w);
            *reinterpret_cast<uint4*>(o + j) = q;
            if (out_seg > 0) *reinterpret_cast<uint4*>(o + out_seg + j) = ql;
          }
        } else {
          for (int j = 0; j < 32 && nb + j < p.N; ++j) {
            __half hi, lo;
            split_half(v[j], hi, lo);
            o[j] = hi;
            if (out_seg > 0) o[out_seg + j] = lo;
          }
        }
      } else if (p.epi == EPI_GATE_F16) {
        __half* o = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(row) * p.ldo + nb / 2;
        float g[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float a = v[2 * j], t = v[2 * j + 1];
          g[j] = a * (p.act == ACT_SELU ? selu_f(t) : gelu_f(t));
        }
        if (full) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 q, ql;
            q.x = pack_hi2(g[j], g[j + 1], ql.x);
            q.y = pack_hi2(g[j + 2], g[j + 3], ql.y);
            q.z = pack_hi2(g[j + 4], g[j + 5], ql.z);
            q.w = pack_hi2(g[j + 6], g[j + 7], ql.w);
            *reinterpret_cast<uint4*>(o + j) = q;
            if (p.out_seg > 0) *reinterpret_cast<uint4*>(o + p.out_seg + j) = ql;
          }
        } else {
          for (int j = 0; j < 16 && nb + 2 * j < p.N; ++j) {
            __half hi, lo;
            split_half(g[j], hi, lo);
            o[j] = hi;
            if (p.out_seg > 0) o[p.out_seg + j] = lo;
          }
        }
      } else {
        float* o = reinterpret_cast<float*>(p.out) + out_off + static_cast<size_t>(row) * p.ldo + nb;
        const bool resid = (p.epi == EPI_RES || p.epi == EPI_RES_LEAKY);
        const bool leaky = (p.epi == EPI_RES_LEAKY || p.epi == EPI_LEAKY_F32);
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 x = resid ? *reinterpret_cast<const float4*>(o + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            x.x += leaky ? leaky_f(v[j]) : v[j];
            x.y += leaky ? leaky_f(v[j + 1]) : v[j + 1];
            x.z += leaky ? leaky_f(v[j + 2]) : v[j + 2];
            x.w += leaky ? leaky_f(v[j + 3]) : v[j + 3];
            *reinterpret_cast<float4*>(o + j) = x;
          }
        } else {
          for (int j = 0; j < 32 && nb + j < p.N; ++j) {
            const float t = leaky ? leaky_f(v[j]) : v[j];
            o[j] = resid ? o[j] + t : t;
          }
        }
      }
      }  // row_ok
      __syncwarp();
    }
    // accumulator drained (every tcgen05.ld above has completed): hand it back to the issuer
    fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&acc_empty[acc]);
    if (p.ln_out != nullptr) {
      // ---- fused PreNorm of the next module. Every CTA holds exactly one tile here (the host checks), so all tiles
      // of a 128-row block are in flight together: each CTA publishes its tile on the block's counter, waits until
      // the block is complete (counter == epoch * tiles_n; the counters only grow within a forward) and then
      // normalises ITS share of the block's rows — the LayerNorm runs on every SM, like the separate kernel did.
      __threadfence();
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      if (threadIdx.x == 64) {
        unsigned* ctr = p.ln_counters + m0 / BM;
        atomicAdd(ctr, 1u);
        const unsigned target = static_cast<unsigned>(p.ln_epoch) * static_cast<unsigned>(p.tiles_n);
        unsigned seen;
        int spins = 0;  // bounded (~1 s): a logic error must never hang the GPU
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
          if (seen < target) __nanosleep(40);
        } while (seen < target && ++spins < (1 << 24));
        if (seen < target) __trap();  // never normalise incomplete rows: fail the launch loudly instead
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      {
        const int ew = warp - 2;                                    // 0 .. EPI_WARPS-1
        const int share = (BM + p.tiles_n - 1) / p.tiles_n;         // rows of the block this CTA normalises
        const int rbeg = (n0 / BN) * share, rend = rbeg + share < BM ? rbeg + share : BM;
        constexpr int NQ = 4;  // float4 per lane: rows of up to 512 columns (the host checks)
        for (int r0 = rbeg + ew; r0 < rend; r0 += EPI_WARPS) {
          const int rr = m0 + r0;
          if (rr >= p.M) break;
          const float* xr = reinterpret_cast<const float*>(p.out) + static_cast<size_t>(rr) * p.ldo;
          float4 xv[NQ];
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int c = (q * 32 + lane) * 4;
            xv[q] = c < p.N ? __ldcg(reinterpret_cast<const float4*>(xr + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float sum = 0.f;
#pragma unroll
          for (int q = 0; q < NQ; ++q) sum += xv[q].x + xv[q].y + xv[q].z + xv[q].w;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float mean = sum / p.N;
          float qq = 0.f;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            if ((q * 32 + lane) * 4 < p.N) {
              const float d0 = xv[q].x - mean, d1 = xv[q].y - mean, d2 = xv[q].z - mean, d3 = xv[q].w - mean;
              qq += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
          const float rstd = rsqrtf(qq / p.N + 1e-5f);
          __half* y = p.ln_out + static_cast<size_t>(rr) * p.ln_ld;
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            const int c = (q * 32 + lane) * 4;
            if (c < p.N) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + c));
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + c));
              uint2 hi, lo;
              hi.x = pack_hi2((xv[q].x - mean) * rstd * g4.x + b4.x, (xv[q].y - mean) * rstd * g4.y + b4.y, lo.x);
              hi.y = pack_hi2((xv[q].z - mean) * rstd * g4.z + b4.z, (xv[q].w - mean) * rstd * g4.w + b4.w, lo.y);
              *reinterpret_cast<uint2*>(y + c) = hi;
              *reinterpret_cast<uint2*>(y + p.ln_seg + c) = lo;
            }
          }
        }
      }
    }
    }  // tile
  }
  fence_before_sync();
  __syncthreads();
  if (CN > 1) cluster_sync_all();  // nobody leaves while a peer may still multicast into, or arrive on, its memory
  if (warp == 1) tmem_dealloc<2 * BN>(tmem);
}

template <int BN, int BK, int CN, bool SHARE_A>
int launch_t(const GemmArgs& a, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  // single-segment operands rely on TMA zero fill beyond column K; split operands declare hi|lo and keep
  // explicit zeros in the pad columns [K, seg) of both segments
  const uint64_t a_cols = a.terms == 3 ? static_cast<uint64_t>(a.a_seg) + a.K : a.K;
  const uint64_t b_cols = a.terms >= 2 ? static_cast<uint64_t>(a.b_seg) + a.K : a.K;
  const CUtensorMapSwizzle swz = BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const uint64_t a_rows = static_cast<uint64_t>(a.a_brows) * (a.nbatch - 1) + a.M;
  const uint64_t b_rows = static_cast<uint64_t>(a.b_brows) * (a.nbatch - 1) + a.N;
  if (!make_tmap_2d_f16(&tmA, a.A, a_rows, a_cols, static_cast<uint64_t>(a.lda) * 2, SHARE_A ? BM / CN : BM, BK, swz) ||
      !make_tmap_2d_f16(&tmB, a.B, b_rows, b_cols, static_cast<uint64_t>(a.ldb) * 2, SHARE_A ? BN : BN / CN, BK, swz)) {
    set_error("gemm: cuTensorMapEncodeTiled failed");
    return -2;
  }
  const bool half_out = (a.epi == EPI_F16 || a.epi == EPI_GATE_F16);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) && (a.ldo % (half_out ? 8 : 4) == 0) &&
                     (a.out_seg % 8 == 0) && (a.out_b1 % 4 == 0) && (a.out_b2 % 4 == 0);
  const int stage_bytes = ((a.terms == 3 ? 2 : 1) * BM + (a.terms >= 2 ? 2 : 1) * BN) * BK * 2;
  int stages = SMEM_BUDGET / stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  const int tiles_m = (a.M + BM - 1) / BM, tiles_n = (a.N + BN - 1) / BN;
  GemmDev p{a.M, a.N, a.K, a.epi, a.act, a.bias, a.out, a.ldo, vec_ok, a.terms, a.a_seg, a.b_seg,
            half_out ? a.out_seg : 0, stages,
            (a.bias != nullptr && (reinterpret_cast<uintptr_t>(a.bias) & 15) == 0) ? 1 : 0,
            a.ln_gamma, a.ln_beta, a.ln_out, a.ln_ld, a.ln_seg, a.ln_counters, a.ln_epoch, tiles_m, tiles_n, a.bf16,
            a.nbatch, a.nb2, static_cast<int>(a.a_brows), static_cast<int>(a.b_brows), a.out_b1, a.out_b2, a.alpha};
  const int smem = stages * stage_bytes + 1024;
  // per device, once: SM count and the kernel's shared-memory opt-in (the backward pass makes ~1000 GEMM launches per
  // step: three runtime calls per launch were a visible share of the host's enqueue time)
  int dev = 0;
  HN_CHECK_CUDA(cudaGetDevice(&dev));
  static int sms_of[64] = {0};
  static bool attr_set[64] = {false};   // (per template instantiation)
  const int slot = dev & 63;
  if (sms_of[slot] == 0) HN_CHECK_CUDA(cudaDeviceGetAttribute(&sms_of[slot], cudaDevAttrMultiProcessorCount, dev));
  const int sms = sms_of[slot];
  if (!attr_set[slot]) {
    HN_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, BK, CN, SHARE_A>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM_BUDGET + 1024));
    attr_set[slot] = true;
  }
  const long n_super = static_cast<long>(tiles_m) * tiles_n * a.nbatch / CN;  // the caller guarantees divisibility
  const long max_clusters = sms / CN;
  const unsigned grid = static_cast<unsigned>((n_super < max_clusters ? n_super : max_clusters) * CN);
  // the fused-LayerNorm epilogue makes the CTAs of the grid wait for each other: launch it cooperatively so that the
  // whole grid is resident whatever else shares the GPU
  HN_CHECK_CUDA(launch_kc(gemm_kernel<BN, BK, CN, SHARE_A>, dim3(grid), dim3((2 + EPI_WARPS) * 32), smem, stream,
                          static_cast<unsigned>(CN), a.ln_out != nullptr, tmA, tmB, p));
  return 0;
}
}  // namespace

bool gemm_can_fuse_ln(const GemmArgs& a) {
  static int on = -1;  // HN_GEMM_LN=0 keeps the separate LayerNorm kernel (A/B measurements)
  if (on < 0) {
    const char* e = getenv("HN_GEMM_LN");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  const bool resid = a.epi == EPI_RES || a.epi == EPI_RES_LEAKY;
  if (!(on && resid && a.ldo == a.N && a.N <= 512 && a.N % 4 == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0))
    return false;
  // every CTA must hold exactly one 128 x 64 tile (they wait for each other): tile count <= SM count, no clusters
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return false;
  const long mt = (a.M + BM - 1) / BM;
  return mt * ((a.N + 63) / 64) <= sms && a.M < 16384 && getenv("HN_GEMM_BN") == nullptr && getenv("HN_GEMM_BK") == nullptr;
}

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  HN_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  HN_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm: operand pitches must be multiples of 8 elements");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  if (a.epi == EPI_GATE_F16) HN_REQUIRE(a.N % 2 == 0, "gemm: gated epilogue needs even N");
  HN_REQUIRE(a.terms >= 1 && a.terms <= 3, "gemm: terms must be 1, 2 or 3");
  HN_REQUIRE(!a.bf16 || a.epi == EPI_RES || a.epi == EPI_F32, "gemm: bf16 operands go with fp32 epilogues");
  HN_REQUIRE(a.nbatch >= 1 && a.nb2 >= 1, "gemm: bad batch count");
  if (a.nbatch > 1) {
    HN_REQUIRE((a.epi == EPI_RES || a.epi == EPI_F32) && a.ln_out == nullptr && a.bias == nullptr,
               "gemm: the batched form has plain fp32 epilogues only");
    HN_REQUIRE(a.a_brows >= a.M && a.b_brows >= a.N &&
                   a.a_brows * a.nbatch < 2147483647L && a.b_brows * a.nbatch < 2147483647L &&
                   static_cast<long>((a.M + BM - 1) / BM) * ((a.N + 63) / 64) * a.nbatch < 2147483647L,
               "gemm: batched operand rows out of range");
  }
  HN_REQUIRE(a.alpha == 1.f || a.epi == EPI_RES || a.epi == EPI_F32, "gemm: alpha goes with the plain fp32 epilogues");
  if (a.ln_out != nullptr) {
    const bool resid = a.epi == EPI_RES || a.epi == EPI_RES_LEAKY;
    const bool aligned = (reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.ln_out) & 7) == 0 &&
                         a.ln_ld % 4 == 0 && a.ln_seg % 4 == 0 && (reinterpret_cast<uintptr_t>(a.ln_gamma) & 15) == 0 &&
                         (reinterpret_cast<uintptr_t>(a.ln_beta) & 15) == 0;
    HN_REQUIRE(resid && a.ldo == a.N && a.N <= 512 && a.N % 4 == 0 && aligned && a.ln_counters != nullptr &&
                   a.ln_epoch >= 1 && gemm_can_fuse_ln(a),
               "gemm: fused LayerNorm needs a residual epilogue over complete rows of <= 512 columns (see gemm_can_fuse_ln)");
  }
  if (a.terms >= 2) HN_REQUIRE(a.b_seg % 64 == 0 && a.b_seg >= a.K, "gemm: B lo segment must start at a multiple of 64 >= K");
  if (a.terms == 3) HN_REQUIRE(a.a_seg % 64 == 0 && a.a_seg >= a.K, "gemm: A lo segment must start at a multiple of 64 >= K");
  // widest tile that still gives (nearly) every SM one
  const long mt = (a.M + BM - 1) / BM;
  const long tiles256 = static_cast<long>((a.N + 255) / 256) * mt * a.nbatch;
  const long tiles128 = static_cast<long>((a.N + 127) / 128) * mt * a.nbatch;
  static int force = -1, force_bk = -1;  // tuning knobs: HN_GEMM_BN=64|128|256, HN_GEMM_BK=32|64
  if (force < 0) {
    const char* e = getenv("HN_GEMM_BN");
    force = e ? atoi(e) : 0;
    const char* k = getenv("HN_GEMM_BK");
    force_bk = k ? atoi(k) : 0;
  }
  const int bn = force ? force : (tiles256 >= 120 ? 256 : tiles128 >= 120 ? 128 : 64);
  const int bk = force_bk ? force_bk : 64;  // 128-byte operand rows: half the L2 requests of BK = 32 (measured faster for every shape)
  // clusters (HN_GEMM_CLUSTER=0 turns them off, =1 forces them for every size): along N sharing the activations for
  // the narrow-output GEMMs, along M sharing the weights for the wide FF1 tile; only when the tile grid divides
  // evenly. Measured (tools/bench_gemm.py): at 2048 latent rows (batch 4) these GEMMs are latency-bound and the
  // cluster hand-shakes cost 2-3 %; from ~16 k rows on they are operand-feed-bound and multicast gains 2-3 %.
  static int cl = -1;
  if (cl < 0) {
    const char* e = getenv("HN_GEMM_CLUSTER");
    cl = e ? atoi(e) : 4;
    if (cl == 1) cl = -4;  // forced
  }
  const int tn = (a.N + bn - 1) / bn;
  const bool want_cluster = a.nbatch == 1 && (cl < 0 || (cl > 1 && a.M >= 16384));
  if (bk == 64 && want_cluster) {
    if (bn == 64 && tn % 4 == 0) return launch_t<64, 64, 4, true>(a, stream);
    if (bn == 64 && tn % 2 == 0) return launch_t<64, 64, 2, true>(a, stream);
    if (bn == 128 && tn % 2 == 0) return launch_t<128, 64, 2, true>(a, stream);
    if (bn == 256 && mt % 2 == 0) return launch_t<256, 64, 2, false>(a, stream);
  }
  if (bn == 256) return bk == 64 ? launch_t<256, 64, 1, true>(a, stream) : launch_t<256, 32, 1, true>(a, stream);
  if (bn == 128) return bk == 64 ? launch_t<128, 64, 1, true>(a, stream) : launch_t<128, 32, 1, true>(a, stream);
  return bk == 64 ? launch_t<64, 64, 1, true>(a, stream) : launch_t<64, 32, 1, true>(a, stream);
}

}  // namespace hn
