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
// One CTA computes a 128 x BN output tile: warp 0 = TMA producer (128B-swizzled K-major tiles of 64
// fp16), warp 1 = single-thread UMMA issuer (+ TMEM allocator), warps 2..5 = epilogue (each thread owns
// one accumulator row = one TMEM lane). M/N/K tails are handled by TMA out-of-bounds zero fill plus
// guards in the epilogue, so any shape whose row pitches are multiples of 8 elements is legal.
#include <cstdlib>

#include "common.cuh"
#include "tc05.cuh"

namespace hn {
namespace {
using namespace tc05;

constexpr int BM = 128;
constexpr int BK = 64;

struct GemmDev {
  int M, N, K;
  int epi, act;
  const float* bias;
  void* out;
  int ldo;
  int vec_ok;  // output rows are 16-byte aligned -> vector stores allowed
  int terms;   // 1, 2 or 3 passes over the K tiles (see header comment)
  int a_seg, b_seg;  // column offset of the lo segment of A / B (elements)
  int out_seg;       // fp16 outputs: > 0 -> also store lo = fp16(v - hi) at column + out_seg
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

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                   const __grid_constant__ CUtensorMap tmB, GemmDev p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
  const int k_tiles_1 = (p.K + BK - 1) / BK;
  const int k_tiles = k_tiles_1 * p.terms;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&acc_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<BN>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tD = tmem_base_s;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      tma_prefetch_desc(&tmB);
      for (int kt = 0; kt < k_tiles; ++kt) {
        const int s = kt % STAGES;
        mbar_wait(&empty_bar[s], ((kt / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        const int pass = kt / k_tiles_1, k0 = (kt - pass * k_tiles_1) * BK;
        // pass 0: A_hi.B_hi; terms == 3: pass 1 = A_lo.B_hi, pass 2 = A_hi.B_lo; terms == 2: pass 1 = A_hi.B_lo
        const int a_col = k0 + ((p.terms == 3 && pass == 1) ? p.a_seg : 0);
        const int b_col = k0 + ((pass == p.terms - 1 && pass > 0) ? p.b_seg : 0);
        tma_load_2d(smem + s * STAGE_BYTES, &tmA, &full_bar[s], a_col, m0);
        tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &tmB, &full_bar[s], b_col, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      constexpr uint32_t idesc = idesc_f16(BM, BN, false, false);
      for (int kt = 0; kt < k_tiles; ++kt) {
        const int s = kt % STAGES;
        mbar_wait(&full_bar[s], (kt / STAGES) & 1);
        fence_after_sync();
        const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t b0 = a0 + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          umma_ss(tD, smem_desc(a0 + k * 32, 16, 1024, SWZ_128B), smem_desc(b0 + k * 32, 16, 1024, SWZ_128B),
                  idesc, (kt | k) != 0);
        }
        umma_commit(&empty_bar[s]);
      }
      umma_commit(&acc_bar);
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const uint32_t lane_base = (warp & 3) * 32;
    const int row = m0 + lane_base + lane;
    const bool row_ok = row < p.M;
    mbar_wait(&acc_bar, 0);
    fence_after_sync();
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_addr(tD, lane_base, c * 32), r);
      tmem_wait_ld();
      const int nb = n0 + c * 32;
      if (nb >= p.N) continue;
      const bool full = (nb + 32 <= p.N) && p.vec_ok;
      float v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float b = 0.f;
        if (p.bias != nullptr && nb + j < p.N) b = __ldg(p.bias + nb + j);
        v[j] = __uint_as_float(r[j]) + b;
      }
      if (row_ok) {
      if (p.epi == EPI_F16) {
        __half* o = reinterpret_cast<__half*>(p.out) + static_cast<size_t>(row) * p.ldo + nb;
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 q, ql;
            q.x = pack_hi2(v[j], v[j + 1], ql.x);
            q.y = pack_hi2(v[j + 2], v[j + 3], ql.y);
            q.z = pack_hi2(v[j + 4], v[j + 5], ql.z);
            q.w = pack_hi2(v[j + 6], v[j + 7], ql.w);
            *reinterpret_cast<uint4*>(o + j) = q;
            if (p.out_seg > 0) *reinterpret_cast<uint4*>(o + p.out_seg + j) = ql;
          }
        } else {
          for (int j = 0; j < 32 && nb + j < p.N; ++j) {
            __half hi, lo;
            split_half(v[j], hi, lo);
            o[j] = hi;
            if (p.out_seg > 0) o[p.out_seg + j] = lo;
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
        float* o = reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + nb;
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
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<BN>(tD);
}

template <int BN, int STAGES>
int launch_t(const GemmArgs& a, cudaStream_t stream) {
  CUtensorMap tmA, tmB;
  // single-segment operands rely on TMA zero fill beyond column K; split operands declare hi|lo and keep
  // explicit zeros in the pad columns [K, seg) of both segments
  const uint64_t a_cols = a.terms == 3 ? static_cast<uint64_t>(a.a_seg) + a.K : a.K;
  const uint64_t b_cols = a.terms >= 2 ? static_cast<uint64_t>(a.b_seg) + a.K : a.K;
  if (!make_tmap_2d_f16(&tmA, a.A, a.M, a_cols, static_cast<uint64_t>(a.lda) * 2, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B) ||
      !make_tmap_2d_f16(&tmB, a.B, a.N, b_cols, static_cast<uint64_t>(a.ldb) * 2, BN, BK, CU_TENSOR_MAP_SWIZZLE_128B)) {
    set_error("gemm: cuTensorMapEncodeTiled failed");
    return -2;
  }
  const bool half_out = (a.epi == EPI_F16 || a.epi == EPI_GATE_F16);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(a.out) & 15) == 0) && (a.ldo % (half_out ? 8 : 4) == 0) &&
                     (a.out_seg % 8 == 0);
  GemmDev p{a.M, a.N, a.K, a.epi, a.act, a.bias, a.out, a.ldo, vec_ok, a.terms, a.a_seg, a.b_seg,
            half_out ? a.out_seg : 0};
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024;
  // per device and cheap: set on every launch so a process driving several GPUs never misses it
  HN_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM);
  gemm_kernel<BN, STAGES><<<grid, 192, SMEM, stream>>>(tmA, tmB, p);
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  HN_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem");
  HN_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm: operand pitches must be multiples of 8 elements");
  HN_REQUIRE((reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0,
             "gemm: operands must be 16-byte aligned");
  if (a.epi == EPI_GATE_F16) HN_REQUIRE(a.N % 2 == 0, "gemm: gated epilogue needs even N");
  HN_REQUIRE(a.terms >= 1 && a.terms <= 3, "gemm: terms must be 1, 2 or 3");
  if (a.terms >= 2) HN_REQUIRE(a.b_seg % 64 == 0 && a.b_seg >= a.K, "gemm: B lo segment must start at a multiple of 64 >= K");
  if (a.terms == 3) HN_REQUIRE(a.a_seg % 64 == 0 && a.a_seg >= a.K, "gemm: A lo segment must start at a multiple of 64 >= K");
  // wide tiles once the grid would still cover the 148 SMs, narrow ones otherwise
  const long mt = (a.M + BM - 1) / BM;
  const long tiles128 = static_cast<long>((a.N + 127) / 128) * mt;
  static int force = -1;  // tuning knob: HN_GEMM_BN=64|128|256
  if (force < 0) {
    const char* e = getenv("HN_GEMM_BN");
    force = e ? atoi(e) : 0;
  }
  if (force == 256) return launch_t<256, 3>(a, stream);
  if (force == 128) return launch_t<128, 3>(a, stream);
  if (force == 64) return launch_t<64, 4>(a, stream);
  if (force == 648) return launch_t<64, 8>(a, stream);
  if (force == 1286) return launch_t<128, 6>(a, stream);
  if (tiles128 >= 148) return launch_t<128, 3>(a, stream);
  return launch_t<64, 4>(a, stream);
}

}  // namespace hn
