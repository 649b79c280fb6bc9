// probe.cu — one-CTA unit kernel that exercises exactly the UMMA/TMA/TMEM conventions the attention and
// GEMM kernels rely on: S = Q·Kᵀ (SS, K-major A and B, TMA-swizzled tiles), P = fp16(S) written back to
// TMEM, U = P·V (TS: A from TMEM, B MN-major from the same kind of TMA tile). Test-only entry point
// (hn_debug_probe); descriptor fields can be overridden from the host so one GPU call can sweep variants.
#include "../../include/healnet_b200.h"
#include "tc05.cuh"

namespace {
using namespace tc05;

struct ProbeCfg {
  int kd, vd;          // inner dims: Q/K width, V width (32 -> 64B swizzle, 64 -> 128B swizzle)
  int q_lbo, q_sbo, k_lbo, k_sbo, v_lbo, v_sbo;  // descriptor byte offsets
  int qk_kadv, v_kadv;  // start-address advance (bytes) per UMMA K step (16 elements)
  int q_layout, v_layout;
};

constexpr int BM = 128;  // latent rows
constexpr int BN = 64;   // tokens per tile

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
             const __grid_constant__ CUtensorMap tmV, ProbeCfg cfg, float* __restrict__ S_out,
             float* __restrict__ U_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // 128 x kd fp16  (<= 16 KB)
  uint8_t* sK = smem + 16384;         // 64 x kd fp16   (<= 8 KB)
  uint8_t* sV = smem + 16384 + 8192;  // 64 x vd fp16   (<= 8 KB)
  __shared__ uint64_t bar_load, bar_s, bar_u;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_s, 1);
    mbar_init(&bar_u, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tS = tmem + 0, tP = tmem + 64, tU = tmem + 128;

  if (threadIdx.x == 0) {
    const uint32_t bytes = (BM * cfg.kd + BN * cfg.kd + BN * cfg.vd) * 2;
    mbar_arrive_expect_tx(&bar_load, bytes);
    tma_load_2d(sQ, &tmQ, &bar_load, 0, 0);
    tma_load_2d(sK, &tmK, &bar_load, 0, 0);
    tma_load_2d(sV, &tmV, &bar_load, 0, 0);
    mbar_wait(&bar_load, 0);
    fence_after_sync();
    const uint32_t id = idesc_f16(BM, BN, false, false);
    for (int k = 0; k < cfg.kd / 16; ++k) {
      uint64_t da = smem_desc(smem_u32(sQ) + k * cfg.qk_kadv, cfg.q_lbo, cfg.q_sbo, cfg.q_layout);
      uint64_t db = smem_desc(smem_u32(sK) + k * cfg.qk_kadv, cfg.k_lbo, cfg.k_sbo, cfg.q_layout);
      umma_ss(tS, da, db, id, k > 0);
    }
    umma_commit(&bar_s);
  }
  __syncwarp();
  mbar_wait(&bar_s, 0);
  fence_after_sync();

  const uint32_t lane_base = (warp & 3) * 32;
  const int row = lane_base + lane;
  {
    uint32_t s[32];
    uint32_t p[32];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      tmem_ld32(tmem_addr(tS, lane_base, half * 32), s);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 32; ++j) S_out[row * BN + half * 32 + j] = __uint_as_float(s[j]);
#pragma unroll
      for (int j = 0; j < 16; ++j)
        p[half * 16 + j] = pack_half2(__uint_as_float(s[2 * j]), __uint_as_float(s[2 * j + 1]));
    }
    tmem_st32(tmem_addr(tP, lane_base, 0), p);
    tmem_wait_st();
  }
  fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_after_sync();
    const uint32_t id = idesc_f16(BM, cfg.vd, false, true);
    for (int k = 0; k < BN / 16; ++k) {
      uint64_t db = smem_desc(smem_u32(sV) + k * cfg.v_kadv, cfg.v_lbo, cfg.v_sbo, cfg.v_layout);
      umma_ts(tU, tP + k * 8, db, id, k > 0);
    }
    umma_commit(&bar_u);
  }
  __syncwarp();
  mbar_wait(&bar_u, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < cfg.vd; c0 += 32) {
    uint32_t u[32];
    tmem_ld32(tmem_addr(tU, lane_base, c0), u);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) U_out[row * cfg.vd + c0 + j] = __uint_as_float(u[j]);
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}
}  // namespace

// Q: [128][kd] fp16, K: [64][kd] fp16, V: [64][vd] fp16 (row-major, device). S_out [128][64], U_out [128][vd] fp32.
// ov: optional 10 ints overriding {q_lbo,q_sbo,k_lbo,k_sbo,v_lbo,v_sbo,qk_kadv,v_kadv,q_layout,v_layout}; <0 keeps default.
extern "C" __attribute__((visibility("default"))) int hn_debug_probe(const void* Q, const void* K, const void* V, int kd, int vd, float* S_out,
                              float* U_out, const int* ov, void* stream) {
  if (!((kd == 32 || kd == 64) && (vd == 32 || vd == 64))) return -1;
  CUtensorMap tq, tk, tv;
  CUtensorMapSwizzle sq = kd == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUtensorMapSwizzle sv = vd == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (!tc05::make_tmap_2d_f16(&tq, Q, 128, kd, kd * 2, 128, kd, sq)) return -2;
  if (!tc05::make_tmap_2d_f16(&tk, K, 64, kd, kd * 2, 64, kd, sq)) return -2;
  if (!tc05::make_tmap_2d_f16(&tv, V, 64, vd, vd * 2, 64, vd, sv)) return -2;
  ProbeCfg c;
  c.kd = kd;
  c.vd = vd;
  c.q_lbo = 16;
  c.q_sbo = 8 * kd * 2;
  c.k_lbo = 16;
  c.k_sbo = 8 * kd * 2;
  c.v_lbo = 16;
  c.v_sbo = 8 * vd * 2;
  c.qk_kadv = 32;
  c.v_kadv = 16 * vd * 2;
  c.q_layout = kd == 64 ? tc05::SWZ_128B : tc05::SWZ_64B;
  c.v_layout = vd == 64 ? tc05::SWZ_128B : tc05::SWZ_64B;
  if (ov) {
    int* f[10] = {&c.q_lbo, &c.q_sbo, &c.k_lbo, &c.k_sbo, &c.v_lbo, &c.v_sbo, &c.qk_kadv, &c.v_kadv, &c.q_layout,
                  &c.v_layout};
    for (int i = 0; i < 10; ++i)
      if (ov[i] >= 0) *f[i] = ov[i];
  }
  const int smem = 16384 + 8192 + 8192 + 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(tq, tk, tv, c, S_out, U_out);
  return cudaGetLastError() == cudaSuccess ? 0 : -3;
}
