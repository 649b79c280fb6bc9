// xattn.cu — streaming (flash-style, split-N) softmax attention of a 128-row latent tile against a long token
// axis, on tcgen05 tensor cores with TMEM accumulators and TMA-staged token tiles.
//
// Replaces Attention.forward's  sim = q k^T * dh^-0.5 ; attn = softmax(sim / 0.5) ; out = attn v
// (reference healnet/models/healnet.py:409-424) without ever materialising sim / attn in HBM.
//
// Two instantiations of one kernel:
//   KD = 64 "generic"  : per-head K and V tiles (64 tokens x 64) — latent self-attention and wide-context
//                        modalities (WSI patch features, tabular rows).
//   KD = 32 "small-C"  : the reassociated form for narrow contexts (C <= 31: image / volume voxels).
//                        With LN(c) = gamma*z + beta:  q.k_t = (Wk' q).z_t + const  and
//                        sum_t p_t v_t = Wv' (sum_t p_t z_t) + Wv beta,  so ONE 64x32 tile of standardised
//                        context z serves as "K" and "V" of every head; Q' = Wk'^T q is 32 wide. Column C of
//                        z is 1.0, so accumulator column C is the softmax denominator (computed by the UMMA).
//
// CTA = (split, sample, head, 128-row latent tile); 6 warps:
//   warp 0   : TMA producer (Q tile once; K/V tiles through a 4-stage mbarrier ring)
//   warp 1   : UMMA issuer (one thread): S = Q K^T (SS) into a 2-deep TMEM ring; U += P V (TS, P from TMEM)
//   warps 2-5: softmax, one thread per latent row (TMEM lane): S -> exp2(S - m_ref) -> fp16 P back to TMEM.
//              The running max is only a reference point: it is raised lazily (when a tile exceeds it by
//              > 2^8), in which case the owning warp rescales its 32 accumulator rows in TMEM itself.
// Scores arrive pre-multiplied by 2/sqrt(dh) * log2(e) (folded into the Q projection), so the softmax is
// a bare ex2. Each CTA writes un-normalised (acc, m, l) partials; combine kernels (rowops.cu) merge splits.
#include "common.cuh"
#include "tc05.cuh"

namespace hn {
namespace {
using namespace tc05;

constexpr int BM = 128;     // latent rows per CTA
constexpr int BT = 64;      // tokens per tile
constexpr float RESCALE_THRESHOLD = 8.f;  // log2 units: P may reach 2^8 before the reference max is raised

__device__ __forceinline__ float ex2_approx(float x) {  // MUFU.EX2; ex2(-inf) = +0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnDev {
  int L, H, batch, nsplit, n_ltiles;
  int tiles_total;  // ceil(N / 64)
  long N;
  int k_col0, v_col0;
  int q_lo_off, kv_lo_off;  // precise mode: column offsets of the lo parts
  int v_lo;                 // precise mode: also contract P with the lo parts of V (short token axes)
  const uint64_t* mask_bits;
  float* part_acc;
  float* part_ml;
  __half* out;  // nsplit == 1: normalised output rows (see AttnArgs::out), else null
  int out_ld, out_lo_seg;
};

// PREC (generic path only): operands arrive split as hi + lo fp16 pairs; S = Qh.Kh + Ql.Kh + Qh.Kl and
// U += P.Vh + P.Vl. Used for short token axes (latent self-attention, tabular rows, small bags of patches),
// where per-token fp16 rounding of Q/K/V would not average out under the softmax.
// NA = 64-column atoms per head (generic path): 1 for dim_head <= 64, 2 for dim_head <= 128 (head pitch 128):
// S accumulates over both atoms, U is NA * 64 columns wide (one N = 64 UMMA per atom and k-step).
template <int KD, bool SHARED, bool PREC, int NA>
__global__ void __launch_bounds__(192, NA == 1 ? 2 : 1)
attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, AttnDev p) {
  static_assert(!(SHARED && PREC), "the precise mode exists on the generic path only");
  static_assert(NA == 1 || (!SHARED && KD == 64), "two atoms per head exist on the generic path only");
  constexpr int NSTAGE = (PREC || NA == 2) ? 2 : 4;  // K/V smem ring depth
  constexpr int VD = KD * NA;
  constexpr int HP = KD * NA;                        // head pitch in columns
  constexpr int Q_TILE = BM * KD * 2;                // one atom of Q
  constexpr int Q_BYTES = (PREC ? 2 : 1) * NA * Q_TILE;   // [Q atoms | Q_lo atoms]
  constexpr int K_BYTES = BT * KD * 2;               // one atom of K or V
  // stage: [K atoms | V atoms | K_lo atoms | V_lo atoms]
  constexpr int STAGE_BYTES = SHARED ? K_BYTES : (PREC ? 4 : 2) * NA * K_BYTES;
  constexpr uint32_t TCOLS = (NA == 1) ? 256 : 512;  // S 2x64 | P 2x32 | U VD
  constexpr uint32_t LAYOUT = (KD == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * KD * 2;        // 8-row group pitch of a swizzled tile
  constexpr uint32_t V_KADV = 16 * KD * 2;    // MN-major B: 16 tokens (one UMMA K step) further down

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sKV = smem + Q_BYTES;
  __shared__ uint64_t q_full, kv_full[NSTAGE], kv_empty[NSTAGE], s_full[2], p_ready[2], u_done, acc_done;
  __shared__ uint32_t tmem_base_s;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int idx = blockIdx.x;
  const int lt = idx % p.n_ltiles;
  idx /= p.n_ltiles;
  const int h = idx % p.H;
  idx /= p.H;
  const int b = idx % p.batch;
  const int split = idx / p.batch;
  const int t_begin = static_cast<int>(static_cast<long>(p.tiles_total) * split / p.nsplit);
  const int t_end = static_cast<int>(static_cast<long>(p.tiles_total) * (split + 1) / p.nsplit);
  const int n = t_end - t_begin;

  const long part_row0 = ((static_cast<long>(b) * p.nsplit + split) * p.H + h) * p.L + lt * BM;
  if (n <= 0) {  // more splits than tiles: publish an empty partial
    HN_PDL_WAIT();
    for (int r = threadIdx.x; r < BM; r += blockDim.x) {
      if (lt * BM + r < p.L) {
        float* acc = p.part_acc + (part_row0 + r) * VD;
        for (int c = 0; c < VD; ++c) acc[c] = 0.f;
        p.part_ml[(part_row0 + r) * 2] = -INFINITY;
        p.part_ml[(part_row0 + r) * 2 + 1] = 0.f;
      }
    }
    return;
  }

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 4);
    }
    mbar_init(&u_done, 1);
    mbar_init(&acc_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<TCOLS>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tS0 = tmem, tP0 = tmem + 128, tU = tmem + 192;
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmKV);
      mbar_arrive_expect_tx(&q_full, Q_BYTES);
#pragma unroll
      for (int a = 0; a < NA; ++a) {
        tma_load_3d(sQ + a * Q_TILE, &tmQ, &q_full, h * HP + a * KD, lt * BM, b);
        if (PREC) tma_load_3d(sQ + (NA + a) * Q_TILE, &tmQ, &q_full, p.q_lo_off + h * HP + a * KD, lt * BM, b);
      }
      const int kcol = SHARED ? 0 : p.k_col0 + h * HP;
      const int vcol = SHARED ? 0 : p.v_col0 + h * HP;
      for (int i = 0; i < n; ++i) {
        const int s = i % NSTAGE;
        mbar_wait(&kv_empty[s], ((i / NSTAGE) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], (PREC && !p.v_lo) ? STAGE_BYTES - NA * K_BYTES : STAGE_BYTES);
        const int tok0 = (t_begin + i) * BT;
        uint8_t* st = sKV + s * STAGE_BYTES;
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          tma_load_3d(st + a * K_BYTES, &tmKV, &kv_full[s], kcol + a * KD, tok0, b);
          if (!SHARED) tma_load_3d(st + (NA + a) * K_BYTES, &tmKV, &kv_full[s], vcol + a * KD, tok0, b);
          if (PREC) {
            tma_load_3d(st + (2 * NA + a) * K_BYTES, &tmKV, &kv_full[s], p.kv_lo_off + kcol + a * KD, tok0, b);
            if (p.v_lo) tma_load_3d(st + (3 * NA + a) * K_BYTES, &tmKV, &kv_full[s], p.kv_lo_off + vcol + a * KD, tok0, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer
    if (elect_one()) {
      constexpr uint32_t idesc_s = idesc_f16(BM, BT, false, false);  // S[128x64]  = Q[128xKD] . K[64xKD]^T
      constexpr uint32_t idesc_u = idesc_f16(BM, KD, false, true);   // U[128xKD] += P[128x64] . V[64xKD] per atom (MN-major B)
      const uint32_t q0 = smem_u32(sQ);
      auto issue_s = [&](int j) {
        const int s = j % NSTAGE;
        mbar_wait(&kv_full[s], (j / NSTAGE) & 1);
        fence_after_sync();
        const uint32_t k0 = smem_u32(sKV + s * STAGE_BYTES);
#pragma unroll
        for (int a = 0; a < NA; ++a) {
#pragma unroll
          for (int k = 0; k < KD / 16; ++k)
            umma_ss(tS0 + (j & 1) * 64, smem_desc(q0 + a * Q_TILE + k * 32, 16, SBO, LAYOUT),
                    smem_desc(k0 + a * K_BYTES + k * 32, 16, SBO, LAYOUT), idesc_s, (a | k) != 0);
          if (PREC) {
#pragma unroll
            for (int k = 0; k < KD / 16; ++k)  // Q_lo . K_hi
              umma_ss(tS0 + (j & 1) * 64, smem_desc(q0 + (NA + a) * Q_TILE + k * 32, 16, SBO, LAYOUT),
                      smem_desc(k0 + a * K_BYTES + k * 32, 16, SBO, LAYOUT), idesc_s, true);
#pragma unroll
            for (int k = 0; k < KD / 16; ++k)  // Q_hi . K_lo
              umma_ss(tS0 + (j & 1) * 64, smem_desc(q0 + a * Q_TILE + k * 32, 16, SBO, LAYOUT),
                      smem_desc(k0 + (2 * NA + a) * K_BYTES + k * 32, 16, SBO, LAYOUT), idesc_s, true);
          }
        }
        umma_commit(&s_full[j & 1]);
      };
      mbar_wait(&q_full, 0);
      issue_s(0);
      if (n > 1) issue_s(1);
      for (int i = 0; i < n; ++i) {
        const int s = i % NSTAGE;
        mbar_wait(&p_ready[i & 1], (i >> 1) & 1);
        fence_after_sync();
        const uint32_t v0 = smem_u32(sKV + s * STAGE_BYTES + (SHARED ? 0 : NA * K_BYTES));
#pragma unroll
        for (int a = 0; a < NA; ++a) {
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)
            umma_ts(tU + a * KD, tP0 + (i & 1) * 32 + k * 8, smem_desc(v0 + a * K_BYTES + k * V_KADV, 16, SBO, LAYOUT),
                    idesc_u, (i | k) != 0);
          if (PREC && p.v_lo) {
#pragma unroll
            for (int k = 0; k < BT / 16; ++k)  // P . V_lo
              umma_ts(tU + a * KD, tP0 + (i & 1) * 32 + k * 8,
                      smem_desc(v0 + (2 * NA + a) * K_BYTES + k * V_KADV, 16, SBO, LAYOUT), idesc_u, true);
          }
        }
        umma_commit(&kv_empty[s]);
        umma_commit(&u_done);
        if (i + 1 == n) umma_commit(&acc_done);  // u_done's parity alone cannot tell PV(n-1) from PV(n-3)
        if (i + 2 < n) issue_s(i + 2);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax warps: thread = latent row
    const uint32_t lane_base = (warp & 3) * 32;
    const int row = lt * BM + lane_base + lane;
    float m_ref = -INFINITY;
    float l_sum = 0.f;
    for (int i = 0; i < n; ++i) {
      const int buf = i & 1;
      const int tile = t_begin + i;
      uint64_t bits = ~0ull;
      if (p.mask_bits != nullptr) bits = p.mask_bits[static_cast<long>(b) * p.tiles_total + tile];
      const long rem = p.N - static_cast<long>(tile) * BT;
      if (rem < BT) bits &= (1ull << rem) - 1ull;

      mbar_wait(&s_full[buf], (i >> 1) & 1);
      fence_after_sync();
      uint32_t s0[32], s1[32];
      tmem_ld32(tmem_addr(tS0 + buf * 64, lane_base, 0), s0);
      tmem_ld32(tmem_addr(tS0 + buf * 64, lane_base, 32), s1);
      tmem_wait_ld();
      if (bits != ~0ull) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (!((bits >> j) & 1ull)) s0[j] = __float_as_uint(-INFINITY);
          if (!((bits >> (32 + j)) & 1ull)) s1[j] = __float_as_uint(-INFINITY);
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(s0[j]), __uint_as_float(s1[j])));

      if (__any_sync(0xffffffffu, mx > m_ref + RESCALE_THRESHOLD)) {
        // raise the reference max (rare after the first tile); rescale this warp's accumulator rows
        const float m_new = fmaxf(m_ref, mx);
        if (i > 0) {
          const float sc = (m_new == -INFINITY) ? 1.f : ex2_approx(m_ref - m_new);
          mbar_wait(&u_done, (i - 1) & 1);  // PV(i-1) has landed; PV(i) cannot start before our arrive
          fence_after_sync();
#pragma unroll
          for (int c = 0; c < VD; c += 32) {
            uint32_t u[32];
            tmem_ld32(tmem_addr(tU, lane_base, c), u);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(__uint_as_float(u[j]) * sc);
            tmem_st32(tmem_addr(tU, lane_base, c), u);
          }
          tmem_wait_st();
          l_sum *= sc;
        }
        m_ref = m_new;
      }
      const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
      uint32_t pk[32];
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e0 = ex2_approx(__uint_as_float(s0[2 * j]) - m_use);
        const float e1 = ex2_approx(__uint_as_float(s0[2 * j + 1]) - m_use);
        if (!SHARED) rs += e0 + e1;
        pk[j] = pack_half2(e0, e1);
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float e0 = ex2_approx(__uint_as_float(s1[2 * j]) - m_use);
        const float e1 = ex2_approx(__uint_as_float(s1[2 * j + 1]) - m_use);
        if (!SHARED) rs += e0 + e1;
        pk[16 + j] = pack_half2(e0, e1);
      }
      l_sum += rs;
      tmem_st32(tmem_addr(tP0 + buf * 32, lane_base, 0), pk);
      tmem_wait_st();
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[buf]);
    }
    // ---- epilogue: un-normalised accumulator rows + (m, l)
    mbar_wait(&acc_done, 0);
    fence_after_sync();
    const float inv_l = 1.f / l_sum;  // (a fully masked row gives NaN, like the reference's softmax over -inf)
#pragma unroll
    for (int c = 0; c < VD; c += 32) {
      uint32_t u[32];
      tmem_ld32(tmem_addr(tU, lane_base, c), u);
      tmem_wait_ld();
      if (row < p.L && p.out != nullptr) {
        // single split: normalise and store the fp16 (hi | lo) output row directly
        __half* o = p.out + (static_cast<long>(b) * p.L + row) * p.out_ld + h * VD + c;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 qh, ql;
          uint32_t* ph = reinterpret_cast<uint32_t*>(&qh);
          uint32_t* pl = reinterpret_cast<uint32_t*>(&ql);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float v0 = __uint_as_float(u[j + 2 * k]) * inv_l, v1 = __uint_as_float(u[j + 2 * k + 1]) * inv_l;
            const __half2 hi = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(hi);
            const __half2 lo = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
            ph[k] = *reinterpret_cast<const uint32_t*>(&hi);
            pl[k] = *reinterpret_cast<const uint32_t*>(&lo);
          }
          *reinterpret_cast<uint4*>(o + j) = qh;
          if (p.out_lo_seg > 0) *reinterpret_cast<uint4*>(o + p.out_lo_seg + j) = ql;
        }
      } else if (row < p.L) {
        float4* dst = reinterpret_cast<float4*>(p.part_acc + (part_row0 + lane_base + lane) * VD + c);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          dst[j] = make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]), __uint_as_float(u[4 * j + 2]),
                               __uint_as_float(u[4 * j + 3]));
      }
      __syncwarp();
    }
    if (row < p.L) {
      float2* ml = reinterpret_cast<float2*>(p.part_ml + (part_row0 + lane_base + lane) * 2);
      *ml = make_float2(m_ref, l_sum);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TCOLS>(tmem);
}

template <int KD, bool SHARED, bool PREC, int NA>
int launch_t(const AttnArgs& a, cudaStream_t stream) {
  CUtensorMap tmQ, tmKV;
  const CUtensorMapSwizzle swz = KD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (!make_tmap_3d_f16(&tmQ, a.Q, a.batch, a.L, a.q_ld, static_cast<uint64_t>(a.q_ld) * 2,
                        static_cast<uint64_t>(a.L) * a.q_ld * 2, BM, KD, swz) ||
      !make_tmap_3d_f16(&tmKV, a.KV, a.batch, a.N, a.kv_ld, static_cast<uint64_t>(a.kv_ld) * 2,
                        static_cast<uint64_t>(a.N) * a.kv_ld * 2, BT, KD, swz)) {
    set_error("attention: cuTensorMapEncodeTiled failed");
    return -2;
  }
  AttnDev p;
  p.L = a.L;
  p.H = a.H;
  p.batch = a.batch;
  p.nsplit = a.nsplit;
  p.n_ltiles = (a.L + BM - 1) / BM;
  p.tiles_total = static_cast<int>((a.N + BT - 1) / BT);
  p.N = a.N;
  p.k_col0 = a.k_col0;
  p.v_col0 = a.v_col0;
  p.q_lo_off = a.q_lo_off;
  p.kv_lo_off = a.kv_lo_off;
  p.v_lo = a.v_hi_only ? 0 : 1;
  p.mask_bits = a.mask_bits;
  p.part_acc = a.part_acc;
  p.part_ml = a.part_ml;
  p.out = (a.nsplit == 1 && !SHARED) ? a.out : nullptr;
  p.out_ld = a.out_ld;
  p.out_lo_seg = a.out_lo_seg;
  if (p.out != nullptr)
    HN_REQUIRE((reinterpret_cast<uintptr_t>(a.out) & 15) == 0 && a.out_ld % 8 == 0 && a.out_lo_seg % 8 == 0,
               "attention: direct output rows must be 16-byte aligned");
  constexpr int NSTAGE = (PREC || NA == 2) ? 2 : 4;
  constexpr int SMEM = (PREC ? 2 : 1) * NA * BM * KD * 2 +
                       NSTAGE * (SHARED ? 1 : (PREC ? 4 : 2) * NA) * BT * KD * 2 + 1024;
  HN_CHECK_CUDA(
      cudaFuncSetAttribute(attn_kernel<KD, SHARED, PREC, NA>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  const long grid = static_cast<long>(p.n_ltiles) * a.H * a.batch * a.nsplit;
  HN_REQUIRE(grid > 0 && grid < 2147483647L, "attention: grid too large");
  HN_CHECK_CUDA(launch_k(attn_kernel<KD, SHARED, PREC, NA>, dim3(static_cast<unsigned>(grid)), dim3(192), SMEM, stream,
                         tmQ, tmKV, p));
  return 0;
}
}  // namespace

// Split the token axis so the grid covers the chip several times over (2 CTAs per SM, 148 SMs) while each
// CTA still streams enough tiles to amortise its prologue.
int attention_pick_nsplit(int batch, int L, int H, long N) {
  const long base = static_cast<long>((L + BM - 1) / BM) * H * batch;
  const long tiles = (N + BT - 1) / BT;
  const long slots = 2 * 148;
  if (tiles <= 16) return 1;
  long best = 1;
  double best_cost = 1e30;
  const long max_split = tiles / 8 > 0 ? tiles / 8 : 1;
  for (long s = 1; s <= max_split && s <= 512; ++s) {
    const long ctas = base * s;
    const long waves = (ctas + slots - 1) / slots;
    const long per = (tiles + s - 1) / s;
    // time ~ waves * (tiles per CTA + fixed prologue/epilogue cost of ~6 tiles)
    const double cost = static_cast<double>(waves) * (per + 6.0);
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = s;
    }
  }
  return static_cast<int>(best);
}

int launch_attention(const AttnArgs& a, cudaStream_t stream) {
  HN_REQUIRE(a.batch > 0 && a.L > 0 && a.H > 0 && a.N > 0 && a.nsplit > 0, "attention: empty problem");
  HN_REQUIRE(a.q_ld % 8 == 0 && a.kv_ld % 8 == 0, "attention: row pitches must be multiples of 8 elements");
  HN_REQUIRE(a.N < (1L << 31), "attention: token axis too long");
  if (a.shared_kv) {
    HN_REQUIRE(a.kd == 32 || a.kd == 64, "attention: shared-context rows must be 32 or 64 wide");
    return launch_small_attention(a, stream);
  }
  HN_REQUIRE(a.hp == 64 || a.hp == 128, "attention: head pitch must be 64 or 128");
  if (a.precise) {
    HN_REQUIRE(a.q_lo_off > 0 && a.kv_lo_off > 0, "attention: precise mode needs the lo-part offsets");
    return a.hp == 64 ? launch_t<64, false, true, 1>(a, stream) : launch_t<64, false, true, 2>(a, stream);
  }
  return a.hp == 64 ? launch_t<64, false, false, 1>(a, stream) : launch_t<64, false, false, 2>(a, stream);
}

}  // namespace hn
