// xattn_small.cu — the hot kernel of the path: streaming softmax attention of the latent array against a LONG
// token axis of NARROW context rows (image pixels / volume voxels: C = channels + Fourier features <= 63).
//
// Replaces  sim = q k^T * dh^-0.5 ; attn = softmax(sim / 0.5) ; out = attn v   (healnet.py:409-424) for such
// modalities in the reassociated form (see xattn.cu): with LN(c) = gamma*z + beta,
//     q.k_t = Q'.z_t + const,   Q' = (Wk' ^T q) (KD = 32 | 64 wide),    sum_t p_t v_t = Wv' (sum_t p_t z_t) + Wv beta
// so ONE 64-token x KD tile of standardised context z is K and V of every head and latent tile. Column C of z
// is 1.0, so accumulator column C is the softmax denominator.
//
// The kernel is bound by the softmax exponentials (H*L*N per pass; MUFU.EX2 = 16/clk/SM measured), not by the
// tensor pipe, so it is organised around keeping the MUFU *and* FMA pipes of all four SM sub-partitions busy:
//   * one CTA per SM owns G = 3 (KD 32) / 2 (KD 64) "row blocks" (128 latent rows of one head) that stream the
//     SAME z tiles: 4 softmax warps per row block -> 12 softmax warps per SM, 3 per scheduler, so that one
//     group's TMEM / mbarrier latency is covered by the others; the z tile is fetched once per G row blocks;
//   * S is double-buffered in TMEM (P(i) is written over the first half of S(i), which frees the columns for a
//     second S buffer), so S(i+1) is computed while the softmax of tile i runs;
//   * softmax threads own one latent row (one TMEM lane) and work in 16-column chunks (four 16-register TMEM loads,
//     four 8-register stores per tile);
//   * half of the exponentials are evaluated on the FMA / ALU pipes, two at a time in half2 arithmetic (range
//     reduction with the magic-constant trick + degree-3 minimax polynomial + integer exponent insert, ~10 issue slots
//     per column pair incl. the fp16 pack: ex2_pair_h2_lean) so MUFU only sees the rest (pipe microbenchmark: 23
//     elem/clk/SM for a 3/8 fp32 mix vs 16 for MUFU alone, tools/microbench/mb_pipes.cu; HN_POLY_MODE sweeps the share in
//     debug builds; 1/2 measured best, also on the round-2 kernel);
//   * the running max is a lazily raised reference: the steady state does no max pass at all — it only tracks
//     the max of the packed fp16 P words of the MUFU lanes (VIMNMX3.U16x2) and the OR of the polynomial lanes' t words,
//     and falls back to the exact two-pass path when a P exceeds 2^15 (or on the first / a masked / the ragged last tile);
//   * the "- max" of the softmax costs nothing: Q' carries -m_ref (fp16) in the column where z holds its 1.0,
//     so the S UMMA delivers q.z - m_ref directly; the owning thread rewrites that smem element on the rare raise;
//   * what bounds it (ncu, profiles/r2b_*): no single pipe (issue 73 %, tensor-core instruction pipe 77 %, XU 56 %) but
//     the chain  P(i) published -> issuer wakes -> PV(i), S(i+2) on a busy tensor pipe -> softmax wakes.  Everything
//     that can leave that chain has: the mid-tile barrier test is non-blocking and pinned late (a hoisted blocking
//     try_wait held P(i) back until S(i+1) had landed), the issuer fetches its descriptors and waits for tile i+2's
//     context rows BEFORE it waits for P(i), and the score product only issues the 16-column steps that hold context
//     columns (merged tail for 17 <= C <= 23).
// Warp roles: warps 0..4G-1 = softmax groups (4 warps each); warps 4G..5G-1 = one UMMA issuer warp per group
// (S = Q'.z^T (SS) ; U += P.z (TS, P from TMEM)), each blocking on its own group's barrier — a single issuer
// polling several groups' mbarriers (~150 clk per test) starved the groups, and letting a softmax warp issue
// stretched that warp's tile and with it the whole group's; last warp = TMA producer (Q' tiles once, z tiles
// through an 8-stage mbarrier ring). TMEM per group: S/P buffer 0 (64) | S/P buffer 1 (64) | U (KD).
//
// SPLIT (the product path): the score operands arrive as fp16 hi + lo pairs — Q' rows [hi | lo at q_lo_off], z rows
// [hi (KD) | lo (KD)] — and S = Q'h.zh + Q'l.zh + Q'h.zl (three K blocks into the same TMEM accumulator). With single
// fp16 operands the score error grows like |s| * 2^-11 and does not average out once the softmax is peaked (measured at
// the full cfg 1 volume with to_q scaled x8 / x32: latent error 1e-3 / 9e-3, tests/test_gpu_fullsize.py); the split
// makes the scores fp32-exact for up to four extra UMMAs per tile (two at the volume's C = 18 with the merged tail, none
// beyond the three terms' single steps at the image's C = 13). P.z (the value side) keeps the hi part only: it is a
// convex combination whose fp16 rounding stays below 2^-11 relative.
#include <cstdlib>
#include <type_traits>

#include "bwd.cuh"
#include "common.cuh"
#include "tc05.cuh"

namespace hn {
namespace {
using namespace tc05;

constexpr int BM = 128;  // latent rows per row block
constexpr int BT = 64;   // tokens per tile
constexpr int NST = 8;   // z ring depth
constexpr int MAXG = 4;
// P is stored in fp16. A raise puts the reference AT the row max seen so far; it then stays put until some P exceeds
// 2^5 times the reference weight (or is inf / garbage), i.e. until the running max has grown by another 5 log2 units;
// that triggers the exact path, which raises it again.
// P' = 2^(s - m_ref + P_SHIFT): the tile the UMMA delivers already carries -m_ref + P_SHIFT (folded into Q'), so the
// half2 polynomial can build its result directly as a NORMAL fp16 (exponent field = rint(x') + 15 >= 1 for x' >= -14,
// i.e. down to 2^-24 of the reference) without the final 2^-10 rescale, and MUFU results use the top of the fp16
// range. Numerator and denominator (the ones column) carry the same 2^P_SHIFT, so it cancels in the combine kernels.
constexpr float P_SHIFT = 10.f;
constexpr uint32_t P_RAISE_BITS = 0x7800u;  // fp16(2^15) = 2^(RESCALE_THRESHOLD + P_SHIFT)
constexpr float RESCALE_THRESHOLD = 5.f;    // log2 units

struct SmallDev {
  int L, H, batch, nsplit, n_ltiles, n_rb;  // n_rb = H * n_ltiles row blocks per (sample, split)
  int q_lo_off;                              // SPLIT: column offset of the lo parts inside a Q' row
  int ctas_per_stream;                       // ceil(n_rb / G)
  int tiles_total;
  int mask_words;  // 64-token mask words per sample
  int c_ones;  // column of z that is 1.0 (= context width C): Q' carries -m_ref there
  int rt_zero;  // 0, but only known at run time (pins the mid-tile barrier test behind the exponentials, see below)
  // Which 16-column steps of the score products hold non-zero columns (Q'h.zh: columns 0..C, the fold / ones column
  // included; Q'l.zh and Q'h.zl: columns 0..C-1); merged: the second step of both lo-order products runs as ONE UMMA on
  // the re-arranged Q'l / zl tails. (ncu: the tensor-core pipe is 82-87 % busy in this kernel.)
  int mode;    // KD 32, split: 0 = C <= 15 (1 + 1 + 1 UMMAs), 1 = C == 16 (2 + 1 + 1), 2 = merged tail (2 + 1 + 1 + 1), 3 = all
  int merged;
  long N;
  const uint64_t* mask_bits;
  float* part_acc;
  float* part_ml;
};

__device__ __forceinline__ float ex2_mufu(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x on the FMA / ALU pipes: n = rint(x) via the 1.5*2^23 trick, degree-3 minimax of 2^f on [-0.5, 0.5],
// exponent inserted with one integer multiply-add. x <= 16 here; x below -125 (masked: -inf) clamps to ~0.
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;
  const float f = x - (t - 12582912.f);
  float p = 0.05517167f;
  p = fmaf(p, f, 0.24261112f);
  p = fmaf(p, f, 0.69326099f);
  p = fmaf(p, f, 0.99992807f);
  return __int_as_float(__float_as_int(t) * 8388608 + __float_as_int(p));
}
__device__ __forceinline__ uint32_t vmaxu2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}
// element j of a 32-column chunk goes to the FMA-pipe polynomial when POLY(j); ~1/3 of the elements, spread so
// that every packed pair mixes the two pipes
// PMODE: 0 = none (all MUFU), 1 = 1/4, 2 = 1/3, 3 = 3/8, 4 = 1/2
template <int PMODE>
__device__ __forceinline__ constexpr bool poly_slot(int j) {
  return PMODE == 0 ? false
         : PMODE == 1 ? (j % 4) == 1
         : PMODE == 2 ? (j % 3) == 1
         : PMODE == 3 ? ((j % 8) == 1 || (j % 8) == 4 || (j % 8) == 6)
                      : (j % 2) == 1;
}
// packed-pair variant: both exponentials of a column pair on the FMA / ALU pipes in half2 arithmetic (13
// instructions per PAIR incl. the fp16 pack, vs 17 for two fp32 polynomials + pack). The argument already carries
// +P_SHIFT (see above). n = rint(x) via the magic constant trick (fp16 ulp is 1 in [1024, 2048)): the integer lands in
// the low mantissa bits and the exponent insert (a lane-wise integer add) builds 2^x — a normal fp16 for every
// x >= -14. Below that the exponent field runs through 0 (a subnormal <= 2^-14, i.e. <= 2^-24 of the reference weight)
// into the sign bit: the lane then reads as a NEGATIVE fp16 (or -inf / NaN), which the final max with +0 turns into an
// exact zero — so tokens far below the reference contribute nothing, like MUFU's flush to zero. (Round 1 clamped the
// argument at -14 instead, which gave every such token a weight of 2^-24: harmless for diffuse softmaxes, but over
// 602 112 voxels a spurious mass of up to 1.8 % once the softmax is peaked — found by tests/test_gpu_fullsize.py.)
// The argument is clamped to [-30, 16]: that keeps the magic-constant sum inside [1024, 2048) and the exponent field
// of an overflowing result at 30 | 31 (>= 2^15.5, inf or NaN with the sign bit CLEAR), which the unsigned max check
// reads as "> 2^15" and sends to the exact path; the zeroing of negative lanes is a SIGNED INTEGER max for the same
// reason (a floating-point max would also erase the NaN). (Round 1 had no upper clamp: a jump of the row max by more
// than 2^38 on a polynomial column could wrap the 6-bit exponent add all the way round and go unnoticed.)
// Relative error ~6e-4 rms (3x the fp16 rounding of P itself, zero-mean).
__device__ __forceinline__ uint32_t ex2_pair_h2(float x0, float x1) {
  __half2 xh = __floats2half2_rn(x0, x1);
  xh = __hmin2(__hmax2(xh, __float2half2_rn(-30.f)), __float2half2_rn(16.f));
  const __half2 magic = __float2half2_rn(1536.f);
  const __half2 t = __hadd2(xh, magic);
  const __half2 f = __hsub2(xh, __hsub2(t, magic));  // t = 1536 + rint(x) exactly
  __half2 p = __float2half2_rn(0.05517167f);
  p = __hfma2(p, f, __float2half2_rn(0.24261112f));
  p = __hfma2(p, f, __float2half2_rn(0.69326099f));
  p = __hfma2(p, f, __float2half2_rn(0.99992807f));
  const uint32_t e = (*reinterpret_cast<const uint32_t*>(&t) << 10) & 0xFC00FC00u;
  uint32_t r;
  asm("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&p)), "r"(e));
  asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(0u));
  return r;
}
// Leaner form of the same polynomial for the steady-state path (10 issue slots per pair instead of 13):
//   * only the LOWER clamp is applied, at -16, as one unsigned integer min on the fp16 bit patterns (negative halves
//     order by magnitude: min.u16 with bits(-16) clamps them, positive ones pass). rint(x) = -16 already flushes
//     to an exact zero below (exponent field 15 - 16 < 0); for -15.5 < x < -14 the field is 0 and the lane reads as a
//     subnormal — a weight below 2^-24 of the reference, under-estimated by up to 40 %, never over-estimated
//     (tests/test_exp2_lane_model.py models the whole sequence on the CPU, exhaustively over the fp16 inputs);
//   * the magic constant is 1536 + 16, so t = 1552 + rint(x) has the bit pattern 0x6600 + n', n' = rint(x) + 16 in
//     [0, 31] for every admissible x: the exponent insert is then ONE 32-bit multiply-add (t * 1024 + p: no carry can
//     cross the lanes, and what the low lane's constant bits push into the high lane is the constant 0x198) followed by
//     one lane-wise add of the correction (-16 << 10 per lane, -0x198 in the high lane) fused with the signed max
//     against 0 that flushes the lanes whose exponent ran through zero;
//   * there is no upper clamp: an x that rounds to 16 or more (P above 2^15, the raise threshold) gives n' >= 32, i.e.
//     a t word with a bit outside 0x661F (every 16-bit number above 0x661F has one; t is monotonic in x up to +inf /
//     NaN). The caller ORs the t words together (one 3-input LOP3 per two pairs) and takes the exact path — where
//     nothing of this result is used — when such a bit shows, exactly as it does for a MUFU lane above 2^15.
constexpr uint32_t POLY_T_OK = 0x661F661Fu;
__device__ __forceinline__ uint32_t ex2_pair_h2_lean(float x0, float x1, uint32_t& tbits) {
  const uint32_t xr = pack_half2(x0, x1);
  uint32_t xb;
  asm("min.u16x2 %0, %1, %2;" : "=r"(xb) : "r"(xr), "r"(0xCC00CC00u));
  const __half2 xh = *reinterpret_cast<const __half2*>(&xb);
  const __half2 magic = __float2half2_rn(1552.f);
  const __half2 t = __hadd2(xh, magic);
  const __half2 f = __hsub2(xh, __hsub2(t, magic));
  __half2 p = __float2half2_rn(0.05517167f);
  p = __hfma2(p, f, __float2half2_rn(0.24261112f));
  p = __hfma2(p, f, __float2half2_rn(0.69326099f));
  p = __hfma2(p, f, __float2half2_rn(0.99992807f));
  tbits = *reinterpret_cast<const uint32_t*>(&t);
  uint32_t r = tbits * 1024u + *reinterpret_cast<const uint32_t*>(&p);
  asm("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(0xBE68C000u));
  asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(r), "r"(0u));
  return r;
}
// PMODE >= 5: which column PAIRS take the half2 polynomial: 5 = 1/3, 6 = 1/2, 7 = 2/3, 8 = 1/4
template <int PMODE>
__device__ __forceinline__ constexpr bool poly_pair(int j) {
  return PMODE == 5 ? (j % 3) == 1 : PMODE == 6 ? (j % 2) == 1 : PMODE == 7 ? (j % 3) != 1 : PMODE == 8 ? (j % 4) == 1 : false;
}
// byte offset of element (row, col) inside a TMA-swizzled [rows][KD] fp16 tile (64-byte rows -> SWIZZLE_64B,
// 128-byte rows -> SWIZZLE_128B): the 16-byte chunk index is XORed with the low bits of (row-pair | row)
template <int KD>
__device__ __forceinline__ uint32_t swizzled_off(int row, int col) {
  const uint32_t chunk = static_cast<uint32_t>(col * 2) >> 4, within = static_cast<uint32_t>(col * 2) & 15u;
  if (KD == 32) return row * 64 + ((chunk ^ ((static_cast<uint32_t>(row) >> 1) & 3u)) << 4) + within;
  return row * 128 + ((chunk ^ (static_cast<uint32_t>(row) & 7u)) << 4) + within;
}

// ---------------------------------------------------------------- softmax warps, 16 rows each (SW = 8)
// Warp wg (0..7) of row block g owns TMEM lanes [32 (wg & 3) + 16 (wg >> 2), +16): the .16x256b fragment gives thread t
// rows t/4 and t/4 + 8 of those, and of each 32-column chunk the column pairs 8k + 2(t%4) + {0, 1}, k = 0..3 — which
// packed to fp16x2 are exactly the .16x128b fragment of P. Row state (reference max, folded offsets) is therefore
// replicated across the four threads of a quad and agreed on with two shuffles; everything else is the algorithm of the
// SW = 4 variant (lazily raised reference, fold in Q', steady-state path without max pass).
template <int KD, int G, int PMODE, bool SPLIT>
__device__ __forceinline__ void softmax_rows16(const SmallDev& p, uint8_t* sQ, uint64_t* s_full_all, uint64_t* p_ready_all,
                                               uint64_t* u_done_all, uint64_t* acc_done_all, uint32_t tmem, int warp,
                                               int lane, int rb0, int n_active, int b, int split, int t_begin,
                                               int t_end) {
  constexpr int VD = KD;
  constexpr int Q_GROUP = (SPLIT ? 2 : 1) * BM * KD * 2;
  constexpr int GCOLS = 2 * 64 + VD;
  const int g = warp >> 3;
  if (g >= n_active) return;
  const int n = t_end - t_begin;
  const int wg = warp & 7;
  const int rb = rb0 + g, h = rb / p.n_ltiles, lt = rb % p.n_ltiles;
  const uint32_t lane_off = (wg & 3) * 32 + (wg >> 2) * 16;
  const int qd = lane >> 2, qc = lane & 3;
  const int trow0 = lane_off + qd;                  // rows inside the 128-row block: trow0 and trow0 + 8
  const uint32_t tG = tmem + g * GCOLS;
  const uint32_t tL = tmem_addr(tG, lane_off, 0);
  const uint32_t tU = tL + 128;
  uint64_t* s_full = s_full_all + g * 2;
  uint64_t* p_ready = p_ready_all + g * 2;
  uint64_t* u_done = u_done_all + g;
  uint64_t* acc_done = acc_done_all + g;
  uint8_t* qrow_fold = sQ + g * Q_GROUP;
  float m_ref[2] = {-INFINITY, -INFINITY};
  float m_in0[2] = {0.f, 0.f}, m_in1[2] = {0.f, 0.f};

  const int n_full = (t_end == p.tiles_total && (p.N % BT) != 0) ? n - 1 : n;
  const bool has_mask = p.mask_bits != nullptr;
  const int n_steady = has_mask ? 0 : n_full;
  bool stale = true;
  bool ready = false;

  auto tile = [&](const int i, auto buf_c, const uint32_t ph) {
    constexpr int buf = decltype(buf_c)::value;
    const uint32_t tS = tL + buf * 64;
    if (!ready) mbar_wait(&s_full[buf], ph);
    fence_after_sync();
    ready = false;

    bool exact = stale || i >= n_steady;
    stale = false;
    uint32_t pk[16];  // P(i): register 2K + a = row a, packed column 4K + qc  (K = 0..7)
    if (!exact) {
      uint32_t pmax = 0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t s[16];
        tmem_ld_16x256b_x4(tS + c * 32, s);
        tmem_wait_ld();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const float x0 = __uint_as_float(s[4 * k + 2 * a]), x1 = __uint_as_float(s[4 * k + 2 * a + 1]);
            const int o = c * 8 + 2 * k + a;
            // PMODE 6: every other pair of every row on the FMA pipe; other shares by running pair index
            if (PMODE == 6 ? ((k + a) & 1) : (PMODE >= 5 && poly_pair<PMODE>(c * 8 + 2 * k + a + (a ? 1 : 0)))) {
              pk[o] = ex2_pair_h2(x0, x1);
            } else {
              pk[o] = pack_half2(ex2_mufu(x0), ex2_mufu(x1));
            }
            pmax = vmaxu2(pmax, pk[o]);
          }
        }
        if (c == 0) ready = mbar_try_wait(&s_full[buf ^ 1], buf ? (ph ^ 1) : ph);
      }
      const bool big = ((pmax & 0xFFFFu) > P_RAISE_BITS) || ((pmax >> 16) > P_RAISE_BITS);
      exact = __any_sync(0xffffffffu, big);
    }
    if (exact) {
      const int tile_idx = t_begin + i;
      uint64_t bits = ~0ull;
      if (has_mask) bits = p.mask_bits[static_cast<long>(b) * p.tiles_total + tile_idx];
      const long rem = p.N - static_cast<long>(tile_idx) * BT;
      if (rem < BT) bits &= (1ull << rem) - 1ull;
      const uint64_t mybits = bits >> (2 * qc);  // bit (8k + e) of the shifted word = my column 8k + 2 qc + e
      float m_in[2] = {buf ? m_in1[0] : m_in0[0], buf ? m_in1[1] : m_in0[1]};
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t s[16];
        tmem_ld_16x256b_x4(tS + c * 32, s);
        tmem_wait_ld();
        const uint32_t mb = static_cast<uint32_t>(mybits >> (32 * c));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const float v0 = ((mb >> (8 * k)) & 1u) ? __uint_as_float(s[4 * k + 2 * a]) : -INFINITY;
            const float v1 = ((mb >> (8 * k + 1)) & 1u) ? __uint_as_float(s[4 * k + 2 * a + 1]) : -INFINITY;
            mx[a] = fmax3(mx[a], v0, v1);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], 1));
        mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], 2));
        mx[a] += m_in[a];
      }
      const bool raise = __any_sync(0xffffffffu, mx[0] > m_ref[0] + RESCALE_THRESHOLD || mx[1] > m_ref[1] + RESCALE_THRESHOLD);
      if (raise) {
        __half fold_h[2];
        float m_new[2], sc[2];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          fold_h[a] = __float2half_rn(P_SHIFT - fmaxf(m_ref[a], mx[a]));
          m_new[a] = (mx[a] == -INFINITY && m_ref[a] == -INFINITY) ? -INFINITY : P_SHIFT - __half2float(fold_h[a]);
          sc[a] = (m_new[a] == -INFINITY || m_ref[a] == m_new[a]) ? 1.f : ex2_mufu(m_ref[a] - m_new[a]);
        }
        if (i > 0) {
          mbar_wait(u_done, (i - 1) & 1);  // PV(i-1) landed; PV(i) cannot start before our p_ready arrive
          fence_after_sync();
          if (__any_sync(0xffffffffu, sc[0] != 1.f || sc[1] != 1.f)) {
#pragma unroll
            for (int c = 0; c < VD; c += 32) {
              uint32_t u[16];
              tmem_ld_16x256b_x4(tU + c, u);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j) u[j] = __float_as_uint(__uint_as_float(u[j]) * sc[(j >> 1) & 1]);
              tmem_st_16x256b_x4(tU + c, u);
            }
          }
        }
        if (i + 1 < n) {  // S(i+1) may still be reading Q': wait for it before changing the folded offset
          mbar_wait(&s_full[buf ^ 1], buf ? (ph ^ 1) : ph);
          ready = true;
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
          m_ref[a] = m_new[a];
          if (qc == 0)
            *reinterpret_cast<__half*>(qrow_fold + swizzled_off<KD>(trow0 + 8 * a, p.c_ones)) =
                (m_ref[a] == -INFINITY) ? __float2half_rn(0.f) : fold_h[a];
        }
        fence_proxy_async_smem();
        stale = true;
      }
      float m_cur[2], delta[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        m_cur[a] = (m_ref[a] == -INFINITY) ? 0.f : m_ref[a] - P_SHIFT;
        delta[a] = m_in[a] - m_cur[a];
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t s[16];
        tmem_ld_16x256b_x4(tS + c * 32, s);
        tmem_wait_ld();
        const uint32_t mb = static_cast<uint32_t>(mybits >> (32 * c));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const float v0 = ((mb >> (8 * k)) & 1u) ? __uint_as_float(s[4 * k + 2 * a]) : -INFINITY;
            const float v1 = ((mb >> (8 * k + 1)) & 1u) ? __uint_as_float(s[4 * k + 2 * a + 1]) : -INFINITY;
            pk[c * 8 + 2 * k + a] = pack_half2(ex2_mufu(v0 + delta[a]), ex2_mufu(v1 + delta[a]));
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (buf) m_in1[a] = m_cur[a]; else m_in0[a] = m_cur[a];
      }
    }
    tmem_st_16x128b_x8(tS, pk);  // P(i) over S columns 0..31 of my 16 lanes (all 64 of them have been read)
    tmem_wait_st();
    fence_before_sync();
    __syncwarp();
    if (lane == 0) mbar_arrive(&p_ready[buf]);
    __syncwarp();
  };
  {
    int i = 0;
    uint32_t ph = 0;
    for (; i + 1 < n; i += 2, ph ^= 1) {
      tile(i, std::integral_constant<int, 0>{}, ph);
      tile(i + 1, std::integral_constant<int, 1>{}, ph);
    }
    if (i < n) tile(i, std::integral_constant<int, 0>{}, ph);
  }
  // ---- epilogue: un-normalised accumulator rows + reference max
  mbar_wait(acc_done, 0);
  fence_after_sync();
#pragma unroll
  for (int c = 0; c < VD; c += 32) {
    uint32_t u[16];
    tmem_ld_16x256b_x4(tU + c, u);
    tmem_wait_ld();
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int row = lt * BM + trow0 + 8 * a;
      if (row < p.L) {
        const long part_row = ((static_cast<long>(b) * p.nsplit + split) * p.H + h) * p.L + row;
        float* dst = p.part_acc + part_row * VD + c + 2 * qc;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<float2*>(dst + 8 * k) =
              make_float2(__uint_as_float(u[4 * k + 2 * a]), __uint_as_float(u[4 * k + 2 * a + 1]));
      }
    }
    __syncwarp();
  }
  if (qc == 0) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int row = lt * BM + trow0 + 8 * a;
      if (row < p.L) {
        const long part_row = ((static_cast<long>(b) * p.nsplit + split) * p.H + h) * p.L + row;
        *reinterpret_cast<float2*>(p.part_ml + part_row * 2) = make_float2(m_ref[a], 0.f);
      }
    }
  }
}

// SW = softmax warps per row block: 4 (thread = latent row, .32x32b TMEM fragments) or 8 (each warp owns 16 rows of
// the block and sees all 64 columns of them through the .16x256b / .16x128b fragments: twice the warps per scheduler to
// hide the per-tile TMEM / mbarrier latencies, half the registers per thread).
template <int KD, int G, int PMODE, bool SPLIT, int SW>
__global__ void __launch_bounds__(((SW + 1) * G + 1) * 32, 1)
attn_small_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmZ, SmallDev p) {
  constexpr int VD = KD;
  constexpr int Q_TILE = BM * KD * 2;            // one [128][KD] operand tile (hi or lo)
  constexpr int Q_GROUP = (SPLIT ? 2 : 1) * Q_TILE;   // per row block: [Q'h | Q'l]
  constexpr int Z_BYTES = BT * KD * 2;           // one [64][KD] tile (hi or lo)
  constexpr int Z_STAGE = (SPLIT ? 2 : 1) * Z_BYTES;  // per ring stage: [zh | zl]
  constexpr uint32_t LAYOUT = (KD == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * KD * 2;
  constexpr uint32_t V_KADV = 16 * VD * 2;
  constexpr int GCOLS = 2 * 64 + VD;  // S/P buffer 0 | S/P buffer 1 | U   (P(i) overwrites the first 32 columns of S(i))
  static_assert(G * GCOLS <= 512, "TMEM budget");
  constexpr uint32_t idesc_s = idesc_f16(BM, BT, false, false);  // S[128x64]  = Q'[128xKD] . z[64xKD]^T
  constexpr uint32_t idesc_u = idesc_f16(BM, VD, false, true);   // U[128xVD] += P[128x64] . z[64xVD] (MN-major B)

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sZ = smem + G * Q_GROUP;
  __shared__ uint64_t q_full, z_full[NST], z_empty[NST];
  __shared__ uint64_t s_full[MAXG][2], p_ready[MAXG][2], u_done[MAXG], acc_done[MAXG], q_fixed[MAXG];
  __shared__ uint32_t tmem_base_s;
  // UMMA shared-memory descriptors of the context-row ring, one row per stage: [0..3] the V operand of the four
  // 16-token steps of P.z, [4..7] the K operand steps zh k0, zh k1, zl k0, zl k1 (K-major). Read back with VOLATILE
  // loads before the issuer's wait for P(i), so that they sit in registers when the wait returns (ptxas otherwise
  // re-materialises the ~45 instructions of descriptor arithmetic behind the wait, i.e. on the P(i) -> S(i+2) chain).
  __shared__ __align__(16) uint64_t z_desc[NST][8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int idx = blockIdx.x;
  const int cta_rb = idx % p.ctas_per_stream;
  idx /= p.ctas_per_stream;
  const int b = idx % p.batch;
  const int split = idx / p.batch;
  const int rb0 = cta_rb * G;
  const int n_active = (p.n_rb - rb0) < G ? (p.n_rb - rb0) : G;
  const int t_begin = static_cast<int>(static_cast<long>(p.tiles_total) * split / p.nsplit);
  const int t_end = static_cast<int>(static_cast<long>(p.tiles_total) * (split + 1) / p.nsplit);
  const int n = t_end - t_begin;

  if (n <= 0) {  // more splits than tiles: publish empty partials
    HN_PDL_WAIT();
    for (int g = 0; g < n_active; ++g) {
      const int rb = rb0 + g, h = rb / p.n_ltiles, lt = rb % p.n_ltiles;
      const long row0 = ((static_cast<long>(b) * p.nsplit + split) * p.H + h) * p.L + lt * BM;
      for (int r = threadIdx.x; r < BM; r += blockDim.x) {
        if (lt * BM + r < p.L) {
          float* acc = p.part_acc + (row0 + r) * VD;
          for (int c = 0; c < VD; ++c) acc[c] = 0.f;
          p.part_ml[(row0 + r) * 2] = -INFINITY;
          p.part_ml[(row0 + r) * 2 + 1] = 0.f;
        }
      }
    }
    return;
  }

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < NST; ++s) {
      mbar_init(&z_full[s], 1);
      mbar_init(&z_empty[s], n_active);
    }
    for (int g = 0; g < MAXG; ++g) {
      mbar_init(&s_full[g][0], 1);
      mbar_init(&s_full[g][1], 1);
      mbar_init(&p_ready[g][0], SW);
      mbar_init(&p_ready[g][1], SW);
      mbar_init(&u_done[g], 1);
      mbar_init(&acc_done[g], 1);
      mbar_init(&q_fixed[g], SW);
    }
    fence_mbar_init();
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + NST) {
    const int st = threadIdx.x - 32;
    const uint32_t z0 = smem_u32(sZ + st * Z_STAGE);
#pragma unroll
    for (int k = 0; k < 4; ++k) z_desc[st][k] = smem_desc(z0 + k * V_KADV, 16, SBO, LAYOUT);
    z_desc[st][4] = smem_desc(z0, 16, SBO, LAYOUT);
    z_desc[st][5] = smem_desc(z0 + 32, 16, SBO, LAYOUT);
    z_desc[st][6] = smem_desc(z0 + Z_BYTES, 16, SBO, LAYOUT);
    z_desc[st][7] = smem_desc(z0 + Z_BYTES + 32, 16, SBO, LAYOUT);
  }
  constexpr int PRODUCER_WARP = (SW + 1) * G;
  if (warp == PRODUCER_WARP) tmem_alloc<512>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();

  if (warp == PRODUCER_WARP) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmZ);
      mbar_arrive_expect_tx(&q_full, n_active * Q_GROUP);
      for (int g = 0; g < n_active; ++g) {
        const int rb = rb0 + g, h = rb / p.n_ltiles, lt = rb % p.n_ltiles;
        tma_load_3d(sQ + g * Q_GROUP, &tmQ, &q_full, h * KD, lt * BM, b);
        if (SPLIT) tma_load_3d(sQ + g * Q_GROUP + Q_TILE, &tmQ, &q_full, p.q_lo_off + h * KD, lt * BM, b);
      }
      for (int i = 0; i < n; ++i) {
        const int s = i % NST;
        mbar_wait_sleepy(&z_empty[s], ((i / NST) & 1) ^ 1, 20000);
        mbar_arrive_expect_tx(&z_full[s], Z_STAGE);
        tma_load_3d(sZ + s * Z_STAGE, &tmZ, &z_full[s], 0, (t_begin + i) * BT, b);
        if (SPLIT) tma_load_3d(sZ + s * Z_STAGE + Z_BYTES, &tmZ, &z_full[s], KD, (t_begin + i) * BT, b);
      }
    }
  } else if (warp >= SW * G) {
    // ------------------------------------------------------------ UMMA issuers: one warp (one elected thread) per
    // group, each on its own scheduler. Once all four softmax warps have published P(i) the thread issues PV(i)
    // and, right behind it, S(i+2) into the buffer P(i) occupied — so S(i+1) is always complete before the
    // softmax needs it and no tensor-pipe or issue latency sits on the softmax warps' critical path.
    const int g = warp - SW * G;
    if (g < n_active && elect_one()) {
      const uint32_t tG = tmem + g * GCOLS;
      const uint32_t q0 = smem_u32(sQ + g * Q_GROUP);
      // S(i) = Q'.z_i^T into S buffer i & 1; z0 = shared address of the tile's ring stage (already landed). Only the
      // 16-column steps that hold non-zero columns are issued (columns above C are zero in Q' and z): KH steps of
      // Q'h.zh, KL of each lo-order product, compile-time per mode so that the issue sequence stays branch-free.
      // dz: the K-operand descriptors of the tile's ring stage (KD 32: zh k0, zh k1, zl k0, zl k1 from the table)
      auto issue_s_c = [&](int i, uint32_t z0, const uint64_t (&dz)[4], auto kh_c, auto kl_c, auto mg_c) {
        constexpr int KH = decltype(kh_c)::value, KL = decltype(kl_c)::value;
        constexpr bool MG = decltype(mg_c)::value;
        const uint32_t tS = tG + (i & 1) * 64;
        auto zh = [&](int k) { return KD == 32 ? dz[k] : smem_desc(z0 + k * 32, 16, SBO, LAYOUT); };
        auto zl = [&](int k) { return KD == 32 ? dz[2 + k] : smem_desc(z0 + Z_BYTES + k * 32, 16, SBO, LAYOUT); };
#pragma unroll
        for (int k = 0; k < KH; ++k) umma_ss(tS, smem_desc(q0 + k * 32, 16, SBO, LAYOUT), zh(k), idesc_s, k != 0);
        if (SPLIT) {
#pragma unroll
          for (int k = 0; k < KL; ++k)  // Q'_lo . z_hi
            umma_ss(tS, smem_desc(q0 + Q_TILE + k * 32, 16, SBO, LAYOUT), zh(k), idesc_s, true);
#pragma unroll
          for (int k = 0; k < KL; ++k)  // Q'_hi . z_lo
            umma_ss(tS, smem_desc(q0 + k * 32, 16, SBO, LAYOUT), zl(k), idesc_s, true);
          if (MG)  // columns 16..C-1 of both lo-order products: [Q'h tail | 0 | Q'l tail] . [zl tail | 0 | zh tail]
            umma_ss(tS, smem_desc(q0 + Q_TILE + 32, 16, SBO, LAYOUT), zl(1), idesc_s, true);
        }
        umma_commit(&s_full[g][i & 1]);
      };
      using std::integral_constant;
      const int mode = p.mode;
      auto issue_s = [&](int i, uint32_t z0, const uint64_t (&dz)[4]) {
        constexpr int KF = KD / 16;
        if (KD == 32 && SPLIT) {
          // (an if-chain: a switch became an indirect branch through a constant-bank table on the critical path)
          if (mode == 2) { issue_s_c(i, z0, dz, integral_constant<int, 2>{}, integral_constant<int, 1>{}, std::true_type{}); return; }
          if (mode == 0) { issue_s_c(i, z0, dz, integral_constant<int, 1>{}, integral_constant<int, 1>{}, std::false_type{}); return; }
          if (mode == 1) { issue_s_c(i, z0, dz, integral_constant<int, 2>{}, integral_constant<int, 1>{}, std::false_type{}); return; }
        }
        issue_s_c(i, z0, dz, integral_constant<int, KF>{}, integral_constant<int, KF>{}, std::false_type{});
      };
      // volatile 16-byte reads of one table row half: [first, first + 4)
      auto load_desc4 = [&](int stage, int first, uint64_t (&d)[4]) {
        const uint32_t a = smem_u32(&z_desc[stage][first]);
        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(d[0]), "=l"(d[1]) : "r"(a));
        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(d[2]), "=l"(d[3]) : "r"(a + 16));
      };
      auto z_stage = [&](int i) {  // wait until tile i's context rows have landed; -> their shared address
        const int s = i % NST;
        mbar_wait(&z_full[s], (i / NST) & 1);
        return smem_u32(sZ + s * Z_STAGE);
      };
      mbar_wait(&q_full, 0);
      if (SPLIT && p.merged) mbar_wait(&q_fixed[g], 0);  // the row owners have re-arranged the tail of the Q'l tile
      for (int j = 0; j < 2 && j < n; ++j) {
        const uint32_t za = z_stage(j);
        uint64_t dz[4];
        load_desc4(j % NST, 4, dz);
        fence_after_sync();
        issue_s(j, za, dz);
      }
      for (int i = 0; i < n; ++i) {
        // The softmax warps' wait for S(i+2) starts the moment they publish P(i): everything on this thread between
        // that arrival and the commit of S(i+2) is on their critical path (ncu: 56 % of the tiles found S not yet
        // complete). So whatever does not depend on P(i) happens BEFORE the wait: the tile's V descriptors, the
        // arrival of the context rows of tile i+2 (landed long ago: the ring is 8 deep) and their descriptors.
        const int s = i % NST;
        uint64_t dv[4], dz[4] = {0, 0, 0, 0};
        if (KD == 32) {
          load_desc4(s, 0, dv);
        } else {
          const uint32_t z0 = smem_u32(sZ + s * Z_STAGE);
#pragma unroll
          for (int k = 0; k < 4; ++k) dv[k] = smem_desc(z0 + k * V_KADV, 16, SBO, LAYOUT);
        }
        const bool more = i + 2 < n;
        uint32_t z2 = 0;
        if (more) {
          z2 = z_stage(i + 2);
          if (KD == 32) load_desc4((i + 2) % NST, 4, dz);
        }
        // all softmax warps of the group: S(i) consumed, P(i) in TMEM, Q' fold up to date. Two alternating barriers: a warp may run
        // one tile ahead of its group but never two, so it cannot arrive twice in one phase of either.
        mbar_wait_sleepy(&p_ready[g][i & 1], (i >> 1) & 1, 20000);
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < BT / 16; ++k)
          umma_ts(tG + 128, tG + (i & 1) * 64 + k * 8, dv[k], idesc_u, (i | k) != 0);
        umma_commit(&z_empty[s]);
        umma_commit(&u_done[g]);
        if (more) issue_s(i + 2, z2, dz);
        if (i + 1 == n) umma_commit(&acc_done[g]);
      }
    }
  } else if (SW == 8) {
    softmax_rows16<KD, G, PMODE, SPLIT>(p, sQ, &s_full[0][0], &p_ready[0][0], u_done, acc_done, tmem, warp, lane, rb0,
                                        n_active, b, split, t_begin, t_end);
  } else {
    // ------------------------------------------------------------ softmax groups: thread = latent row
    const int g = warp >> 2;
    if (g < n_active) {
      const int rb = rb0 + g, h = rb / p.n_ltiles, lt = rb % p.n_ltiles;
      const uint32_t lane_base = (warp & 3) * 32;
      const int trow = lane_base + lane;  // row inside the 128-row block
      const uint32_t tG = tmem + g * GCOLS;  // group columns (lane field 0)
      const uint32_t tL = tmem_addr(tG, lane_base, 0);
      const uint32_t tU = tL + 128;
      uint8_t* qrow_fold = sQ + g * Q_GROUP;  // Q'h[trow][C_ones] holds -m_ref (fp16): the UMMA subtracts the max for us
      float m_ref = -INFINITY;  // reference max (log2 units), always exactly representable in fp16
      float m_in0 = 0.f, m_in1 = 0.f;  // offset baked into S buffer 0 / 1 by the fold (0: none; else m_ref - P_SHIFT)

      if (SPLIT && KD == 32 && p.merged) {
        // Merged tail (17 <= C <= 23): columns 16..C-1 of the two lo-order products Q'l.zh and Q'h.zl go through ONE
        // UMMA on the second 16-column step of the Q'l and zl tiles. The context rows arrive with zl[16..32) =
        // [zl_16 .. zl_C-1 | 0 | zh_16 .. zh_C-1 | 0 ..] (rowops.cu); every thread re-arranges ITS row of the Q'l tile
        // to match, once: [Q'h_16 .. Q'h_C-1 | 0 | Q'l_16 .. Q'l_C-1 | 0 ..].
        mbar_wait(&q_full, 0);
        uint8_t* qh_t = sQ + g * Q_GROUP;
        uint8_t* ql_t = qh_t + Q_TILE;
        const int e = p.c_ones - 16;
        __half hi_t[8], lo_t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hi_t[j] = __float2half_rn(0.f);
          lo_t[j] = __float2half_rn(0.f);
          if (j < e) {
            hi_t[j] = *reinterpret_cast<const __half*>(qh_t + swizzled_off<KD>(trow, 16 + j));
            lo_t[j] = *reinterpret_cast<const __half*>(ql_t + swizzled_off<KD>(trow, 16 + j));
          }
        }
#pragma unroll
        for (int c = 16; c < 32; ++c) {
          __half v = __float2half_rn(0.f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < e && c == 16 + j) v = hi_t[j];
            if (j < e && c == p.c_ones + 1 + j) v = lo_t[j];
          }
          *reinterpret_cast<__half*>(ql_t + swizzled_off<KD>(trow, c)) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_fixed[g]);
        __syncwarp();
      }
      // tiles [0, n_steady) can take the steady-state path (complete, unmasked 64-token tiles; only the globally last
      // tile of the token axis can be ragged)
      const int n_full = (t_end == p.tiles_total && (p.N % BT) != 0) ? n - 1 : n;
      const bool has_mask = p.mask_bits != nullptr;
      const int n_steady = has_mask ? 0 : n_full;
      // warp-uniform: true while the S buffer about to be read was computed before the last raise (and for tile 0,
      // which has no reference yet)
      bool stale = true;
      bool ready = false;  // s_full of the tile about to start was already seen complete (tested inside the previous tile)

      // One tile; BUF is a compile-time constant (the loop below is unrolled by two) so that every barrier / TMEM
      // address is an immediate offset and the per-tile bookkeeping stays off the serial path.
      auto tile = [&](const int i, auto buf_c, const uint32_t ph) {
        constexpr int buf = decltype(buf_c)::value;
        const uint32_t tS = tL + buf * 64;
        if (!ready) mbar_wait(&s_full[g][buf], ph);
        fence_after_sync();
        ready = false;

        bool exact = stale || i >= n_steady;
        stale = false;
        // P(i) as packed fp16 pairs goes over S columns 0..31 once the whole row is known good (this thread has read
        // all 64 of them)
        auto published = [&]() {
          tmem_wait_st();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_ready[g][buf]);
          __syncwarp();
        };
        if (!exact) {
          // ---------------- steady state: P = 2^S in four 16-column chunks; no max pass, no subtraction, no mask.
          // 16-register loads and 8-register stores: with 32-register tuples ptxas spent ~30 MOVs per tile on moving
          // results into the store tuple (tools/sass_tile_mix.py counts the body)
          uint32_t pk[4][8];
          uint32_t pmax = 0, tor = 0;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t s[16];
            tmem_ld16(tS + c * 16, s);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float x0 = __uint_as_float(s[2 * j]), x1 = __uint_as_float(s[2 * j + 1]);
#ifdef HN_DEBUG
              if (PMODE == 9) {  // timing experiment only (debug builds): no exponentials at all (skeleton cost)
                pk[c][j] = pack_half2(x0, x1) & 0x3FFF3FFFu;
              } else
#endif
              if (PMODE >= 5 && poly_pair<PMODE>(j)) {
                uint32_t tb;
                pk[c][j] = ex2_pair_h2_lean(x0, x1, tb);
                tor |= tb;  // an overflowing polynomial lane shows in its t word, not in its result
              } else {
                const float e0 = (PMODE < 5 && poly_slot<PMODE>(2 * j)) ? ex2_poly(x0) : ex2_mufu(x0);
                const float e1 = (PMODE < 5 && poly_slot<PMODE>(2 * j + 1)) ? ex2_poly(x1) : ex2_mufu(x1);
                pk[c][j] = pack_half2(e0, e1);
              }
            }
            // (kept together so that ptxas pairs them into 3-input VIMNMX3)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (!(PMODE >= 5 && poly_pair<PMODE>(j))) pmax = vmaxu2(pmax, pk[c][j]);
            // S(i+1) is normally complete well before this tile ends: test its barrier late in the tile, where the
            // ~100 clk of the test hide behind the last chunk's exponentials, instead of opening the next tile with it.
            // NON-blocking test, and its address is made to depend on this chunk's results (x rt_zero): ptxas hoists a
            // free-standing barrier instruction to the top of the tile, where S(i+1) cannot be complete yet, and a
            // blocking try_wait there held back the publication of P(i) until S(i+1) had landed (ncu source view:
            // 12 % of the softmax warps' time on the instruction consuming its predicate) — which in turn delayed
            // PV(i) and S(i+2): the whole pipeline ran in lock-step with the tensor pipe.
            if (c == 3)  // (anchored on an early result of chunk 2: lands ~80 % through the tile body)
              ready = mbar_test_wait_addr(smem_u32(&s_full[g][buf ^ 1]) + pk[2][1] * static_cast<uint32_t>(p.rt_zero),
                                          buf ? (ph ^ 1) : ph);
          }
          const bool big = ((pmax & 0xFFFFu) > P_RAISE_BITS) || ((pmax >> 16) > P_RAISE_BITS) || (tor & ~POLY_T_OK) != 0u;
          exact = __any_sync(0xffffffffu, big);  // some P above 2^15 (or inf / garbage): redo with a raised reference
          if (!exact) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tmem_st8(tS + c * 8, pk[c]);
            published();
          }
        }
        if (exact) {
          uint32_t pk[32];
          // ---------------- exact path (first tiles, masked / ragged tile, stale fold, or the reference max has
          // to be raised): max pass, rescale, exp pass. TMEM holds s - m_in (nothing has been overwritten yet).
          const int tile_idx = t_begin + i;
          uint64_t bits = ~0ull;
          if (has_mask) bits = p.mask_bits[static_cast<long>(b) * p.tiles_total + tile_idx];
          const long rem = p.N - static_cast<long>(tile_idx) * BT;
          if (rem < BT) bits &= (1ull << rem) - 1ull;
          const float m_in = buf ? m_in1 : m_in0;  // what the fold subtracted from the raw scores of this buffer
          float mx = -INFINITY;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t s[32];
            tmem_ld32(tS + c * 32, s);
            tmem_wait_ld();
            const uint32_t mb = static_cast<uint32_t>(bits >> (32 * c));
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float a = ((mb >> j) & 1u) ? __uint_as_float(s[j]) : -INFINITY;
              const float bb = ((mb >> (j + 1)) & 1u) ? __uint_as_float(s[j + 1]) : -INFINITY;
              mx = fmax3(mx, a, bb);
            }
          }
          mx += m_in;  // true row max of this tile (-inf stays -inf)
          // raise lazily: keep the reference while the tile stays within 2^5 of it
          const bool raise = __any_sync(0xffffffffu, mx > m_ref + RESCALE_THRESHOLD);
          if (raise) {
            // the fold (-m_ref + P_SHIFT) lives in Q' as fp16: round IT, and define the reference from the rounded
            // value, so that the folded offset is exactly the reference the steady state assumes
            const __half fold_h = __float2half_rn(P_SHIFT - fmaxf(m_ref, mx));
            const float m_new = (mx == -INFINITY && m_ref == -INFINITY) ? -INFINITY : P_SHIFT - __half2float(fold_h);
            if (i > 0) {
              mbar_wait(&u_done[g], (i - 1) & 1);  // PV(i-1) landed; PV(i) cannot start before our p_ready arrive
              fence_after_sync();
              const float sc = (m_new == -INFINITY || m_ref == m_new) ? 1.f : ex2_mufu(m_ref - m_new);
              if (__any_sync(0xffffffffu, sc != 1.f)) {
#pragma unroll
                for (int c = 0; c < VD; c += 32) {
                  uint32_t u[32];
                  tmem_ld32(tU + c, u);
                  tmem_wait_ld();
#pragma unroll
                  for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(__uint_as_float(u[j]) * sc);
                  tmem_st32(tU + c, u);
                }
              }
            }
            // S(i+1) may still be reading Q': wait for it before changing the folded offset
            if (i + 1 < n) {
              mbar_wait(&s_full[g][buf ^ 1], buf ? (ph ^ 1) : ph);
              ready = true;
            }
            m_ref = m_new;
            *reinterpret_cast<__half*>(qrow_fold + swizzled_off<KD>(trow, p.c_ones)) =
                (m_ref == -INFINITY) ? __float2half_rn(0.f) : fold_h;
            fence_proxy_async_smem();
            stale = true;  // S(i+1) was computed with the previous offset
          }
          const float m_cur = (m_ref == -INFINITY) ? 0.f : m_ref - P_SHIFT;  // offset the current fold subtracts
          const float delta = m_in - m_cur;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t s[32];
            tmem_ld32(tS + c * 32, s);
            tmem_wait_ld();
            const uint32_t mb = static_cast<uint32_t>(bits >> (32 * c));
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = ((mb >> (2 * j)) & 1u) ? __uint_as_float(s[2 * j]) : -INFINITY;
              const float bb = ((mb >> (2 * j + 1)) & 1u) ? __uint_as_float(s[2 * j + 1]) : -INFINITY;
              pk[c * 16 + j] = pack_half2(ex2_mufu(a + delta), ex2_mufu(bb + delta));
            }
          }
          // the tile S(i+2) that will land in this buffer carries the current reference
          if (buf) m_in1 = m_cur; else m_in0 = m_cur;
          tmem_st32(tS, pk);
          published();
        }
      };
      {
        int i = 0;
        uint32_t ph = 0;
        for (; i + 1 < n; i += 2, ph ^= 1) {
          tile(i, std::integral_constant<int, 0>{}, ph);
          tile(i + 1, std::integral_constant<int, 1>{}, ph);
        }
        if (i < n) tile(i, std::integral_constant<int, 0>{}, ph);
      }
      // ---- epilogue: un-normalised accumulator rows + reference max
      const int row = lt * BM + trow;
      const long part_row = ((static_cast<long>(b) * p.nsplit + split) * p.H + h) * p.L + row;
      mbar_wait(&acc_done[g], 0);
      fence_after_sync();
#pragma unroll
      for (int c = 0; c < VD; c += 32) {
        uint32_t u[32];
        tmem_ld32(tU + c, u);
        tmem_wait_ld();
        if (row < p.L) {
          float4* dst = reinterpret_cast<float4*>(p.part_acc + part_row * VD + c);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]),
                                 __uint_as_float(u[4 * j + 2]), __uint_as_float(u[4 * j + 3]));
        }
        __syncwarp();
      }
      if (row < p.L) {
        float2* ml = reinterpret_cast<float2*>(p.part_ml + part_row * 2);
        *ml = make_float2(m_ref, 0.f);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == PRODUCER_WARP) tmem_dealloc<512>(tmem);
}


// =====================================================================================================================
// Streaming BACKWARD of the small-context cross-attention on the same machinery (training step).
// No gradient flows into the context rows z, so the backward of the whole streaming pass is, per (sample, head, row l),
//     dr_l = sum_t dt_lt z_t ,   dt_lt = p_lt (du_l . z_t - delta_l) ,   p_lt = 2^(r_l . z_t - M_l) / den_l
// (r = log2-unit score vector of the row, du = gradient w.r.t. u_l = sum_t p_lt z_t, delta_l = du_l . u_l; M, den saved
// by the forward). One more pass over the token axis with the forward's tile pipeline:
//     S'(i) = R . z_i^T   (SS, three fp16 terms; R carries P_SHIFT - M in the column where z holds its 1.0)
//     G (i) = DU . z_i^T  (SS, DU split hi + lo: two terms)
//     softmax warps: dt = 2^S' * a_l * (G - delta_l), a_l = 2^-P_SHIFT / den_l  -> fp16 over the first half of S'
//     dR += dt . z_i      (TS, dt from TMEM) — accumulator [128 x KD] fp32 per row block
// DU rows arrive scaled by a per-row power of two (so that fp16 dt keeps its precision whatever the size of the
// incoming gradient); the caller undoes the scale. There is no running max here (M is known), hence no exact path:
// every tile is the steady state. S' and G are single-buffered (TMEM: 64 + 64 + KD columns per row block, three row
// blocks per CTA as in the forward): a row block's tensor and softmax phases alternate, the other two fill the gaps.
// Out-of-range tokens of a ragged last tile have z = 0 (TMA zero fill) and drop out of dt . z by themselves; masked
// tokens are zeroed in dt.
struct SmallBwdDev {
  int L, H, batch, nsplit, n_ltiles, n_rb, ctas_per_stream, tiles_total;
  int lo_off;           // column offset of the lo parts inside R / DU rows
  int mode;             // KD 32: 0 = C <= 15 (every product fits one 16-column step), 1 = C == 16 (only R_hi.z_hi, which
                        // carries the fold column, needs the second step), 2 = merged tail of S' (17 <= C <= 23), else all steps
  long N;
  const uint64_t* mask_bits;
  const float* row_a;   // [(b*L + l)*H + h]  2^-P_SHIFT / den
  const float* row_d;   // [(b*L + l)*H + h]  delta (scaled like DU)
  float* part;          // [b][nsplit][H][L][KD] fp32
};

template <int KD, int G>
__global__ void __launch_bounds__((5 * G + 1) * 32, 1)
attn_small_bwd_kernel(const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmDU,
                      const __grid_constant__ CUtensorMap tmZ, SmallBwdDev p) {
  constexpr int NSB = KD == 64 ? 4 : NST;  // z ring depth (the four operand tiles per row block take the room)
  constexpr int Q_TILE = BM * KD * 2;     // one [128][KD] fp16 tile
  constexpr int Q_GROUP = 4 * Q_TILE;     // per row block: [R_hi | R_lo | DU_hi | DU_lo]
  constexpr int Z_BYTES = BT * KD * 2;
  constexpr int Z_STAGE = 2 * Z_BYTES;    // [z_hi | z_lo]
  constexpr uint32_t LAYOUT = (KD == 64) ? SWZ_128B : SWZ_64B;
  constexpr uint32_t SBO = 8 * KD * 2;
  constexpr uint32_t V_KADV = 16 * KD * 2;
  constexpr int GCOLS = 2 * 64 + KD;      // S' | G | dR
  static_assert(G * GCOLS <= 512, "TMEM budget");
  constexpr uint32_t idesc_s = idesc_f16(BM, BT, false, false);
  constexpr uint32_t idesc_u = idesc_f16(BM, KD, false, true);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sZ = smem + G * Q_GROUP;
  __shared__ uint64_t q_full, z_full[NST], z_empty[NST];
  __shared__ uint64_t sg_full[MAXG], p_ready[MAXG], acc_done[MAXG];
  __shared__ uint32_t tmem_base_s;
  // UMMA descriptors of the context-row ring (as in the forward kernel): [0..3] the dt.z operand of the four 16-token
  // steps, [4..7] z_hi k0, z_hi k1, z_lo k0, z_lo k1; read with volatile loads before the issuer's wait for dt(i)
  __shared__ __align__(16) uint64_t z_desc[NST][8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int idx = blockIdx.x;
  const int cta_rb = idx % p.ctas_per_stream;
  idx /= p.ctas_per_stream;
  const int b = idx % p.batch;
  const int split = idx / p.batch;
  const int rb0 = cta_rb * G;
  const int n_active = (p.n_rb - rb0) < G ? (p.n_rb - rb0) : G;
  const int t_begin = static_cast<int>(static_cast<long>(p.tiles_total) * split / p.nsplit);
  const int t_end = static_cast<int>(static_cast<long>(p.tiles_total) * (split + 1) / p.nsplit);
  const int n = t_end - t_begin;

  if (n <= 0) {  // more splits than tiles: empty partial
    HN_PDL_WAIT();
    for (int g = 0; g < n_active; ++g) {
      const int rb = rb0 + g, h = rb / p.n_ltiles, lt = rb % p.n_ltiles;
      const long row0 = ((static_cast<long>(b) * p.nsplit + split) * p.H + h) * p.L + lt * BM;
      for (int r = threadIdx.x; r < BM; r += blockDim.x)
        if (lt * BM + r < p.L)
          for (int c = 0; c < KD; ++c) p.part[(row0 + r) * KD + c] = 0.f;
    }
    return;
  }
  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < NSB; ++s) {
      mbar_init(&z_full[s], 1);
      mbar_init(&z_empty[s], n_active);
    }
    for (int g = 0; g < MAXG; ++g) {
      mbar_init(&sg_full[g], 1);
      mbar_init(&p_ready[g], 4);
      mbar_init(&acc_done[g], 1);
    }
    fence_mbar_init();
  }
  if (KD == 32 && threadIdx.x >= 32 && threadIdx.x < 32 + NSB) {
    const int st = threadIdx.x - 32;
    const uint32_t z0 = smem_u32(sZ + st * Z_STAGE);
#pragma unroll
    for (int k = 0; k < 4; ++k) z_desc[st][k] = smem_desc(z0 + k * V_KADV, 16, SBO, LAYOUT);
    z_desc[st][4] = smem_desc(z0, 16, SBO, LAYOUT);
    z_desc[st][5] = smem_desc(z0 + 32, 16, SBO, LAYOUT);
    z_desc[st][6] = smem_desc(z0 + Z_BYTES, 16, SBO, LAYOUT);
    z_desc[st][7] = smem_desc(z0 + Z_BYTES + 32, 16, SBO, LAYOUT);
  }
  constexpr int PRODUCER_WARP = 5 * G;
  if (warp == PRODUCER_WARP) tmem_alloc<512>(&tmem_base_s);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  HN_PDL_LAUNCH();
  HN_PDL_WAIT();

  if (warp == PRODUCER_WARP) {
    if (elect_one()) {
      tma_prefetch_desc(&tmR);
      tma_prefetch_desc(&tmDU);
      tma_prefetch_desc(&tmZ);
      mbar_arrive_expect_tx(&q_full, n_active * Q_GROUP);
      for (int g = 0; g < n_active; ++g) {
        const int rb = rb0 + g, h = rb / p.n_ltiles, lt = rb % p.n_ltiles;
        uint8_t* q = sQ + g * Q_GROUP;
        tma_load_3d(q, &tmR, &q_full, h * KD, lt * BM, b);
        tma_load_3d(q + Q_TILE, &tmR, &q_full, p.lo_off + h * KD, lt * BM, b);
        tma_load_3d(q + 2 * Q_TILE, &tmDU, &q_full, h * KD, lt * BM, b);
        tma_load_3d(q + 3 * Q_TILE, &tmDU, &q_full, p.lo_off + h * KD, lt * BM, b);
      }
      for (int i = 0; i < n; ++i) {
        const int s = i % NSB;
        mbar_wait_sleepy(&z_empty[s], ((i / NSB) & 1) ^ 1, 20000);
        mbar_arrive_expect_tx(&z_full[s], Z_STAGE);
        tma_load_3d(sZ + s * Z_STAGE, &tmZ, &z_full[s], 0, (t_begin + i) * BT, b);
        tma_load_3d(sZ + s * Z_STAGE + Z_BYTES, &tmZ, &z_full[s], KD, (t_begin + i) * BT, b);
      }
    }
  } else if (warp >= 4 * G) {
    const int g = warp - 4 * G;
    if (g < n_active && elect_one()) {
      const uint32_t tG = tmem + g * GCOLS;
      const uint32_t q0 = smem_u32(sQ + g * Q_GROUP);
      // S' = R_hi.z_hi + R_lo.z_hi + R_hi.z_lo ; G = DU_hi.z_hi + DU_lo.z_hi over the 16-column steps that hold context
      // columns (KH for the product that carries the fold column, KL for the others), compile-time per mode; dz = the
      // tile's K-operand descriptors (KD 32: from the table)
      auto issue_sg_c = [&](uint32_t z0, const uint64_t (&dz)[4], auto kh_c, auto kl_c, auto kg_c, auto mg_c) {
        constexpr int KH = decltype(kh_c)::value, KL = decltype(kl_c)::value, KG = decltype(kg_c)::value;
        constexpr bool MG = decltype(mg_c)::value;
        auto zh = [&](int k) { return KD == 32 ? dz[k] : smem_desc(z0 + k * 32, 16, SBO, LAYOUT); };
        auto zl = [&](int k) { return KD == 32 ? dz[2 + k] : smem_desc(z0 + Z_BYTES + k * 32, 16, SBO, LAYOUT); };
#pragma unroll
        for (int k = 0; k < KH; ++k) umma_ss(tG, smem_desc(q0 + k * 32, 16, SBO, LAYOUT), zh(k), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < KL; ++k) umma_ss(tG, smem_desc(q0 + Q_TILE + k * 32, 16, SBO, LAYOUT), zh(k), idesc_s, true);
#pragma unroll
        for (int k = 0; k < KL; ++k) umma_ss(tG, smem_desc(q0 + k * 32, 16, SBO, LAYOUT), zl(k), idesc_s, true);
        if (MG)  // merged tail (as in the forward): [R_hi tail | 0 | R_lo tail] . [z_lo tail | 0 | z_hi tail], R_lo rows
                 // arrive re-arranged from launch_small_bwd_prep, the z rows from the context-row builder
          umma_ss(tG, smem_desc(q0 + Q_TILE + 32, 16, SBO, LAYOUT), zl(1), idesc_s, true);
#pragma unroll
        for (int k = 0; k < KG; ++k) umma_ss(tG + 64, smem_desc(q0 + 2 * Q_TILE + k * 32, 16, SBO, LAYOUT), zh(k), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < KG; ++k) umma_ss(tG + 64, smem_desc(q0 + 3 * Q_TILE + k * 32, 16, SBO, LAYOUT), zh(k), idesc_s, true);
        umma_commit(&sg_full[g]);
      };
      using std::integral_constant;
      const int mode = p.mode;
      auto issue_sg = [&](uint32_t z0, const uint64_t (&dz)[4]) {
        constexpr int KF = KD / 16;
        using I1 = integral_constant<int, 1>;
        using I2 = integral_constant<int, 2>;
        if (KD == 32) {
          if (mode == 2) { issue_sg_c(z0, dz, I2{}, I1{}, I2{}, std::true_type{}); return; }
          if (mode == 0) { issue_sg_c(z0, dz, I1{}, I1{}, I1{}, std::false_type{}); return; }
          if (mode == 1) { issue_sg_c(z0, dz, I2{}, I1{}, I1{}, std::false_type{}); return; }
        }
        issue_sg_c(z0, dz, integral_constant<int, KF>{}, integral_constant<int, KF>{}, integral_constant<int, KF>{},
                   std::false_type{});
      };
      auto load_desc4 = [&](int stage, int first, uint64_t (&d)[4]) {
        const uint32_t a = smem_u32(&z_desc[stage][first]);
        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(d[0]), "=l"(d[1]) : "r"(a));
        asm volatile("ld.volatile.shared.v2.u64 {%0, %1}, [%2];" : "=l"(d[2]), "=l"(d[3]) : "r"(a + 16));
      };
      auto z_stage = [&](int i) {  // wait until tile i's context rows have landed; -> their shared address
        const int s = i % NSB;
        mbar_wait(&z_full[s], (i / NSB) & 1);
        return smem_u32(sZ + s * Z_STAGE);
      };
      mbar_wait(&q_full, 0);
      {
        const uint32_t za = z_stage(0);
        uint64_t dz[4] = {0, 0, 0, 0};
        if (KD == 32) load_desc4(0, 4, dz);
        fence_after_sync();
        issue_sg(za, dz);
      }
      for (int i = 0; i < n; ++i) {
        // single-buffered S' / G: the softmax warps of this row block idle from the moment they publish dt(i) until
        // S'(i+1) and G(i+1) are complete, so everything that does not depend on dt(i) happens BEFORE the wait for it
        // (the dt.z descriptors, the arrival of tile i+1's context rows and their descriptors). The tensor pipe executes
        // in issue order: dt(i).z has consumed S' columns 0..31 before S'(i+1) overwrites them.
        const int s = i % NSB;
        const uint32_t z0 = smem_u32(sZ + s * Z_STAGE);
        uint64_t dv[4], dz[4] = {0, 0, 0, 0};
        if (KD == 32) {
          load_desc4(s, 0, dv);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) dv[k] = smem_desc(z0 + k * V_KADV, 16, SBO, LAYOUT);
        }
        const bool more = i + 1 < n;
        uint32_t z1 = 0;
        if (more) {
          z1 = z_stage(i + 1);
          if (KD == 32) load_desc4((i + 1) % NSB, 4, dz);
        }
        mbar_wait_sleepy(&p_ready[g], i & 1, 20000);
        fence_after_sync();
#pragma unroll
        for (int k = 0; k < BT / 16; ++k) umma_ts(tG + 128, tG + k * 8, dv[k], idesc_u, (i | k) != 0);
        umma_commit(&z_empty[s]);
        if (more) issue_sg(z1, dz);
        if (i + 1 == n) umma_commit(&acc_done[g]);
      }
    }
  } else {
    const int g = warp >> 2;
    if (g < n_active) {
      const int rb = rb0 + g, h = rb / p.n_ltiles, lt = rb % p.n_ltiles;
      const uint32_t lane_base = (warp & 3) * 32;
      const int trow = lane_base + lane;
      const int row = lt * BM + trow;
      const uint32_t tL = tmem_addr(tmem + g * GCOLS, lane_base, 0);
      const bool has_mask = p.mask_bits != nullptr;
      float a_l = 0.f, d_l = 0.f;
      if (row < p.L) {
        const long R = (static_cast<long>(b) * p.L + row) * p.H + h;
        a_l = p.row_a[R];
        d_l = p.row_d[R];
      }
      const float ad = -a_l * d_l;
      for (int i = 0; i < n; ++i) {
        mbar_wait(&sg_full[g], i & 1);
        fence_after_sync();
        uint64_t bits = ~0ull;
        if (has_mask) bits = p.mask_bits[static_cast<long>(b) * p.tiles_total + t_begin + i];
        // four 16-column chunks (16-register loads, 8-register stores: a 32-register store tuple costs ~30 MOVs per tile,
        // see the forward kernel); x = S' <= P_SHIFT here (M is the true row max), so the lean polynomial needs no
        // overflow test
        uint32_t pk[4][8];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t s[16], gq[16];
          tmem_ld16(tL + c * 16, s);
          tmem_ld16(tL + 64 + c * 16, gq);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float x0 = __uint_as_float(s[2 * j]), x1 = __uint_as_float(s[2 * j + 1]);
            // w = a (G - delta)
            const float w0 = fmaf(__uint_as_float(gq[2 * j]), a_l, ad), w1 = fmaf(__uint_as_float(gq[2 * j + 1]), a_l, ad);
            if (j & 1) {  // half2 polynomial exponentials, product in half2
              uint32_t tb;
              const uint32_t e = ex2_pair_h2_lean(x0, x1, tb);
              const __half2 wh = __floats2half2_rn(w0, w1);
              const __half2 r = __hmul2(*reinterpret_cast<const __half2*>(&e), wh);
              pk[c][j] = *reinterpret_cast<const uint32_t*>(&r);
            } else {
              pk[c][j] = pack_half2(ex2_mufu(x0) * w0, ex2_mufu(x1) * w1);
            }
          }
          if (has_mask) {
            const uint32_t mb = static_cast<uint32_t>(bits >> (16 * c)) & 0xFFFFu;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t keep = (((mb >> (2 * j)) & 1u) ? 0x0000FFFFu : 0u) | (((mb >> (2 * j + 1)) & 1u) ? 0xFFFF0000u : 0u);
              pk[c][j] &= keep;
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st8(tL + c * 8, pk[c]);  // dt(i) over S' columns 0..31 (all 64 columns of S' and G have been read)
        tmem_wait_st();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[g]);
        __syncwarp();
      }
      const long part_row = ((static_cast<long>(b) * p.nsplit + split) * p.H + h) * p.L + row;
      mbar_wait(&acc_done[g], 0);
      fence_after_sync();
#pragma unroll
      for (int c = 0; c < KD; c += 32) {
        uint32_t u[32];
        tmem_ld32(tL + 128 + c, u);
        tmem_wait_ld();
        if (row < p.L) {
          float4* dst = reinterpret_cast<float4*>(p.part + part_row * KD + c);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            dst[j] = make_float4(__uint_as_float(u[4 * j]), __uint_as_float(u[4 * j + 1]), __uint_as_float(u[4 * j + 2]),
                                 __uint_as_float(u[4 * j + 3]));
        }
        __syncwarp();
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == PRODUCER_WARP) tmem_dealloc<512>(tmem);
}

template <int KD, int G>
int launch_small_bwd_t(const SmallBwdTcArgs& a, cudaStream_t stream) {
  CUtensorMap tmR, tmDU, tmZ;
  const CUtensorMapSwizzle swz = KD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  const uint64_t ld = a.rq_ld;
  if (!make_tmap_3d_f16(&tmR, a.rq, a.batch, a.L, ld, ld * 2, static_cast<uint64_t>(a.L) * ld * 2, BM, KD, swz) ||
      !make_tmap_3d_f16(&tmDU, a.duq, a.batch, a.L, ld, ld * 2, static_cast<uint64_t>(a.L) * ld * 2, BM, KD, swz) ||
      !make_tmap_3d_f16(&tmZ, a.z, a.batch, a.N, 2 * KD, static_cast<uint64_t>(2 * KD) * 2,
                        static_cast<uint64_t>(a.N) * 2 * KD * 2, BT, KD, swz)) {
    set_error("attention backward: cuTensorMapEncodeTiled failed");
    return -2;
  }
  SmallBwdDev p;
  p.L = a.L;
  p.H = a.H;
  p.batch = a.batch;
  p.nsplit = a.nsplit;
  p.n_ltiles = (a.L + BM - 1) / BM;
  p.n_rb = p.n_ltiles * a.H;
  p.ctas_per_stream = (p.n_rb + G - 1) / G;
  p.tiles_total = static_cast<int>((a.N + BT - 1) / BT);
  p.lo_off = a.lo_off;
  p.mode = a.C <= 15 ? 0 : a.C == 16 ? 1 : (KD == 32 && a.merged_tail && a.C >= 17 && a.C <= 23) ? 2 : 3;
  p.N = a.N;
  p.mask_bits = a.mask_bits;
  p.row_a = a.row_a;
  p.row_d = a.row_d;
  p.part = a.part;
  constexpr int SMEM_NEED = G * 4 * BM * KD * 2 + (KD == 64 ? 4 : NST) * 2 * BT * KD * 2 + 1024;
  constexpr int SMEM = SMEM_NEED > 120 * 1024 ? SMEM_NEED : 120 * 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  const long grid = static_cast<long>(p.ctas_per_stream) * a.batch * a.nsplit;
  HN_REQUIRE(grid > 0 && grid < 2147483647L, "attention backward: grid too large");
  HN_CHECK_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel<KD, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  HN_CHECK_CUDA(launch_k(attn_small_bwd_kernel<KD, G>, dim3(static_cast<unsigned>(grid)), dim3((5 * G + 1) * 32), SMEM, stream,
                         tmR, tmDU, tmZ, p));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <int KD, int G, int PMODE, bool SPLIT, int SW>
int launch_small_t(const AttnArgs& a, cudaStream_t stream) {
  CUtensorMap tmQ, tmZ;
  const CUtensorMapSwizzle swz = KD == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (!make_tmap_3d_f16(&tmQ, a.Q, a.batch, a.L, a.q_ld, static_cast<uint64_t>(a.q_ld) * 2,
                        static_cast<uint64_t>(a.L) * a.q_ld * 2, BM, KD, swz) ||
      !make_tmap_3d_f16(&tmZ, a.KV, a.batch, a.N, a.kv_ld, static_cast<uint64_t>(a.kv_ld) * 2,
                        static_cast<uint64_t>(a.N) * a.kv_ld * 2, BT, KD, swz)) {
    set_error("attention: cuTensorMapEncodeTiled failed");
    return -2;
  }
  SmallDev p;
  p.L = a.L;
  p.H = a.H;
  p.batch = a.batch;
  p.nsplit = a.nsplit;
  p.n_ltiles = (a.L + BM - 1) / BM;
  p.n_rb = p.n_ltiles * a.H;
  p.q_lo_off = a.q_lo_off;
  p.ctas_per_stream = (p.n_rb + G - 1) / G;
  p.tiles_total = static_cast<int>((a.N + BT - 1) / BT);
  p.N = a.N;
  p.c_ones = a.c_ones;
  p.rt_zero = 0;
  p.merged = (a.z_tail_merged && SPLIT && KD == 32 && SW == 4 && a.c_ones >= 17 && a.c_ones <= 23) ? 1 : 0;
  p.mode = a.c_ones <= 15 ? 0 : a.c_ones == 16 ? 1 : p.merged ? 2 : 3;
  p.mask_bits = a.mask_bits;
  p.part_acc = a.part_acc;
  p.part_ml = a.part_ml;
  // at least 120 KB so that a second CTA can never share the SM (each CTA allocates all 512 TMEM columns)
  constexpr int SMEM_NEED = (SPLIT ? 2 : 1) * (G * BM * KD * 2 + NST * BT * KD * 2) + 1024;
  constexpr int SMEM = SMEM_NEED > 120 * 1024 ? SMEM_NEED : 120 * 1024;
  static_assert(SMEM <= 227 * 1024, "shared memory budget");
  const long grid = static_cast<long>(p.ctas_per_stream) * a.batch * a.nsplit;
  HN_REQUIRE(grid > 0 && grid < 2147483647L, "attention: grid too large");
  p.mask_words = p.tiles_total;
  HN_CHECK_CUDA(cudaFuncSetAttribute(attn_small_kernel<KD, G, PMODE, SPLIT, SW>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  HN_CHECK_CUDA(launch_k(attn_small_kernel<KD, G, PMODE, SPLIT, SW>, dim3(static_cast<unsigned>(grid)),
                         dim3(((SW + 1) * G + 1) * 32), SMEM, stream, tmQ, tmZ, p));
  HN_CHECK_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

static int poly_mode() {
#ifdef HN_DEBUG
  static int pmode = -1;  // tuning knob of debug builds (HN_POLY_MODE): share of the exponentials on the FMA pipe
  if (pmode < 0) {
    const char* e = getenv("HN_POLY_MODE");
    pmode = (e != nullptr && e[0] >= '0' && e[0] <= '9') ? e[0] - '0' : 6;
  }
  return pmode;
#else
  return 6;  // every other column pair on the FMA pipe (half2 polynomial)
#endif
}

int small_attention_groups(int kd) { return kd == 32 ? 3 : 2; }

int launch_small_attention_bwd(const SmallBwdTcArgs& a, cudaStream_t stream) {
  HN_REQUIRE(a.batch > 0 && a.L > 0 && a.H > 0 && a.N > 0 && a.nsplit > 0, "attention backward: empty problem");
  HN_REQUIRE(a.kd == 32 || a.kd == 64, "attention backward: context rows must be 32 or 64 wide");
  HN_REQUIRE(a.rq_ld % 8 == 0 && a.lo_off >= a.H * a.kd && a.rq_ld >= a.lo_off + a.H * a.kd, "attention backward: bad row layout");
  if (a.kd == 64) return launch_small_bwd_t<64, 2>(a, stream);
  return launch_small_bwd_t<32, 3>(a, stream);
}

// Split the token axis so that the grid is a whole number of waves of one CTA per SM while each CTA still
// streams enough tiles to amortise its prologue / epilogue.
int small_attention_pick_nsplit(int batch, int L, int H, long N, int kd) {
#ifdef HN_DEBUG
  {  // tuning knob of debug builds: force the number of token splits of long axes
    static const char* e = getenv("HN_SMALL_NSPLIT");
    if (e != nullptr && atoi(e) > 0 && N > 100000) return atoi(e);
  }
#endif
  const int G = small_attention_groups(kd);
  const long n_rb = static_cast<long>((L + BM - 1) / BM) * H;
  const long base = ((n_rb + G - 1) / G) * batch;
  const long tiles = (N + BT - 1) / BT;
  const long slots = 148;
  if (tiles <= 16) return 1;
  long best = 1;
  double best_cost = 1e30;
  const long max_split = tiles / 16 > 0 ? tiles / 16 : 1;
  for (long s = 1; s <= max_split && s <= 1024; ++s) {
    const long ctas = base * s;
    const long waves = (ctas + slots - 1) / slots;
    const long per = (tiles + s - 1) / s;
    // every extra split also costs a set of partials (written here, re-read by the combine kernel)
    const double cost = static_cast<double>(waves) * (per + 8.0) * (1.0 + 0.001 * s);
    if (cost < best_cost * 0.999) {
      best_cost = cost;
      best = s;
    }
  }
  return static_cast<int>(best);
}

// Softmax warps per row block. The product runs 4 (thread = latent row). The 8-warp layout (16 rows per warp, built on
// the .16x256b / .16x128b TMEM fragments) doubles the warps per scheduler but was measured SLOWER on the B200 (cfg 1
// volume, batch 4: 2.45 ms vs 2.36 ms; poly-share sweep in profiles/r2_attn_small_experiments.md): the kernel is bound
// by instruction issue (4.8 instructions per element at thread-per-row, 5.6 at 16 rows per warp), not by latency.
// It stays in DEBUG builds (HN_SMALL_SW=8) as the record of that experiment.
#ifdef HN_DEBUG
static int softmax_warps() {
  static int sw = 0;
  if (sw == 0) {
    const char* e = getenv("HN_SMALL_SW");
    sw = (e != nullptr && e[0] == '8') ? 8 : 4;
  }
  return sw;
}
#endif

template <int PMODE>
static int launch_small_variant(const AttnArgs& a, cudaStream_t stream) {
#ifdef HN_DEBUG
  if (softmax_warps() == 8) {
    if (a.precise) {
      if (a.kd == 64) return launch_small_t<64, 2, PMODE, true, 8>(a, stream);
      return launch_small_t<32, 3, PMODE, true, 8>(a, stream);
    }
    if (a.kd == 64) return launch_small_t<64, 2, PMODE, false, 8>(a, stream);
    return launch_small_t<32, 3, PMODE, false, 8>(a, stream);
  }
#endif
  if (a.precise) {
    if (a.kd == 64) return launch_small_t<64, 2, PMODE, true, 4>(a, stream);
    return launch_small_t<32, 3, PMODE, true, 4>(a, stream);
  }
  if (a.kd == 64) return launch_small_t<64, 2, PMODE, false, 4>(a, stream);
  return launch_small_t<32, 3, PMODE, false, 4>(a, stream);
}

int launch_small_attention(const AttnArgs& a, cudaStream_t stream) {
  HN_REQUIRE(a.batch > 0 && a.L > 0 && a.H > 0 && a.N > 0 && a.nsplit > 0, "attention: empty problem");
  HN_REQUIRE(a.kd == 32 || a.kd == 64, "attention: shared-context rows must be 32 or 64 wide");
#ifndef HN_DEBUG
  HN_REQUIRE(a.kv_ld == (a.precise ? 2 * a.kd : a.kd), "attention: shared-context rows must be dense ([hi | lo] when split)");
#endif
  HN_REQUIRE(!a.precise || a.q_lo_off >= a.H * a.kd, "attention: split Q' rows need the lo-part offset");
  HN_REQUIRE(a.q_ld % 8 == 0, "attention: row pitches must be multiples of 8 elements");
  HN_REQUIRE(a.N < (1L << 31), "attention: token axis too long");
  HN_REQUIRE(a.c_ones >= 1 && a.c_ones < a.kd, "attention: ones column must lie inside the context row");
  switch (poly_mode()) {
#ifdef HN_DEBUG
    case 0: return launch_small_variant<0>(a, stream);
    case 5: return launch_small_variant<5>(a, stream);
    case 7: return launch_small_variant<7>(a, stream);
    case 8: return launch_small_variant<8>(a, stream);
#endif
    default: return launch_small_variant<6>(a, stream);
  }
}

}  // namespace hn

