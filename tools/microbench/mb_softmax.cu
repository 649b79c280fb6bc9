// mb_softmax.cu — upper bound of the softmax side of attn_small_kernel without any cross-warp synchronisation:
// every warp loops  tcgen05.ld 2 x 32 columns -> 64 exponentials (MUFU / half2 polynomial mix) -> fp16 pack ->
// packed max -> tcgen05.st 32 columns,  on its own TMEM lane quadrant. Sweeps the polynomial share and the number of
// warps per scheduler. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_softmax mb_softmax.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ float ex2_mufu(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t vmaxu2(uint32_t a, uint32_t b) { uint32_t r; asm("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) { __half2 h = __floats2half2_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ uint32_t ex2_pair_h2(float x0, float x1) {
  __half2 xh = __floats2half2_rn(x0, x1);
  xh = __hmax2(xh, __float2half2_rn(-24.f));
  const __half2 magic = __float2half2_rn(1546.f);
  const __half2 t = __hadd2(xh, magic);
  const __half2 f = __hsub2(xh, __hsub2(t, magic));
  __half2 p = __float2half2_rn(0.05517167f);
  p = __hfma2(p, f, __float2half2_rn(0.24261112f));
  p = __hfma2(p, f, __float2half2_rn(0.69326099f));
  p = __hfma2(p, f, __float2half2_rn(0.99992807f));
  const uint32_t e = (*reinterpret_cast<const uint32_t*>(&t) << 10) & 0xFC00FC00u;
  uint32_t r;
  asm("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&p)), "r"(e));
  const __half2 scaled = __hmul2(*reinterpret_cast<const __half2*>(&r), __float2half2_rn(0.0009765625f));
  return *reinterpret_cast<const uint32_t*>(&scaled);
}
// variant: clamp at -14 (the caller folds +10 into S, so this is x >= -24), result built directly as a normal fp16
// (no final scale); SHF selects the shifter (ALU pipe) for the exponent move
template <bool SHF, int DEG>
__device__ __forceinline__ uint32_t ex2_pair_h2_v(float x0, float x1) {
  __half2 xh = __floats2half2_rn(x0, x1);
  xh = __hmax2(xh, __float2half2_rn(-14.f));
  const __half2 magic = __float2half2_rn(1536.f);
  const __half2 t = __hadd2(xh, magic);
  const __half2 f = __hsub2(xh, __hsub2(t, magic));
  __half2 p;
  if (DEG == 3) {
    p = __float2half2_rn(0.05517167f);
    p = __hfma2(p, f, __float2half2_rn(0.24261112f));
    p = __hfma2(p, f, __float2half2_rn(0.69326099f));
    p = __hfma2(p, f, __float2half2_rn(0.99992807f));
  } else {
    p = __float2half2_rn(0.2402265f);
    p = __hfma2(p, f, __float2half2_rn(0.6931472f));
    p = __hfma2(p, f, __float2half2_rn(1.0f));
  }
  uint32_t e;
  const uint32_t tb = *reinterpret_cast<const uint32_t*>(&t);
  if (SHF) asm("shf.l.clamp.b32 %0, %1, %1, 10;" : "=r"(e) : "r"(tb));
  else e = tb << 10;
  e &= 0xFC00FC00u;
  uint32_t r;
  asm("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&p)), "r"(e));
  return r;
}
// MUFU pair through the packed half2 form: one pack of the arguments, ex2.approx.f16x2, result already packed
__device__ __forceinline__ uint32_t ex2_pair_mufu_h2(float x0, float x1) {
  __half2 xh = __floats2half2_rn(x0, x1);
  uint32_t r;
  asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(*reinterpret_cast<const uint32_t*>(&xh)));
  return r;
}
// bf16 pack by truncation: one PRMT per pair, no conversion unit
__device__ __forceinline__ uint32_t pack_bf16_trunc(float lo, float hi) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(lo)), "r"(__float_as_uint(hi)));
  return r;
}
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]),
      "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
}

// MODE: 0 all MUFU + F2FP | 5 every third pair half2 polynomial | 6 every second pair | 9 no exponentials (pack only)
//       10 all MUFU, bf16 truncation pack (no F2FP) | 11 nothing but ld / st
//       2x: variant polynomial without final scale: 20 = 1/3, 21 = 1/2, 22 = 1/2 + ALU shift, 23 = 5/8 + ALU shift,
//           24 = 3/8 + ALU shift, 25 = 1/2 degree 2 + ALU shift, 26 = 3/4 + ALU shift, 27 = all polynomial + ALU shift
template <int MODE>
__device__ __forceinline__ bool poly_pair(int j) {
  return MODE == 5 || MODE == 20 ? (j % 3) == 1
         : MODE == 6 || MODE == 21 || MODE == 22 || MODE == 25 ? (j % 2) == 1
         : MODE == 23 ? ((j % 8) != 0 && (j % 8) != 3 && (j % 8) != 6)
         : MODE == 24 ? ((j % 8) == 1 || (j % 8) == 4 || (j % 8) == 6)
         : MODE == 28 ? (j % 2) == 1
         : MODE == 26 ? (j % 4) != 0
         : MODE == 27;
}

template <int MODE, int NW>
__global__ void __launch_bounds__(NW * 32, 1) k_softmax(float* out, int iters, long long* clk) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int g = warp >> 2;
  const uint32_t t = tbase + (((warp & 3) * 32u) << 16) + g * 128;  // 128 columns per group of four warps
  {
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(-0.01f * ((threadIdx.x + 7 * j) % 97));
    st32(t, r); st32(t + 32, r); st32(t + 64, r); st32(t + 96, r);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const uint32_t tS = t + (i & 1) * 64;
    uint32_t pk[32];
    uint32_t pmax = 0;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t s[32];
      ld32(tS + c * 32, s);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float x0 = __uint_as_float(s[2 * j]), x1 = __uint_as_float(s[2 * j + 1]);
        if (MODE == 11) pk[c * 16 + j] = s[2 * j] ^ s[2 * j + 1];
        else if (MODE == 9) pk[c * 16 + j] = pack_half2(x0, x1) & 0x3FFF3FFFu;
        else if (MODE == 10) pk[c * 16 + j] = pack_bf16_trunc(ex2_mufu(x0), ex2_mufu(x1));
        else if (MODE >= 20 && poly_pair<MODE>(j)) pk[c * 16 + j] = ex2_pair_h2_v<(MODE >= 22), (MODE == 25 ? 2 : 3)>(x0, x1);
        else if (poly_pair<MODE>(j)) pk[c * 16 + j] = ex2_pair_h2(x0, x1);
        else if (MODE == 28) pk[c * 16 + j] = ex2_pair_mufu_h2(x0, x1);
        else pk[c * 16 + j] = pack_half2(ex2_mufu(x0), ex2_mufu(x1));
        if (MODE != 11) pmax = vmaxu2(pmax, pk[c * 16 + j]);
      }
    }
    const bool big = ((pmax & 0xFFFFu) > 0x5000u) || ((pmax >> 16) > 0x5000u);
    if (__any_sync(0xffffffffu, big)) acc += 1;
    // keep the stored data equal to the loaded data so every iteration does the same work: store elsewhere
    st32(t + 64 * 0 + 96 * 0 + (i & 1) * 64 + 0, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    // restore the S values this tile overwrote (cheap: one more st of constants would change the mix; instead the
    // next read of this buffer simply sees fp16 pairs reinterpreted as small floats - same instruction stream)
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

template <int MODE, int NW>
static int run(const char* name) {
  const int nblk = 148, iters = 2000;
  float* out; long long* clk;
  CK(cudaMalloc(&out, sizeof(float) * nblk * NW * 32));
  CK(cudaMalloc(&clk, sizeof(long long) * nblk));
  k_softmax<MODE, NW><<<nblk, NW * 32>>>(out, iters, clk);
  CK(cudaDeviceSynchronize());
  k_softmax<MODE, NW><<<nblk, NW * 32>>>(out, iters, clk);
  CK(cudaDeviceSynchronize());
  long long h[148];
  CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < nblk; ++i) avg += h[i];
  avg /= nblk;
  const double per_round = avg / iters;  // clocks for every warp of the SM to finish one 64-column tile
  const double elems = 64.0 * 32 * NW;
  printf("%-46s warps/SM %2d (%d per scheduler): %7.0f clk per round, %5.2f elem/clk/SM, %6.0f clk per 3-row-block round\n",
         name, NW, NW / 4, per_round, elems / per_round, per_round * 12.0 / NW);
  cudaFree(out); cudaFree(clk);
  return 0;
}

int main() {
  run<11, 12>("ld/st only");
  run<9, 12>("pack + max only (skeleton)");
  run<0, 12>("all MUFU + F2FP pack");
  run<10, 12>("all MUFU + bf16 truncation pack (PRMT)");
  run<5, 12>("1/3 half2 polynomial (production)");
  run<6, 12>("1/2 half2 polynomial");
  run<20, 12>("1/3 polynomial, no final scale");
  run<21, 12>("1/2 polynomial, no final scale");
  run<22, 12>("1/2 polynomial, no final scale, ALU shift");
  run<24, 12>("3/8 polynomial, no final scale, ALU shift");
  run<23, 12>("5/8 polynomial, no final scale, ALU shift");
  run<26, 12>("3/4 polynomial, no final scale, ALU shift");
  run<27, 12>("all polynomial, no final scale, ALU shift");
  run<25, 12>("1/2 degree-2 polynomial, no scale, ALU shift");
  run<28, 12>("1/2 polynomial (no scale, ALU shift) + MUFU pairs as ex2.f16x2");
  run<22, 16>("1/2 polynomial, no final scale, ALU shift");
  run<23, 16>("5/8 polynomial, no final scale, ALU shift");
  return 0;
}
