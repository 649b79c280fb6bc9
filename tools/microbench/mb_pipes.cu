// mb_pipes.cu — B200 pipe microbenchmarks that size the softmax side of the streaming attention kernel:
//   MUFU.EX2 rate (f32 / f16 / bf16 forms), FMA-pipe polynomial exp2 rate, tcgen05.ld / tcgen05.st bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mb_pipes mb_pipes.cu ; run on the GPU box.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_ex2_f32(float* out, float a, int iters, long long* clk) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = a + threadIdx.x * 1e-3f + j * 0.1f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j]));
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
__global__ void k_ex2_h2(float* out, float a, int iters, long long* clk) {
  uint32_t x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = 0x30003000u + threadIdx.x + j;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(x[j]));
  }
  long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(s);
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// degree-3 Cody-Waite exp2 on the FMA/ALU pipes (no MUFU)
__device__ __forceinline__ float poly_exp2(float x) {
  const float t = x + 12582912.f;               // round to nearest integer (magic 1.5*2^23)
  const float f = x - (t - 12582912.f);          // f in [-0.5, 0.5]
  float p = 0.0555041086f;                       // minimax-ish coefficients (placeholders for timing)
  p = fmaf(p, f, 0.2402265069f);
  p = fmaf(p, f, 0.6931471805f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__global__ void k_poly(float* out, float a, int iters, long long* clk) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = a + threadIdx.x * 1e-3f + j * 0.1f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = poly_exp2(x[j]) - 1.0f;
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
// mix: per 8 elements, NP via poly and 8-NP via MUFU
template <int NP>
__global__ void k_mix(float* out, float a, int iters, long long* clk) {
  float x[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) x[j] = a + threadIdx.x * 1e-3f + j * 0.1f;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < NP) x[j] = poly_exp2(x[j]) - 1.0f;
      else { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[j])); x[j] -= 1.0f; }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// TMEM load / store bandwidth: NW warps (multiple of 4), each reads its 32-lane quadrant, 32 columns x iters
__global__ void k_tmem_ld(float* out, int iters, long long* clk, int mode) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t = tbase + (((warp & 3) * 32u) << 16);
  uint32_t r[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) r[j] = threadIdx.x + j;
  // initialise the columns we read
  for (int c = 0; c < 256; c += 32) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(t + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  uint32_t acc = 0;
  long long t0 = clock64();
  if (mode == 0) {
    for (int i = 0; i < iters; ++i) {
      const uint32_t c = (i & 7) * 32;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(t + c) : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      acc ^= r[0] ^ r[31];
    }
  } else {
    for (int i = 0; i < iters; ++i) {
      const uint32_t c = (i & 7) * 32;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(t + c), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tbase) : "memory");
}

template <typename F>
static double run(F launch, long long* dclk, int nblk) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch();  // warm
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  launch();
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> h(nblk);
  cudaMemcpy(h.data(), dclk, nblk * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (auto v : h) avg += v; avg /= nblk;
  printf("   [%.3f ms, avg %.0f clk/block]", ms, avg);
  return avg;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs\n", prop.name, sms);
  float* out; long long* clk;
  CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 1024)); CK(cudaMalloc(&clk, sizeof(long long) * sms * 8));
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    const int thr = warps * 32;
    printf("warps/SM=%d\n", warps);
    double c;
    c = run([&] { k_ex2_f32<<<sms, thr>>>(out, 0.5f, iters, clk); }, clk, sms);
    printf(" ex2.f32      : %.2f elem/clk/SM\n", (double)thr * 8 * iters / c);
    c = run([&] { k_ex2_h2<<<sms, thr>>>(out, 0.5f, iters, clk); }, clk, sms);
    printf(" ex2.f16x2    : %.2f elem/clk/SM\n", (double)thr * 16 * iters / c);
    c = run([&] { k_poly<<<sms, thr>>>(out, 0.5f, iters, clk); }, clk, sms);
    printf(" poly exp2    : %.2f elem/clk/SM\n", (double)thr * 8 * iters / c);
    c = run([&] { k_mix<2><<<sms, thr>>>(out, 0.5f, iters, clk); }, clk, sms);
    printf(" mix 2/8 poly : %.2f elem/clk/SM\n", (double)thr * 8 * iters / c);
    c = run([&] { k_mix<3><<<sms, thr>>>(out, 0.5f, iters, clk); }, clk, sms);
    printf(" mix 3/8 poly : %.2f elem/clk/SM\n", (double)thr * 8 * iters / c);
    c = run([&] { k_mix<4><<<sms, thr>>>(out, 0.5f, iters, clk); }, clk, sms);
    printf(" mix 4/8 poly : %.2f elem/clk/SM\n", (double)thr * 8 * iters / c);
  }
  for (int warps : {4, 8}) {
    const int thr = warps * 32;
    double c;
    c = run([&] { k_tmem_ld<<<sms, thr>>>(out, iters, clk, 0); }, clk, sms);
    printf(" tcgen05.ld x32, %d warps, 1 CTA/SM: %.1f B/clk/SM\n", warps, (double)thr * 32 * 4 * iters / c);
    c = run([&] { k_tmem_ld<<<sms, thr>>>(out, iters, clk, 1); }, clk, sms);
    printf(" tcgen05.st x32, %d warps, 1 CTA/SM: %.1f B/clk/SM\n", warps, (double)thr * 32 * 4 * iters / c);
  }
  {
    double c = run([&] { k_tmem_ld<<<2 * sms, 128>>>(out, iters, clk, 0); }, clk, 2 * sms);
    printf(" tcgen05.ld x32, 4 warps, 2 CTA/SM: %.1f B/clk/SM\n", 2.0 * 128 * 32 * 4 * iters / c);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
