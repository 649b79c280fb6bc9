// Probe (run on a B200): register <-> TMEM mapping of tcgen05.ld / tcgen05.st shapes .16x256b and .16x128b, found by
// writing a known pattern with the .32x32b shape (thread i <-> lane i, register j <-> column j) and reading it back.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_tmem_layout probe_tmem_layout.cu && ./probe_tmem_layout
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../healnet_b200/csrc/tc05.cuh"
using namespace tc05;

__global__ void probe(uint32_t* out_ld, uint32_t* out_st) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<128>(&tbase);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tm = tbase;
  // pattern: value(lane, col) = lane * 1000 + col, 64 columns
  uint32_t v[32];
  for (int half = 0; half < 2; ++half) {
    for (int j = 0; j < 32; ++j) v[j] = (warp * 32 + lane) * 1000 + half * 32 + j;
    tmem_st32(tmem_addr(tm, warp * 32, half * 32), v);
  }
  tmem_wait_st();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  // read 16 lanes x 64 columns with .16x256b.x8 (32 registers) from lane offsets 0 and 16 of this warp's quadrant
  for (int h = 0; h < 2; ++h) {
    uint32_t r[32];
    const uint32_t a = tmem_addr(tm, warp * 32 + h * 16, 0);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(a)
        : "memory");
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out_ld[((warp * 2 + h) * 32 + lane) * 32 + j] = r[j];
  }
  __syncthreads();
  // store with .16x128b.x8 (16 registers: 16 lanes x 32 columns) into columns 64..95, register j of thread t = t*100 + j
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
    for (int j = 0; j < 16; ++j) r[j] = (h * 32 + lane) * 100 + j;
    const uint32_t a = tmem_addr(tm, warp * 32 + h * 16, 64);
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
  }
  tmem_wait_st();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  uint32_t w[32];
  tmem_ld32(tmem_addr(tm, warp * 32, 64), w);
  tmem_wait_ld();
  for (int j = 0; j < 32; ++j) out_st[(warp * 32 + lane) * 32 + j] = w[j];
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tm);
}

int main() {
  uint32_t *d_ld, *d_st;
  cudaMalloc(&d_ld, 4 * 2 * 32 * 32 * 4);
  cudaMalloc(&d_st, 128 * 32 * 4);
  probe<<<1, 128>>>(d_ld, d_st);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h_ld[4 * 2 * 32 * 32], h_st[128 * 32];
  cudaMemcpy(h_ld, d_ld, sizeof(h_ld), cudaMemcpyDeviceToHost);
  cudaMemcpy(h_st, d_st, sizeof(h_st), cudaMemcpyDeviceToHost);
  printf("LD .16x256b.x8, warp 0, lane-offset 0: thread t register j -> (lane, col)\n");
  for (int t = 0; t < 8; ++t) {
    printf("t=%2d:", t);
    for (int j = 0; j < 8; ++j) printf(" r%d=(%u,%u)", j, h_ld[t * 32 + j] / 1000, h_ld[t * 32 + j] % 1000);
    printf("\n");
  }
  // check the conjectured layout: reg 4k + 2*a + e of thread t = (lane base + t/4 + 8a, col 8k + 2(t%4) + e)
  int bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int h = 0; h < 2; ++h)
      for (int t = 0; t < 32; ++t)
        for (int k = 0; k < 8; ++k)
          for (int a = 0; a < 2; ++a)
            for (int ee = 0; ee < 2; ++ee) {
              const uint32_t got = h_ld[((w * 2 + h) * 32 + t) * 32 + 4 * k + 2 * a + ee];
              const uint32_t want = (w * 32 + h * 16 + t / 4 + 8 * a) * 1000 + 8 * k + 2 * (t % 4) + ee;
              if (got != want) ++bad;
            }
  printf("LD conjecture (reg 4k+2a+e of thread t = lane t/4+8a, col 8k+2(t%%4)+e): %d mismatches\n", bad);
  printf("ST .16x128b.x8 readback (32x32b), warp 0 lanes 0..3 and 16..17: column -> t*100 + j\n");
  for (int l : {0, 1, 2, 8, 16, 17}) {
    printf("lane %2d:", l);
    for (int c = 0; c < 8; ++c) printf(" c%d=%u", c, h_st[l * 32 + c]);
    printf("\n");
  }
  // conjecture: reg 2k + a of thread t -> (lane t/4 + 8a, col 4k + t%4)
  bad = 0;
  for (int w = 0; w < 4; ++w)
    for (int l = 0; l < 32; ++l)
      for (int c = 0; c < 32; ++c) {
        const int h = l / 16, lr = l % 16, a = lr / 8, t = (lr % 8) * 4 + c % 4, k = c / 4;
        const uint32_t want = (h * 32 + t) * 100 + 2 * k + a;
        if (h_st[(w * 32 + l) * 32 + c] != want) ++bad;
      }
  printf("ST conjecture (reg 2k+a of thread t -> lane t/4+8a, col 4k+t%%4): %d mismatches\n", bad);
  return 0;
}
