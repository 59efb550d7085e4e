#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// microbench: mma.sync m16n8k32 u8 throughput + PRMT/IDP pipes on sm_100a
__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int NACC>
__global__ void k_imma(int* out, int iters, uint32_t seed) {
  int c[NACC][4];
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = 0;
  uint32_t a = 0x01010101u, b0 = seed + threadIdx.x, b1 = seed * 3 + threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++) imma(c[i], a, a, a, a, b0, b1);
  }
  int s = 0;
  for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_prmt(int* out, int iters, uint32_t seed) {
  uint32_t x[8];
  for (int i = 0; i < 8; i++) x[i] = seed * (i + 1) + threadIdx.x;
  uint32_t t0 = seed ^ 0x12345678u, t1 = seed ^ 0x9abcdef0u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = __byte_perm(t0, t1, x[i]);
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_idp(int* out, int iters, uint32_t seed) {
  int x[8];
  for (int i = 0; i < 8; i++) x[i] = seed * (i + 1) + threadIdx.x;
  uint32_t t0 = seed ^ 0x12345678u;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = (int)__dp4a(t0, 0x01000000u, (unsigned)x[i]);
  }
  int s = 0;
  for (int i = 0; i < 8; i++) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: 6 PRMT : 1 IMMA (the planned inner loop ratio)
__global__ void k_mix(int* out, int iters, uint32_t seed) {
  uint32_t x[12];
  for (int i = 0; i < 12; i++) x[i] = seed * (i + 1) + threadIdx.x;
  uint32_t t0 = seed ^ 0x12345678u, t1 = seed ^ 0x9abcdef0u;
  int c[2][4] = {};
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 12; i++) x[i] = __byte_perm(t0, t1, x[i]);
    imma(c[0], 0x01010101u, 0x01010101u, 0, 0, x[0], x[1]);
    imma(c[1], 0x01010101u, 0x01010101u, 0, 0, x[6], x[7]);
  }
  uint32_t s = 0;
  for (int i = 0; i < 12; i++) s += x[i];
  for (int j = 0; j < 4; j++) s += c[0][j] + c[1][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F>
float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  int* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
  int iters = 20000;
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock attr %d kHz\n", clk);
  for (int warps = 4; warps <= 16; warps *= 2) {
    int grid = 148, block = warps * 32;
    float ms;
    ms = timeit([&] { k_imma<4><<<grid, block>>>(out, iters, 7); });
    printf("warps/SM %2d  IMMA16832.u8 x4acc: %.3f ms -> %.2f warp-inst/ns/SM  (%.1f cyc/inst/SMSP @1.9GHz)\n", warps, ms,
           (double)iters * 4 * warps / (ms * 1e6), ms * 1e6 * 1.9 / ((double)iters * 4 * warps / 4));
    ms = timeit([&] { k_prmt<<<grid, block>>>(out, iters, 7); });
    printf("             PRMT x8: %.3f ms -> %.1f cyc/inst/SMSP\n", ms, ms * 1e6 * 1.9 / ((double)iters * 8 * warps / 4));
    ms = timeit([&] { k_idp<<<grid, block>>>(out, iters, 7); });
    printf("             IDP x8: %.3f ms -> %.1f cyc/inst/SMSP\n", ms, ms * 1e6 * 1.9 / ((double)iters * 8 * warps / 4));
    ms = timeit([&] { k_mix<<<grid, block>>>(out, iters, 7); });
    printf("             MIX 12 PRMT+2 IMMA: %.3f ms -> %.1f cyc/iter/SMSP\n", ms, ms * 1e6 * 1.9 / ((double)iters * warps / 4));
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
