// Pipe-rate microbenchmarks (sm_100a): which instructions can carry the FastScan lookup loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("prmt.b32 %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("lop3.b32 %0,%1,%2,%3,0xCA;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t shr16(uint32_t a) { uint32_t d; asm volatile("shr.u32 %0,%1,16;" : "=r"(d) : "r"(a)); return d; }
__device__ __forceinline__ uint32_t mulhi(uint32_t a, uint32_t b) { uint32_t d; asm volatile("mul.hi.u32 %0,%1,%2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t mad(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.u32 %0,%1,%2,%3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
__device__ __forceinline__ uint32_t add3(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.u32 %0,%1,%2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}
#define KERNEL(name, body)                                                           \
  __global__ void name(int* out, int iters, uint32_t seed) {                         \
    uint32_t x[8];                                                                   \
    for (int i = 0; i < 8; i++) x[i] = seed * (i + 1) + threadIdx.x;                 \
    uint32_t t0 = seed ^ 0x12345678u, t1 = seed ^ 0x9abcdef0u;                       \
    for (int it = 0; it < iters; it++) {                                             \
      _Pragma("unroll") for (int i = 0; i < 8; i++) { body; }                        \
    }                                                                                \
    uint32_t s = 0;                                                                  \
    for (int i = 0; i < 8; i++) s += x[i];                                           \
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + t0 + t1;                        \
  }
KERNEL(k_prmt, x[i] = prmt(t0, t1, x[i]))
KERNEL(k_prmt_imm, x[i] = prmt(x[i], t1, 0x5410u))
KERNEL(k_lop3, x[i] = lop3(x[i], t0, t1))
KERNEL(k_shr, x[i] = shr16(x[i]) + 0 * t0; x[i] |= seed)
KERNEL(k_mulhi, x[i] = mulhi(x[i], t0))
KERNEL(k_mad, x[i] = mad(x[i], t0, t1))
KERNEL(k_add, x[i] = add3(x[i], t0))
// the planned inner step for 8 lookups: 4 PRMT + 2 LOP3 + 2 shifts (ALU or FMA pipe) + 1 IMMA
template <int SHIFT_FMA>
__global__ void k_step(int* out, int iters, uint32_t seed, const uint4* __restrict__ g) {
  uint32_t t0 = seed ^ 0x12345678u, t1 = seed ^ 0x9abcdef0u, t2 = seed * 77u, t3 = seed * 91u;
  uint4 s4 = g[threadIdx.x], m4 = g[threadIdx.x + 256], m5 = g[threadIdx.x + 512];
  int c[4][4] = {};
  for (int it = 0; it < iters; it++) {
    uint32_t S[4] = {s4.x + it, s4.y + it, s4.z, s4.w}, M[8] = {m4.x, m4.y, m4.z, m4.w, m5.x, m5.y, m5.z, m5.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t s = S[k], sh = SHIFT_FMA ? mulhi(s, 0x10000u) : shr16(s);
      uint32_t lo0 = prmt(t0, t1, s), hi0 = prmt(t2, t3, s), lo1 = prmt(t0, t1, sh), hi1 = prmt(t2, t3, sh);
      uint32_t r0 = lop3(M[2 * k], hi0, lo0), r1 = lop3(M[2 * k + 1], hi1, lo1);
      imma(c[k], 1u, 1u, r0, r1);
    }
  }
  int s = 0;
  for (int k = 0; k < 4; k++) for (int j = 0; j < 4; j++) s += c[k][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  int* out; cudaMalloc(&out, 148 * 1024 * 4 * 2);
  uint4* g; cudaMalloc(&g, 1024 * 16); cudaMemset(g, 0x35, 1024 * 16);
  const int iters = 20000, grid = 148;
  const double ghz = 1.965;
  for (int warps = 8; warps <= 16; warps *= 2) {
    int block = warps * 32;
    printf("== %d warps/SM ==\n", warps);
#define RUN(name) { float ms = timeit([&] { name<<<grid, block>>>(out, iters, 7); }); \
    printf("%-12s %.3f ms -> %.2f cyc/inst/SMSP\n", #name, ms, ms * 1e6 * ghz / ((double)iters * 8 * warps / 4)); }
    RUN(k_prmt) RUN(k_prmt_imm) RUN(k_lop3) RUN(k_shr) RUN(k_mulhi) RUN(k_mad) RUN(k_add)
    { float ms = timeit([&] { k_step<0><<<grid, block>>>(out, iters, 7, g); });
      printf("step(shr)    %.3f ms -> %.1f cyc per 8 lookups /SMSP\n", ms, ms * 1e6 * ghz / ((double)iters * 4 * warps / 4)); }
    { float ms = timeit([&] { k_step<1><<<grid, block>>>(out, iters, 7, g); });
      printf("step(mulhi)  %.3f ms -> %.1f cyc per 8 lookups /SMSP\n", ms, ms * 1e6 * ghz / ((double)iters * 4 * warps / 4)); }
  }
  printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
