// mbarrier ping-pong latency (sm_100a): P producer warps and one consumer warp exchange full/empty barriers over a 2-stage
// ring with no work in between -- the synchronisation skeleton of tail_tc_kernel.  Variants:
//   wait_all   1: every lane polls the barrier, 0: lane 0 polls, then __syncwarp
//   arrive_all 1: every producer thread arrives (count 32 P), 0: one lane per warp (count P)
//   poll       0: mbarrier.try_wait loop, 1: mbarrier.test_wait loop (pure spin)
//   extra      bit mask of per-chunk producer extras: 1 cp.async commit/wait_group, 2 tcgen05 fences, 4 fence.proxy.async, 8 __syncwarp
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mbar_pingpong mbar_pingpong.cu && ./mbar_pingpong
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int POLL>
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    if (POLL == 0)
        asm volatile("{\n.reg .pred P1;\nW%=:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D%=;\nbra W%=;\nD%=:\n}" ::"r"(bar), "r"(parity)
                     : "memory");
    else
        asm volatile("{\n.reg .pred P1;\nW%=:\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D%=;\nbra W%=;\nD%=:\n}" ::"r"(bar), "r"(parity)
                     : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

template <int WAIT_ALL, int ARRIVE_ALL, int POLL, int STAGES>
__global__ void pingpong(int P, int iters, long long* out, int extra) {
    __shared__ uint64_t bars[2 * 8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[8]);
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(full0 + 8 * s), "r"(ARRIVE_ALL ? 32 * P : P) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty0 + 8 * s) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const long long t0 = clock64();
    if (warp == P) {  // consumer
        for (int c = 0; c < iters; ++c) {
            const uint32_t s = c % STAGES;
            if (WAIT_ALL || lane == 0) bar_wait<POLL>(full0 + 8 * s, (c / STAGES) & 1);
            if (extra & 2) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(empty0 + 8 * s);
            __syncwarp();
        }
    } else {
        for (int c = 0; c < iters; ++c) {
            const uint32_t s = c % STAGES, use = c / STAGES;
            if (use > 0) {
                if (WAIT_ALL || lane == 0) bar_wait<POLL>(empty0 + 8 * s, (use - 1) & 1);
                __syncwarp();
            }
            if (extra & 1) {
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 2;" ::: "memory");
            }
            if (extra & 8) __syncwarp();
            if (extra & 2) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            if (extra & 4) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (ARRIVE_ALL) bar_arrive(full0 + 8 * s);
            else {
                __syncwarp();
                if (lane == 0) bar_arrive(full0 + 8 * s);
            }
        }
    }
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) out[0] = clock64() - t0;
}

template <int WAIT_ALL, int ARRIVE_ALL, int POLL, int STAGES>
void run(int P, int ctas, long long* d_out, int extra = 0) {
    const int iters = 20000;
    pingpong<WAIT_ALL, ARRIVE_ALL, POLL, STAGES><<<148 * ctas, (P + 1) * 32>>>(P, iters, d_out, extra);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("extra=%2d P=%d ctas/sm=%d stages=%d wait_all=%d arrive_all=%d poll=%s : %7.1f clk per chunk %s\n", extra, P, ctas, STAGES, WAIT_ALL, ARRIVE_ALL,
           POLL ? "test_wait" : "try_wait", (double)c / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    for (int ctas = 1; ctas <= 2; ++ctas) {
        run<1, 1, 0, 2>(8, ctas, d_out);
        run<1, 0, 0, 2>(8, ctas, d_out);
        run<0, 0, 0, 2>(8, ctas, d_out);
        run<1, 1, 1, 2>(8, ctas, d_out);
        run<0, 0, 1, 2>(8, ctas, d_out);
        run<1, 0, 0, 4>(8, ctas, d_out);
        run<0, 0, 0, 4>(8, ctas, d_out);
        run<0, 0, 1, 4>(8, ctas, d_out);
        run<0, 0, 0, 2>(4, ctas, d_out);
        for (int extra : {1, 2, 4, 8, 15}) run<1, 1, 0, 2>(8, ctas, d_out, extra);
    }
    return 0;
}
