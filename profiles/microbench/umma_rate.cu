// tcgen05.mma issue-rate microbenchmark (sm_100a): cycles per 128 x N x 32B MMA for kind::i8 / kind::f8f6f4, A from shared
// memory (SS) or tensor memory (TS), one or two CTAs per SM.  Operand contents are irrelevant (zeros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate umma_rate.cu && ./umma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc128(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// KIND 0: i8 (u8 x u8 -> s32), 1: f8f6f4 (e4m3 x e4m3 -> f32)
template <int KIND>
__device__ __forceinline__ uint32_t idesc(uint32_t n) {
    if (KIND == 0) return (2u << 4) | ((n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    return (1u << 4) | ((n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // D = F32 (1), A = B = E4M3 (0)
}

template <int KIND, int TS>
__global__ void __launch_bounds__(128) rate_kernel(int n, int iters, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_base;
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint64_t da = desc128(smem_u32(sm)), db = desc128(smem_u32(sm + 16384));
        const uint32_t id = idesc<KIND>((uint32_t)n);
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (TS) {
                    if (KIND == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem),
                                     "r"(tmem + 128u + 8u * k), "l"(db + (uint64_t)(2 * k)), "r"(id), "r"(1u)
                                     : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem),
                                     "r"(tmem + 128u + 8u * k), "l"(db + (uint64_t)(2 * k)), "r"(id), "r"(1u)
                                     : "memory");
                } else {
                    if (KIND == 0)
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                                     "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(id), "r"(1u)
                                     : "memory");
                    else
                        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n}" ::"r"(tmem),
                                     "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(id), "r"(1u)
                                     : "memory");
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile(
            "{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar))
            : "memory");
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

template <int KIND, int TS>
void run(const char* name, int n, int ctas_per_sm, long long* d_out) {
    const int iters = 4096, sms = 148;
    const size_t smem = 16384 + 32768 + 1024 + (ctas_per_sm == 1 ? 64 * 1024 : 0);  // pad so that only one CTA fits when asked
    cudaFuncSetAttribute(rate_kernel<KIND, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    rate_kernel<KIND, TS><<<sms * ctas_per_sm, 128, smem>>>(n, iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-10s N=%3d ctas/sm=%d : %7.1f clk per MMA (128xNx32B)%s\n", name, n, ctas_per_sm, (double)c / (iters * 4.0),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}


// i8 TS MMAs (N = 64) issued by warp 4 while warps 0-3 generate background traffic:
//   bg = 0 none, 1 tcgen05.st.x32 into other TMEM columns (the A ring being refilled), 2 st.shared.v4 into the B region,
//   3 both.  rot = 1 rotates the A columns (2 stages) and the B stage (8 x 8 KB) per group of 4 MMAs like the real kernel.
__global__ void __launch_bounds__(160) bg_kernel(int bg, int rot, int iters, long long* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint32_t tmem_base;
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 160) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
    if (tid == 0) {
        stop = 0;
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (warp == 4) {
        if ((tid & 31) == 0) {
            const uint32_t id = idesc<0>(64u);
            long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                const uint32_t ta = tmem + 128u + (rot ? (uint32_t)(it & 1) * 64u : 0u);
                const uint64_t db = desc128(smem_u32(sm + (rot ? (it & 7) * 8192 : 0)));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}" ::"r"(tmem),
                                 "r"(ta + 8u * k), "l"(db + (uint64_t)(2 * k)), "r"(id), "r"(1u)
                                 : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            asm volatile(
                "{\n.reg .pred P1;\nW2:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D2;\nbra W2;\nD2:\n}" ::"r"(smem_u32(&bar))
                : "memory");
            long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
            stop = 1;
        }
    } else {
        uint32_t x = tid;
        while (!stop) {
            if (bg & 1) {
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + 192u + (x & 32u);
                asm volatile(
                    "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                    "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr),
                    "r"(0u)
                    : "memory");
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            }
            if (bg & 2) {
                const uint32_t dst = smem_u32(sm + 32768) + (uint32_t)tid * 16u + ((x & 7u) * 2048u);
                asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0u) : "memory");
            }
            x = x * 1664525u + 1013904223u;
        }
    }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

void run_bg(int bg, int rot, int ctas_per_sm, long long* d_out) {
    const int iters = 4096, sms = 148;
    const size_t smem = 65536 + 1024 + (ctas_per_sm == 1 ? 48 * 1024 : 0);
    cudaFuncSetAttribute(bg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bg_kernel<<<sms * ctas_per_sm, 160, smem>>>(bg, rot, iters, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
    printf("i8 TS N=64 bg=%d rot=%d ctas/sm=%d : %7.1f clk per MMA %s\n", bg, rot, ctas_per_sm, (double)c / (iters * 4.0),
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char** argv) {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    for (int cps = 1; cps <= 2; ++cps)
        for (int rot = 0; rot <= 1; ++rot)
            for (int bg = 0; bg <= 3; ++bg) run_bg(bg, rot, cps, d_out);
    if (argc < 2) return 0;
    for (int cps = 1; cps <= 2; ++cps)
        for (int n : {16, 32, 48, 64, 128, 256}) {
            run<0, 0>("i8 SS", n, cps, d_out);
            run<0, 1>("i8 TS", n, cps, d_out);
            run<1, 0>("e4m3 SS", n, cps, d_out);
            run<1, 1>("e4m3 TS", n, cps, d_out);
        }
    return 0;
}
