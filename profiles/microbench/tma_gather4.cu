// TMA tile::gather4 probe (sm_100a): which tensor-map box shape does cp.async.bulk.tensor.2d...tile::gather4 want, and where do
// the four gathered rows land under 128-byte swizzle?  Prints the verdict for box = {128, 1} and {128, 4}.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_gather4 tma_gather4.cu -lcuda && ./tma_gather4
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void gather_kernel(const __grid_constant__ CUtensorMap map, int col, int r0, int r1, int r2, int r3, uint8_t* out) {
    extern __shared__ unsigned char raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 2048; i += blockDim.x) sm[i] = 0xEE;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(512) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                smem_u32(sm + 512)),  // second 4-row group of an 8-row swizzle atom: rows 4..7
            "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
            : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = tid; i < 2048; i += blockDim.x) out[i] = sm[i];
}

int main() {
    const int rows = 64, rowbytes = 3840;
    uint8_t* h = (uint8_t*)malloc((size_t)rows * rowbytes);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < rowbytes; ++c) h[(size_t)r * rowbytes + c] = (uint8_t)((r * 16) ^ (c / 16));  // byte = f(row, 16-byte chunk)
    uint8_t *d, *dout;
    cudaMalloc(&d, (size_t)rows * rowbytes);
    cudaMalloc(&dout, 2048);
    cudaMemcpy(d, h, (size_t)rows * rowbytes, cudaMemcpyHostToDevice);
    for (int boxrows : {1, 4}) {
        CUtensorMap map;
        cuuint64_t dims[2] = {(cuuint64_t)rowbytes, (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)rowbytes};
        cuuint32_t box[2] = {128, (cuuint32_t)boxrows};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box rows %d: encode rc=%d\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        const int col = 256, rr[4] = {3, 10, 7, 0};
        cudaMemset(dout, 0, 2048);
        gather_kernel<<<1, 128, 4096>>>(map, col, rr[0], rr[1], rr[2], rr[3], dout);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  kernel: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        uint8_t o[2048];
        cudaMemcpy(o, dout, 2048, cudaMemcpyDeviceToHost);
        // expectation: gathered row i lands in smem row 4 + i (128 B each), 16-byte chunk j at position j ^ ((4 + i) & 7)
        int ok = 1;
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 8; ++j) {
                const int pos = j ^ ((4 + i) & 7);
                const uint8_t want = (uint8_t)((rr[i] * 16) ^ ((col + 16 * j) / 16));
                for (int b = 0; b < 16; ++b)
                    if (o[512 + i * 128 + pos * 16 + b] != want) ok = 0;
            }
        int untouched = 1;
        for (int i = 0; i < 512; ++i)
            if (o[i] != 0xEE) untouched = 0;
        printf("  rows land swizzled at rows 4..7 as expected: %s; rows 0..3 untouched: %s\n", ok ? "YES" : "no", untouched ? "yes" : "NO");
        if (!ok) {
            for (int i = 0; i < 4; ++i) {
                printf("  smem row %d:", 4 + i);
                for (int j = 0; j < 8; ++j) printf(" %02x", o[512 + i * 128 + j * 16]);
                printf("\n");
            }
        }
    }
    return 0;
}
