// scan_common.cuh -- device helpers shared by the scan kernels (scan.cu: one warp walks one query's probe
// sequence; scan_tail.cu: list-major FastScan over the tail of every query's probe sequence).
#pragma once
#include <cfloat>

#include "rbq_internal.h"

namespace rbq {

constexpr int kRefineSlots = 4;  // candidates refined per round (4 x 8 lanes)

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds_f32x2(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
    uint32_t v;
    asm volatile("{\n.reg .u16 t;\nld.shared.u16 t, [%1];\ncvt.u32.u16 %0, t;\n}" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t ldg32(const void* p) { return __ldg(reinterpret_cast<const uint32_t*>(p)); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ---- K10: packed ex-code dot, AVX2 lane order -------------------------------------------------------
// Staging: the 8 lanes of a group expand one candidate's packed ex-code (global memory) into one byte per
// code in shared memory, 16 bytes per 16-dim chunk ordered (c0,c8,c1,c9,...,c7,c15) so that "AVX lane" j
// later reads its two codes of the chunk (dims 16c+j and 16c+8+j) as one 16-bit load.
// EXK: 2 / 6 = the reference's C++-compatible layouts (src/simd.rs:2478-2695); 1 = generic LSB-first
// bit stream (src/simd.rs:166-191), used for bit widths the reference cannot search (extension).
template <int EXK>
__device__ __forceinline__ void stage_expand(const uint8_t* __restrict__ src, uint32_t stg, int D, int j, int ex_bits) {
#pragma unroll 4
    for (int c = j; c < D / 16; c += 8) {
        uint32_t A, Bq, Cq, Dq;  // codes 0-3, 4-7, 8-11, 12-15 of the chunk, one per byte
        if (EXK == 2) {
            const uint32_t w = ldg32(src + 4 * c);  // byte b: codes b, b+4, b+8, b+12 (2 bits each)
            A = w & 0x03030303u;
            Bq = (w >> 2) & 0x03030303u;
            Cq = (w >> 4) & 0x03030303u;
            Dq = (w >> 6) & 0x03030303u;
        } else if (EXK == 6) {
            const uint32_t w0 = ldg32(src + 12 * c), w1 = ldg32(src + 12 * c + 4), w2 = ldg32(src + 12 * c + 8);
            // bytes 0-7: low nibble = low 4 bits of code b, high nibble = low 4 bits of code b+8; w2: the 2-bit layout
            A = (w0 & 0x0F0F0F0Fu) | ((w2 << 4) & 0x30303030u);
            Bq = (w1 & 0x0F0F0F0Fu) | ((w2 << 2) & 0x30303030u);
            Cq = ((w0 >> 4) & 0x0F0F0F0Fu) | (w2 & 0x30303030u);
            Dq = ((w1 >> 4) & 0x0F0F0F0Fu) | ((w2 >> 2) & 0x30303030u);
        } else {
            const uint32_t mask = (1u << ex_bits) - 1u;
            uint32_t x[4] = {0u, 0u, 0u, 0u};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const uint32_t pos = (uint32_t)(16 * c + k) * (uint32_t)ex_bits;
                const uint32_t two = (uint32_t)__ldg(src + (pos >> 3)) | ((uint32_t)__ldg(src + (pos >> 3) + 1) << 8);
                x[k >> 2] |= ((two >> (pos & 7u)) & mask) << (8 * (k & 3));
            }
            A = x[0];
            Bq = x[1];
            Cq = x[2];
            Dq = x[3];
        }
        sts128(stg + 16u * (uint32_t)c, prmt(A, Cq, 0x5140u), prmt(A, Cq, 0x7362u), prmt(Bq, Dq, 0x5140u), prmt(Bq, Dq, 0x7362u));
    }
}
// One of the 8 "AVX lanes" (j): dims j, j+8, j+16, ... accumulated with fma, in the order of the two fmadd
// steps per 16 dims of the reference (src/simd.rs:1749-1757, 1804-1812).  rq2 holds the rotated query
// interleaved the same way: float2 (rq[16c+j], rq[16c+8+j]) at index 8c+j.
__device__ __forceinline__ float ex_dot_lane(uint32_t stg, uint32_t rq2, int D, int j) {
    float acc = 0.0f;
#pragma unroll 4
    for (int c = 0; c < D / 16; ++c) {
        const uint32_t pair = lds_u16(stg + 16u * (uint32_t)c + 2u * (uint32_t)j);
        const float2 qv = lds_f32x2(rq2 + 8u * (uint32_t)(8 * c + j));
        acc = __fmaf_rn((float)(pair & 0xffu), qv.x, acc);
        acc = __fmaf_rn((float)(pair >> 8), qv.y, acc);
    }
    return acc;
}
// horizontal sum of the 8 lanes exactly as the AVX2 code: ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7))
__device__ __forceinline__ float hsum8(float a) {
    a = a + __shfl_xor_sync(0xffffffffu, a, 4);
    a = a + __shfl_xor_sync(0xffffffffu, a, 2);
    a = a + __shfl_xor_sync(0xffffffffu, a, 1);
    return a;
}

// ---- K11: warp-cooperative insertion into an ascending list of at most k (distance, id) pairs -----
// Equal distances keep the earlier-visited entry first (and drop the newcomer at the boundary).
__device__ __forceinline__ void topk_insert(float* sd, unsigned long long* si, int& cnt, int k, float d,
                                            unsigned long long id, int lane) {
    int pos = 0;
    for (int base = 0; base < cnt; base += 32) {
        const int i = base + lane;
        pos += __popc(__ballot_sync(0xffffffffu, i < cnt && sd[i] <= d));
    }
    if (pos >= k) return;
    const int newcnt = cnt < k ? cnt + 1 : k;
    for (int base = ((newcnt - 1) >> 5) << 5; base >= 0 && base + 32 > pos; base -= 32) {
        const int i = base + lane;
        const bool mv = i >= pos && i < newcnt - 1;
        float td = 0.0f;
        unsigned long long ti = 0;
        if (mv) {
            td = sd[i];
            ti = si[i];
        }
        __syncwarp();
        if (mv) {
            sd[i + 1] = td;
            si[i + 1] = ti;
        }
        __syncwarp();
    }
    if (lane == 0) {
        sd[pos] = d;
        si[pos] = id;
    }
    __syncwarp();
    cnt = newcnt;
}

}  // namespace rbq
