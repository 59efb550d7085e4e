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
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar),
        "r"(parity), "r"(kMbarSuspendHintNs)
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- K7: lookups ---------------------------------------------------------------------------------
// 32 nibble lookups of one codebook.  C = its 16 code bytes, T = its 16 LUT bytes.  acc[v] += lut[nibble(v)].
// Byte j of C holds vector KPERM0[j] (low nibble) and KPERM0[j]+16 (high nibble); bytes 4k+2h, 4k+2h+1
// form PRMT selector half h of register k and carry vectors m, m+16, m+8, m+24 with m = 2k+h.
__device__ __forceinline__ void lookup_accumulate(const uint4& Cv, const uint4& T, uint32_t (&acc)[32]) {
    const uint32_t C[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = C[k];
        const uint32_t s = c & 0x77777777u;  // 3-bit byte selectors (bit 3 of a PRMT selector = sign-replicate mode)
        const uint32_t sh = c << 4;          // brings the low nibbles' msb into byte-sign position
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t sel = h ? (s >> 16) : s;
            const uint32_t lo = prmt(T.x, T.y, sel);                // entries 0..7
            const uint32_t hi = prmt(T.z, T.w, sel);                // entries 8..15
            const uint32_t m = prmt(c, sh, h ? 0xBFAEu : 0x9D8Cu);  // 0xFF where the nibble's msb is set
            const uint32_t r = (lo & ~m) | (hi & m);                // 4 looked-up bytes
            const int v = 2 * k + h;
            acc[v] = __dp4a(r, 0x00000001u, acc[v]);
            acc[v + 16] = __dp4a(r, 0x00000100u, acc[v + 16]);
            acc[v + 8] = __dp4a(r, 0x00010000u, acc[v + 8]);
            acc[v + 24] = __dp4a(r, 0x01000000u, acc[v + 24]);
        }
    }
}

// Cross-lane reduce-scatter: on return lane v holds the sum over lanes of acc[v].
template <bool WIDE>
__device__ __forceinline__ uint32_t reduce_scatter(uint32_t (&acc)[32], int lane) {
    const unsigned full = 0xffffffffu;
    if (!WIDE) {
        // totals fit 16 bits (padded_dim <= 1024): pack vector pairs (v, v+8) and halve the shuffles
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
        uint32_t X[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const uint32_t E = acc[m] + (acc[m + 8] << 16), O = acc[m + 16] + (acc[m + 24] << 16);
            const uint32_t keep = b4 ? O : E, send = b4 ? E : O;
            X[m] = keep + __shfl_xor_sync(full, send, 16);
        }
        uint32_t Y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t keep = b2 ? X[i + 4] : X[i], send = b2 ? X[i] : X[i + 4];
            Y[i] = keep + __shfl_xor_sync(full, send, 4);
        }
        uint32_t Z[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t keep = b1 ? Y[i + 2] : Y[i], send = b1 ? Y[i] : Y[i + 2];
            Z[i] = keep + __shfl_xor_sync(full, send, 2);
        }
        const uint32_t W = (b0 ? Z[1] : Z[0]) + __shfl_xor_sync(full, b0 ? Z[0] : Z[1], 1);
        const uint32_t keep = b3 ? (W >> 16) : (W & 0xffffu), send = b3 ? (W & 0xffffu) : (W >> 16);
        return keep + __shfl_xor_sync(full, send, 8);
    } else {
        // padded_dim > 1024: totals can exceed 16 bits, reduce in 32-bit (the caller applies the u16 wrap)
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
            const bool up = lane & w;
#pragma unroll
            for (int i = 0; i < w; ++i) {
                const uint32_t keep = up ? acc[i + w] : acc[i], send = up ? acc[i] : acc[i + w];
                acc[i] = keep + __shfl_xor_sync(full, send, w);
            }
        }
        return acc[0];
    }
}

// K7 for one block resident in shared memory (blk = 32-bit shared address): returns accu[lane] (exact
// integer sum, before the u16 wrap).  Branch-free: lanes past the last codebook hold an all-zero LUT row, so
// whatever (clamped, valid) code bytes they read contribute nothing.
template <int NCB, bool WIDE>
__device__ __forceinline__ uint32_t accumulate_block(uint32_t blk, const uint4 (&T)[NCB], int ncb, int lane) {
    uint32_t acc[32];
#pragma unroll
    for (int v = 0; v < 32; ++v) acc[v] = 0u;
#pragma unroll
    for (int i = 0; i < NCB; ++i) {
        const int cb = min(lane + 32 * i, ncb - 1);
        const uint4 C = lds128(blk + 16u * (uint32_t)cb);
        lookup_accumulate(C, T[i], acc);
    }
    return reduce_scatter<WIDE>(acc, lane);
}

// The same for a block read straight from global memory (head scan: every block is visited once per query, so
// there is nothing to stage).
template <int NCB>
__device__ __forceinline__ void load_block_codes(const uint8_t* __restrict__ blk, uint4 (&C)[NCB], int ncb, int lane) {
#pragma unroll
    for (int i = 0; i < NCB; ++i) C[i] = ldg128(blk + 16 * min(lane + 32 * i, ncb - 1));
}
template <int NCB, bool WIDE>
__device__ __forceinline__ uint32_t accumulate_block_regs(const uint4 (&C)[NCB], const uint4 (&T)[NCB], int lane) {
    uint32_t acc[32];
#pragma unroll
    for (int v = 0; v < 32; ++v) acc[v] = 0u;
#pragma unroll
    for (int i = 0; i < NCB; ++i) lookup_accumulate(C[i], T[i], acc);
    return reduce_scatter<WIDE>(acc, lane);
}

// ---- K10: packed ex-code dot, AVX2 lane order -------------------------------------------------------
// Staging: the 8 lanes of a group expand one candidate's packed ex-code (global memory) into one byte per
// code in shared memory, 16 bytes per 16-dim chunk ordered (c0,c8,c1,c9,...,c7,c15) so that "AVX lane" j
// later reads its two codes of the chunk (dims 16c+j and 16c+8+j) as one 16-bit load.
// EXK: 2 / 6 = the reference's C++-compatible layouts (src/simd.rs:2478-2695); 1 = generic LSB-first
// bit stream (src/simd.rs:166-191), used for bit widths the reference cannot search (extension).
// SMEM: src is a generic pointer into shared memory (raw packed code copied there by cp.async) instead of global.
// the 16 codes of chunk c of one packed ex-code, one per byte: A = codes 0-3, Bq = 4-7, Cq = 8-11, Dq = 12-15
template <int EXK, bool SMEM = false>
__device__ __forceinline__ void decode_chunk(const uint8_t* __restrict__ src, int c, int ex_bits, uint32_t& A, uint32_t& Bq, uint32_t& Cq,
                                             uint32_t& Dq) {
    auto ldw = [&](const uint8_t* p) -> uint32_t { return SMEM ? *reinterpret_cast<const uint32_t*>(p) : ldg32(p); };
    auto ldb = [&](const uint8_t* p) -> uint32_t { return SMEM ? (uint32_t)*p : (uint32_t)__ldg(p); };
    if (EXK == 2) {
        const uint32_t w = ldw(src + 4 * c);  // byte b: codes b, b+4, b+8, b+12 (2 bits each)
        A = w & 0x03030303u;
        Bq = (w >> 2) & 0x03030303u;
        Cq = (w >> 4) & 0x03030303u;
        Dq = (w >> 6) & 0x03030303u;
    } else if (EXK == 6) {
        const uint32_t w0 = ldw(src + 12 * c), w1 = ldw(src + 12 * c + 4), w2 = ldw(src + 12 * c + 8);
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
            const uint32_t two = ldb(src + (pos >> 3)) | (ldb(src + (pos >> 3) + 1) << 8);
            x[k >> 2] |= ((two >> (pos & 7u)) & mask) << (8 * (k & 3));
        }
        A = x[0];
        Bq = x[1];
        Cq = x[2];
        Dq = x[3];
    }
}
template <int EXK, bool SMEM = false>
__device__ __forceinline__ void stage_expand(const uint8_t* __restrict__ src, uint32_t stg, int D, int j, int ex_bits) {
#pragma unroll 4
    for (int c = j; c < D / 16; c += 8) {
        uint32_t A, Bq, Cq, Dq;  // codes 0-3, 4-7, 8-11, 12-15 of the chunk, one per byte
        decode_chunk<EXK, SMEM>(src, c, ex_bits, A, Bq, Cq, Dq);
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
        // byte -> float without the conversion pipe: 0x4B0000xx is 2^23 + xx exactly, and the subtraction is exact
        const float ca = __uint_as_float(prmt(pair, 0x4B000000u, 0x7540u)) - 8388608.0f;
        const float cb = __uint_as_float(prmt(pair, 0x4B000000u, 0x7541u)) - 8388608.0f;
        acc = __fmaf_rn(ca, qv.x, acc);
        acc = __fmaf_rn(cb, qv.y, acc);
    }
    return acc;
}
// ---- K10 on the device's lane-major ex-code layout (DevIndex::exl, built at load time by resolve.cu) ----------
// The reference's ex-dot runs 8 independent FMA chains ("AVX lane" j owns dims j, j+8, j+16, ...; src/simd.rs:1749-1757,
// 1804-1812).  exl stores every vector as 8 rows of one byte per code in exactly that order -- row j, position t =
// code of dim 8t + j -- zero padded to `exl_lane` bytes, so a chain reads its own contiguous row and nothing has to be
// unpacked or exchanged between lanes at search time.  The rotated query is staged the same way (row j, position t =
// q[8t + j], zero padded): padding adds fma(0, 0, acc) = acc (acc is never -0: it starts at +0 and x + (-x) = +0).
__host__ __device__ inline uint32_t exl_lane_bytes(uint32_t D) { return ((D / 8u) + 15u) / 16u * 16u; }
// shared-memory row strides (bytes) that keep 128-bit loads of 8 consecutive rows on disjoint banks: 16 * odd
__host__ __device__ inline uint32_t exl_row_stride(uint32_t D) { return ((exl_lane_bytes(D) / 16u) | 1u) * 16u; }
__host__ __device__ inline uint32_t rql_row_stride(uint32_t D) { return ((exl_lane_bytes(D) * 4u / 16u) | 1u) * 16u; }

// one chain: `row` = shared address of the lane's code bytes, `qrow` = shared address of its query floats
__device__ __forceinline__ float ex_dot_chain(uint32_t row, uint32_t qrow, uint32_t lane_bytes) {
    float acc = 0.0f;
#pragma unroll 1
    for (uint32_t t = 0; t < lane_bytes; t += 16) {
        const uint4 w = lds128(row + t);
        const uint32_t W[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 qv = lds128(qrow + 4u * t + 16u * (uint32_t)k);
            // byte -> float without the conversion pipe: 0x4B0000xx is 2^23 + xx exactly, and the subtraction is exact
            const float c0 = __uint_as_float(prmt(W[k], 0x4B000000u, 0x7540u)) - 8388608.0f;
            const float c1 = __uint_as_float(prmt(W[k], 0x4B000000u, 0x7541u)) - 8388608.0f;
            const float c2 = __uint_as_float(prmt(W[k], 0x4B000000u, 0x7542u)) - 8388608.0f;
            const float c3 = __uint_as_float(prmt(W[k], 0x4B000000u, 0x7543u)) - 8388608.0f;
            acc = __fmaf_rn(c0, __uint_as_float(qv.x), acc);
            acc = __fmaf_rn(c1, __uint_as_float(qv.y), acc);
            acc = __fmaf_rn(c2, __uint_as_float(qv.z), acc);
            acc = __fmaf_rn(c3, __uint_as_float(qv.w), acc);
        }
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

// ---- eight-lane refine (one chain per lane, 4 candidates per round): used where the paired form's larger staging would cost
// resident warps (padded_dim > 640)
// rotated query in chain order: row j (stride rql_row bytes), position t = q[8t + j]; positions past D/8 are zero
__device__ __forceinline__ void load_rql(unsigned char* rql, uint32_t rql_row, const float* __restrict__ rot, int D, uint32_t lane_bytes,
                                         int lane) {
    for (int i = lane; i < 8 * (int)lane_bytes; i += 32) {
        const int j = i & 7, t = i >> 3;
        reinterpret_cast<float*>(rql + (size_t)j * rql_row)[t] = i < D ? __ldg(rot + i) : 0.0f;
    }
}

// K10 for up to 32 candidates: lane i holds the global vector index of candidate i (i < nb); returns that candidate's
// ex-code dot in lane i.  Candidates are served 4 at a time by the four 8-lane groups (= the 8 AVX lanes): every lane
// copies ITS chain's code row (DevIndex::exl) into its own shared-memory row with cp.async and multiplies it against
// its query row -- no unpacking and no exchange between lanes.  With two staging buffers (stage_bufs == 2) the rows of
// round r+1 travel while round r is multiplied; the default is one buffer (the candidates' rows were already pulled into
// L2 when they were queued, and the smaller footprint lets more warps be resident, which measured faster).
__device__ __forceinline__ float refine_batch(const DevIndex& ix, unsigned long long gv, int nb, uint32_t stage_u32, uint32_t rql_u32,
                                              uint32_t exl_row, uint32_t rql_row, uint32_t stage_bufs, int lane) {
    const int g = lane >> 3, j = lane & 7;
    const uint32_t LB = ix.exl_lane;
    const uint32_t my_row = (uint32_t)lane * exl_row, buf_bytes = 32u * exl_row;
    const uint32_t qrow = rql_u32 + (uint32_t)j * rql_row;
    const uint32_t npr = LB >> 4;  // 16-byte pieces per code row
    auto issue = [&](int r0, uint32_t buf) {
        const int c = r0 + g;
        const unsigned long long gv_c = __shfl_sync(0xffffffffu, gv, c & 31);
        if (c < nb) {
            if (ix.exl_copy_coalesced) {
                // The candidate's 8 rows are one contiguous block of 8 * LB bytes: the group's 8 lanes copy it in address order,
                // 128 contiguous bytes per warp instruction and candidate (4 lines instead of 32 half-used sectors when every lane
                // walks its own row).  A piece lands in the staging row of the chain it belongs to, where that chain's lane reads it.
                const uint8_t* src = ix.exl + gv_c * ix.exl_stride + 16u * (uint32_t)j;  // this lane's piece of stripe 0
                const uint32_t dst = stage_u32 + buf * buf_bytes + (uint32_t)(g * 8) * exl_row;
#ifdef RBQ_EXL_COPY_REGULAR  // cheaper address arithmetic for the regular row sizes.  Built with resolve_lazy_kernel capped at 80 registers (the extra
                             // branches otherwise raise it to 118) it measured no better than the one loop below (GIST-1M head 0.401 vs 0.405 ms,
                             // replay 0.235 vs 0.223), so it stays off; verified bit-exact (157 GPU tests) for the record.
                if ((LB & 127u) == 0u) {  // every 128-byte stripe lies inside one row (padded_dim % 1024 == 0 or 960, ...)
                    // (kept rolled: unrolling the 8 rows raised resolve_lazy_kernel from 80 to 121 registers and cost two resident CTAs per SM)
                    uint32_t d = dst + 16u * (uint32_t)j;
#pragma unroll 1
                    for (int r = 0; r < 8; ++r) {
#pragma unroll 1
                        for (uint32_t o = 0; o < LB; o += 128u) cp_async16(d + o, src + o);
                        d += exl_row;
                        src += LB;
                    }
                } else if (LB == 64u || LB == 32u) {  // a stripe covers 2 or 4 whole rows
                    const uint32_t sh = LB == 64u ? 2u : 1u;  // 16-byte pieces per row = 1 << sh
                    uint32_t d = dst + ((uint32_t)j >> sh) * exl_row + (((uint32_t)j & ((1u << sh) - 1u)) << 4);
                    const uint32_t dstep = (8u >> sh) * exl_row;
                    for (uint32_t o = 0; o < 8u * LB; o += 128u) {
                        cp_async16(d, src + o);
                        d += dstep;
                    }
                } else
#endif
                {  // piece f = j, j + 8, ... sits in row f / npr (any row size)
                    uint32_t row = (uint32_t)j / npr, pc = (uint32_t)j - row * npr;
                    for (uint32_t f = (uint32_t)j; f < 8u * npr; f += 8u) {
                        cp_async16(dst + row * exl_row + (pc << 4), src + ((size_t)(f - (uint32_t)j) << 4));
                        pc += 8u;
                        while (pc >= npr) {
                            pc -= npr;
                            row += 1u;
                        }
                    }
                }
            } else {
                const uint8_t* src = ix.exl + gv_c * ix.exl_stride + (size_t)j * LB;
                const uint32_t dst = stage_u32 + buf * buf_bytes + my_row;
                for (uint32_t p = 0; p < LB; p += 16) cp_async16(dst + p, src + p);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const bool dbl = stage_bufs > 1u;
    issue(0, 0u);
    float exdot = 0.0f;
    uint32_t buf = 0;
    for (int r0 = 0; r0 < nb; r0 += kRefineSlots) {
        const int c = r0 + g;  // candidate served by this 8-lane group
        if (dbl && r0 + kRefineSlots < nb) {
            issue(r0 + kRefineSlots, buf ^ 1u);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (ix.exl_copy_coalesced) __syncwarp();  // a lane's row was copied by the other lanes of its group
        float part = 0.0f;
        if (c < nb) part = ex_dot_chain(stage_u32 + buf * buf_bytes + my_row, qrow, LB);
        part = hsum8(part);  // (full-mask shuffles: every lane is past its reads of the staging rows when the next copies are issued)
        const float v = __shfl_sync(0xffffffffu, part, ((lane - r0) & 3) * 8);
        if (lane >= r0 && lane < r0 + kRefineSlots) exdot = v;
        if (dbl) buf ^= 1u;
        else if (r0 + kRefineSlots < nb) issue(r0 + kRefineSlots, 0u);  // the lane's own row is free again (only it reads it)
    }
    return exdot;
}


// ---- paired chains: one GPU lane runs TWO of the 8 AVX lanes (j and j + 4) with the packed FP32 instructions of sm_100
// (add.rn.f32x2 / fma.rn.f32x2 = FADD2 / FFMA2: two IEEE operations per issue slot, each element rounded exactly like the scalar
// instruction).  A candidate then needs 4 lanes instead of 8, a warp refines 8 candidates per round instead of 4, and the
// byte->float subtraction and the fma cost one instruction per TWO codes: the refine loops are issue-bound, so this is where
// their time goes.  The first step of the AVX2 horizontal sum (a_j + a_{j+4}) becomes an in-lane addition.
constexpr int kRefineSlots2 = 8;  // candidates refined per round (8 x 4 lanes)
// shared-memory strides (bytes): a lane's two staged code rows (row p at +0, row p + 4 at +exl_lane_bytes), and a pair-row of
// the query (float2 (q[8t + p], q[8t + p + 4]) at index t); both 16 * odd, so 128-bit loads of consecutive lanes / the four
// pair-rows fall on disjoint banks
__host__ __device__ inline uint32_t exl2_lane_stride(uint32_t D) { return ((2u * exl_lane_bytes(D) / 16u) | 1u) * 16u; }
__host__ __device__ inline uint32_t rql2_row_stride(uint32_t D) { return ((exl_lane_bytes(D) * 8u / 16u) | 1u) * 16u; }

__device__ __forceinline__ unsigned long long f32x2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ unsigned long long f32x2_add(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// rotated query in paired chain order: pair-row p (stride rql_row bytes), float2 t = (q[8t + p], q[8t + p + 4]); zero past D
__device__ __forceinline__ void load_rql2(unsigned char* rql, uint32_t rql_row, const float* __restrict__ rot, int D, uint32_t lane_bytes,
                                          int lane) {
    for (int i = lane; i < 8 * (int)lane_bytes; i += 32) {
        const int j = i & 7, t = i >> 3;
        reinterpret_cast<float*>(rql + (size_t)(j & 3) * rql_row)[2 * t + (j >> 2)] = i < D ? __ldg(rot + i) : 0.0f;
    }
}
// two chains: row0 / row1 = shared addresses of the code bytes of chains p and p + 4, qrow = the pair-row; returns a_p + a_{p+4}
__device__ __forceinline__ float ex_dot_chain2(uint32_t row0, uint32_t row1, uint32_t qrow, uint32_t lane_bytes) {
    unsigned long long acc = 0ull;  // (+0.0f, +0.0f)
    const unsigned long long m23 = f32x2_pack(-8388608.0f, -8388608.0f);
#pragma unroll 1
    for (uint32_t t = 0; t < lane_bytes; t += 16) {
        const uint4 w0 = lds128(row0 + t), w1 = lds128(row1 + t);
        const uint32_t W0[4] = {w0.x, w0.y, w0.z, w0.w}, W1[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 qa = lds128(qrow + 8u * t + 32u * (uint32_t)k), qb = lds128(qrow + 8u * t + 32u * (uint32_t)k + 16u);
            // byte -> float without the conversion pipe: 0x4B0000xx is 2^23 + xx exactly, and the subtraction is exact
            const unsigned long long c0 = f32x2_add(f32x2_pack(__uint_as_float(prmt(W0[k], 0x4B000000u, 0x7540u)), __uint_as_float(prmt(W1[k], 0x4B000000u, 0x7540u))), m23);
            const unsigned long long c1 = f32x2_add(f32x2_pack(__uint_as_float(prmt(W0[k], 0x4B000000u, 0x7541u)), __uint_as_float(prmt(W1[k], 0x4B000000u, 0x7541u))), m23);
            const unsigned long long c2 = f32x2_add(f32x2_pack(__uint_as_float(prmt(W0[k], 0x4B000000u, 0x7542u)), __uint_as_float(prmt(W1[k], 0x4B000000u, 0x7542u))), m23);
            const unsigned long long c3 = f32x2_add(f32x2_pack(__uint_as_float(prmt(W0[k], 0x4B000000u, 0x7543u)), __uint_as_float(prmt(W1[k], 0x4B000000u, 0x7543u))), m23);
            acc = f32x2_fma(c0, f32x2_pack(__uint_as_float(qa.x), __uint_as_float(qa.y)), acc);
            acc = f32x2_fma(c1, f32x2_pack(__uint_as_float(qa.z), __uint_as_float(qa.w)), acc);
            acc = f32x2_fma(c2, f32x2_pack(__uint_as_float(qb.x), __uint_as_float(qb.y)), acc);
            acc = f32x2_fma(c3, f32x2_pack(__uint_as_float(qb.z), __uint_as_float(qb.w)), acc);
        }
    }
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
    return lo + hi;  // a_p + a_{p+4}: the first step of the AVX2 horizontal sum
}
// K10 for up to 32 candidates: lane i holds the global vector index of candidate i (i < nb); returns that candidate's ex-code
// dot in lane i.  Candidates are served 8 at a time by the eight 4-lane groups: every lane copies ITS two chains' code rows
// (DevIndex::exl rows p and p + 4) into its own shared-memory region with cp.async and multiplies them against its query
// pair-row -- no unpacking and no exchange between lanes; the remaining two steps of the AVX2 horizontal sum
// ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7)) are two shuffles.  stage_bufs == 2: the rows of round r+1 travel while round r is multiplied.
__device__ __forceinline__ float refine_batch2(const DevIndex& ix, unsigned long long gv, int nb, uint32_t stage_u32, uint32_t rql_u32,
                                               uint32_t lane_stride, uint32_t rql_row, uint32_t stage_bufs, int lane) {
    const int g = lane >> 2, p = lane & 3;
    const uint32_t LB = ix.exl_lane;
    const uint32_t my = (uint32_t)lane * lane_stride, buf_bytes = 32u * lane_stride;
    const uint32_t qrow = rql_u32 + (uint32_t)p * rql_row;
    const uint32_t npr = LB >> 4;  // 16-byte pieces per code row
    auto issue = [&](int r0, uint32_t buf) {
        const int c = r0 + g;
        const unsigned long long gv_c = __shfl_sync(0xffffffffu, gv, c & 31);
        if (c < nb) {
            if (ix.exl_copy_coalesced) {
                // the group's 4 lanes copy the candidate's contiguous 8 * LB bytes in address order, 64 bytes per instruction (see
                // refine_batch); row r of the block belongs to lane r & 3 of the group, first (r < 4) or second (r >= 4) staged row
                const uint8_t* src = ix.exl + gv_c * ix.exl_stride + 16u * (uint32_t)p;  // this lane's piece of stripe 0
                const uint32_t dst = stage_u32 + buf * buf_bytes + (uint32_t)(g * 4) * lane_stride;
#ifdef RBQ_EXL_COPY_REGULAR
                if ((LB & 63u) == 0u) {  // every 64-byte stripe lies inside one row
#pragma unroll 1
                    for (int r = 0; r < 8; ++r) {
                        const uint32_t d = dst + (uint32_t)(r & 3) * lane_stride + (uint32_t)(r >> 2) * LB + 16u * (uint32_t)p;
#pragma unroll 1
                        for (uint32_t o = 0; o < LB; o += 64u) cp_async16(d + o, src + o);
                        src += LB;
                    }
                } else if (LB == 32u) {  // a stripe covers two whole rows
                    const uint32_t rl = (uint32_t)p >> 1, ol = ((uint32_t)p & 1u) << 4;
#pragma unroll 1
                    for (int st = 0; st < 4; ++st) {
                        const uint32_t row = 2u * (uint32_t)st + rl;
                        cp_async16(dst + (row & 3u) * lane_stride + (row >> 2) * LB + ol, src + 64u * (uint32_t)st);
                    }
                } else
#endif
                {  // piece f = p, p + 4, ... sits in row f / npr (any row size)
                    uint32_t row = (uint32_t)p / npr, pc = (uint32_t)p - row * npr;
                    for (uint32_t f = (uint32_t)p; f < 8u * npr; f += 4u) {
                        cp_async16(dst + (row & 3u) * lane_stride + (row >> 2) * LB + (pc << 4), src + ((size_t)(f - (uint32_t)p) << 4));
                        pc += 4u;
                        while (pc >= npr) {
                            pc -= npr;
                            row += 1u;
                        }
                    }
                }
            } else {
                const uint8_t* src = ix.exl + gv_c * ix.exl_stride + (size_t)p * LB;
                const uint32_t dst = stage_u32 + buf * buf_bytes + my;
                for (uint32_t o = 0; o < LB; o += 16) {
                    cp_async16(dst + o, src + o);
                    cp_async16(dst + LB + o, src + 4u * LB + o);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const bool dbl = stage_bufs > 1u;
    issue(0, 0u);
    float exdot = 0.0f;
    uint32_t buf = 0;
    for (int r0 = 0; r0 < nb; r0 += kRefineSlots2) {
        const int c = r0 + g;  // candidate served by this 4-lane group
        if (dbl && r0 + kRefineSlots2 < nb) {
            issue(r0 + kRefineSlots2, buf ^ 1u);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (ix.exl_copy_coalesced) __syncwarp();  // a lane's rows were copied by the other lanes of its group
        float part = 0.0f;
        if (c < nb) {
            const uint32_t row = stage_u32 + buf * buf_bytes + my;
            part = ex_dot_chain2(row, row + LB, qrow, LB);
        }
        part = part + __shfl_xor_sync(0xffffffffu, part, 2);
        part = part + __shfl_xor_sync(0xffffffffu, part, 1);
        const float v = __shfl_sync(0xffffffffu, part, ((lane - r0) & 7) * 4);
        if (lane >= r0 && lane < r0 + kRefineSlots2) exdot = v;
        if (dbl) buf ^= 1u;
        else if (r0 + kRefineSlots2 < nb) issue(r0 + kRefineSlots2, 0u);  // the lane's own region is free again (only it reads it)
    }
    return exdot;
}

// PAIRED selects the form at compile time (the kernels are instantiated for both; the host picks by padded_dim)
__host__ __device__ inline bool refine_paired_for(uint32_t D) { return D <= 640u; }
template <bool PAIRED>
__device__ __forceinline__ void load_rql_any(unsigned char* rql, uint32_t rql_row, const float* __restrict__ rot, int D, uint32_t lane_bytes, int lane) {
    if (PAIRED) load_rql2(rql, rql_row, rot, D, lane_bytes, lane);
    else load_rql(rql, rql_row, rot, D, lane_bytes, lane);
}
template <bool PAIRED>
__device__ __forceinline__ float refine_batch_any(const DevIndex& ix, unsigned long long gv, int nb, uint32_t stage_u32, uint32_t rql_u32,
                                                  uint32_t exl_row, uint32_t rql_row, uint32_t stage_bufs, int lane) {
    if (PAIRED) return refine_batch2(ix, gv, nb, stage_u32, rql_u32, exl_row, rql_row, stage_bufs, lane);
    return refine_batch(ix, gv, nb, stage_u32, rql_u32, exl_row, rql_row, stage_bufs, lane);
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

// The same list with a register fast path: for k <= 32 entry i lives in lane i (no shared-memory traffic, no
// __syncwarp; insertion = one ballot and two shuffles), otherwise in the shared arrays sd / si.  Same order semantics.
struct TopK {
    float* sd;
    unsigned long long* si;
    int k, cnt;
    bool reg;
    float rd;
    unsigned long long ri;
    __device__ __forceinline__ void init(float* sd_, unsigned long long* si_, int k_) {
        sd = sd_;
        si = si_;
        k = k_;
        cnt = 0;
        reg = k_ <= 32;
        rd = 0.0f;
        ri = 0ull;
    }
    // the k-th distance (the pruning threshold), +inf until the list is full
    __device__ __forceinline__ float theta() const {
        if (cnt < k) return INFINITY;
        return reg ? __shfl_sync(0xffffffffu, rd, k - 1) : sd[k - 1];
    }
    __device__ __forceinline__ void insert(float d, unsigned long long id, int lane) {
        if (!reg) {
            topk_insert(sd, si, cnt, k, d, id, lane);
            return;
        }
        const int pos = __popc(__ballot_sync(0xffffffffu, lane < cnt && rd <= d));
        const float up_d = __shfl_up_sync(0xffffffffu, rd, 1);
        const unsigned long long up_i = __shfl_up_sync(0xffffffffu, ri, 1);
        if (pos >= k) return;
        const int newcnt = cnt < k ? cnt + 1 : k;
        if (lane == pos) {
            rd = d;
            ri = id;
        } else if (lane > pos && lane < newcnt) {
            rd = up_d;
            ri = up_i;
        }
        cnt = newcnt;
    }
    // entry i (i == lane + 32 j): only valid for i < cnt
    __device__ __forceinline__ float dist_at(int i) const { return reg ? rd : sd[i]; }
    __device__ __forceinline__ unsigned long long id_at(int i) const { return reg ? ri : si[i]; }
    __device__ __forceinline__ void set_at(int i, float d, unsigned long long id) {
        if (reg) {
            rd = d;
            ri = id;
        } else {
            sd[i] = d;
            si[i] = id;
        }
    }
};

}  // namespace rbq
