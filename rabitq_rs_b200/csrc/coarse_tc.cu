// coarse_tc.cu -- K4 on the tensor cores: query x centroid GEMM (the only dense contraction of the
// search path) as a hand-written sm_100a kernel: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared
// memory -> tcgen05.mma (kind::f16, BF16 in / FP32 accumulate in TMEM) -> tcgen05.ld epilogue.
//
// Replaces the all-pairs loop of the reference (src/ivf.rs:1782-1789 -> math::l2_distance_sqr / dot)
// as a CANDIDATE GENERATOR only: probe selection must be bit-exact, so the approximate scores are
// followed by an exact FP32 re-score of the near-threshold centroids in the reference's float order
// (coarse.cu, probe_select_tc_kernel).
//
// Precision: operands are split into two bf16 terms, x = hi + lo (+ O(2^-18 |x|)), and the GEMM runs
// over the K-concatenation A' = [q_hi | q_hi | q_lo], B' = [c_hi | c_lo | c_hi], so that
// A'.B'^T = q_hi.c_hi + q_hi.c_lo + q_lo.c_hi = q.c up to 2^-16-relative terms -- one plain GEMM with
// K' = 3*D, no second accumulator.  Scores written: L2: |q|^2 + |c|^2 - 2 q.c ; IP: q.c.
#include <cuda.h>
#include <cuda_bf16.h>

#include <algorithm>

#include "rbq_internal.h"

namespace rbq {

namespace tc {
constexpr int BM = 128;      // queries per tile (UMMA M)
constexpr int BN = 256;      // centroids per tile (UMMA N), fp32 accumulator = 256 TMEM columns
constexpr int BK = 64;       // bf16 per k-block = 128 bytes = one swizzle row
constexpr int UK = 16;       // UMMA K for 16-bit inputs
constexpr int STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_BYTES = BN * BK * 2;  // 32 KB
// Epilogue warps: several per scheduler (a lone epilogue warp per scheduler ran at ~0.13 IPC: dependent-issue latency).  Warp w
// drains lane quarter w % 4 of column part (w - 2) / 4.  The score-matrix drain is bound by its scattered stores and runs best
// with 8 warps, the filter / arg-min drains are instruction-latency bound and take 16 (measured: DESIGN.md).
__host__ __device__ constexpr int epi_warps(int mode) { return mode == kGemmScores ? 8 : 16; }
__host__ __device__ constexpr int threads(int mode) { return (2 + epi_warps(mode)) * 32; }  // warp 0: TMA, warp 1: MMA + TMEM alloc, the rest: epilogue
constexpr int ACC_BUFS = 2;           // accumulator buffers in tensor memory: the epilogue of tile i overlaps the MMAs of tile i+1
constexpr int TMEM_COLS = ACC_BUFS * BN;  // 512 = the whole tensor memory of the SM (one CTA per SM: 193 KB of shared memory)
constexpr size_t SMEM = (size_t)STAGES * (A_BYTES + B_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
}  // namespace tc

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void bar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// shared-memory matrix descriptor, K-major operand, 128-byte swizzle (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) = 8 rows * 128 B |
// version=1 [46,48) | layout_type=SWIZZLE_128B(2) [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// one leader lane of the (converged) warp; unlike `lane == 0` the compiler keeps the guarded code on the uniform path
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}"
        : "=r"(pred));
    return pred != 0;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1,
// A,B K-major, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// monotone map of f32::total_cmp onto u32 (same as coarse.cu::order_key, ascending)
__device__ __forceinline__ uint32_t tc_order_key(float f) {
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// 32 accumulator columns of this warp's TMEM lane quarter -> registers (asynchronous: valid after tmem_ld_wait on the same array)
__device__ __forceinline__ void tmem_ld32(uint32_t (&r)[32], uint32_t taddr) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// waits for every tcgen05.ld of this thread; the array is an in/out operand so that no use of it can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])::"memory");
}

// Persistent GEMM: every CTA walks a sequence of 128 x 256 output tiles.  Three roles: one thread feeds the 4-stage
// shared-memory ring with TMA, one thread issues tcgen05.mma into one of the two 256-column accumulator buffers, four warps
// drain the other buffer.  MODE picks what the drain does with a score:
//   kGemmScores  store it (dense nq x ncols matrix: small centroid tables, the centroid sample of the filter mode)
//   kGemmFilter  compare with the row's threshold and append the few that pass to the row's candidate list
//   kGemmArgmin  keep the row's best (k-means assignment); tiles of one row block are consecutive so it lives in registers
template <int MODE>
__global__ void __launch_bounds__(tc::threads(MODE), 1)
    coarse_gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const GemmEpi e, int num_kb,
                       int mtiles, int ntiles) {
    using namespace tc;
    constexpr int EPI_WARPS = epi_warps(MODE), PARTS = EPI_WARPS / 4, PART_COLS = BN / PARTS;
    extern __shared__ unsigned char gsm_raw[];
    unsigned char* gsm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gsm_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = gsm;
    unsigned char* sB = gsm + (size_t)STAGES * A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(gsm + (size_t)STAGES * (A_BYTES + B_BYTES));
    // bars[0..S) full, [S..2S) empty, [2S..2S+2) accumulator buffer full, [2S+2..2S+4) accumulator buffer drained; then the TMEM base
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * ACC_BUFS);
    const uint32_t full0 = s_u32(bars), empty0 = s_u32(bars + STAGES), tfull0 = s_u32(bars + 2 * STAGES),
                   tempty0 = s_u32(bars + 2 * STAGES + ACC_BUFS);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // this CTA's tiles: strided over the grid, row blocks fastest (CTAs running together share a centroid tile); for the
    // arg-min drain a contiguous range of the column-fastest order (a row block's tiles follow each other)
    const int total = mtiles * ntiles;
    int t_begin, t_end, t_step;
    if (MODE == kGemmArgmin) {
        const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
        t_begin = min(total, (int)blockIdx.x * per);
        t_end = min(total, t_begin + per);
        t_step = 1;
    } else {
        t_begin = (int)blockIdx.x;
        t_end = total;
        t_step = (int)gridDim.x;
    }
    auto tile_mn = [&](int t, int& m, int& n) {
        if (MODE == kGemmArgmin) {
            m = t / ntiles;
            n = t - m * ntiles;
        } else {
            n = t / mtiles;
            m = t - n * mtiles;
        }
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < STAGES; ++s) {
            bar_init(full0 + 8 * s, 1);
            bar_init(empty0 + 8 * s, 1);
        }
        for (int b = 0; b < ACC_BUFS; ++b) {
            bar_init(tfull0 + 8 * b, 1);
            bar_init(tempty0 + 8 * b, EPI_WARPS);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 2 x 256 columns x 128 lanes of fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: the warp walks the loop, one elected lane issues (operands stay in uniform registers) =====
        uint32_t it = 0;
        for (int t = t_begin; t < t_end; t += t_step) {
            int m_blk, n_blk;
            tile_mn(t, m_blk, n_blk);
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % STAGES;
                bar_wait(empty0 + 8 * s, ((it / STAGES) & 1u) ^ 1u);
                if (elect_one()) {
                    bar_expect_tx(full0 + 8 * s, A_BYTES + B_BYTES);
                    tma_load_2d(s_u32(sA + (size_t)s * A_BYTES), &map_a, full0 + 8 * s, kb * BK, m_blk * BM);
                    tma_load_2d(s_u32(sB + (size_t)s * B_BYTES), &map_b, full0 + 8 * s, kb * BK, n_blk * BN);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the whole warp walks the loop, one elected lane issues =====
        // tcgen05.mma takes its operands from uniform registers.  Behind `lane == 0`, with the TMEM base loaded from shared memory,
        // the compiler wraps every MMA in an ELECT + R2UR sequence (~13 instructions, ~190 clk per MMA for one thread -- more than
        // the 128 clk a 128x256x16 MMA occupies the tensor pipe).  A warp-wide OR of the (identical) base is uniform by construction
        // and elect.sync keeps the branch on the uniform path: the MMAs of a k-block then issue back to back.
        constexpr uint32_t idesc = umma_idesc(BM, BN);
        const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
        uint32_t it = 0, ti = 0;
        for (int t = t_begin; t < t_end; t += t_step, ++ti) {
            const uint32_t buf = ti & 1u;
            bar_wait(tempty0 + 8 * buf, ((ti >> 1) & 1u) ^ 1u);  // the epilogue has drained this buffer (free at first use)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc_addr = tmem_u + buf * (uint32_t)BN;
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const uint32_t s = it % STAGES;
                bar_wait(full0 + 8 * s, (it / STAGES) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint64_t da = umma_desc(s_u32(sA + (size_t)s * A_BYTES));
                    const uint64_t db = umma_desc(s_u32(sB + (size_t)s * B_BYTES));
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        const uint32_t acc = (kb | k) ? 1u : 0u;
                        // advancing K by 16 bf16 = 32 bytes inside the swizzle row: +2 in the (addr >> 4) field
                        asm volatile(
                            "{\n"
                            ".reg .pred p;\n"
                            "setp.ne.b32 p, %4, 0;\n"
                            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
                            "}" ::"r"(acc_addr),
                            "l"(da + (uint64_t)(2 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc), "r"(acc)
                            : "memory");
                    }
                    // frees the smem stage once the MMAs above have read it (implies fence::before_thread_sync)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8 * s)
                                 : "memory");
                    if (kb + 1 == num_kb)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tfull0 + 8 * buf)
                                     : "memory");
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> (scores | candidate lists | running arg-min) =====
        const int quarter = warp & 3;  // a warp may only touch TMEM lanes 32*(warp%4) .. +31
        const int half = (warp - 2) >> 2;  // its part of the tile's 256 columns
        const bool l2 = e.metric == RBQ_METRIC_L2;
        uint32_t ti = 0;
        int best_m = -1;  // kGemmArgmin: row block the running best belongs to
        unsigned long long best = ~0ull;
        auto flush_best = [&]() {
            const int row = best_m * BM + quarter * 32 + lane;
            if (best_m >= 0 && row < e.nq && best != ~0ull) atomicMin(e.best + row, best);
        };
        for (int t = t_begin; t < t_end; t += t_step, ++ti) {
            int m_blk, n_blk;
            tile_mn(t, m_blk, n_blk);
            const uint32_t buf = ti & 1u;
            const int row = m_blk * BM + quarter * 32 + lane;
            const bool row_ok = row < e.nq;
            const float qq = (row_ok && l2 && !e.shifted) ? __ldg(e.qn2 + row) : 0.0f;
            float thr = 0.0f;
            if (MODE == kGemmFilter) thr = row_ok ? __ldg(e.thr + row) : (l2 ? -INFINITY : INFINITY);
            if (MODE == kGemmArgmin && m_blk != best_m) {
                flush_best();
                best_m = m_blk;
                best = ~0ull;
            }
            bar_wait(tfull0 + 8 * buf, (ti >> 1) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int ncols_tile = min(BN, e.ncols - n_blk * BN);  // columns of this tile that exist
            // one 32-column group of the tile: r = its accumulators
            auto process = [&](const uint32_t (&r)[32], const int c0) {
                const int n0 = n_blk * BN + c0;
                const bool full32 = n0 + 32 <= e.ncols && (e.ncols & 3) == 0;
                if (MODE == kGemmScores) {
                    if (row_ok) {
                        float* out = e.scores + (size_t)row * e.ncols + n0;
                        if (full32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                float4 v;
                                float* pv = &v.x;
                                float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (l2) cc = __ldg(reinterpret_cast<const float4*>(e.cn2 + n0 + j));
                                const float* pc = &cc.x;
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float g = __uint_as_float(r[j + u]);
                                    pv[u] = !l2 ? g : e.shifted ? __fmaf_rn(-2.0f, g, pc[u]) : (qq + pc[u]) - 2.0f * g;
                                }
                                *reinterpret_cast<float4*>(out + j) = v;
                            }
                        } else {
                            for (int j = 0; j < 32; ++j)
                                if (n0 + j < e.ncols) {
                                    const float g = __uint_as_float(r[j]);
                                    out[j] = !l2 ? g : e.shifted ? __fmaf_rn(-2.0f, g, e.cn2[n0 + j]) : (qq + e.cn2[n0 + j]) - 2.0f * g;
                                }
                        }
                    }
                } else {
                    const int lim = min(32, e.ncols - n0);
                    // scores of the 32 columns (independent: no branches between them)
                    float sc[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (l2) {
                            if (full32) cc = __ldg(reinterpret_cast<const float4*>(e.cn2 + n0 + j));
                            else {
                                float* pc = &cc.x;
                                for (int u = 0; u < 4; ++u) pc[u] = n0 + j + u < e.ncols ? e.cn2[n0 + j + u] : 0.0f;
                            }
                        }
                        const float* pc = &cc.x;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float g = __uint_as_float(r[j + u]);
                            sc[j + u] = !l2 ? g : e.shifted ? __fmaf_rn(-2.0f, g, pc[u]) : (qq + pc[u]) - 2.0f * g;
                        }
                    }
                    if (MODE == kGemmFilter) {
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) mask |= (uint32_t)(l2 ? sc[j] <= thr : sc[j] >= thr) << j;
                        if (lim < 32) mask &= (1u << lim) - 1u;
                        if (mask) {
                            // One bump of the row's counter for all hits of the group, then straight-line predicated stores in
                            // column order.  (The first version staged single hits in shared memory behind a branch per column:
                            // with ~1 % of the scores passing, SOME lane of a warp has a hit in nearly every group, so every warp
                            // walked all 32 branches -- 18 instructions per score, and the drain, not the tensor pipe, set the
                            // kernel's time on the 16 384 / 65 536-list tables.  A two-pass variant -- masks first, one bump per
                            // tile, scores re-read from tensor memory -- measured slower: more instructions.)
                            const uint32_t n = (uint32_t)__popc(mask);
                            const uint32_t base = atomicAdd(e.cand_cnt + row, n);
                            if (base + n <= e.cap) {
                                unsigned long long addr = reinterpret_cast<unsigned long long>(e.cand + (size_t)row * e.cap + base);
#pragma unroll
                                for (int j = 0; j < 32; ++j) {
                                    const uint32_t hit = (mask >> j) & 1u;
                                    asm volatile(
                                        "{\n"
                                        ".reg .pred p;\n"
                                        "setp.ne.u32 p, %0, 0;\n"
                                        "@p st.global.v2.b32 [%1], {%2, %3};\n"
                                        "}" ::"r"(hit),
                                        "l"(addr), "r"(__float_as_uint(sc[j])), "r"((uint32_t)(n0 + j))
                                        : "memory");
                                    addr += 8ull * hit;
                                }
                            } else {  // the list overflows (the query takes the exact fallback): keep what fits
                                const uint32_t room = base < e.cap ? e.cap - base : 0u;
                                CandRec* out = e.cand + (size_t)row * e.cap + base;
                                uint32_t w = 0;
                                for (int j = 0; j < 32; ++j) {
                                    const bool hit = (mask >> j) & 1u;
                                    if (hit && w < room) out[w] = CandRec{sc[j], (uint32_t)(n0 + j)};
                                    w += hit ? 1u : 0u;
                                }
                            }
                        }
                    } else {  // arg-min over max(score, 0), ties to the lower column (reference src/kmeans.rs:505-516)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const unsigned long long key = ((unsigned long long)tc_order_key(fmaxf(sc[j], 0.0f)) << 32) | (uint32_t)(n0 + j);
                            if (j < lim && key < best) best = key;
                        }
                    }
                }
            };
            {
                const int cbeg = half * PART_COLS;  // this warp's columns of the tile: [cbeg, cbeg + PART_COLS)
                const uint32_t tbase = tmem + ((uint32_t)(quarter * 32) << 16) + buf * (uint32_t)BN + (uint32_t)cbeg;
                const int ngroups = max(0, min(PART_COLS, ncols_tile - cbeg) + 31) / 32;
                if (PARTS >= 4) {
                    // many warps per scheduler hide the TMEM latency: one buffer keeps the register count low enough for 18 warps
                    uint32_t ra[32];
                    for (int g = 0; g < ngroups; ++g) {
                        tmem_ld32(ra, tbase + 32u * (uint32_t)g);
                        tmem_ld_wait(ra);
                        process(ra, cbeg + 32 * g);
                    }
                } else {
                    // software pipeline over the groups: the TMEM load of group g+1 is in flight while group g is processed
                    uint32_t ra[32], rb[32];
                    if (ngroups > 0) {
                        tmem_ld32(ra, tbase);
                        tmem_ld_wait(ra);
                    }
                    for (int g = 0; g < ngroups; g += 2) {
                        if (g + 1 < ngroups) tmem_ld32(rb, tbase + 32u * (uint32_t)(g + 1));
                        process(ra, cbeg + 32 * g);
                        if (g + 1 < ngroups) {
                            tmem_ld_wait(rb);
                            if (g + 2 < ngroups) tmem_ld32(ra, tbase + 32u * (uint32_t)(g + 2));
                            process(rb, cbeg + 32 * (g + 1));
                            if (g + 2 < ngroups) tmem_ld_wait(ra);
                        }
                    }
                }
            }
            // this warp's TMEM reads are complete: hand the buffer back to the MMA issuer
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(tempty0 + 8 * buf);
        }
        if (MODE == kGemmArgmin) flush_best();
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

// ---- operand preparation ------------------------------------------------------------------------------
// rows x D fp32 -> rows x 3D bf16 in the order the GEMM wants, plus |x|^2 per row.
// query side: [hi | hi | lo]; centroid side: [hi | lo | hi]
__global__ void split_bf16_kernel(const float* __restrict__ x, int rows, int in_dim, int D, int centroid_side,
                                  __nv_bfloat16* __restrict__ out, float* __restrict__ n2) {
    const int row = blockIdx.x;
    if (row >= rows) return;
    const float* xr = x + (size_t)row * in_dim;
    __nv_bfloat16* o = out + (size_t)row * 3 * D;
    float acc = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float v = i < in_dim ? xr[i] : 0.0f;  // columns in_dim..D are zero padding (k-means in the original space)
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
        o[i] = hi;
        o[D + i] = centroid_side ? lo : hi;
        o[2 * D + i] = centroid_side ? hi : lo;
        acc += v * v;
    }
    __shared__ float red[32];
    for (int o2 = 16; o2 > 0; o2 >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o2);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        n2[row] = t;
    }
}

int launch_split_bf16_pad(const float* d_x, size_t rows, int in_dim, int D, int centroid_side, void* d_out, float* d_n2, cudaStream_t st) {
    if (rows == 0) return RBQ_OK;
    split_bf16_kernel<<<(unsigned)rows, 128, 0, st>>>(d_x, (int)rows, in_dim, D, centroid_side, reinterpret_cast<__nv_bfloat16*>(d_out), d_n2);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}
int launch_split_bf16(const float* d_x, size_t rows, int D, int centroid_side, void* d_out, float* d_n2, cudaStream_t st) {
    return launch_split_bf16_pad(d_x, rows, D, D, centroid_side, d_out, d_n2, st);
}

// rows idx[0..n) of a row-major byte matrix -> dst (16-byte granules); builds the centroid sample of the filter mode
__global__ void gather_rows_kernel(const uint4* __restrict__ src, size_t row_u4, const uint32_t* __restrict__ idx, uint4* __restrict__ dst) {
    const uint4* s = src + (size_t)idx[blockIdx.x] * row_u4;
    uint4* d = dst + (size_t)blockIdx.x * row_u4;
    for (size_t i = threadIdx.x; i < row_u4; i += blockDim.x) d[i] = s[i];
}
int launch_gather_rows(const void* d_src, size_t row_bytes, const uint32_t* d_idx, size_t n, void* d_dst, cudaStream_t st) {
    if (n == 0) return RBQ_OK;
    if (row_bytes % 16 != 0) return fail(RBQ_INVALID_CONFIG, "gather_rows: row size must be a multiple of 16 bytes");
    gather_rows_kernel<<<(unsigned)n, 128, 0, st>>>(reinterpret_cast<const uint4*>(d_src), row_bytes / 16, d_idx, reinterpret_cast<uint4*>(d_dst));
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// ---- tensor maps + launch ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int get_encode() {
    if (g_encode) return RBQ_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    RBQ_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !fn) return fail(RBQ_CUDA_ERROR, "cuTensorMapEncodeTiled is not available");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return RBQ_OK;
}

// 2-D bf16 tensor [rows][k_cols] with row pitch `pitch_elems`: the terms = 1 view reads only the leading D columns of a 3D-wide split
static int make_map_pitch(CUtensorMap* map, const void* base, size_t rows, size_t k_cols, size_t pitch_elems, int box_rows) {
    int rc = get_encode();
    if (rc) return rc;
    cuuint64_t dims[2] = {(cuuint64_t)k_cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
    cuuint32_t box[2] = {(cuuint32_t)tc::BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RBQ_CUDA_ERROR, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return RBQ_OK;
}

static int gemm_sms() {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 148;
    return sms > 0 ? sms : 148;
}

template <int MODE>
static int launch_gemm_mode(const CUtensorMap& ma, const CUtensorMap& mb, const GemmEpi& epi, int num_kb, int mtiles, int ntiles, cudaStream_t st) {
    const size_t smem = tc::SMEM;
    RBQ_CUDA(cudaFuncSetAttribute(coarse_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long total = (long long)mtiles * ntiles;
    const int grid = (int)std::min<long long>(total, gemm_sms());
    coarse_gemm_kernel<MODE><<<grid, tc::threads(MODE), smem, st>>>(ma, mb, epi, num_kb, mtiles, ntiles);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int launch_coarse_gemm(int mode, const void* d_a, size_t rows, const void* d_b, size_t cols, int D, int terms, const GemmEpi& epi,
                       cudaStream_t st) {
    if (rows == 0 || cols == 0) return RBQ_OK;
    const size_t pitch = (size_t)3 * D, K = terms == 1 ? (size_t)D : pitch;
    CUtensorMap ma, mb;
    int rc;
    if ((rc = make_map_pitch(&ma, d_a, rows, K, pitch, tc::BM))) return rc;
    if ((rc = make_map_pitch(&mb, d_b, cols, K, pitch, tc::BN))) return rc;
    const int num_kb = (int)((K + tc::BK - 1) / tc::BK);
    const size_t mt = (rows + tc::BM - 1) / tc::BM, nt = (cols + tc::BN - 1) / tc::BN;
    if (mt * nt > 0x7fffffffull) return fail(RBQ_INVALID_CONFIG, "coarse GEMM: too many tiles in one launch");
    switch (mode) {
        case kGemmScores: return launch_gemm_mode<kGemmScores>(ma, mb, epi, num_kb, (int)mt, (int)nt, st);
        case kGemmFilter: return launch_gemm_mode<kGemmFilter>(ma, mb, epi, num_kb, (int)mt, (int)nt, st);
        default: return launch_gemm_mode<kGemmArgmin>(ma, mb, epi, num_kb, (int)mt, (int)nt, st);
    }
}

// scores[nq][nlist] (approximate).  d_qsplit: nq x 3D bf16, d_csplit: nlist x 3D bf16.
int launch_coarse_tc(const DevIndex& ix, const void* d_qsplit, const float* d_qn2, size_t nq, float* d_scores, cudaStream_t st, int terms) {
    GemmEpi e;
    e.nq = (int)nq;
    e.ncols = (int)ix.nlist;
    e.metric = ix.metric;
    e.qn2 = d_qn2;
    e.cn2 = ix.cent_n2;
    e.scores = d_scores;
    return launch_coarse_gemm(kGemmScores, d_qsplit, nq, ix.cent_split, ix.nlist, ix.D, terms, e, st);
}

}  // namespace rbq
