// tail_tc.cu -- the list-major FastScan of the tail stage as an exact integer GEMM on the 5th-gen tensor cores.
//
// K7 (simd::accumulate_batch_avx2, reference src/simd.rs:972-1184; scalar ground truth :1462-1525) is
//     accu[v][q] = sum over codebooks cb of lut_u8[q][16*cb + nibble(v, cb)]
// i.e. a product of a one-hot matrix (vectors x 16*ncb, one 1 per codebook) with the queries' LUT bytes -- u8 x u8
// with s32 accumulation, which is exact, so the tensor cores return the reference's integer bit for bit.  The PRMT
// kernel (scan_tail.cu) spends ~3 ALU instructions per 4 lookups and is bound by the integer pipe; here a lookup
// costs 16 bytes of a one-hot row written once to tensor memory and shared by all the queries that probe the list, and
// the sums run on tcgen05.mma (kind::i8, M = 128 vectors, N = 64 queries, K = 32 bytes = 2 codebooks per instruction)
// with the accumulators in TMEM.
//
// Two CTAs per SM, one work item (a list and <= 64 of the (query, rank) pairs probing it) at a time:
//   * the A operand (one-hot rows) lives in TENSOR MEMORY, not shared memory: producer thread (warp w, lane v) owns
//     vector v of block w of the group = TMEM lane 32*(w%4)+v of tile w/4, builds the 128 one-hot bytes of a K-chunk
//     (8 codebooks) in 32 registers (one BMSK per word) and writes them with one tcgen05.st.32x32b.x32 -- the 32x
//     expansion of the packed codes never touches the shared-memory pipe, which only carries the B operand;
//   * the 64 queries' LUT slice of the chunk is copied into a 128B-swizzled K-major B tile with cp.async (the LUT row
//     of a query is already K-major: lut[16*cb + nibble], reference src/simd.rs:818-840);
//   * one thread issues the MMAs of the chunk (tcgen05.mma kind::i8, A from TMEM, B from shared memory; 4 per
//     128-vector tile); tcgen05.commit frees the stage, so the next chunk is built while the tensor core works;
//   * after the last chunk the 8 warps read the sums back (tcgen05.ld), evaluate K8 (compute_batch_distances_u16,
//     src/simd.rs:2090-2140, AVX2 operation order) per (vector, query) and keep the candidates whose lower bound
//     beats the query's head threshold -- the same survivors as the PRMT kernel, appended in any order (the replay
//     sorts them).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "scan_common.cuh"

namespace rbq {

namespace tt {
#ifndef RBQ_TT_MT
#define RBQ_TT_MT 2
#endif
constexpr int MT_MAX = RBQ_TT_MT;                    // 128-vector tiles (4 blocks each) accumulated concurrently
constexpr int NQ = 64;                       // queries per item = UMMA N
constexpr int KCH = 128;                     // bytes of K per chunk = 8 codebooks = one swizzle row
constexpr int STAGES = 2;                    // A stages (tensor memory)
constexpr int PD_MAX = 6;                    // largest prefetch distance in chunks (LUT slices from L2, packed codes from HBM)
// With prefetch distance PD the B ring has PD + STAGES stages: chunk c + PD lands in the stage chunk c - STAGES released, and
// the producers prefetch right after they got A stage c % STAGES back, i.e. after that MMA has finished (the A ring lets the
// tensor core lag the producers by at most STAGES chunks).  The raw packed-code ring (4 bytes per producer thread and chunk,
// read by the producers only) has PD + 2 slots.
constexpr int PRODUCER_WARPS = 4 * MT_MAX;   // one block of the group per warp
constexpr int ISSUER_WARPS = MT_MAX;         // one MMA issuer warp per 128-vector tile: the tiles accumulate independently, and one thread
                                             // needs ~190 clk to issue one of these small MMAs (descriptor moves through the uniform datapath)
constexpr int THREADS = (PRODUCER_WARPS + ISSUER_WARPS) * 32;
constexpr int A_COLS = KCH / 4;              // TMEM columns of one tile's K-chunk (4 one-hot bytes per 32-bit column)
constexpr int B_STAGE = NQ * KCH;            // 8 KB
constexpr int SURV_CAP = 256;                // survivors staged in shared memory between flushes
constexpr int ACC_COLS = MT_MAX * NQ;        // s32 accumulators: tile mt at column mt * NQ
constexpr int CTAS_PER_SM = MT_MAX == 1 ? 4 : 2;
constexpr int TMEM_COLS = 512 / CTAS_PER_SM;               // accumulators + STAGES x MT_MAX A chunks (power of two; two CTAs share the SM's 512)
static_assert(ACC_COLS + STAGES * MT_MAX * A_COLS <= TMEM_COLS, "tensor memory budget");
struct Misc {
    float4 c0[NQ];   // delta, sum_vl, k1x, g_add
    float4 c1[NQ];   // g_error, tau, non-finite fallback, unused
    uint32_t q[NQ];  // query of pair slot n (0xffffffff: unused slot)
    uint32_t rank[NQ];
    uint32_t sq[SURV_CAP];
    Survivor ss[SURV_CAP];
    uint64_t bars[2 * STAGES + 1];  // full[stage], empty[stage], accumulators done
    alignas(16) uint32_t raw[PD_MAX + 2][PRODUCER_WARPS * 32];  // packed code bytes in flight (cp.async)
    uint32_t tmem_base, item, surv_n, pad;
};
constexpr size_t smem_bytes(int pd) { return (size_t)(pd + STAGES) * B_STAGE + sizeof(Misc) + 1024 /*alignment slack*/; }
}  // namespace tt

// K-major operand, 128-byte swizzle: start>>4 [0,14) | LBO>>4 [16,30) (unused) | SBO>>4 [32,46) = 8 rows * 128 B |
// version 1 [46,48) | SWIZZLE_128B (2) [61,64)   (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t tt_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// kind::i8 instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2) [4,6), A = B = unsigned 8-bit (0) [7,10) / [10,13),
// both K-major, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t tt_idesc(uint32_t n) { return (2u << 4) | ((n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); }

__device__ __forceinline__ void tt_bar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "TT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra TT_DONE;\n"
        "bra TT_WAIT;\n"
        "TT_DONE:\n"
        "}" ::"r"(bar),
        "r"(parity), "r"(kMbarSuspendHintNs)
        : "memory");
}
// 1 << pos for pos in [0, 32), 0 otherwise (pos is taken as unsigned, so "negative" positions give 0)
__device__ __forceinline__ uint32_t onehot32(uint32_t pos) {
    uint32_t d;
    asm("bmsk.clamp.b32 %0, %1, 1;" : "=r"(d) : "r"(pos));
    return d;
}

// one leader lane of the (converged) warp; unlike `lane == 0` the compiler keeps the guarded code on the uniform path
__device__ __forceinline__ bool tt_elect_one() {
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
__device__ __forceinline__ void tt_bar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// barrier among the 8 producer warps only (the issuer warp never joins it)
__device__ __forceinline__ void tt_sync_producers() { asm volatile("bar.sync 1, %0;" ::"n"(tt::PRODUCER_WARPS * 32) : "memory"); }

#ifdef RBQ_TT_PROBE
// dev-only phase clocks of producer warp 0 lane 0 of every CTA (build with -DRBQ_TT_PROBE; printed by launch_tail_tc):
// [0] item set-up, [1] wait empty, [2] cp.async wait, [3] build row, [4] st + wait::st, [5] fences + arrive, [6] wait accd,
// [7] epilogue, [8] chunks, [9] items, [10] total
__device__ unsigned long long g_tt_probe[16];
__device__ int g_tt_variant;  // timing experiments (results become wrong): bit 0 no proxy fence, bit 1 no raw-code copies, bit 2 no B copies
#define TT_VARIANT(bit) ((g_tt_variant >> (bit)) & 1)
#define TT_CLK(i)                                            \
    do {                                                     \
        if (tid == 0) {                                      \
            const long long t_ = clock64();                  \
            pr_[i] += (unsigned long long)(t_ - t_last_);    \
            t_last_ = t_;                                    \
        }                                                    \
    } while (0)
#else
#define TT_CLK(i) do { } while (0)
#define TT_VARIANT(bit) 0
#endif

// FULLK: the codebook count is a multiple of 8 (padded_dim % 32 == 0), so no K-chunk is partial and the producers skip the
// per-codebook "does it exist" selects
template <bool WIDE, int PD, bool FULLK>
__global__ void __launch_bounds__(tt::THREADS, tt::CTAS_PER_SM) tail_tc_kernel(DevIndex ix, TailArgs a) {
    using namespace tt;
#ifdef RBQ_TT_PROBE
    unsigned long long pr_[13] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long t_last_ = clock64();
    const long long t_begin_ = t_last_;
#endif
    constexpr int BSTAGES = PD + STAGES, RSTAGES = PD + 2;
    extern __shared__ unsigned char tt_raw[];
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(tt_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sB = sm;                                  // [BSTAGES][NQ rows][128 B]
    Misc* mi = reinterpret_cast<Misc*>(sm + (size_t)BSTAGES * B_STAGE);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = ix.D, ncb = D / 4;
    const uint32_t B = ix.block_stride;
    const uint32_t nkc = ((uint32_t)ncb + 7u) / 8u;  // K-chunks of 8 codebooks
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    const uint32_t sB_u32 = smem_u32(sB);
    // barriers: full[stage] (256 producer arrivals), empty[stage] (one tcgen05.commit per issuer warp), accumulators done (ditto)
    const uint32_t full0 = smem_u32(&mi->bars[0]), empty0 = smem_u32(&mi->bars[STAGES]), accd = smem_u32(&mi->bars[2 * STAGES]);
    const bool issuer = warp >= PRODUCER_WARPS;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, PRODUCER_WARPS * 32);  // every producer thread arrives (measured faster than syncwarp + one lane)
            mbar_init(empty0 + 8 * s, ISSUER_WARPS);
        }
        mbar_init(accd, ISSUER_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mi->surv_n = 0;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&mi->tmem_base)), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = mi->tmem_base;

    uint32_t chunk_seq = 0;  // running chunk counter: A stage = chunk_seq % 2, B stage = chunk_seq % 4
    uint32_t groups = 0;     // accumulator groups finished so far (phase of the `accd` barrier)

    // fetch role of a producer thread: byte quad qd of codebook row cbl of its warp's block-chunk (cp.async into `raw`)
    const int qd = lane >> 3, cbl = lane & 7;
    // A production role: vector `lane` of the block.  Its nibble of a codebook sits in byte vp of the codebook's 16 code
    // bytes (KPERM0[vp] = lane & 15; reference src/simd.rs:774,876-902), high nibble for vectors 16..31; in `raw` that byte
    // is byte (vp & 3) of word (vp >> 2) * 8 + codebook.
    const int vp = ((lane & 7) << 1) | ((lane & 15) >> 3);
    const uint32_t nib_rot = (8u * (uint32_t)(vp & 3) + 4u * (uint32_t)(lane >> 4) + 29u) & 31u;  // rotate right by shift - 3
    // epilogue role (and the TMEM lanes this warp may touch): lane quarter lq of tile warp >> 2
    const int lq = warp & 3;

    // copies the staged survivors to the per-query buffers (one global atomic each, all in flight together)
    auto flush_survivors = [&]() {
        tt_sync_producers();
        const uint32_t n = min(mi->surv_n, (uint32_t)SURV_CAP);
        for (uint32_t i = tid; i < n; i += PRODUCER_WARPS * 32) {
            const uint32_t q = mi->sq[i];
            const uint32_t slot = atomicAdd(&a.surv_cnt[q], 1u);
            if (slot < a.surv_cap) a.surv[(size_t)q * a.surv_cap + slot] = mi->ss[i];
        }
        tt_sync_producers();
        if (tid == 0) {
            if (a.stats && n) atomicAdd(&a.stats->survivors, (unsigned long long)n);
            mi->surv_n = 0;
        }
        tt_sync_producers();
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) mi->item = atomicAdd(&a.counters[1], 1u);
        __syncthreads();
        const uint32_t item = mi->item;
        if (item >= a.counters[0]) break;
#ifdef RBQ_TT_PROBE
        if (tid == 0) pr_[9] += 1;
#endif
        // the staged survivors go out when the buffer is half full (everyone reads the same count here: nobody appends
        // between the barriers above and the first epilogue of the item)
        if (!issuer && mi->surv_n > (uint32_t)SURV_CAP / 2) flush_survivors();
        const TailItem it = a.items[item];
        const uint32_t nv = ix.list_n[it.cid], nb = (nv + kBatch - 1) / kBatch;
        const uint32_t P = it.pair_count;
        // UMMA N = the item's pairs rounded up to 16: the integer pipe's time is proportional to it
        const uint32_t idesc = tt_idesc(min((uint32_t)NQ, (P + 15u) & ~15u));

        if (issuer) {
            // ===== MMA issuer of tile `mt`: waits for a produced stage, issues the tile's MMAs, commits the stage back =====
            // Every operand of tcgen05.mma travels through the uniform datapath.  Values the compiler cannot prove warp-uniform
            // (anything loaded from memory or derived from threadIdx) cost an ELECT + six R2UR moves per instruction -- ~190 clk
            // per MMA for one thread, which made the issuer the bottleneck of these small MMAs.  A warp-wide OR of identical
            // values (REDUX) is uniform by construction, and elect.sync instead of `lane == 0` keeps the branch uniform: the MMAs
            // then issue back to back from uniform registers.
            const uint32_t mt = __reduce_or_sync(0xffffffffu, (uint32_t)(warp - PRODUCER_WARPS));
            const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem);
            const uint32_t nb_u = __reduce_or_sync(0xffffffffu, nb), idesc_u = __reduce_or_sync(0xffffffffu, idesc);
            const uint32_t nkc_u = __reduce_or_sync(0xffffffffu, nkc);
            const uint32_t d_tm = tmem_u + mt * (uint32_t)NQ;
            for (uint32_t b0 = 0; b0 < nb_u; b0 += 4 * MT_MAX) {
                const uint32_t nbg = min((uint32_t)(4 * MT_MAX), nb_u - b0), mt_cnt = (nbg + 3) / 4;
                const bool have_tile = mt < mt_cnt;
                for (uint32_t kc = 0; kc < nkc_u; ++kc, ++chunk_seq) {
                    const uint32_t s = chunk_seq % STAGES;
                    tt_bar_wait(full0 + 8 * s, (chunk_seq / STAGES) & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (tt_elect_one()) {
                        if (have_tile) {
                            const uint64_t db = tt_desc(sB_u32 + (chunk_seq % BSTAGES) * B_STAGE);
                            const uint32_t ta = tmem_u + (uint32_t)ACC_COLS + (s * MT_MAX + mt) * (uint32_t)A_COLS;
#pragma unroll
                            for (int k = 0; k < KCH / 32; ++k) {
                                const uint32_t acc = (kc | (uint32_t)k) ? 1u : 0u;
                                asm volatile(
                                    "{\n"
                                    ".reg .pred p;\n"
                                    "setp.ne.b32 p, %4, 0;\n"
                                    "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
                                    "}" ::"r"(d_tm),
                                    "r"(ta + (uint32_t)(8 * k)), "l"(db + (uint64_t)(2 * k)), "r"(idesc_u), "r"(acc)
                                    : "memory");
                            }
                            // the commits track this thread's MMAs only: every issuer arrives once per stage / group
                            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8 * s)
                                         : "memory");
                            if (kc + 1 == nkc_u)
                                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(accd)
                                             : "memory");
                        } else {  // the group has no such tile: plain arrivals keep the barrier counts
                            tt_bar_arrive(empty0 + 8 * s);
                            if (kc + 1 == nkc_u) tt_bar_arrive(accd);
                        }
                    }
                    __syncwarp();
                }
            }
            continue;
        }

        // ===== producers / epilogue (8 warps) =====
        const uint8_t* lbase = ix.blocks + (size_t)ix.blk_off[it.cid] * B;
        const unsigned long long vbase = ix.vec_off[it.cid];
        if (tid == 0 && a.stats) {
            atomicAdd(&a.stats->tail_blocks, (unsigned long long)P * nb);
            atomicAdd(&a.stats->tail_pairs, (unsigned long long)P);
            atomicAdd(&a.stats->candidates, (unsigned long long)P * nv);
        }
        // per-pair constants of the item
        if (tid < NQ) {
            uint32_t q = 0xffffffffu, rank = 0;
            float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = make_float4(0.f, -INFINITY, 0.f, 0.f);
            if ((uint32_t)tid < P) {
                const uint32_t pid = a.pairs[it.pair_begin + tid];
                q = pid / a.nprobe;
                rank = pid - q * a.nprobe;
                const QueryScalars s = a.qs[q];
                const Probe p = a.probes[pid];
                c0 = make_float4(s.delta, s.sum_vl, s.k1x, p.g_add);
                c1 = make_float4(p.g_error, a.tau[q], l2 ? 0.0f : -(p.dot_qc + s.qnorm), 0.f);
            }
            mi->q[tid] = q;
            mi->rank[tid] = rank;
            mi->c0[tid] = c0;
            mi->c1[tid] = c1;
        }
        tt_sync_producers();
        // this thread's two 16-byte pieces of every B chunk (row r = pair slot, piece j): source row, swizzled destination and
        // the K range in which the piece exists -- fixed for the whole item
        const uint8_t* bsrc[NQ * 8 / (PRODUCER_WARPS * 32)];
        uint32_t bdst[NQ * 8 / (PRODUCER_WARPS * 32)], blim[NQ * 8 / (PRODUCER_WARPS * 32)];
#pragma unroll
        for (int i = 0; i < NQ * 8 / (PRODUCER_WARPS * 32); ++i) {
            const int piece = tid + PRODUCER_WARPS * 32 * i, r = piece >> 3, j = piece & 7;
            const uint32_t q = mi->q[r];
            bsrc[i] = a.lut + (size_t)(q == 0xffffffffu ? 0u : q) * D * 4 + 16u * (uint32_t)j;
            bdst[i] = sB_u32 + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
            blim[i] = (q != 0xffffffffu && (uint32_t)D * 4u > 16u * (uint32_t)j) ? (uint32_t)D * 4u - 16u * (uint32_t)j : 0u;  // kc * KCH < blim
        }

        for (uint32_t b0 = 0; b0 < nb; b0 += 4 * MT_MAX) {  // accumulator group: up to 16 blocks = 4 tiles of 128 vectors
            const uint32_t nbg = min((uint32_t)(4 * MT_MAX), nb - b0), mt_cnt = (nbg + 3) / 4;
            // chunk kc of the group: B = LUT slice [kc*128, kc*128+128) of the item's queries (row r = pair slot) straight into
            // its swizzled stage, and this thread's 4 packed code bytes (byte quad qd of codebook row 8*kc + cbl of block
            // bg = warp) into the raw ring -- both asynchronous, one commit group per chunk
            const uint8_t* rsrc = lbase + (size_t)(b0 + min((uint32_t)warp, nbg - 1u)) * B + 16u * (uint32_t)cbl + 4u * (uint32_t)qd;
            const uint32_t rlim = (uint32_t)warp < nbg && (uint32_t)cbl < (uint32_t)ncb ? ((uint32_t)ncb - (uint32_t)cbl + 7u) / 8u : 0u;  // kc < rlim
            const uint32_t rdst = smem_u32(&mi->raw[0][tid]);
            auto prefetch = [&](uint32_t kc, uint32_t seq) {
                if (kc < nkc) {
                    const uint32_t koff = kc * KCH, bst = (seq % BSTAGES) * B_STAGE;
#pragma unroll
                    for (int i = 0; i < NQ * 8 / (PRODUCER_WARPS * 32); ++i)
                        if (koff < blim[i] && !TT_VARIANT(2)) cp_async16(bdst[i] + bst, bsrc[i] + koff);
                    if (kc < rlim && !TT_VARIANT(1))
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(rdst + (seq % RSTAGES) * (uint32_t)(PRODUCER_WARPS * 32 * 4)),
                                     "l"(rsrc + koff)
                                     : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            // every earlier MMA has completed (accumulator barrier), so all stages are free
#pragma unroll
            for (int d = 0; d < PD; ++d) prefetch((uint32_t)d, chunk_seq + d);
            TT_CLK(0);
            for (uint32_t kc = 0; kc < nkc; ++kc, ++chunk_seq) {
                const uint32_t s = chunk_seq % STAGES, use = chunk_seq / STAGES;
#ifdef RBQ_TT_PROBE
                if (tid == 0) pr_[8] += 1;
#endif
                // chunk_seq - 2 was the last reader of A stage s.  (Building the row before this wait was measured slower:
                // 0.65 vs 0.61 ms at GIST/10k -- the row would stay live in 32 registers across the wait.)
                if (use > 0) tt_bar_wait(empty0 + 8 * s, (use - 1) & 1u);
                TT_CLK(1);
                prefetch(kc + PD, chunk_seq + PD);  // the operands of chunk kc + PD start travelling
                TT_CLK(11);
                asm volatile("cp.async.wait_group %0;" ::"n"(PD) : "memory");  // this thread's pieces of chunk kc have landed
                TT_CLK(2);
                __syncwarp();  // the other lanes' pieces of the warp's raw words have landed too
                TT_CLK(12);
                // the one-hot row of this thread's vector for codebooks 8*kc .. 8*kc+7, built in registers
                uint32_t r[32];
                if ((uint32_t)warp < nbg) {
                    const uint32_t rw = smem_u32(&mi->raw[chunk_seq % RSTAGES][warp * 32 + (vp >> 2) * 8]);
                    const uint4 w0 = lds128(rw), w1 = lds128(rw + 16u);
                    const uint32_t W[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    const uint32_t live = min(8u, (uint32_t)ncb - kc * 8u);  // codebooks of this chunk that exist (uniform)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        // bit position of the 1 inside the codebook's 128-bit one-hot: 8 * nibble = (W ror (shift - 3)) & 0x78;
                        // past the last codebook: no bit at all
                        uint32_t pos = __funnelshift_r(W[c], W[c], nib_rot) & 0x78u;
                        if (!FULLK && (uint32_t)c >= live) pos = 0xffffff00u;
                        r[4 * c + 0] = onehot32(pos);
                        r[4 * c + 1] = onehot32(pos - 32u);
                        r[4 * c + 2] = onehot32(pos - 64u);
                        r[4 * c + 3] = onehot32(pos - 96u);
                    }
                }
#ifdef RBQ_TT_PROBE
                if (tid == 0) {  // keep the row's last word live up to here so that the build is timed, not sunk below the clock read
                    asm volatile("" ::"r"(r[31]), "r"(r[0]));
                }
#endif
                TT_CLK(3);
                if ((uint32_t)warp < nbg) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)ACC_COLS + (s * MT_MAX + (uint32_t)(warp >> 2)) * (uint32_t)A_COLS;
                    asm volatile(
                        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
                        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
                        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
                        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                        : "memory");
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                TT_CLK(4);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                if (!TT_VARIANT(0)) fence_proxy_async();  // its B pieces (generic-proxy writes) -> visible to the tensor core's async-proxy reads
                tt_bar_arrive(full0 + 8 * s);
                TT_CLK(5);
            }
            // ---- epilogue: sums -> K8 -> survivors ----
            tt_bar_wait(accd, groups & 1u);
            TT_CLK(6);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            for (uint32_t mt = (uint32_t)(warp >> 2); mt < mt_cnt; mt += PRODUCER_WARPS / 4) {
                const uint32_t bg = mt * 4u + (uint32_t)lq;  // this warp's block of the tile; TMEM lane = 32*lq + vector
                const bool have_blk = bg < nbg;
                const uint32_t li = (b0 + bg) * kBatch + (uint32_t)lane;
                bool valid = have_blk && li < nv;
                float f_add = 0.f, f_rescale = 0.f, f_error = 0.f;
                if (have_blk) {
                    const float* fac = reinterpret_cast<const float*>(lbase + (size_t)(b0 + bg) * B + (size_t)D * 4);
                    f_add = __ldg(fac + lane);
                    f_rescale = __ldg(fac + 32 + lane);
                    f_error = __ldg(fac + 64 + lane);
                }
                if (a.filter != nullptr && valid) {
                    const uint32_t id32 = (uint32_t)ix.ids[vbase + li];
                    valid = (unsigned long long)id32 < a.filter_nbits && ((a.filter[id32 >> 6] >> (id32 & 63u)) & 1ull);
                }
#pragma unroll 1
                for (int c0i = 0; c0i < NQ; c0i += 32) {
                    if ((uint32_t)c0i >= P) break;
                    uint32_t r[32];
                    const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + mt * NQ + (uint32_t)c0i;
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const uint32_t n = (uint32_t)(c0i + j);
                        if (n >= P) break;  // uniform: unused pair slots hold garbage sums
                        uint32_t accu = r[j];
                        if (WIDE) accu &= 0xffffu;  // the reference accumulates in wrapping u16
                        const float4 k0 = mi->c0[n], k1 = mi->c1[n];
                        // K8 (AVX2 variant): ip = fmadd(delta, accu, sum_vl); est = (f_add+g_add) + f_rescale*(ip+k1x)
                        const float ip = __fmaf_rn(k0.x, (float)accu, k0.y);
                        const float t1 = ip + k0.z;
                        const float t2 = f_rescale * t1;
                        const float t3 = f_add + k0.w;
                        const float est = t3 + t2;
                        const float t4 = f_error * k1.x;
                        float lower = est - t4;
                        if (!isfinite(lower)) lower = k1.z;
                        if (valid && lower < k1.y) {
                            const Survivor sv = Survivor{mi->rank[n], li, lower, a.has_ex ? ip : est};
                            const uint32_t q = mi->q[n];
                            const uint32_t slot = atomicAdd(&mi->surv_n, 1u);
                            if (slot < (uint32_t)SURV_CAP) {
                                mi->sq[slot] = q;
                                mi->ss[slot] = sv;
                            } else {  // staging full: straight to the query's buffer
                                const uint32_t gs = atomicAdd(&a.surv_cnt[q], 1u);
                                if (gs < a.surv_cap) a.surv[(size_t)q * a.surv_cap + gs] = sv;
                                if (a.stats) atomicAdd(&a.stats->survivors, 1ull);
                            }
                        }
                    }
                }
            }
            // TMEM reads done before this thread's next arrival on a `full` barrier lets the issuer overwrite the accumulators
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            groups += 1;
            TT_CLK(7);
        }
    }
#ifdef RBQ_TT_PROBE
    if (tid == 0) {
        pr_[10] = (unsigned long long)(clock64() - t_begin_);
        for (int i = 0; i < 13; ++i) atomicAdd(&g_tt_probe[i], pr_[i]);
    }
#endif
    if (!issuer) flush_survivors();
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
    }
}

bool tail_tc_supported(const DevIndex& ix) {
    static const bool off = [] {
        const char* e = getenv("RBQ_TAIL_PRMT");
        return e != nullptr && atoi(e) != 0;
    }();
    return !off && ix.D % 16 == 0;
}

template <bool WIDE, int PD, bool FULLK>
static int launch_tail_tc_k(const DevIndex& ix, const TailArgs& a, int sms, cudaStream_t st) {
    const size_t smem = tt::smem_bytes(PD);
    RBQ_CUDA(cudaFuncSetAttribute(tail_tc_kernel<WIDE, PD, FULLK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#ifdef RBQ_TT_PROBE
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_tt_probe, z, sizeof(z));
    const int variant = getenv("RBQ_TT_VARIANT") ? atoi(getenv("RBQ_TT_VARIANT")) : 0;
    cudaMemcpyToSymbol(g_tt_variant, &variant, sizeof(variant));
#endif
    tail_tc_kernel<WIDE, PD, FULLK><<<tt::CTAS_PER_SM * sms, tt::THREADS, smem, st>>>(ix, a);
    RBQ_CUDA(cudaGetLastError());
#ifdef RBQ_TT_PROBE
    cudaStreamSynchronize(st);
    cudaMemcpyFromSymbol(z, g_tt_probe, sizeof(z));
    const double nc = (double)(tt::CTAS_PER_SM * sms);
    fprintf(stderr, "[tt probe] per CTA: setup %.0f wait_empty %.0f cpasync %.0f build %.0f st %.0f arrive %.0f wait_acc %.0f epilogue %.0f | chunks %.0f items %.1f total %.0f clk\n",
            z[0] / nc, z[1] / nc, z[2] / nc, z[3] / nc, z[4] / nc, z[5] / nc, z[6] / nc, z[7] / nc, z[8] / nc, z[9] / nc, z[10] / nc);
    fprintf(stderr, "[tt probe]   prefetch issue %.0f, syncwarp %.0f\n", z[11] / nc, z[12] / nc);
#endif
    return RBQ_OK;
}

template <bool WIDE, int PD>
static int launch_tail_tc_ex(const DevIndex& ix, const TailArgs& a, int sms, cudaStream_t st) {
    return ix.D % 32 == 0 ? launch_tail_tc_k<WIDE, PD, true>(ix, a, sms, st) : launch_tail_tc_k<WIDE, PD, false>(ix, a, sms, st);
}

int launch_tail_tc(const DevIndex& ix, const TailArgs& a, cudaStream_t st) {
    int dev = 0, sms = 0;
    RBQ_CUDA(cudaGetDevice(&dev));
    RBQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static const int pd = [] {
        const char* e = getenv("RBQ_TAIL_PD");
        return e ? atoi(e) : 2;
    }();
    if (ix.D > 1024) return launch_tail_tc_ex<true, 2>(ix, a, sms, st);
    switch (pd) {
        case 3: return launch_tail_tc_ex<false, 3>(ix, a, sms, st);
        case 4: return launch_tail_tc_ex<false, 4>(ix, a, sms, st);
        default: return launch_tail_tc_ex<false, 2>(ix, a, sms, st);
    }
}

}  // namespace rbq
