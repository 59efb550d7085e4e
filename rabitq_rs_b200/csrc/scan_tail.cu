// scan_tail.cu -- list-major FastScan over the tail of every query's probe sequence.
//
// The reference walks a query's probed lists nearest-first and prunes with the running k-th distance
// (search_cluster_v2_batched, reference src/ivf.rs:2013-2127).  That threshold never increases, so once the
// head pass (scan.cu, kScanHead) has walked enough lists to fill the heap, tau = its k-th distance bounds every
// threshold the reference will use later: a tail vector can only be admitted if lower_bound < tau.  The tail
// stage therefore evaluates K7+K8 (simd::accumulate_batch_avx2 src/simd.rs:972-1184, compute_batch_distances_u16
// :2090-2140) for all remaining (query, list) pairs in ANY order, keeps the few candidates with
// lower_bound < tau ("survivors"), and the replay pass (scan.cu, kScanReplay) visits them in the reference's
// order against the live threshold -- the same decisions, bit for bit.
//
// Order freedom is what makes the stage fast: pairs are grouped by list, a CTA stages a list segment in shared
// memory ONCE (expanded into ready-made PRMT selectors and blend masks, work that depends on the codes only)
// and every warp then runs one query at a time over it:
//   * lane l owns codebooks l, l+32, ... with their 16 LUT bytes in registers (as in scan.cu);
//   * per 8 lookups: 4 PRMT (entries 0-7 / 8-15 for two selector halves) + 2 LOP3 blends + 1 shift;
//   * the looked-up bytes are summed on the tensor pipe: one mma.sync.m16n8k32.u8 (SASS IMMA.16832.U8.U8) against
//     a constant one-hot A matrix adds the two result registers of a lane over 4 lanes per byte position, and the
//     accumulator fragment carries the sum over codebooks -- no per-byte adds on the ALU/FMA pipes, and the
//     end-of-block cross-lane reduction shrinks to 8 shuffles;
//   * integer sums are exact, K8 is evaluated in the reference's (AVX2 variant) operation order.
#include <algorithm>
#include <cstdlib>

#include "scan_common.cuh"

namespace rbq {

constexpr int kTailWarps = 8;
constexpr uint32_t kPairsPerItem = 64;
constexpr uint32_t kMaxSegBlocks = 16;
constexpr int kSurvBuf = 32;  // per-warp survivor staging (flushed with one atomic when a block could overflow it)

// ---- grouping the (query, rank) pairs of the tail by list ------------------------------------------
__global__ void tail_count_kernel(const Probe* __restrict__ probes, const uint32_t* __restrict__ tail_start, uint32_t nq,
                                  uint32_t nprobe, uint32_t* __restrict__ list_cnt) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nq * nprobe) return;
    const uint32_t q = (uint32_t)(idx / nprobe), rank = (uint32_t)(idx % nprobe);
    if (rank < tail_start[q]) return;
    const Probe& p = probes[idx];
    if (p.nv != 0) atomicAdd(&list_cnt[p.cid], 1u);
}

// Work planning: exclusive scans of the per-list pair counts and item counts, then the work items.  Items of long lists (more
// than one accumulator group of `big_blocks` blocks) come first: they take two or more passes over the K dimension in the
// tensor-core kernel, and starting them early shortens the under-filled end of the persistent grid.
// Two launches of 1024-thread CTAs, one list per thread: (1) per-CTA scans + CTA totals, (2) every CTA re-scans the (<= 1024)
// CTA totals for its own base and writes its lists' offsets and items.  (The first version walked all lists in ONE CTA:
// 0.36 ms at 65536 lists, a cost every shard of a multi-GPU search paid in full.)
struct PlanRec {
    uint32_t pairs, items, big;
};
__device__ __forceinline__ PlanRec plan_block_scan(PlanRec v, PlanRec* s_warp, PlanRec& total) {
    // inclusive scan of (pairs, items, big) over the 1024 threads of the CTA; returns the inclusive value, total = CTA sum
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1) {
        const uint32_t ap = __shfl_up_sync(0xffffffffu, v.pairs, o), ai = __shfl_up_sync(0xffffffffu, v.items, o),
                       ab = __shfl_up_sync(0xffffffffu, v.big, o);
        if (lane >= o) {
            v.pairs += ap;
            v.items += ai;
            v.big += ab;
        }
    }
    if (lane == 31) s_warp[wid] = v;
    __syncthreads();
    if (wid == 0) {
        PlanRec w = s_warp[lane];
#pragma unroll
        for (uint32_t o = 1; o < 32; o <<= 1) {
            const uint32_t ap = __shfl_up_sync(0xffffffffu, w.pairs, o), ai = __shfl_up_sync(0xffffffffu, w.items, o),
                           ab = __shfl_up_sync(0xffffffffu, w.big, o);
            if (lane >= o) {
                w.pairs += ap;
                w.items += ai;
                w.big += ab;
            }
        }
        s_warp[32 + lane] = w;  // inclusive totals of warps 0..lane
    }
    __syncthreads();
    if (wid) {
        const PlanRec b = s_warp[32 + wid - 1];
        v.pairs += b.pairs;
        v.items += b.items;
        v.big += b.big;
    }
    total = s_warp[63];
    __syncthreads();
    return v;
}
__global__ void __launch_bounds__(1024) tail_plan_local_kernel(const uint32_t* __restrict__ list_cnt, const uint32_t* __restrict__ list_n,
                                                              uint32_t nlist, uint32_t per_item, uint32_t big_blocks, PlanRec* __restrict__ cta_tot) {
    __shared__ PlanRec s_warp[64];
    const uint32_t c = blockIdx.x * 1024u + threadIdx.x;
    PlanRec v{0u, 0u, 0u};
    if (c < nlist) {
        const uint32_t n = list_cnt[c], k = (n + per_item - 1) / per_item;
        v.pairs = n;
        v.items = k;
        v.big = (list_n[c] + kBatch - 1) / kBatch > big_blocks ? k : 0u;
    }
    PlanRec total;
    plan_block_scan(v, s_warp, total);
    if (threadIdx.x == 0) cta_tot[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) tail_plan_kernel(const uint32_t* __restrict__ list_cnt, const uint32_t* __restrict__ list_n,
                                                         uint32_t nlist, uint32_t per_item, uint32_t big_blocks, const PlanRec* __restrict__ cta_tot,
                                                         uint32_t* __restrict__ list_off, TailItem* __restrict__ items, uint32_t max_items,
                                                         uint32_t* __restrict__ counters) {
    __shared__ PlanRec s_warp[64];
    __shared__ PlanRec s_base, s_all;
    const uint32_t t = threadIdx.x, nctas = gridDim.x;
    // base of this CTA = sum of the totals of the CTAs before it; grand totals = sum of all (nctas <= 1024)
    PlanRec mine{0u, 0u, 0u};
    if (t < nctas) mine = cta_tot[t];
    PlanRec all;
    const PlanRec inc = plan_block_scan(mine, s_warp, all);
    if (t == blockIdx.x) s_base = PlanRec{inc.pairs - mine.pairs, inc.items - mine.items, inc.big - mine.big};
    if (t == 0) s_all = all;
    __syncthreads();
    const PlanRec base = s_base, grand = s_all;
    const uint32_t c = blockIdx.x * 1024u + t;
    PlanRec v{0u, 0u, 0u};
    uint32_t n = 0;
    bool big = false;
    if (c < nlist) {
        n = list_cnt[c];
        const uint32_t k = (n + per_item - 1) / per_item;
        big = (list_n[c] + kBatch - 1) / kBatch > big_blocks;
        v = PlanRec{n, k, big ? k : 0u};
    }
    PlanRec total;
    const PlanRec loc = plan_block_scan(v, s_warp, total);
    if (c < nlist) {
        const uint32_t po = base.pairs + loc.pairs - v.pairs;                       // first pair slot of the list
        const uint32_t before_items = base.items + loc.items - v.items, before_big = base.big + loc.big - v.big;
        uint32_t io = big ? before_big : grand.big + (before_items - before_big);  // next slot among the big / the other items
        list_off[c] = po;
        if (n) {
            const uint32_t chunks = (n + per_item - 1) / per_item, sz = (n + chunks - 1) / chunks;  // balanced chunks
            for (uint32_t j = 0; j < chunks; ++j, ++io) {
                const uint32_t b = j * sz, e = min(n, b + sz);
                if (io < max_items && b < e) items[io] = TailItem{c, po + b, e - b, 0u};
            }
        }
    }
    if (blockIdx.x == 0 && t == 0) {
        list_off[nlist] = grand.pairs;
        counters[0] = min(grand.items, max_items);
        counters[1] = 0u;
    }
}

__global__ void tail_scatter_kernel(const Probe* __restrict__ probes, const uint32_t* __restrict__ tail_start, uint32_t nq,
                                    uint32_t nprobe, const uint32_t* __restrict__ list_off, uint32_t* __restrict__ list_fill,
                                    uint32_t* __restrict__ pairs) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)nq * nprobe) return;
    const uint32_t q = (uint32_t)(idx / nprobe), rank = (uint32_t)(idx % nprobe);
    if (rank < tail_start[q]) return;
    const Probe& p = probes[idx];
    if (p.nv == 0) return;
    const uint32_t pos = list_off[p.cid] + atomicAdd(&list_fill[p.cid], 1u);
    pairs[pos] = (uint32_t)idx;
}

// ---- the tail kernel ----------------------------------------------------------------------------------

__device__ __forceinline__ void imma_u8(int (&c)[4], uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
    // C[16x8] += A[16x32] * B[32x8], u8 x u8 -> s32.  a1 = a3 = 0 (rows 8..15 unused).
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

// 8 lookups of one codebook: selector word s (two 16-bit halves, bit 3 of every nibble cleared), byte masks m0/m1
// (0xFF where the code nibble's msb is set), LUT row T.  The two result registers go to the tensor pipe.
__device__ __forceinline__ void lookup8(int (&acc)[4], const uint4& T, uint32_t s, uint32_t m0, uint32_t m1, uint32_t a0,
                                        uint32_t a2) {
    const uint32_t lo0 = prmt(T.x, T.y, s), hi0 = prmt(T.z, T.w, s);
    const uint32_t s1 = prmt(s, 0u, 0x4432u);  // s >> 16 on the byte-permute path (shifts issue at a quarter of its rate)
    const uint32_t lo1 = prmt(T.x, T.y, s1), hi1 = prmt(T.z, T.w, s1);
    const uint32_t r0 = (lo0 & ~m0) | (hi0 & m0), r1 = (lo1 & ~m1) | (hi1 & m1);
    imma_u8(acc, a0, a2, r0, r1);
}

template <int NCB, bool WIDE>
__global__ void __launch_bounds__(kTailWarps * 32, 2) tail_kernel(DevIndex ix, TailArgs a) {
    extern __shared__ __align__(128) unsigned char tail_smem[];
    __shared__ uint32_t s_item;
    __shared__ Survivor s_surv[kTailWarps][kSurvBuf];  // survivors of the pair a warp is working on
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = ix.D, ncb = D / 4;
    const uint32_t B = ix.block_stride;
    const uint32_t EB = 12u * (uint32_t)D + 384u;  // expanded block: selectors 4D | masks 8D | factors 384
    const uint32_t S = a.seg_blocks;
    const uint32_t sm_u32 = smem_u32(tail_smem);
    const bool l2 = ix.metric == RBQ_METRIC_L2;

    // constant A fragment: row r < 4 picks byte r of b0, row 4 + r picks byte r of b1 (see lookup8)
    const int g = lane >> 2, t4 = lane & 3;
    const uint32_t a0 = g < 4 ? (1u << (8 * g)) : 0u, a2 = g >= 4 ? (1u << (8 * (g - 4))) : 0u;
    // after the reduction lane (g, t4) holds accu of vector vmap: register k = t4, half h = g >> 2, byte g & 3
    const int vmap = 2 * t4 + (g >> 2) + ((g & 1) ? 16 : 0) + ((g & 2) ? 8 : 0);
    unsigned long long st_surv = 0;
    int nbuf = 0;  // entries waiting in s_surv[warp]
    // one atomic per flush instead of one per block: the slot base is a global round trip
    auto flush_survivors = [&](uint32_t q) {
        if (nbuf == 0) return;
        __syncwarp();
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&a.surv_cnt[q], (uint32_t)nbuf);
        base = __shfl_sync(0xffffffffu, base, 0);
        for (int i = lane; i < nbuf; i += 32)
            if (base + (uint32_t)i < a.surv_cap) a.surv[(size_t)q * a.surv_cap + base + i] = s_surv[warp][i];
        st_surv += (unsigned long long)nbuf;
        nbuf = 0;
        __syncwarp();
    };

    for (;;) {
        __syncthreads();  // everyone is done with s_item and the staged segment
        if (tid == 0) s_item = atomicAdd(&a.counters[1], 1u);
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= a.counters[0]) break;
        const TailItem it = a.items[item];
        const uint32_t nv = ix.list_n[it.cid], nb = (nv + kBatch - 1) / kBatch;
        const uint8_t* lbase = ix.blocks + (size_t)ix.blk_off[it.cid] * B;
        const unsigned long long vbase = ix.vec_off[it.cid];
        if (tid == 0 && a.stats) {
            atomicAdd(&a.stats->tail_blocks, (unsigned long long)it.pair_count * nb);
            atomicAdd(&a.stats->tail_pairs, (unsigned long long)it.pair_count);
            atomicAdd(&a.stats->candidates, (unsigned long long)it.pair_count * nv);
        }
        for (uint32_t seg0 = 0; seg0 < nb; seg0 += S) {
            const uint32_t ns = min(S, nb - seg0);
            if (seg0) __syncthreads();  // previous segment fully consumed
            // ---- stage + expand (codes-only work, shared by every query of the item) ----
            for (uint32_t b = warp; b < ns; b += kTailWarps) {
                const uint8_t* src = lbase + (size_t)(seg0 + b) * B;
                const uint32_t dst = sm_u32 + b * EB;
                for (int cb = lane; cb < ncb; cb += 32) {
                    const uint4 c = ldg128(src + 16 * cb);
                    const uint32_t C[4] = {c.x, c.y, c.z, c.w};
                    uint32_t M[8];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint32_t sh = C[k] << 4;  // brings the low nibbles' msb into byte-sign position
                        M[2 * k] = prmt(C[k], sh, 0x9D8Cu);
                        M[2 * k + 1] = prmt(C[k], sh, 0xBFAEu);
                    }
                    sts128(dst + 16u * cb, C[0] & 0x77777777u, C[1] & 0x77777777u, C[2] & 0x77777777u, C[3] & 0x77777777u);
                    sts128(dst + 4u * D + 16u * cb, M[0], M[1], M[2], M[3]);
                    sts128(dst + 8u * D + 16u * cb, M[4], M[5], M[6], M[7]);
                }
                for (int j = lane; j < 96; j += 32)
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(dst + 12u * D + 4u * j), "r"(ldg32(src + 4 * D + 4 * j)) : "memory");
            }
            __syncthreads();
            // ---- every warp: one (query, rank) pair at a time over the staged blocks ----
            for (uint32_t pi = warp; pi < it.pair_count; pi += kTailWarps) {
                const uint32_t pid = a.pairs[it.pair_begin + pi];
                const uint32_t q = pid / a.nprobe, rank = pid - q * a.nprobe;
                if (pi + kTailWarps < it.pair_count) {  // pull the next pair's LUT towards L1 while this one is scanned
                    const uint32_t qn = a.pairs[it.pair_begin + pi + kTailWarps] / a.nprobe;
                    const uint8_t* nl = a.lut + (size_t)qn * D * 4;
                    for (int o = lane * 128; o < D * 4; o += 32 * 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(nl + o));
                }
                uint4 T[NCB];
#pragma unroll
                for (int i = 0; i < NCB; ++i) {
                    const int cb = lane + 32 * i;
                    T[i] = (cb < ncb) ? ldg128(a.lut + (size_t)q * D * 4 + 16 * cb) : make_uint4(0, 0, 0, 0);
                }
                const QueryScalars s = a.qs[q];
                const Probe p = a.probes[pid];
                const float tau = a.tau[q];
                for (uint32_t b = 0; b < ns; ++b) {
                    const uint32_t blk = sm_u32 + b * EB;
                    int acc[4][4];
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[k][j] = 0;
#pragma unroll
                    for (int i = 0; i < NCB; ++i) {
                        const uint32_t cb = (uint32_t)min(lane + 32 * i, ncb - 1);  // padding lanes: zero LUT row
                        const uint4 Sx = lds128(blk + 16u * cb);
                        const uint4 Ma = lds128(blk + 4u * D + 16u * cb);
                        const uint4 Mb = lds128(blk + 8u * D + 16u * cb);
                        lookup8(acc[0], T[i], Sx.x, Ma.x, Ma.y, a0, a2);
                        lookup8(acc[1], T[i], Sx.y, Ma.z, Ma.w, a0, a2);
                        lookup8(acc[2], T[i], Sx.z, Mb.x, Mb.y, a0, a2);
                        lookup8(acc[3], T[i], Sx.w, Mb.z, Mb.w, a0, a2);
                    }
                    // fragment (row g, cols 2*t4, 2*t4+1): sum the 8 columns (= lane groups) of row g
                    int tot[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        int x = acc[k][0] + acc[k][1];
                        x += __shfl_xor_sync(0xffffffffu, x, 1);
                        x += __shfl_xor_sync(0xffffffffu, x, 2);
                        tot[k] = x;
                    }
                    uint32_t accu = (uint32_t)(t4 == 0 ? tot[0] : t4 == 1 ? tot[1] : t4 == 2 ? tot[2] : tot[3]);
                    if (WIDE) accu &= 0xffffu;  // the reference accumulates in wrapping u16
                    const uint32_t fac = blk + 12u * D + 4u * (uint32_t)vmap;
                    const float f_add = lds_f32(fac), f_rescale = lds_f32(fac + 128u), f_error = lds_f32(fac + 256u);
                    // K8 (AVX2 variant): ip = fmadd(delta, accu, sum_vl); est = (f_add+g_add) + f_rescale*(ip+k1x)
                    const float ip = __fmaf_rn(s.delta, (float)accu, s.sum_vl);
                    const float t1 = ip + s.k1x;
                    const float t2 = f_rescale * t1;
                    const float t3 = f_add + p.g_add;
                    const float est = t3 + t2;
                    const float t4f = f_error * p.g_error;
                    float lower = est - t4f;
                    const uint32_t li = (seg0 + b) * kBatch + (uint32_t)vmap;
                    bool valid = li < nv;
                    if (a.filter != nullptr && valid) {
                        const uint32_t id32 = (uint32_t)ix.ids[vbase + li];
                        valid = (unsigned long long)id32 < a.filter_nbits && ((a.filter[id32 >> 6] >> (id32 & 63u)) & 1ull);
                    }
                    if (!isfinite(lower)) lower = l2 ? 0.0f : -(p.dot_qc + s.qnorm);
                    const bool cand = valid && (lower < tau);
                    const unsigned mask = __ballot_sync(0xffffffffu, cand);
                    if (mask != 0u) {
                        if (nbuf + __popc(mask) > kSurvBuf) flush_survivors(q);
                        if (cand) s_surv[warp][nbuf + __popc(mask & ((1u << lane) - 1u))] = Survivor{rank, li, lower, a.has_ex ? ip : est};
                        nbuf += __popc(mask);
                    }
                }
                flush_survivors(q);
            }
        }
    }
    if (lane == 0 && a.stats && st_surv) atomicAdd(&a.stats->survivors, st_surv);
}

// limits of the CURRENT device, refreshed by every launcher (a process may drive several GPUs from several threads)
static thread_local int g_tail_sms = 0;
static thread_local size_t g_tail_smem_optin = 0;
static int tail_limits() {
    int dev = 0, v = 0;
    RBQ_CUDA(cudaGetDevice(&dev));
    RBQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    g_tail_sms = v;
    RBQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    g_tail_smem_optin = (size_t)v;
    return RBQ_OK;
}

static uint32_t g_surv_cap_override = 0;  // test knob: forces survivor-buffer overflows
void tail_debug_set_survivor_cap(uint32_t cap) { g_surv_cap_override = cap; }
// survivors per query the lazy replay sorts in shared memory (10-bit slot in its sort keys)
static uint32_t tail_sort_cap(size_t top_k) {
    if (g_surv_cap_override) return g_surv_cap_override;
    (void)top_k;
    return 1024;
}
// slots per query in the survivor buffer: queries between the two caps are replayed by the overflow tier (one CTA per query,
// resolve.cu), queries beyond it by a sequential re-walk of their tail
static uint32_t tail_surv_cap(size_t top_k) { return 4u * tail_sort_cap(top_k); }

// dense head buffer: room for the longest list of the shard (lists beyond 64Ki vectors send their queries to the sequential
// fallback), and as many query rows as fit 512 MiB -- the head pass runs over row-sized sub-chunks of the batch
static uint32_t tail_head_cap(const DevIndex& ix) {
    const size_t want = ((size_t)ix.max_list_n + 31) / 32 * 32;
    return (uint32_t)std::max<size_t>(32, std::min<size_t>(want, 65536));
}
static uint32_t tail_head_rows(const DevIndex& ix, size_t nq) {
    const size_t afford = ((size_t)512 << 20) / 8 / tail_head_cap(ix);
    return (uint32_t)std::max<size_t>(1, std::min<size_t>(nq, std::max<size_t>(afford, 1024)));
}

size_t tail_ws_bytes(const DevIndex& ix, size_t nq, size_t nprobe, size_t top_k) {
    const size_t cap = tail_surv_cap(top_k), max_items = nq * nprobe / kPairsPerItem + ix.nlist + 1;
    size_t n = 4096;
    n += nq * 4 + 256;                               // tail_start
    n += nq * 4 + 256;                               // tau
    n += (nq + 2 * (size_t)ix.nlist + kTailCounters + kHeadCursors) * 4 + 256;  // surv_cnt | list_cnt | list_fill | counters (zeroed together)
    n += ((size_t)ix.nlist + 1) * 4 + 256;           // list_off
    n += 1024 * 12 + 256;                            // plan_tot
    n += nq * nprobe * 4 + 256;                      // pairs
    n += max_items * sizeof(TailItem) + 256;
    n += nq * cap * sizeof(Survivor) + 256;
    n += (size_t)tail_head_rows(ix, nq) * tail_head_cap(ix) * 8 + 256;  // head_buf
    n += 2 * (nq * 4 + 256);                             // fb_list, fb2_list
    n += nq * 4 + 256;                                   // ovf_list
    n += (size_t)kOvfCtas * 4096 * 24 + 256;              // ovf_recs
    n += nq * 4 + 512;                                   // qlist + qcount
    return n;
}

void tail_ws_carve(const DevIndex& ix, size_t nq, size_t nprobe, size_t top_k, char* base, TailWs& tw) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        off = (off + 255) & ~(size_t)255;
        char* r = base + off;
        off += bytes;
        return r;
    };
    tw.surv_cap = tail_surv_cap(top_k);
    tw.sort_cap = tail_sort_cap(top_k);
    tw.max_items = (uint32_t)(nq * nprobe / kPairsPerItem + ix.nlist + 1);
    tw.pairs_per_item = kPairsPerItem;
    tw.tail_start = reinterpret_cast<uint32_t*>(take(nq * 4));
    tw.tau = reinterpret_cast<float*>(take(nq * 4));
    uint32_t* z = reinterpret_cast<uint32_t*>(take((nq + 2 * (size_t)ix.nlist + kTailCounters + kHeadCursors) * 4));
    tw.surv_cnt = z;
    tw.list_cnt = z + nq;
    tw.list_fill = tw.list_cnt + ix.nlist;
    tw.counters = tw.list_fill + ix.nlist;
    tw.list_off = reinterpret_cast<uint32_t*>(take(((size_t)ix.nlist + 1) * 4));
    tw.plan_tot = take(1024 * 12);
    tw.pairs = reinterpret_cast<uint32_t*>(take(nq * nprobe * 4));
    tw.items = reinterpret_cast<TailItem*>(take((size_t)tw.max_items * sizeof(TailItem)));
    tw.surv = reinterpret_cast<Survivor*>(take(nq * (size_t)tw.surv_cap * sizeof(Survivor)));
    tw.head_cap = tail_head_cap(ix);
    tw.head_rows = tail_head_rows(ix, nq);
    tw.head_buf = reinterpret_cast<float2*>(take((size_t)tw.head_rows * tw.head_cap * 8));
    tw.fb_list = reinterpret_cast<uint32_t*>(take(nq * 4));
    tw.fb2_list = reinterpret_cast<uint32_t*>(take(nq * 4));
    tw.ovf_list = reinterpret_cast<uint32_t*>(take(nq * 4));
    tw.ovf_recs = take((size_t)kOvfCtas * 4096 * 24);
    tw.qlist = reinterpret_cast<uint32_t*>(take(nq * 4));
    tw.qcount = reinterpret_cast<uint32_t*>(take(16));
}

template <int NCB, bool WIDE>
static int launch_tail_ex(const DevIndex& ix, TailArgs& a, cudaStream_t st) {
    const uint32_t EB = 12u * (uint32_t)ix.D + 384u;
    // two CTAs per SM: each may use half of the shared memory (minus the per-CTA reservation)
    const size_t budget = std::min<size_t>(g_tail_smem_optin, (227 * 1024 - 2 * 1024) / 2 - 256) - sizeof(Survivor) * kSurvBuf * kTailWarps;
    const uint32_t S = (uint32_t)std::min<size_t>(kMaxSegBlocks, std::max<size_t>(1, budget / EB));
    const size_t smem = (size_t)S * EB;
    if (smem > g_tail_smem_optin) return fail(RBQ_INVALID_CONFIG, "tail kernel shared memory exceeds the device limit");
    a.seg_blocks = S;
    RBQ_CUDA(cudaFuncSetAttribute(tail_kernel<NCB, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tail_kernel<NCB, WIDE><<<2 * g_tail_sms, kTailWarps * 32, smem, st>>>(ix, a);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

static int launch_tail_fastscan(const DevIndex& ix, TailArgs& a, cudaStream_t st) {
    if (tail_tc_supported(ix)) return launch_tail_tc(ix, a, st);
    const int ncb_lane = (ix.D / 4 + 31) / 32;
    if (ix.D > 1024) {
        if (ncb_lane <= 12) return launch_tail_ex<12, true>(ix, a, st);
        return launch_tail_ex<16, true>(ix, a, st);
    }
    switch (ncb_lane) {
        case 1: return launch_tail_ex<1, false>(ix, a, st);
        case 2: return launch_tail_ex<2, false>(ix, a, st);
        case 3: return launch_tail_ex<3, false>(ix, a, st);
        case 4: return launch_tail_ex<4, false>(ix, a, st);
        case 5:
        case 6: return launch_tail_ex<6, false>(ix, a, st);
        default: return launch_tail_ex<8, false>(ix, a, st);
    }
}

int launch_tail(const DevIndex& ix, const uint8_t* d_lut, const QueryScalars* d_qs, const Probe* d_probes, size_t nq,
                size_t nprobe, const uint64_t* d_filter, size_t filter_nbits, DevStats* d_stats, const TailWs& tw,
                cudaStream_t st, uint64_t* launches, cudaEvent_t ev_begin, cudaEvent_t ev_end) {
    if (nq == 0) return RBQ_OK;
    int rc = tail_limits();
    if (rc) return rc;
    const size_t npairs = nq * nprobe;
    const unsigned tb = 256, gb = (unsigned)((npairs + tb - 1) / tb);
    tail_count_kernel<<<gb, tb, 0, st>>>(d_probes, tw.tail_start, (uint32_t)nq, (uint32_t)nprobe, tw.list_cnt);
    const unsigned plan_ctas = (ix.nlist + 1023u) / 1024u;
    if (plan_ctas > 1024u) return fail(RBQ_INVALID_CONFIG, "more than 2^20 inverted lists are not supported by the tail planner");
    PlanRec* plan_tot = reinterpret_cast<PlanRec*>(tw.plan_tot);
    tail_plan_local_kernel<<<plan_ctas, 1024, 0, st>>>(tw.list_cnt, ix.list_n, ix.nlist, tw.pairs_per_item, 8u, plan_tot);
    tail_plan_kernel<<<plan_ctas, 1024, 0, st>>>(tw.list_cnt, ix.list_n, ix.nlist, tw.pairs_per_item, 8u, plan_tot, tw.list_off, tw.items,
                                                 tw.max_items, tw.counters);
    tail_scatter_kernel<<<gb, tb, 0, st>>>(d_probes, tw.tail_start, (uint32_t)nq, (uint32_t)nprobe, tw.list_off, tw.list_fill,
                                           tw.pairs);
    RBQ_CUDA(cudaGetLastError());
    TailArgs a;
    a.lut = d_lut;
    a.qs = d_qs;
    a.probes = d_probes;
    a.nprobe = (uint32_t)nprobe;
    a.tau = tw.tau;
    a.pairs = tw.pairs;
    a.items = tw.items;
    a.counters = tw.counters;
    a.surv = tw.surv;
    a.surv_cnt = tw.surv_cnt;
    a.surv_cap = tw.surv_cap;
    a.filter = reinterpret_cast<const unsigned long long*>(d_filter);
    a.filter_nbits = filter_nbits;
    a.stats = d_stats;
    a.seg_blocks = 1;
    a.has_ex = ix.ex_bits != 0;
    if (launches) *launches += 5;
    if (ev_begin) cudaEventRecord(ev_begin, st);
    rc = launch_tail_fastscan(ix, a, st);
    if (ev_end) cudaEventRecord(ev_end, st);
    return rc;
}

}  // namespace rbq
