// scan.cu -- the hot path: one warp walks one query's probed lists in the reference's order.
//
//   K7   FastScan accumulate   simd::accumulate_batch_avx2 / _scalar   (reference src/simd.rs:972-1184, 1462-1525)
//   K8   batch distances       simd::compute_batch_distances_u16 (AVX2) (reference src/simd.rs:2090-2140)
//   K9   prune + refine        search_cluster_v2_batched                (reference src/ivf.rs:2013-2127)
//   K10  packed ex-code dot    ip_packed_ex2_f32 / ip_packed_ex6_f32 (AVX2 lane order) (src/simd.rs:1722-1825)
//   K11  top-k                 BinaryHeap<HeapEntry> keep-k-smallest    (reference src/ivf.rs:1844, 2116-2126)
//
// Layout consumed as stored in the index file: a 32-vector block = 4*D code
// bytes + f_add[32] + f_rescale[32] + f_error[32].  The 16 code bytes at offset 16*cb belong to
// codebook cb (dims 4cb..4cb+3), like the 16 LUT bytes at the same offset of the query's LUT.
//
// Kernel structure (persistent CTAs, warps fetch queries from a global counter):
//   * every warp owns a ring of NST block-sized shared-memory stages fed by 1-D TMA bulk copies
//     (cp.async.bulk + mbarrier complete_tx), so the next blocks of the query's probe sequence are in
//     flight while the current one is processed -- the stream never waits on a register load;
//   * lane l owns codebooks l, l+32, ...: its LUT rows stay in registers for the whole query, its
//     code bytes are one LDS.128 per codebook; the 16-entry byte lookup is two PRMTs (entries 0-7 /
//     8-15) blended by a PRMT-generated mask (the GPU analogue of pshufb); the four looked-up bytes of
//     a PRMT are accumulated with dp4a against one-hot selectors (FMA pipe, balancing the ALU pipe);
//   * per-lane partial sums are packed to u16 pairs and reduced across lanes with a 16-shuffle
//     reduce-scatter that leaves lane v holding accu[v]; factors are read from the same stage;
//   * pruning is exact: lanes whose lower bound beats a (possibly stale, hence looser)
//     threshold are queued; at the end of each list the queue is refined 4 candidates x 8 lanes at a
//     time (ex-codes staged through shared memory with 128-bit loads) and then replayed in visit
//     order against the live threshold -- the reference's sequential decisions, reproduced exactly.
//
// Integer sums are exact (== the reference's wrapping-u16 value); float ops follow the AVX2
// variants' order (fma only where the reference uses fmadd).  Compiled with -fmad=false.
#include <algorithm>

#include "scan_common.cuh"

namespace rbq {

constexpr int kWarps = 4;
constexpr int kMaxTopK = 1024;
size_t scan_max_topk() { return kMaxTopK; }

struct ScanArgs {
    const float* rot;
    const uint8_t* lut;
    const QueryScalars* qs;
    const Probe* probes;
    uint32_t nq, nprobe, top_k;
    const unsigned long long* filter;
    unsigned long long filter_nbits;
    unsigned long long* out_ids;
    float* out_scores;
    uint32_t* out_counts;
    DevStats* stats;
    unsigned int* work_counter;  // next query to hand out
    uint32_t nst;                // ring stages per warp
    uint32_t ex_stage_stride;    // bytes per refine staging slot (16-byte multiple)
    // head / replay passes of the list-major pipeline (see scan_tail.cu)
    uint32_t mode;               // ScanMode
    uint32_t* tail_start;
    float* tau;
    const Survivor* surv;
    const uint32_t* surv_cnt;
    uint32_t surv_cap;
    // kScanFallback: the queries the list-major fast path handed back (see resolve.cu)
    const uint32_t* fb_list;
    const uint32_t* fb_count;
};

// Per-warp shared memory carve-up (bytes), all offsets 16-byte aligned.
struct WarpSmem {
    uint32_t ring, bars, exst, rq, si, sd, ord, total;
};
__host__ __device__ inline WarpSmem warp_smem_layout(uint32_t block_stride, uint32_t nst, uint32_t ex_stage_stride,
                                                     uint32_t D, uint32_t k, bool has_ex, uint32_t surv_cap) {
    WarpSmem w;
    uint32_t o = 0;
    w.ring = o;
    o += nst * block_stride;
    w.bars = o;
    o += ((nst * 8 + 15) / 16) * 16;
    w.exst = o;
    o += has_ex ? kRefineSlots * ex_stage_stride : 0;
    w.rq = o;
    o += has_ex ? ((D * 4 + 15) / 16) * 16 : 0;
    w.si = o;
    o += ((k * 8 + 15) / 16) * 16;
    w.sd = o;
    o += ((k * 4 + 15) / 16) * 16;
    w.ord = o;  // replay pass: sort keys of the survivors (capacity rounded up to a power of two >= 32)
    uint32_t cap2 = surv_cap ? 32 : 0;
    while (cap2 < surv_cap) cap2 <<= 1;
    o += cap2 * 8;
    w.total = o;
    return w;
}

// Walks a query's block sequence: probes in order, blocks of a list in order; lists with no local
// vectors (empty, or owned by another shard) are skipped.
struct Cursor {
    uint32_t pi, b, nb;  // probe rank, block inside the list, blocks in the list
    const uint8_t* base; // first block of the current list
    __device__ __forceinline__ void seek(const DevIndex& ix, const Probe* pr, uint32_t nprobe) {
        while (pi < nprobe) {
            const uint32_t nv = pr[pi].nv;
            if (nv != 0) {
                nb = (nv + kBatch - 1) / kBatch;
                base = ix.blocks + (size_t)pr[pi].blk_off * ix.block_stride;
                b = 0;
                return;
            }
            ++pi;
        }
        nb = 0;
    }
    __device__ __forceinline__ bool valid(uint32_t nprobe) const { return pi < nprobe; }
    __device__ __forceinline__ void next(const DevIndex& ix, const Probe* pr, uint32_t nprobe) {
        if (++b >= nb) {
            ++pi;
            seek(ix, pr, nprobe);
        }
    }
};

template <int NCB, int EXK, bool WIDE>
__global__ void __launch_bounds__(kWarps * 32) scan_kernel(DevIndex ix, ScanArgs a) {
    extern __shared__ __align__(128) unsigned char scan_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = ix.D, ncb = D / 4, k = (int)a.top_k;
    const uint32_t B = ix.block_stride, NST = a.nst;
    const WarpSmem L = warp_smem_layout(B, NST, a.ex_stage_stride, D, k, EXK != 0, a.mode == kScanReplay ? a.surv_cap : 0);
    unsigned char* wbase = scan_smem + (size_t)warp * L.total;
    const uint32_t ring_u32 = smem_u32(wbase + L.ring), bars_u32 = smem_u32(wbase + L.bars);
    const uint32_t exst_u32 = smem_u32(wbase + L.exst), rq2_u32 = smem_u32(wbase + L.rq);
    float* rq2 = reinterpret_cast<float*>(wbase + L.rq);  // rotated query, (j, j+8) interleaved per 16 dims
    unsigned long long* si = reinterpret_cast<unsigned long long*>(wbase + L.si);
    float* sd = reinterpret_cast<float*>(wbase + L.sd);
    unsigned long long* ord = reinterpret_cast<unsigned long long*>(wbase + L.ord);  // replay: survivor sort keys
    const bool l2 = ix.metric == RBQ_METRIC_L2;

    if (lane == 0) {
        for (uint32_t s = 0; s < NST; ++s) mbar_init(bars_u32 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t iss_stage = 0, use_stage = 0, use_phase = 0;  // ring cursors (persist across queries)
    unsigned long long st_blocks = 0, st_cand = 0, st_ref = 0, st_adm = 0, st_ovf = 0;

    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(a.work_counter, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        bool resume = false;
        if (a.mode == kScanFallback) {
            if (q >= *a.fb_count) break;
            const uint32_t entry = a.fb_list[q];
            q = entry & ~kFbResume;
            resume = (entry & kFbResume) != 0u;
        }
        if (q >= a.nq) break;
        const Probe* pr = a.probes + (size_t)q * a.nprobe;

        // what this pass does for the query
        uint32_t start_pi = 0, n_surv = 0;
        bool walk = true;  // walk lists start_pi.. sequentially (full / head / overflowed replay)
        int cnt = 0;
        if (a.mode == kScanReplay) {
            start_pi = a.tail_start[q];
            n_surv = a.surv_cnt[q];
            if (start_pi >= a.nprobe || n_surv == 0) continue;  // the head result is already final
            walk = n_surv > a.surv_cap;                          // survivor buffer overflowed: re-walk the tail
            if (walk) st_ovf += 1;
        } else if (resume) {
            start_pi = a.tail_start[q];
        }
        if (a.mode == kScanReplay || resume) {
            // resume from the head pass' top-k (stored best-first in the output arrays)
            cnt = (int)a.out_counts[q];
            for (int i = lane; i < cnt; i += 32) {
                const float sc = a.out_scores[(size_t)q * k + i];
                sd[i] = l2 ? sc : -sc;
                si[i] = a.out_ids[(size_t)q * k + i];
            }
        }

        // producer side: prime the ring (all stages are free: everything issued so far was consumed)
        Cursor pc;
        pc.pi = start_pi;
        uint32_t inflight = 0;
        if (walk) {
            if (lane == 0) fence_proxy_async();
            pc.seek(ix, pr, a.nprobe);
            for (uint32_t s = 0; s < NST && pc.valid(a.nprobe); ++s) {
                if (lane == 0) {
                    mbar_expect_tx(bars_u32 + 8 * iss_stage, B);
                    tma_load_1d(ring_u32 + iss_stage * B, pc.base + (size_t)pc.b * B, B, bars_u32 + 8 * iss_stage);
                }
                if (++iss_stage == NST) iss_stage = 0;
                ++inflight;
                pc.next(ix, pr, a.nprobe);
            }
        }

        if (EXK != 0)
            for (int i = lane; i < D; i += 32)  // dim 16c + r -> float2 slot 8c + (r & 7), component r >> 3
                rq2[2 * (8 * (i >> 4) + (i & 7)) + ((i >> 3) & 1)] = a.rot[(size_t)q * D + i];
        uint4 T[NCB];
        if (walk) {
#pragma unroll
            for (int i = 0; i < NCB; ++i) {
                const int cb = lane + 32 * i;
                T[i] = (cb < ncb) ? ldg128(a.lut + (size_t)q * D * 4 + 16 * cb) : make_uint4(0, 0, 0, 0);
            }
        }
        const QueryScalars s = a.qs[q];
        __syncwarp();

        // candidate queue: slot i lives in lane i, in visit order
        int qn = 0;
        float q_lower = 0.0f, q_ip = 0.0f, q_gadd = 0.0f;
        unsigned long long q_gv = 0;  // global vector index (list's vec_off + position)

        // refine + replay everything queued (reference order, live threshold)
        auto flush = [&]() {
            if (qn == 0) return;
            float dist = 0.0f;
            const bool mine = lane < qn;
            float fae = 0.0f, fre = 0.0f;
            unsigned long long q_vid = 0;
            if (mine) {
                fae = __ldg(ix.f_add_ex + q_gv);
                fre = __ldg(ix.f_rescale_ex + q_gv);
                q_vid = ix.ids[q_gv];
            }
            float exdot = 0.0f;
            const int g = lane >> 3, j = lane & 7;
            for (int r0 = 0; r0 < qn; r0 += kRefineSlots) {
                const int c = r0 + g;  // candidate served by this 8-lane group
                const unsigned long long gv_c = __shfl_sync(0xffffffffu, q_gv, c & 31);
                const uint32_t stg = exst_u32 + (uint32_t)g * a.ex_stage_stride;
                if (c < qn) stage_expand<EXK>(ix.ex + gv_c * ix.ex_stride, stg, D, j, ix.ex_bits);
                __syncwarp();
                float part = 0.0f;
                if (c < qn) part = ex_dot_lane(stg, rq2_u32, D, j);
                part = hsum8(part);
                const float v = __shfl_sync(0xffffffffu, part, ((lane - r0) & 3) * 8);
                if (lane >= r0 && lane < r0 + kRefineSlots) exdot = v;
                __syncwarp();
            }
            st_ref += qn;
            if (mine) {
                // distance = f_add_ex + g_add + f_rescale_ex * (binary_scale*ip + ex_dot + kbx)  (ivf.rs:2095-2099)
                float tt = s.bscale * q_ip;
                tt = tt + exdot;
                tt = tt + s.kbx;
                const float mm2 = fre * tt;
                const float aa = fae + q_gadd;
                dist = aa + mm2;
            }
            for (int c = 0; c < qn; ++c) {
                const float lb_s = __shfl_sync(0xffffffffu, q_lower, c);
                const float d_s = __shfl_sync(0xffffffffu, dist, c);
                const unsigned long long id_s = __shfl_sync(0xffffffffu, q_vid, c);
                const float theta = cnt >= k ? sd[k - 1] : INFINITY;
                if (lb_s >= theta) continue;  // skipped_by_lower_bound
                st_adm += 1;
                if (!isfinite(d_s)) continue;
                topk_insert(sd, si, cnt, k, d_s, id_s, lane);
            }
            qn = 0;
        };
        // append the lanes of `mask` to the queue in lane order (queue slot qn + r takes the r-th set lane)
        auto enqueue = [&](unsigned mask, float lower, float ipv, unsigned long long gv, float gadd) {
            const int n_new = __popc(mask);
            if (qn + n_new > 32) flush();
            const int r = lane - qn;
            const int src = (r >= 0 && r < n_new) ? (int)__fns(mask, 0, r + 1) : 0;
            const float nl = __shfl_sync(0xffffffffu, lower, src);
            const float nip = __shfl_sync(0xffffffffu, ipv, src);
            const unsigned long long ngv = __shfl_sync(0xffffffffu, gv, src);
            const float nga = __shfl_sync(0xffffffffu, gadd, src);
            if (r >= 0 && r < n_new) {
                q_lower = nl;
                q_ip = nip;
                q_gv = ngv;
                q_gadd = nga;
                // warm L2 with the candidate's ex-code while the scan goes on
                const uint8_t* ep = ix.ex + q_gv * ix.ex_stride;
                for (uint32_t o = 0; o < ix.ex_stride; o += 128) prefetch_l2(ep + o);
            }
            qn += n_new;
            if (qn >= 2 * kRefineSlots) flush();  // keeps rounds full and the threshold fresh
        };
        // 1-bit index (distance == estimate): replay the lanes of `mask` right away
        auto replay_direct = [&](unsigned mask, float lower, float est, unsigned long long gv) {
            unsigned long long vid = 0;
            if ((mask >> lane) & 1u) vid = ix.ids[gv];
            unsigned m = mask;
            while (m) {
                const int sl = __ffs(m) - 1;
                m &= m - 1;
                const float lb_s = __shfl_sync(0xffffffffu, lower, sl);
                const float d_s = __shfl_sync(0xffffffffu, est, sl);
                const unsigned long long id_s = __shfl_sync(0xffffffffu, vid, sl);
                const float theta = cnt >= k ? sd[k - 1] : INFINITY;
                if (lb_s >= theta) continue;
                st_adm += 1;
                if (!isfinite(d_s)) continue;
                topk_insert(sd, si, cnt, k, d_s, id_s, lane);
            }
        };

        uint32_t next_pi = a.nprobe;  // head pass: first rank left to the tail stage
        if (walk) {
            Cursor cc;
            cc.pi = start_pi;
            cc.seek(ix, pr, a.nprobe);
            while (cc.valid(a.nprobe)) {
                const Probe p = pr[cc.pi];
                const uint32_t nv = p.nv;
                const unsigned long long vbase = p.vec_off;
                st_blocks += cc.nb;
                const uint32_t list_pi = cc.pi;
                while (cc.valid(a.nprobe) && cc.pi == list_pi) {
                    const uint32_t b = cc.b;
                    mbar_wait(bars_u32 + 8 * use_stage, use_phase);
                    const uint32_t blk = ring_u32 + use_stage * B;
                    uint32_t accu = accumulate_block<NCB, WIDE>(blk, T, ncb, lane);
                    if (WIDE) accu &= 0xffffu;  // the reference accumulates in wrapping u16
                    const uint32_t fac = blk + (uint32_t)D * 4u + 4u * (uint32_t)lane;
                    const float f_add = lds_f32(fac), f_rescale = lds_f32(fac + 128u), f_error = lds_f32(fac + 256u);
                    // the stage has been read: hand it back to the TMA engine for the block NST ahead
                    __syncwarp();
                    if (++use_stage == NST) {
                        use_stage = 0;
                        use_phase ^= 1u;
                    }
                    --inflight;
                    if (pc.valid(a.nprobe)) {
                        if (lane == 0) {
                            fence_proxy_async();
                            mbar_expect_tx(bars_u32 + 8 * iss_stage, B);
                            tma_load_1d(ring_u32 + iss_stage * B, pc.base + (size_t)pc.b * B, B, bars_u32 + 8 * iss_stage);
                        }
                        if (++iss_stage == NST) iss_stage = 0;
                        ++inflight;
                        pc.next(ix, pr, a.nprobe);
                    }
                    // K8 (AVX2 variant): ip = fmadd(delta, accu, sum_vl); est = (f_add+g_add) + f_rescale*(ip+k1x)
                    const float ip = __fmaf_rn(s.delta, (float)accu, s.sum_vl);
                    const float t1 = ip + s.k1x;
                    const float t2 = f_rescale * t1;
                    const float t3 = f_add + p.g_add;
                    const float est = t3 + t2;
                    const float t4 = f_error * p.g_error;
                    float lower = est - t4;
                    // K9
                    const uint32_t li = b * kBatch + lane;
                    bool valid = li < nv;
                    if (a.filter != nullptr && valid) {
                        const uint32_t id32 = (uint32_t)ix.ids[vbase + li];
                        valid = (unsigned long long)id32 < a.filter_nbits && ((a.filter[id32 >> 6] >> (id32 & 63u)) & 1ull);
                    }
                    if (!isfinite(lower)) lower = l2 ? 0.0f : -(p.dot_qc + s.qnorm);
                    const float theta0 = cnt >= k ? sd[k - 1] : INFINITY;  // stale w.r.t. queued candidates => superset
                    const bool cand = valid && (lower < theta0);
                    const unsigned mask = __ballot_sync(0xffffffffu, cand);
                    st_cand += __popc(__ballot_sync(0xffffffffu, valid));
                    if (mask != 0u) {
                        if (EXK == 0) replay_direct(mask, lower, est, vbase + li);
                        else enqueue(mask, lower, ip, vbase + li, p.g_add);
                    }
                    cc.next(ix, pr, a.nprobe);
                }
                if (EXK != 0) flush();
                if (a.mode == kScanHead && cnt >= k) {  // heap full at a list boundary: the rest is the tail stage's
                    next_pi = cc.pi;
                    break;
                }
            }
            // head pass stopped early: drain the blocks still in flight so the ring is free for the next query
            while (inflight) {
                mbar_wait(bars_u32 + 8 * use_stage, use_phase);
                if (++use_stage == NST) {
                    use_stage = 0;
                    use_phase ^= 1u;
                }
                --inflight;
            }
        } else {
            // replay pass: survivors of the tail kernel, visited in (rank, position) order -- the order the
            // reference meets them -- against the live threshold.  Keys are unique, so ranking by counting sorts.
            const Survivor* sv = a.surv + (size_t)q * a.surv_cap;
            // sort key: rank (12 bits, nprobe <= 4096) | position (32) | slot in the buffer (10, cap <= 1024)
            uint32_t npad = 32;
            while (npad < n_surv) npad <<= 1;
            for (uint32_t i = lane; i < npad; i += 32)
                ord[i] = i < n_surv ? ((unsigned long long)sv[i].rank << 42) | ((unsigned long long)sv[i].pos << 10) | i : ~0ull;
            __syncwarp();
            for (uint32_t size = 2; size <= npad; size <<= 1) {  // bitonic sort, ascending
                for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                    for (uint32_t t = lane; t < npad / 2; t += 32) {
                        const uint32_t i = 2 * t - (t & (stride - 1)), j2 = i + stride;
                        const unsigned long long x = ord[i], y = ord[j2];
                        if ((x > y) == ((i & size) == 0)) {
                            ord[i] = y;
                            ord[j2] = x;
                        }
                    }
                    __syncwarp();
                }
            }
            for (uint32_t base = 0; base < n_surv; base += 32) {
                const uint32_t i = base + lane;
                const bool have = i < n_surv;
                Survivor rec = {0u, 0u, 0.0f, 0.0f};
                if (have) rec = sv[(uint32_t)ord[i] & 1023u];
                const Probe* pp = pr + (have ? rec.rank : 0u);
                const unsigned long long gv = have ? pp->vec_off + rec.pos : 0ull;
                const float gadd = have ? pp->g_add : 0.0f;
                const float theta0 = cnt >= k ? sd[k - 1] : INFINITY;
                const bool cand = have && (rec.lower < theta0);
                const unsigned mask = __ballot_sync(0xffffffffu, cand);
                if (mask != 0u) {
                    if (EXK == 0) replay_direct(mask, rec.lower, rec.x, gv);
                    else enqueue(mask, rec.lower, rec.x, gv, gadd);
                }
            }
            if (EXK != 0) flush();
        }
        for (int i = lane; i < k; i += 32) {
            const bool have = i < cnt;
            a.out_ids[(size_t)q * k + i] = have ? si[i] : ~0ull;
            a.out_scores[(size_t)q * k + i] = have ? (l2 ? sd[i] : -sd[i]) : 0.0f;
        }
        if (lane == 0) {
            a.out_counts[q] = (uint32_t)cnt;
            if (a.mode == kScanHead) {
                a.tail_start[q] = next_pi;
                a.tau[q] = cnt >= k ? sd[k - 1] : INFINITY;
            }
        }
        __syncwarp();
    }
    if (lane == 0 && a.stats) {
        atomicAdd(&a.stats->blocks, st_blocks);
        atomicAdd(&a.stats->candidates, st_cand);
        atomicAdd(&a.stats->refined, st_ref);
        atomicAdd(&a.stats->admitted, st_adm);
        if (st_ovf) atomicAdd(&a.stats->overflow_queries, st_ovf);
    }
}

static int g_num_sms = 0;
static size_t g_smem_optin = 0;
static int device_limits() {
    if (g_num_sms) return RBQ_OK;
    int dev = 0, v = 0;
    RBQ_CUDA(cudaGetDevice(&dev));
    RBQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    g_num_sms = v;
    RBQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    g_smem_optin = (size_t)v;
    return RBQ_OK;
}

template <int NCB, bool WIDE>
static int launch_scan_ex(const DevIndex& ix, ScanArgs& a, cudaStream_t st) {
    int rc = device_limits();
    if (rc) return rc;
    const bool has_ex = ix.ex_bits != 0;
    // refine staging: one byte per code (16 bytes per 16-dim chunk); stride = 32 (mod 128) so the four
    // 8-lane groups hit disjoint shared-memory banks
    a.ex_stage_stride = ((uint32_t)ix.D + 127u) / 128u * 128u + 32u;
    // ring depth: as many stages as fit 3 CTAs/SM, between 2 and 4
    const size_t per_sm = 227 * 1024;
    uint32_t nst = 4;
    const uint32_t ord_cap = a.mode == kScanReplay ? a.surv_cap : 0;
    for (; nst > 2; --nst) {
        const WarpSmem w = warp_smem_layout(ix.block_stride, nst, a.ex_stage_stride, ix.D, a.top_k, has_ex, ord_cap);
        if ((size_t)w.total * kWarps * 3 + 3 * 1024 <= per_sm) break;
    }
    a.nst = nst;
    const WarpSmem w = warp_smem_layout(ix.block_stride, nst, a.ex_stage_stride, ix.D, a.top_k, has_ex, ord_cap);
    const size_t smem = (size_t)w.total * kWarps;
    if (smem > g_smem_optin) return fail(RBQ_INVALID_CONFIG, "scan kernel shared memory exceeds the device limit");
    const unsigned ctas_per_sm = (unsigned)std::max<size_t>(1, std::min<size_t>(4, per_sm / (smem + 1024)));
    unsigned grid = (unsigned)std::min<size_t>(((size_t)a.nq + kWarps - 1) / kWarps, (size_t)g_num_sms * ctas_per_sm);
    if (a.mode == kScanFallback) grid = std::min<unsigned>(grid, (unsigned)g_num_sms);  // normally nothing to do: keep the launch light
#define RBQ_LAUNCH(EXK)                                                                                        \
    do {                                                                                                       \
        RBQ_CUDA(cudaFuncSetAttribute(scan_kernel<NCB, EXK, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                             \
        scan_kernel<NCB, EXK, WIDE><<<grid, kWarps * 32, smem, st>>>(ix, a);                                   \
    } while (0)
    if (ix.ex_bits == 0) RBQ_LAUNCH(0);
    else if (ix.ex_bits == 2) RBQ_LAUNCH(2);
    else if (ix.ex_bits == 6) RBQ_LAUNCH(6);
    else RBQ_LAUNCH(1);
#undef RBQ_LAUNCH
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int launch_scan(const DevIndex& ix, const float* d_rot, const uint8_t* d_lut, const QueryScalars* d_qs,
                const Probe* d_probes, size_t nq, size_t nprobe, size_t top_k, const uint64_t* d_filter,
                size_t filter_nbits, uint64_t* d_ids, float* d_scores, uint32_t* d_counts, DevStats* d_stats,
                unsigned int* d_work_counter, int mode, const TailWs* tw, cudaStream_t st) {
    if (nq == 0) return RBQ_OK;
    if (mode != kScanFull && tw == nullptr) return fail(RBQ_INVALID_CONFIG, "head/replay scan needs the tail workspace");
    if (top_k > (size_t)kMaxTopK) return fail(RBQ_INVALID_CONFIG, "top_k exceeds the device limit (1024)");
    if (ix.D > 2048) return fail(RBQ_INVALID_CONFIG, "padded_dim > 2048 (high-accuracy LUT path) is not supported");
    if (mode != kScanFallback && mode != kScanFallbackResume) RBQ_CUDA(cudaMemsetAsync(d_work_counter, 0, sizeof(unsigned int), st));
    ScanArgs a;
    a.rot = d_rot;
    a.lut = d_lut;
    a.qs = d_qs;
    a.probes = d_probes;
    a.nq = (uint32_t)nq;
    a.nprobe = (uint32_t)nprobe;
    a.top_k = (uint32_t)top_k;
    a.filter = reinterpret_cast<const unsigned long long*>(d_filter);
    a.filter_nbits = filter_nbits;
    a.out_ids = reinterpret_cast<unsigned long long*>(d_ids);
    a.out_scores = d_scores;
    a.out_counts = d_counts;
    a.stats = d_stats;
    a.work_counter = d_work_counter;
    a.nst = 2;
    a.ex_stage_stride = 16;
    a.mode = (uint32_t)mode;
    a.tail_start = tw ? tw->tail_start : nullptr;
    a.tau = tw ? tw->tau : nullptr;
    a.surv = tw ? tw->surv : nullptr;
    a.surv_cnt = tw ? tw->surv_cnt : nullptr;
    a.surv_cap = tw ? tw->surv_cap : 0;
    a.fb_list = tw ? tw->fb_list : nullptr;
    a.fb_count = tw ? tw->counters + 2 : nullptr;
    if (mode == kScanFallback) a.work_counter = tw->counters + 6;
    if (mode == kScanFallbackResume) {  // same kernel path, the replay tiers' list
        a.mode = (uint32_t)kScanFallback;
        a.fb_list = tw->fb2_list;
        a.fb_count = tw->counters + 9;
        a.work_counter = tw->counters + 10;
    }
    const int ncb_lane = (ix.D / 4 + 31) / 32;
    if (ix.D > 1024) {
        if (ncb_lane <= 12) return launch_scan_ex<12, true>(ix, a, st);
        return launch_scan_ex<16, true>(ix, a, st);
    }
    switch (ncb_lane) {
        case 1: return launch_scan_ex<1, false>(ix, a, st);
        case 2: return launch_scan_ex<2, false>(ix, a, st);
        case 3: return launch_scan_ex<3, false>(ix, a, st);
        case 4: return launch_scan_ex<4, false>(ix, a, st);
        case 5:
        case 6: return launch_scan_ex<6, false>(ix, a, st);
        default: return launch_scan_ex<8, false>(ix, a, st);
    }
}

// ---- stage probe: FastScan over one list for one query (tests compare every lane with the oracle) ----
__global__ void scan_debug_kernel(DevIndex ix, const uint8_t* __restrict__ lut, const QueryScalars* __restrict__ qs,
                                  uint32_t cluster, float g_add, float g_error, uint32_t* __restrict__ accu_out,
                                  float* __restrict__ ip_out, float* __restrict__ est_out, float* __restrict__ lb_out) {
    const int lane = threadIdx.x & 31, D = ix.D, ncb = D / 4;
    const uint32_t nv = ix.list_n[cluster], nb = (nv + kBatch - 1) / kBatch;
    const uint8_t* base = ix.blocks + (size_t)ix.blk_off[cluster] * ix.block_stride;
    const QueryScalars s = qs[0];
    for (uint32_t b = blockIdx.x; b < nb; b += gridDim.x) {
        const uint8_t* blk = base + (size_t)b * ix.block_stride;
        uint32_t total = 0;
        for (int c0 = 0; c0 < ncb; c0 += 32) {  // 32 codebooks per pass so that any D works
            const int cb = c0 + lane;
            uint4 C = make_uint4(0, 0, 0, 0), T = make_uint4(0, 0, 0, 0);
            if (cb < ncb) {
                C = ldg128(blk + 16 * cb);
                T = ldg128(lut + 16 * cb);
            }
            uint32_t acc[32];
#pragma unroll
            for (int v = 0; v < 32; ++v) acc[v] = 0u;
            lookup_accumulate(C, T, acc);
            total += reduce_scatter<true>(acc, lane);
        }
        const uint32_t accu = total & 0xffffu;
        const float* fac = reinterpret_cast<const float*>(blk + (size_t)D * 4);
        const float ip = __fmaf_rn(s.delta, (float)accu, s.sum_vl);
        const float t1 = ip + s.k1x;
        const float t2 = fac[32 + lane] * t1;
        const float t3 = fac[lane] + g_add;
        const float est = t3 + t2;
        const float t4 = fac[64 + lane] * g_error;
        accu_out[b * 32 + lane] = accu;
        ip_out[b * 32 + lane] = ip;
        est_out[b * 32 + lane] = est;
        lb_out[b * 32 + lane] = est - t4;
    }
}

int launch_scan_debug(const DevIndex& ix, const uint8_t* d_lut, const QueryScalars* d_qs, uint32_t cluster, float g_add,
                      float g_error, uint32_t* d_accu, float* d_ip, float* d_est, float* d_lb, cudaStream_t st) {
    scan_debug_kernel<<<64, 32, 0, st>>>(ix, d_lut, d_qs, cluster, g_add, g_error, d_accu, d_ip, d_est, d_lb);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// ---- multi-GPU: k-way merge of per-shard sorted top-k lists, one warp per query -------------------
__global__ void merge_kernel(int metric, int nshards, uint32_t nq, uint32_t k, const unsigned long long* __restrict__ in_ids,
                             const float* __restrict__ in_scores, const uint32_t* __restrict__ in_counts,
                             unsigned long long* __restrict__ out_ids, float* __restrict__ out_scores,
                             uint32_t* __restrict__ out_counts, size_t ids_stride, size_t sc_stride, size_t cn_stride) {
    // in_*[shard] start ids_stride / sc_stride / cn_stride ELEMENTS apart (separate [nshards][nq][k] arrays, or one packed
    // all-gather buffer holding ids | scores | counts per shard).  Shards hold disjoint id sets, each list is sorted best-first; repeatedly take the best head.
    // Order = the reference's result order: L2 ascending distance, IP descending score; ties -> lower shard.
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint32_t head = 0, cnt = 0;  // lane s (< nshards) tracks shard s; nshards <= 32
    if (lane < nshards) cnt = in_counts[(size_t)lane * cn_stride + q];
    uint32_t n = 0;
    for (; n < k; ++n) {
        float key = INFINITY;
        bool have = lane < nshards && head < cnt;
        if (have) {
            float sc = in_scores[(size_t)lane * sc_stride + (size_t)q * k + head];
            key = metric == RBQ_METRIC_L2 ? sc : -sc;
        }
        float best = key;
        int who = have ? lane : 64;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ow = __shfl_xor_sync(0xffffffffu, who, o);
            if (ow < 64 && (who >= 64 || ob < best || (ob == best && ow < who))) {
                best = ob;
                who = ow;
            }
        }
        if (who >= 64) break;
        if (lane == who) {
            out_ids[(size_t)q * k + n] = in_ids[(size_t)lane * ids_stride + (size_t)q * k + head];
            out_scores[(size_t)q * k + n] = in_scores[(size_t)lane * sc_stride + (size_t)q * k + head];
            head++;
        }
    }
    for (uint32_t i = n + lane; i < k; i += 32) {
        out_ids[(size_t)q * k + i] = ~0ull;
        out_scores[(size_t)q * k + i] = 0.0f;
    }
    if (lane == 0) out_counts[q] = n;
}

int launch_merge(int metric, int nshards, size_t nq, size_t top_k, const uint64_t* in_ids, const float* in_scores,
                 const uint32_t* in_counts, uint64_t* out_ids, float* out_scores, uint32_t* out_counts,
                 cudaStream_t st, size_t ids_stride, size_t sc_stride, size_t cn_stride) {
    if (nq == 0 || top_k == 0) return RBQ_OK;
    if (nshards < 1 || nshards > 32) return fail(RBQ_INVALID_CONFIG, "merge supports 1..32 shards");
    const unsigned grid = (unsigned)((nq + 3) / 4);
    merge_kernel<<<grid, 128, 0, st>>>(metric, nshards, (uint32_t)nq, (uint32_t)top_k,
                                       reinterpret_cast<const unsigned long long*>(in_ids), in_scores, in_counts,
                                       reinterpret_cast<unsigned long long*>(out_ids), out_scores, out_counts,
                                       ids_stride ? ids_stride : nq * top_k, sc_stride ? sc_stride : nq * top_k, cn_stride ? cn_stride : nq);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
