// resolve.cu -- the stages of the list-major pipeline that surround the tail kernel (scan_tail.cu).
//
//   head scan     K7+K8 (simd::accumulate_batch_avx2, compute_batch_distances_u16; reference src/simd.rs:972-1184,
//                 2090-2140) over every query's first owned list, one warp per query, blocks read straight from
//                 global memory; writes (lower bound, ip | estimate) of every vector to a dense per-query buffer.
//   head resolve  K9-K11 over that buffer: the reference's sequential prune / refine / keep-k-smallest loop
//                 (search_cluster_v2_batched, reference src/ivf.rs:2013-2127) with on-demand ex-code refinement.
//                 Leaves the heap state in the output arrays, tau (the k-th distance) and tail_start.
//   lazy replay   survivors sorted into the reference's visit order (probe rank, position), then the head-resolve loop over
//                 them: K10 (ip_packed_ex2_f32 / ip_packed_ex6_f32, AVX2 lane order, src/simd.rs:1722-1825) only for the
//                 candidates that beat the live threshold, the reference's decisions in order.
//   replay        the same for 1-bit indexes (distance == estimate: nothing to refine).
//
// The kernels are lean on purpose (no LUT registers, no block ring): 2-3x the resident warps of the sequential
// scan kernel, which is what the latency-bound refine rounds need.  Queries the fast path cannot take (head list
// longer than the dense buffer, heap not full after the head list, survivor overflow) are appended to a
// fallback list and walked by the sequential kernel (scan.cu, kScanFallback) -- same results, bit for bit.
#include <algorithm>
#include <cstdlib>

#include "scan_common.cuh"

namespace rbq {

constexpr int kResWarps = 4;

struct ResolveArgs {
    const float* rot;
    const uint8_t* lut;
    const QueryScalars* qs;
    const Probe* probes;
    uint32_t nq, nprobe, top_k;
    const uint8_t* head_owner;  // phased multi-GPU search: 0 = another shard runs this query's head pass (nullptr: all ours)
    const uint32_t* qlist;      // phased multi-GPU search: the queries whose head pass runs here, compacted (nullptr: identity);
    const uint32_t* qcount;     //   their number.  q_begin / q_count then address SLOTS of qlist
    uint32_t q_begin, q_count, cursor;  // head stage: the queries [q_begin, q_begin + q_count) of this launch; cursor = its work counter slot
    const unsigned long long* filter;
    unsigned long long filter_nbits;
    unsigned long long* out_ids;
    float* out_scores;
    uint32_t* out_counts;
    DevStats* stats;
    uint32_t* counters;  // TailWs::counters
    float2* head_buf;
    uint32_t head_cap;
    uint32_t* tail_start;
    float* tau;
    uint32_t* fb_list;   // head pass: queries it cannot take (from scratch)
    uint32_t* fb2_list;  // replay tiers: queries handed back (resume entries)
    Survivor* surv;
    const uint32_t* surv_cnt;
    uint32_t surv_cap;   // slots per query in the survivor buffer
    uint32_t sort_cap;   // survivors the lazy replay sorts per query; sort_cap < n <= surv_cap: overflow tier
    uint32_t* ovf_list;  // queries handed to the overflow tier (counters[7] entries)
    XrRec* ovf_recs;     // overflow tier scratch: kOvfMaxRecs records per CTA
    uint32_t exl_row;   // shared-memory stride of a lane's two staged code rows (exl2_lane_stride)
    uint32_t rql_row;   // shared-memory stride of a query pair-row (rql2_row_stride)
    uint32_t stage_bufs;  // 2: the rows of round r+1 travel while round r is multiplied; 1: one round at a time, less shared memory
    uint32_t has_ex;
    uint32_t flush_at;  // head resolve: refine as soon as this many candidates are queued
    uint32_t lazy_flush_at;  // lazy replay: queue length that triggers a refine round
};

// first probe rank >= from whose list has vectors on this shard (nprobe if none)
__device__ __forceinline__ uint32_t first_owned_rank(const Probe* __restrict__ pr, uint32_t nprobe, uint32_t from, int lane) {
    for (uint32_t base = from; base < nprobe; base += 32) {
        const uint32_t r = base + (uint32_t)lane;
        const unsigned m = __ballot_sync(0xffffffffu, r < nprobe && pr[r].nv != 0);
        if (m) return base + (uint32_t)(__ffs(m) - 1);
    }
    return nprobe;
}

// ---- head scan ---------------------------------------------------------------------------------------------
template <int NCB, bool WIDE>
__global__ void __launch_bounds__(128, 3) head_scan_kernel(DevIndex ix, ResolveArgs a) {
    const int lane = threadIdx.x & 31;
    const uint32_t slot = a.q_begin + blockIdx.x * 4u + (threadIdx.x >> 5);
    if (slot >= a.q_begin + a.q_count) return;
    uint32_t q = slot;
    if (a.qlist != nullptr) {
        if (slot >= *a.qcount) return;
        q = a.qlist[slot];
    }
    const Probe* pr = a.probes + (size_t)q * a.nprobe;
    const uint32_t h = first_owned_rank(pr, a.nprobe, 0, lane);
    if (h >= a.nprobe) return;
    const Probe p = pr[h];
    if (p.nv > a.head_cap) return;  // fallback query (resolve_head_kernel files it)
    const int D = ix.D, ncb = D / 4;
    uint4 T[NCB];
#pragma unroll
    for (int i = 0; i < NCB; ++i) {
        const int cb = lane + 32 * i;
        T[i] = (cb < ncb) ? ldg128(a.lut + (size_t)q * D * 4 + 16 * cb) : make_uint4(0, 0, 0, 0);
    }
    const QueryScalars s = a.qs[q];
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    const uint32_t nb = (p.nv + kBatch - 1) / kBatch;
    const uint8_t* base = ix.blocks + (size_t)p.blk_off * ix.block_stride;
    float2* out = a.head_buf + (size_t)(slot - a.q_begin) * a.head_cap;
    // the next block's codes and factors are in flight while the current block is looked up
    uint4 Cn[NCB];
    load_block_codes<NCB>(base, Cn, ncb, lane);
    const float* fac0 = reinterpret_cast<const float*>(base + (size_t)D * 4);
    float fn_add = __ldg(fac0 + lane), fn_rescale = __ldg(fac0 + 32 + lane), fn_error = __ldg(fac0 + 64 + lane);
    for (uint32_t b = 0; b < nb; ++b) {
        uint4 C[NCB];
#pragma unroll
        for (int i = 0; i < NCB; ++i) C[i] = Cn[i];
        const float f_add = fn_add, f_rescale = fn_rescale, f_error = fn_error;
        if (b + 1 < nb) {
            const uint8_t* nblk = base + (size_t)(b + 1) * ix.block_stride;
            load_block_codes<NCB>(nblk, Cn, ncb, lane);
            const float* fac = reinterpret_cast<const float*>(nblk + (size_t)D * 4);
            fn_add = __ldg(fac + lane);
            fn_rescale = __ldg(fac + 32 + lane);
            fn_error = __ldg(fac + 64 + lane);
        }
        uint32_t accu = accumulate_block_regs<NCB, WIDE>(C, T, lane);
        if (WIDE) accu &= 0xffffu;  // the reference accumulates in wrapping u16
        // K8 (AVX2 variant): ip = fmadd(delta, accu, sum_vl); est = (f_add+g_add) + f_rescale*(ip+k1x)
        const float ip = __fmaf_rn(s.delta, (float)accu, s.sum_vl);
        const float t1 = ip + s.k1x;
        const float t2 = f_rescale * t1;
        const float t3 = f_add + p.g_add;
        const float est = t3 + t2;
        const float t4 = f_error * p.g_error;
        float lower = est - t4;
        if (!isfinite(lower)) lower = l2 ? 0.0f : -(p.dot_qc + s.qnorm);
        out[b * kBatch + lane] = make_float2(lower, a.has_ex ? ip : est);
    }
}

// ---- per-warp shared memory of the resolve kernels ----------------------------------------------------------
struct ResSmem {
    uint32_t stage, rq, si, sd, ord, slots, total;
};
__host__ __device__ inline ResSmem res_smem_layout(uint32_t exl_row, uint32_t rql_row, uint32_t k, bool refine, bool topk, uint32_t surv_cap,
                                                   uint32_t stage_bufs = 2, bool paired = false) {
    ResSmem w;
    uint32_t o = 0;
    w.stage = o;
    o += refine ? stage_bufs * 32u * exl_row : 0;  // stage_bufs rounds of 32 lane regions (exl_row each): 4 candidates x 8 rows, or 8 x 4 x 2 rows (paired)
    w.rq = o;
    o += refine ? (paired ? 4u : 8u) * rql_row : 0;  // the query: 8 chain rows, or 4 pair-rows
    w.si = o;
    o += topk ? ((k * 8 + 15) / 16) * 16 : 0;
    w.sd = o;
    o += topk ? ((k * 4 + 15) / 16) * 16 : 0;
    uint32_t cap2 = surv_cap ? 32 : 0;
    while (cap2 < surv_cap) cap2 <<= 1;
    // Lazy replay (refine && surv_cap): the 64-bit sort keys are only needed until the survivors are sorted, which happens before
    // the first refinement, so they live in the refine staging + query rows when those are large enough; what the replay loop
    // reads afterwards is the sorted buffer slots, 2 bytes each.
    const bool alias = refine && surv_cap != 0 && (w.si - w.stage) >= cap2 * 8;
    w.slots = o;
    o += (refine && surv_cap != 0) ? ((cap2 * 2 + 15) / 16) * 16 : 0;
    w.ord = alias ? w.stage : o;
    o += alias ? 0 : cap2 * 8;
    w.total = o;
    return w;
}

// load_rql2 / refine_batch2 (paired chains, 8 candidates per round): scan_common.cuh

// ---- lane-major ex-codes (DevIndex::exl) ------------------------------------------------------------------------------
// One thread per (vector, 16-dim chunk): decodes the chunk of the packed code (the reference's layouts, src/simd.rs:
// 2478-2695, or the generic LSB-first stream) and scatters its 16 bytes to the 8 chain rows.  Runs once per index load.
template <int EXK>
__global__ void relayout_ex_kernel(const uint8_t* __restrict__ ex, uint32_t ex_stride, int ex_bits, int D, size_t nvec, uint32_t lane_bytes,
                                   uint8_t* __restrict__ exl) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int chunks = D / 16;
    if (idx >= nvec * (size_t)chunks) return;
    const size_t v = idx / chunks;
    const int c = (int)(idx % chunks);
    uint32_t X[4];
    decode_chunk<EXK, false>(ex + v * ex_stride, c, ex_bits, X[0], X[1], X[2], X[3]);
    uint8_t* dst = exl + v * (size_t)(8u * lane_bytes);
#pragma unroll
    for (int r = 0; r < 16; ++r)  // dim 16c + r = 8t + j with j = r & 7, t = 2c + (r >> 3)
        dst[(size_t)(r & 7) * lane_bytes + 2 * c + (r >> 3)] = (uint8_t)(X[r >> 2] >> (8 * (r & 3)));
}

// ---- head resolve ---------------------------------------------------------------------------------------------
template <int EXK>
__global__ void __launch_bounds__(kResWarps * 32, 6) resolve_head_kernel(DevIndex ix, ResolveArgs a) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = ix.D, k = (int)a.top_k;
    const ResSmem L = res_smem_layout(a.exl_row, a.rql_row, k, EXK != 0, true, 0, a.stage_bufs, EXK == 2);
    unsigned char* wbase = res_smem + (size_t)warp * L.total;
    const uint32_t stage_u32 = smem_u32(wbase + L.stage), rql_u32 = smem_u32(wbase + L.rq);
    unsigned char* rql = wbase + L.rq;
    unsigned long long* si = reinterpret_cast<unsigned long long*>(wbase + L.si);
    float* sd = reinterpret_cast<float*>(wbase + L.sd);
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    unsigned long long st_blocks = 0, st_cand = 0, st_ref = 0, st_adm = 0;

    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(&a.counters[a.cursor], 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= a.q_count) break;
        q += a.q_begin;
        const uint32_t slot = q;  // row of the dense head buffer = slot - q_begin
        if (a.qlist != nullptr) {
            if (slot >= *a.qcount) break;
            q = a.qlist[slot];
        }
        const Probe* pr = a.probes + (size_t)q * a.nprobe;
        const uint32_t h = first_owned_rank(pr, a.nprobe, 0, lane);
        TopK tk;
        tk.init(sd, si, k);
        bool fallback = false;
        uint32_t next_start = a.nprobe;
        unsigned long long q_blocks = 0, q_cand = 0, q_ref = 0, q_adm = 0;
        if (h < a.nprobe) {
            const Probe p = pr[h];
            if (p.nv > a.head_cap) {
                fallback = true;
            } else {
                if (EXK != 0) load_rql_any<EXK == 2>(rql, a.rql_row, a.rot + (size_t)q * D, D, ix.exl_lane, lane);
                const QueryScalars s = a.qs[q];
                __syncwarp();
                const uint32_t nv = p.nv, nb = (nv + kBatch - 1) / kBatch;
                const unsigned long long vbase = p.vec_off;
                const float2* hb = a.head_buf + (size_t)(slot - a.q_begin) * a.head_cap;
                q_blocks = nb;

                // candidate queue: slot i lives in lane i, in visit order
                int qn = 0;
                float q_lower = 0.0f, q_ip = 0.0f;
                unsigned long long q_gv = 0;
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
                    const float exdot = refine_batch_any<EXK == 2>(ix, q_gv, qn, stage_u32, rql_u32, a.exl_row, a.rql_row, a.stage_bufs, lane);
                    q_ref += qn;
                    if (mine) {
                        // distance = f_add_ex + g_add + f_rescale_ex * (binary_scale*ip + ex_dot + kbx)  (ivf.rs:2095-2099)
                        float tt = s.bscale * q_ip;
                        tt = tt + exdot;
                        tt = tt + s.kbx;
                        const float mm2 = fre * tt;
                        const float aa = fae + p.g_add;
                        dist = aa + mm2;
                    }
                    for (int c = 0; c < qn; ++c) {
                        const float lb_s = __shfl_sync(0xffffffffu, q_lower, c);
                        const float d_s = __shfl_sync(0xffffffffu, dist, c);
                        const unsigned long long id_s = __shfl_sync(0xffffffffu, q_vid, c);
                        const float theta = tk.theta();
                        if (lb_s >= theta) continue;  // skipped_by_lower_bound
                        q_adm += 1;
                        if (!isfinite(d_s)) continue;
                        tk.insert(d_s, id_s, lane);
                    }
                    qn = 0;
                };
                auto enqueue = [&](unsigned mask, float lower, float ipv, unsigned long long gv) {
                    const int n_new = __popc(mask);
                    if (qn + n_new > 32) flush();
                    const int r = lane - qn;
                    const int src = (r >= 0 && r < n_new) ? (int)__fns(mask, 0, r + 1) : 0;
                    const float nl = __shfl_sync(0xffffffffu, lower, src);
                    const float nip = __shfl_sync(0xffffffffu, ipv, src);
                    const unsigned long long ngv = __shfl_sync(0xffffffffu, gv, src);
                    if (r >= 0 && r < n_new) {
                        q_lower = nl;
                        q_ip = nip;
                        q_gv = ngv;
                        const uint8_t* ep = ix.exl + q_gv * ix.exl_stride;  // warm L2 with the candidate's ex-code rows
                        for (uint32_t o = 0; o < ix.exl_stride; o += 128) prefetch_l2(ep + o);
                    }
                    qn += n_new;
                    if (qn >= (int)a.flush_at) flush();  // keeps rounds full and the threshold fresh
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
                        const float theta = tk.theta();
                        if (lb_s >= theta) continue;
                        q_adm += 1;
                        if (!isfinite(d_s)) continue;
                        tk.insert(d_s, id_s, lane);
                    }
                };

                float2 cur = hb[lane];
                for (uint32_t b = 0; b < nb; ++b) {
                    const float2 rec = cur;
                    if (b + 1 < nb) cur = hb[(b + 1) * kBatch + lane];
                    const uint32_t li = b * kBatch + lane;
                    bool valid = li < nv;
                    if (a.filter != nullptr && valid) {
                        const uint32_t id32 = (uint32_t)ix.ids[vbase + li];
                        valid = (unsigned long long)id32 < a.filter_nbits && ((a.filter[id32 >> 6] >> (id32 & 63u)) & 1ull);
                    }
                    const float theta0 = tk.theta();  // stale w.r.t. queued candidates => superset
                    const bool cand = valid && (rec.x < theta0);
                    const unsigned mask = __ballot_sync(0xffffffffu, cand);
                    q_cand += __popc(__ballot_sync(0xffffffffu, valid));
                    if (mask != 0u) {
                        if (EXK == 0) replay_direct(mask, rec.x, rec.y, vbase + li);
                        else enqueue(mask, rec.x, rec.y, vbase + li);
                    }
                }
                if (EXK != 0) flush();
                const bool more = first_owned_rank(pr, a.nprobe, h + 1, lane) < a.nprobe;
                if (tk.cnt >= k) next_start = h + 1;
                else if (more) fallback = true;  // the heap is not full yet: the sequential kernel walks on
            }
        }
        if (fallback) {
            if (lane == 0) {
                if (a.stats) atomicAdd(&a.stats->fallback_queries, 1ull);
                a.fb_list[atomicAdd(&a.counters[2], 1u)] = q;  // from scratch
                a.tail_start[q] = a.nprobe;
                a.tau[q] = INFINITY;
                a.out_counts[q] = 0u;
            }
            __syncwarp();
            continue;
        }
        st_blocks += q_blocks;
        st_cand += q_cand;
        st_ref += q_ref;
        st_adm += q_adm;
        const float tau_q = tk.theta();
        __syncwarp();
        for (int i = lane; i < k; i += 32) {
            const bool have = i < tk.cnt;
            const float dv = tk.dist_at(i);
            a.out_ids[(size_t)q * k + i] = have ? tk.id_at(i) : ~0ull;
            a.out_scores[(size_t)q * k + i] = have ? (l2 ? dv : -dv) : 0.0f;
        }
        if (lane == 0) {
            a.out_counts[q] = (uint32_t)tk.cnt;
            a.tail_start[q] = next_start;
            a.tau[q] = tau_q;
        }
        __syncwarp();
    }
    if (lane == 0 && a.stats) {
        if (st_blocks) atomicAdd(&a.stats->blocks, st_blocks);
        if (st_cand) atomicAdd(&a.stats->candidates, st_cand);
        if (st_ref) atomicAdd(&a.stats->refined, st_ref);
        if (st_adm) atomicAdd(&a.stats->admitted, st_adm);
    }
}

// ---- replay ------------------------------------------------------------------------------------------------------
template <int EXK>
__global__ void __launch_bounds__(kResWarps * 32) resolve_replay_kernel(DevIndex ix, ResolveArgs a) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = (int)a.top_k;
    const ResSmem L = res_smem_layout(0, 0, k, false, true, a.sort_cap);
    unsigned char* wbase = res_smem + (size_t)warp * L.total;
    unsigned long long* si = reinterpret_cast<unsigned long long*>(wbase + L.si);
    float* sd = reinterpret_cast<float*>(wbase + L.sd);
    unsigned long long* ord = reinterpret_cast<unsigned long long*>(wbase + L.ord);
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    unsigned long long st_adm = 0, st_ovf = 0;
    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(&a.counters[5], 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= a.nq) break;
        const uint32_t start_pi = a.tail_start[q], n_surv = a.surv_cnt[q];
        if (start_pi >= a.nprobe || n_surv == 0) continue;  // the head result is already final
        if (n_surv > a.sort_cap) {  // more than this kernel sorts: the overflow tier replays it; beyond the buffer: sequential re-walk of the tail
            if (lane == 0) {
                if (n_surv <= a.surv_cap) a.ovf_list[atomicAdd(&a.counters[7], 1u)] = q;
                else a.fb2_list[atomicAdd(&a.counters[9], 1u)] = q | kFbResume;
            }
            st_ovf += 1;
            continue;
        }
        const Probe* pr = a.probes + (size_t)q * a.nprobe;
        const float tau_q = a.tau[q];  // caps the live threshold (see resolve_lazy_kernel)
        __syncwarp();
        int cnt = (int)a.out_counts[q];
        for (int i = lane; i < cnt; i += 32) {  // resume from the head pass' top-k (stored best-first)
            const float sc = a.out_scores[(size_t)q * k + i];
            sd[i] = l2 ? sc : -sc;
            si[i] = a.out_ids[(size_t)q * k + i];
        }
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
            unsigned long long vid = 0;
            if (have) {
                const uint32_t slot = (uint32_t)ord[i] & 1023u;
                rec = sv[slot];
                vid = ix.ids[pr[rec.rank].vec_off + rec.pos];
            }
            const float theta0 = fminf(cnt >= k ? sd[k - 1] : INFINITY, tau_q);
            unsigned m = __ballot_sync(0xffffffffu, have && (rec.lower < theta0));
            while (m) {
                const int sl = __ffs(m) - 1;
                m &= m - 1;
                const float lb_s = __shfl_sync(0xffffffffu, rec.lower, sl);
                const float d_s = __shfl_sync(0xffffffffu, rec.x, sl);
                const unsigned long long id_s = __shfl_sync(0xffffffffu, vid, sl);
                const float theta = fminf(cnt >= k ? sd[k - 1] : INFINITY, tau_q);
                if (lb_s >= theta) continue;  // skipped_by_lower_bound
                st_adm += 1;
                if (!isfinite(d_s)) continue;
                topk_insert(sd, si, cnt, k, d_s, id_s, lane);
            }
        }
        for (int i = lane; i < k; i += 32) {
            const bool have = i < cnt;
            a.out_ids[(size_t)q * k + i] = have ? si[i] : ~0ull;
            a.out_scores[(size_t)q * k + i] = have ? (l2 ? sd[i] : -sd[i]) : 0.0f;
        }
        if (lane == 0) a.out_counts[q] = (uint32_t)cnt;
        __syncwarp();
    }
    if (lane == 0 && a.stats) {
        if (st_adm) atomicAdd(&a.stats->admitted, st_adm);
        if (st_ovf) atomicAdd(&a.stats->overflow_queries, st_ovf);
    }
}

// ---- lazy replay: on-demand refinement of the sorted survivors ---------------------------------------------------
// Every survivor has lower bound < the head threshold tau, but the reference only
// refines a candidate whose lower bound beats the LIVE k-th distance, which keeps falling while the tail is walked
// (reference src/ivf.rs:2044-2052): at GIST/nprobe 16 that is ~15 of ~75 survivors per query.  This kernel sorts the
// survivors into the reference's visit order first and then runs the head-resolve loop over them: candidates that beat
// the current (stale => looser => superset) threshold are queued, a full queue is refined in one batch and replayed
// against the live threshold.  Same decisions as the reference, a fraction of the ex-code traffic and FMA chains.
template <int EXK>
#ifdef RBQ_EXL_COPY_REGULAR
__global__ void __launch_bounds__(kResWarps * 32, 6) resolve_lazy_kernel(DevIndex ix, ResolveArgs a) {
#else
__global__ void __launch_bounds__(kResWarps * 32) resolve_lazy_kernel(DevIndex ix, ResolveArgs a) {
#endif
    extern __shared__ __align__(16) unsigned char res_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = ix.D, k = (int)a.top_k;
    const ResSmem L = res_smem_layout(a.exl_row, a.rql_row, k, true, true, a.sort_cap, a.stage_bufs, EXK == 2);
    unsigned char* wbase = res_smem + (size_t)warp * L.total;
    const uint32_t stage_u32 = smem_u32(wbase + L.stage), rql_u32 = smem_u32(wbase + L.rq);
    unsigned char* rql = wbase + L.rq;
    unsigned long long* si = reinterpret_cast<unsigned long long*>(wbase + L.si);
    float* sd = reinterpret_cast<float*>(wbase + L.sd);
    unsigned long long* ord = reinterpret_cast<unsigned long long*>(wbase + L.ord);  // may alias the staging + query rows
    uint16_t* slots = reinterpret_cast<uint16_t*>(wbase + L.slots);
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    unsigned long long st_adm = 0, st_ovf = 0, st_ref = 0;
    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(&a.counters[5], 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= a.nq) break;
        const uint32_t start_pi = a.tail_start[q], n_surv = a.surv_cnt[q];
        if (start_pi >= a.nprobe || n_surv == 0) continue;  // the head result is already final
        if (n_surv > a.sort_cap) {  // more than this kernel sorts: the overflow tier replays it; beyond the buffer: sequential re-walk of the tail
            if (lane == 0) {
                if (n_surv <= a.surv_cap) a.ovf_list[atomicAdd(&a.counters[7], 1u)] = q;
                else a.fb2_list[atomicAdd(&a.counters[9], 1u)] = q | kFbResume;
            }
            st_ovf += 1;
            continue;
        }
        const Probe* pr = a.probes + (size_t)q * a.nprobe;
        __syncwarp();
        const QueryScalars s = a.qs[q];
        // the head threshold caps the live one: a no-op on one GPU (the heap's k-th distance starts at tau and only falls), the
        // bound another shard's head pass established in the phased multi-GPU search
        const float tau_q = a.tau[q];
        TopK tk;
        tk.init(sd, si, k);
        tk.cnt = (int)a.out_counts[q];
        for (int i = lane; i < tk.cnt; i += 32) {  // resume from the head pass' top-k (stored best-first)
            const float sc = a.out_scores[(size_t)q * k + i];
            tk.set_at(i, l2 ? sc : -sc, a.out_ids[(size_t)q * k + i]);
        }
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
        for (uint32_t i = lane; i < n_surv; i += 32) slots[i] = (uint16_t)((uint32_t)ord[i] & 1023u);
        __syncwarp();  // the keys are dead: their memory becomes the query rows and the refine staging
        load_rql_any<EXK == 2>(rql, a.rql_row, a.rot + (size_t)q * D, D, ix.exl_lane, lane);
        __syncwarp();
        // candidate queue: slot i lives in lane i, in visit order
        int qn = 0;
        float q_lower = 0.0f, q_ip = 0.0f, q_gadd = 0.0f;
        unsigned long long q_gv = 0;
        unsigned long long q_ref = 0, q_adm = 0;
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
            const float exdot = refine_batch_any<EXK == 2>(ix, q_gv, qn, stage_u32, rql_u32, a.exl_row, a.rql_row, a.stage_bufs, lane);
            q_ref += qn;
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
                const float theta = fminf(tk.theta(), tau_q);
                if (lb_s >= theta) continue;  // skipped_by_lower_bound
                q_adm += 1;
                if (!isfinite(d_s)) continue;
                tk.insert(d_s, id_s, lane);
            }
            qn = 0;
        };
        const int fl = (int)a.lazy_flush_at;
        for (uint32_t base = 0; base < n_surv; base += 32) {
            const uint32_t i = base + lane;
            bool have = i < n_surv;
            Survivor rec = {0u, 0u, 0.0f, 0.0f};
            unsigned long long gv = 0;
            float g_add = 0.0f;
            if (have) {
                rec = sv[slots[i]];
                const Probe* pp = pr + rec.rank;
                gv = pp->vec_off + rec.pos;
                g_add = pp->g_add;
            }
            {   // the first few likely candidates of the batch: start their ex-codes towards L2 now
                const float th = fminf(tk.theta(), tau_q);
                const bool likely = have && (rec.lower < th);
                const unsigned m0 = __ballot_sync(0xffffffffu, likely);
                if (likely && __popc(m0 & ((1u << lane) - 1u)) < 2 * fl) {
                    const uint8_t* ep = ix.exl + gv * ix.exl_stride;
                    for (uint32_t o = 0; o < ix.exl_stride; o += 128) prefetch_l2(ep + o);
                }
            }
            // the batch is consumed in visit order, a queue-full at a time, so that the threshold is refreshed between
            // refine rounds; a lane that fails the test once is out for good (the threshold never rises)
            for (;;) {
                const float theta0 = fminf(tk.theta(), tau_q);  // stale w.r.t. queued candidates => superset
                have = have && (rec.lower < theta0);
                unsigned mask = __ballot_sync(0xffffffffu, have);
                if (mask == 0u) break;
                const int room = fl - qn;  // >= 1: a full queue is flushed right away
                if (__popc(mask) > room) mask &= (2u << __fns(mask, 0, room)) - 1u;  // the first `room` candidates
                const int n_new = __popc(mask);
                const int r = lane - qn;
                const int src = (r >= 0 && r < n_new) ? (int)__fns(mask, 0, r + 1) : 0;
                const float nl = __shfl_sync(0xffffffffu, rec.lower, src);
                const float nip = __shfl_sync(0xffffffffu, rec.x, src);
                const float nga = __shfl_sync(0xffffffffu, g_add, src);
                const unsigned long long ngv = __shfl_sync(0xffffffffu, gv, src);
                if (r >= 0 && r < n_new) {
                    q_lower = nl;
                    q_ip = nip;
                    q_gadd = nga;
                    q_gv = ngv;
                }
                if ((mask >> lane) & 1u) have = false;  // queued
                qn += n_new;
                if (qn >= fl) flush();
            }
        }
        flush();
        st_ref += q_ref;
        st_adm += q_adm;
        __syncwarp();
        for (int i = lane; i < k; i += 32) {
            const bool have = i < tk.cnt;
            const float dv = tk.dist_at(i);
            a.out_ids[(size_t)q * k + i] = have ? tk.id_at(i) : ~0ull;
            a.out_scores[(size_t)q * k + i] = have ? (l2 ? dv : -dv) : 0.0f;
        }
        if (lane == 0) a.out_counts[q] = (uint32_t)tk.cnt;
        __syncwarp();
    }
    if (lane == 0 && a.stats) {
        if (st_adm) atomicAdd(&a.stats->admitted, st_adm);
        if (st_ref) atomicAdd(&a.stats->refined, st_ref);
        if (st_ovf) atomicAdd(&a.stats->overflow_queries, st_ovf);
    }
}

// ---- overflow tier: queries with more survivors than the lazy replay sorts (sort_cap < n <= surv_cap) -------------------
// Such a query sits where the head threshold is loose (its nearest list is tiny, or -- phased multi-GPU search -- another
// shard's): nearly every vector of its other lists survives.  The sequential kernel needs ~0.6 ms for one of them, a latency
// chain that does not shrink with the number of GPUs.  Here one CTA takes the query: all survivors are refined EAGERLY by the
// CTA's warps (independent, 8 or 4 candidates per warp and round), {visit key, lower bound, distance, id} records go to a
// per-CTA scratch, the keys are bitonic-sorted in shared memory, and one warp replays the records in the reference's visit
// order against the live threshold -- the same decisions as the lazy replay (src/ivf.rs:2044-2127), the distance of a
// candidate does not depend on when it is computed.
constexpr int kOvfWarps = 8;
constexpr uint32_t kOvfMaxRecs = 4096;  // == the survivor buffer's slots per query at the default caps
struct OvfSmem {
    uint32_t keys, idx, warp0, warp_stride, si, sd, total;
};
__host__ __device__ inline OvfSmem ovf_smem_layout(uint32_t exl_row, uint32_t rql_row, uint32_t k, bool refine, bool paired, uint32_t stage_bufs) {
    OvfSmem w;
    uint32_t o = 0;
    w.keys = o;
    o += kOvfMaxRecs * 8u;
    w.idx = o;
    o += kOvfMaxRecs * 4u;
    w.warp0 = o;
    w.warp_stride = refine ? stage_bufs * 32u * exl_row + (paired ? 4u : 8u) * rql_row : 0u;
    o += (uint32_t)kOvfWarps * w.warp_stride;
    w.si = o;
    o += ((k * 8 + 15) / 16) * 16;
    w.sd = o;
    o += ((k * 4 + 15) / 16) * 16;
    w.total = o;
    return w;
}

template <int EXK>
__global__ void __launch_bounds__(kOvfWarps * 32) overflow_replay_kernel(DevIndex ix, ResolveArgs a) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    __shared__ uint32_t s_q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int D = ix.D, k = (int)a.top_k;
    const OvfSmem L = ovf_smem_layout(a.exl_row, a.rql_row, a.top_k, EXK != 0, EXK == 2, a.stage_bufs);
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(res_smem + L.keys);
    uint32_t* ridx = reinterpret_cast<uint32_t*>(res_smem + L.idx);
    unsigned char* wbase = res_smem + L.warp0 + (size_t)warp * L.warp_stride;
    const uint32_t stage_u32 = smem_u32(wbase), rql_u32 = smem_u32(wbase + a.stage_bufs * 32u * a.exl_row);
    unsigned char* rql = wbase + a.stage_bufs * 32u * a.exl_row;
    unsigned long long* si = reinterpret_cast<unsigned long long*>(res_smem + L.si);
    float* sd = reinterpret_cast<float*>(res_smem + L.sd);
    XrRec* recs = a.ovf_recs + (size_t)blockIdx.x * kOvfMaxRecs;
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    for (;;) {
        __syncthreads();
        if (tid == 0) s_q = atomicAdd(&a.counters[8], 1u);
        __syncthreads();
        const uint32_t slot = s_q;
        if (slot >= a.counters[7]) break;
        const uint32_t q = a.ovf_list[slot];
        const uint32_t m = a.surv_cnt[q];  // sort_cap < m <= surv_cap (the lazy replay filed it)
        if (m > kOvfMaxRecs) {             // cannot happen at the default caps; keep the exact path anyway
            if (tid == 0) a.fb2_list[atomicAdd(&a.counters[9], 1u)] = q | kFbResume;
            continue;
        }
        const Probe* pr = a.probes + (size_t)q * a.nprobe;
        const Survivor* sv = a.surv + (size_t)q * a.surv_cap;
        QueryScalars s{};
        if (EXK != 0) {
            s = a.qs[q];
            load_rql_any<EXK == 2>(rql, a.rql_row, a.rot + (size_t)q * D, D, ix.exl_lane, lane);  // every warp stages its own copy
            __syncwarp();
        }
        // 1. records: eager refinement, 32 survivors per warp and pass
        for (uint32_t b0 = (uint32_t)warp * 32u; b0 < m; b0 += (uint32_t)kOvfWarps * 32u) {
            const uint32_t i = b0 + (uint32_t)lane;
            const int nb = (int)min(32u, m - b0);
            Survivor rec = {0u, 0u, 0.0f, 0.0f};
            unsigned long long gv = 0;
            float g_add = 0.0f;
            if (i < m) {
                rec = sv[i];
                const Probe* pp = pr + rec.rank;
                gv = pp->vec_off + rec.pos;
                g_add = pp->g_add;
            }
            float dist = rec.x;  // 1-bit index: the estimate is the distance
            if (EXK != 0) {
                const float exdot = refine_batch_any<EXK == 2>(ix, gv, nb, stage_u32, rql_u32, a.exl_row, a.rql_row, a.stage_bufs, lane);
                if (i < m) {
                    // distance = f_add_ex + g_add + f_rescale_ex * (binary_scale*ip + ex_dot + kbx)  (ivf.rs:2095-2099)
                    const float fae = __ldg(ix.f_add_ex + gv), fre = __ldg(ix.f_rescale_ex + gv);
                    float tt = s.bscale * rec.x;
                    tt = tt + exdot;
                    tt = tt + s.kbx;
                    const float mm2 = fre * tt;
                    const float aa = fae + g_add;
                    dist = aa + mm2;
                }
            }
            if (i < m) {
                const unsigned long long key = ((unsigned long long)rec.rank << 32) | rec.pos;
                recs[i] = XrRec{key, rec.lower, dist, ix.ids[gv]};
                keys[i] = key;
                ridx[i] = i;
            }
        }
        uint32_t npad = 32;
        while (npad < m) npad <<= 1;
        for (uint32_t i = m + (uint32_t)tid; i < npad; i += (uint32_t)kOvfWarps * 32u) {
            keys[i] = ~0ull;
            ridx[i] = 0u;
        }
        __syncthreads();
        // 2. visit order
        for (uint32_t size = 2; size <= npad; size <<= 1)
            for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
                for (uint32_t t = (uint32_t)tid; t < npad / 2; t += (uint32_t)kOvfWarps * 32u) {
                    const uint32_t i = 2 * t - (t & (stride - 1)), j2 = i + stride;
                    const unsigned long long x = keys[i], y = keys[j2];
                    if ((x > y) == ((i & size) == 0)) {
                        keys[i] = y;
                        keys[j2] = x;
                        const uint32_t u = ridx[i];
                        ridx[i] = ridx[j2];
                        ridx[j2] = u;
                    }
                }
                __syncthreads();
            }
        if (warp != 0) continue;
        // 3. warp 0: the reference's loop over the candidates in visit order, resumed from the head pass' top-k
        __threadfence_block();
        const float tau_q = a.tau[q];
        TopK tk;
        tk.init(sd, si, k);
        tk.cnt = (int)a.out_counts[q];
        for (int i = lane; i < tk.cnt; i += 32) {
            const float sc = a.out_scores[(size_t)q * k + i];
            tk.set_at(i, l2 ? sc : -sc, a.out_ids[(size_t)q * k + i]);
        }
        // The threshold only moves when a candidate enters a full heap, and only a candidate whose distance beats the k-th
        // distance AT THE START of the batch can do that (the k-th distance never rises).  So the batch is walked from one such
        // candidate to the next; for the lanes in between the threshold is constant and "admitted" is a population count.  A
        // query of this tier has thousands of candidates below a loose lower-bound threshold and a handful that change the heap.
        unsigned long long q_adm = 0;
        for (uint32_t b0 = 0; b0 < m; b0 += 32) {
            const uint32_t i = b0 + (uint32_t)lane;
            XrRec r{0ull, 0.0f, 0.0f, 0ull};
            if (i < m) r = recs[ridx[i]];
            const float kth0 = tk.theta();  // warp-collective for k <= 32: every lane calls it, outside any short-circuit
            float theta = fminf(kth0, tau_q);
            const bool below = i < m && r.lower < theta;  // stale threshold: superset of what the live one admits
            const bool mover = below && isfinite(r.dist) && (tk.cnt < k || r.dist < kth0);
            unsigned movers = __ballot_sync(0xffffffffu, mover);
            int done = 0;  // lanes [0, done) are accounted for
            while (movers) {
                const int sl = __ffs(movers) - 1;
                movers &= movers - 1;
                // lanes [done, sl): admitted iff their lower bound beats the current threshold; none of them changes the heap
                q_adm += (unsigned long long)__popc(__ballot_sync(0xffffffffu, lane >= done && lane < sl && below && r.lower < theta));
                const float lb_s = __shfl_sync(0xffffffffu, r.lower, sl);
                const float d_s = __shfl_sync(0xffffffffu, r.dist, sl);
                const unsigned long long id_s = __shfl_sync(0xffffffffu, r.id, sl);
                done = sl + 1;
                if (lb_s >= theta) continue;  // skipped_by_lower_bound
                q_adm += 1;
                tk.insert(d_s, id_s, lane);   // finite by construction of `mover`
                theta = fminf(tk.theta(), tau_q);
            }
            q_adm += (unsigned long long)__popc(__ballot_sync(0xffffffffu, lane >= done && below && r.lower < theta));
        }
        __syncwarp();
        for (int i = lane; i < k; i += 32) {
            const bool have = i < tk.cnt;
            const float dv = tk.dist_at(i);
            a.out_ids[(size_t)q * k + i] = have ? tk.id_at(i) : ~0ull;
            a.out_scores[(size_t)q * k + i] = have ? (l2 ? dv : -dv) : 0.0f;
        }
        if (lane == 0) {
            a.out_counts[q] = (uint32_t)tk.cnt;
            if (a.stats) {
                atomicAdd(&a.stats->admitted, q_adm);
                if (EXK != 0) atomicAdd(&a.stats->refined, (unsigned long long)m);
            }
        }
    }
}

// Stage probe for K10: the ex-code dot of `n` stored vectors (global positions gv[i]) against one rotated query, through the
// product's own refine path (the form the resolve kernels use for this padded_dim: lane-major rows, 8 FMA chains, AVX2-order sum).
template <bool PAIRED>
__global__ void __launch_bounds__(32) ex_dot_debug_kernel(DevIndex ix, ResolveArgs a, const unsigned long long* __restrict__ gv, int n,
                                                         float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char res_smem[];
    const int lane = threadIdx.x;
    const ResSmem L = res_smem_layout(a.exl_row, a.rql_row, 1, true, false, 0, a.stage_bufs, PAIRED);
    load_rql_any<PAIRED>(res_smem + L.rq, a.rql_row, a.rot, ix.D, ix.exl_lane, lane);
    __syncwarp();
    for (int b0 = 0; b0 < n; b0 += 32) {
        const int m = min(32, n - b0);
        const unsigned long long g = lane < m ? gv[b0 + lane] : 0ull;
        const float d = refine_batch_any<PAIRED>(ix, g, m, smem_u32(res_smem + L.stage), smem_u32(res_smem + L.rq), a.exl_row, a.rql_row, a.stage_bufs, lane);
        if (lane < m) out[b0 + lane] = d;
        __syncwarp();
    }
}
__global__ void fill_u32_kernel(uint32_t* p, size_t n, uint32_t v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
int launch_fill_u32(uint32_t* d_p, size_t n, uint32_t v, cudaStream_t st) {
    if (n == 0) return RBQ_OK;
    fill_u32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_p, n, v);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int prepare_ex_lanes(rbq_index* h) {
    DevIndex& d = h->dev;
    d.exl = nullptr;
    d.exl_lane = exl_lane_bytes((uint32_t)d.D);
    d.exl_stride = 8u * d.exl_lane;
    {   // Refine staging: a candidate's rows copied in address order by its lane group (rows of >= 32 bytes), or every lane its own
        // row (16-byte rows: the 8 rows are one 128-byte line anyway).  Measured, head + replay: GIST-1M (128-byte rows) 0.644 ->
        // 0.628 ms in address order, SIFT-1M (16-byte rows) 0.497 -> 0.512.  RBQ_EXL_COALESCED=0/1 (read at load) overrides.
        const char* e = getenv("RBQ_EXL_COALESCED");
        d.exl_copy_coalesced = e != nullptr ? (atoi(e) != 0 ? 1u : 0u) : (d.exl_lane >= 32u ? 1u : 0u);
    }
    const size_t nvec = h->host.vec_off.empty() ? 0 : (size_t)h->host.vec_off.back();
    if (d.ex_bits == 0 || nvec == 0) return RBQ_OK;
    uint8_t* out = nullptr;
    const size_t bytes = nvec * (size_t)d.exl_stride;
    RBQ_CUDA(cudaMalloc(&out, bytes + 16));
    h->allocations.push_back(out);
    RBQ_CUDA(cudaMemset(out, 0, bytes + 16));
    const size_t work = nvec * (size_t)(d.D / 16);
    const unsigned tb = 256;
    const unsigned gb = (unsigned)((work + tb - 1) / tb);
    if (d.ex_bits == 2) relayout_ex_kernel<2><<<gb, tb>>>(d.ex, d.ex_stride, d.ex_bits, d.D, nvec, d.exl_lane, out);
    else if (d.ex_bits == 6) relayout_ex_kernel<6><<<gb, tb>>>(d.ex, d.ex_stride, d.ex_bits, d.D, nvec, d.exl_lane, out);
    else relayout_ex_kernel<1><<<gb, tb>>>(d.ex, d.ex_stride, d.ex_bits, d.D, nvec, d.exl_lane, out);
    RBQ_CUDA(cudaGetLastError());
    RBQ_CUDA(cudaDeviceSynchronize());
    d.exl = out;
    return RBQ_OK;
}

// ---- phased multi-GPU search: probe lists travel between shards without the shard-local list geometry ------------------
__global__ void probe_export_kernel(const Probe* __restrict__ probes, size_t first, size_t count, rbq_probe_rec* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Probe p = probes[first + i];
    out[first + i] = rbq_probe_rec{p.cid, p.g_add, p.g_error, p.dot_qc};
}
__global__ void probe_import_kernel(DevIndex ix, const rbq_probe_rec* __restrict__ in, size_t nq, uint32_t nprobe, Probe* __restrict__ probes,
                                    uint8_t* __restrict__ head_owner) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq * nprobe) return;
    const rbq_probe_rec r = in[i];
    Probe p;
    p.cid = r.cid;
    p.g_add = r.g_add;
    p.g_error = r.g_error;
    p.dot_qc = r.dot_qc;
    p.nv = ix.list_n[r.cid];
    p.blk_off = ix.blk_off[r.cid];
    p.vec_off = ix.vec_off[r.cid];
    probes[i] = p;
    if (i % nprobe == 0) head_owner[i / nprobe] = ix.list_owner == nullptr || ix.list_owner[r.cid] == (uint8_t)ix.shard_rank;
}
// Phased multi-GPU search: the queries whose nearest probed list this shard owns, compacted (any order), so that the head
// kernels run on dense grids; every other query gets the "nothing done here" head state (all its pairs are tail pairs).
__global__ void head_compact_kernel(const uint8_t* __restrict__ head_owner, uint32_t nq, uint32_t* __restrict__ qlist, uint32_t* __restrict__ qcount,
                                    uint32_t* __restrict__ out_counts, uint32_t* __restrict__ tail_start, float* __restrict__ tau) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool mine = q < nq && head_owner[q] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    uint32_t base = 0;
    if ((threadIdx.x & 31) == 0 && m) base = atomicAdd(qcount, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (mine) qlist[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = q;
    else if (q < nq) {
        out_counts[q] = 0u;
        tail_start[q] = 0u;
        tau[q] = INFINITY;
    }
}

int launch_probe_export(const Probe* d_probes, size_t q_begin, size_t q_count, size_t nprobe, rbq_probe_rec* d_out, cudaStream_t st) {
    const size_t count = q_count * nprobe;
    if (count == 0) return RBQ_OK;
    probe_export_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(d_probes, q_begin * nprobe, count, d_out);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}
int launch_probe_import(const DevIndex& ix, const rbq_probe_rec* d_in, size_t nq, size_t nprobe, Probe* d_probes, uint8_t* d_head_owner,
                        cudaStream_t st) {
    const size_t count = nq * nprobe;
    if (count == 0) return RBQ_OK;
    probe_import_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(ix, d_in, nq, (uint32_t)nprobe, d_probes, d_head_owner);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// ---- launchers ---------------------------------------------------------------------------------------------------
// limits of the CURRENT device, refreshed by every launcher (a process may drive several GPUs from several threads)
static thread_local int g_res_sms = 0;
static thread_local size_t g_res_smem_optin = 0;
static int res_limits() {
    int dev = 0, v = 0;
    RBQ_CUDA(cudaGetDevice(&dev));
    RBQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
    g_res_sms = v;
    RBQ_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    g_res_smem_optin = (size_t)v;
    return RBQ_OK;
}

// which refine form the resolve kernels use for this index (RBQ_REFINE_PAIRED=0/1 overrides the padded_dim rule)
static bool refine_paired(const DevIndex& ix) {
    static const int forced = [] {
        const char* e = getenv("RBQ_REFINE_PAIRED");
        return e ? atoi(e) : -1;
    }();
    return forced >= 0 ? forced != 0 : refine_paired_for((uint32_t)ix.D);
}

static void fill_args(ResolveArgs& a, const DevIndex& ix, const float* d_rot, const uint8_t* d_lut, const QueryScalars* d_qs,
                      const Probe* d_probes, size_t nq, size_t nprobe, size_t top_k, const uint64_t* d_filter, size_t filter_nbits,
                      uint64_t* d_ids, float* d_scores, uint32_t* d_counts, DevStats* d_stats, const TailWs& tw) {
    a.rot = d_rot;
    a.lut = d_lut;
    a.qs = d_qs;
    a.probes = d_probes;
    a.nq = (uint32_t)nq;
    a.q_begin = 0;
    a.q_count = (uint32_t)nq;
    a.cursor = 3;
    a.head_owner = nullptr;
    a.qlist = nullptr;
    a.qcount = nullptr;
    a.nprobe = (uint32_t)nprobe;
    a.top_k = (uint32_t)top_k;
    a.filter = reinterpret_cast<const unsigned long long*>(d_filter);
    a.filter_nbits = filter_nbits;
    a.out_ids = reinterpret_cast<unsigned long long*>(d_ids);
    a.out_scores = d_scores;
    a.out_counts = d_counts;
    a.stats = d_stats;
    a.counters = tw.counters;
    a.head_buf = tw.head_buf;
    a.head_cap = tw.head_cap;
    a.tail_start = tw.tail_start;
    a.tau = tw.tau;
    a.fb_list = tw.fb_list;
    a.fb2_list = tw.fb2_list;
    a.surv = tw.surv;
    a.surv_cnt = tw.surv_cnt;
    a.surv_cap = tw.surv_cap;
    a.sort_cap = tw.sort_cap;
    a.ovf_list = tw.ovf_list;
    a.ovf_recs = reinterpret_cast<XrRec*>(tw.ovf_recs);
    const bool paired = refine_paired(ix);
    a.exl_row = paired ? exl2_lane_stride((uint32_t)ix.D) : exl_row_stride((uint32_t)ix.D);
    a.rql_row = paired ? rql2_row_stride((uint32_t)ix.D) : rql_row_stride((uint32_t)ix.D);
    static const uint32_t stage_bufs = [] {
        const char* e = getenv("RBQ_STAGE_BUFS");
        return (uint32_t)(e && atoi(e) == 2 ? 2 : 1);
    }();
    a.stage_bufs = stage_bufs;
    a.has_ex = ix.ex_bits != 0;
    static const uint32_t flush_at = [] {
        const char* e = getenv("RBQ_FLUSH_AT");
        return (uint32_t)std::min(32, std::max(0, e ? atoi(e) : 0));  // 0: one refine round per flush (measured best)
    }();
    a.flush_at = flush_at ? flush_at : (paired ? kRefineSlots2 : kRefineSlots);
    static const uint32_t lazy_flush_at = [] {
        const char* e = getenv("RBQ_LAZY_FLUSH");
        return (uint32_t)std::min(32, std::max(0, e ? atoi(e) : 0));
    }();
    a.lazy_flush_at = lazy_flush_at ? lazy_flush_at : (paired ? kRefineSlots2 : kRefineSlots);
}

template <int NCB, bool WIDE>
static int launch_head_scan_ex(const DevIndex& ix, const ResolveArgs& a, cudaStream_t st) {
    head_scan_kernel<NCB, WIDE><<<(a.q_count + 3) / 4, 128, 0, st>>>(ix, a);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// grid of persistent CTAs for a kernel with `smem` bytes per CTA
static unsigned res_grid(size_t nq, size_t smem) {
    const size_t per_sm = std::max<size_t>(1, std::min<size_t>(16, (227 * 1024) / (smem + 1024)));
    return (unsigned)std::min<size_t>((nq + kResWarps - 1) / kResWarps, (size_t)g_res_sms * per_sm);
}

// kernel instances: <0> 1-bit index (nothing to refine), <1> eight-lane refine, <2> paired refine
#define RBQ_RES_LAUNCH(KERNEL, smem, grid)                                                                               \
    do {                                                                                                                  \
        if (ix.ex_bits == 0) {                                                                                            \
            RBQ_CUDA(cudaFuncSetAttribute(KERNEL<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)));           \
            KERNEL<0><<<grid, kResWarps * 32, smem, st>>>(ix, a);                                                         \
        } else if (refine_paired(ix)) {                                                                                   \
            RBQ_CUDA(cudaFuncSetAttribute(KERNEL<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)));           \
            KERNEL<2><<<grid, kResWarps * 32, smem, st>>>(ix, a);                                                         \
        } else {                                                                                                          \
            RBQ_CUDA(cudaFuncSetAttribute(KERNEL<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(smem)));           \
            KERNEL<1><<<grid, kResWarps * 32, smem, st>>>(ix, a);                                                         \
        }                                                                                                                 \
        RBQ_CUDA(cudaGetLastError());                                                                                     \
    } while (0)

int launch_head(const DevIndex& ix, const float* d_rot, const uint8_t* d_lut, const QueryScalars* d_qs, const Probe* d_probes,
                size_t nq, size_t nprobe, size_t top_k, const uint64_t* d_filter, size_t filter_nbits, uint64_t* d_ids,
                float* d_scores, uint32_t* d_counts, DevStats* d_stats, const TailWs& tw, cudaStream_t st, uint64_t* launches,
                size_t q_begin, size_t q_count, int* launch_index, const uint8_t* d_head_owner) {
    if (nq == 0 || q_count == 0) return RBQ_OK;
    int rc = res_limits();
    if (rc) return rc;
    ResolveArgs a;
    fill_args(a, ix, d_rot, d_lut, d_qs, d_probes, nq, nprobe, top_k, d_filter, filter_nbits, d_ids, d_scores, d_counts, d_stats, tw);
    a.head_owner = d_head_owner;
    if (d_head_owner != nullptr) {  // compact the queries of this shard: slots [q_begin, q_begin + q_count) of tw.qlist
        RBQ_CUDA(cudaMemsetAsync(tw.qcount, 0, 4, st));
        head_compact_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(d_head_owner, (uint32_t)nq, tw.qlist, tw.qcount, d_counts, tw.tail_start,
                                                                       tw.tau);
        RBQ_CUDA(cudaGetLastError());
        a.qlist = tw.qlist;
        a.qcount = tw.qcount;
        if (launches) *launches += 1;
    }
    const int ncb_lane = (ix.D / 4 + 31) / 32;
    const ResSmem w = res_smem_layout(a.exl_row, a.rql_row, a.top_k, ix.ex_bits != 0, true, 0, a.stage_bufs, refine_paired(ix));
    const size_t smem = (size_t)w.total * kResWarps;
    if (smem > g_res_smem_optin) return fail(RBQ_INVALID_CONFIG, "resolve kernel shared memory exceeds the device limit");
    // the dense head buffer holds tw.head_rows queries: the slice is walked in sub-chunks that reuse it (same stream)
    for (size_t s0 = q_begin; s0 < q_begin + q_count; s0 += tw.head_rows) {
        a.q_begin = (uint32_t)s0;
        a.q_count = (uint32_t)std::min<size_t>(tw.head_rows, q_begin + q_count - s0);
        // work cursor of this launch: the first kHeadCursors launches of a tile use pre-zeroed slots, later ones recycle
        const int li = launch_index ? (*launch_index)++ : 0;
        a.cursor = kTailCounters + (uint32_t)li % kHeadCursors;
        if (li >= (int)kHeadCursors) RBQ_CUDA(cudaMemsetAsync(tw.counters + a.cursor, 0, 4, st));
        if (ix.D > 1024) {
            rc = ncb_lane <= 12 ? launch_head_scan_ex<12, true>(ix, a, st) : launch_head_scan_ex<16, true>(ix, a, st);
        } else {
            switch (ncb_lane) {
                case 1: rc = launch_head_scan_ex<1, false>(ix, a, st); break;
                case 2: rc = launch_head_scan_ex<2, false>(ix, a, st); break;
                case 3: rc = launch_head_scan_ex<3, false>(ix, a, st); break;
                case 4: rc = launch_head_scan_ex<4, false>(ix, a, st); break;
                case 5:
                case 6: rc = launch_head_scan_ex<6, false>(ix, a, st); break;
                default: rc = launch_head_scan_ex<8, false>(ix, a, st); break;
            }
        }
        if (rc) return rc;
        const unsigned grid = res_grid(a.q_count, smem);
        RBQ_RES_LAUNCH(resolve_head_kernel, smem, grid);
        if (launches) *launches += 2;
    }
    return RBQ_OK;
}

// queries the lazy replay filed for the overflow tier (normally none: the kernel then exits at once)
static int launch_overflow_tier(const DevIndex& ix, const ResolveArgs& a, cudaStream_t st, uint64_t* launches) {
    static_assert(sizeof(XrRec) == 24, "overflow scratch is sized for 24-byte records");
    const bool paired = refine_paired(ix);
    const OvfSmem w = ovf_smem_layout(a.exl_row, a.rql_row, a.top_k, ix.ex_bits != 0, paired, a.stage_bufs);
    const size_t smem = w.total;
    if (smem > g_res_smem_optin) return fail(RBQ_INVALID_CONFIG, "overflow-tier shared memory exceeds the device limit");
    const unsigned grid = std::min<unsigned>((unsigned)g_res_sms, kOvfCtas);
    if (ix.ex_bits == 0) {
        RBQ_CUDA(cudaFuncSetAttribute(overflow_replay_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        overflow_replay_kernel<0><<<grid, kOvfWarps * 32, smem, st>>>(ix, a);
    } else if (paired) {
        RBQ_CUDA(cudaFuncSetAttribute(overflow_replay_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        overflow_replay_kernel<2><<<grid, kOvfWarps * 32, smem, st>>>(ix, a);
    } else {
        RBQ_CUDA(cudaFuncSetAttribute(overflow_replay_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        overflow_replay_kernel<1><<<grid, kOvfWarps * 32, smem, st>>>(ix, a);
    }
    RBQ_CUDA(cudaGetLastError());
    if (launches) *launches += 1;
    return RBQ_OK;
}

int launch_refine_replay(const DevIndex& ix, const float* d_rot, const QueryScalars* d_qs, const Probe* d_probes, size_t nq,
                         size_t nprobe, size_t top_k, uint64_t* d_ids, float* d_scores, uint32_t* d_counts, DevStats* d_stats,
                         const TailWs& tw, cudaStream_t st, uint64_t* launches) {
    if (nq == 0) return RBQ_OK;
    int rc = res_limits();
    if (rc) return rc;
    ResolveArgs a;
    fill_args(a, ix, d_rot, nullptr, d_qs, d_probes, nq, nprobe, top_k, nullptr, 0, d_ids, d_scores, d_counts, d_stats, tw);
    if (ix.ex_bits != 0) {
        // sorted survivors, refinement on demand against the live threshold (one kernel)
        const ResSmem w = res_smem_layout(a.exl_row, a.rql_row, a.top_k, true, true, a.sort_cap, a.stage_bufs, refine_paired(ix));
        const size_t smem = (size_t)w.total * kResWarps;
        if (smem > g_res_smem_optin) return fail(RBQ_INVALID_CONFIG, "replay kernel shared memory exceeds the device limit");
        const unsigned grid = res_grid(nq, smem);
        if (refine_paired(ix)) {
            RBQ_CUDA(cudaFuncSetAttribute(resolve_lazy_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            resolve_lazy_kernel<2><<<grid, kResWarps * 32, smem, st>>>(ix, a);
        } else {
            RBQ_CUDA(cudaFuncSetAttribute(resolve_lazy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            resolve_lazy_kernel<1><<<grid, kResWarps * 32, smem, st>>>(ix, a);
        }
        RBQ_CUDA(cudaGetLastError());
        if (launches) *launches += 1;
        return launch_overflow_tier(ix, a, st, launches);
    }
    const ResSmem w = res_smem_layout(0, 0, a.top_k, false, true, a.sort_cap);
    const size_t smem = (size_t)w.total * kResWarps;
    if (smem > g_res_smem_optin) return fail(RBQ_INVALID_CONFIG, "replay kernel shared memory exceeds the device limit");
    const unsigned grid = res_grid(nq, smem);
    RBQ_RES_LAUNCH(resolve_replay_kernel, smem, grid);
    if (launches) *launches += 1;
    return launch_overflow_tier(ix, a, st, launches);
}

int launch_ex_dot_debug(const DevIndex& ix, const float* d_rot, const unsigned long long* d_gv, int n, float* d_out, const TailWs& tw, cudaStream_t st) {
    if (n == 0 || ix.ex_bits == 0) return RBQ_OK;
    int rc = res_limits();
    if (rc) return rc;
    ResolveArgs a;
    fill_args(a, ix, d_rot, nullptr, nullptr, nullptr, 1, 1, 1, nullptr, 0, nullptr, nullptr, nullptr, nullptr, tw);
    const ResSmem w = res_smem_layout(a.exl_row, a.rql_row, 1, true, false, 0, a.stage_bufs, refine_paired(ix));
    if (refine_paired(ix)) {
        RBQ_CUDA(cudaFuncSetAttribute(ex_dot_debug_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.total));
        ex_dot_debug_kernel<true><<<1, 32, w.total, st>>>(ix, a, d_gv, n, d_out);
    } else {
        RBQ_CUDA(cudaFuncSetAttribute(ex_dot_debug_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.total));
        ex_dot_debug_kernel<false><<<1, 32, w.total, st>>>(ix, a, d_gv, n, d_out);
    }
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
