// exact_merge.cu -- bit-identical multi-GPU results: the global replay of the one-call sharded search (api.cu,
// rbq_search_batch_sharded_device with rbq_set_exact_merge).
//
// The phased search lets every shard replay its own survivors against min(tau, local k-th distance).  That threshold is never
// below the single sequence's, so a shard may admit a candidate the reference skipped (lower bound >= the global threshold at
// that point); when such a candidate's refined distance violates its lower bound it can enter the merged top-k -- the only way
// a sharded answer differs from the reference's (class D2: ~4 ids in 10 000 at GIST-1M, ~1 in 1 000 at
// 100M x 128).  The global replay removes it: every shard refines ALL its survivors eagerly (the tail kernel's survivor set
// under the all-reduced head threshold is exactly the single-GPU survivor set, split by list owner), ships them -- with the
// head pass' heap state from the shard that ran it -- to the query's HOME rank (the rank that selected its probes), and the
// home rank replays the union in the reference's visit order (probe rank, position) against the live threshold:
// search_cluster_v2_batched's decisions (src/ivf.rs:2013-2127), bit for bit, with one replay per query instead of one per
// (query, shard).
//
//   xr_count      records per query on this shard (survivors + head state where the head pass ran here); overflowed
//                 survivor buffers are announced with a marker so that every rank drops the query from the exact path
//   xr_scan       exclusive scan -> packed offsets, totals per destination rank
//   xr_records    eager refinement (K10 through resolve.cu's refine path) + packing: {visit key, lower bound, distance, id}
//   (NCCL)        all-gather of the counts, grouped send/recv of the record ranges (sizes read back once: the call syncs here)
//   xr_plan       per source shard: where each home query's records sit in the receive buffer
//   xr_replay     sort by visit key, sequential replay with the precomputed distances -> final top-k of the home queries
//   (NCCL)        all-gather of the home results;  xr_override writes them over the phased search's merged answer
// Queries the exact path cannot take (survivor overflow on some shard, head pass that did not fill the heap, more than
// kMaxRecs records) keep the phased answer and are counted in rbq_search_stats::inexact_queries.
#include <algorithm>

#include "scan_common.cuh"

namespace rbq {

constexpr uint32_t kXrMarker = 0xffffffffu;  // count slot of a query whose survivors overflowed on this shard
constexpr int kXrMaxRecs = 4096;             // records one home query may gather (sort buffer in shared memory)

// ---- counts + scan ---------------------------------------------------------------------------------------------------
__global__ void xr_count_kernel(const uint32_t* __restrict__ surv_cnt, uint32_t surv_cap, const uint8_t* __restrict__ head_owner,
                                const uint32_t* __restrict__ head_cnt, const float* __restrict__ tau, uint32_t nq, uint32_t* __restrict__ cnt) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint32_t s = surv_cnt[q];
    uint32_t c;
    if (s > surv_cap || !(tau[q] < INFINITY)) c = kXrMarker;  // overflow here, or nobody filled the heap in the head pass
    else c = s + (head_owner[q] ? head_cnt[q] : 0u);
    cnt[q] = c;
}
// one CTA per 1024 queries: local exclusive scan + CTA total; second kernel adds the CTA bases
__global__ void __launch_bounds__(1024) xr_scan_local_kernel(const uint32_t* __restrict__ cnt, uint32_t nq, uint32_t* __restrict__ off,
                                                            uint32_t* __restrict__ cta_tot) {
    __shared__ uint32_t s_w[64];
    const uint32_t q = blockIdx.x * 1024u + threadIdx.x, lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t v = 0;
    if (q < nq) {
        v = cnt[q];
        if (v == kXrMarker) v = 0;
    }
    uint32_t x = v;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) s_w[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t y = s_w[lane];
#pragma unroll
        for (uint32_t o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, y, o);
            if (lane >= o) y += t;
        }
        s_w[32 + lane] = y;
    }
    __syncthreads();
    const uint32_t base = wid ? s_w[32 + wid - 1] : 0u;
    if (q < nq) off[q] = base + x - v;
    if (threadIdx.x == 1023) cta_tot[blockIdx.x] = base + x;
}
__global__ void __launch_bounds__(1024) xr_scan_fix_kernel(uint32_t nq, uint32_t* __restrict__ off, const uint32_t* __restrict__ cta_tot,
                                                          uint32_t nctas, uint32_t per, uint32_t world, unsigned long long* __restrict__ send_tot) {
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0) {
        uint32_t b = 0;
        for (uint32_t i = 0; i < blockIdx.x; ++i) b += cta_tot[i];
        s_base = b;
    }
    __syncthreads();
    const uint32_t q = blockIdx.x * 1024u + threadIdx.x;
    if (q < nq) off[q] += s_base;
    if (q == nq - 1 || (q < nq && false)) {}
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {  // grand total behind the last query
        uint32_t t = 0;
        for (uint32_t i = 0; i < nctas; ++i) t += cta_tot[i];
        off[nq] = t;
    }
    (void)per;
    (void)world;
    (void)send_tot;
}
// totals per destination rank from the finished offsets: send_tot[d] = off[min((d+1)*per, nq)] - off[min(d*per, nq)]
__global__ void xr_send_totals_kernel(const uint32_t* __restrict__ off, uint32_t nq, uint32_t per, uint32_t world,
                                      unsigned long long* __restrict__ send_tot) {
    const uint32_t d = threadIdx.x;
    if (d >= world) return;
    const uint32_t b = min(d * per, nq), e = min((d + 1) * per, nq);
    send_tot[d] = (unsigned long long)(off[e] - off[b]);
}

// ---- records ---------------------------------------------------------------------------------------------------------
struct XrArgs {
    const float* rot;
    const QueryScalars* qs;
    const Probe* probes;
    uint32_t nq, nprobe, top_k;
    const uint32_t* cnt;        // xr_count
    const uint32_t* off;        // packed offsets
    const Survivor* surv;
    const uint32_t* surv_cnt;
    uint32_t surv_cap;
    const uint8_t* head_owner;
    const unsigned long long* head_ids;  // head pass' heap state (snapshot), [nq][k]
    const float* head_sc;
    const uint32_t* head_cnt;
    XrRec* recs;
    uint32_t exl_row, rql_row, stage_bufs, has_ex;
    uint32_t* cursor;
};

__global__ void __launch_bounds__(128) xr_records_kernel(DevIndex ix, XrArgs a) {
    extern __shared__ __align__(16) unsigned char xr_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = ix.D, k = (int)a.top_k;
    const uint32_t per_warp = 32u * a.exl_row + 4u * a.rql_row;
    unsigned char* wbase = xr_smem + (size_t)warp * per_warp;
    const uint32_t stage_u32 = smem_u32(wbase), rql_u32 = smem_u32(wbase + 32u * a.exl_row);
    unsigned char* rql = wbase + 32u * a.exl_row;
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    for (;;) {
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(a.cursor, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= a.nq) break;
        const uint32_t c = a.cnt[q];
        if (c == kXrMarker || c == 0u) continue;
        XrRec* out = a.recs + a.off[q];
        uint32_t nhead = 0;
        if (a.head_owner[q]) {  // the heap after the head pass, best first: re-inserted unconditionally at the home rank
            nhead = a.head_cnt[q];
            for (uint32_t i = lane; i < nhead; i += 32) {
                const float sc = a.head_sc[(size_t)q * k + i];
                out[i] = XrRec{(unsigned long long)i, -INFINITY, l2 ? sc : -sc, a.head_ids[(size_t)q * k + i]};
            }
        }
        const uint32_t ns = c - nhead;
        if (ns == 0) continue;
        const Probe* pr = a.probes + (size_t)q * a.nprobe;
        const Survivor* sv = a.surv + (size_t)q * a.surv_cap;
        QueryScalars s{};
        if (a.has_ex) {
            s = a.qs[q];
            __syncwarp();
            load_rql2(rql, a.rql_row, a.rot + (size_t)q * D, D, ix.exl_lane, lane);  // rotated query in paired chain order
            __syncwarp();
        }
        for (uint32_t b0 = 0; b0 < ns; b0 += 32) {
            const uint32_t i = b0 + lane;
            const int m = (int)min(32u, ns - b0);
            Survivor rec = {0u, 0u, 0.0f, 0.0f};
            unsigned long long gv = 0;
            float g_add = 0.0f;
            if (i < ns) {
                rec = sv[i];
                const Probe* pp = pr + rec.rank;
                gv = pp->vec_off + rec.pos;
                g_add = pp->g_add;
            }
            float dist = rec.x;  // 1-bit index: the estimate is the distance
            if (a.has_ex) {
                const float exdot = refine_batch2(ix, gv, m, stage_u32, rql_u32, a.exl_row, a.rql_row, 1u, lane);
                if (i < ns) {
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
            if (i < ns)
                out[nhead + i] = XrRec{((unsigned long long)(rec.rank + 1u) << 32) | rec.pos, rec.lower, dist, ix.ids[gv]};
        }
    }
}

// ---- home side ---------------------------------------------------------------------------------------------------------
// all-gathered counts C[s][q] -> per source: records of my slice (recv_tot[s]) and each home query's offset inside that source's
// range (qoff[s][j]); flagged[j] = some shard announced a marker.  One CTA per source shard.
__global__ void __launch_bounds__(1024) xr_plan_kernel(const uint32_t* __restrict__ all_cnt, uint32_t nq, uint32_t q_begin, uint32_t q_count,
                                                      uint32_t* __restrict__ qoff, unsigned long long* __restrict__ recv_tot,
                                                      uint32_t* __restrict__ flagged) {
    __shared__ uint32_t s_w[64];
    __shared__ uint32_t s_run;
    const uint32_t s = blockIdx.x, lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t* c = all_cnt + (size_t)s * nq + q_begin;
    uint32_t* o = qoff + (size_t)s * q_count;
    if (threadIdx.x == 0) s_run = 0;
    __syncthreads();
    for (uint32_t j0 = 0; j0 < q_count; j0 += 1024u) {
        const uint32_t j = j0 + threadIdx.x;
        uint32_t v = 0;
        if (j < q_count) {
            v = c[j];
            if (v == kXrMarker) {
                v = 0;
                atomicOr(&flagged[j], 1u);
            }
        }
        uint32_t x = v;
#pragma unroll
        for (uint32_t t = 1; t < 32; t <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, x, t);
            if (lane >= t) x += u;
        }
        if (lane == 31) s_w[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t y = s_w[lane];
#pragma unroll
            for (uint32_t t = 1; t < 32; t <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, y, t);
                if (lane >= t) y += u;
            }
            s_w[32 + lane] = y;
        }
        __syncthreads();
        const uint32_t base = s_run + (wid ? s_w[32 + wid - 1] : 0u);
        if (j < q_count) o[j] = base + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_run = base + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) recv_tot[s] = s_run;
}

struct XrReplayArgs {
    const XrRec* recv;               // received records, source ranges back to back
    const unsigned long long* rbase; // [world] first record of every source's range
    const uint32_t* all_cnt;         // [world][nq]
    const uint32_t* qoff;            // [world][q_count]
    const uint32_t* flagged;         // [q_count]
    uint32_t nq, q_begin, q_count, world, top_k, metric;
    unsigned long long* out_ids;     // home results [q_count][k]
    float* out_sc;
    uint32_t* out_cn;                // count, or kXrMarker = "keep the phased answer"
    DevStats* stats;
    unsigned long long* inexact;     // device counter
};

// one CTA per home query: gather (key, index) pairs, bitonic sort, warp 0 replays in visit order
__global__ void __launch_bounds__(128) xr_replay_kernel(XrReplayArgs a) {
    extern __shared__ __align__(16) unsigned char xr_smem[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(xr_smem);                 // [npad] key
    uint32_t* ridx = reinterpret_cast<uint32_t*>(keys + kXrMaxRecs);                           // [npad] record index in recv
    float* sd = reinterpret_cast<float*>(ridx + kXrMaxRecs);                                   // [k]
    unsigned long long* si = reinterpret_cast<unsigned long long*>(sd + ((a.top_k + 1) & ~1u));  // [k]
    __shared__ uint32_t s_src_off[kMaxShards + 1];
    const uint32_t j = blockIdx.x, tid = threadIdx.x, lane = tid & 31u;
    const uint32_t q = a.q_begin + j;
    const int k = (int)a.top_k;
    if (tid == 0) {
        uint32_t run = 0;
        for (uint32_t s = 0; s < a.world; ++s) {
            s_src_off[s] = run;
            const uint32_t c = a.all_cnt[(size_t)s * a.nq + q];
            run += c == kXrMarker ? 0u : c;
        }
        s_src_off[a.world] = run;
    }
    __syncthreads();
    const uint32_t m = s_src_off[a.world];
    if (a.flagged[j] || m > (uint32_t)kXrMaxRecs) {
        if (tid == 0) {
            a.out_cn[j] = kXrMarker;
            atomicAdd(a.inexact, 1ull);
        }
        return;
    }
    uint32_t npad = 32;
    while (npad < m) npad <<= 1;
    for (uint32_t s = 0; s < a.world; ++s) {
        const uint32_t c = s_src_off[s + 1] - s_src_off[s];
        const unsigned long long first = a.rbase[s] + a.qoff[(size_t)s * a.q_count + j];
        for (uint32_t i = tid; i < c; i += 128u) {
            keys[s_src_off[s] + i] = a.recv[first + i].key;
            ridx[s_src_off[s] + i] = (uint32_t)(first + i);
        }
    }
    for (uint32_t i = m + tid; i < npad; i += 128u) {
        keys[i] = ~0ull;
        ridx[i] = 0u;
    }
    __syncthreads();
    for (uint32_t size = 2; size <= npad; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            for (uint32_t t = tid; t < npad / 2; t += 128u) {
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
    if (tid >= 32) return;
    // warp 0: the reference's loop over the candidates in visit order (src/ivf.rs:2044-2127) with precomputed distances
    TopK tk;
    tk.init(sd, si, k);
    unsigned long long adm = 0;
    for (uint32_t b0 = 0; b0 < m; b0 += 32) {
        const uint32_t i = b0 + lane;
        XrRec r{0ull, 0.0f, 0.0f, 0ull};
        if (i < m) r = a.recv[ridx[i]];
        const float th0 = tk.theta();
        unsigned mask = __ballot_sync(0xffffffffu, i < m && r.lower < th0);  // stale threshold: superset of what the live one admits
        while (mask) {
            const int sl = __ffs(mask) - 1;
            mask &= mask - 1;
            const float lb_s = __shfl_sync(0xffffffffu, r.lower, sl);
            const float d_s = __shfl_sync(0xffffffffu, r.dist, sl);
            const unsigned long long id_s = __shfl_sync(0xffffffffu, r.id, sl);
            const float theta = tk.theta();
            if (lb_s >= theta) continue;  // skipped_by_lower_bound
            if (lb_s > -INFINITY) adm += 1;  // head-state records were admitted (and counted) by the head pass
            if (!isfinite(d_s)) continue;
            tk.insert(d_s, id_s, (int)lane);
        }
    }
    const bool l2 = a.metric == RBQ_METRIC_L2;
    for (int i = (int)lane; i < k; i += 32) {
        const bool have = i < tk.cnt;
        const float dv = tk.dist_at(i);
        a.out_ids[(size_t)j * k + i] = have ? tk.id_at(i) : ~0ull;
        a.out_sc[(size_t)j * k + i] = have ? (l2 ? dv : -dv) : 0.0f;
    }
    if (lane == 0) a.out_cn[j] = (uint32_t)tk.cnt;
    (void)adm;
}

// all-gathered home results (one chunk per rank: ids | scores | counts of its slice) over the phased answer
__global__ void xr_override_kernel(const char* __restrict__ gath, size_t chunk, uint32_t per, uint32_t nq, uint32_t k, unsigned long long* __restrict__ ids,
                                   float* __restrict__ sc, uint32_t* __restrict__ cn) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    const uint32_t r = q / per, j = q - r * per;
    const char* base = gath + (size_t)r * chunk;
    const unsigned long long* gi = reinterpret_cast<const unsigned long long*>(base);
    const float* gs = reinterpret_cast<const float*>(base + (size_t)per * k * 8);
    const uint32_t* gc = reinterpret_cast<const uint32_t*>(base + (size_t)per * k * 12);
    const uint32_t c = gc[j];
    if (c == kXrMarker) return;  // not taken by the exact path: the phased answer stays
    for (uint32_t i = 0; i < k; ++i) {
        ids[(size_t)q * k + i] = gi[(size_t)j * k + i];
        sc[(size_t)q * k + i] = gs[(size_t)j * k + i];
    }
    cn[q] = c;
}

// ---- launchers -------------------------------------------------------------------------------------------------------
int xr_launch_count_scan(const TailWs& tw, const uint8_t* d_head_owner, const uint32_t* d_head_cnt, size_t nq, uint32_t per, uint32_t world,
                         uint32_t* d_cnt, uint32_t* d_off, uint32_t* d_cta_tot, unsigned long long* d_send_tot, cudaStream_t st) {
    const unsigned gb = (unsigned)((nq + 255) / 256), nctas = (unsigned)((nq + 1023) / 1024);
    xr_count_kernel<<<gb, 256, 0, st>>>(tw.surv_cnt, tw.surv_cap, d_head_owner, d_head_cnt, tw.tau, (uint32_t)nq, d_cnt);
    xr_scan_local_kernel<<<nctas, 1024, 0, st>>>(d_cnt, (uint32_t)nq, d_off, d_cta_tot);
    xr_scan_fix_kernel<<<nctas, 1024, 0, st>>>((uint32_t)nq, d_off, d_cta_tot, nctas, per, world, d_send_tot);
    xr_send_totals_kernel<<<1, 32, 0, st>>>(d_off, (uint32_t)nq, per, world, d_send_tot);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int xr_launch_records(const DevIndex& ix, const float* d_rot, const QueryScalars* d_qs, const Probe* d_probes, size_t nq, size_t nprobe, size_t top_k,
                      const TailWs& tw, const uint8_t* d_head_owner, const uint64_t* d_head_ids, const float* d_head_sc, const uint32_t* d_head_cnt,
                      const uint32_t* d_cnt, const uint32_t* d_off, XrRec* d_recs, uint32_t* d_cursor, cudaStream_t st) {
    XrArgs a;
    a.rot = d_rot;
    a.qs = d_qs;
    a.probes = d_probes;
    a.nq = (uint32_t)nq;
    a.nprobe = (uint32_t)nprobe;
    a.top_k = (uint32_t)top_k;
    a.cnt = d_cnt;
    a.off = d_off;
    a.surv = tw.surv;
    a.surv_cnt = tw.surv_cnt;
    a.surv_cap = tw.surv_cap;
    a.head_owner = d_head_owner;
    a.head_ids = reinterpret_cast<const unsigned long long*>(d_head_ids);
    a.head_sc = d_head_sc;
    a.head_cnt = d_head_cnt;
    a.recs = d_recs;
    a.exl_row = exl2_lane_stride((uint32_t)ix.D);
    a.rql_row = rql2_row_stride((uint32_t)ix.D);
    a.stage_bufs = 1;
    a.has_ex = ix.ex_bits != 0;
    a.cursor = d_cursor;
    RBQ_CUDA(cudaMemsetAsync(d_cursor, 0, 4, st));
    const size_t smem = (size_t)4 * (32u * a.exl_row + 4u * a.rql_row);
    RBQ_CUDA(cudaFuncSetAttribute(xr_records_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 0;
    RBQ_CUDA(cudaGetDevice(&dev));
    RBQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t per_sm = std::max<size_t>(1, std::min<size_t>(12, (200 * 1024) / (smem + 1024)));
    const unsigned grid = (unsigned)std::min<size_t>((nq + 3) / 4, (size_t)sms * per_sm);
    xr_records_kernel<<<grid, 128, smem, st>>>(ix, a);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int xr_launch_plan(const uint32_t* d_all_cnt, size_t nq, size_t q_begin, size_t q_count, int world, uint32_t* d_qoff, unsigned long long* d_recv_tot,
                   uint32_t* d_flagged, cudaStream_t st) {
    RBQ_CUDA(cudaMemsetAsync(d_flagged, 0, std::max<size_t>(q_count, 1) * 4, st));
    xr_plan_kernel<<<world, 1024, 0, st>>>(d_all_cnt, (uint32_t)nq, (uint32_t)q_begin, (uint32_t)q_count, d_qoff, d_recv_tot, d_flagged);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int xr_launch_replay(const XrRec* d_recv, const unsigned long long* d_rbase, const uint32_t* d_all_cnt, const uint32_t* d_qoff, const uint32_t* d_flagged,
                     size_t nq, size_t q_begin, size_t q_count, int world, size_t top_k, int metric, uint64_t* d_out_ids, float* d_out_sc,
                     uint32_t* d_out_cn, unsigned long long* d_inexact, cudaStream_t st) {
    if (q_count == 0) return RBQ_OK;
    XrReplayArgs a;
    a.recv = d_recv;
    a.rbase = d_rbase;
    a.all_cnt = d_all_cnt;
    a.qoff = d_qoff;
    a.flagged = d_flagged;
    a.nq = (uint32_t)nq;
    a.q_begin = (uint32_t)q_begin;
    a.q_count = (uint32_t)q_count;
    a.world = (uint32_t)world;
    a.top_k = (uint32_t)top_k;
    a.metric = (uint32_t)metric;
    a.out_ids = reinterpret_cast<unsigned long long*>(d_out_ids);
    a.out_sc = d_out_sc;
    a.out_cn = d_out_cn;
    a.stats = nullptr;
    a.inexact = d_inexact;
    const size_t smem = (size_t)kXrMaxRecs * 12 + ((top_k + 1) & ~(size_t)1) * 4 + top_k * 8 + 16;
    RBQ_CUDA(cudaFuncSetAttribute(xr_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    xr_replay_kernel<<<(unsigned)q_count, 128, smem, st>>>(a);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int xr_launch_override(const void* d_gath, size_t chunk, size_t per, size_t nq, size_t top_k, uint64_t* d_ids, float* d_sc, uint32_t* d_cn, cudaStream_t st) {
    xr_override_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, st>>>(static_cast<const char*>(d_gath), chunk, (uint32_t)per, (uint32_t)nq, (uint32_t)top_k,
                                                                    reinterpret_cast<unsigned long long*>(d_ids), d_sc, d_cn);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
