// scan.cu -- the hot path: one warp walks one query's probed lists in the reference's order.
//
//   K7   FastScan accumulate   simd::accumulate_batch_avx2 / _scalar   (reference src/simd.rs:972-1184, 1462-1525)
//   K8   batch distances       simd::compute_batch_distances_u16 (AVX2) (reference src/simd.rs:2090-2140)
//   K9   prune + refine        search_cluster_v2_batched                (reference src/ivf.rs:2013-2127)
//   K10  packed ex-code dot    ip_packed_ex2_f32 / ip_packed_ex6_f32 (AVX2 lane order) (src/simd.rs:1722-1825)
//   K11  top-k                 BinaryHeap<HeapEntry> keep-k-smallest    (reference src/ivf.rs:1844, 2116-2126)
//
// Data layout consumed as stored in the index file: a 32-vector block is
// 4*D code bytes followed by f_add[32], f_rescale[32], f_error[32].  The 16 code bytes at offset
// 16*cb belong to codebook cb (dims 4cb..4cb+3) and so do the 16 LUT bytes at the same offset:
// lane l of the warp owns codebooks l, l+32, ... -> one coalesced 128-bit load per lane per 512 B of
// block, its LUT rows live in registers for the whole query, and the 16-entry byte lookup is two
// PRMTs (entries 0-7 / 8-15) blended by a PRMT-generated mask (the GPU analogue of pshufb).  Per-lane
// partial sums are kept as packed u16 pairs and reduced across lanes with a 16-shuffle
// reduce-scatter that leaves lane v holding accu[v]; the block's factors are then read by lane v.
//
// Exact pruning order: lanes whose lower bound beats the threshold at block entry
// (a superset of what the reference admits) are refined in parallel, then replayed in lane order
// against the live threshold, which reproduces the reference's sequential decisions exactly.
//
// Integer sums are exact (== the reference's u16 wrap-around value); float ops follow the AVX2
// variants' order (fma only where the reference uses fmadd).  Compiled with -fmad=false.
#include <cfloat>

#include "rbq_internal.h"

namespace rbq {

constexpr int kWarps = 4;
constexpr int kMaxTopK = 1024;
size_t scan_max_topk() { return kMaxTopK; }

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint4 ldg128(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// 32 nibble lookups of one codebook.  C = the 16 code bytes (4 regs), T = the codebook's 16 LUT bytes.
// E[m] accumulates (vector m | vector m+8 << 16), O[m] (vector m+16 | vector m+24 << 16), m = 0..7.
__device__ __forceinline__ void lookup_accumulate(const uint4& Cv, const uint4& T, uint32_t (&E)[8], uint32_t (&O)[8]) {
    const uint32_t C[4] = {Cv.x, Cv.y, Cv.z, Cv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t c = C[k];
        const uint32_t s = c & 0x77777777u;  // 3-bit byte selectors (bit 3 of a PRMT selector = sign mode)
        const uint32_t sh = c << 4;          // moves the low nibbles' msb into byte-sign position
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t sel = h ? (s >> 16) : s;
            const uint32_t lo = prmt(T.x, T.y, sel);                    // entries 0..7
            const uint32_t hi = prmt(T.z, T.w, sel);                    // entries 8..15
            const uint32_t m = prmt(c, sh, h ? 0xBFAEu : 0x9D8Cu);      // 0xFF where the nibble's msb is set
            const uint32_t r = (lo & ~m) | (hi & m);                    // 4 looked-up bytes
            E[2 * k + h] += prmt(r, 0u, 0x4240u);                       // bytes 0,2 -> u16 pair
            O[2 * k + h] += prmt(r, 0u, 0x4341u);                       // bytes 1,3 -> u16 pair
        }
    }
}

// Cross-lane reduce-scatter: on return lane v holds sum over lanes of the partial for vector v.
template <bool WIDE>
__device__ __forceinline__ uint32_t reduce_scatter(uint32_t (&E)[8], uint32_t (&O)[8], int lane) {
    const unsigned full = 0xffffffffu;
    if (!WIDE) {
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2, b0 = lane & 1;
        uint32_t X[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            const uint32_t keep = b4 ? O[m] : E[m], send = b4 ? E[m] : O[m];
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
        uint32_t x[32];
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            x[m] = E[m] & 0xffffu;
            x[m + 8] = E[m] >> 16;
            x[m + 16] = O[m] & 0xffffu;
            x[m + 24] = O[m] >> 16;
        }
#pragma unroll
        for (int w = 16; w >= 1; w >>= 1) {
            const bool up = lane & w;
#pragma unroll
            for (int i = 0; i < w; ++i) {
                const uint32_t keep = up ? x[i + w] : x[i], send = up ? x[i] : x[i + w];
                x[i] = keep + __shfl_xor_sync(full, send, w);
            }
        }
        return x[0];
    }
}

// K7 for one block: returns accu[lane] (exact integer sum, before the u16 wrap)
template <int NCB, bool WIDE>
__device__ __forceinline__ uint32_t accumulate_block(const uint8_t* __restrict__ blk, const uint4 (&T)[NCB], int ncb,
                                                     int lane) {
    uint32_t E[8], O[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) E[m] = O[m] = 0u;
    uint4 C[NCB];
#pragma unroll
    for (int i = 0; i < NCB; ++i) {
        const int cb = lane + 32 * i;
        C[i] = (cb < ncb) ? ldg128(blk + 16 * cb) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < NCB; ++i) lookup_accumulate(C[i], T[i], E, O);
    return reduce_scatter<WIDE>(E, O, lane);
}

// K10: one of the 8 "AVX lanes" (j) of the packed ex-code dot product: dims j, j+8, j+16, ... with fma.
// EXK: 2 / 6 = the reference's C++-compatible layouts; 1 = generic LSB-first bit stream (extension).
template <int EXK>
__device__ __forceinline__ float ex_dot_lane(const uint8_t* __restrict__ p, const float* __restrict__ rq, int D, int j,
                                             int ex_bits) {
    float acc = 0.0f;
    const int sh_lo = 8 * (j & 3) + 2 * (j >> 2);  // bit position of code j in the 2-bit word; code j+8: +4
    if (EXK == 2) {
        for (int c = 0; c < D / 16; ++c) {
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p + 4 * c));
            acc = __fmaf_rn((float)((w >> sh_lo) & 3u), rq[16 * c + j], acc);
            acc = __fmaf_rn((float)((w >> (sh_lo + 4)) & 3u), rq[16 * c + 8 + j], acc);
        }
    } else if (EXK == 6) {
        for (int c = 0; c < D / 16; ++c) {
            const uint32_t b = __ldg(p + 12 * c + j);
            const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(p + 12 * c + 8));
            const uint32_t c0 = (b & 15u) | (((w >> sh_lo) & 3u) << 4);
            const uint32_t c1 = (b >> 4) | (((w >> (sh_lo + 4)) & 3u) << 4);
            acc = __fmaf_rn((float)c0, rq[16 * c + j], acc);
            acc = __fmaf_rn((float)c1, rq[16 * c + 8 + j], acc);
        }
    } else {
        const uint32_t mask = (1u << ex_bits) - 1u;
        for (int d = j; d < D; d += 8) {
            const uint32_t pos = (uint32_t)d * (uint32_t)ex_bits;
            const uint32_t two = (uint32_t)__ldg(p + (pos >> 3)) | ((uint32_t)__ldg(p + (pos >> 3) + 1) << 8);
            acc = __fmaf_rn((float)((two >> (pos & 7u)) & mask), rq[d], acc);
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

// K11: warp-cooperative insertion into an ascending list of at most k (distance, id) pairs.
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
};

template <int NCB, int EXK, bool WIDE>
__global__ void __launch_bounds__(kWarps * 32) scan_kernel(DevIndex ix, ScanArgs a) {
    extern __shared__ __align__(16) unsigned char scan_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * kWarps + warp;
    if (q >= a.nq) return;  // whole warp exits together; no block-level sync below
    const int D = ix.D, ncb = D / 4, k = (int)a.top_k;
    // per-warp shared memory: top-k ids (8B) | top-k distances | rotated query
    unsigned long long* si = reinterpret_cast<unsigned long long*>(scan_smem) + (size_t)warp * k;
    float* sd = reinterpret_cast<float*>(scan_smem + (size_t)kWarps * k * 8) + (size_t)warp * k;
    float* rq = reinterpret_cast<float*>(scan_smem + (size_t)kWarps * k * 12) + (size_t)warp * D;
    if (EXK != 0)
        for (int i = lane; i < D; i += 32) rq[i] = a.rot[(size_t)q * D + i];
    __syncwarp();

    uint4 T[NCB];
#pragma unroll
    for (int i = 0; i < NCB; ++i) {
        const int cb = lane + 32 * i;
        T[i] = (cb < ncb) ? ldg128(a.lut + (size_t)q * D * 4 + 16 * cb) : make_uint4(0, 0, 0, 0);
    }
    const QueryScalars s = a.qs[q];
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    int cnt = 0;
    unsigned long long st_blocks = 0, st_cand = 0, st_ref = 0, st_adm = 0;

    for (uint32_t pi = 0; pi < a.nprobe; ++pi) {
        const Probe pr = a.probes[(size_t)q * a.nprobe + pi];
        const uint32_t nv = ix.list_n[pr.cid];
        if (nv == 0) continue;  // empty list, or a list owned by another shard
        const uint8_t* base = ix.blocks + (size_t)ix.blk_off[pr.cid] * ix.block_stride;
        const unsigned long long vbase = ix.vec_off[pr.cid];
        const uint32_t nb = (nv + kBatch - 1) / kBatch;
        st_blocks += nb;
        for (uint32_t b = 0; b < nb; ++b) {
            const uint8_t* blk = base + (size_t)b * ix.block_stride;
            uint32_t accu = accumulate_block<NCB, WIDE>(blk, T, ncb, lane);
            if (WIDE) accu &= 0xffffu;  // the reference accumulates in wrapping u16
            const float* fac = reinterpret_cast<const float*>(blk + (size_t)D * 4);
            const float f_add = __ldg(fac + lane), f_rescale = __ldg(fac + 32 + lane), f_error = __ldg(fac + 64 + lane);
            // K8 (AVX2 variant): ip = fmadd(delta, accu, sum_vl); est = (f_add+g_add) + f_rescale*(ip+k1x)
            const float ip = __fmaf_rn(s.delta, (float)accu, s.sum_vl);
            const float t1 = ip + s.k1x;
            const float t2 = f_rescale * t1;
            const float t3 = f_add + pr.g_add;
            const float est = t3 + t2;
            const float t4 = f_error * pr.g_error;
            float lower = est - t4;
            // K9
            const uint32_t li = b * kBatch + lane;
            bool valid = li < nv;
            unsigned long long vid = 0;
            if (a.filter != nullptr) {
                if (valid) {
                    vid = ix.ids[vbase + li];
                    const uint32_t id32 = (uint32_t)vid;
                    valid = (unsigned long long)id32 < a.filter_nbits && ((a.filter[id32 >> 6] >> (id32 & 63u)) & 1ull);
                }
            }
            if (!isfinite(lower)) lower = l2 ? 0.0f : -(pr.dot_qc + s.qnorm);
            const float theta0 = cnt >= k ? sd[k - 1] : INFINITY;
            const bool cand = valid && (lower < theta0);
            const unsigned mask = __ballot_sync(0xffffffffu, cand);
            st_cand += __popc(__ballot_sync(0xffffffffu, valid));
            if (mask == 0u) continue;
            if (cand && a.filter == nullptr) vid = ix.ids[vbase + li];
            float dist = est;
            if (EXK != 0) {
                // refine the superset: 4 candidates at a time, 8 lanes each
                float exdot = 0.0f;
                unsigned m = mask;
                const int g = lane >> 3, j = lane & 7;
                while (m) {
                    int src = -1;  // the candidate lane served by this 8-lane group in this round
                    unsigned mm = m;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (mm) {
                            const int sl = __ffs(mm) - 1;
                            mm &= mm - 1;
                            if (t == g) src = sl;
                        }
                    }
                    float part = 0.0f;
                    if (src >= 0) {
                        const unsigned long long gv = vbase + (unsigned long long)b * kBatch + src;
                        part = ex_dot_lane<EXK>(ix.ex + gv * ix.ex_stride, rq, D, j, ix.ex_bits);
                    }
                    part = hsum8(part);
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int sl = __shfl_sync(0xffffffffu, src, t * 8);
                        const float v = __shfl_sync(0xffffffffu, part, t * 8);
                        if (sl == lane) exdot = v;
                    }
                    st_ref += __popc(m) < 4 ? __popc(m) : 4;
                    m = mm;
                }
                if (cand) {
                    // distance = f_add_ex + g_add + f_rescale_ex * (binary_scale*ip + ex_dot + kbx)  (ivf.rs:2095-2099)
                    const unsigned long long gv = vbase + li;
                    float tt = s.bscale * ip;
                    tt = tt + exdot;
                    tt = tt + s.kbx;
                    const float mm2 = __ldg(ix.f_rescale_ex + gv) * tt;
                    const float aa = __ldg(ix.f_add_ex + gv) + pr.g_add;
                    dist = aa + mm2;
                }
            }
            // replay in reference order against the live threshold
            unsigned m = mask;
            while (m) {
                const int sl = __ffs(m) - 1;
                m &= m - 1;
                const float lb_s = __shfl_sync(0xffffffffu, lower, sl);
                const float d_s = __shfl_sync(0xffffffffu, dist, sl);
                const unsigned long long id_s = __shfl_sync(0xffffffffu, vid, sl);
                const float theta = cnt >= k ? sd[k - 1] : INFINITY;
                if (lb_s >= theta) continue;  // skipped_by_lower_bound
                st_adm += 1;
                if (!isfinite(d_s)) continue;
                topk_insert(sd, si, cnt, k, d_s, id_s, lane);
            }
        }
    }
    for (int i = lane; i < k; i += 32) {
        const bool have = i < cnt;
        a.out_ids[(size_t)q * k + i] = have ? si[i] : ~0ull;
        a.out_scores[(size_t)q * k + i] = have ? (l2 ? sd[i] : -sd[i]) : 0.0f;
    }
    if (lane == 0) {
        a.out_counts[q] = (uint32_t)cnt;
        if (a.stats) {
            atomicAdd(&a.stats->blocks, st_blocks);
            atomicAdd(&a.stats->candidates, st_cand);
            atomicAdd(&a.stats->refined, st_ref);
            atomicAdd(&a.stats->admitted, st_adm);
        }
    }
}

template <int NCB, bool WIDE>
static int launch_scan_ex(const DevIndex& ix, const ScanArgs& a, size_t smem, cudaStream_t st) {
    const unsigned grid = (a.nq + kWarps - 1) / kWarps;
#define RBQ_LAUNCH(EXK)                                                                                        \
    do {                                                                                                       \
        if (smem > 48 * 1024)                                                                                  \
            RBQ_CUDA(cudaFuncSetAttribute(scan_kernel<NCB, EXK, WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          (int)smem));                                                         \
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
                cudaStream_t st) {
    if (nq == 0) return RBQ_OK;
    if (top_k > (size_t)kMaxTopK) return fail(RBQ_INVALID_CONFIG, "top_k exceeds the device limit (1024)");
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
    const size_t smem = (size_t)kWarps * top_k * 12 + (size_t)kWarps * ix.D * 4;
    const int ncb_lane = (ix.D / 4 + 31) / 32;
    if (ix.D > 2048) return fail(RBQ_INVALID_CONFIG, "padded_dim > 2048 (high-accuracy LUT path) is not supported");
    if (ix.D > 1024) {
        if (ncb_lane <= 12) return launch_scan_ex<12, true>(ix, a, smem, st);
        return launch_scan_ex<16, true>(ix, a, smem, st);
    }
    switch (ncb_lane) {
        case 1: return launch_scan_ex<1, false>(ix, a, smem, st);
        case 2: return launch_scan_ex<2, false>(ix, a, smem, st);
        case 3: return launch_scan_ex<3, false>(ix, a, smem, st);
        case 4: return launch_scan_ex<4, false>(ix, a, smem, st);
        case 5:
        case 6: return launch_scan_ex<6, false>(ix, a, smem, st);
        default: return launch_scan_ex<8, false>(ix, a, smem, st);
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
        // generic (slow) formulation of the same mapping: 16 codebooks per pass so any D works
        uint32_t E[8], O[8];
#pragma unroll
        for (int m = 0; m < 8; ++m) E[m] = O[m] = 0u;
        uint32_t total = 0;
        for (int c0 = 0; c0 < ncb; c0 += 32) {
            const int cb = c0 + lane;
            uint4 C = make_uint4(0, 0, 0, 0), T = make_uint4(0, 0, 0, 0);
            if (cb < ncb) {
                C = ldg128(blk + 16 * cb);
                T = ldg128(lut + 16 * cb);
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) E[m] = O[m] = 0u;
            lookup_accumulate(C, T, E, O);
            total += reduce_scatter<true>(E, O, lane);
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
                             uint32_t* __restrict__ out_counts) {
    // Shards hold disjoint id sets, each list is sorted best-first; repeatedly take the best head.
    // Order = the reference's result order: L2 ascending distance, IP descending score; ties -> lower shard.
    const uint32_t q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint32_t head = 0, cnt = 0;  // lane s (< nshards) tracks shard s; nshards <= 32
    if (lane < nshards) cnt = in_counts[(size_t)lane * nq + q];
    uint32_t n = 0;
    for (; n < k; ++n) {
        float key = INFINITY;
        bool have = lane < nshards && head < cnt;
        if (have) {
            float sc = in_scores[((size_t)lane * nq + q) * k + head];
            key = metric == RBQ_METRIC_L2 ? sc : -sc;
        }
        // warp argmin over (have, key, lane)
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
            out_ids[(size_t)q * k + n] = in_ids[((size_t)lane * nq + q) * k + head];
            out_scores[(size_t)q * k + n] = in_scores[((size_t)lane * nq + q) * k + head];
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
                 cudaStream_t st) {
    if (nq == 0 || top_k == 0) return RBQ_OK;
    if (nshards < 1 || nshards > 32) return fail(RBQ_INVALID_CONFIG, "merge supports 1..32 shards");
    const unsigned grid = (unsigned)((nq + 3) / 4);
    merge_kernel<<<grid, 128, 0, st>>>(metric, nshards, (uint32_t)nq, (uint32_t)top_k,
                                       reinterpret_cast<const unsigned long long*>(in_ids), in_scores, in_counts,
                                       reinterpret_cast<unsigned long long*>(out_ids), out_scores, out_counts);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
