// coarse.cu -- probe selection on the device.
//   K4  centroid scoring   reference src/ivf.rs:1782-1789 -> math::l2_distance_sqr / dot
//                          (AVX2 variants, src/math.rs:154-181, 216-245)
//   K5  top-nprobe select  reference src/ivf.rs:1803-1835 (score.total_cmp, then cluster id)
//   K6  per-list constants reference src/ivf.rs:1850-1857 (g_add, g_error, dot_query_centroid)
//
// The probed-list set and its visit order must be bit-exact, so scores are produced in the
// reference's float order: 8 strided partial sums ("AVX lanes") updated with a separate multiply
// and add per element, lanes then summed 0..7 starting from 0.0f.  Compiled with -fmad=false.
//
// coarse_exact_kernel scores every (query, centroid) pair that way on the CUDA cores (mode 0).
// probe_select_kernel picks the nprobe best per query with an exact radix select on the
// total_cmp-ordered key (ties broken by cluster id like the reference), sorts them, and re-derives
// the K6 constants with the same lane order.
#include <algorithm>
#include <cmath>

#include "rbq_internal.h"

namespace rbq {

// ---- K4: exact all-pairs scoring ---------------------------------------------------------------
constexpr int TQ = 32, TC = 64, KC = 16, CPAD = 4;

__global__ void __launch_bounds__(256) coarse_exact_kernel(DevIndex ix, const float* __restrict__ rot, int nq,
                                                           float* __restrict__ scores) {
    __shared__ __align__(16) float qs[KC][TQ + CPAD];
    __shared__ __align__(16) float cs[KC][TC + CPAD];
    const int D = ix.D, nl = (int)ix.nlist;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.y * TQ, c0 = blockIdx.x * TC;
    const bool l2 = ix.metric == RBQ_METRIC_L2;
    float acc[2][4][8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
#pragma unroll
            for (int l = 0; l < 8; ++l) acc[a][b][l] = 0.0f;

    const int lrow = tid >> 2, lk = (tid & 3) * 4;  // loader mapping: one float4 of one row
    for (int k0 = 0; k0 < D; k0 += KC) {
        {
            float4 v = make_float4(0, 0, 0, 0);
            if (c0 + lrow < nl) v = *reinterpret_cast<const float4*>(ix.centroids + (size_t)(c0 + lrow) * D + k0 + lk);
            cs[lk][lrow] = v.x;
            cs[lk + 1][lrow] = v.y;
            cs[lk + 2][lrow] = v.z;
            cs[lk + 3][lrow] = v.w;
            if (tid < TQ * 4) {
                float4 u = make_float4(0, 0, 0, 0);
                if (q0 + lrow < nq) u = *reinterpret_cast<const float4*>(rot + (size_t)(q0 + lrow) * D + k0 + lk);
                qs[lk][lrow] = u.x;
                qs[lk + 1][lrow] = u.y;
                qs[lk + 2][lrow] = u.z;
                qs[lk + 3][lrow] = u.w;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const float4 cv = *reinterpret_cast<const float4*>(&cs[k][tx * 4]);
            const float2 qv = *reinterpret_cast<const float2*>(&qs[k][ty * 2]);
            const float cc[4] = {cv.x, cv.y, cv.z, cv.w};
            const float qq[2] = {qv.x, qv.y};
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    float p;
                    if (l2) {
                        float d = qq[a] - cc[b];
                        p = d * d;
                    } else {
                        p = qq[a] * cc[b];
                    }
                    acc[a][b][k & 7] = acc[a][b][k & 7] + p;
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            int q = q0 + ty * 2 + a, c = c0 + tx * 4 + b;
            if (q < nq && c < nl) {
                float s = 0.0f;
#pragma unroll
                for (int l = 0; l < 8; ++l) s = s + acc[a][b][l];
                scores[(size_t)q * nl + c] = s;
            }
        }
}

int launch_coarse_exact(const DevIndex& ix, const float* d_rot, size_t nq, float* d_scores, cudaStream_t st) {
    if (nq == 0) return RBQ_OK;
    dim3 grid((ix.nlist + TC - 1) / TC, (unsigned)((nq + TQ - 1) / TQ));
    coarse_exact_kernel<<<grid, 256, 0, st>>>(ix, d_rot, (int)nq, d_scores);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// ---- K5 + K6: exact top-nprobe and per-list constants -------------------------------------------
constexpr int kSelThreads = 256;
// launch_probe_select_cand, candidate capacity <= 512: CTA shape of the selection kernel.  Measured at GIST-1M / 10k queries / nprobe 16
// (select stage, ms): 0 = 256 threads x 4 CTAs/SM, 16 centroid elements in flight per lane: 0.169; 1 = 32 in flight: 0.186;
// 2 = 128 threads x 10 CTAs/SM: 0.164; 3 = 128 x 8, 32 in flight: 0.167; 4 = 256 x 5: 0.170; 5 = 128 threads x 12 CTAs/SM: 0.161.
// 6 = 5 with the lane-grouped centroid rows (exact_pair_q4): 0.109; 7 = 6 with 32 in flight: 0.119; 8 = 256 x 5 lane-grouped: 0.137.
constexpr int kSelVariantDefault = 6;
constexpr int kMaxNprobe = 4096;
size_t probe_select_max_nprobe() { return kMaxNprobe; }

__device__ __forceinline__ uint32_t order_key(float f, bool descending) {
    // monotone map of f32::total_cmp onto u32 (negative NaN < -inf < ... < +inf < positive NaN)
    uint32_t b = __float_as_uint(f);
    uint32_t u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return descending ? ~u : u;
}

__device__ __forceinline__ float key_to_float(uint32_t u, bool descending) {
    if (descending) u = ~u;
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

struct SelShared {
    unsigned int hist[256];
    unsigned int prefix, remaining, nless, eqbase, warp_tot[kSelThreads / 32];
    unsigned int count;
};

// Radix select (8 bits per pass, most significant first): key of the n-th smallest order_key among
// sc[0..nl).  Returns the key; *take_eq = how many entries equal to it belong to the n smallest.
__device__ __forceinline__ uint32_t radix_select(const float* __restrict__ sc, int nl, int n, bool desc, SelShared& sh,
                                                 unsigned int* take_eq) {
    const int tid = threadIdx.x;
    if (tid == 0) {
        sh.prefix = 0;
        sh.remaining = (unsigned)n;
    }
    __syncthreads();
    uint32_t mask = 0;
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = tid; i < 256; i += kSelThreads) sh.hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = sh.prefix;
        for (int c = tid; c < nl; c += kSelThreads) {
            const uint32_t u = order_key(sc[c], desc);
            if ((u & mask) == prefix) atomicAdd(&sh.hist[(u >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int cum = 0, rem = sh.remaining;
            int b = 0;
            for (; b < 256; ++b) {
                if (cum + sh.hist[b] >= rem) break;
                cum += sh.hist[b];
            }
            sh.remaining = rem - cum;
            sh.prefix = prefix | ((uint32_t)b << (8 * pass));
        }
        mask |= 255u << (8 * pass);
        __syncthreads();
    }
    *take_eq = sh.remaining;
    return sh.prefix;
}

// Exact top-nprobe keys from exact scores: keys < vstar in any order, keys == vstar in increasing
// cluster id (the reference's tie-break, src/ivf.rs:1808-1823).  Leaves nprobe keys in sel[0..nprobe).
__device__ __forceinline__ void gather_exact(const float* __restrict__ sc, int nl, int nprobe, bool desc, SelShared& sh,
                                             unsigned long long* sel) {
    const int tid = threadIdx.x;
    unsigned int take_eq;
    const uint32_t vstar = radix_select(sc, nl, nprobe, desc, sh, &take_eq);
    const unsigned int n_less = (unsigned)nprobe - take_eq;
    if (tid == 0) {
        sh.nless = 0;
        sh.eqbase = 0;
    }
    __syncthreads();
    for (int base = 0; base < nl; base += kSelThreads) {
        const int c = base + tid;
        uint32_t u = 0;
        bool less = false, eq = false;
        if (c < nl) {
            u = order_key(sc[c], desc);
            less = u < vstar;
            eq = u == vstar;
        }
        if (less) {
            const unsigned int slot = atomicAdd(&sh.nless, 1u);
            sel[slot] = ((unsigned long long)u << 32) | (unsigned)c;
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, eq);
        if ((tid & 31) == 0) sh.warp_tot[tid >> 5] = __popc(bal);
        __syncthreads();
        unsigned int before = sh.eqbase;
        for (int w = 0; w < (tid >> 5); ++w) before += sh.warp_tot[w];
        const unsigned int pos = before + __popc(bal & ((1u << (tid & 31)) - 1u));
        if (eq && pos < take_eq) sel[n_less + pos] = ((unsigned long long)u << 32) | (unsigned)c;
        __syncthreads();
        if (tid == 0) {
            unsigned int t = 0;
            for (int w = 0; w < kSelThreads / 32; ++w) t += sh.warp_tot[w];
            sh.eqbase += t;
        }
        __syncthreads();
    }
}

// l2_distance_sqr and dot of the query against one centroid, AVX2 lane order; all 8 lanes of the
// group return both values.
template <int N, bool NEED_L2, bool NEED_IP>
__device__ __forceinline__ void exact_batch(const float* __restrict__ rq, const float* __restrict__ ce, int i, float& al2, float& aip) {
    float b[N];  // N centroid elements in flight per lane (the row comes from L2): the adds stay in index order
#pragma unroll
    for (int u = 0; u < N; ++u) b[u] = __ldg(ce + i + 8 * u);
#pragma unroll
    for (int u = 0; u < N; ++u) {
        const float a = rq[i + 8 * u];
        if (NEED_L2) {
            const float d = a - b[u];
            const float p = d * d;
            al2 = al2 + p;
        }
        if (NEED_IP) {
            const float m = a * b[u];
            aip = aip + m;
        }
    }
}
template <bool NEED_L2 = true, bool NEED_IP = true, int DEPTH = 16>
__device__ __forceinline__ void exact_pair(const float* __restrict__ rq, const float* __restrict__ ce, int D, int lane8,
                                           unsigned gmask, float* l2_out, float* ip_out) {
    float al2 = 0.0f, aip = 0.0f;
    int i = lane8;
    for (; i + 8 * (DEPTH - 1) < D; i += 8 * DEPTH) exact_batch<DEPTH, NEED_L2, NEED_IP>(rq, ce, i, al2, aip);
    if (DEPTH > 8)
        for (; i + 8 * 7 < D; i += 8 * 8) exact_batch<8, NEED_L2, NEED_IP>(rq, ce, i, al2, aip);
    for (; i < D; i += 8) exact_batch<1, NEED_L2, NEED_IP>(rq, ce, i, al2, aip);
    float l2 = 0.0f, ip = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        if (NEED_L2) l2 = l2 + __shfl_sync(gmask, al2, l, 8);
        if (NEED_IP) ip = ip + __shfl_sync(gmask, aip, l, 8);
    }
    *l2_out = l2;
    *ip_out = ip;
}

// The same two sums from the lane-grouped layouts (DevIndex::cent_q4, and the query staged the same way): "AVX lane" j reads its
// elements j, j+8, j+16, ... four at a time as one float4, so a warp instruction of four 8-lane groups touches four 128-byte lines
// instead of four 32-byte pieces per ELEMENT (the re-score was bound by those L1 wavefronts), and the query costs one LDS.128 per
// four elements.  Same operations in the same order per lane: bit-identical to exact_pair.
__host__ __device__ __forceinline__ int q4_index(int i) {  // position of element i in the lane-grouped row
    const int j = i & 7, kk = i >> 3;
    return (((kk >> 2) * 8 + j) << 2) + (kk & 3);
}
template <bool NEED_L2 = true, bool NEED_IP = true, int DEPTH4 = 4>
__device__ __forceinline__ void exact_pair_q4(const float* __restrict__ rq4, const float* __restrict__ ce4, int D, int lane8,
                                              unsigned gmask, float* l2_out, float* ip_out) {
    float al2 = 0.0f, aip = 0.0f;
    const float4* c4 = reinterpret_cast<const float4*>(ce4) + lane8;
    const float4* r4 = reinterpret_cast<const float4*>(rq4) + lane8;
    const int n4 = D >> 5;  // float4 per lane
    for (int k = 0; k < n4; k += DEPTH4) {
        float4 b[DEPTH4];
#pragma unroll
        for (int u = 0; u < DEPTH4; ++u)
            if (k + u < n4) b[u] = __ldg(c4 + (k + u) * 8);
#pragma unroll
        for (int u = 0; u < DEPTH4; ++u) {
            if (k + u < n4) {
                const float4 a = r4[(k + u) * 8];
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (NEED_L2) {
                        const float d = av[e] - bv[e];
                        const float p = d * d;
                        al2 = al2 + p;
                    }
                    if (NEED_IP) {
                        const float m = av[e] * bv[e];
                        aip = aip + m;
                    }
                }
            }
        }
    }
    float l2 = 0.0f, ip = 0.0f;
#pragma unroll
    for (int l = 0; l < 8; ++l) {
        if (NEED_L2) l2 = l2 + __shfl_sync(gmask, al2, l, 8);
        if (NEED_IP) ip = ip + __shfl_sync(gmask, aip, l, 8);
    }
    *l2_out = l2;
    *ip_out = ip;
}
__global__ void __launch_bounds__(256) centroid_q4_kernel(const float* __restrict__ in, size_t total, int D, float* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const size_t row = idx / (size_t)D;
    const int i = (int)(idx - row * (size_t)D);
    out[row * (size_t)D + q4_index(i)] = in[idx];
}
int launch_centroid_q4(const float* d_in, size_t rows, int D, float* d_out, cudaStream_t st) {
    const size_t total = rows * (size_t)D;
    if (total == 0) return RBQ_OK;
    if (D % 32 != 0) return fail(RBQ_INVALID_CONFIG, "lane-grouped centroids need padded_dim % 32 == 0");
    centroid_q4_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(d_in, total, D, d_out);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// bitonic sort of sel[0..sort_n) ascending, then K6 for the first nprobe entries
__device__ __forceinline__ void sort_and_emit(const DevIndex& ix, const float* __restrict__ rq, unsigned long long* sel,
                                              int sort_n, int nprobe, bool desc, Probe* __restrict__ out) {
    const int tid = threadIdx.x, D = ix.D;
    for (int k = 2; k <= sort_n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < sort_n; i += kSelThreads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = sel[i], b = sel[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        sel[i] = b;
                        sel[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    const int lane8 = tid & 7, grp = tid >> 3;
    const unsigned int gmask = 0xffu << ((tid & 31) & ~7);
    for (int r = grp; r < nprobe; r += kSelThreads / 8) {
        const uint32_t cid = (uint32_t)(sel[r] & 0xffffffffull);
        float l2, ip;
        exact_pair(rq, ix.centroids + (size_t)cid * D, D, lane8, gmask, &l2, &ip);
        if (lane8 == 0) {
            Probe pr;
            pr.cid = cid;
            pr.g_add = desc ? -ip : l2;
            pr.g_error = sqrtf(l2);
            pr.dot_qc = ip;
            pr.nv = ix.list_n[cid];
            pr.blk_off = ix.blk_off[cid];
            pr.vec_off = ix.vec_off[cid];
            out[r] = pr;
        }
    }
}

// mode 0: scores are exact (coarse_exact_kernel)
__global__ void __launch_bounds__(kSelThreads) probe_select_kernel(DevIndex ix, const float* __restrict__ rot,
                                                                  const float* __restrict__ scores, int nprobe,
                                                                  int sort_n, Probe* __restrict__ probes) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(sel_smem);  // sort_n keys
    float* rq = reinterpret_cast<float*>(sel + sort_n);                         // D floats
    __shared__ SelShared sh;
    const int tid = threadIdx.x, nl = (int)ix.nlist, D = ix.D;
    const size_t q = blockIdx.x;
    const bool desc = ix.metric == RBQ_METRIC_INNER_PRODUCT;
    for (int i = tid; i < D; i += kSelThreads) rq[i] = rot[q * D + i];
    for (int i = tid; i < sort_n; i += kSelThreads) sel[i] = ~0ull;
    __syncthreads();
    gather_exact(scores + q * (size_t)nl, nl, nprobe, desc, sh, sel);
    sort_and_emit(ix, rq, sel, sort_n, nprobe, desc, probes + q * (size_t)nprobe);
}

// mode 1: scores are the tensor-core approximations.  Every centroid whose approximate score is within
// 2*delta of the nprobe-th best is re-scored exactly (delta bounds |approx - reference score|, see
// DESIGN.md), which provably contains the reference's top-nprobe; the exact keys are then sorted with the
// reference's comparator.  If the candidate set overflows the sort buffer the query falls back to exact
// scoring of every centroid inside this kernel.
__global__ void __launch_bounds__(kSelThreads) probe_select_tc_kernel(DevIndex ix, const float* __restrict__ rot,
                                                                     float* __restrict__ scores,
                                                                     const QueryScalars* __restrict__ qs, int nprobe,
                                                                     int sort_n, float eps_g, Probe* __restrict__ probes,
                                                                     unsigned int* __restrict__ fallbacks) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(sel_smem);  // sort_n keys
    float* rq = reinterpret_cast<float*>(sel + sort_n);                         // D floats
    uint32_t* cand = reinterpret_cast<uint32_t*>(rq + ix.D);                    // sort_n candidate ids
    __shared__ SelShared sh;
    const int tid = threadIdx.x, nl = (int)ix.nlist, D = ix.D;
    const size_t q = blockIdx.x;
    const bool desc = ix.metric == RBQ_METRIC_INNER_PRODUCT;
    float* sc = scores + q * (size_t)nl;
    for (int i = tid; i < D; i += kSelThreads) rq[i] = rot[q * D + i];
    for (int i = tid; i < sort_n; i += kSelThreads) sel[i] = ~0ull;
    __syncthreads();

    unsigned int take_eq;
    const uint32_t tkey = radix_select(sc, nl, nprobe, desc, sh, &take_eq);
    const float T = key_to_float(tkey, desc);
    const float qn = qs[q].qnorm, cm = ix.cmax_norm;
    const float round_terms = (1.2f * (float)D + 20.0f) * 5.9604645e-8f;
    float thr;
    if (!desc) {
        const float s2 = (qn + cm) * (qn + cm);
        thr = T + 2.0f * (0.5f * eps_g + round_terms) * s2;
    } else {
        thr = T - 2.0f * (eps_g + round_terms) * qn * cm;
    }
    const uint32_t thr_key = order_key(thr, desc);
    if (tid == 0) sh.count = 0;
    __syncthreads();
    for (int c = tid; c < nl; c += kSelThreads) {
        if (order_key(sc[c], desc) <= thr_key) {
            const unsigned int slot = atomicAdd(&sh.count, 1u);
            if (slot < (unsigned)sort_n) cand[slot] = (uint32_t)c;
        }
    }
    __syncthreads();
    const unsigned int m = sh.count;
    const int lane8 = tid & 7, grp = tid >> 3;
    const unsigned int gmask = 0xffu << ((tid & 31) & ~7);
    if (m <= (unsigned)sort_n && m >= (unsigned)nprobe) {
        for (unsigned int i = grp; i < m; i += kSelThreads / 8) {
            const uint32_t cid = cand[i];
            float l2, ip;
            exact_pair(rq, ix.centroids + (size_t)cid * D, D, lane8, gmask, &l2, &ip);
            if (lane8 == 0) sel[i] = ((unsigned long long)order_key(desc ? ip : l2, desc) << 32) | cid;
        }
        __syncthreads();
    } else {
        // rare: exact scores for every centroid of this query, then the exact selection
        if (tid == 0) atomicAdd(fallbacks, 1u);
        for (int c = grp; c < nl; c += kSelThreads / 8) {
            float l2, ip;
            exact_pair(rq, ix.centroids + (size_t)c * D, D, lane8, gmask, &l2, &ip);
            if (lane8 == 0) sc[c] = desc ? ip : l2;
        }
        __syncthreads();
        gather_exact(sc, nl, nprobe, desc, sh, sel);
    }
    sort_and_emit(ix, rq, sel, sort_n, nprobe, desc, probes + q * (size_t)nprobe);
}

// mode 1, fast path (nlist <= 256*KPT, small candidate sets): the query's approximate keys live in registers (KPT per
// thread), the radix select skips the bit prefix all keys share (scores of one query span a narrow range, so a
// fixed top-down digit order would pile every key on one histogram bin), the bin scan is a warp prefix sum, the
// candidates' exact (l2, ip) are computed once and reused for K6, and the exact keys are ranked by counting.
// LIST: the keys come from the query's candidate list (coarse filter mode: the GEMM epilogue kept only the centroids whose
// approximate score beat a per-query threshold `fthr` estimated from a centroid sample) instead of a dense score row.  The
// list provably holds every centroid with approximate score <= fthr, so the selection below is exact whenever the 2*delta
// band around the nprobe-th best candidate stays inside fthr; otherwise (or when the list overflowed / is too short) the
// query goes to the exact fallback kernel.
struct SelList {
    const CandRec* cand;
    const uint32_t* cand_cnt;
    const float* fthr;
    uint32_t cap;
    uint32_t* fb_list;
    uint32_t* fb_count;
};
// NT threads per CTA, MINB resident CTAs per SM asked of the compiler, DEPTH centroid elements in flight per lane of the exact re-score
// Q4: the query is staged, and the centroid rows are read, in the lane-grouped layout (exact_pair_q4)
template <int KPT, bool NEED_IP, bool LIST, int NT = kSelThreads, int MINB = 4, int DEPTH = 16, bool Q4 = false>
__global__ void __launch_bounds__(NT, MINB) probe_select_fast_kernel(DevIndex ix, const float* __restrict__ rot,
                                                                       float* __restrict__ scores,
                                                                       const QueryScalars* __restrict__ qs, int nprobe,
                                                                       int sort_n, float eps_g, Probe* __restrict__ probes,
                                                                       unsigned int* __restrict__ fallbacks, SelList sl) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(sel_smem);  // sort_n keys
    float* rq = reinterpret_cast<float*>(sel + sort_n);                         // D floats
    uint32_t* cand = reinterpret_cast<uint32_t*>(rq + ix.D);                    // sort_n candidate ids
    float* cl2 = reinterpret_cast<float*>(cand + sort_n);                       // sort_n exact l2
    float* cip = cl2 + sort_n;                                                  // sort_n exact ip
    __shared__ SelShared sh;
    static_assert(LIST || NT == kSelThreads, "the dense variant's rare path uses the kSelThreads-strided helpers");
    static_assert(LIST || !Q4, "the dense variant's rare path reads the query in its natural order");
    __shared__ uint32_t s_min[NT / 32], s_max[NT / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, D = ix.D;
    const size_t q = blockIdx.x;
    const bool desc = ix.metric == RBQ_METRIC_INNER_PRODUCT;
    int nl = (int)ix.nlist;
    float* sc = nullptr;
    const CandRec* cl = nullptr;
    if (LIST) {
        const uint32_t cnt = sl.cand_cnt[q];
        if (cnt > sl.cap || cnt < (uint32_t)nprobe) {  // overflowed or too short: exact fallback (uniform branch)
            if (tid == 0) sl.fb_list[atomicAdd(sl.fb_count, 1u)] = (uint32_t)q;
            return;
        }
        nl = (int)cnt;
        cl = sl.cand + q * (size_t)sl.cap;
    } else {
        sc = scores + q * (size_t)nl;
    }
    // scalars of the band test, loaded up front (their latency would otherwise sit between the select and the gather)
    const float qn = qs[q].qnorm, cm = ix.cmax_norm;
    const float fthr = LIST ? sl.fthr[q] : 0.0f;
    for (int i = tid; i < D; i += NT) rq[Q4 ? q4_index(i) : i] = rot[q * D + i];

    uint32_t key[KPT], col[LIST ? KPT : 1];
    uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        const int c = tid + NT * j;
        key[j] = 0u;
        if (LIST) col[j] = 0u;
        if (c < nl) {
            if (LIST) {
                const CandRec rec = cl[c];
                key[j] = order_key(rec.score, desc);
                col[j] = rec.cid;
            } else {
                key[j] = order_key(sc[c], desc);
            }
            kmin = min(kmin, key[j]);
            kmax = max(kmax, key[j]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kmin = min(kmin, __shfl_xor_sync(0xffffffffu, kmin, o));
        kmax = max(kmax, __shfl_xor_sync(0xffffffffu, kmax, o));
    }
    if (lane == 0) {
        s_min[warp] = kmin;
        s_max[warp] = kmax;
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        kmin = min(kmin, s_min[w]);
        kmax = max(kmax, s_max[w]);
    }
    // radix select of the nprobe-th smallest key below the shared prefix
    int hi = 32 - __clz(kmin ^ kmax);  // bits [hi, 32) are common to all keys (clz(0) == 32 -> hi = 0)
    uint32_t prefix = hi >= 32 ? 0u : (kmin >> hi) << hi;
    uint32_t remaining = (uint32_t)nprobe;
    while (hi > 0) {
        const int lo = hi > 8 ? hi - 8 : 0;
        const uint32_t dmask = (1u << (hi - lo)) - 1u;
        for (int i = tid; i < 256; i += NT) sh.hist[i] = 0;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < KPT; ++j) {
            const int c = tid + NT * j;
            if (c < nl && (((unsigned long long)(key[j] ^ prefix)) >> hi) == 0ull) atomicAdd(&sh.hist[(key[j] >> lo) & dmask], 1u);
        }
        __syncthreads();
        if (warp == 0) {  // bins 8*lane .. 8*lane+7
            uint32_t hcnt[8], tot = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                hcnt[b] = sh.hist[8 * lane + b];
                tot += hcnt[b];
            }
            uint32_t inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            uint32_t before = inc - tot;
            if (before < remaining && remaining <= inc) {  // exactly one lane
                int b = 0;
                for (; b < 7; ++b) {
                    if (before + hcnt[b] >= remaining) break;
                    before += hcnt[b];
                }
                sh.prefix = prefix | ((uint32_t)(8 * lane + b) << lo);
                sh.remaining = remaining - before;
                sh.nless = hcnt[b];  // keys left in the chosen bin
                sh.count = 0;
            }
        }
        __syncthreads();
        prefix = sh.prefix;
        remaining = sh.remaining;
        hi = lo;
        if (hi > 0 && sh.nless <= 64u) {
            // few keys left: finish by ranking them directly (ties ordered by list position; only the value matters)
            const unsigned nleft = sh.nless;
            __syncthreads();  // everyone has read the scan results; the histogram becomes the key list
#pragma unroll
            for (int j = 0; j < KPT; ++j) {
                const int c = tid + NT * j;
                if (c < nl && (((unsigned long long)(key[j] ^ prefix)) >> hi) == 0ull) sh.hist[atomicAdd(&sh.count, 1u)] = key[j];
            }
            __syncthreads();
            if ((unsigned)tid < nleft) {
                const uint32_t mine = sh.hist[tid];
                unsigned rank = 0;
                for (unsigned j = 0; j < nleft; ++j) {
                    const uint32_t o = sh.hist[j];
                    rank += (o < mine) || (o == mine && j < (unsigned)tid);
                }
                if (rank + 1 == remaining) sh.prefix = mine;
            }
            __syncthreads();
            prefix = sh.prefix;
            hi = 0;
        }
    }
    const float T = key_to_float(prefix, desc);
    const float round_terms = (1.2f * (float)D + 20.0f) * 5.9604645e-8f;
    float thr;
    if (!desc) {
        const float s2 = (qn + cm) * (qn + cm);
        thr = T + 2.0f * (0.5f * eps_g + round_terms) * s2;
    } else {
        thr = T - 2.0f * (eps_g + round_terms) * qn * cm;
    }
    const uint32_t thr_key = order_key(thr, desc);
    // LIST: the candidate list is complete only up to the filter threshold
    const bool covered = !LIST || thr_key <= order_key(fthr, desc);
    if (tid == 0) sh.count = 0;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < KPT; ++j) {
        const int c = tid + NT * j;
        if (c < nl && key[j] <= thr_key) {
            const unsigned int slot = atomicAdd(&sh.count, 1u);
            if (slot < (unsigned)sort_n) cand[slot] = LIST ? col[j] : (uint32_t)c;
        }
    }
    __syncthreads();
    const unsigned int m = sh.count;
    const int lane8 = tid & 7, grp = tid >> 3;
    const unsigned int gmask = 0xffu << ((tid & 31) & ~7);
    if (covered && m <= (unsigned)sort_n && m >= (unsigned)nprobe) {
        for (unsigned int i = grp; i < m; i += NT / 8) {
            const uint32_t cid = cand[i];
            float l2, ip;  // L2 searches never read dot_query_centroid (src/ivf.rs:2031-2042 uses it for InnerProduct only)
            if (Q4) exact_pair_q4<true, NEED_IP, (DEPTH + 3) / 4>(rq, ix.cent_q4 + (size_t)cid * D, D, lane8, gmask, &l2, &ip);
            else exact_pair<true, NEED_IP, DEPTH>(rq, ix.centroids + (size_t)cid * D, D, lane8, gmask, &l2, &ip);
            if (lane8 == 0) {
                sel[i] = ((unsigned long long)order_key(desc ? ip : l2, desc) << 32) | cid;
                cl2[i] = l2;
                cip[i] = ip;
            }
        }
        __syncthreads();
        // exact keys are unique (cluster id in the low word): rank by counting, the nprobe smallest emit K6
        for (unsigned int i = tid; i < m; i += NT) {
            const unsigned long long mine = sel[i];
            unsigned int rank = 0;
            for (unsigned int j = 0; j < m; ++j) rank += sel[j] < mine;
            if (rank < (unsigned)nprobe) {
                const uint32_t cid = (uint32_t)(mine & 0xffffffffull);
                const float l2 = cl2[i], ip = cip[i];
                Probe pr;
                pr.cid = cid;
                pr.g_add = desc ? -ip : l2;
                pr.g_error = sqrtf(l2);
                pr.dot_qc = ip;
                pr.nv = ix.list_n[cid];
                pr.blk_off = ix.blk_off[cid];
                pr.vec_off = ix.vec_off[cid];
                probes[q * (size_t)nprobe + rank] = pr;
            }
        }
        return;
    }
    if (LIST) {  // rare: the exact fallback kernel scores every centroid of this query
        if (tid == 0) sl.fb_list[atomicAdd(sl.fb_count, 1u)] = (uint32_t)q;
        return;
    }
    // rare: exact scores for every centroid of this query, then the exact selection
    if (tid == 0) atomicAdd(fallbacks, 1u);
    for (int i = tid; i < sort_n; i += NT) sel[i] = ~0ull;
    for (int c = grp; c < nl; c += NT / 8) {
        float l2, ip;
        exact_pair(rq, ix.centroids + (size_t)c * D, D, lane8, gmask, &l2, &ip);
        if (lane8 == 0) sc[c] = desc ? ip : l2;
    }
    __syncthreads();
    gather_exact(sc, nl, nprobe, desc, sh, sel);
    sort_and_emit(ix, rq, sel, sort_n, nprobe, desc, probes + q * (size_t)nprobe);
}

// Exact fallback, cooperative form: a fallback query scored by ONE CTA costs ~1 us per 5 centroids x 1000 dims (2.5 ms at 16384
// lists x 768 dims -- one such query stalled the whole batch).  Rounds of up to kFbRows queries: every (query, centroid slice)
// gets its own CTA for the scoring, then one CTA per query selects.  fb_count[0] = number of fallback queries.
constexpr int kFbRows = 64, kFbSlices = 64, kFbRounds = 2;
__global__ void __launch_bounds__(kSelThreads) fallback_score_kernel(DevIndex ix, const float* __restrict__ rot, const uint32_t* __restrict__ fb_list,
                                                                    const uint32_t* __restrict__ fb_count, uint32_t first, float* __restrict__ scratch) {
    const uint32_t idx = first + blockIdx.y;
    if (idx >= fb_count[0]) return;
    extern __shared__ __align__(16) unsigned char sel_smem[];
    float* rq = reinterpret_cast<float*>(sel_smem);
    const int tid = threadIdx.x, nl = (int)ix.nlist, D = ix.D;
    const bool desc = ix.metric == RBQ_METRIC_INNER_PRODUCT;
    const size_t q = fb_list[idx];
    for (int i = tid; i < D; i += kSelThreads) rq[i] = rot[q * D + i];
    __syncthreads();
    const int per = (nl + kFbSlices - 1) / kFbSlices, c0 = (int)blockIdx.x * per, c1 = min(nl, c0 + per);
    float* sc = scratch + (size_t)blockIdx.y * nl;
    const int lane8 = tid & 7, grp = tid >> 3;
    const unsigned int gmask = 0xffu << ((tid & 31) & ~7);
    for (int c = c0 + grp; c < c1; c += kSelThreads / 8) {
        float l2, ip;
        exact_pair(rq, ix.centroids + (size_t)c * D, D, lane8, gmask, &l2, &ip);
        if (lane8 == 0) sc[c] = desc ? ip : l2;
    }
}
__global__ void __launch_bounds__(kSelThreads) fallback_select_kernel(DevIndex ix, const float* __restrict__ rot, int nprobe, int sort_n,
                                                                     Probe* __restrict__ probes, const uint32_t* __restrict__ fb_list,
                                                                     const uint32_t* __restrict__ fb_count, uint32_t first, const float* __restrict__ scratch,
                                                                     unsigned int* __restrict__ fallbacks) {
    const uint32_t idx = first + blockIdx.x;
    if (idx >= fb_count[0]) return;
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(sel_smem);  // sort_n keys
    float* rq = reinterpret_cast<float*>(sel + sort_n);                         // D floats
    __shared__ SelShared sh;
    const int tid = threadIdx.x, nl = (int)ix.nlist, D = ix.D;
    const bool desc = ix.metric == RBQ_METRIC_INNER_PRODUCT;
    const size_t q = fb_list[idx];
    if (tid == 0) atomicAdd(fallbacks, 1u);
    for (int i = tid; i < D; i += kSelThreads) rq[i] = rot[q * D + i];
    for (int i = tid; i < sort_n; i += kSelThreads) sel[i] = ~0ull;
    __syncthreads();
    gather_exact(scratch + (size_t)blockIdx.x * nl, nl, nprobe, desc, sh, sel);
    sort_and_emit(ix, rq, sel, sort_n, nprobe, desc, probes + q * (size_t)nprobe);
}

// Exact fallback of the filter mode: the queries listed in fb_list get exact scores for every centroid (into a per-CTA scratch
// row) and the exact selection, like the dense kernels' rare path.  Persistent: fb_count[0] = how many, fb_count[1] = cursor.
__global__ void __launch_bounds__(kSelThreads) probe_select_exact_kernel(DevIndex ix, const float* __restrict__ rot, int nprobe, int sort_n,
                                                                        Probe* __restrict__ probes, const uint32_t* __restrict__ fb_list,
                                                                        uint32_t* __restrict__ fb_count, float* __restrict__ scratch,
                                                                        unsigned int* __restrict__ fallbacks, uint32_t first) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(sel_smem);  // sort_n keys
    float* rq = reinterpret_cast<float*>(sel + sort_n);                         // D floats
    __shared__ SelShared sh;
    __shared__ uint32_t s_idx;
    if (fb_count[0] <= first) return;  // the cooperative rounds took them all
    const int tid = threadIdx.x, nl = (int)ix.nlist, D = ix.D;
    const bool desc = ix.metric == RBQ_METRIC_INNER_PRODUCT;
    float* sc = scratch + (size_t)blockIdx.x * nl;
    const int lane8 = tid & 7, grp = tid >> 3;
    const unsigned int gmask = 0xffu << ((tid & 31) & ~7);
    const uint32_t total = fb_count[0];
    for (;;) {
        __syncthreads();
        if (tid == 0) s_idx = first + atomicAdd(&fb_count[1], 1u);
        __syncthreads();
        const uint32_t idx = s_idx;
        if (idx >= total) break;
        const size_t q = fb_list[idx];
        if (tid == 0) atomicAdd(fallbacks, 1u);
        for (int i = tid; i < D; i += kSelThreads) rq[i] = rot[q * D + i];
        for (int i = tid; i < sort_n; i += kSelThreads) sel[i] = ~0ull;
        __syncthreads();
        for (int c = grp; c < nl; c += kSelThreads / 8) {
            float l2, ip;
            exact_pair(rq, ix.centroids + (size_t)c * D, D, lane8, gmask, &l2, &ip);
            if (lane8 == 0) sc[c] = desc ? ip : l2;
        }
        __syncthreads();
        gather_exact(sc, nl, nprobe, desc, sh, sel);
        sort_and_emit(ix, rq, sel, sort_n, nprobe, desc, probes + q * (size_t)nprobe);
    }
}

// Filter mode, step 2: the rank-th best of a query's sample scores = the filter threshold.  One warp per query, the sample
// (<= 1024 scores) in registers, bisection on the order key between the sample's min and max (no block barrier).
template <int VPT>
__global__ void __launch_bounds__(128) sample_threshold_kernel(const float* __restrict__ ss, uint32_t nq, uint32_t samp_n, uint32_t rank, int desc,
                                                              float* __restrict__ thr) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const float* row = ss + (size_t)q * samp_n;
    uint32_t key[VPT];
    uint32_t kmin = 0xffffffffu, kmax = 0u;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const uint32_t c = (uint32_t)lane + 32u * j;
        key[j] = 0xffffffffu;
        if (c < samp_n) {
            key[j] = order_key(row[c], desc != 0);
            kmin = min(kmin, key[j]);
            kmax = max(kmax, key[j]);
        }
    }
    uint32_t lo = __reduce_min_sync(0xffffffffu, kmin), hi = __reduce_max_sync(0xffffffffu, kmax);
    while (lo < hi) {  // smallest K with #(key <= K) >= rank (rank <= samp_n, so K <= max key)
        const uint32_t mid = lo + ((hi - lo) >> 1);
        uint32_t c = 0;
#pragma unroll
        for (int j = 0; j < VPT; ++j) c += key[j] <= mid;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= rank) hi = mid;
        else lo = mid + 1u;
    }
    if (lane == 0) thr[q] = key_to_float(hi, desc != 0);
}
int launch_sample_threshold(const float* d_samp_scores, size_t nq, uint32_t samp_n, uint32_t rank, int metric, float* d_thr, cudaStream_t st) {
    if (nq == 0) return RBQ_OK;
    if (samp_n > 1024u || rank == 0 || rank > samp_n) return fail(RBQ_INVALID_CONFIG, "sample threshold: bad sample size or rank");
    const int desc = metric == RBQ_METRIC_INNER_PRODUCT;
    const unsigned grid = (unsigned)((nq + 3) / 4);
    if (samp_n <= 512u) sample_threshold_kernel<16><<<grid, 128, 0, st>>>(d_samp_scores, (uint32_t)nq, samp_n, rank, desc, d_thr);
    else sample_threshold_kernel<32><<<grid, 128, 0, st>>>(d_samp_scores, (uint32_t)nq, samp_n, rank, desc, d_thr);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// Sample rank whose score, used as the filter threshold, leaves about rank * nlist / samp_n candidates per query and
// falls below the needed population rank (nprobe plus the re-score band) with probability < ~1e-3 (3 sigma of the
// order statistic).  0: the filter mode does not pay (nprobe too large for this centroid table).
uint32_t filter_sample_rank(uint32_t nlist, uint32_t samp_n, size_t nprobe, int terms) {
    if (samp_n == 0 || nlist == 0) return 0;
    const double f = (double)samp_n / (double)nlist;
    const double need = ((terms == 1 ? 2.5 : 1.25) * (double)nprobe + 8.0) * f;
    uint32_t r = 4;
    while ((double)r - 3.0 * std::sqrt((double)r) < need) ++r;
    return r <= samp_n / 4 ? r : 0;
}
uint32_t filter_cand_cap(uint32_t nlist, uint32_t samp_n, size_t nprobe, int terms) {
    const uint32_t r = filter_sample_rank(nlist, samp_n, nprobe, terms);
    if (r == 0) return 0;
    const double expect = (double)r * (double)nlist / (double)samp_n;
    uint32_t cap = 512;
    while ((double)cap < 2.0 * expect + 64.0) cap <<= 1;
    return cap <= 4096 ? cap : 0;
}

static int sort_size
(size_t n) {
    int s = 32;
    while ((size_t)s < n) s <<= 1;
    return s;
}

int launch_probe_select(const DevIndex& ix, const float* d_rot, const float* d_scores, size_t nq, size_t nprobe,
                        Probe* d_probes, cudaStream_t st) {
    if (nq == 0) return RBQ_OK;
    if (nprobe > (size_t)kMaxNprobe)
        return fail(RBQ_INVALID_CONFIG, "nprobe exceeds the device probe-selection limit (4096)");
    const int sort_n = sort_size(nprobe);
    const size_t smem = (size_t)sort_n * 8 + (size_t)ix.D * 4;
    if (smem > 48 * 1024)
        RBQ_CUDA(cudaFuncSetAttribute(probe_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_select_kernel<<<(unsigned)nq, kSelThreads, smem, st>>>(ix, d_rot, d_scores, (int)nprobe, sort_n, d_probes);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

int launch_probe_select_tc(const DevIndex& ix, const float* d_rot, float* d_scores, const QueryScalars* d_qs, size_t nq,
                           size_t nprobe, float eps_g, Probe* d_probes, unsigned int* d_fallbacks, cudaStream_t st, bool need_ip) {
    if (nq == 0) return RBQ_OK;
    if (nprobe > (size_t)kMaxNprobe)
        return fail(RBQ_INVALID_CONFIG, "nprobe exceeds the device probe-selection limit (4096)");
    // room for the re-score band around the nprobe-th score: a handful of centroids with fp32-class scores, up to ~nprobe more
    // with bf16-class scores (1-term GEMM)
    const int sort_n = sort_size(std::min<size_t>(eps_g > 1e-3f ? 2 * nprobe + 64 : nprobe + 48, (size_t)kMaxNprobe));
    if (ix.nlist <= 16u * kSelThreads && sort_n <= 256) {
        const size_t smem_f = (size_t)sort_n * 20 + (size_t)ix.D * 4;
        const bool ipn = need_ip || ix.metric == RBQ_METRIC_INNER_PRODUCT;
#define RBQ_SEL(KPT, IP)                                                                                                     \
    probe_select_fast_kernel<KPT, IP, false><<<(unsigned)nq, kSelThreads, smem_f, st>>>(ix, d_rot, d_scores, d_qs, (int)nprobe, sort_n, \
                                                                                       eps_g, d_probes, d_fallbacks, SelList{})
        if (ix.nlist <= 4u * kSelThreads) {
            if (ipn) RBQ_SEL(4, true);
            else RBQ_SEL(4, false);
        } else {
            if (ipn) RBQ_SEL(16, true);
            else RBQ_SEL(16, false);
        }
#undef RBQ_SEL
        RBQ_CUDA(cudaGetLastError());
        return RBQ_OK;
    }
    const size_t smem = (size_t)sort_n * 12 + (size_t)ix.D * 4;
    if (smem > 48 * 1024)
        RBQ_CUDA(cudaFuncSetAttribute(probe_select_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_select_tc_kernel<<<(unsigned)nq, kSelThreads, smem, st>>>(ix, d_rot, d_scores, d_qs, (int)nprobe, sort_n, eps_g,
                                                                   d_probes, d_fallbacks);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// Filter mode, steps 4-5: selection from the candidate lists, then the exact fallback for the queries it could not settle.
int launch_probe_select_cand(const DevIndex& ix, const float* d_rot, const QueryScalars* d_qs, size_t nq, size_t nprobe, float eps_g,
                             const FilterWs& fw, Probe* d_probes, unsigned int* d_fallbacks, cudaStream_t st, bool need_ip) {
    if (nq == 0) return RBQ_OK;
    if (nprobe > (size_t)kMaxNprobe)
        return fail(RBQ_INVALID_CONFIG, "nprobe exceeds the device probe-selection limit (4096)");
    const int sort_n = sort_size(std::min<size_t>(2 * nprobe + 64, (size_t)kMaxNprobe));
    const size_t smem_f = (size_t)sort_n * 20 + (size_t)ix.D * 4;
    const bool ipn = need_ip || ix.metric == RBQ_METRIC_INNER_PRODUCT;
    SelList sl{fw.cand, fw.cand_cnt, fw.thr, fw.cap, fw.fb_list, fw.fb_count};
#define RBQ_SELL(KPT, IP)                                                                                                              \
    do {                                                                                                                              \
        if (smem_f > 48 * 1024)                                                                                                       \
            RBQ_CUDA(cudaFuncSetAttribute(probe_select_fast_kernel<KPT, IP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f)); \
        probe_select_fast_kernel<KPT, IP, true><<<(unsigned)nq, kSelThreads, smem_f, st>>>(ix, d_rot, nullptr, d_qs, (int)nprobe, sort_n, eps_g, \
                                                                                          d_probes, d_fallbacks, sl);                \
    } while (0)
    const uint32_t kpt = (fw.cap + kSelThreads - 1) / kSelThreads;
    // small candidate lists (cap <= 512): variants of the CTA shape / resident CTAs / re-score prefetch depth (RBQ_SEL_VARIANT, a
    // tuning knob read per launch; default = the measured best)
#define RBQ_SELV(KPT, IP, NTV, MINB, DEPTH, Q4V)                                                                                            \
    do {                                                                                                                              \
        if (smem_f > 48 * 1024)                                                                                                       \
            RBQ_CUDA(cudaFuncSetAttribute(probe_select_fast_kernel<KPT, IP, true, NTV, MINB, DEPTH, Q4V>,                                  \
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f));                                 \
        probe_select_fast_kernel<KPT, IP, true, NTV, MINB, DEPTH, Q4V><<<(unsigned)nq, NTV, smem_f, st>>>(ix, d_rot, nullptr, d_qs, (int)nprobe, \
                                                                                                    sort_n, eps_g, d_probes, d_fallbacks, sl); \
    } while (0)
    const char* ve = getenv("RBQ_SEL_VARIANT");
    int variant = ve ? atoi(ve) : kSelVariantDefault;
    if (variant >= 6 && ix.cent_q4 == nullptr) variant = 5;  // no lane-grouped centroids (padded_dim % 32 != 0)
    if (kpt <= 2 && variant > 0) {
        switch (variant) {
            case 6:
                if (ipn) RBQ_SELV(4, true, 128, 12, 16, true);
                else RBQ_SELV(4, false, 128, 12, 16, true);
                break;
            case 7:
                if (ipn) RBQ_SELV(4, true, 128, 10, 32, true);
                else RBQ_SELV(4, false, 128, 10, 32, true);
                break;
            case 8:
                if (ipn) RBQ_SELV(2, true, 256, 5, 16, true);
                else RBQ_SELV(2, false, 256, 5, 16, true);
                break;
            case 1:
                if (ipn) RBQ_SELV(2, true, 256, 4, 32, false);
                else RBQ_SELV(2, false, 256, 4, 32, false);
                break;
            case 2:
                if (ipn) RBQ_SELV(4, true, 128, 10, 16, false);
                else RBQ_SELV(4, false, 128, 10, 16, false);
                break;
            case 3:
                if (ipn) RBQ_SELV(4, true, 128, 8, 32, false);
                else RBQ_SELV(4, false, 128, 8, 32, false);
                break;
            case 4:
                if (ipn) RBQ_SELV(2, true, 256, 5, 16, false);
                else RBQ_SELV(2, false, 256, 5, 16, false);
                break;
            default:
                if (ipn) RBQ_SELV(4, true, 128, 12, 16, false);
                else RBQ_SELV(4, false, 128, 12, 16, false);
                break;
        }
    } else if (kpt <= 2) {
        if (ipn) RBQ_SELL(2, true);
        else RBQ_SELL(2, false);
    } else if (kpt <= 4) {
        if (ipn) RBQ_SELL(4, true);
        else RBQ_SELL(4, false);
    } else if (kpt <= 8) {
        if (ipn) RBQ_SELL(8, true);
        else RBQ_SELL(8, false);
    } else if (kpt <= 16) {
        if (ipn) RBQ_SELL(16, true);
        else RBQ_SELL(16, false);
    } else {
        return fail(RBQ_INVALID_CONFIG, "candidate capacity exceeds the selection kernel's limit (4096)");
    }
#undef RBQ_SELV
#undef RBQ_SELL
    RBQ_CUDA(cudaGetLastError());
    const int sort_x = sort_size(nprobe);
    const size_t smem_x = (size_t)sort_x * 8 + (size_t)ix.D * 4;
    if (smem_x > 48 * 1024) {
        RBQ_CUDA(cudaFuncSetAttribute(probe_select_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
        RBQ_CUDA(cudaFuncSetAttribute(fallback_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_x));
    }
    if (fw.fb_ctas < (uint32_t)kFbRows) return fail(RBQ_INVALID_CONFIG, "fallback scratch too small");
    for (int r = 0; r < kFbRounds; ++r) {  // cooperative rounds: kFbRows queries each, every (query, centroid slice) on its own CTA
        fallback_score_kernel<<<dim3(kFbSlices, kFbRows), kSelThreads, (size_t)ix.D * 4, st>>>(ix, d_rot, fw.fb_list, fw.fb_count, (uint32_t)(r * kFbRows),
                                                                                               fw.fb_scratch);
        fallback_select_kernel<<<kFbRows, kSelThreads, smem_x, st>>>(ix, d_rot, (int)nprobe, sort_x, d_probes, fw.fb_list, fw.fb_count,
                                                                     (uint32_t)(r * kFbRows), fw.fb_scratch, d_fallbacks);
    }
    // whatever is left (more than kFbRounds * kFbRows fallback queries in one chunk): one CTA per query
    probe_select_exact_kernel<<<fw.fb_ctas, kSelThreads, smem_x, st>>>(ix, d_rot, (int)nprobe, sort_x, d_probes, fw.fb_list, fw.fb_count,
                                                                       fw.fb_scratch, d_fallbacks, (uint32_t)(kFbRounds * kFbRows));
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
