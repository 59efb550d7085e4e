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
constexpr int kMaxNprobe = 4096;
size_t probe_select_max_nprobe() { return kMaxNprobe; }

__device__ __forceinline__ uint32_t order_key(float f, bool descending) {
    // monotone map of f32::total_cmp onto u32 (negative NaN < -inf < ... < +inf < positive NaN)
    uint32_t b = __float_as_uint(f);
    uint32_t u = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return descending ? ~u : u;
}

__global__ void __launch_bounds__(kSelThreads) probe_select_kernel(DevIndex ix, const float* __restrict__ rot,
                                                                  const float* __restrict__ scores, int nprobe,
                                                                  int sort_n, Probe* __restrict__ probes) {
    extern __shared__ __align__(16) unsigned char sel_smem[];
    unsigned long long* sel = reinterpret_cast<unsigned long long*>(sel_smem);  // sort_n keys
    float* rq = reinterpret_cast<float*>(sel + sort_n);                         // D floats
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_remaining, s_nless, s_eqbase, s_warp_tot[kSelThreads / 32];
    const int tid = threadIdx.x, nl = (int)ix.nlist, D = ix.D;
    const size_t q = blockIdx.x;
    const float* sc = scores + q * (size_t)nl;
    const bool desc = ix.metric == RBQ_METRIC_INNER_PRODUCT;

    for (int i = tid; i < D; i += kSelThreads) rq[i] = rot[q * D + i];
    for (int i = tid; i < sort_n; i += kSelThreads) sel[i] = ~0ull;
    if (tid == 0) {
        s_prefix = 0;
        s_remaining = (unsigned)nprobe;
        s_nless = 0;
        s_eqbase = 0;
    }
    __syncthreads();

    // radix select (8 bits per pass, most significant first) of the nprobe-th smallest key
    uint32_t mask = 0;
    for (int pass = 3; pass >= 0; --pass) {
        for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix;
        for (int c = tid; c < nl; c += kSelThreads) {
            uint32_t u = order_key(sc[c], desc);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            unsigned int cum = 0, rem = s_remaining;
            int b = 0;
            for (; b < 256; ++b) {
                if (cum + hist[b] >= rem) break;
                cum += hist[b];
            }
            s_remaining = rem - cum;
            s_prefix = prefix | ((uint32_t)b << (8 * pass));
        }
        mask |= 255u << (8 * pass);
        __syncthreads();
    }
    const uint32_t vstar = s_prefix;          // key of the nprobe-th best list
    const unsigned int take_eq = s_remaining;  // how many lists with key == vstar belong to the result
    const unsigned int n_less = (unsigned)nprobe - take_eq;

    // gather: keys < vstar in any order; keys == vstar in increasing cluster id (the reference's tie-break)
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
            unsigned int slot = atomicAdd(&s_nless, 1u);
            sel[slot] = ((unsigned long long)u << 32) | (unsigned)c;
        }
        const unsigned int bal = __ballot_sync(0xffffffffu, eq);
        if ((tid & 31) == 0) s_warp_tot[tid >> 5] = __popc(bal);
        __syncthreads();
        unsigned int before = s_eqbase;
        for (int w = 0; w < (tid >> 5); ++w) before += s_warp_tot[w];
        const unsigned int pos = before + __popc(bal & ((1u << (tid & 31)) - 1u));
        if (eq && pos < take_eq) sel[n_less + pos] = ((unsigned long long)u << 32) | (unsigned)c;
        __syncthreads();
        if (tid == 0) {
            unsigned int t = 0;
            for (int w = 0; w < kSelThreads / 32; ++w) t += s_warp_tot[w];
            s_eqbase += t;
        }
        __syncthreads();
    }

    // bitonic sort of the selected keys (ascending; padding = all ones)
    for (int k = 2; k <= sort_n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < sort_n; i += kSelThreads) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = sel[i], b = sel[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        sel[i] = b;
                        sel[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }

    // K6 for every selected list: 8 threads = the 8 AVX lanes of l2_distance_sqr / dot
    const int lane8 = tid & 7, grp = tid >> 3;
    const unsigned int gmask = 0xffu << ((tid & 31) & ~7);
    for (int r = grp; r < nprobe; r += kSelThreads / 8) {
        const uint32_t cid = (uint32_t)(sel[r] & 0xffffffffull);
        const float* ce = ix.centroids + (size_t)cid * D;
        float al2 = 0.0f, aip = 0.0f;
        for (int i = lane8; i < D; i += 8) {
            const float a = rq[i], b = ce[i];
            const float d = a - b;
            const float p = d * d;
            al2 = al2 + p;
            const float m = a * b;
            aip = aip + m;
        }
        float l2 = 0.0f, ip = 0.0f;
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            l2 = l2 + __shfl_sync(gmask, al2, l, 8);
            ip = ip + __shfl_sync(gmask, aip, l, 8);
        }
        if (lane8 == 0) {
            Probe pr;
            pr.cid = cid;
            pr.g_add = desc ? -ip : l2;
            pr.g_error = sqrtf(l2);
            pr.dot_qc = ip;
            pr.nv = ix.list_n[cid];
            pr.blk_off = ix.blk_off[cid];
            pr.vec_off = ix.vec_off[cid];
            probes[q * (size_t)nprobe + r] = pr;
        }
    }
}

int launch_probe_select(const DevIndex& ix, const float* d_rot, const float* d_scores, size_t nq, size_t nprobe,
                        Probe* d_probes, cudaStream_t st) {
    if (nq == 0) return RBQ_OK;
    if (nprobe > (size_t)kMaxNprobe)
        return fail(RBQ_INVALID_CONFIG, "nprobe exceeds the device probe-selection limit (4096)");
    int sort_n = 32;
    while ((size_t)sort_n < nprobe) sort_n <<= 1;
    size_t smem = (size_t)sort_n * 8 + (size_t)ix.D * 4;
    if (smem > 48 * 1024)
        RBQ_CUDA(cudaFuncSetAttribute(probe_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_select_kernel<<<(unsigned)nq, kSelThreads, smem, st>>>(ix, d_rot, d_scores, (int)nprobe, sort_n, d_probes);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace rbq
