// build.cu -- index construction on the GPU: IvfRabitqIndex::train_with_clusters
// (reference src/ivf.rs:1025-1215) = rotate data + centroids, quantize_with_centroid per vector
// (reference src/quantizer.rs:140-535), FastScan block packing (ClusterData::from_quantized_vectors,
// reference src/ivf.rs:409-696; simd::pack_codes src/simd.rs:864-904) straight into the device-resident
// index.  The result serialises to the same RBQ1 v3 bytes the reference's `save` would write for the
// same rotator state and rescale constant.
//
// Float order: math::dot is the AVX2 variant (8 strided lanes, separate mul+add, lanes summed 0..7);
// the |r| norm and the f64 ipnorm are sequential folds like the reference's iterators.  -fmad=false.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <queue>
#include <thread>

#include "rbq_internal.h"
#include "rotate.cuh"

namespace rbq {

// ---- host: rescale-factor search (reference src/quantizer.rs:337-427, 563-592) ---------------------
static const double kTightStart[9] = {0.0, 0.15, 0.20, 0.52, 0.59, 0.71, 0.75, 0.77, 0.81};
static const double kEps = 1e-5, kNenum = 10.0;
static const float kF32Eps = 1.1920929e-7f;

double best_rescale_factor_host(const float* o_abs, size_t dim, int ex_bits) {
    float mx = 0.0f;
    for (size_t i = 0; i < dim; ++i) mx = std::fmax(mx, o_abs[i]);
    const double max_o = mx;
    if (max_o <= 2.220446049250313e-16) return 1.0;
    const int top = (1 << ex_bits) - 1;
    const double t_end = ((double)top + kNenum) / max_o;
    const double t_start = t_end * kTightStart[std::min(ex_bits, 8)];
    std::vector<int32_t> cur(dim);
    double sqr_den = (double)dim * 0.25, num = 0.0;
    typedef std::pair<double, size_t> Ev;  // next breakpoint of a coordinate, ordered by (t, idx)
    std::priority_queue<Ev, std::vector<Ev>, std::greater<Ev>> events;
    for (size_t i = 0; i < dim; ++i) {
        const int32_t c = (int32_t)((t_start * (double)o_abs[i]) + kEps);
        cur[i] = c;
        sqr_den += (double)(c * c + c);
        num += ((double)c + 0.5) * (double)o_abs[i];
    }
    for (size_t i = 0; i < dim; ++i)
        if (o_abs[i] > 0.0f) events.push(Ev((double)(cur[i] + 1) / (double)o_abs[i], i));
    double best_ip = 0.0, best_t = t_start;
    while (!events.empty()) {
        const Ev e = events.top();
        events.pop();
        if (e.first >= t_end) continue;
        const size_t i = e.second;
        const int32_t u = ++cur[i];
        sqr_den += 2.0 * (double)u;
        num += (double)o_abs[i];
        const double ip = num / std::sqrt(sqr_den);
        if (ip > best_ip) {
            best_ip = ip;
            best_t = e.first;
        }
        if (u < top && o_abs[i] > 0.0f) {
            const double tn = (double)(u + 1) / (double)o_abs[i];
            if (tn < t_end) events.push(Ev(tn, i));
        }
    }
    return best_t <= 0.0 ? std::max(t_start, 2.220446049250313e-16) : best_t;
}

static inline uint64_t splitmix64(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
static inline double uniform01(uint64_t& s) { return ((splitmix64(s) >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static inline double gaussian(uint64_t& s) {
    const double u = uniform01(s), v = uniform01(s);
    return std::sqrt(-2.0 * std::log(u)) * std::cos(6.283185307179586 * v);
}

// Mean optimal rescale factor over 100 random unit vectors.  The reference draws them from
// StdRng(seed) (ChaCha12); any seeded Gaussian source gives a statistically equivalent constant.
float const_scaling_factor_host(size_t D, int ex_bits, uint64_t seed) {
    uint64_t st = seed;
    double sum_t = 0.0;
    std::vector<float> v(D), oa(D);
    for (int s = 0; s < 100; ++s) {
        float n2 = 0.0f;
        for (size_t i = 0; i < D; ++i) {
            v[i] = (float)gaussian(st);
            n2 += v[i] * v[i];
        }
        const float norm = std::sqrt(n2);
        if (norm <= kF32Eps) continue;
        for (size_t i = 0; i < D; ++i) oa[i] = std::fabs(v[i] / norm);
        sum_t += best_rescale_factor_host(oa.data(), D, ex_bits);
    }
    return (float)(sum_t / 100.0);
}

// ---- device kernels --------------------------------------------------------------------------------
// src (optional): input row of output row v (streaming build: the chunk's vectors in list order); rows with src == ~0 are skipped
__global__ void __launch_bounds__(256) rotate_only_kernel(DevIndex ix, const float* __restrict__ in, const uint32_t* __restrict__ src,
                                                          float* __restrict__ out) {
    extern __shared__ float rsm[];
    const int D = ix.D, tid = threadIdx.x, nt = blockDim.x;
    float* buf = rsm;
    float* tmp = rsm + D;
    const size_t v = blockIdx.x;
    size_t row = v;
    if (src != nullptr) {
        const uint32_t r = src[v];
        if (r == 0xffffffffu) return;
        row = r;
    }
    rotate_block(ix, in + row * ix.dim, buf, tmp, tid, nt);
    for (int i = tid; i < D; i += nt) out[v * D + i] = buf[i];
}

int launch_rotate_only(const DevIndex& ix, const float* d_in, size_t n, float* d_out, cudaStream_t st, const uint32_t* d_src) {
    if (n == 0) return RBQ_OK;
    const size_t smem = (size_t)ix.D * 2 * sizeof(float);
    const int threads = ix.D >= 512 ? 256 : 128;
    rotate_only_kernel<<<(unsigned)n, threads, smem, st>>>(ix, d_in, d_src, d_out);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// math::dot, AVX2 variant: every lane computes AVX lane (lane & 7); the lane sums are then folded 0..7.
template <class FA, class FB>
__device__ __forceinline__ float dot_avx2_order(int D, int lane, FA a, FB b) {
    const int l = lane & 7;
    float acc = 0.0f;
    for (int i = l; i < D; i += 8) {
        const float p = a(i) * b(i);
        acc = acc + p;
    }
    float sum = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sum = sum + __shfl_sync(0xffffffffu, acc, k);
    return sum;
}

constexpr int kQWarps = 4;

// one warp per vector.  rot: chunk of rotated vectors [n][D]; cents: rotated centroids; list_of[v] = list.
__global__ void __launch_bounds__(kQWarps * 32) quantize_kernel(DevIndex ix, const float* __restrict__ rot,
                                                               const uint32_t* __restrict__ list_of,
                                                               const float* __restrict__ cents, unsigned n,
                                                               float t_const, const double* __restrict__ t_per_vec,
                                                               BuildOut out, const unsigned long long* __restrict__ dst_pos) {
    extern __shared__ __align__(16) unsigned char qsm[];
    const int D = ix.D, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ex = ix.ex_bits;
    const unsigned v = blockIdx.x * kQWarps + warp;
    if (v >= n) return;
    // o: where the vector's outputs go (streaming build: its final position inside its list; ~0 = not on this shard)
    unsigned long long o = v;
    if (dst_pos != nullptr) {
        o = dst_pos[v];
        if (o == ~0ull) return;
    }
    float* r = reinterpret_cast<float*>(qsm) + (size_t)warp * 2 * D;  // residual
    float* oa = r + D;                                               // |r| / norm
    unsigned short* code = reinterpret_cast<unsigned short*>(qsm + (size_t)kQWarps * 2 * D * 4) + (size_t)warp * D;
    const float* ce = cents + (size_t)list_of[v] * D;
    const float* x = rot + (size_t)v * D;
    const int maxv = (1 << ex) - 1;

    for (int i = lane; i < D; i += 32) {
        r[i] = x[i] - ce[i];  // math::subtract
        code[i] = 0;
    }
    __syncwarp();
    float ipnorm_inv = 1.0f;
    if (ex > 0) {  // ex_bits_code_with_inv + quantize_ex_with_inv
        float s = 0.0f;
        if (lane == 0)
            for (int i = 0; i < D; ++i) {
                const float a = fabsf(r[i]);
                const float p = a * a;
                s = s + p;
            }
        const float norm = sqrtf(__shfl_sync(0xffffffffu, s, 0));
        if (norm > kF32Eps) {
            const double t = t_per_vec ? t_per_vec[v] : (double)t_const;
            for (int i = lane; i < D; i += 32) {
                const float o = fabsf(r[i]) / norm;
                oa[i] = o;
                const double tv = t * (double)o;
                int c = (int)(tv + 1e-5);
                if (c > maxv) c = maxv;
                code[i] = (unsigned short)c;
            }
            __syncwarp();
            double ipn = 0.0;
            if (lane == 0)
                for (int i = 0; i < D; ++i) {
                    const double m = ((double)code[i] + 0.5) * (double)oa[i];
                    ipn = ipn + m;
                }
            ipn = __shfl_sync(0xffffffffu, ipn, 0);
            ipnorm_inv = (isfinite(ipn) && ipn > 0.0) ? (float)(1.0 / ipn) : 1.0f;
            if (!isfinite(ipnorm_inv)) ipnorm_inv = 1.0f;
            for (int i = lane; i < D; i += 32)
                if (r[i] < 0.0f) code[i] = (unsigned short)((~code[i]) & maxv);
            __syncwarp();
        }
    }
    const float cb = -((float)(1 << ex) - 0.5f);
    auto R = [&](int i) { return r[i]; };
    auto C = [&](int i) { return ce[i]; };
    auto XU = [&](int i) { return (r[i] >= 0.0f ? 1.0f : 0.0f) - 0.5f; };
    auto QS = [&](int i) { return (float)(unsigned short)(code[i] + ((r[i] >= 0.0f ? 1 : 0) << ex)) + cb; };
    // compute_one_bit_factors (quantizer.rs:264-308)
    const float l2 = dot_avx2_order(D, lane, R, R);
    const float l2n = sqrtf(l2);
    const float xun = dot_avx2_order(D, lane, XU, XU);
    const float ip_r = dot_avx2_order(D, lane, R, XU);
    const float ip_c = dot_avx2_order(D, lane, C, XU);
    const float drc = dot_avx2_order(D, lane, R, C);
    float denom = ip_r;
    if (fabsf(denom) <= kF32Eps) denom = INFINITY;
    float tmp_err = 0.0f;
    if (D > 1) {
        const float dd = denom * denom;
        const float ratio = ((l2 * xun) / dd) - 1.0f;
        if (isfinite(ratio) && ratio > 0.0f) {
            const float q = fmaxf(ratio / (float)(D - 1), 0.0f);
            const float a = l2n * 1.9f;
            tmp_err = a * sqrtf(q);
        }
    }
    float f_add, f_rescale, f_error;
    if (ix.metric == RBQ_METRIC_L2) {
        const float t0 = 2.0f * l2;
        const float t1 = t0 * ip_c;
        f_add = l2 + t1 / denom;
        f_rescale = (-2.0f * l2) / denom;
        f_error = 2.0f * tmp_err;
    } else {
        const float t0 = 1.0f - drc;
        const float t1 = l2 * ip_c;
        f_add = t0 + t1 / denom;
        f_rescale = (-l2) / denom;
        f_error = tmp_err;
    }
    // delta / vl (reconstruction parameters, quantizer.rs:172-187) and the ex factors (:475-535)
    const float nq2 = dot_avx2_order(D, lane, QS, QS);
    const float drq = dot_avx2_order(D, lane, R, QS);
    const float nq = sqrtf(nq2);
    const float den2 = fmaxf(l2n * nq, kF32Eps);
    const float cosv = fminf(fmaxf(drq / den2, -1.0f), 1.0f);
    const float delta = nq <= kF32Eps ? 0.0f : (l2n / nq) * cosv;
    const float vl = delta * cb;
    float f_add_ex = 0.0f, f_rescale_ex = 0.0f;
    if (ex > 0) {
        const float ip_cx = dot_avx2_order(D, lane, C, QS);
        const float safe = fabsf(drq) <= kF32Eps ? INFINITY : drq;
        if (ix.metric == RBQ_METRIC_L2) {
            const float t0 = 2.0f * l2;
            const float t1 = t0 * ip_cx;
            f_add_ex = l2 + t1 / safe;
            const float t2 = -2.0f * l2n;
            f_rescale_ex = t2 * ipnorm_inv;
        } else {
            const float t0 = 1.0f - drc;
            const float t1 = l2 * ip_cx;
            f_add_ex = t0 + t1 / safe;
            f_rescale_ex = (-l2n) * ipnorm_inv;
        }
    }
    if (lane == 0) {
        out.f_add[o] = f_add;
        out.f_rescale[o] = f_rescale;
        out.f_error[o] = f_error;
        out.f_add_ex[o] = f_add_ex;
        out.f_rescale_ex[o] = f_rescale_ex;
        out.delta[o] = delta;
        out.vl[o] = vl;
        if (out.rnorm) out.rnorm[o] = l2n;
    }
    // sign codes, MSB-first (simd.rs:141-150)
    uint8_t* brow = out.bin_rows + (size_t)o * (D / 8);
    for (int b = lane; b < D / 8; b += 32) {
        unsigned byte = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) byte |= (r[8 * b + k] >= 0.0f ? 1u : 0u) << (7 - k);
        brow[b] = (uint8_t)byte;
    }
    // ex codes (quantizer.rs:212-243 -> simd.rs:2406-2695 for 1/2/6 bits, generic LSB-first otherwise)
    if (ex > 0) {
        uint8_t* erow = out.ex + (size_t)o * ix.ex_stride;
        for (int ch = lane; ch < D / 16; ch += 32) {
            const unsigned short* c = code + 16 * ch;
            if (ex == 2) {
                uint32_t w = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) w |= (uint32_t)(c[k] & 3) << (8 * (k & 3) + 2 * (k >> 2));
                *reinterpret_cast<uint32_t*>(erow + 4 * ch) = w;
            } else if (ex == 6) {
                uint32_t lo0 = 0, lo1 = 0, w = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    lo0 |= (uint32_t)((c[k] & 15) | ((c[k + 8] & 15) << 4)) << (8 * k);
                    lo1 |= (uint32_t)((c[k + 4] & 15) | ((c[k + 12] & 15) << 4)) << (8 * k);
                }
#pragma unroll
                for (int k = 0; k < 16; ++k) w |= (uint32_t)((c[k] >> 4) & 3) << (8 * (k & 3) + 2 * (k >> 2));
                uint32_t* o = reinterpret_cast<uint32_t*>(erow + 12 * ch);
                o[0] = lo0;
                o[1] = lo1;
                o[2] = w;
            } else {
                unsigned long long lo = 0, hi = 0;  // 16*ex bits, LSB first
                for (int k = 0; k < 16; ++k) {
                    const int pos = k * ex;
                    const unsigned long long cv = c[k] & maxv;
                    if (pos < 64) {
                        lo |= cv << pos;
                        if (pos + ex > 64) hi |= cv >> (64 - pos);
                    } else {
                        hi |= cv << (pos - 64);
                    }
                }
                uint8_t* o = erow + 2 * ex * ch;
                for (int b = 0; b < 2 * ex; ++b) o[b] = (uint8_t)((b < 8 ? lo >> (8 * b) : hi >> (8 * (b - 8))) & 0xff);
            }
        }
    }
}

int launch_build_quantize(const DevIndex& ix, const float* d_rot, const uint32_t* d_list_of, size_t n,
                          const float* d_cents, float t_const, const double* d_t_per_vec, BuildOut out,
                          cudaStream_t st, const unsigned long long* d_dst_pos) {
    if (n == 0) return RBQ_OK;
    const size_t smem = (size_t)kQWarps * ix.D * (2 * 4 + 2);
    if (smem > 48 * 1024)
        RBQ_CUDA(cudaFuncSetAttribute(quantize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    quantize_kernel<<<(unsigned)((n + kQWarps - 1) / kQWarps), kQWarps * 32, smem, st>>>(ix, d_rot, d_list_of, d_cents,
                                                                                      (unsigned)n, t_const, d_t_per_vec, out, d_dst_pos);
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

// FastScan block packing: one CTA per 32-vector block, one thread per byte column (8 dims).
// simd::pack_codes (reference src/simd.rs:864-904) + factor arrays (reference src/ivf.rs:566-592).
__global__ void pack_blocks_kernel(int D, uint32_t block_stride, const uint8_t* __restrict__ bin_rows,
                                   const float* __restrict__ f_add, const float* __restrict__ f_rescale,
                                   const float* __restrict__ f_error, const uint32_t* __restrict__ blk_list,
                                   const uint32_t* __restrict__ list_n, const uint32_t* __restrict__ blk_off,
                                   const uint64_t* __restrict__ vec_off, uint8_t* __restrict__ blocks) {
    const uint32_t gb = blockIdx.x, c = blk_list[gb], b = gb - blk_off[c];
    const uint32_t cnt = min(32u, list_n[c] - 32u * b);
    const size_t v0 = vec_off[c] + (size_t)32 * b;
    uint8_t* blk = blocks + (size_t)gb * block_stride;
    const int db = D / 8;
    for (int col = threadIdx.x; col < db; col += blockDim.x) {
        uint8_t cd[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) cd[i] = (uint32_t)i < cnt ? bin_rows[(v0 + i) * db + col] : (uint8_t)0;
        uint32_t w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) w[k] = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int a = (j >> 1) + ((j & 1) << 3);  // KPERM0[j]
            const uint32_t up = (uint32_t)(cd[a] >> 4) | ((uint32_t)(cd[a + 16] >> 4) << 4);
            const uint32_t lo = (uint32_t)(cd[a] & 15) | ((uint32_t)(cd[a + 16] & 15) << 4);
            w[j >> 2] |= up << (8 * (j & 3));
            w[4 + (j >> 2)] |= lo << (8 * (j & 3));
        }
        uint4* o = reinterpret_cast<uint4*>(blk + (size_t)col * 32);
        o[0] = make_uint4(w[0], w[1], w[2], w[3]);
        o[1] = make_uint4(w[4], w[5], w[6], w[7]);
    }
    if (threadIdx.x < 32) {
        float* fac = reinterpret_cast<float*>(blk + (size_t)D * 4);
        const bool have = threadIdx.x < cnt;
        fac[threadIdx.x] = have ? f_add[v0 + threadIdx.x] : 0.0f;
        fac[32 + threadIdx.x] = have ? f_rescale[v0 + threadIdx.x] : 0.0f;
        fac[64 + threadIdx.x] = have ? f_error[v0 + threadIdx.x] : 0.0f;
    }
}

}  // namespace rbq

// ---- C entry ---------------------------------------------------------------------------------------
using namespace rbq;

namespace {
struct DeviceGuardLite {
    int prev = -1;
    explicit DeviceGuardLite(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuardLite() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};
struct Scratch {  // frees temporary device buffers on every exit path
    std::vector<void*> ptrs;
    ~Scratch() {
        for (void* p : ptrs) cudaFree(p);
    }
    template <class T>
    int alloc(T** out, size_t count) {
        void* d = nullptr;
        RBQ_CUDA(cudaMalloc(&d, std::max<size_t>(count * sizeof(T), 16)));
        ptrs.push_back(d);
        *out = reinterpret_cast<T*>(d);
        return RBQ_OK;
    }
};
template <class T>
int persist(rbq_index* h, T** out, size_t count, size_t pad = 0) {
    void* d = nullptr;
    RBQ_CUDA(cudaMalloc(&d, std::max<size_t>(count * sizeof(T) + pad, 16)));
    h->allocations.push_back(d);
    RBQ_CUDA(cudaMemset(d, 0, std::max<size_t>(count * sizeof(T) + pad, 16)));
    *out = reinterpret_cast<T*>(d);
    return RBQ_OK;
}

// Orthonormal matrix for the MatrixRotator (Gaussian rows + Gram-Schmidt, reference src/rotation.rs:87-139).
void make_matrix(size_t D, uint64_t seed, std::vector<uint8_t>& bytes) {
    uint64_t st = seed;
    std::vector<double> m(D * D);
    for (size_t r = 0; r < D; ++r) {
        for (;;) {
            for (size_t k = 0; k < D; ++k) m[r * D + k] = gaussian(st);
            for (size_t p = 0; p < r; ++p) {
                double dp = 0;
                for (size_t k = 0; k < D; ++k) dp += m[r * D + k] * m[p * D + k];
                for (size_t k = 0; k < D; ++k) m[r * D + k] -= dp * m[p * D + k];
            }
            double nn = 0;
            for (size_t k = 0; k < D; ++k) nn += m[r * D + k] * m[r * D + k];
            nn = std::sqrt(nn);
            if (nn > 1e-9) {
                for (size_t k = 0; k < D; ++k) m[r * D + k] /= nn;
                break;
            }
        }
    }
    bytes.resize(D * D * 4);
    float* f = reinterpret_cast<float*>(bytes.data());
    for (size_t i = 0; i < D * D; ++i) f[i] = (float)m[i];
}
}  // namespace

extern "C" int rbq_index_build(const float* data, size_t n, size_t dim, const float* centroids, size_t nlist,
                               const uint32_t* assignments, int total_bits, int metric, int rotator_type, uint64_t seed,
                               int faster_config, const uint8_t* rotator_state, int device, rbq_index** out) {
    if (!out) return fail(RBQ_INVALID_CONFIG, "null argument");
    *out = nullptr;
    // validation order and messages: reference src/ivf.rs:1035-1079
    if (n == 0 || !data) return fail(RBQ_INVALID_CONFIG, "training data must be non-empty");
    if (nlist == 0 || !centroids) return fail(RBQ_INVALID_CONFIG, "centroids must be non-empty");
    if (!assignments) return fail(RBQ_INVALID_CONFIG, "assignments length must match data length");
    if (total_bits < 1 || total_bits > 16) return fail(RBQ_INVALID_CONFIG, "total_bits must be between 1 and 16");
    if (total_bits > 9) return fail(RBQ_INVALID_CONFIG, "total_bits above 9 (ex_bits > 8) is not supported");
    if (dim == 0) return fail(RBQ_INVALID_CONFIG, "input vectors must share the same dimension");
    if (nlist > n) return fail(RBQ_INVALID_CONFIG, "nlist cannot exceed number of vectors");
    if (metric != RBQ_METRIC_L2 && metric != RBQ_METRIC_INNER_PRODUCT) return fail(RBQ_INVALID_CONFIG, "unknown metric");
    if (rotator_type != RBQ_ROTATOR_MATRIX && rotator_type != RBQ_ROTATOR_FHT_KAC)
        return fail(RBQ_INVALID_CONFIG, "unknown rotator type");
    for (size_t i = 0; i < n; ++i)
        if (assignments[i] >= nlist) return fail(RBQ_INVALID_CONFIG, "assignments reference invalid cluster ids");

    rbq_index* h = new rbq_index();
    struct Guard {
        rbq_index* h;
        bool ok = false;
        ~Guard() {
            if (!ok) rbq_index_free(h);
        }
    } guard{h};
    h->device = device;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    RBQ_CUDA(cudaSetDevice(device));
    struct Restore {
        int d;
        ~Restore() { cudaSetDevice(d); }
    } restore{prev_dev};

    HostIndex& hi = h->host;
    hi.dim = (uint32_t)dim;
    hi.D = (uint32_t)(rotator_type == RBQ_ROTATOR_FHT_KAC ? (dim + 63) / 64 * 64 : dim);  // rotation.rs:27-32
    hi.metric = metric;
    hi.rot_type = rotator_type;
    hi.ex_bits = total_bits - 1;
    hi.nlist = nlist;
    hi.nvec_total = n;
    const size_t D = hi.D, exs = hi.ex_stride(), stride = hi.block_stride();
    if (D % 16 != 0)
        return fail(RBQ_INVALID_CONFIG, "padded_dim must be a multiple of 16 (FastScan requirement, reference src/simd.rs:978-981)");
    if (D > 2048) return fail(RBQ_INVALID_CONFIG, "padded_dim > 2048 (high-accuracy LUT path) is not supported");
    // rotator state
    if (rotator_type == RBQ_ROTATOR_FHT_KAC) {
        hi.rot_bytes.resize(4 * D / 8);
        if (rotator_state) std::memcpy(hi.rot_bytes.data(), rotator_state, hi.rot_bytes.size());
        else {
            uint64_t st = seed;
            for (auto& b : hi.rot_bytes) b = (uint8_t)(splitmix64(st) >> 56);
        }
    } else if (rotator_state) {
        hi.rot_bytes.assign(rotator_state, rotator_state + D * D * 4);
    } else {
        make_matrix(D, seed, hi.rot_bytes);
    }
    const float t_const = (faster_config && hi.ex_bits > 0) ? const_scaling_factor_host(D, hi.ex_bits, seed) : -1.0f;
    const bool per_vec_t = hi.ex_bits > 0 && !faster_config;

    // group vector indices by list, ascending index inside a list (ivf.rs:1141-1149)
    hi.list_n_all.assign(nlist, 0);
    for (size_t i = 0; i < n; ++i) hi.list_n_all[assignments[i]]++;
    hi.list_n = hi.list_n_all;
    hi.blk_off.assign(nlist + 1, 0);
    hi.vec_off.assign(nlist + 1, 0);
    uint64_t nblk = 0;
    for (size_t c = 0; c < nlist; ++c) {
        if (hi.list_n[c] > 1000000)
            return fail(RBQ_INVALID_CONFIG, "a list holds more than 1,000,000 vectors (the RBQ1 loader's per-cluster limit)");
        hi.blk_off[c] = (uint32_t)nblk;
        hi.vec_off[c + 1] = hi.vec_off[c] + hi.list_n[c];
        nblk += (hi.list_n[c] + kBatch - 1) / kBatch;
    }
    if (nblk > 0xFFFFFFFFull) return fail(RBQ_INVALID_CONFIG, "more than 2^32 blocks");
    hi.blk_off[nlist] = (uint32_t)nblk;
    std::vector<uint64_t> order(n);
    std::vector<uint32_t> list_of(n), blk_list(nblk);
    {
        std::vector<uint64_t> cursor(hi.vec_off.begin(), hi.vec_off.end() - 1);
        for (size_t i = 0; i < n; ++i) {
            const uint64_t p = cursor[assignments[i]]++;
            order[p] = i;
            list_of[p] = assignments[i];
        }
        for (size_t c = 0; c < nlist; ++c)
            for (uint32_t b = hi.blk_off[c]; b < hi.blk_off[c + 1]; ++b) blk_list[b] = (uint32_t)c;
    }

    // geometry + rotator on the device
    auto fill = [&]() {
        DevIndex& d = h->dev;
        d.dim = (int)dim;
        d.D = (int)D;
        d.metric = metric;
        d.ex_bits = hi.ex_bits;
        d.rot_type = rotator_type;
        int lg = 0;
        while ((2u << lg) <= dim) ++lg;
        d.trunc = 1 << lg;
        d.fac = 1.0f / std::sqrt((float)d.trunc);
        d.nlist = (uint32_t)nlist;
        d.block_stride = (uint32_t)stride;
        d.ex_stride = (uint32_t)exs;
    };
    fill();
    DevIndex& dv = h->dev;
    int rc;
    {
        uint8_t* d_flip = nullptr;
        float* d_mt = nullptr;
        if (rotator_type == RBQ_ROTATOR_FHT_KAC) {
            if ((rc = persist(h, &d_flip, hi.rot_bytes.size()))) return rc;
            RBQ_CUDA(cudaMemcpy(d_flip, hi.rot_bytes.data(), hi.rot_bytes.size(), cudaMemcpyHostToDevice));
        } else {
            std::vector<float> mt(D * D);
            const float* m = reinterpret_cast<const float*>(hi.rot_bytes.data());
            for (size_t r = 0; r < D; ++r)
                for (size_t k = 0; k < D; ++k) mt[k * D + r] = m[r * D + k];
            if ((rc = persist(h, &d_mt, D * D))) return rc;
            RBQ_CUDA(cudaMemcpy(d_mt, mt.data(), D * D * 4, cudaMemcpyHostToDevice));
        }
        dv.flip = d_flip;
        dv.matrix_t = d_mt;
    }
    // persistent arrays
    float *d_cent = nullptr, *d_fae = nullptr, *d_fre = nullptr;
    uint32_t *d_list_n = nullptr, *d_blk_off = nullptr;
    uint64_t *d_vec_off = nullptr, *d_ids = nullptr;
    uint8_t *d_blocks = nullptr, *d_ex = nullptr;
    if ((rc = persist(h, &d_cent, nlist * D))) return rc;
    if ((rc = persist(h, &d_list_n, nlist))) return rc;
    if ((rc = persist(h, &d_blk_off, nlist + 1))) return rc;
    if ((rc = persist(h, &d_vec_off, nlist + 1))) return rc;
    if ((rc = persist(h, &d_blocks, nblk * stride))) return rc;
    if ((rc = persist(h, &d_ids, n))) return rc;
    if ((rc = persist(h, &d_ex, n * exs, 16))) return rc;
    if ((rc = persist(h, &d_fae, n))) return rc;
    if ((rc = persist(h, &d_fre, n))) return rc;
    RBQ_CUDA(cudaMemcpy(d_list_n, hi.list_n.data(), nlist * 4, cudaMemcpyHostToDevice));
    RBQ_CUDA(cudaMemcpy(d_blk_off, hi.blk_off.data(), (nlist + 1) * 4, cudaMemcpyHostToDevice));
    RBQ_CUDA(cudaMemcpy(d_vec_off, hi.vec_off.data(), (nlist + 1) * 8, cudaMemcpyHostToDevice));
    RBQ_CUDA(cudaMemcpy(d_ids, order.data(), n * 8, cudaMemcpyHostToDevice));  // ids = positions in the training slice

    Scratch tmp;
    const size_t CH = std::min<size_t>(n, std::max<size_t>(4096, ((size_t)384 << 20) / (D * 4)));
    float *d_in = nullptr, *d_rot = nullptr, *d_fa = nullptr, *d_fr = nullptr, *d_fe = nullptr, *d_delta = nullptr, *d_vl = nullptr;
    uint8_t* d_bin = nullptr;
    uint32_t *d_list_of = nullptr, *d_blk_list = nullptr;
    double* d_t = nullptr;
    if ((rc = tmp.alloc(&d_in, std::max(CH, nlist) * dim))) return rc;
    if ((rc = tmp.alloc(&d_rot, CH * D))) return rc;
    if ((rc = tmp.alloc(&d_bin, n * (D / 8)))) return rc;
    if ((rc = tmp.alloc(&d_fa, n))) return rc;
    if ((rc = tmp.alloc(&d_fr, n))) return rc;
    if ((rc = tmp.alloc(&d_fe, n))) return rc;
    if ((rc = tmp.alloc(&d_delta, n))) return rc;
    if ((rc = tmp.alloc(&d_vl, n))) return rc;
    if ((rc = tmp.alloc(&d_list_of, n))) return rc;
    if ((rc = tmp.alloc(&d_blk_list, std::max<size_t>(nblk, 1)))) return rc;
    if (per_vec_t && (rc = tmp.alloc(&d_t, CH))) return rc;
    RBQ_CUDA(cudaMemcpy(d_list_of, list_of.data(), n * 4, cudaMemcpyHostToDevice));
    if (nblk) RBQ_CUDA(cudaMemcpy(d_blk_list, blk_list.data(), nblk * 4, cudaMemcpyHostToDevice));

    // rotate centroids (ivf.rs:1088-1089)
    for (size_t c0 = 0; c0 < nlist; c0 += CH) {
        const size_t m = std::min(CH, nlist - c0);
        RBQ_CUDA(cudaMemcpy(d_in, centroids + c0 * dim, m * dim * 4, cudaMemcpyHostToDevice));
        if ((rc = launch_rotate_only(dv, d_in, m, d_cent + c0 * D, nullptr))) return rc;
    }
    hi.centroids.resize(nlist * D);
    RBQ_CUDA(cudaMemcpy(hi.centroids.data(), d_cent, nlist * D * 4, cudaMemcpyDeviceToHost));

    // rotate + quantise the vectors in list order, chunk by chunk
    float* staging = nullptr;
    RBQ_CUDA(cudaMallocHost(&staging, CH * dim * 4));
    struct Pinned {
        float* p;
        ~Pinned() { cudaFreeHost(p); }
    } pinned{staging};
    std::vector<float> rot_host;
    std::vector<double> t_host;
    for (size_t p0 = 0; p0 < n; p0 += CH) {
        const size_t m = std::min(CH, n - p0);
        for (size_t i = 0; i < m; ++i) std::memcpy(staging + i * dim, data + order[p0 + i] * dim, dim * 4);
        RBQ_CUDA(cudaMemcpy(d_in, staging, m * dim * 4, cudaMemcpyHostToDevice));
        if ((rc = launch_rotate_only(dv, d_in, m, d_rot, nullptr))) return rc;
        if (per_vec_t) {
            // precise mode (RabitqConfig::new): optimal rescale factor per vector, computed on the host
            rot_host.resize(m * D);
            t_host.resize(m);
            RBQ_CUDA(cudaMemcpy(rot_host.data(), d_rot, m * D * 4, cudaMemcpyDeviceToHost));
            const unsigned nth = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
            std::vector<std::thread> pool;
            const int ex_bits = hi.ex_bits;
            for (unsigned t = 0; t < nth; ++t)
                pool.emplace_back([&, t]() {
                    std::vector<float> oa(D);
                    for (size_t i = t; i < m; i += nth) {
                        const float* x = &rot_host[i * D];
                        const float* ce = &hi.centroids[(size_t)list_of[p0 + i] * D];
                        float s = 0.0f;
                        for (size_t k = 0; k < D; ++k) {
                            const float a = std::fabs(x[k] - ce[k]);
                            oa[k] = a;
                            const float pr = a * a;
                            s = s + pr;
                        }
                        const float norm = std::sqrt(s);
                        if (norm <= kF32Eps) {
                            t_host[i] = 1.0;
                            continue;
                        }
                        for (size_t k = 0; k < D; ++k) oa[k] = oa[k] / norm;
                        t_host[i] = best_rescale_factor_host(oa.data(), D, ex_bits);
                    }
                });
            for (auto& th : pool) th.join();
            RBQ_CUDA(cudaMemcpy(d_t, t_host.data(), m * 8, cudaMemcpyHostToDevice));
        }
        BuildOut bo;
        bo.bin_rows = d_bin + p0 * (D / 8);
        bo.ex = d_ex + p0 * exs;
        bo.f_add = d_fa + p0;
        bo.f_rescale = d_fr + p0;
        bo.f_error = d_fe + p0;
        bo.f_add_ex = d_fae + p0;
        bo.f_rescale_ex = d_fre + p0;
        bo.delta = d_delta + p0;
        bo.vl = d_vl + p0;
        if ((rc = launch_build_quantize(dv, d_rot, d_list_of + p0, m, d_cent, t_const, per_vec_t ? d_t : nullptr, bo, nullptr)))
            return rc;
        RBQ_CUDA(cudaDeviceSynchronize());
    }
    if (nblk) {
        pack_blocks_kernel<<<(unsigned)nblk, 128>>>((int)D, (uint32_t)stride, d_bin, d_fa, d_fr, d_fe, d_blk_list, d_list_n,
                                                    d_blk_off, d_vec_off, d_blocks);
        RBQ_CUDA(cudaGetLastError());
    }
    hi.delta.resize(n);
    hi.vl.resize(n);
    RBQ_CUDA(cudaMemcpy(hi.delta.data(), d_delta, n * 4, cudaMemcpyDeviceToHost));
    RBQ_CUDA(cudaMemcpy(hi.vl.data(), d_vl, n * 4, cudaMemcpyDeviceToHost));
    RBQ_CUDA(cudaDeviceSynchronize());

    dv.centroids = d_cent;
    dv.list_n = d_list_n;
    dv.max_list_n = 0;
    for (uint32_t c : hi.list_n) dv.max_list_n = std::max(dv.max_list_n, c);
    dv.blk_off = d_blk_off;
    dv.vec_off = d_vec_off;
    dv.blocks = d_blocks;
    dv.ids = d_ids;
    dv.ex = d_ex;
    dv.f_add_ex = d_fae;
    dv.f_rescale_ex = d_fre;
    if ((rc = prepare_coarse_tc(h))) return rc;
    if ((rc = prepare_coarse_sample(h))) return rc;
    if ((rc = prepare_ex_lanes(h))) return rc;
    void* stp = nullptr;
    RBQ_CUDA(cudaMalloc(&stp, sizeof(DevStats) + 64));
    RBQ_CUDA(cudaMemset(stp, 0, sizeof(DevStats) + 64));
    h->allocations.push_back(stp);
    h->d_stats = reinterpret_cast<DevStats*>(stp);
    for (auto& e : h->ev) cudaEventCreate(&e);
    guard.ok = true;
    *out = h;
    return RBQ_OK;
}

// ---- streaming build: device-resident data, chunk by chunk -----------------------------------------------------------
// For indexes too large to hand over as one host array (BASELINE config 5: 100M x 128).  The caller clusters the data first
// (any k-means: rbq_kmeans_* below or its own), tells the builder how many vectors every list will receive, and then feeds
// chunks of (vectors, assignments) that already live on the device, in ascending id order.  Every chunk is sorted by list
// (stable), rotated, quantised and scattered to its final place inside the list-concatenated arrays -- the same kernels and
// the same bytes as rbq_index_build; with shard_count > 1 only the lists of shard_rank are kept (the same shard a load of the
// complete file would keep).
#include <cub/device/device_radix_sort.cuh>

struct rbq_builder {
    rbq_index* h = nullptr;
    float t_const = -1.0f;
    size_t n_local = 0, nblk = 0, chunk_cap = 0;
    uint64_t added = 0;  // vectors offered so far (all shards)
    // persistent outputs (owned by h->allocations) and temporaries (freed at finish)
    float *d_cent = nullptr, *d_fae = nullptr, *d_fre = nullptr;
    uint32_t *d_list_n = nullptr, *d_blk_off = nullptr;
    uint64_t *d_vec_off = nullptr, *d_ids = nullptr;
    uint8_t *d_blocks = nullptr, *d_ex = nullptr, *d_owner = nullptr;
    std::vector<void*> tmp;
    uint8_t* d_bin = nullptr;
    float *d_fa = nullptr, *d_fr = nullptr, *d_fe = nullptr, *d_delta = nullptr, *d_vl = nullptr, *d_rot = nullptr;
    uint32_t *d_cursor = nullptr, *d_keys_in = nullptr, *d_keys = nullptr, *d_idx_in = nullptr, *d_idx = nullptr, *d_blk_list = nullptr;
    unsigned long long* d_dst = nullptr;
    void* d_sort_tmp = nullptr;
    size_t sort_tmp_bytes = 0;
    int key_bits = 32;
    ~rbq_builder() {
        for (void* p : tmp) cudaFree(p);
    }
};

namespace {
// key = list id for vectors this shard keeps, ~0 otherwise (sorts last); idx = position in the chunk
__global__ void builder_keys_kernel(const uint32_t* __restrict__ assign, const uint8_t* __restrict__ owner, int shard_rank, uint32_t nlist,
                                    uint32_t m, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx, unsigned int* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint32_t c = assign[i];
    uint32_t k = 0xffffffffu;
    if (c >= nlist) atomicAdd(bad, 1u);
    else if (owner == nullptr || owner[c] == (uint8_t)shard_rank) k = c;
    keys[i] = k;
    idx[i] = i;
}
// sorted (key, idx): final position of every kept vector = list start + vectors of the list that arrived earlier + rank inside
// the chunk's run of that list; the head of every run bumps the list's cursor by the run length
__global__ void builder_place_kernel(const uint32_t* __restrict__ keys, uint32_t* __restrict__ idx, uint32_t m, const uint64_t* __restrict__ vec_off,
                                     const uint32_t* __restrict__ list_n, uint32_t* __restrict__ cursor, unsigned long long id_base,
                                     unsigned long long* __restrict__ dst, uint64_t* __restrict__ ids, unsigned int* __restrict__ bad) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t c = keys[j];
    if (c == 0xffffffffu) {
        dst[j] = ~0ull;
        idx[j] = 0xffffffffu;
        return;
    }
    uint32_t lo = 0, hi = j;  // first position holding key c
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] < c) lo = mid + 1;
        else hi = mid;
    }
    const uint32_t within = cursor[c] + (j - lo);  // cursor is only advanced by the next kernel
    if (within >= list_n[c]) {
        atomicAdd(bad, 1u);
        dst[j] = ~0ull;
        idx[j] = 0xffffffffu;
        return;
    }
    const unsigned long long p = vec_off[c] + within;
    dst[j] = p;
    ids[p] = id_base + idx[j];
}
__global__ void builder_advance_kernel(const uint32_t* __restrict__ keys, uint32_t m, uint32_t* __restrict__ cursor) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const uint32_t c = keys[j];
    if (c == 0xffffffffu || (j > 0 && keys[j - 1] == c)) return;  // heads of runs only
    uint32_t lo = j, hi = m;  // first position past the run
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (keys[mid] <= c) lo = mid + 1;
        else hi = mid;
    }
    cursor[c] += lo - j;
}
template <class T>
int btmp(rbq_builder* b, T** out, size_t count) {
    void* d = nullptr;
    RBQ_CUDA(cudaMalloc(&d, std::max<size_t>(count * sizeof(T), 16)));
    b->tmp.push_back(d);
    *out = reinterpret_cast<T*>(d);
    return RBQ_OK;
}
}  // namespace

extern "C" void rbq_builder_free(rbq_builder* b) {
    if (!b) return;
    if (b->h) {
        int prev = 0;
        cudaGetDevice(&prev);
        cudaSetDevice(b->h->device);
        for (void* p : b->tmp) cudaFree(p);
        b->tmp.clear();
        rbq_index_free(b->h);
        cudaSetDevice(prev);
    }
    delete b;
}

extern "C" int rbq_builder_create(size_t dim, const float* centroids, size_t nlist, const uint32_t* list_sizes, int total_bits, int metric,
                                  int rotator_type, uint64_t seed, const uint8_t* rotator_state, int device, int shard_rank, int shard_count,
                                  size_t max_chunk, rbq_builder** out) {
    if (!out) return fail(RBQ_INVALID_CONFIG, "null argument");
    *out = nullptr;
    if (nlist == 0 || !centroids || !list_sizes) return fail(RBQ_INVALID_CONFIG, "centroids must be non-empty");
    if (total_bits < 1 || total_bits > 16) return fail(RBQ_INVALID_CONFIG, "total_bits must be between 1 and 16");
    if (total_bits > 9) return fail(RBQ_INVALID_CONFIG, "total_bits above 9 (ex_bits > 8) is not supported");
    if (dim == 0) return fail(RBQ_INVALID_CONFIG, "input vectors must share the same dimension");
    if (metric != RBQ_METRIC_L2 && metric != RBQ_METRIC_INNER_PRODUCT) return fail(RBQ_INVALID_CONFIG, "unknown metric");
    if (rotator_type != RBQ_ROTATOR_MATRIX && rotator_type != RBQ_ROTATOR_FHT_KAC) return fail(RBQ_INVALID_CONFIG, "unknown rotator type");
    if (shard_count < 1 || shard_count > kMaxShards || shard_rank < 0 || shard_rank >= shard_count)
        return fail(RBQ_INVALID_CONFIG, "shard_rank/shard_count out of range");
    if (max_chunk == 0 || max_chunk > 0x7fffffffull) return fail(RBQ_INVALID_CONFIG, "max_chunk must be in [1, 2^31)");
    uint64_t n_total = 0;
    for (size_t c = 0; c < nlist; ++c) {
        if (list_sizes[c] > 1000000)
            return fail(RBQ_INVALID_CONFIG, "a list holds more than 1,000,000 vectors (the RBQ1 loader's per-cluster limit)");
        n_total += list_sizes[c];
    }
    if (n_total == 0) return fail(RBQ_INVALID_CONFIG, "training data must be non-empty");
    if (nlist > n_total) return fail(RBQ_INVALID_CONFIG, "nlist cannot exceed number of vectors");

    rbq_builder* b = new rbq_builder();
    rbq_index* h = new rbq_index();
    b->h = h;
    struct Guard {
        rbq_builder* b;
        bool ok = false;
        ~Guard() {
            if (!ok) rbq_builder_free(b);
        }
    } guard{b};
    h->device = device;
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    RBQ_CUDA(cudaSetDevice(device));
    struct Restore {
        int d;
        ~Restore() { cudaSetDevice(d); }
    } restore{prev_dev};

    HostIndex& hi = h->host;
    hi.dim = (uint32_t)dim;
    hi.D = (uint32_t)(rotator_type == RBQ_ROTATOR_FHT_KAC ? (dim + 63) / 64 * 64 : dim);
    hi.metric = metric;
    hi.rot_type = rotator_type;
    hi.ex_bits = total_bits - 1;
    hi.nlist = nlist;
    hi.nvec_total = n_total;
    hi.shard_rank = shard_rank;
    hi.shard_count = shard_count;
    const size_t D = hi.D, exs = hi.ex_stride(), stride = hi.block_stride();
    if (D % 16 != 0)
        return fail(RBQ_INVALID_CONFIG, "padded_dim must be a multiple of 16 (FastScan requirement, reference src/simd.rs:978-981)");
    if (D > 2048) return fail(RBQ_INVALID_CONFIG, "padded_dim > 2048 (high-accuracy LUT path) is not supported");
    if (rotator_type == RBQ_ROTATOR_FHT_KAC) {
        hi.rot_bytes.resize(4 * D / 8);
        if (rotator_state) std::memcpy(hi.rot_bytes.data(), rotator_state, hi.rot_bytes.size());
        else {
            uint64_t st = seed;
            for (auto& x : hi.rot_bytes) x = (uint8_t)(splitmix64(st) >> 56);
        }
    } else if (rotator_state) {
        hi.rot_bytes.assign(rotator_state, rotator_state + D * D * 4);
    } else {
        make_matrix(D, seed, hi.rot_bytes);
    }
    b->t_const = hi.ex_bits > 0 ? const_scaling_factor_host(D, hi.ex_bits, seed) : -1.0f;  // RabitqConfig::faster
    hi.list_n_all.assign(list_sizes, list_sizes + nlist);
    int rc;
    if ((rc = shard_layout(hi))) return rc;
    if (shard_count == 1) hi.list_owner.clear();
    const size_t n_local = hi.vec_off[nlist], nblk = hi.blk_off[nlist];
    b->n_local = n_local;
    b->nblk = nblk;
    b->chunk_cap = max_chunk;

    DevIndex& dv = h->dev;
    dv.dim = (int)dim;
    dv.D = (int)D;
    dv.metric = metric;
    dv.ex_bits = hi.ex_bits;
    dv.rot_type = rotator_type;
    int lg = 0;
    while ((2u << lg) <= dim) ++lg;
    dv.trunc = 1 << lg;
    dv.fac = 1.0f / std::sqrt((float)dv.trunc);
    dv.nlist = (uint32_t)nlist;
    dv.block_stride = (uint32_t)stride;
    dv.ex_stride = (uint32_t)exs;
    dv.shard_rank = shard_rank;
    {
        uint8_t* d_flip = nullptr;
        float* d_mt = nullptr;
        if (rotator_type == RBQ_ROTATOR_FHT_KAC) {
            if ((rc = persist(h, &d_flip, hi.rot_bytes.size()))) return rc;
            RBQ_CUDA(cudaMemcpy(d_flip, hi.rot_bytes.data(), hi.rot_bytes.size(), cudaMemcpyHostToDevice));
        } else {
            std::vector<float> mt(D * D);
            const float* m = reinterpret_cast<const float*>(hi.rot_bytes.data());
            for (size_t r = 0; r < D; ++r)
                for (size_t k = 0; k < D; ++k) mt[k * D + r] = m[r * D + k];
            if ((rc = persist(h, &d_mt, D * D))) return rc;
            RBQ_CUDA(cudaMemcpy(d_mt, mt.data(), D * D * 4, cudaMemcpyHostToDevice));
        }
        dv.flip = d_flip;
        dv.matrix_t = d_mt;
    }
    if ((rc = persist(h, &b->d_cent, nlist * D))) return rc;
    if ((rc = persist(h, &b->d_list_n, nlist))) return rc;
    if ((rc = persist(h, &b->d_blk_off, nlist + 1))) return rc;
    if ((rc = persist(h, &b->d_vec_off, nlist + 1))) return rc;
    if ((rc = persist(h, &b->d_blocks, nblk * stride))) return rc;
    if ((rc = persist(h, &b->d_ids, n_local))) return rc;
    if ((rc = persist(h, &b->d_ex, n_local * exs, 16))) return rc;
    if ((rc = persist(h, &b->d_fae, n_local))) return rc;
    if ((rc = persist(h, &b->d_fre, n_local))) return rc;
    if (shard_count > 1) {
        if ((rc = persist(h, &b->d_owner, nlist))) return rc;
        RBQ_CUDA(cudaMemcpy(b->d_owner, hi.list_owner.data(), nlist, cudaMemcpyHostToDevice));
    }
    RBQ_CUDA(cudaMemcpy(b->d_list_n, hi.list_n.data(), nlist * 4, cudaMemcpyHostToDevice));
    RBQ_CUDA(cudaMemcpy(b->d_blk_off, hi.blk_off.data(), (nlist + 1) * 4, cudaMemcpyHostToDevice));
    RBQ_CUDA(cudaMemcpy(b->d_vec_off, hi.vec_off.data(), (nlist + 1) * 8, cudaMemcpyHostToDevice));
    // rotate the centroids (ivf.rs:1088-1089)
    {
        float* d_in = nullptr;
        const size_t CH = std::min<size_t>(nlist, 65536);
        RBQ_CUDA(cudaMalloc(&d_in, CH * dim * 4));
        for (size_t c0 = 0; c0 < nlist; c0 += CH) {
            const size_t m = std::min(CH, nlist - c0);
            cudaError_t e = cudaMemcpy(d_in, centroids + c0 * dim, m * dim * 4, cudaMemcpyHostToDevice);
            if (e == cudaSuccess) rc = launch_rotate_only(dv, d_in, m, b->d_cent + c0 * D, nullptr);
            if (e != cudaSuccess || rc) {
                cudaFree(d_in);
                return rc ? rc : fail(RBQ_CUDA_ERROR, std::string("CUDA error: ") + cudaGetErrorString(e));
            }
        }
        RBQ_CUDA(cudaDeviceSynchronize());
        cudaFree(d_in);
    }
    hi.centroids.resize(nlist * D);
    RBQ_CUDA(cudaMemcpy(hi.centroids.data(), b->d_cent, nlist * D * 4, cudaMemcpyDeviceToHost));
    // temporaries
    if ((rc = btmp(b, &b->d_bin, n_local * (D / 8)))) return rc;
    if ((rc = btmp(b, &b->d_fa, n_local))) return rc;
    if ((rc = btmp(b, &b->d_fr, n_local))) return rc;
    if ((rc = btmp(b, &b->d_fe, n_local))) return rc;
    if ((rc = btmp(b, &b->d_delta, n_local))) return rc;
    if ((rc = btmp(b, &b->d_vl, n_local))) return rc;
    if ((rc = btmp(b, &b->d_rot, max_chunk * D))) return rc;
    if ((rc = btmp(b, &b->d_cursor, nlist + 1))) return rc;
    RBQ_CUDA(cudaMemset(b->d_cursor, 0, (nlist + 1) * 4));
    if ((rc = btmp(b, &b->d_keys_in, max_chunk))) return rc;
    if ((rc = btmp(b, &b->d_keys, max_chunk))) return rc;
    if ((rc = btmp(b, &b->d_idx_in, max_chunk))) return rc;
    if ((rc = btmp(b, &b->d_idx, max_chunk))) return rc;
    if ((rc = btmp(b, &b->d_dst, max_chunk))) return rc;
    if ((rc = btmp(b, &b->d_blk_list, std::max<size_t>(nblk, 1)))) return rc;
    {
        std::vector<uint32_t> blk_list(nblk);
        for (size_t c = 0; c < nlist; ++c)
            for (uint32_t x = hi.blk_off[c]; x < hi.blk_off[c + 1]; ++x) blk_list[x] = (uint32_t)c;
        if (nblk) RBQ_CUDA(cudaMemcpy(b->d_blk_list, blk_list.data(), nblk * 4, cudaMemcpyHostToDevice));
    }
    RBQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, b->sort_tmp_bytes, b->d_keys_in, b->d_keys, b->d_idx_in, b->d_idx, (int)max_chunk, 0, 32,
                                             (cudaStream_t) nullptr));
    {
        char* p = nullptr;
        if ((rc = btmp(b, &p, b->sort_tmp_bytes + 16))) return rc;
        b->d_sort_tmp = p;
    }
    dv.centroids = b->d_cent;
    dv.list_n = b->d_list_n;
    dv.blk_off = b->d_blk_off;
    dv.vec_off = b->d_vec_off;
    guard.ok = true;
    *out = b;
    return RBQ_OK;
}

extern "C" int rbq_builder_add_device(rbq_builder* b, const float* d_data, const uint32_t* d_assign, size_t m, uint64_t id_base, void* stream) {
    if (!b || !b->h) return fail(RBQ_INVALID_CONFIG, "null builder");
    if (m == 0) return RBQ_OK;
    if (!d_data || !d_assign) return fail(RBQ_INVALID_CONFIG, "null argument");
    if (m > b->chunk_cap) return fail(RBQ_INVALID_CONFIG, "chunk larger than the builder's max_chunk");
    rbq_index* h = b->h;
    DeviceGuardLite g(h->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const DevIndex& dv = h->dev;
    const unsigned tb = 256, gb = (unsigned)((m + tb - 1) / tb);
    unsigned int* d_bad = reinterpret_cast<unsigned int*>(b->d_cursor + dv.nlist);  // error counter behind the cursors
    builder_keys_kernel<<<gb, tb, 0, st>>>(d_assign, b->d_owner, dv.shard_rank, dv.nlist, (uint32_t)m, b->d_keys_in, b->d_idx_in, d_bad);
    size_t tmp_bytes = b->sort_tmp_bytes;
    RBQ_CUDA(cub::DeviceRadixSort::SortPairs(b->d_sort_tmp, tmp_bytes, b->d_keys_in, b->d_keys, b->d_idx_in, b->d_idx, (int)m, 0, 32, st));
    builder_place_kernel<<<gb, tb, 0, st>>>(b->d_keys, b->d_idx, (uint32_t)m, b->d_vec_off, b->d_list_n, b->d_cursor, id_base, b->d_dst, b->d_ids,
                                            d_bad);
    builder_advance_kernel<<<gb, tb, 0, st>>>(b->d_keys, (uint32_t)m, b->d_cursor);
    RBQ_CUDA(cudaGetLastError());
    int rc;
    if ((rc = launch_rotate_only(dv, d_data, m, b->d_rot, st, b->d_idx))) return rc;
    BuildOut bo;
    bo.bin_rows = b->d_bin;
    bo.ex = b->d_ex;
    bo.f_add = b->d_fa;
    bo.f_rescale = b->d_fr;
    bo.f_error = b->d_fe;
    bo.f_add_ex = b->d_fae;
    bo.f_rescale_ex = b->d_fre;
    bo.delta = b->d_delta;
    bo.vl = b->d_vl;
    if ((rc = launch_build_quantize(dv, b->d_rot, b->d_keys, m, b->d_cent, b->t_const, nullptr, bo, st, b->d_dst))) return rc;
    b->added += m;
    return RBQ_OK;
}

extern "C" int rbq_builder_finish(rbq_builder* b, rbq_index** out) {
    if (!b || !b->h || !out) return fail(RBQ_INVALID_CONFIG, "null argument");
    *out = nullptr;
    rbq_index* h = b->h;
    DeviceGuardLite g(h->device);
    HostIndex& hi = h->host;
    DevIndex& dv = h->dev;
    const size_t nlist = hi.nlist, D = hi.D;
    RBQ_CUDA(cudaDeviceSynchronize());
    // every list must have received exactly the announced number of vectors, and no chunk may have carried a bad list id
    std::vector<uint32_t> cur(nlist + 1);
    RBQ_CUDA(cudaMemcpy(cur.data(), b->d_cursor, (nlist + 1) * 4, cudaMemcpyDeviceToHost));
    if (cur[nlist] != 0) return fail(RBQ_INVALID_CONFIG, "assignments reference invalid cluster ids or overflow the announced list sizes");
    for (size_t c = 0; c < nlist; ++c)
        if (cur[c] != hi.list_n[c]) return fail(RBQ_INVALID_CONFIG, "the vectors added do not match the announced list sizes");
    if (b->nblk) {
        pack_blocks_kernel<<<(unsigned)b->nblk, 128>>>((int)D, dv.block_stride, b->d_bin, b->d_fa, b->d_fr, b->d_fe, b->d_blk_list, b->d_list_n,
                                                       b->d_blk_off, b->d_vec_off, b->d_blocks);
        RBQ_CUDA(cudaGetLastError());
    }
    hi.delta.resize(b->n_local);
    hi.vl.resize(b->n_local);
    if (b->n_local) {
        RBQ_CUDA(cudaMemcpy(hi.delta.data(), b->d_delta, b->n_local * 4, cudaMemcpyDeviceToHost));
        RBQ_CUDA(cudaMemcpy(hi.vl.data(), b->d_vl, b->n_local * 4, cudaMemcpyDeviceToHost));
    }
    RBQ_CUDA(cudaDeviceSynchronize());
    for (void* p : b->tmp) cudaFree(p);
    b->tmp.clear();
    dv.max_list_n = 0;
    for (uint32_t c : hi.list_n) dv.max_list_n = std::max(dv.max_list_n, c);
    dv.blocks = b->d_blocks;
    dv.ids = b->d_ids;
    dv.ex = b->d_ex;
    dv.f_add_ex = b->d_fae;
    dv.f_rescale_ex = b->d_fre;
    dv.list_owner = hi.shard_count > 1 ? b->d_owner : nullptr;
    int rc;
    if ((rc = prepare_coarse_tc(h))) return rc;
    if ((rc = prepare_coarse_sample(h))) return rc;
    if ((rc = prepare_ex_lanes(h))) return rc;
    void* stp = nullptr;
    RBQ_CUDA(cudaMalloc(&stp, sizeof(DevStats) + 64));
    RBQ_CUDA(cudaMemset(stp, 0, sizeof(DevStats) + 64));
    h->allocations.push_back(stp);
    h->d_stats = reinterpret_cast<DevStats*>(stp);
    for (auto& e : h->ev) cudaEventCreate(&e);
    *out = h;
    b->h = nullptr;
    delete b;
    return RBQ_OK;
}
