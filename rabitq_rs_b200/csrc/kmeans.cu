// kmeans.cu -- k-means for `train` on the device (reference src/kmeans.rs: run_kmeans_flat :71-186, Lloyd iterations
// :291-326, assign_points_for_update :439-547, update_centroids :564-602, assign_full_dataset :604-643).
//
// Same pipeline as the reference: a training subset of at most max_points_per_centroid * k points, random (Forgy)
// initialisation from it, niter Lloyd iterations, then the assignment of the full data set.  The assignment -- the
// reference's sgemm + arg-min over |x|^2 + |c|^2 - 2 x.c clamped at 0, first minimum wins -- runs on the engine's tcgen05
// GEMM with the arg-min fused into its epilogue (coarse_tc.cu, kGemmArgmin): operands are bf16 hi/lo splits, so the scores
// are fp32-class and no n x k matrix exists.  The centroid update is a deterministic segmented mean: points are sorted by
// cluster (stable radix sort) and every cluster's members are summed in index order, so a given (data, seed) always yields
// the same centroids (the reference's rayon fold is order-dependent).  Empty clusters are re-seeded with the points farthest
// from their centroids, like the reference's candidate pool.  The random streams are splitmix64, not the reference's ChaCha12:
// k-means results are never part of the parity contract (any clustering yields a valid index).
#include <algorithm>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>

#include "rbq_internal.h"

namespace rbq {
namespace {

struct Tmp {
    std::vector<void*> ptrs;
    ~Tmp() {
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

inline uint64_t sm64(uint64_t& s) {
    uint64_t z = (s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

// rows idx[i] of src (row-major, dim floats) -> dst row i
__global__ void gather_f32_rows_kernel(const float* __restrict__ src, const uint64_t* __restrict__ idx, int dim, float* __restrict__ dst) {
    const float* s = src + idx[blockIdx.x] * (size_t)dim;
    float* d = dst + (size_t)blockIdx.x * dim;
    for (int i = threadIdx.x; i < dim; i += blockDim.x) d[i] = s[i];
}
__global__ void fill_u64_kernel(unsigned long long* p, size_t n, unsigned long long v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// best[i] = (order key of the clamped distance) << 32 | cluster  ->  assignment, and the sort inputs of the update step
__global__ void unpack_best_kernel(const unsigned long long* __restrict__ best, size_t n, uint32_t* __restrict__ assign,
                                   uint32_t* __restrict__ far_key, uint32_t* __restrict__ idx) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long b = best[i];
    assign[i] = (uint32_t)(b & 0xffffffffull);
    if (far_key) far_key[i] = ~(uint32_t)(b >> 32);  // ascending sort of ~key = farthest first
    if (idx) idx[i] = (uint32_t)i;
}
// sorted (cluster, point) pairs -> start of every cluster's run (CSR offsets, k + 1 entries)
__global__ void run_offsets_kernel(const uint32_t* __restrict__ sorted_cluster, uint32_t n, uint32_t k, uint32_t* __restrict__ off) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > k) return;
    uint32_t lo = 0, hi = n;  // first position with cluster >= c
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sorted_cluster[mid] < c) lo = mid + 1;
        else hi = mid;
    }
    off[c] = lo;
}
// exclusive scan of the empty-cluster flags (one CTA): empty_rank[c] = number of empty clusters before c
__global__ void __launch_bounds__(1024) empty_rank_kernel(const uint32_t* __restrict__ off, uint32_t k, uint32_t* __restrict__ empty_rank) {
    __shared__ uint32_t s_part[1024];
    const uint32_t t = threadIdx.x, per = (k + 1023u) / 1024u, c0 = min(t * per, k), c1 = min(c0 + per, k);
    uint32_t cnt = 0;
    for (uint32_t c = c0; c < c1; ++c) cnt += off[c + 1] == off[c];
    s_part[t] = cnt;
    __syncthreads();
    if (t == 0) {
        uint32_t run = 0;
        for (int i = 0; i < 1024; ++i) {
            const uint32_t v = s_part[i];
            s_part[i] = run;
            run += v;
        }
    }
    __syncthreads();
    uint32_t run = s_part[t];
    for (uint32_t c = c0; c < c1; ++c) {
        empty_rank[c] = run;
        run += off[c + 1] == off[c];
    }
}
// one CTA per cluster: mean of its members, summed in ascending point order (deterministic); an empty cluster takes the
// empty_rank-th farthest point of the training set (reference src/kmeans.rs:564-602)
__global__ void __launch_bounds__(128) update_centroids_kernel(const float* __restrict__ data, int dim, const uint32_t* __restrict__ sorted_point,
                                                              const uint32_t* __restrict__ off, const uint32_t* __restrict__ empty_rank,
                                                              const uint32_t* __restrict__ far_point, uint32_t n, float* __restrict__ cents) {
    const uint32_t c = blockIdx.x;
    const uint32_t b = off[c], e = off[c + 1];
    float* out = cents + (size_t)c * dim;
    if (b == e) {
        const uint32_t src = far_point[min(empty_rank[c], n - 1)];
        for (int d = threadIdx.x; d < dim; d += blockDim.x) out[d] = data[(size_t)src * dim + d];
        return;
    }
    const float inv = 1.0f / (float)(e - b);
    for (int d = threadIdx.x; d < dim; d += blockDim.x) {
        float s = 0.0f;
        for (uint32_t i = b; i < e; ++i) s = s + data[(size_t)sorted_point[i] * dim + d];
        out[d] = s * inv;
    }
}

// arg-min assignment of `rows` points (device, row-major dim floats) against k centroids whose split operand is ready
int assign_chunked(const float* d_x, size_t rows, int dim, int Dp, const void* d_csplit, const float* d_cn2, size_t k, void* d_xsplit, float* d_xn2,
                   unsigned long long* d_best, size_t chunk, cudaStream_t st) {
    int rc;
    for (size_t r0 = 0; r0 < rows; r0 += chunk) {
        const size_t m = std::min(chunk, rows - r0);
        if ((rc = launch_split_bf16_pad(d_x + r0 * dim, m, dim, Dp, 0, d_xsplit, d_xn2, st))) return rc;
        fill_u64_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(d_best + r0, m, ~0ull);
        GemmEpi e;
        e.nq = (int)m;
        e.ncols = (int)k;
        e.metric = RBQ_METRIC_L2;
        e.qn2 = d_xn2;
        e.cn2 = d_cn2;
        e.best = d_best + r0;
        if ((rc = launch_coarse_gemm(kGemmArgmin, d_xsplit, m, d_csplit, k, Dp, 3, e, st))) return rc;
    }
    RBQ_CUDA(cudaGetLastError());
    return RBQ_OK;
}

}  // namespace
}  // namespace rbq

using namespace rbq;

extern "C" int rbq_kmeans_assign_device(const float* d_data, size_t n, size_t dim, const float* d_centroids, size_t k, uint32_t* d_assign,
                                        int device, void* stream) {
    if (!d_data || !d_centroids || !d_assign) return fail(RBQ_INVALID_CONFIG, "null argument");
    if (n == 0) return RBQ_OK;
    if (k == 0 || dim == 0) return fail(RBQ_INVALID_CONFIG, "centroids must be non-empty");
    if (n > 0xffffffffull) return fail(RBQ_INVALID_CONFIG, "assign at most 2^32 points per call");
    int prev = 0;
    cudaGetDevice(&prev);
    RBQ_CUDA(cudaSetDevice(device));
    struct Restore {
        int d;
        ~Restore() { cudaSetDevice(d); }
    } restore{prev};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int Dp = (int)((dim + 7) / 8 * 8);
    const size_t chunk = std::min<size_t>(n, std::max<size_t>(4096, ((size_t)1 << 30) / ((size_t)Dp * 6)));
    Tmp t;
    void *d_cs = nullptr, *d_xs = nullptr;
    float *d_cn2 = nullptr, *d_xn2 = nullptr;
    unsigned long long* d_best = nullptr;
    int rc;
    if ((rc = t.alloc((uint16_t**)&d_cs, k * 3 * (size_t)Dp))) return rc;
    if ((rc = t.alloc(&d_cn2, k))) return rc;
    if ((rc = t.alloc((uint16_t**)&d_xs, chunk * 3 * (size_t)Dp))) return rc;
    if ((rc = t.alloc(&d_xn2, chunk))) return rc;
    if ((rc = t.alloc(&d_best, n))) return rc;
    if ((rc = launch_split_bf16_pad(d_centroids, k, (int)dim, Dp, 1, d_cs, d_cn2, st))) return rc;
    if ((rc = assign_chunked(d_data, n, (int)dim, Dp, d_cs, d_cn2, k, d_xs, d_xn2, d_best, chunk, st))) return rc;
    unpack_best_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_best, n, d_assign, nullptr, nullptr);
    RBQ_CUDA(cudaGetLastError());
    RBQ_CUDA(cudaStreamSynchronize(st));  // the temporaries die with this frame
    return RBQ_OK;
}

extern "C" int rbq_kmeans_device(const float* d_data, size_t n, size_t dim, size_t k, int niter, uint64_t seed, size_t max_points_per_centroid,
                                 float* d_centroids, int device, void* stream) {
    if (!d_data || !d_centroids) return fail(RBQ_INVALID_CONFIG, "null argument");
    if (n == 0) return fail(RBQ_INVALID_CONFIG, "training data must be non-empty");
    if (k == 0 || dim == 0) return fail(RBQ_INVALID_CONFIG, "k must be positive");
    if (k > n) return fail(RBQ_INVALID_CONFIG, "nlist cannot exceed number of vectors");
    if (niter < 0) return fail(RBQ_INVALID_CONFIG, "niter must be non-negative");
    if (max_points_per_centroid == 0) max_points_per_centroid = 256;  // DEFAULT_MAX_POINTS_PER_CENTROID (src/kmeans.rs:10)
    int prev = 0;
    cudaGetDevice(&prev);
    RBQ_CUDA(cudaSetDevice(device));
    struct Restore {
        int d;
        ~Restore() { cudaSetDevice(d); }
    } restore{prev};
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Tmp t;
    int rc;
    // training subset (src/kmeans.rs:210-227): one point per stride, at a seeded offset inside the stride
    size_t nt = std::min<size_t>(n, k * max_points_per_centroid);
    if (nt > 0xfffffff0ull) return fail(RBQ_INVALID_CONFIG, "training subset too large");
    const float* d_train = d_data;
    uint64_t rng = seed ^ 0x5bd1e995u;
    if (nt < n) {
        std::vector<uint64_t> idx(nt);
        for (size_t j = 0; j < nt; ++j) {
            const size_t lo = (size_t)((unsigned __int128)j * n / nt), hi = (size_t)((unsigned __int128)(j + 1) * n / nt);
            idx[j] = lo + (size_t)(sm64(rng) % std::max<size_t>(hi - lo, 1));
        }
        uint64_t* d_idx = nullptr;
        float* d_sub = nullptr;
        if ((rc = t.alloc(&d_idx, nt))) return rc;
        if ((rc = t.alloc(&d_sub, nt * dim))) return rc;
        RBQ_CUDA(cudaMemcpyAsync(d_idx, idx.data(), nt * 8, cudaMemcpyHostToDevice, st));
        gather_f32_rows_kernel<<<(unsigned)nt, 128, 0, st>>>(d_data, d_idx, (int)dim, d_sub);
        RBQ_CUDA(cudaStreamSynchronize(st));
        d_train = d_sub;
    }
    // Forgy initialisation: k distinct training points, one per stride of the subset
    {
        std::vector<uint64_t> idx(k);
        for (size_t j = 0; j < k; ++j) {
            const size_t lo = (size_t)((unsigned __int128)j * nt / k), hi = (size_t)((unsigned __int128)(j + 1) * nt / k);
            idx[j] = lo + (size_t)(sm64(rng) % std::max<size_t>(hi - lo, 1));
        }
        uint64_t* d_idx = nullptr;
        if ((rc = t.alloc(&d_idx, k))) return rc;
        RBQ_CUDA(cudaMemcpyAsync(d_idx, idx.data(), k * 8, cudaMemcpyHostToDevice, st));
        gather_f32_rows_kernel<<<(unsigned)k, 128, 0, st>>>(d_train, d_idx, (int)dim, d_centroids);
        RBQ_CUDA(cudaStreamSynchronize(st));
    }
    if (niter == 0) return RBQ_OK;
    const int Dp = (int)((dim + 7) / 8 * 8);
    const size_t chunk = std::min<size_t>(nt, std::max<size_t>(4096, ((size_t)1 << 30) / ((size_t)Dp * 6)));
    void *d_cs = nullptr, *d_xs = nullptr;
    float *d_cn2 = nullptr, *d_xn2 = nullptr;
    unsigned long long* d_best = nullptr;
    uint32_t *d_assign = nullptr, *d_far = nullptr, *d_idx = nullptr, *d_sa = nullptr, *d_sp = nullptr, *d_sf = nullptr, *d_fp = nullptr, *d_off = nullptr,
             *d_er = nullptr;
    if ((rc = t.alloc((uint16_t**)&d_cs, k * 3 * (size_t)Dp))) return rc;
    if ((rc = t.alloc(&d_cn2, k))) return rc;
    if ((rc = t.alloc((uint16_t**)&d_xs, chunk * 3 * (size_t)Dp))) return rc;
    if ((rc = t.alloc(&d_xn2, chunk))) return rc;
    if ((rc = t.alloc(&d_best, nt))) return rc;
    if ((rc = t.alloc(&d_assign, nt))) return rc;
    if ((rc = t.alloc(&d_far, nt))) return rc;
    if ((rc = t.alloc(&d_idx, nt))) return rc;
    if ((rc = t.alloc(&d_sa, nt))) return rc;
    if ((rc = t.alloc(&d_sp, nt))) return rc;
    if ((rc = t.alloc(&d_sf, nt))) return rc;
    if ((rc = t.alloc(&d_fp, nt))) return rc;
    if ((rc = t.alloc(&d_off, k + 1))) return rc;
    if ((rc = t.alloc(&d_er, k))) return rc;
    size_t sort_bytes = 0;
    RBQ_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, d_assign, d_sa, d_idx, d_sp, (int)nt, 0, 32, st));
    char* d_sort = nullptr;
    if ((rc = t.alloc(&d_sort, sort_bytes + 16))) return rc;
    int kbits = 1;
    while (((size_t)1 << kbits) < k) ++kbits;
    for (int it = 0; it < niter; ++it) {
        if ((rc = launch_split_bf16_pad(d_centroids, k, (int)dim, Dp, 1, d_cs, d_cn2, st))) return rc;
        if ((rc = assign_chunked(d_train, nt, (int)dim, Dp, d_cs, d_cn2, k, d_xs, d_xn2, d_best, chunk, st))) return rc;
        unpack_best_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(d_best, nt, d_assign, d_far, d_idx);
        size_t sb = sort_bytes;
        // members of every cluster in ascending point order (LSD radix sort is stable), and the points farthest first
        RBQ_CUDA(cub::DeviceRadixSort::SortPairs(d_sort, sb, d_assign, d_sa, d_idx, d_sp, (int)nt, 0, kbits, st));
        sb = sort_bytes;
        RBQ_CUDA(cub::DeviceRadixSort::SortPairs(d_sort, sb, d_far, d_sf, d_idx, d_fp, (int)nt, 0, 32, st));
        run_offsets_kernel<<<(unsigned)((k + 1 + 255) / 256), 256, 0, st>>>(d_sa, (uint32_t)nt, (uint32_t)k, d_off);
        empty_rank_kernel<<<1, 1024, 0, st>>>(d_off, (uint32_t)k, d_er);
        update_centroids_kernel<<<(unsigned)k, 128, 0, st>>>(d_train, (int)dim, d_sp, d_off, d_er, d_fp, (uint32_t)nt, d_centroids);
        RBQ_CUDA(cudaGetLastError());
    }
    RBQ_CUDA(cudaStreamSynchronize(st));
    return RBQ_OK;
}
