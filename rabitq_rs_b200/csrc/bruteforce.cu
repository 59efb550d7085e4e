// bruteforce.cu -- BruteForceRabitqIndex on the device (reference src/brute_force.rs): no clustering, every vector quantised
// against the zero centroid, exhaustive scan.  SURVEY.md section 8 row f-4.
//
//   train   :214-285  rotate + quantize_with_centroid(zero) per vector, input order = ids  -> build.cu's kernels
//   save    :305-385  "RBF1" v1: header | rotator bytes | per vector: packed sign code, packed ex-code, 8 x f32 | CRC-32
//   load    :395-523  same validation order and messages
//   search  :545-650  per vector: binary_dot and ex_dot as SEQUENTIAL scalar sums over the padded dimension (separate multiply
//                     and add), distance from the 1-bit or the extended factors, BinaryHeap of the k smallest, final sort
//
// The search kernel keeps the reference's float order exactly (one thread walks one vector's dimensions in index order;
// -fmad=false), so distances are bit-identical; one CTA per query, every warp takes 32 vectors at a time, each warp keeps a
// top-k list (registers for k <= 32) pruned by its own k-th distance, the 8 lists are merged at the end.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "scan_common.cuh"

using namespace rbq;

struct rbq_bf_index {
    size_t dim = 0, D = 0, n = 0;
    int metric = 0, rot_type = 1, ex_bits = 0, device = 0;
    std::vector<uint8_t> rot_bytes;
    DevIndex dev{};          // rotator fields only
    size_t bin_stride = 0, ex_stride = 0, ex_row = 0;  // packed bytes per vector; ex_row = ex_stride rounded up to 16 (device rows)
    uint8_t* d_bin = nullptr;    // n * bin_row (bin_row = D/8 rounded up to 16)
    uint8_t* d_ex = nullptr;     // n * ex_row
    float* d_meta = nullptr;     // n * 8: delta, vl, f_add, f_rescale, f_error, residual_norm, f_add_ex, f_rescale_ex
    size_t bin_row = 0;
    std::vector<void*> allocations;
    std::mutex mu;
};

namespace {
constexpr int kBfWarps = 8;

struct BfArgs {
    const uint8_t* bin;
    const uint8_t* ex;
    const float* meta;
    uint32_t n, bin_row, ex_row;
    int D, ex_bits, metric;
    const float* rot;  // nq * D rotated queries
    uint32_t top_k;
    const unsigned long long* filter;
    unsigned long long filter_nbits;
    unsigned long long* out_ids;
    float* out_scores;
    uint32_t* out_counts;
};

// the 16 codes of chunk c of a packed ex-code, in dimension order (decode_chunk returns them 4 per word: A = dims 0-3 ...)
template <int EXK>
__global__ void __launch_bounds__(kBfWarps * 32) bf_search_kernel(BfArgs a) {
    extern __shared__ __align__(16) unsigned char bf_smem[];
    float* rq = reinterpret_cast<float*>(bf_smem);                                       // D floats
    float* sd = rq + a.D;                                                                // kBfWarps * k distances
    unsigned long long* si = reinterpret_cast<unsigned long long*>(sd + ((kBfWarps * a.top_k + 1) & ~1u));  // kBfWarps * k ids
    __shared__ float s_sum;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, D = a.D, k = (int)a.top_k;
    const size_t q = blockIdx.x;
    for (int i = tid; i < D; i += kBfWarps * 32) rq[i] = a.rot[q * D + i];
    __syncthreads();
    if (tid == 0) {  // QueryPrecomputed::new: sequential iterator sum (src/brute_force.rs:84-96)
        float s = 0.0f;
        for (int i = 0; i < D; ++i) s = s + rq[i];
        s_sum = s;
    }
    __syncthreads();
    const float sum_q = s_sum;
    const float k1x = -0.5f * sum_q, cb = -((float)(1 << a.ex_bits) - 0.5f), kbx = cb * sum_q, bscale = (float)(1 << a.ex_bits);
    TopK tk;
    tk.init(sd + (size_t)warp * k, si + (size_t)warp * k, k);
    for (uint32_t v0 = (uint32_t)warp * 32u; v0 < a.n; v0 += kBfWarps * 32u) {
        const uint32_t v = v0 + (uint32_t)lane;
        bool live = v < a.n;
        if (live && a.filter != nullptr) live = (unsigned long long)v < a.filter_nbits && ((a.filter[v >> 6] >> (v & 63u)) & 1ull);
        float dist = INFINITY;
        if (live) {
            const uint8_t* brow = a.bin + (size_t)v * a.bin_row;
            float bdot = 0.0f;
            for (int b0 = 0; b0 < D / 8; b0 += 16) {  // 16 bytes = 128 dims per load
                const uint4 w = ldg128(brow + b0);
                const uint32_t W[4] = {w.x, w.y, w.z, w.w};
                const int nb = min(16, D / 8 - b0);
                for (int j = 0; j < nb; ++j) {
                    const uint32_t byte = (W[j >> 2] >> (8 * (j & 3))) & 0xffu;
#pragma unroll
                    for (int t = 0; t < 8; ++t) {  // MSB-first (src/simd.rs:141-150); bit * q is q or +-0, and s + (+-0) == s
                        const float pr = (float)((byte >> (7 - t)) & 1u) * rq[8 * (b0 + j) + t];
                        bdot = bdot + pr;
                    }
                }
            }
            const float* m = a.meta + (size_t)v * 8;
            float t0, t1;
            if (EXK == 0) {
                const float bterm = bdot + k1x;
                t0 = m[2] + 0.0f;
                t1 = m[3] * bterm;
            } else {
                const uint8_t* erow = a.ex + (size_t)v * a.ex_row;
                float edot = 0.0f;
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t X[4];
                    decode_chunk<EXK, false>(erow, c, a.ex_bits, X[0], X[1], X[2], X[3]);
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float pr = (float)((X[r >> 2] >> (8 * (r & 3))) & 0xffu) * rq[16 * c + r];
                        edot = edot + pr;
                    }
                }
                float tt = bscale * bdot;
                tt = tt + edot;
                tt = tt + kbx;
                t0 = m[6] + 0.0f;
                t1 = m[7] * tt;
            }
            dist = t0 + t1;
            if (!isfinite(dist)) live = false;
        }
        // the k smallest: candidates below the warp's current k-th distance, in lane (= id) order
        const float th = tk.theta();  // warp-collective: must not sit behind a per-lane short circuit
        unsigned mask = __ballot_sync(0xffffffffu, live && dist < th);
        while (mask) {
            const int sl = __ffs(mask) - 1;
            mask &= mask - 1;
            const float d_s = __shfl_sync(0xffffffffu, dist, sl);
            if (d_s < tk.theta()) tk.insert(d_s, (unsigned long long)(v0 + sl), lane);
        }
    }
    // merge the warps' lists: every warp stores its (sorted) list, warp 0 inserts the others' entries into its own
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
        const bool have = i < tk.cnt;
        const float dv = tk.dist_at(i);
        const unsigned long long iv = tk.id_at(i);
        sd[(size_t)warp * k + i] = have ? dv : INFINITY;
        si[(size_t)warp * k + i] = have ? iv : ~0ull;
    }
    __shared__ int s_cnt[kBfWarps];
    if (lane == 0) s_cnt[warp] = tk.cnt;
    __syncthreads();
    if (warp == 0) {
        // warp 0's own list stays in place (registers for k <= 32, its shared rows otherwise)
        for (int w = 1; w < kBfWarps; ++w)
            for (int i = 0; i < s_cnt[w]; ++i) {
                const float d_s = sd[(size_t)w * k + i];
                if (!(d_s < tk.theta())) break;  // the list is sorted: nothing further can enter
                tk.insert(d_s, si[(size_t)w * k + i], lane);
            }
        __syncwarp();
        const bool l2 = a.metric == RBQ_METRIC_L2;
        for (int i = lane; i < k; i += 32) {
            const bool have = i < tk.cnt;
            const float dv = tk.dist_at(i);
            a.out_ids[q * k + i] = have ? tk.id_at(i) : ~0ull;
            a.out_scores[q * k + i] = have ? (l2 ? dv : -dv) : 0.0f;
        }
        if (lane == 0) a.out_counts[q] = (uint32_t)tk.cnt;
    }
}

struct Guard {
    int prev = -1;
    explicit Guard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~Guard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
int keep(rbq_bf_index* h, T** out, size_t count) {
    void* d = nullptr;
    RBQ_CUDA(cudaMalloc(&d, std::max<size_t>(count * sizeof(T), 16)));
    h->allocations.push_back(d);
    RBQ_CUDA(cudaMemset(d, 0, std::max<size_t>(count * sizeof(T), 16)));
    *out = reinterpret_cast<T*>(d);
    return RBQ_OK;
}

int setup_geometry(rbq_bf_index* h) {
    if (h->D % 16 != 0) return fail(RBQ_INVALID_CONFIG, "padded_dim must be a multiple of 16");
    if (h->D > 2048) return fail(RBQ_INVALID_CONFIG, "padded_dim > 2048 is not supported");
    if (h->ex_bits > 8) return fail(RBQ_INVALID_CONFIG, "ex_bits > 8 is not supported");
    h->bin_stride = h->D / 8;
    h->ex_stride = h->ex_bits > 0 ? h->D * h->ex_bits / 8 : 0;
    h->bin_row = (h->bin_stride + 15) / 16 * 16;
    h->ex_row = (h->ex_stride + 15) / 16 * 16 + 16;  // + slack: the generic decoder reads one byte past a code
    DevIndex& d = h->dev;
    d.dim = (int)h->dim;
    d.D = (int)h->D;
    d.metric = h->metric;
    d.ex_bits = h->ex_bits;
    d.rot_type = h->rot_type;
    int lg = 0;
    while ((2u << lg) <= h->dim) ++lg;
    d.trunc = 1 << lg;
    d.fac = 1.0f / std::sqrt((float)d.trunc);
    d.ex_stride = (uint32_t)h->ex_stride;
    d.flip = nullptr;
    d.matrix_t = nullptr;
    int rc;
    if (h->rot_type == RBQ_ROTATOR_FHT_KAC) {
        uint8_t* f = nullptr;
        if ((rc = keep(h, &f, h->rot_bytes.size()))) return rc;
        RBQ_CUDA(cudaMemcpy(f, h->rot_bytes.data(), h->rot_bytes.size(), cudaMemcpyHostToDevice));
        d.flip = f;
    } else {
        const size_t D = h->D;
        std::vector<float> mt(D * D);
        const float* m = reinterpret_cast<const float*>(h->rot_bytes.data());
        for (size_t r = 0; r < D; ++r)
            for (size_t c = 0; c < D; ++c) mt[c * D + r] = m[r * D + c];
        float* f = nullptr;
        if ((rc = keep(h, &f, D * D))) return rc;
        RBQ_CUDA(cudaMemcpy(f, mt.data(), D * D * 4, cudaMemcpyHostToDevice));
        d.matrix_t = f;
    }
    return RBQ_OK;
}

// packed rows (file / builder layout, stride bytes each) -> device rows of row bytes (zero padded)
__global__ void bf_spread_rows_kernel(const uint8_t* __restrict__ src, size_t stride, uint8_t* __restrict__ dst, size_t row, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * row) return;
    const size_t v = i / row, b = i % row;
    dst[i] = b < stride ? src[v * stride + b] : (uint8_t)0;
}
}  // namespace

extern "C" {

void rbq_bf_free(rbq_bf_index* h) {
    if (!h) return;
    {
        Guard g(h->device);
        for (void* p : h->allocations) cudaFree(p);
    }
    delete h;
}

// BruteForceRabitqIndex::train (src/brute_force.rs:214-285); validation messages as the reference
int rbq_bf_train(const float* data, size_t n, size_t dim, int total_bits, int metric, int rotator_type, uint64_t seed, int faster_config,
                 const uint8_t* rotator_state, int device, rbq_bf_index** out) {
    if (!out) return fail(RBQ_INVALID_CONFIG, "null argument");
    *out = nullptr;
    if (n == 0 || !data) return fail(RBQ_INVALID_CONFIG, "training data must be non-empty");
    if (total_bits < 1 || total_bits > 16) return fail(RBQ_INVALID_CONFIG, "total_bits must be between 1 and 16");
    if (total_bits > 9) return fail(RBQ_INVALID_CONFIG, "total_bits above 9 (ex_bits > 8) is not supported");
    if (dim == 0) return fail(RBQ_INVALID_CONFIG, "input vectors must share the same dimension");
    if (metric != RBQ_METRIC_L2 && metric != RBQ_METRIC_INNER_PRODUCT) return fail(RBQ_INVALID_CONFIG, "unknown metric");
    if (rotator_type != RBQ_ROTATOR_MATRIX && rotator_type != RBQ_ROTATOR_FHT_KAC) return fail(RBQ_INVALID_CONFIG, "unknown rotator type");
    if (!faster_config && total_bits > 1)
        return fail(RBQ_INVALID_CONFIG, "the device brute-force builder implements RabitqConfig::faster only (use_faster_config = true)");
    rbq_bf_index* h = new rbq_bf_index();
    struct Fin {
        rbq_bf_index* h;
        bool ok = false;
        ~Fin() {
            if (!ok) rbq_bf_free(h);
        }
    } fin{h};
    h->device = device;
    Guard g(device);
    h->dim = dim;
    h->D = rotator_type == RBQ_ROTATOR_FHT_KAC ? (dim + 63) / 64 * 64 : dim;
    h->metric = metric;
    h->rot_type = rotator_type;
    h->ex_bits = total_bits - 1;
    h->n = n;
    const size_t D = h->D;
    if (rotator_type == RBQ_ROTATOR_FHT_KAC) {
        h->rot_bytes.resize(4 * D / 8);
        if (rotator_state) std::memcpy(h->rot_bytes.data(), rotator_state, h->rot_bytes.size());
        else {
            uint64_t st = seed;
            for (auto& b : h->rot_bytes) {
                st += 0x9e3779b97f4a7c15ULL;
                uint64_t z = st;
                z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
                z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
                b = (uint8_t)((z ^ (z >> 31)) >> 56);
            }
        }
    } else {
        if (!rotator_state) return fail(RBQ_INVALID_CONFIG, "a MatrixRotator brute-force index needs an explicit rotator_state (padded^2 f32)");
        h->rot_bytes.assign(rotator_state, rotator_state + D * D * 4);
    }
    int rc;
    if ((rc = setup_geometry(h))) return rc;
    const float t_const = h->ex_bits > 0 ? const_scaling_factor_host(D, h->ex_bits, seed) : -1.0f;
    if ((rc = keep(h, &h->d_bin, n * h->bin_row))) return rc;
    if ((rc = keep(h, &h->d_ex, n * h->ex_row + 16))) return rc;
    if ((rc = keep(h, &h->d_meta, n * 8))) return rc;
    // scratch: chunks of vectors through rotate_only + quantize (zero centroid, every vector in list 0)
    const size_t CH = std::min<size_t>(n, std::max<size_t>(4096, ((size_t)256 << 20) / (D * 4)));
    std::vector<void*> tmp;
    struct Free {
        std::vector<void*>& t;
        ~Free() {
            for (void* p : t) cudaFree(p);
        }
    } fr{tmp};
    auto talloc = [&](void** p, size_t bytes) -> int {
        RBQ_CUDA(cudaMalloc(p, std::max<size_t>(bytes, 16)));
        tmp.push_back(*p);
        return RBQ_OK;
    };
    float *d_in, *d_rot, *d_zero, *d_f[8];
    uint8_t *d_binp, *d_exp;
    uint32_t* d_list;
    if ((rc = talloc((void**)&d_in, CH * dim * 4))) return rc;
    if ((rc = talloc((void**)&d_rot, CH * D * 4))) return rc;
    if ((rc = talloc((void**)&d_zero, D * 4))) return rc;
    if ((rc = talloc((void**)&d_binp, CH * h->bin_stride))) return rc;
    if ((rc = talloc((void**)&d_exp, CH * h->ex_stride + 16))) return rc;
    if ((rc = talloc((void**)&d_list, CH * 4))) return rc;
    for (auto& p : d_f)
        if ((rc = talloc((void**)&p, CH * 4))) return rc;
    RBQ_CUDA(cudaMemset(d_zero, 0, D * 4));
    RBQ_CUDA(cudaMemset(d_list, 0, CH * 4));
    std::vector<float> hf(8 * CH), hm(8 * CH);
    for (size_t p0 = 0; p0 < n; p0 += CH) {
        const size_t m = std::min(CH, n - p0);
        RBQ_CUDA(cudaMemcpy(d_in, data + p0 * dim, m * dim * 4, cudaMemcpyHostToDevice));
        if ((rc = launch_rotate_only(h->dev, d_in, m, d_rot, nullptr))) return rc;
        BuildOut bo;
        bo.bin_rows = d_binp;
        bo.ex = d_exp;
        bo.delta = d_f[0];
        bo.vl = d_f[1];
        bo.f_add = d_f[2];
        bo.f_rescale = d_f[3];
        bo.f_error = d_f[4];
        bo.rnorm = d_f[5];
        bo.f_add_ex = d_f[6];
        bo.f_rescale_ex = d_f[7];
        if ((rc = launch_build_quantize(h->dev, d_rot, d_list, m, d_zero, t_const, nullptr, bo, nullptr))) return rc;
        const unsigned tb = 256;
        bf_spread_rows_kernel<<<(unsigned)((m * h->bin_row + tb - 1) / tb), tb>>>(d_binp, h->bin_stride, h->d_bin + p0 * h->bin_row, h->bin_row, m);
        if (h->ex_stride)
            bf_spread_rows_kernel<<<(unsigned)((m * h->ex_row + tb - 1) / tb), tb>>>(d_exp, h->ex_stride, h->d_ex + p0 * h->ex_row, h->ex_row, m);
        RBQ_CUDA(cudaGetLastError());
        for (int f = 0; f < 8; ++f) RBQ_CUDA(cudaMemcpy(hf.data() + (size_t)f * CH, d_f[f], m * 4, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < m; ++i)
            for (int f = 0; f < 8; ++f) hm[i * 8 + f] = hf[(size_t)f * CH + i];
        RBQ_CUDA(cudaMemcpy(h->d_meta + p0 * 8, hm.data(), m * 32, cudaMemcpyHostToDevice));
    }
    RBQ_CUDA(cudaDeviceSynchronize());
    fin.ok = true;
    *out = h;
    return RBQ_OK;
}

// BruteForceRabitqIndex::load_from_reader (src/brute_force.rs:395-523)
int rbq_bf_load_mem(const uint8_t* p, size_t n, int device, rbq_bf_index** out) {
    if (!p || !out) return fail(RBQ_INVALID_CONFIG, "null argument");
    *out = nullptr;
    const char* kEof = "failed to fill whole buffer";
    size_t off = 0;
    auto need = [&](size_t k) { return off + k <= n; };
    if (!need(4)) return fail(RBQ_IO, kEof);
    if (std::memcmp(p, "RBF1", 4) != 0) return fail(RBQ_INVALID_PERSISTENCE, "unrecognized file header");
    off = 4;
    auto rd = [&](void* v, size_t k) {
        if (!need(k)) return false;
        std::memcpy(v, p + off, k);
        off += k;
        return true;
    };
    uint32_t version, dim, D;
    uint8_t metric, rt, exb, tb;
    uint64_t cnt, rlen;
    if (!rd(&version, 4)) return fail(RBQ_IO, kEof);
    if (version != 1) return fail(RBQ_INVALID_PERSISTENCE, "unsupported index format version");
    if (!rd(&dim, 4)) return fail(RBQ_IO, kEof);
    if (dim == 0) return fail(RBQ_INVALID_PERSISTENCE, "dimension must be positive");
    if (!rd(&D, 4)) return fail(RBQ_IO, kEof);
    if (D < dim) return fail(RBQ_INVALID_PERSISTENCE, "padded_dim must be >= dim");
    if (!rd(&metric, 1)) return fail(RBQ_IO, kEof);
    if (metric > 1) return fail(RBQ_INVALID_PERSISTENCE, "unknown metric tag");
    if (!rd(&rt, 1)) return fail(RBQ_IO, kEof);
    if (rt > 1) return fail(RBQ_INVALID_PERSISTENCE, "unknown rotator type tag");
    if (!rd(&exb, 1)) return fail(RBQ_IO, kEof);
    if (exb > 16) return fail(RBQ_INVALID_PERSISTENCE, "ex_bits out of range");
    if (!rd(&tb, 1)) return fail(RBQ_IO, kEof);
    if (tb == 0 || tb > 16) return fail(RBQ_INVALID_PERSISTENCE, "total_bits out of range");
    if ((int)tb - 1 != (int)exb) return fail(RBQ_INVALID_PERSISTENCE, "total_bits does not match ex_bits");
    if (!rd(&cnt, 8) || !rd(&rlen, 8)) return fail(RBQ_IO, kEof);
    if (!need(rlen)) return fail(RBQ_IO, kEof);
    if (rt == 1 && rlen != 4 * (uint64_t)D / 8) return fail(RBQ_INVALID_PERSISTENCE, "FHT rotator flip bits length mismatch");
    if (rt == 0 && rlen != (uint64_t)D * D * 4) return fail(RBQ_INVALID_PERSISTENCE, "rotator matrix length mismatch");
    rbq_bf_index* h = new rbq_bf_index();
    struct Fin {
        rbq_bf_index* h;
        bool ok = false;
        ~Fin() {
            if (!ok) rbq_bf_free(h);
        }
    } fin{h};
    h->device = device;
    Guard g(device);
    h->dim = dim;
    h->D = D;
    h->metric = metric;
    h->rot_type = rt;
    h->ex_bits = exb;
    h->rot_bytes.assign(p + off, p + off + rlen);
    off += rlen;
    int rc;
    if ((rc = setup_geometry(h))) return rc;
    const size_t bsz = ((size_t)D + 7) / 8, esz = exb > 0 ? ((size_t)D * exb + 7) / 8 : 0, rec = bsz + esz + 32;
    if (cnt > (n - off) / rec + 1 || !need(cnt * rec)) return fail(RBQ_IO, kEof);
    h->n = cnt;
    const size_t crc_end = off + cnt * rec;
    uint32_t stored;
    if (crc_end + 4 > n) return fail(RBQ_IO, kEof);
    std::memcpy(&stored, p + crc_end, 4);
    if (crc32_ieee(0, p + 8, crc_end - 8) != stored) return fail(RBQ_INVALID_PERSISTENCE, "checksum mismatch");
    if ((rc = keep(h, &h->d_bin, cnt * h->bin_row))) return rc;
    if ((rc = keep(h, &h->d_ex, cnt * h->ex_row + 16))) return rc;
    if ((rc = keep(h, &h->d_meta, cnt * 8))) return rc;
    std::vector<uint8_t> hb(cnt * h->bin_row, 0), he(cnt * h->ex_row + 16, 0);
    std::vector<float> hm(cnt * 8);
    for (uint64_t i = 0; i < cnt; ++i) {
        const uint8_t* r = p + off + i * rec;
        std::memcpy(&hb[i * h->bin_row], r, bsz);
        std::memcpy(&he[i * h->ex_row], r + bsz, esz);
        std::memcpy(&hm[i * 8], r + bsz + esz, 32);
    }
    if (cnt) {
        RBQ_CUDA(cudaMemcpy(h->d_bin, hb.data(), hb.size(), cudaMemcpyHostToDevice));
        RBQ_CUDA(cudaMemcpy(h->d_ex, he.data(), he.size(), cudaMemcpyHostToDevice));
        RBQ_CUDA(cudaMemcpy(h->d_meta, hm.data(), hm.size() * 4, cudaMemcpyHostToDevice));
    }
    fin.ok = true;
    *out = h;
    return RBQ_OK;
}

int rbq_bf_load(const char* path, int device, rbq_bf_index** out) {
    if (!path || !out) return fail(RBQ_INVALID_CONFIG, "null argument");
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(RBQ_IO, std::string(strerror(errno)) + ": " + path);
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        close(fd);
        return fail(RBQ_IO, std::string(strerror(errno)) + ": " + path);
    }
    const size_t len = (size_t)sb.st_size;
    void* map = len ? mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (len && map == MAP_FAILED) return fail(RBQ_IO, std::string("mmap failed: ") + path);
    static const uint8_t empty = 0;
    const int rc = rbq_bf_load_mem(len ? (const uint8_t*)map : &empty, len, device, out);
    if (len) munmap(map, len);
    return rc;
}

// BruteForceRabitqIndex::save_to_writer (src/brute_force.rs:305-385); out == NULL queries the size
int rbq_bf_save_mem(rbq_bf_index* h, uint8_t* out, size_t cap, size_t* written) {
    if (!h || !written) return fail(RBQ_INVALID_CONFIG, "null argument");
    Guard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    const size_t bsz = h->bin_stride, esz = h->ex_stride, rec = bsz + esz + 32;
    const size_t total = 4 + 4 + 4 + 4 + 4 + 8 + 8 + h->rot_bytes.size() + h->n * rec + 4;
    *written = total;
    if (!out) return RBQ_OK;
    if (cap < total) return fail(RBQ_IO, "output buffer too small");
    std::vector<uint8_t> hb(h->n * h->bin_row), he(h->n * h->ex_row + 16);
    std::vector<float> hm(h->n * 8);
    if (h->n) {
        RBQ_CUDA(cudaMemcpy(hb.data(), h->d_bin, hb.size(), cudaMemcpyDeviceToHost));
        RBQ_CUDA(cudaMemcpy(he.data(), h->d_ex, he.size(), cudaMemcpyDeviceToHost));
        RBQ_CUDA(cudaMemcpy(hm.data(), h->d_meta, hm.size() * 4, cudaMemcpyDeviceToHost));
    }
    size_t off = 0;
    auto put = [&](const void* p, size_t k) {
        std::memcpy(out + off, p, k);
        off += k;
    };
    const uint32_t version = 1, dim = (uint32_t)h->dim, D = (uint32_t)h->D;
    const uint8_t tags[4] = {(uint8_t)h->metric, (uint8_t)h->rot_type, (uint8_t)h->ex_bits, (uint8_t)(h->ex_bits + 1)};
    const uint64_t cnt = h->n, rlen = h->rot_bytes.size();
    put("RBF1", 4);
    put(&version, 4);
    put(&dim, 4);
    put(&D, 4);
    put(tags, 4);
    put(&cnt, 8);
    put(&rlen, 8);
    put(h->rot_bytes.data(), rlen);
    for (size_t i = 0; i < h->n; ++i) {
        put(&hb[i * h->bin_row], bsz);
        put(&he[i * h->ex_row], esz);
        put(&hm[i * 8], 32);
    }
    const uint32_t crc = crc32_ieee(0, out + 8, off - 8);
    put(&crc, 4);
    return RBQ_OK;
}

int rbq_bf_save(rbq_bf_index* h, const char* path) {
    if (!h || !path) return fail(RBQ_INVALID_CONFIG, "null argument");
    size_t n = 0;
    int rc = rbq_bf_save_mem(h, nullptr, 0, &n);
    if (rc) return rc;
    std::vector<uint8_t> buf(n);
    if ((rc = rbq_bf_save_mem(h, buf.data(), n, &n))) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(RBQ_IO, std::string(strerror(errno)) + ": " + path);
    const size_t w = fwrite(buf.data(), 1, n, f);
    if (fclose(f) != 0 || w != n) return fail(RBQ_IO, std::string("short write: ") + path);
    return RBQ_OK;
}

size_t rbq_bf_len(const rbq_bf_index* h) { return h ? h->n : 0; }
size_t rbq_bf_dim(const rbq_bf_index* h) { return h ? h->dim : 0; }
size_t rbq_bf_padded_dim(const rbq_bf_index* h) { return h ? h->D : 0; }

// BruteForceRabitqIndex::search / search_filtered for a batch (src/brute_force.rs:526-650).  Host buffers.
int rbq_bf_search_batch(rbq_bf_index* h, const float* queries, size_t nq, size_t dim, size_t top_k, const uint64_t* filter_bits,
                        size_t filter_nbits, uint64_t* ids, float* scores, uint32_t* counts) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    if (h->n == 0) return fail(RBQ_EMPTY_INDEX, "index is empty; call `train` first");
    if (dim != h->dim) return fail(RBQ_DIMENSION_MISMATCH, "expected " + std::to_string(h->dim) + ", got " + std::to_string(dim));
    if (nq && (!queries || !counts || (top_k && (!ids || !scores)))) return fail(RBQ_INVALID_CONFIG, "null buffer");
    if (top_k > 1024) return fail(RBQ_INVALID_CONFIG, "top_k exceeds the device limit (1024)");
    if (nq == 0) return RBQ_OK;
    if (top_k == 0) {
        std::memset(counts, 0, nq * 4);
        return RBQ_OK;
    }
    Guard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    const size_t D = h->D, fwords = filter_bits ? std::max<size_t>((filter_nbits + 63) / 64, 1) : 0;
    std::vector<void*> tmp;
    struct Free {
        std::vector<void*>& t;
        ~Free() {
            for (void* p : t) cudaFree(p);
        }
    } fr{tmp};
    auto talloc = [&](void** p, size_t bytes) -> int {
        RBQ_CUDA(cudaMalloc(p, std::max<size_t>(bytes, 16)));
        tmp.push_back(*p);
        return RBQ_OK;
    };
    float *d_q, *d_rot, *d_sc;
    uint64_t *d_ids, *d_f = nullptr;
    uint32_t* d_cn;
    int rc;
    if ((rc = talloc((void**)&d_q, nq * dim * 4))) return rc;
    if ((rc = talloc((void**)&d_rot, nq * D * 4))) return rc;
    if ((rc = talloc((void**)&d_ids, nq * top_k * 8))) return rc;
    if ((rc = talloc((void**)&d_sc, nq * top_k * 4))) return rc;
    if ((rc = talloc((void**)&d_cn, nq * 4))) return rc;
    if (fwords) {
        if ((rc = talloc((void**)&d_f, fwords * 8))) return rc;
        RBQ_CUDA(cudaMemset(d_f, 0, fwords * 8));
        if (filter_nbits) RBQ_CUDA(cudaMemcpy(d_f, filter_bits, (filter_nbits + 63) / 64 * 8, cudaMemcpyHostToDevice));
    }
    RBQ_CUDA(cudaMemcpy(d_q, queries, nq * dim * 4, cudaMemcpyHostToDevice));
    if ((rc = launch_rotate_only(h->dev, d_q, nq, d_rot, nullptr))) return rc;
    BfArgs a;
    a.bin = h->d_bin;
    a.ex = h->d_ex;
    a.meta = h->d_meta;
    a.n = (uint32_t)h->n;
    a.bin_row = (uint32_t)h->bin_row;
    a.ex_row = (uint32_t)h->ex_row;
    a.D = (int)D;
    a.ex_bits = h->ex_bits;
    a.metric = h->metric;
    a.rot = d_rot;
    a.top_k = (uint32_t)top_k;
    a.filter = reinterpret_cast<const unsigned long long*>(d_f);
    a.filter_nbits = filter_nbits;
    a.out_ids = reinterpret_cast<unsigned long long*>(d_ids);
    a.out_scores = d_sc;
    a.out_counts = d_cn;
    const size_t smem = D * 4 + ((kBfWarps * top_k + 1) & ~(size_t)1) * 4 + kBfWarps * top_k * 8 + 16;
#define RBQ_BF_LAUNCH(EXK)                                                                                            \
    do {                                                                                                              \
        RBQ_CUDA(cudaFuncSetAttribute(bf_search_kernel<EXK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        bf_search_kernel<EXK><<<(unsigned)nq, kBfWarps * 32, smem>>>(a);                                              \
    } while (0)
    if (h->ex_bits == 0) RBQ_BF_LAUNCH(0);
    else if (h->ex_bits == 2) RBQ_BF_LAUNCH(2);
    else if (h->ex_bits == 6) RBQ_BF_LAUNCH(6);
    else RBQ_BF_LAUNCH(1);
#undef RBQ_BF_LAUNCH
    RBQ_CUDA(cudaGetLastError());
    RBQ_CUDA(cudaMemcpy(ids, d_ids, nq * top_k * 8, cudaMemcpyDeviceToHost));
    RBQ_CUDA(cudaMemcpy(scores, d_sc, nq * top_k * 4, cudaMemcpyDeviceToHost));
    RBQ_CUDA(cudaMemcpy(counts, d_cn, nq * 4, cudaMemcpyDeviceToHost));
    return RBQ_OK;
}

}  // extern "C"
