// api.cu -- C entry points of librbq.so: handle lifetime, upload, the batched search pipeline.
// Interfaces replaced: IvfRabitqIndex::{load_from_path, load_from_reader, save_to_path, search,
// search_filtered, batch_search, len, cluster_count} (reference src/ivf.rs:1218-1230, 1310-1752).
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <ctime>
#include <cstdlib>
#include <cstring>

#include "nccl_loader.h"
#include "rbq_internal.h"

using namespace rbq;

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

template <class T>
int upload(rbq_index* h, const T* src, size_t count, const T** dst, size_t pad_bytes = 0) {
    void* d = nullptr;
    size_t bytes = count * sizeof(T);
    RBQ_CUDA(cudaMalloc(&d, std::max<size_t>(bytes + pad_bytes, 16)));
    h->allocations.push_back(d);
    if (bytes) RBQ_CUDA(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
    if (pad_bytes) RBQ_CUDA(cudaMemset((char*)d + bytes, 0, pad_bytes));
    *dst = reinterpret_cast<const T*>(d);
    return RBQ_OK;
}

int validate_geometry(const HostIndex& hi) {
    if (hi.D % 16 != 0)
        return fail(RBQ_INVALID_CONFIG, "padded_dim must be a multiple of 16 (FastScan requirement, reference src/simd.rs:978-981)");
    if (hi.D > 2048)
        return fail(RBQ_INVALID_CONFIG, "padded_dim > 2048 (high-accuracy LUT path) is not supported");
    if (hi.ex_bits > 8) return fail(RBQ_INVALID_CONFIG, "ex_bits > 8 is not supported");
    return RBQ_OK;
}

void fill_geometry(rbq_index* h) {
    const HostIndex& hi = h->host;
    DevIndex& d = h->dev;
    d.dim = (int)hi.dim;
    d.D = (int)hi.D;
    d.metric = hi.metric;
    d.ex_bits = hi.ex_bits;
    d.rot_type = hi.rot_type;
    int lg = 0;
    while ((2u << lg) <= hi.dim) ++lg;  // floor(log2(dim)), reference src/rotation.rs:263-266
    d.trunc = 1 << lg;
    d.fac = 1.0f / std::sqrt((float)d.trunc);
    d.nlist = (uint32_t)hi.nlist;
    d.block_stride = (uint32_t)hi.block_stride();
    d.max_list_n = 0;
    for (uint32_t n : hi.list_n) d.max_list_n = std::max(d.max_list_n, n);
    d.ex_stride = (uint32_t)hi.ex_stride();
}

// Move the host image to the device; bulk host arrays are released afterwards.
int upload_index(rbq_index* h) {
    HostIndex& hi = h->host;
    int rc = validate_geometry(hi);
    if (rc) return rc;
    fill_geometry(h);
    DevIndex& d = h->dev;
    d.flip = nullptr;
    d.matrix_t = nullptr;
    if (hi.rot_type == RBQ_ROTATOR_FHT_KAC) {
        if ((rc = upload(h, hi.rot_bytes.data(), hi.rot_bytes.size(), &d.flip))) return rc;
    } else {
        size_t D = hi.D;
        std::vector<float> mt(D * D);
        const float* m = reinterpret_cast<const float*>(hi.rot_bytes.data());
        for (size_t r = 0; r < D; ++r)
            for (size_t k = 0; k < D; ++k) mt[k * D + r] = m[r * D + k];
        if ((rc = upload(h, mt.data(), mt.size(), &d.matrix_t))) return rc;
    }
    if ((rc = upload(h, hi.centroids.data(), hi.centroids.size(), &d.centroids))) return rc;
    if ((rc = prepare_coarse_tc(h))) return rc;
    if ((rc = prepare_coarse_sample(h))) return rc;
    if ((rc = upload(h, hi.list_n.data(), hi.list_n.size(), &d.list_n))) return rc;
    d.list_owner = nullptr;
    d.shard_rank = hi.shard_rank;
    if (hi.shard_count > 1 && (rc = upload(h, hi.list_owner.data(), hi.list_owner.size(), &d.list_owner))) return rc;
    if ((rc = upload(h, hi.blk_off.data(), hi.blk_off.size(), &d.blk_off))) return rc;
    if ((rc = upload(h, hi.vec_off.data(), hi.vec_off.size(), &d.vec_off))) return rc;
    if ((rc = upload(h, hi.blocks.data(), hi.blocks.size(), &d.blocks))) return rc;
    if ((rc = upload(h, hi.ids.data(), hi.ids.size(), &d.ids))) return rc;
    if ((rc = upload(h, hi.ex.data(), hi.ex.size(), &d.ex, 16))) return rc;
    if ((rc = prepare_ex_lanes(h))) return rc;
    if ((rc = upload(h, hi.f_add_ex.data(), hi.f_add_ex.size(), &d.f_add_ex))) return rc;
    if ((rc = upload(h, hi.f_rescale_ex.data(), hi.f_rescale_ex.size(), &d.f_rescale_ex))) return rc;
    std::vector<uint8_t>().swap(hi.blocks);
    std::vector<uint64_t>().swap(hi.ids);
    std::vector<uint8_t>().swap(hi.ex);
    std::vector<float>().swap(hi.f_add_ex);
    std::vector<float>().swap(hi.f_rescale_ex);
    void* st = nullptr;
    RBQ_CUDA(cudaMalloc(&st, sizeof(DevStats) + 64));
    RBQ_CUDA(cudaMemset(st, 0, sizeof(DevStats) + 64));
    h->allocations.push_back(st);
    h->d_stats = reinterpret_cast<DevStats*>(st);
    return RBQ_OK;
}

}  // namespace

// Tensor-core coarse stage operands: bf16 split of the (rotated) centroids, |c|^2 and max |c|.
int rbq::prepare_coarse_tc(rbq_index* h) {
    DevIndex& d = h->dev;
    const size_t nl = d.nlist, D = d.D;
    void* sp = nullptr;
    float* n2 = nullptr;
    RBQ_CUDA(cudaMalloc(&sp, std::max<size_t>(nl * 3 * D * 2, 16)));
    h->allocations.push_back(sp);
    RBQ_CUDA(cudaMalloc(&n2, std::max<size_t>(nl * 4, 16)));
    h->allocations.push_back(n2);
    int rc = launch_split_bf16(d.centroids, nl, (int)D, 1, sp, n2, nullptr);
    if (rc) return rc;
    d.cent_q4 = nullptr;
    if (D % 32 == 0 && nl > 0) {  // lane-grouped copy for the exact re-score of the probe selection (coarse.cu)
        float* q4 = nullptr;
        RBQ_CUDA(cudaMalloc(&q4, nl * D * 4));
        h->allocations.push_back(q4);
        if ((rc = launch_centroid_q4(d.centroids, nl, (int)D, q4, nullptr))) return rc;
        d.cent_q4 = q4;
    }
    RBQ_CUDA(cudaDeviceSynchronize());
    double mx = 0.0;
    for (size_t c = 0; c < nl; ++c) {
        double s = 0.0;
        for (size_t k = 0; k < D; ++k) s += (double)h->host.centroids[c * D + k] * (double)h->host.centroids[c * D + k];
        mx = std::max(mx, s);
    }
    d.cent_split = sp;
    d.cent_n2 = n2;
    d.cmax_norm = (float)(std::sqrt(mx) * (1.0 + 1e-6));
    return RBQ_OK;
}

// Centroid sample of the coarse filter mode: samp_n rows of cent_split, one per stride with a hashed offset inside the
// stride (robust against centroid tables stored in a sorted order).  Tables below 2048 lists keep the dense path only.
int rbq::prepare_coarse_sample(rbq_index* h) {
    DevIndex& d = h->dev;
    d.samp_split = nullptr;
    d.samp_n2 = nullptr;
    d.samp_n = 0;
    const size_t nl = d.nlist, D = d.D;
    if (nl < 2048) return RBQ_OK;
    // 512 scores per query for tables up to 8192 lists, 1024 beyond: one warp ranks them in registers; a sparser sample
    // only loosens the threshold (more candidates per query), never the guarantee
    size_t S = std::min<size_t>(1024, std::max<size_t>(512, (nl / 16 + 255) / 256 * 256));
    const size_t stride = nl / S;
    std::vector<uint32_t> idx(S);
    uint64_t st = 0x9e3779b97f4a7c15ull ^ (uint64_t)nl;
    for (size_t j = 0; j < S; ++j) {
        st = st * 6364136223846793005ull + 1442695040888963407ull;
        idx[j] = (uint32_t)(j * stride + (size_t)((st >> 33) % stride));
    }
    uint32_t* d_idx = nullptr;
    void* sp = nullptr;
    float* n2 = nullptr;
    RBQ_CUDA(cudaMalloc(&d_idx, S * 4));
    h->allocations.push_back(d_idx);
    RBQ_CUDA(cudaMalloc(&sp, S * 3 * D * 2));
    h->allocations.push_back(sp);
    RBQ_CUDA(cudaMalloc(&n2, S * 4));
    h->allocations.push_back(n2);
    RBQ_CUDA(cudaMemcpy(d_idx, idx.data(), S * 4, cudaMemcpyHostToDevice));
    int rc = launch_gather_rows(d.cent_split, 3 * D * 2, d_idx, S, sp, nullptr);
    if (rc) return rc;
    std::vector<float> hn2(nl), sn2(S);
    RBQ_CUDA(cudaMemcpy(hn2.data(), d.cent_n2, nl * 4, cudaMemcpyDeviceToHost));
    for (size_t j = 0; j < S; ++j) sn2[j] = hn2[idx[j]];
    RBQ_CUDA(cudaMemcpy(n2, sn2.data(), S * 4, cudaMemcpyHostToDevice));
    RBQ_CUDA(cudaDeviceSynchronize());
    d.samp_split = sp;
    d.samp_n2 = n2;
    d.samp_n = (uint32_t)S;
    return RBQ_OK;
}

namespace {

int ensure_ws(const rbq_index* h, size_t bytes) {
    if (h->ws_bytes >= bytes) return RBQ_OK;
    if (h->ws) {
        RBQ_CUDA(cudaDeviceSynchronize());  // earlier asynchronous calls may still be using the old workspace
        cudaFree(h->ws);
    }
    h->ws = nullptr;
    h->ws_bytes = 0;
    RBQ_CUDA(cudaMalloc(&h->ws, bytes));
    h->ws_bytes = bytes;
    return RBQ_OK;
}

struct Carver {
    char* p;
    size_t off = 0;
    template <class T>
    T* take(size_t count) {
        off = (off + 255) & ~(size_t)255;
        T* r = reinterpret_cast<T*>(p + off);
        off += count * sizeof(T);
        return r;
    }
};

// How one call runs.  The SCAN works on tiles of `qt` queries (the list-major schedule wants as many (query, list) pairs per
// list as it can get: the whole batch when it fits), the FRONT END on chunks of `cq` queries inside a tile (its per-query
// scratch -- a dense score row or a candidate list -- is what limits a chunk, not the tile).
struct Plan {
    size_t qt = 0, cq = 0;
    int coarse = 1;        // 0 exact FP32, 1 dense tensor-core scores, 2 tensor-core scores filtered in the GEMM epilogue
    int terms = 3;         // bf16 split terms of the GEMM (3: fp32-class, 1: bf16-class with a wider re-score band)
    float eps = 0.0f;      // bound on |gemm(q.c) - q.c| / (|q||c|)
    uint32_t rank = 0, cap = 0, fb_ctas = 0;  // filter mode: sample rank of the threshold, candidate capacity, fallback CTAs
    int slots = 1;         // host entry: front-end scratch sets / streams the feed chunks alternate over (1: one stream)
};
constexpr int kMaxSlots = 4;

Plan make_plan(const rbq_index* h, size_t nq, size_t nprobe) {
    const DevIndex& ix = h->dev;
    Plan p;
    const size_t pair_cap = std::max<size_t>(1, ((size_t)1 << 26) / std::max<size_t>(nprobe, 1));
    p.qt = std::min<size_t>({nq, (size_t)131072, pair_cap});
    p.terms = h->coarse_terms == 1 ? 1 : h->coarse_terms == 3 ? 3 : (ix.D >= 512 ? 1 : 3);  // 0 = auto
    p.eps = p.terms == 1 ? 4.39453125e-3f /* 2^-8 + 2^-11: bf16 rounding of both operands */ : h->coarse_eps;
    p.coarse = h->coarse_mode;
    if (p.coarse < 0 || p.coarse == 2) {
        p.rank = filter_sample_rank(ix.nlist, ix.samp_n, nprobe, p.terms);
        p.cap = filter_cand_cap(ix.nlist, ix.samp_n, nprobe, p.terms);
        // auto: the filter whenever the index has a centroid sample (>= 2048 lists) and nprobe leaves it selective
        const bool ok = p.rank != 0 && p.cap != 0;
        p.coarse = ok ? 2 : 1;
    }
    size_t per_q = (size_t)ix.nlist * 4;
    if (p.coarse == 2) per_q = (size_t)p.cap * sizeof(CandRec) + (size_t)ix.samp_n * 4 + 16;
    size_t cq = std::max<size_t>(128, (((size_t)1 << 30) / std::max<size_t>(per_q, 1)) / 128 * 128);
    cq = std::min<size_t>(cq, 32768);
    p.cq = std::min(cq, p.qt);
    p.fb_ctas = 64;
    return p;
}

struct WsLayout {
    float* d_rot;
    uint8_t* d_lut;
    QueryScalars* d_qs;
    Probe* d_pr;
    uint16_t* d_qsplit;
    float* d_qn2;
    TailWs tw;
    uint8_t* d_head_owner;
    float* d_sc;   // coarse 0/1: [cq][nlist]
    FilterWs fw;   // coarse 2
    float* d_scx[kMaxSlots - 1];   // the same scratch for slots 1.. (plan.slots > 1)
    FilterWs fwx[kMaxSlots - 1];
    size_t end;
    float* sc(int slot) const { return slot == 0 ? d_sc : d_scx[slot - 1]; }
    const FilterWs& filter(int slot) const { return slot == 0 ? fw : fwx[slot - 1]; }
};
// One tile's device workspace.  The layout is a pure function of (plan, nprobe, top_k), so the phases of a multi-GPU search
// find each other's data by carving again.  base == nullptr only measures.
WsLayout carve_ws(const rbq_index* h, char* ws_base, const Plan& pl, size_t nprobe, size_t top_k) {
    const DevIndex& ix = h->dev;
    const size_t D = ix.D, qt = pl.qt, cq = pl.cq;
    Carver cv{ws_base};
    WsLayout L;
    L.d_rot = cv.take<float>(qt * D);
    L.d_lut = cv.take<uint8_t>(qt * D * 4);
    L.d_qs = cv.take<QueryScalars>(qt);
    L.d_pr = cv.take<Probe>(qt * nprobe);
    L.d_qsplit = cv.take<uint16_t>(qt * 3 * D);
    L.d_qn2 = cv.take<float>(qt);
    const size_t tb = tail_ws_bytes(ix, qt, nprobe, top_k);
    char* tbase = cv.take<char>(tb);
    if (ws_base) tail_ws_carve(ix, qt, nprobe, top_k, tbase, L.tw);
    L.d_head_owner = cv.take<uint8_t>(qt);
    for (int s = 0; s < std::max(1, std::min(pl.slots, kMaxSlots)); ++s) {  // slot 0 first: its offsets do not depend on plan.slots
        float*& sc = s == 0 ? L.d_sc : L.d_scx[s - 1];
        FilterWs& fw = s == 0 ? L.fw : L.fwx[s - 1];
        sc = nullptr;
        fw = FilterWs{};
        if (pl.coarse == 2) {
            fw.cap = pl.cap;
            fw.fb_ctas = pl.fb_ctas;
            fw.samp_scores = cv.take<float>(cq * ix.samp_n);
            fw.thr = cv.take<float>(cq);
            fw.cand = cv.take<CandRec>(cq * pl.cap);
            fw.cand_cnt = cv.take<uint32_t>(cq + 2);
            fw.fb_count = fw.cand_cnt + cq;
            fw.fb_list = cv.take<uint32_t>(cq);
            fw.fb_scratch = cv.take<float>((size_t)pl.fb_ctas * ix.nlist);
        } else {
            sc = cv.take<float>(cq * (size_t)ix.nlist);
        }
    }
    L.end = cv.off + 4096;
    return L;
}

size_t ws_need(const rbq_index* h, const Plan& pl, size_t nprobe, size_t top_k, size_t dim, bool host_io, size_t filter_words) {
    size_t n = carve_ws(h, nullptr, pl, nprobe, top_k).end;
    if (host_io) {
        n += pl.qt * dim * 4 + 256;
        n += pl.qt * top_k * 12 + pl.qt * 4 + 768;
        n += filter_words * 8 + 256;
    }
    return n;
}

// Front end for the queries [c0, c0 + m) of a tile (rows c0.. of the tile buffers): rotation + LUT (+ bf16 operand split),
// centroid scores, probe selection with the per-list constants.  m <= plan.cq.
int run_front(const rbq_index* h, const WsLayout& L, const Plan& pl, const float* dq, size_t c0, size_t m, size_t nprobe, cudaStream_t st,
              uint64_t* launches, bool prep, cudaEvent_t ev_prep, cudaEvent_t ev_coarse, bool need_ip = false, int slot = 0) {
    const DevIndex& ix = h->dev;
    const size_t D = ix.D;
    float* const d_sc = L.sc(slot);  // this slot's front-end scratch (chunks on different streams use different slots)
    int rc;
    const bool tc = pl.coarse != 0;
    bool split_done = false;
    if (prep) {
        if ((rc = launch_query_prep(ix, dq, m, L.d_rot + c0 * D, L.d_lut + c0 * D * 4, L.d_qs + c0, st, tc ? L.d_qsplit + c0 * 3 * D : nullptr,
                                    tc ? L.d_qn2 + c0 : nullptr, &split_done)))
            return rc;
        *launches += 1;
    }
    if (ev_prep) cudaEventRecord(ev_prep, st);
    if (pl.coarse == 0) {
        if ((rc = launch_coarse_exact(ix, L.d_rot + c0 * D, m, d_sc, st))) return rc;
        if (ev_coarse) cudaEventRecord(ev_coarse, st);
        if ((rc = launch_probe_select(ix, L.d_rot + c0 * D, d_sc, m, nprobe, L.d_pr + c0 * nprobe, st))) return rc;
        *launches += 2;
        return RBQ_OK;
    }
    if (!split_done) {
        if ((rc = launch_split_bf16(L.d_rot + c0 * D, m, (int)D, 0, L.d_qsplit + c0 * 3 * D, L.d_qn2 + c0, st))) return rc;
        *launches += 1;
    }
    if (pl.coarse == 1) {
        if ((rc = launch_coarse_tc(ix, L.d_qsplit + c0 * 3 * D, L.d_qn2 + c0, m, d_sc, st, pl.terms))) return rc;
        if (ev_coarse) cudaEventRecord(ev_coarse, st);
        if ((rc = launch_probe_select_tc(ix, L.d_rot + c0 * D, d_sc, L.d_qs + c0, m, nprobe, pl.eps, L.d_pr + c0 * nprobe,
                                         h->fallback_counter(), st, need_ip)))
            return rc;
        *launches += 2;
        return RBQ_OK;
    }
    // filter mode: sample scores -> per-query threshold -> full GEMM keeping the centroids that beat it -> selection
    const FilterWs& fw = L.filter(slot);
    RBQ_CUDA(cudaMemsetAsync(fw.cand_cnt, 0, (pl.cq + 2) * 4, st));  // candidate counters | fallback count | fallback cursor
    GemmEpi e;
    e.nq = (int)m;
    e.metric = ix.metric;
    e.shifted = 1;  // sample scores, filter threshold and candidate scores all live in the shifted domain (|c|^2 - 2 q.c)
    e.qn2 = L.d_qn2 + c0;
    e.ncols = (int)ix.samp_n;
    e.cn2 = ix.samp_n2;
    e.scores = fw.samp_scores;
    if ((rc = launch_coarse_gemm(kGemmScores, L.d_qsplit + c0 * 3 * D, m, ix.samp_split, ix.samp_n, (int)D, pl.terms, e, st))) return rc;
    if ((rc = launch_sample_threshold(fw.samp_scores, m, ix.samp_n, pl.rank, ix.metric, fw.thr, st))) return rc;
    e.ncols = (int)ix.nlist;
    e.cn2 = ix.cent_n2;
    e.scores = nullptr;
    e.thr = fw.thr;
    e.cand = fw.cand;
    e.cand_cnt = fw.cand_cnt;
    e.cap = fw.cap;
    if ((rc = launch_coarse_gemm(kGemmFilter, L.d_qsplit + c0 * 3 * D, m, ix.cent_split, ix.nlist, (int)D, pl.terms, e, st))) return rc;
    if (ev_coarse) cudaEventRecord(ev_coarse, st);
    // rows of rot / qs / probes are tile rows c0.. ; the candidate lists are chunk rows 0..
    if ((rc = launch_probe_select_cand(ix, L.d_rot + c0 * D, L.d_qs + c0, m, nprobe, pl.eps, fw, L.d_pr + c0 * nprobe, h->fallback_counter(), st,
                                       need_ip)))
        return rc;
    *launches += 5;
    return RBQ_OK;
}

// Queries still in host memory: chunks are copied on a dedicated non-blocking stream, each followed by an event the
// compute stream waits on before that chunk's front end.
struct HostFeed {
    static constexpr int kEvents = 16;
    const float* h_q = nullptr;   // host queries (whole call)
    float* d_q = nullptr;         // device staging for one tile
    size_t dim = 0, chunk = 0;
    size_t first = 0;             // queries in the first chunk (0: same as the others): a short first copy starts the GPU earlier
    bool taper = false;           // the last chunks shrink (halving): the tail stage waits for the LAST chunk's front end + head pass
    cudaStream_t copy = nullptr;
    cudaEvent_t* ev = nullptr;
    int issue(size_t q_abs, size_t m, size_t off_in_tile, int slot) {
        RBQ_CUDA(cudaMemcpyAsync(d_q + off_in_tile * dim, h_q + q_abs * dim, m * dim * 4, cudaMemcpyHostToDevice, copy));
        RBQ_CUDA(cudaEventRecord(ev[slot % kEvents], copy));
        return RBQ_OK;
    }
};

// The queries the head pass could not take (nearest list longer than the dense buffer, heap not full after it) are known once
// the head kernels are enqueued; their sequential walk is one query's latency chain (~0.6 ms at 100M x 128) that nothing else
// depends on, so it runs on a side stream beside the tail and replay kernels and is joined before the call's last kernel.
int side_fork(const rbq_index* h, cudaStream_t st) {
    if (!h->side_stream) {
        int lo = 0, hi = 0;  // highest priority: its few CTAs must become resident before the persistent tail / replay grids fill the SMs
        RBQ_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        RBQ_CUDA(cudaStreamCreateWithPriority(&h->side_stream, cudaStreamNonBlocking, hi));
        RBQ_CUDA(cudaEventCreateWithFlags(&h->side_fork, cudaEventDisableTiming));
        RBQ_CUDA(cudaEventCreateWithFlags(&h->side_join, cudaEventDisableTiming));
    }
    RBQ_CUDA(cudaEventRecord(h->side_fork, st));
    RBQ_CUDA(cudaStreamWaitEvent(h->side_stream, h->side_fork, 0));
    return RBQ_OK;
}
int side_join(const rbq_index* h, cudaStream_t st) {
    RBQ_CUDA(cudaEventRecord(h->side_join, h->side_stream));
    RBQ_CUDA(cudaStreamWaitEvent(st, h->side_join, 0));
    return RBQ_OK;
}

// The pipeline on device buffers for one call (tiles internally).  d_filter may be null.
int search_device(const rbq_index* h, const float* d_queries, size_t nq, size_t top_k, size_t nprobe,
                  const uint64_t* d_filter, size_t filter_nbits, uint64_t* d_ids, float* d_scores, uint32_t* d_counts,
                  char* ws_base, const Plan& pl, cudaStream_t st, uint64_t* launches, HostFeed* feed = nullptr) {
    const DevIndex& ix = h->dev;
    const size_t qt = pl.qt;
    const WsLayout L = carve_ws(h, ws_base, pl, nprobe, top_k);
    const TailWs& tw = L.tw;
    float ms[7] = {0, 0, 0, 0, 0, 0, 0};  // [6]: the tail FastScan kernel alone
    bool any_list_major = false;
    for (size_t q0 = 0; q0 < nq; q0 += qt) {
        const size_t n = std::min(qt, nq - q0);
        int rc;
        if (h->profiling) cudaEventRecord(h->ev[0], st);
        // Scan-stage schedule.  Sequential: one warp walks one query's whole probe sequence.  List-major (large batches):
        // head pass (the first owned list, sequential, fills the heap) -> tail kernel (all remaining pairs grouped by list)
        // -> replay pass (survivors in reference order).  Both are exact.
        const bool list_major = h->scan_mode == 2 || (h->scan_mode == 0 && nprobe >= 4 && n * nprobe >= 8 * (size_t)ix.nlist);
        any_list_major |= list_major;
        if (list_major)  // surv_cnt | list_cnt | list_fill | counters
            RBQ_CUDA(cudaMemsetAsync(tw.surv_cnt, 0, (qt + 2 * (size_t)ix.nlist + kTailCounters + kHeadCursors) * 4, st));
        // Front end (rotate + LUT, coarse scores, probe selection) and the head pass are independent per query: they run chunk
        // by chunk (a chunk = what the front end's scratch holds; smaller when the queries are still arriving from the host, so
        // that the H2D transfer of chunk c+1 overlaps the work on chunk c) and only the tail stage waits for the whole tile.
        size_t chunk = std::min(pl.cq, n);
        if (feed) chunk = std::max<size_t>(128, std::min(chunk, feed->chunk));
        // Chunk slots (host entry, unprofiled): chunk ci runs on stream ci % nslots with that slot's front-end scratch and its own
        // part of the dense head buffer.  A chunk of a few thousand queries fills the GPU only partly (0.6 waves of the prep kernel
        // at 2 500 queries); on one stream that idle share is lost in every kernel (front end + head pass of a 10 000-query batch:
        // 0.84 ms in one chunk, 1.12 in two, 1.47 in four), on alternating streams the next chunk's kernels take it.
        int nslots = (feed && !h->profiling) ? std::max(1, std::min(pl.slots, kMaxSlots)) : 1;
        const uint32_t slot_rows = tw.head_rows / (uint32_t)nslots;
        if (chunk >= n || slot_rows < 128 || ((n + chunk - 1) / chunk + 7) * ((chunk + slot_rows - 1) / std::max(slot_rows, 1u)) > kHeadCursors) nslots = 1;
        cudaStream_t cs[kMaxSlots] = {st, nullptr, nullptr, nullptr};
        TailWs tws[kMaxSlots] = {tw, tw, tw, tw};
        if (nslots > 1) {
            if (!h->slot_fork) {
                RBQ_CUDA(cudaEventCreateWithFlags(&h->slot_fork, cudaEventDisableTiming));
                for (int i = 0; i < kMaxSlots - 1; ++i) {
                    RBQ_CUDA(cudaStreamCreateWithFlags(&h->slot_stream[i], cudaStreamNonBlocking));
                    RBQ_CUDA(cudaEventCreateWithFlags(&h->slot_join[i], cudaEventDisableTiming));
                }
            }
            RBQ_CUDA(cudaEventRecord(h->slot_fork, st));  // after the tile's memset (and whatever the call enqueued before)
            for (int i = 1; i < nslots; ++i) {
                cs[i] = h->slot_stream[i - 1];
                RBQ_CUDA(cudaStreamWaitEvent(cs[i], h->slot_fork, 0));
            }
            for (int i = 0; i < nslots; ++i) {
                tws[i].head_rows = slot_rows;
                tws[i].head_buf = tw.head_buf + (size_t)i * slot_rows * tw.head_cap;
            }
        }
        int head_launch = 0, ci = 0;
        for (size_t c0 = 0, m = 0; c0 < n; c0 += m, ++ci) {
            m = std::min((feed && ci == 0 && feed->first) ? std::min(feed->first, chunk) : chunk, n - c0);
            if (feed && feed->taper && nslots > 1 && n - c0 <= 2 * chunk) {
                // Tapered end of the tile: everything after the last chunk's arrival is on the critical path (its front end and head
                // pass, then the tail stage), and a chunk's kernels take time roughly in proportion to its size, so the final chunks
                // halve down to a few hundred queries; the slot streams absorb the extra launches.
                const size_t rem = n - c0;
                m = rem <= 640 ? rem : std::min(chunk, std::max<size_t>(256, (rem / 2 + 127) / 128 * 128));
            }
            const int slot = ci % nslots;
            cudaStream_t cst = cs[slot];
            if (feed) {
                if ((rc = feed->issue(q0 + c0, m, c0, ci))) return rc;
                RBQ_CUDA(cudaStreamWaitEvent(cst, feed->ev[ci % HostFeed::kEvents], 0));
            }
            // profiled calls time every chunk's stages with its own events (read back after the tile)
            cudaEvent_t* ce = nullptr;
            if (h->profiling && ci < 32) {
                ce = h->ev_chunk[ci];
                for (int i = 0; i < 5; ++i)
                    if (!ce[i]) RBQ_CUDA(cudaEventCreate(&ce[i]));
                cudaEventRecord(ce[0], cst);
            }
            const float* dq = feed ? feed->d_q + c0 * ix.dim : d_queries + (q0 + c0) * ix.dim;
            if ((rc = run_front(h, L, pl, dq, c0, m, nprobe, cst, launches, true, ce ? ce[1] : nullptr, ce ? ce[2] : nullptr, false, slot))) return rc;
            if (ce) cudaEventRecord(ce[3], cst);
            // head: FastScan of every query's first owned list + the reference's sequential loop over it
            if (list_major && (rc = launch_head(ix, L.d_rot, L.d_lut, L.d_qs, L.d_pr, n, nprobe, top_k, d_filter, filter_nbits, d_ids + q0 * top_k,
                                                d_scores + q0 * top_k, d_counts + q0, h->d_stats, tws[slot], cst, launches, c0, m, &head_launch)))
                return rc;
            if (ce) cudaEventRecord(ce[4], cst);
        }
        for (int i = 1; i < nslots; ++i) {  // the tail stage needs every chunk's probes and head state
            RBQ_CUDA(cudaEventRecord(h->slot_join[i - 1], cs[i]));
            RBQ_CUDA(cudaStreamWaitEvent(st, h->slot_join[i - 1], 0));
        }
        const int timed_chunks = std::min(ci, 32);
        if (!list_major) {
            if (h->profiling) cudaEventRecord(h->ev[4], st);
            if ((rc = launch_scan(ix, L.d_rot, L.d_lut, L.d_qs, L.d_pr, n, nprobe, top_k, d_filter, filter_nbits, d_ids + q0 * top_k,
                                  d_scores + q0 * top_k, d_counts + q0, h->d_stats, h->work_counter(), kScanFull, nullptr, st)))
                return rc;
            *launches += 1;
            if (h->profiling) {  // sequential schedule: the whole scan is reported as the head stage (ms[3] += ev4 -> ev5)
                cudaEventRecord(h->ev[5], st);
                cudaEventRecord(h->ev[6], st);
            }
        } else {
            if (h->profiling) cudaEventRecord(h->ev[4], st);
            // the head pass' fallback queries: sequential walk on the side stream (normally an empty launch)
            if ((rc = side_fork(h, st))) return rc;
            if ((rc = launch_scan(ix, L.d_rot, L.d_lut, L.d_qs, L.d_pr, n, nprobe, top_k, d_filter, filter_nbits, d_ids + q0 * top_k,
                                  d_scores + q0 * top_k, d_counts + q0, h->d_stats, h->work_counter(), kScanFallback, &tw, h->side_stream)))
                return rc;
            // tail: all remaining (query, list) pairs grouped by list -> survivors
            if ((rc = launch_tail(ix, L.d_lut, L.d_qs, L.d_pr, n, nprobe, d_filter, filter_nbits, h->d_stats, tw, st, launches,
                                  h->profiling ? h->ev[7] : nullptr, h->profiling ? h->ev[8] : nullptr)))
                return rc;
            if (h->profiling) cudaEventRecord(h->ev[5], st);
            // ordered replay of the survivors with refinement on demand (+ the overflow tier), then the queries the replay tiers
            // handed back (normally none), sequentially
            if ((rc = launch_refine_replay(ix, L.d_rot, L.d_qs, L.d_pr, n, nprobe, top_k, d_ids + q0 * top_k, d_scores + q0 * top_k,
                                           d_counts + q0, h->d_stats, tw, st, launches)))
                return rc;
            if ((rc = launch_scan(ix, L.d_rot, L.d_lut, L.d_qs, L.d_pr, n, nprobe, top_k, d_filter, filter_nbits, d_ids + q0 * top_k,
                                  d_scores + q0 * top_k, d_counts + q0, h->d_stats, h->work_counter(), kScanFallbackResume, &tw, st)))
                return rc;
            if ((rc = side_join(h, st))) return rc;
            *launches += 2;
            if (h->profiling) cudaEventRecord(h->ev[6], st);
        }
        if (h->profiling) {
            RBQ_CUDA(cudaEventSynchronize(h->ev[6]));
            auto el = [&](cudaEvent_t a, cudaEvent_t b) {
                float t = 0;
                return cudaEventElapsedTime(&t, a, b) == cudaSuccess ? t : 0.0f;
            };
            for (int c = 0; c < timed_chunks; ++c) {  // prep | coarse | select | head of every chunk
                cudaEvent_t* ce = h->ev_chunk[c];
                for (int i = 0; i < 4; ++i) ms[i] += el(ce[i], ce[i + 1]);
            }
            ms[4] += el(h->ev[4], h->ev[5]);
            ms[5] += el(h->ev[5], h->ev[6]);
            if (list_major) ms[6] += el(h->ev[7], h->ev[8]);
        }
    }
    if (h->profiling) {
        h->last_stats.ms_prep = ms[0];
        h->last_stats.ms_coarse = ms[1];
        h->last_stats.ms_select = ms[2];
        if (!any_list_major) {  // sequential schedule: everything sits in the "tail" slot of the events
            ms[3] += ms[4];
            ms[4] = 0.0f;
        }
        h->last_stats.ms_scan = ms[3] + ms[4] + ms[5];
        h->last_stats.ms_scan_head = ms[3];
        h->last_stats.ms_scan_tail = ms[4];
        h->last_stats.ms_scan_replay = ms[5];
        h->last_stats.ms_tail_kernel = ms[6];
    }
    return RBQ_OK;
}

// Calls on one handle share its workspace: every entry point that enqueues work first makes its stream wait for the
// previous call's last kernel (busy event) and records the event again when it is done enqueuing, so calls from other
// threads or on other streams are ordered on the device, not only on the host.
struct Serial {
    const rbq_index* h;
    cudaStream_t st;
    Serial(const rbq_index* h_, cudaStream_t st_) : h(h_), st(st_) {
        if (!h->busy_ev) cudaEventCreateWithFlags(&h->busy_ev, cudaEventDisableTiming);
        else cudaStreamWaitEvent(st, h->busy_ev, 0);
    }
    ~Serial() { cudaEventRecord(h->busy_ev, st); }
};

int check_search_args(const rbq_index* ix, size_t dim, size_t top_k, size_t* nprobe) {
    if (!ix) return fail(RBQ_INVALID_CONFIG, "null index handle");
    if (ix->host.nvec_total == 0) return fail(RBQ_EMPTY_INDEX, "index is empty; call `train` first");
    if (dim != ix->host.dim)
        return fail(RBQ_DIMENSION_MISMATCH, "expected " + std::to_string(ix->host.dim) + ", got " + std::to_string(dim));
    *nprobe = std::min(std::max<size_t>(*nprobe, 1), ix->host.nlist);  // reference src/ivf.rs:1791
    if (top_k > scan_max_topk()) return fail(RBQ_INVALID_CONFIG, "top_k exceeds the device limit (1024)");
    if (*nprobe > probe_select_max_nprobe())
        return fail(RBQ_INVALID_CONFIG, "nprobe exceeds the device probe-selection limit (4096)");
    return RBQ_OK;
}

int finish_load(rbq_index* h, int device, rbq_index** out) {
    h->device = device;
    DeviceGuard g(device);
    int rc = upload_index(h);
    if (rc) {
        rbq_index_free(h);
        return rc;
    }
    for (auto& e : h->ev) cudaEventCreate(&e);
    *out = h;
    return RBQ_OK;
}

}  // namespace

extern "C" {

int rbq_index_load_mem(const uint8_t* bytes, size_t len, int device, int shard_rank, int shard_count, rbq_index** out) {
    if (!bytes || !out) return fail(RBQ_INVALID_CONFIG, "null argument");
    *out = nullptr;
    rbq_index* h = new rbq_index();
    int rc = parse_rbq1(bytes, len, shard_rank, shard_count, h->host);
    if (rc) {
        delete h;
        return rc;
    }
    return finish_load(h, device, out);
}

int rbq_index_load(const char* path, int device, int shard_rank, int shard_count, rbq_index** out) {
    if (!path || !out) return fail(RBQ_INVALID_CONFIG, "null argument");
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(RBQ_IO, std::string(strerror(errno)) + ": " + path);
    struct stat sb;
    if (fstat(fd, &sb) != 0) {
        close(fd);
        return fail(RBQ_IO, std::string(strerror(errno)) + ": " + path);
    }
    size_t len = (size_t)sb.st_size;
    void* map = len ? mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0) : nullptr;
    close(fd);
    if (len && map == MAP_FAILED) return fail(RBQ_IO, std::string("mmap failed: ") + path);
    static const uint8_t empty = 0;
    int rc = rbq_index_load_mem(len ? (const uint8_t*)map : &empty, len, device, shard_rank, shard_count, out);
    if (len) munmap(map, len);
    return rc;
}

void rbq_index_free(rbq_index* h) {
    if (!h) return;
    {
        DeviceGuard g(h->device);
        for (void* p : h->allocations) cudaFree(p);
        if (h->ws) cudaFree(h->ws);
        for (auto& e : h->ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : h->feed_ev)
            if (e) cudaEventDestroy(e);
        for (auto& row : h->ev_chunk)
            for (auto& e : row)
                if (e) cudaEventDestroy(e);
        if (h->comm) nccl_api().comm_destroy(h->comm);
        if (h->dist_ws) cudaFree(h->dist_ws);
        if (h->xr_ws) cudaFree(h->xr_ws);
        if (h->xr_recs) cudaFree(h->xr_recs);
        if (h->xr_host) cudaFreeHost(h->xr_host);
        for (auto& x : h->slot_stream)
            if (x) cudaStreamDestroy(x);
        for (auto& x : h->slot_join)
            if (x) cudaEventDestroy(x);
        if (h->slot_fork) cudaEventDestroy(h->slot_fork);
        if (h->side_stream) cudaStreamDestroy(h->side_stream);
        if (h->side_fork) cudaEventDestroy(h->side_fork);
        if (h->side_join) cudaEventDestroy(h->side_join);
        if (h->busy_ev) cudaEventDestroy(h->busy_ev);
        if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
        if (h->compute_stream) cudaStreamDestroy(h->compute_stream);
    }
    delete h;
}

// keep (optional): nlist flags; lists with keep[c] == 0 are written empty (their centroid stays)
static int save_impl(const rbq_index* h, const uint8_t* keep, uint8_t* out, size_t cap, size_t* written) {
    if (!h || !written) return fail(RBQ_INVALID_CONFIG, "null argument");
    if (h->host.shard_count != 1) return fail(RBQ_INVALID_CONFIG, "only a complete (unsharded) index can be saved");
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->busy_ev) RBQ_CUDA(cudaEventSynchronize(h->busy_ev));
    const HostIndex& src = h->host;
    const size_t nlist = src.nlist, stride = src.block_stride(), exs = src.ex_stride();
    HostIndex tmp;  // metadata + centroids; bulk arrays come back from the device
    tmp.dim = src.dim;
    tmp.D = src.D;
    tmp.metric = src.metric;
    tmp.rot_type = src.rot_type;
    tmp.ex_bits = src.ex_bits;
    tmp.rot_bytes = src.rot_bytes;
    tmp.nlist = nlist;
    tmp.centroids = src.centroids;
    tmp.list_n.assign(nlist, 0);
    tmp.blk_off.assign(nlist + 1, 0);
    tmp.vec_off.assign(nlist + 1, 0);
    uint64_t nv = 0, nb = 0;
    for (size_t c = 0; c < nlist; ++c) {
        tmp.blk_off[c] = (uint32_t)nb;
        tmp.vec_off[c] = nv;
        if (!keep || keep[c]) {
            tmp.list_n[c] = src.list_n[c];
            nv += src.list_n[c];
            nb += (src.list_n[c] + kBatch - 1) / kBatch;
        }
    }
    tmp.blk_off[nlist] = (uint32_t)nb;
    tmp.vec_off[nlist] = nv;
    tmp.list_n_all = tmp.list_n;
    tmp.nvec_total = nv;
    tmp.blocks.resize(nb * stride);
    tmp.ids.resize(nv);
    tmp.ex.resize(nv * exs);
    tmp.f_add_ex.resize(nv);
    tmp.f_rescale_ex.resize(nv);
    tmp.delta.resize(nv);
    tmp.vl.resize(nv);
    if (!keep) {
        if (!tmp.blocks.empty()) RBQ_CUDA(cudaMemcpy(tmp.blocks.data(), h->dev.blocks, tmp.blocks.size(), cudaMemcpyDeviceToHost));
        if (nv) {
            RBQ_CUDA(cudaMemcpy(tmp.ids.data(), h->dev.ids, nv * 8, cudaMemcpyDeviceToHost));
            if (!tmp.ex.empty()) RBQ_CUDA(cudaMemcpy(tmp.ex.data(), h->dev.ex, tmp.ex.size(), cudaMemcpyDeviceToHost));
            RBQ_CUDA(cudaMemcpy(tmp.f_add_ex.data(), h->dev.f_add_ex, nv * 4, cudaMemcpyDeviceToHost));
            RBQ_CUDA(cudaMemcpy(tmp.f_rescale_ex.data(), h->dev.f_rescale_ex, nv * 4, cudaMemcpyDeviceToHost));
        }
        tmp.delta = src.delta;
        tmp.vl = src.vl;
    } else {
        for (size_t c = 0; c < nlist; ++c) {
            const size_t n = tmp.list_n[c];
            if (!n) continue;
            const size_t so = src.vec_off[c], to = tmp.vec_off[c], nbl = (n + kBatch - 1) / kBatch;
            RBQ_CUDA(cudaMemcpy(tmp.blocks.data() + (size_t)tmp.blk_off[c] * stride, h->dev.blocks + (size_t)src.blk_off[c] * stride, nbl * stride,
                                cudaMemcpyDeviceToHost));
            RBQ_CUDA(cudaMemcpy(tmp.ids.data() + to, h->dev.ids + so, n * 8, cudaMemcpyDeviceToHost));
            if (exs) RBQ_CUDA(cudaMemcpy(tmp.ex.data() + to * exs, h->dev.ex + so * exs, n * exs, cudaMemcpyDeviceToHost));
            RBQ_CUDA(cudaMemcpy(tmp.f_add_ex.data() + to, h->dev.f_add_ex + so, n * 4, cudaMemcpyDeviceToHost));
            RBQ_CUDA(cudaMemcpy(tmp.f_rescale_ex.data() + to, h->dev.f_rescale_ex + so, n * 4, cudaMemcpyDeviceToHost));
            std::memcpy(tmp.delta.data() + to, src.delta.data() + so, n * 4);
            std::memcpy(tmp.vl.data() + to, src.vl.data() + so, n * 4);
        }
    }
    std::vector<uint8_t> buf;
    write_rbq1(tmp, buf);
    *written = buf.size();
    if (out) {
        if (cap < buf.size()) return fail(RBQ_IO, "output buffer too small");
        std::memcpy(out, buf.data(), buf.size());
    }
    return RBQ_OK;
}

int rbq_index_save_mem(const rbq_index* h, uint8_t* out, size_t cap, size_t* written) { return save_impl(h, nullptr, out, cap, written); }

int rbq_index_save_lists_mem(const rbq_index* h, const uint8_t* keep_list, size_t nlist, uint8_t* out, size_t cap, size_t* written) {
    if (!h || !keep_list) return fail(RBQ_INVALID_CONFIG, "null argument");
    if (nlist != h->host.nlist) return fail(RBQ_INVALID_CONFIG, "keep_list must hold one flag per cluster");
    return save_impl(h, keep_list, out, cap, written);
}

int rbq_index_save(const rbq_index* h, const char* path) {
    if (!h || !path) return fail(RBQ_INVALID_CONFIG, "null argument");
    size_t n = 0;
    int rc = rbq_index_save_mem(h, nullptr, 0, &n);
    if (rc) return rc;
    std::vector<uint8_t> buf(n);
    if ((rc = rbq_index_save_mem(h, buf.data(), n, &n))) return rc;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(RBQ_IO, std::string(strerror(errno)) + ": " + path);
    size_t w = fwrite(buf.data(), 1, n, f);
    if (fclose(f) != 0 || w != n) return fail(RBQ_IO, std::string("short write: ") + path);
    return RBQ_OK;
}

size_t rbq_index_len(const rbq_index* h) { return h ? (size_t)h->host.nvec_total : 0; }
size_t rbq_index_local_len(const rbq_index* h) { return h ? (size_t)h->host.vec_off[h->host.nlist] : 0; }
size_t rbq_index_dim(const rbq_index* h) { return h ? h->host.dim : 0; }
size_t rbq_index_padded_dim(const rbq_index* h) { return h ? h->host.D : 0; }
size_t rbq_index_cluster_count(const rbq_index* h) { return h ? h->host.nlist : 0; }
int rbq_index_metric(const rbq_index* h) { return h ? h->host.metric : -1; }
int rbq_index_ex_bits(const rbq_index* h) { return h ? h->host.ex_bits : -1; }
int rbq_index_rotator_type(const rbq_index* h) { return h ? h->host.rot_type : -1; }
int rbq_index_device(const rbq_index* h) { return h ? h->device : -1; }

int rbq_set_profiling(rbq_index* h, int on) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    h->profiling = on != 0;
    return RBQ_OK;
}
int rbq_set_coarse_mode(rbq_index* h, int mode) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    if (mode < -1 || mode > 2)
        return fail(RBQ_INVALID_CONFIG, "coarse mode must be -1 (auto), 0 (exact), 1 (dense tensor-core scores) or 2 (filtered in the GEMM epilogue)");
    std::lock_guard<std::mutex> lk(h->mu);
    h->coarse_mode = mode;
    return RBQ_OK;
}
int rbq_set_coarse_terms(rbq_index* h, int terms) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    if (terms != 0 && terms != 1 && terms != 3) return fail(RBQ_INVALID_CONFIG, "coarse terms must be 0 (auto), 1 or 3");
    std::lock_guard<std::mutex> lk(h->mu);
    h->coarse_terms = terms;
    return RBQ_OK;
}

int rbq_set_scan_mode(rbq_index* h, int mode) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    if (mode < 0 || mode > 2) return fail(RBQ_INVALID_CONFIG, "scan mode must be 0 (auto), 1 (sequential) or 2 (list-major)");
    h->scan_mode = mode;
    return RBQ_OK;
}

int rbq_debug_set_survivor_cap(uint32_t cap) {
    if (cap > 1024) return fail(RBQ_INVALID_CONFIG, "survivor cap must be <= 1024");
    tail_debug_set_survivor_cap(cap);
    return RBQ_OK;
}

int rbq_last_search_stats(const rbq_index* h, rbq_search_stats* out) {
    if (!h || !out) return fail(RBQ_INVALID_CONFIG, "null argument");
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    DevStats ds;
    if (h->busy_ev) RBQ_CUDA(cudaEventSynchronize(h->busy_ev));  // asynchronous entry points may still be running
    if (h->dist_phase == 3) {  // stage times of a profiled phased search (the events were recorded by the three rbq_dist_* calls)
        h->dist_phase = 0;
        auto el = [&](int a, int b) {
            float t = 0;
            return cudaEventElapsedTime(&t, h->ev[a], h->ev[b]) == cudaSuccess ? t : 0.0f;
        };
        h->last_stats.ms_prep = el(0, 1);
        h->last_stats.ms_coarse = el(1, 2);  // coarse scores + probe selection + export of this rank's query slice
        h->last_stats.ms_select = 0.0f;
        h->last_stats.ms_scan_head = el(3, 4);
        h->last_stats.ms_scan_tail = el(5, 9);
        h->last_stats.ms_scan_replay = el(9, 6);
        h->last_stats.ms_tail_kernel = el(7, 8);
        h->last_stats.ms_scan = h->last_stats.ms_scan_head + h->last_stats.ms_scan_tail + h->last_stats.ms_scan_replay;
    }
    RBQ_CUDA(cudaMemcpy(&ds, h->d_stats, sizeof(ds), cudaMemcpyDeviceToHost));
    h->last_stats.blocks_scanned = ds.blocks;
    h->last_stats.bytes_scanned = ds.blocks * (uint64_t)h->dev.block_stride;
    h->last_stats.candidates = ds.candidates;
    h->last_stats.refined = ds.refined;
    h->last_stats.admitted = ds.admitted;
    h->last_stats.tail_blocks = ds.tail_blocks;
    h->last_stats.tail_bytes = ds.tail_blocks * (uint64_t)h->dev.block_stride;
    h->last_stats.tail_pairs = ds.tail_pairs;
    h->last_stats.survivors = ds.survivors;
    h->last_stats.overflow_queries = ds.overflow_queries;
    h->last_stats.fallback_queries = ds.fallback_queries;
    unsigned int fb = 0;
    RBQ_CUDA(cudaMemcpy(&fb, h->fallback_counter(), sizeof(fb), cudaMemcpyDeviceToHost));
    h->last_stats.coarse_fallbacks = fb;
    if (h->xr_inexact) RBQ_CUDA(cudaMemcpy(&h->last_stats.inexact_queries, h->xr_inexact, 8, cudaMemcpyDeviceToHost));
    *out = h->last_stats;
    return RBQ_OK;
}

int rbq_search_batch_device(const rbq_index* h, const float* d_queries, size_t nq, size_t dim, size_t top_k,
                            size_t nprobe, const uint64_t* d_filter_bits, size_t filter_nbits, uint64_t* d_ids,
                            float* d_scores, uint32_t* d_counts, void* stream) {
    int rc = check_search_args(h, dim, top_k, &nprobe);
    if (rc) return rc;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Serial serial(h, st);
    h->last_stats = rbq_search_stats{};
    h->last_stats.queries = nq;
    if (nq == 0) return RBQ_OK;
    if (top_k == 0) {  // reference src/ivf.rs:1792-1794
        RBQ_CUDA(cudaMemsetAsync(d_counts, 0, nq * 4, st));
        return RBQ_OK;
    }
    const Plan pl = make_plan(h, nq, nprobe);
    if ((rc = ensure_ws(h, ws_need(h, pl, nprobe, top_k, dim, false, 0)))) return rc;
    RBQ_CUDA(cudaMemsetAsync(h->d_stats, 0, sizeof(DevStats) + 16, st));
    uint64_t launches = 0;
    rc = search_device(h, d_queries, nq, top_k, nprobe, d_filter_bits, filter_nbits, d_ids, d_scores, d_counts,
                       (char*)h->ws, pl, st, &launches);
    h->last_stats.kernel_launches = launches;
    h->last_stats.coarse_mode_used = (uint32_t)pl.coarse;
    h->last_stats.coarse_terms_used = (uint32_t)pl.terms;
    h->last_stats.front_chunk = (uint32_t)pl.cq;
    return rc;
}

int rbq_search_batch_filtered(const rbq_index* h, const float* queries, size_t nq, size_t dim, size_t top_k,
                              size_t nprobe, const uint64_t* filter_bits, size_t filter_nbits, uint64_t* ids,
                              float* scores, uint32_t* counts) {
    int rc = check_search_args(h, dim, top_k, &nprobe);
    if (rc) return rc;
    if (nq && (!queries || !counts || (top_k && (!ids || !scores)))) return fail(RBQ_INVALID_CONFIG, "null buffer");
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    h->last_stats = rbq_search_stats{};
    h->last_stats.queries = nq;
    if (nq == 0) return RBQ_OK;
    if (top_k == 0) {
        std::memset(counts, 0, nq * 4);
        return RBQ_OK;
    }
    Plan pl = make_plan(h, nq, nprobe);
    const size_t qt = pl.qt;
    {   // feed chunks alternate over `slots` streams, each with its own front-end scratch (search_device); RBQ_FEED_SLOTS overrides
        const char* fs = getenv("RBQ_FEED_SLOTS");
        pl.slots = fs ? (int)std::min<long>(kMaxSlots, std::max(1L, atol(fs))) : kMaxSlots;
        // every slot carries a front-end scratch of plan.cq rows: no larger than a feed chunk can be (about a quarter of a tile)
        if (pl.slots > 1) pl.cq = std::min(pl.cq, std::max<size_t>(5120, (qt / 4 + 255) / 128 * 128));
    }
    // a filter that is present but empty (RoaringBitmap::new()) admits nothing (reference src/tests.rs
    // filtered_search_with_empty_filter): keep one zero word so the kernels see a non-null filter
    const size_t fwords = filter_bits ? std::max<size_t>((filter_nbits + 63) / 64, 1) : 0;
    if ((rc = ensure_ws(h, ws_need(h, pl, nprobe, top_k, dim, true, fwords)))) return rc;
    // host-io buffers live behind the pipeline buffers
    Carver cv{(char*)h->ws};
    cv.off = ws_need(h, pl, nprobe, top_k, dim, false, 0);
    float* d_q = cv.take<float>(qt * dim);
    uint64_t* d_ids = cv.take<uint64_t>(qt * top_k);
    float* d_sc = cv.take<float>(qt * top_k);
    uint32_t* d_cn = cv.take<uint32_t>(qt);
    uint64_t* d_f = fwords ? cv.take<uint64_t>(fwords) : nullptr;
    if (!h->copy_stream) {
        RBQ_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        RBQ_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
        for (auto& e : h->feed_ev) RBQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // the host entry point runs on the handle's own streams (the legacy default stream would serialise with every
    // blocking stream of the process); the call is synchronous, so nothing outlives it
    cudaStream_t st = getenv("RBQ_LEGACY_STREAM") ? nullptr : h->compute_stream;
    Serial serial(h, st);
    if (d_f) {
        const size_t src_words = (filter_nbits + 63) / 64;
        if (src_words) RBQ_CUDA(cudaMemcpyAsync(d_f, filter_bits, src_words * 8, cudaMemcpyHostToDevice, st));
        else RBQ_CUDA(cudaMemsetAsync(d_f, 0, 8, st));
    }
    RBQ_CUDA(cudaMemsetAsync(h->d_stats, 0, sizeof(DevStats) + 16, st));
    uint64_t launches = 0;
    HostFeed feed;
    feed.h_q = queries;
    feed.d_q = d_q;
    feed.dim = dim;
    feed.copy = getenv("RBQ_FEED_SAME_STREAM") ? st : h->copy_stream;
    feed.ev = h->feed_ev;
    // the copy stream reuses the staging buffer of the previous call: it must not run ahead of that call's kernels
    if (feed.copy != st) RBQ_CUDA(cudaStreamWaitEvent(feed.copy, h->busy_ev, 0));
    for (size_t q0 = 0; q0 < nq; q0 += qt) {
        const size_t n = std::min(qt, nq - q0);
        // The H2D copy of chunk c+1 overlaps the front end and the head pass of chunk c.  More chunks hide more of the copy, but a
        // chunk of a few thousand queries runs the per-query kernels on partial waves: on ONE stream the front end + head pass of
        // a 10 000-query batch costs 0.84 ms in one chunk, 1.12 in two, 1.47 in four, 2.28 in eight.  The chunks therefore
        // alternate over plan.slots streams (search_device), where the next chunk's kernels fill the idle share.  Measured at
        // GIST-1M / 10k queries with profiles/e2e_probe.py (ms per call; H2D alone 0.72, device-resident step 1.53):
        //   chunks      2     3     4     5     6     8    12
        //   1 slot    2.15  2.14  2.27  2.41  2.56  2.89  3.61
        //   2 slots   2.08  1.97  2.02  2.02  2.15  2.26  2.55
        //   3 slots   2.16  1.98  1.96  1.96  2.05  2.13  2.33
        //   4 slots   2.17  2.10  1.98  1.95  1.94  2.04  2.20
        // Default: 4 slots, chunks of ~1 700 queries, at most 8 per tile; RBQ_FEED_CHUNKS / RBQ_FEED_SLOTS override (1 = off).
        const char* fe = getenv("RBQ_FEED_CHUNKS");  // read per call: tuning scripts sweep it inside one process
        const long auto_chunks = pl.slots > 1 ? std::min(8L, std::max(1L, (long)((n + 800) / 1700))) : std::min(4L, std::max(1L, (long)((n + 1700) / 3400)));
        const long forced = fe ? std::min(16L, std::max(1L, atol(fe))) : auto_chunks;
        feed.chunk = ((n + forced - 1) / forced + 127) / 128 * 128;
        if (n < 2048) feed.chunk = n;
        // RBQ_FEED_TAPER=1: the last chunks of a tile halve down to a few hundred queries (a shorter critical path after the last
        // copy, in theory).  Measured worse at GIST-1M / 10k queries (2.00 vs 1.95 ms per call, 4.31 vs 5.26 M QPS in the bench's
        // L2-flushed timing: the small chunks' kernels cost their fixed latency each), so it is off.
        const char* ft = getenv("RBQ_FEED_TAPER");
        feed.taper = ft ? atoi(ft) != 0 : false;
        const char* ff = getenv("RBQ_FEED_FIRST");  // tuning knob: queries in the first chunk
        feed.first = ff ? (size_t)std::max(128L, atol(ff) / 128 * 128) : 0;
        if (feed.first && feed.first < n && n >= 2048)  // the rest of the tile in `forced` equal chunks
            feed.chunk = std::max<size_t>(feed.first, ((n - feed.first + forced - 1) / forced + 127) / 128 * 128);
        // search_device copies queries [q0, q0+n) itself (feed) and indexes outputs from the tile start
        const bool trace = getenv("RBQ_TRACE") != nullptr;
        timespec t0, t1, t2;
        if (trace) clock_gettime(CLOCK_MONOTONIC, &t0);
        rc = search_device(h, nullptr, n, top_k, nprobe, d_f, filter_nbits, d_ids, d_sc, d_cn, (char*)h->ws, pl, st, &launches, &feed);
        if (rc) return rc;
        if (trace) clock_gettime(CLOCK_MONOTONIC, &t1);
        feed.h_q += n * dim;
        RBQ_CUDA(cudaMemcpyAsync(ids + q0 * top_k, d_ids, n * top_k * 8, cudaMemcpyDeviceToHost, st));
        RBQ_CUDA(cudaMemcpyAsync(scores + q0 * top_k, d_sc, n * top_k * 4, cudaMemcpyDeviceToHost, st));
        RBQ_CUDA(cudaMemcpyAsync(counts + q0, d_cn, n * 4, cudaMemcpyDeviceToHost, st));
        RBQ_CUDA(cudaStreamSynchronize(st));
        if (trace) {
            clock_gettime(CLOCK_MONOTONIC, &t2);
            fprintf(stderr, "[rbq trace] tile %zu queries: enqueue %.3f ms, total %.3f ms, chunk %zu\n", n,
                    (t1.tv_sec - t0.tv_sec) * 1e3 + (t1.tv_nsec - t0.tv_nsec) * 1e-6,
                    (t2.tv_sec - t0.tv_sec) * 1e3 + (t2.tv_nsec - t0.tv_nsec) * 1e-6, feed.chunk);
        }
    }
    h->last_stats.kernel_launches = launches;
    h->last_stats.coarse_mode_used = (uint32_t)pl.coarse;
    h->last_stats.coarse_terms_used = (uint32_t)pl.terms;
    h->last_stats.front_chunk = (uint32_t)pl.cq;
    return RBQ_OK;
}

int rbq_search_batch(const rbq_index* h, const float* queries, size_t nq, size_t dim, size_t top_k, size_t nprobe,
                     uint64_t* ids, float* scores, uint32_t* counts) {
    return rbq_search_batch_filtered(h, queries, nq, dim, top_k, nprobe, nullptr, 0, ids, scores, counts);
}

// ---- multi-GPU search in three phases (include/rbq.h) ------------------------------------------------
namespace {
int dist_args(const rbq_index* h, size_t nq, size_t dim_or_zero, size_t top_k, size_t* nprobe, Plan* pl) {
    int rc = check_search_args(h, dim_or_zero ? dim_or_zero : (size_t)h->dev.dim, top_k, nprobe);
    if (rc) return rc;
    if (top_k == 0 || nq == 0) return fail(RBQ_INVALID_CONFIG, "phased search needs nq > 0 and top_k > 0");
    *pl = make_plan(h, nq, *nprobe);
    if (pl->qt < nq) return fail(RBQ_INVALID_CONFIG, "phased search handles one tile of queries per batch (131072 queries, 2^26 probes): split the batch");
    return RBQ_OK;
}

// The three phases without locking / argument checks (the public entry points and the one-call sharded search share them).
int dist_front_impl(const rbq_index* h, const Plan& pl, const float* d_queries, size_t nq, size_t dim, size_t top_k, size_t nprobe, size_t q_begin,
                    size_t q_count, rbq_probe_rec* d_probes, cudaStream_t st) {
    int rc;
    h->last_stats = rbq_search_stats{};
    h->dist_phase = 0;
    h->last_stats.queries = nq;
    h->last_stats.coarse_mode_used = (uint32_t)pl.coarse;
    h->last_stats.coarse_terms_used = (uint32_t)pl.terms;
    h->last_stats.front_chunk = (uint32_t)pl.cq;
    if ((rc = ensure_ws(h, ws_need(h, pl, nprobe, top_k, dim, false, 0)))) return rc;
    RBQ_CUDA(cudaMemsetAsync(h->d_stats, 0, sizeof(DevStats) + 16, st));
    const DevIndex& ix = h->dev;
    const WsLayout L = carve_ws(h, (char*)h->ws, pl, nprobe, top_k);
    if (h->profiling) cudaEventRecord(h->ev[0], st);
    // every shard scans its lists for ALL queries, so it needs every query's rotation, LUT and scalars ...
    if ((rc = launch_query_prep(ix, d_queries, nq, L.d_rot, L.d_lut, L.d_qs, st))) return rc;
    if (h->profiling) cudaEventRecord(h->ev[1], st);
    uint64_t launches = 1;
    // ... but the probe lists are the same on every shard: each one computes a slice (in front-end chunks)
    for (size_t c0 = q_begin; c0 < q_begin + q_count; c0 += pl.cq) {
        const size_t m = std::min(pl.cq, q_begin + q_count - c0);
        if ((rc = run_front(h, L, pl, nullptr, c0, m, nprobe, st, &launches, false, nullptr, nullptr))) return rc;
    }
    if (q_count) {
        if ((rc = launch_probe_export(L.d_pr, q_begin, q_count, nprobe, d_probes, st))) return rc;
        launches += 1;
    }
    if (h->profiling) cudaEventRecord(h->ev[2], st);
    h->last_stats.kernel_launches = launches;
    return RBQ_OK;
}

int dist_head_impl(const rbq_index* h, const Plan& pl, size_t nq, size_t top_k, size_t nprobe, const rbq_probe_rec* d_probes, float* d_tau,
                   uint64_t* d_ids, float* d_scores, uint32_t* d_counts, cudaStream_t st) {
    int rc;
    if (h->ws_bytes < ws_need(h, pl, nprobe, top_k, h->dev.dim, false, 0)) return fail(RBQ_INVALID_CONFIG, "rbq_dist_front must run first");
    const DevIndex& ix = h->dev;
    const WsLayout L = carve_ws(h, (char*)h->ws, pl, nprobe, top_k);
    uint64_t launches = h->last_stats.kernel_launches;
    if (h->profiling) cudaEventRecord(h->ev[3], st);
    if ((rc = launch_probe_import(ix, d_probes, nq, nprobe, L.d_pr, L.d_head_owner, st))) return rc;
    RBQ_CUDA(cudaMemsetAsync(L.tw.surv_cnt, 0, (pl.qt + 2 * (size_t)ix.nlist + kTailCounters + kHeadCursors) * 4, st));
    int head_launch = 0;
    if ((rc = launch_head(ix, L.d_rot, L.d_lut, L.d_qs, L.d_pr, nq, nprobe, top_k, nullptr, 0, d_ids, d_scores, d_counts, h->d_stats, L.tw,
                          st, &launches, 0, nq, &head_launch, L.d_head_owner)))
        return rc;
    RBQ_CUDA(cudaMemcpyAsync(d_tau, L.tw.tau, nq * 4, cudaMemcpyDeviceToDevice, st));
    if (h->profiling) cudaEventRecord(h->ev[4], st);
    h->last_stats.kernel_launches = launches + 1;
    return RBQ_OK;
}

int dist_tail_impl(const rbq_index* h, const Plan& pl, size_t nq, size_t top_k, size_t nprobe, const float* d_tau, uint64_t* d_ids, float* d_scores,
                   uint32_t* d_counts, cudaStream_t st) {
    int rc;
    if (h->ws_bytes < ws_need(h, pl, nprobe, top_k, h->dev.dim, false, 0)) return fail(RBQ_INVALID_CONFIG, "rbq_dist_front must run first");
    const DevIndex& ix = h->dev;
    const WsLayout L = carve_ws(h, (char*)h->ws, pl, nprobe, top_k);
    uint64_t launches = h->last_stats.kernel_launches;
    if (h->profiling) cudaEventRecord(h->ev[5], st);
    RBQ_CUDA(cudaMemcpyAsync(L.tw.tau, d_tau, nq * 4, cudaMemcpyDeviceToDevice, st));
    // the head pass' fallback queries (this shard's): sequential walk on the side stream, beside the tail and replay kernels
    if ((rc = side_fork(h, st))) return rc;
    if ((rc = launch_scan(ix, L.d_rot, L.d_lut, L.d_qs, L.d_pr, nq, nprobe, top_k, nullptr, 0, d_ids, d_scores, d_counts, h->d_stats,
                          h->work_counter(), kScanFallback, &L.tw, h->side_stream)))
        return rc;
    if ((rc = launch_tail(ix, L.d_lut, L.d_qs, L.d_pr, nq, nprobe, nullptr, 0, h->d_stats, L.tw, st, &launches,
                          h->profiling ? h->ev[7] : nullptr, h->profiling ? h->ev[8] : nullptr)))
        return rc;
    if (h->profiling) cudaEventRecord(h->ev[9], st);
    if ((rc = launch_refine_replay(ix, L.d_rot, L.d_qs, L.d_pr, nq, nprobe, top_k, d_ids, d_scores, d_counts, h->d_stats, L.tw, st, &launches)))
        return rc;
    if ((rc = launch_scan(ix, L.d_rot, L.d_lut, L.d_qs, L.d_pr, nq, nprobe, top_k, nullptr, 0, d_ids, d_scores, d_counts, h->d_stats,
                          h->work_counter(), kScanFallbackResume, &L.tw, st)))
        return rc;
    if ((rc = side_join(h, st))) return rc;
    if (h->profiling) cudaEventRecord(h->ev[6], st);
    h->dist_phase = h->profiling ? 3 : 0;
    h->last_stats.kernel_launches = launches + 3;
    return RBQ_OK;
}
}  // namespace

int rbq_dist_front(const rbq_index* h, const float* d_queries, size_t nq, size_t dim, size_t top_k, size_t nprobe, size_t q_begin,
                   size_t q_count, rbq_probe_rec* d_probes, void* stream) {
    Plan pl;
    int rc = dist_args(h, nq, dim, top_k, &nprobe, &pl);
    if (rc) return rc;
    if (q_begin > nq || q_count > nq - q_begin) return fail(RBQ_INVALID_CONFIG, "query slice out of range");
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Serial serial(h, st);
    return dist_front_impl(h, pl, d_queries, nq, dim, top_k, nprobe, q_begin, q_count, d_probes, st);
}

int rbq_dist_head(const rbq_index* h, size_t nq, size_t top_k, size_t nprobe, const rbq_probe_rec* d_probes, float* d_tau,
                  uint64_t* d_ids, float* d_scores, uint32_t* d_counts, void* stream) {
    Plan pl;
    int rc = dist_args(h, nq, 0, top_k, &nprobe, &pl);
    if (rc) return rc;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Serial serial(h, st);
    return dist_head_impl(h, pl, nq, top_k, nprobe, d_probes, d_tau, d_ids, d_scores, d_counts, st);
}

int rbq_dist_tail(const rbq_index* h, size_t nq, size_t top_k, size_t nprobe, const float* d_tau, uint64_t* d_ids, float* d_scores,
                  uint32_t* d_counts, void* stream) {
    Plan pl;
    int rc = dist_args(h, nq, 0, top_k, &nprobe, &pl);
    if (rc) return rc;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Serial serial(h, st);
    return dist_tail_impl(h, pl, nq, top_k, nprobe, d_tau, d_ids, d_scores, d_counts, st);
}

// ---- multi-GPU search as ONE call: the three phases with their NCCL exchanges enqueued on one stream ---------------------
// No reference counterpart (src/ivf.rs is single-process).  librbq owns the communicator: rank 0 creates an id
// (rbq_comm_unique_id), the caller ships its 128 bytes to the other ranks by whatever means it has, every rank calls
// rbq_comm_init on its shard handle.  NCCL is loaded at run time (dlopen libnccl.so.2): single-GPU users do not need it.
namespace {
int nccl_check(int r, const char* what) {
    if (r == 0) return RBQ_OK;
    return fail(RBQ_CUDA_ERROR, std::string("NCCL error in ") + what + ": " + nccl_api().get_error_string(r));
}
size_t dist_slice(size_t nq, int world) { return (((nq + world - 1) / world) + 127) / 128 * 128; }  // padded slice length (GEMM row tile)
}  // namespace

int rbq_comm_unique_id(uint8_t* id_out) {
    if (!id_out) return fail(RBQ_INVALID_CONFIG, "null argument");
    int rc = nccl_load();
    if (rc) return rc;
    return nccl_check(nccl_api().get_unique_id(id_out), "ncclGetUniqueId");
}

int rbq_comm_init(rbq_index* h, const uint8_t* id, int rank, int world) {
    if (!h || !id) return fail(RBQ_INVALID_CONFIG, "null argument");
    if (rank != h->host.shard_rank || world != h->host.shard_count)
        return fail(RBQ_INVALID_CONFIG, "communicator rank/size must equal the handle's shard_rank/shard_count");
    int rc = nccl_load();
    if (rc) return rc;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->comm) return fail(RBQ_INVALID_CONFIG, "the handle already has a communicator");
    return nccl_check(nccl_api().comm_init_rank(&h->comm, world, id, rank), "ncclCommInitRank");
}

int rbq_set_exact_merge(rbq_index* h, int on) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    std::lock_guard<std::mutex> lk(h->mu);
    h->exact_merge = on != 0;
    return RBQ_OK;
}

int rbq_comm_destroy(rbq_index* h) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->busy_ev) cudaEventSynchronize(h->busy_ev);
    if (h->comm) {
        nccl_api().comm_destroy(h->comm);
        h->comm = nullptr;
    }
    if (h->dist_ws) cudaFree(h->dist_ws);
    h->dist_ws = nullptr;
    h->dist_ws_bytes = 0;
    if (h->xr_ws) cudaFree(h->xr_ws);
    if (h->xr_recs) cudaFree(h->xr_recs);
    if (h->xr_host) cudaFreeHost(h->xr_host);
    h->xr_ws = h->xr_recs = nullptr;
    h->xr_host = nullptr;
    h->xr_inexact = nullptr;
    h->xr_ws_bytes = h->xr_recs_bytes = 0;
    return RBQ_OK;
}

namespace {
// exact merge: fixed-size part of its workspace (a pure function of nq, top_k, world)
struct XrWs {
    uint64_t* snap_ids;
    float* snap_sc;
    uint32_t* snap_cn;
    uint32_t *all_cnt, *off, *cta_tot, *qoff, *flagged, *cursor;
    unsigned long long *send_tot, *recv_tot, *rbase, *inexact;
    char* gath2;
    size_t chunk2;
};
int xr_carve(const rbq_index* h, size_t nq, size_t top_k, size_t per, int world, XrWs* X) {
    X->chunk2 = (per * top_k * 12 + per * 4 + 15) / 16 * 16;
    for (int pass = 0; pass < 2; ++pass) {
        Carver cv{pass ? (char*)h->xr_ws : nullptr};
        X->snap_ids = cv.take<uint64_t>(nq * top_k);
        X->snap_sc = cv.take<float>(nq * top_k);
        X->snap_cn = cv.take<uint32_t>(nq);
        X->all_cnt = cv.take<uint32_t>(nq * world);
        X->off = cv.take<uint32_t>(nq + 1);
        X->cta_tot = cv.take<uint32_t>((nq + 1023) / 1024 + 1);
        X->qoff = cv.take<uint32_t>(per * world);
        X->flagged = cv.take<uint32_t>(per + 1);
        X->cursor = cv.take<uint32_t>(4);
        X->send_tot = cv.take<unsigned long long>(world);
        X->recv_tot = cv.take<unsigned long long>(world);
        X->rbase = cv.take<unsigned long long>(world);
        X->inexact = cv.take<unsigned long long>(2);
        X->gath2 = cv.take<char>(X->chunk2 * world);
        if (pass == 0) {
            const size_t need = cv.off + 256;
            if (h->xr_ws_bytes < need) {
                if (h->xr_ws) {
                    RBQ_CUDA(cudaDeviceSynchronize());
                    cudaFree(h->xr_ws);
                }
                h->xr_ws = nullptr;
                h->xr_ws_bytes = 0;
                RBQ_CUDA(cudaMalloc(&h->xr_ws, need));
                h->xr_ws_bytes = need;
            }
            if (!h->xr_host) RBQ_CUDA(cudaMallocHost((void**)&h->xr_host, (size_t)3 * kMaxShards * 8));
        }
    }
    return RBQ_OK;
}

// the caller holds the handle's lock and has validated the arguments; extra_bytes of the exchange workspace are reserved in front
// (host entry: staging of the queries and of the merged result)
int sharded_search_impl(const rbq_index* h, const Plan& pl, const float* d_queries, size_t nq, size_t dim, size_t top_k, size_t nprobe,
                        uint64_t* d_ids, float* d_scores, uint32_t* d_counts, cudaStream_t st, size_t extra_bytes, char** extra_out) {
    int rc;
    const int world = h->host.shard_count, rank = h->host.shard_rank;
    // exchange buffers: probe records of all (padded) slices | thresholds | every shard's packed local top-k
    const size_t per = dist_slice(nq, world);
    const size_t chunk = (nq * top_k * 12 + nq * 4 + 15) / 16 * 16;
    const size_t need = extra_bytes + 256 + per * world * nprobe * sizeof(rbq_probe_rec) + 256 + nq * 4 + 256 + chunk * world + 256;
    if (h->dist_ws_bytes < need) {
        if (h->dist_ws) {
            RBQ_CUDA(cudaDeviceSynchronize());
            cudaFree(h->dist_ws);
        }
        h->dist_ws = nullptr;
        h->dist_ws_bytes = 0;
        RBQ_CUDA(cudaMalloc(&h->dist_ws, need));
        h->dist_ws_bytes = need;
    }
    Carver cv{(char*)h->dist_ws};
    char* extra = cv.take<char>(extra_bytes);
    if (extra_out) *extra_out = extra;
    if (d_queries == nullptr) return RBQ_OK;  // sizing / staging pass of the host entry
    rbq_probe_rec* d_rec = cv.take<rbq_probe_rec>(per * world * nprobe);
    float* d_tau = cv.take<float>(nq);
    char* d_gath = cv.take<char>(chunk * world);
    char* mine = d_gath + (size_t)rank * chunk;
    uint64_t* l_ids = reinterpret_cast<uint64_t*>(mine);
    float* l_sc = reinterpret_cast<float*>(mine + nq * top_k * 8);
    uint32_t* l_cn = reinterpret_cast<uint32_t*>(mine + nq * top_k * 12);
    const NcclApi& nc = nccl_api();
    const size_t q_begin = std::min((size_t)rank * per, nq), q_count = std::min((size_t)(rank + 1) * per, nq) - q_begin;
    const bool exact = h->exact_merge != 0;
    XrWs X{};
    h->xr_inexact = nullptr;
    if (exact) {
        if (nq >= ((size_t)1 << 31) / std::max<size_t>(top_k, 1)) return fail(RBQ_INVALID_CONFIG, "exact merge: batch too large");
        if ((rc = xr_carve(h, nq, top_k, per, world, &X))) return rc;
    }
    // 1. front end for this rank's slice; the slices (16 B per probe) are all-gathered in place
    if ((rc = dist_front_impl(h, pl, d_queries, nq, dim, top_k, nprobe, q_begin, q_count, d_rec, st))) return rc;
    const size_t slice_bytes = per * nprobe * sizeof(rbq_probe_rec);
    if ((rc = nccl_check(nc.all_gather((const char*)d_rec + (size_t)rank * slice_bytes, d_rec, slice_bytes, /*ncclChar*/ 0, h->comm, st), "ncclAllGather")))
        return rc;
    // 2. head pass where this shard owns the query's nearest list; thresholds MIN-reduced
    if ((rc = dist_head_impl(h, pl, nq, top_k, nprobe, d_rec, d_tau, l_ids, l_sc, l_cn, st))) return rc;
    if ((rc = nccl_check(nc.all_reduce(d_tau, d_tau, nq, /*ncclFloat32*/ 7, /*ncclMin*/ 3, h->comm, st), "ncclAllReduce"))) return rc;
    if (exact) {  // the heap after the head pass (the tail's replay overwrites it): travels to the home rank with the survivors
        RBQ_CUDA(cudaMemcpyAsync(X.snap_ids, l_ids, nq * top_k * 8, cudaMemcpyDeviceToDevice, st));
        RBQ_CUDA(cudaMemcpyAsync(X.snap_sc, l_sc, nq * top_k * 4, cudaMemcpyDeviceToDevice, st));
        RBQ_CUDA(cudaMemcpyAsync(X.snap_cn, l_cn, nq * 4, cudaMemcpyDeviceToDevice, st));
    }
    // 3. tail + replay on this shard's lists; the packed local top-k (one chunk per rank) all-gathered in place and merged
    if ((rc = dist_tail_impl(h, pl, nq, top_k, nprobe, d_tau, l_ids, l_sc, l_cn, st))) return rc;
    if ((rc = nccl_check(nc.all_gather(mine, d_gath, chunk, /*ncclChar*/ 0, h->comm, st), "ncclAllGather"))) return rc;
    rc = launch_merge(h->host.metric, world, nq, top_k, reinterpret_cast<const uint64_t*>(d_gath), reinterpret_cast<const float*>(d_gath + nq * top_k * 8),
                      reinterpret_cast<const uint32_t*>(d_gath + nq * top_k * 12), d_ids, d_scores, d_counts, st, chunk / 8, chunk / 4, chunk / 4);
    h->last_stats.kernel_launches += 1;
    if (rc || !exact) return rc;

    // 4. exact merge (exact_merge.cu): every survivor refined, records shipped to the query's home rank, one global replay there
    const WsLayout L = carve_ws(h, (char*)h->ws, pl, nprobe, top_k);
    if ((rc = xr_launch_count_scan(L.tw, L.d_head_owner, X.snap_cn, nq, (uint32_t)per, (uint32_t)world, X.all_cnt + (size_t)rank * nq, X.off, X.cta_tot,
                                   X.send_tot, st)))
        return rc;
    if ((rc = nccl_check(nc.all_gather(X.all_cnt + (size_t)rank * nq, X.all_cnt, nq * 4, /*ncclChar*/ 0, h->comm, st), "ncclAllGather"))) return rc;
    if ((rc = xr_launch_plan(X.all_cnt, nq, q_begin, q_count, world, X.qoff, X.recv_tot, X.flagged, st))) return rc;
    unsigned long long* hs = h->xr_host;  // [0, world): records to send per home rank, [world, 2 world): to receive per source, then bases
    RBQ_CUDA(cudaMemcpyAsync(hs, X.send_tot, (size_t)world * 8, cudaMemcpyDeviceToHost, st));
    RBQ_CUDA(cudaMemcpyAsync(hs + world, X.recv_tot, (size_t)world * 8, cudaMemcpyDeviceToHost, st));
    RBQ_CUDA(cudaStreamSynchronize(st));  // the one host synchronisation of the call: message sizes
    unsigned long long n_send = 0, n_recv = 0;
    std::vector<unsigned long long> sbase(world);
    for (int d = 0; d < world; ++d) {
        sbase[d] = n_send;
        n_send += hs[d];
        hs[2 * world + d] = n_recv;
        n_recv += hs[world + d];
    }
    const size_t rec_need = (size_t)(n_send + n_recv + 2) * sizeof(XrRec);
    if (h->xr_recs_bytes < rec_need) {
        if (h->xr_recs) cudaFree(h->xr_recs);
        h->xr_recs = nullptr;
        h->xr_recs_bytes = 0;
        const size_t grow = rec_need + rec_need / 4;
        RBQ_CUDA(cudaMalloc(&h->xr_recs, grow));
        h->xr_recs_bytes = grow;
    }
    XrRec* d_send = reinterpret_cast<XrRec*>(h->xr_recs);
    XrRec* d_recv = d_send + n_send + 1;
    RBQ_CUDA(cudaMemcpyAsync(X.rbase, hs + 2 * world, (size_t)world * 8, cudaMemcpyHostToDevice, st));
    if ((rc = xr_launch_records(h->dev, L.d_rot, L.d_qs, L.d_pr, nq, nprobe, top_k, L.tw, L.d_head_owner, X.snap_ids, X.snap_sc, X.snap_cn,
                                X.all_cnt + (size_t)rank * nq, X.off, d_send, X.cursor, st)))
        return rc;
    if ((rc = nccl_check(nc.group_start(), "ncclGroupStart"))) return rc;
    for (int d = 0; d < world; ++d) {
        if (hs[d]) nc.send(d_send + sbase[d], (size_t)hs[d] * sizeof(XrRec), /*ncclChar*/ 0, d, h->comm, st);
        if (hs[world + d]) nc.recv(d_recv + hs[2 * world + d], (size_t)hs[world + d] * sizeof(XrRec), /*ncclChar*/ 0, d, h->comm, st);
    }
    if ((rc = nccl_check(nc.group_end(), "ncclGroupEnd"))) return rc;
    RBQ_CUDA(cudaMemsetAsync(X.inexact, 0, 8, st));
    char* home = X.gath2 + (size_t)rank * X.chunk2;
    if ((rc = xr_launch_replay(d_recv, X.rbase, X.all_cnt, X.qoff, X.flagged, nq, q_begin, q_count, world, top_k, h->host.metric,
                               reinterpret_cast<uint64_t*>(home), reinterpret_cast<float*>(home + per * top_k * 8),
                               reinterpret_cast<uint32_t*>(home + per * top_k * 12), X.inexact, st)))
        return rc;
    if ((rc = nccl_check(nc.all_gather(home, X.gath2, X.chunk2, /*ncclChar*/ 0, h->comm, st), "ncclAllGather"))) return rc;
    if ((rc = xr_launch_override(X.gath2, X.chunk2, per, nq, top_k, d_ids, d_scores, d_counts, st))) return rc;
    h->xr_inexact = X.inexact;
    h->last_stats.exchanged_records = n_recv;
    h->last_stats.kernel_launches += 9;
    return RBQ_OK;
}
}  // namespace


int rbq_search_batch_sharded_device(const rbq_index* h, const float* d_queries, size_t nq, size_t dim, size_t top_k, size_t nprobe,
                                    uint64_t* d_ids, float* d_scores, uint32_t* d_counts, void* stream) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    if (!h->comm) return fail(RBQ_INVALID_CONFIG, "rbq_comm_init must be called on the shard handle first");
    if (!d_queries) return fail(RBQ_INVALID_CONFIG, "null buffer");
    Plan pl;
    int rc = dist_args(h, nq, dim, top_k, &nprobe, &pl);
    if (rc) return rc;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    Serial serial(h, st);
    return sharded_search_impl(h, pl, d_queries, nq, dim, top_k, nprobe, d_ids, d_scores, d_counts, st, 0, nullptr);
}

// Host buffers: every rank passes the SAME batch; each uploads only its 1/world slice over its own host link and the slices
// are all-gathered over NVLink, the merged result comes back to every rank's host buffers.  Synchronous.
int rbq_search_batch_sharded(const rbq_index* h, const float* queries, size_t nq, size_t dim, size_t top_k, size_t nprobe, uint64_t* ids,
                             float* scores, uint32_t* counts) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    if (!h->comm) return fail(RBQ_INVALID_CONFIG, "rbq_comm_init must be called on the shard handle first");
    if (!queries || !ids || !scores || !counts) return fail(RBQ_INVALID_CONFIG, "null buffer");
    Plan pl;
    int rc = dist_args(h, nq, dim, top_k, &nprobe, &pl);
    if (rc) return rc;
    const int world = h->host.shard_count, rank = h->host.shard_rank;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    if (!h->compute_stream) {
        RBQ_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        RBQ_CUDA(cudaStreamCreateWithFlags(&h->compute_stream, cudaStreamNonBlocking));
        for (auto& e : h->feed_ev) RBQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t st = h->compute_stream;
    Serial serial(h, st);
    const size_t per = (nq + world - 1) / world;
    const size_t q_bytes = per * world * dim * 4, out_bytes = nq * top_k * 12 + nq * 4 + 64;
    char* extra = nullptr;
    if ((rc = sharded_search_impl(h, pl, nullptr, nq, dim, top_k, nprobe, nullptr, nullptr, nullptr, st, q_bytes + 256 + out_bytes, &extra))) return rc;
    float* d_q = reinterpret_cast<float*>(extra);
    char* d_out = extra + ((q_bytes + 255) & ~(size_t)255);
    uint64_t* d_ids = reinterpret_cast<uint64_t*>(d_out);
    float* d_sc = reinterpret_cast<float*>(d_out + nq * top_k * 8);
    uint32_t* d_cn = reinterpret_cast<uint32_t*>(d_out + nq * top_k * 12);
    const size_t lo = std::min((size_t)rank * per, nq), hi = std::min((size_t)(rank + 1) * per, nq);
    if (hi > lo) RBQ_CUDA(cudaMemcpyAsync(d_q + lo * dim, queries + lo * dim, (hi - lo) * dim * 4, cudaMemcpyHostToDevice, st));
    if ((rc = nccl_check(nccl_api().all_gather(d_q + (size_t)rank * per * dim, d_q, per * dim * 4, /*ncclChar*/ 0, h->comm, st), "ncclAllGather")))
        return rc;
    if ((rc = sharded_search_impl(h, pl, d_q, nq, dim, top_k, nprobe, d_ids, d_sc, d_cn, st, q_bytes + 256 + out_bytes, nullptr))) return rc;
    RBQ_CUDA(cudaMemcpyAsync(ids, d_ids, nq * top_k * 8, cudaMemcpyDeviceToHost, st));
    RBQ_CUDA(cudaMemcpyAsync(scores, d_sc, nq * top_k * 4, cudaMemcpyDeviceToHost, st));
    RBQ_CUDA(cudaMemcpyAsync(counts, d_cn, nq * 4, cudaMemcpyDeviceToHost, st));
    RBQ_CUDA(cudaStreamSynchronize(st));
    return RBQ_OK;
}

int rbq_fetch_embedding(const rbq_index* h, uint64_t vector_id, float* out, int* found) {
    if (!h || !out || !found) return fail(RBQ_INVALID_CONFIG, "null argument");
    *found = 0;
    const HostIndex& hi = h->host;
    const size_t nvec = hi.vec_off.empty() ? 0 : (size_t)hi.vec_off.back();
    if (nvec == 0) return RBQ_OK;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->busy_ev) RBQ_CUDA(cudaEventSynchronize(h->busy_ev));  // the workspace may still be in use by an asynchronous search
    int rc = ensure_ws(h, 4096 + (size_t)h->dev.dim * 4);
    if (rc) return rc;
    unsigned long long* d_pos = reinterpret_cast<unsigned long long*>(h->ws);
    float* d_out = reinterpret_cast<float*>((char*)h->ws + 256);
    if ((rc = launch_find_id(h->dev, nvec, vector_id, d_pos, nullptr))) return rc;
    unsigned long long pos = ~0ull;
    RBQ_CUDA(cudaMemcpy(&pos, d_pos, 8, cudaMemcpyDeviceToHost));
    if (pos >= nvec) return RBQ_OK;  // None
    // the list that holds the position, and the reconstruction scalars (kept on the host: search never reads them)
    const size_t c = (size_t)(std::upper_bound(hi.vec_off.begin(), hi.vec_off.end(), (uint64_t)pos) - hi.vec_off.begin()) - 1;
    if (hi.delta.size() != nvec || hi.vl.size() != nvec) return fail(RBQ_INVALID_CONFIG, "reconstruction factors are not resident");
    if ((rc = launch_fetch_embedding(h->dev, (uint32_t)c, (uint32_t)(pos - hi.vec_off[c]), hi.delta[pos], hi.vl[pos], d_out, nullptr))) return rc;
    RBQ_CUDA(cudaMemcpy(out, d_out, (size_t)h->dev.dim * 4, cudaMemcpyDeviceToHost));
    *found = 1;
    return RBQ_OK;
}

int rbq_merge_topk_device(const rbq_index* h, int nshards, size_t nq, size_t top_k, const uint64_t* in_ids,
                          const float* in_scores, const uint32_t* in_counts, uint64_t* out_ids, float* out_scores,
                          uint32_t* out_counts, void* stream) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    DeviceGuard g(h->device);
    return launch_merge(h->host.metric, nshards, nq, top_k, in_ids, in_scores, in_counts, out_ids, out_scores,
                        out_counts, reinterpret_cast<cudaStream_t>(stream));
}

int rbq_merge_topk_packed_device(const rbq_index* h, int nshards, size_t nq, size_t top_k, const void* packed, size_t chunk_bytes,
                                 uint64_t* out_ids, float* out_scores, uint32_t* out_counts, void* stream) {
    if (!h) return fail(RBQ_INVALID_CONFIG, "null index handle");
    if (chunk_bytes % 8 != 0 || chunk_bytes < nq * top_k * 12 + nq * 4) return fail(RBQ_INVALID_CONFIG, "packed top-k chunk too small or misaligned");
    DeviceGuard g(h->device);
    const char* base = static_cast<const char*>(packed);
    return launch_merge(h->host.metric, nshards, nq, top_k, reinterpret_cast<const uint64_t*>(base),
                        reinterpret_cast<const float*>(base + nq * top_k * 8), reinterpret_cast<const uint32_t*>(base + nq * top_k * 12),
                        out_ids, out_scores, out_counts, reinterpret_cast<cudaStream_t>(stream), chunk_bytes / 8, chunk_bytes / 4,
                        chunk_bytes / 4);
}

// ---- stage probes ------------------------------------------------------------------------------
int rbq_debug_query_prep(const rbq_index* h, const float* queries, size_t nq, size_t dim, float* rotated, uint8_t* lut,
                         float* scalars) {
    size_t np = 1;
    int rc = check_search_args(h, dim, 1, &np);
    if (rc) return rc;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    const size_t D = h->dev.D;
    if ((rc = ensure_ws(h, nq * (dim * 4 + D * 8 + 32) + 4096))) return rc;
    Carver cv{(char*)h->ws};
    float* d_q = cv.take<float>(nq * dim);
    float* d_rot = cv.take<float>(nq * D);
    uint8_t* d_lut = cv.take<uint8_t>(nq * D * 4);
    QueryScalars* d_qs = cv.take<QueryScalars>(nq);
    RBQ_CUDA(cudaMemcpy(d_q, queries, nq * dim * 4, cudaMemcpyHostToDevice));
    if ((rc = launch_query_prep(h->dev, d_q, nq, d_rot, d_lut, d_qs, nullptr))) return rc;
    if (rotated) RBQ_CUDA(cudaMemcpy(rotated, d_rot, nq * D * 4, cudaMemcpyDeviceToHost));
    if (lut) RBQ_CUDA(cudaMemcpy(lut, d_lut, nq * D * 4, cudaMemcpyDeviceToHost));
    if (scalars) RBQ_CUDA(cudaMemcpy(scalars, d_qs, nq * sizeof(QueryScalars), cudaMemcpyDeviceToHost));
    RBQ_CUDA(cudaDeviceSynchronize());
    return RBQ_OK;
}

int rbq_debug_probe(const rbq_index* h, const float* queries, size_t nq, size_t dim, size_t nprobe, uint32_t* probe_cids,
                    float* probe_consts) {
    int rc = check_search_args(h, dim, 1, &nprobe);
    if (rc) return rc;
    if (nq == 0) return RBQ_OK;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->busy_ev) RBQ_CUDA(cudaEventSynchronize(h->busy_ev));
    // the product's own front end (run_front), whatever coarse mode the handle is in
    const Plan pl = make_plan(h, nq, nprobe);
    const size_t base = ws_need(h, pl, nprobe, 1, dim, false, 0);
    if ((rc = ensure_ws(h, base + pl.qt * dim * 4 + 4096))) return rc;
    const WsLayout L = carve_ws(h, (char*)h->ws, pl, nprobe, 1);
    Carver cv{(char*)h->ws};
    cv.off = base;
    float* d_q = cv.take<float>(pl.qt * dim);
    RBQ_CUDA(cudaMemset(h->fallback_counter(), 0, 4));
    std::vector<Probe> pr(pl.qt * nprobe);
    uint64_t launches = 0;
    for (size_t q0 = 0; q0 < nq; q0 += pl.qt) {
        const size_t n = std::min(pl.qt, nq - q0);
        RBQ_CUDA(cudaMemcpy(d_q, queries + q0 * dim, n * dim * 4, cudaMemcpyHostToDevice));
        for (size_t c0 = 0; c0 < n; c0 += pl.cq) {
            const size_t m = std::min(pl.cq, n - c0);
            if ((rc = run_front(h, L, pl, d_q + c0 * dim, c0, m, nprobe, nullptr, &launches, true, nullptr, nullptr, true))) return rc;
        }
        RBQ_CUDA(cudaMemcpy(pr.data(), L.d_pr, n * nprobe * sizeof(Probe), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n * nprobe; ++i) {
            const size_t o = q0 * nprobe + i;
            probe_cids[o] = pr[i].cid;
            probe_consts[3 * o] = pr[i].g_add;
            probe_consts[3 * o + 1] = pr[i].g_error;
            probe_consts[3 * o + 2] = pr[i].dot_qc;
        }
    }
    h->last_stats = rbq_search_stats{};
    h->last_stats.queries = nq;
    h->last_stats.coarse_mode_used = (uint32_t)pl.coarse;
    h->last_stats.coarse_terms_used = (uint32_t)pl.terms;
    h->last_stats.front_chunk = (uint32_t)pl.cq;
    return RBQ_OK;
}

// Stage probes through the PRODUCT kernels of the list-major schedule (the kernels a search launches, not a debug twin).
// which == 0: head_scan_kernel -- the dense (lower bound, ip | estimate) rows of every query's nearest list:
//             out_a[q*cap + i] = lower bound, out_b[q*cap + i] = ip (ex_bits > 0) or estimate, out_n[q] = list length.
// which == 1: the tail FastScan kernel (tail_tc_kernel) over ALL probed lists with the threshold at +inf, so every vector
//             survives: records (rank, position, lower bound, ip | estimate) in arbitrary order, out_n[q] of them per query
//             (cap must hold them; <= 1024).  out_a / out_b / out_r / out_p are nq*cap arrays.
int rbq_debug_stage(const rbq_index* h, int which, const float* queries, size_t nq, size_t dim, size_t nprobe, size_t cap, float* out_a,
                    float* out_b, uint32_t* out_r, uint32_t* out_p, uint32_t* out_n) {
    int rc = check_search_args(h, dim, 1, &nprobe);
    if (rc) return rc;
    if (nq == 0) return RBQ_OK;
    if (which != 0 && which != 1) return fail(RBQ_INVALID_CONFIG, "stage must be 0 (head scan) or 1 (tail scan)");
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->busy_ev) RBQ_CUDA(cudaEventSynchronize(h->busy_ev));
    const size_t top_k = 1;
    const Plan pl = make_plan(h, nq, nprobe);
    if (pl.qt < nq) return fail(RBQ_INVALID_CONFIG, "too many queries for a stage probe");
    const size_t base = ws_need(h, pl, nprobe, top_k, dim, false, 0);
    if ((rc = ensure_ws(h, base + nq * dim * 4 + nq * 16 + 4096))) return rc;
    const WsLayout L = carve_ws(h, (char*)h->ws, pl, nprobe, top_k);
    const DevIndex& ix = h->dev;
    Carver cv{(char*)h->ws};
    cv.off = base;
    float* d_q = cv.take<float>(nq * dim);
    uint64_t* d_ids = cv.take<uint64_t>(nq);
    float* d_sc = cv.take<float>(nq);
    uint32_t* d_cn = cv.take<uint32_t>(nq);
    RBQ_CUDA(cudaMemcpy(d_q, queries, nq * dim * 4, cudaMemcpyHostToDevice));
    uint64_t launches = 0;
    for (size_t c0 = 0; c0 < nq; c0 += pl.cq) {
        const size_t m = std::min(pl.cq, nq - c0);
        if ((rc = run_front(h, L, pl, d_q + c0 * dim, c0, m, nprobe, nullptr, &launches, true, nullptr, nullptr, true))) return rc;
    }
    RBQ_CUDA(cudaMemset(L.tw.surv_cnt, 0, (pl.qt + 2 * (size_t)ix.nlist + kTailCounters + kHeadCursors) * 4));
    std::vector<Probe> pr(nq * nprobe);
    RBQ_CUDA(cudaMemcpy(pr.data(), L.d_pr, pr.size() * sizeof(Probe), cudaMemcpyDeviceToHost));
    if (which == 0) {
        if (nq > L.tw.head_rows) return fail(RBQ_INVALID_CONFIG, "too many queries for one head sub-chunk");
        int hl = 0;
        if ((rc = launch_head(ix, L.d_rot, L.d_lut, L.d_qs, L.d_pr, nq, nprobe, top_k, nullptr, 0, d_ids, d_sc, d_cn, h->d_stats, L.tw, nullptr, &launches, 0,
                              nq, &hl)))
            return rc;
        std::vector<float2> row(L.tw.head_cap);
        for (size_t q = 0; q < nq; ++q) {
            uint32_t nv = 0;
            for (size_t r = 0; r < nprobe; ++r)
                if (pr[q * nprobe + r].nv) {
                    nv = pr[q * nprobe + r].nv;
                    break;
                }
            out_n[q] = nv;
            if (nv > cap || nv > L.tw.head_cap) return fail(RBQ_INVALID_CONFIG, "output buffers too small for the head list");
            RBQ_CUDA(cudaMemcpy(row.data(), L.tw.head_buf + q * (size_t)L.tw.head_cap, (size_t)nv * 8, cudaMemcpyDeviceToHost));
            for (uint32_t i = 0; i < nv; ++i) {
                out_a[q * cap + i] = row[i].x;
                out_b[q * cap + i] = row[i].y;
            }
        }
        return RBQ_OK;
    }
    // tail over every pair, threshold +inf
    if (cap > L.tw.surv_cap) return fail(RBQ_INVALID_CONFIG, "cap exceeds the survivor buffer");
    if ((rc = launch_fill_u32(L.tw.tail_start, nq, 0u, nullptr))) return rc;
    if ((rc = launch_fill_u32(reinterpret_cast<uint32_t*>(L.tw.tau), nq, 0x7f800000u, nullptr))) return rc;
    if ((rc = launch_tail(ix, L.d_lut, L.d_qs, L.d_pr, nq, nprobe, nullptr, 0, h->d_stats, L.tw, nullptr, &launches))) return rc;
    std::vector<uint32_t> cnt(nq);
    RBQ_CUDA(cudaMemcpy(cnt.data(), L.tw.surv_cnt, nq * 4, cudaMemcpyDeviceToHost));
    std::vector<Survivor> sv(L.tw.surv_cap);
    for (size_t q = 0; q < nq; ++q) {
        out_n[q] = cnt[q];
        if (cnt[q] > cap) return fail(RBQ_INVALID_CONFIG, "more survivors than the output buffers hold");
        RBQ_CUDA(cudaMemcpy(sv.data(), L.tw.surv + q * (size_t)L.tw.surv_cap, (size_t)cnt[q] * sizeof(Survivor), cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < cnt[q]; ++i) {
            out_r[q * cap + i] = sv[i].rank;
            out_p[q * cap + i] = sv[i].pos;
            out_a[q * cap + i] = sv[i].lower;
            out_b[q * cap + i] = sv[i].x;
        }
    }
    return RBQ_OK;
}

// Stage probe for K10 (ip_packed_ex2_f32 / ip_packed_ex6_f32, src/simd.rs:1722-1825): the ex-code dot of the first n vectors of
// `cluster` against one query, through the product's refine path.
int rbq_debug_ex_dot(const rbq_index* h, const float* query, size_t dim, size_t cluster, size_t n, float* out) {
    size_t np = 1;
    int rc = check_search_args(h, dim, 1, &np);
    if (rc) return rc;
    if (cluster >= h->host.nlist || n > h->host.list_n[cluster]) return fail(RBQ_INVALID_CONFIG, "cluster / count out of range");
    if (n == 0 || h->dev.ex_bits == 0) return RBQ_OK;
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    if (h->busy_ev) RBQ_CUDA(cudaEventSynchronize(h->busy_ev));
    const size_t D = h->dev.D;
    if ((rc = ensure_ws(h, dim * 4 + D * 8 + 64 + n * 12 + 16384))) return rc;
    Carver cv{(char*)h->ws};
    float* d_q = cv.take<float>(dim);
    float* d_rot = cv.take<float>(D);
    uint8_t* d_lut = cv.take<uint8_t>(D * 4);
    QueryScalars* d_qs = cv.take<QueryScalars>(1);
    unsigned long long* d_gv = cv.take<unsigned long long>(n);
    float* d_out = cv.take<float>(n);
    std::vector<unsigned long long> gv(n);
    for (size_t i = 0; i < n; ++i) gv[i] = h->host.vec_off[cluster] + i;
    RBQ_CUDA(cudaMemcpy(d_q, query, dim * 4, cudaMemcpyHostToDevice));
    RBQ_CUDA(cudaMemcpy(d_gv, gv.data(), n * 8, cudaMemcpyHostToDevice));
    if ((rc = launch_query_prep(h->dev, d_q, 1, d_rot, d_lut, d_qs, nullptr))) return rc;
    TailWs tw{};
    if ((rc = launch_ex_dot_debug(h->dev, d_rot, d_gv, (int)n, d_out, tw, nullptr))) return rc;
    RBQ_CUDA(cudaMemcpy(out, d_out, n * 4, cudaMemcpyDeviceToHost));
    return RBQ_OK;
}

int rbq_debug_scan_list(const rbq_index* h, const float* query, size_t dim, size_t cluster, uint32_t* accu, float* ip,
                        float* est, float* lb, size_t cap_vectors) {
    size_t np = 1;
    int rc = check_search_args(h, dim, 1, &np);
    if (rc) return rc;
    if (cluster >= h->host.nlist) return fail(RBQ_INVALID_CONFIG, "cluster out of range");
    DeviceGuard g(h->device);
    std::lock_guard<std::mutex> lk(h->mu);
    const size_t D = h->dev.D, nl = h->dev.nlist;
    const size_t nv = h->host.list_n[cluster], nb = (nv + kBatch - 1) / kBatch, slots = nb * kBatch;
    if (cap_vectors < slots) return fail(RBQ_INVALID_CONFIG, "output buffers too small");
    if ((rc = ensure_ws(h, dim * 4 + D * 8 + 64 + nl * 4 + nl * sizeof(Probe) + slots * 16 + 16384))) return rc;
    Carver cv{(char*)h->ws};
    float* d_q = cv.take<float>(dim);
    float* d_rot = cv.take<float>(D);
    uint8_t* d_lut = cv.take<uint8_t>(D * 4);
    QueryScalars* d_qs = cv.take<QueryScalars>(1);
    float* d_sc = cv.take<float>(nl);
    Probe* d_pr = cv.take<Probe>(nl);
    uint32_t* d_accu = cv.take<uint32_t>(slots + 1);
    float* d_ip = cv.take<float>(slots + 1);
    float* d_est = cv.take<float>(slots + 1);
    float* d_lb = cv.take<float>(slots + 1);
    RBQ_CUDA(cudaMemcpy(d_q, query, dim * 4, cudaMemcpyHostToDevice));
    if ((rc = launch_query_prep(h->dev, d_q, 1, d_rot, d_lut, d_qs, nullptr))) return rc;
    // constants of this list in the reference's float order: reuse the probe kernel over all lists
    if (nl > probe_select_max_nprobe()) return fail(RBQ_INVALID_CONFIG, "debug probe needs nlist <= 4096");
    if ((rc = launch_coarse_exact(h->dev, d_rot, 1, d_sc, nullptr))) return rc;
    if ((rc = launch_probe_select(h->dev, d_rot, d_sc, 1, nl, d_pr, nullptr))) return rc;
    std::vector<Probe> pr(nl);
    RBQ_CUDA(cudaMemcpy(pr.data(), d_pr, nl * sizeof(Probe), cudaMemcpyDeviceToHost));
    float g_add = 0, g_error = 0;
    for (auto& p : pr)
        if (p.cid == cluster) {
            g_add = p.g_add;
            g_error = p.g_error;
        }
    if (slots == 0) return RBQ_OK;
    if ((rc = launch_scan_debug(h->dev, d_lut, d_qs, (uint32_t)cluster, g_add, g_error, d_accu, d_ip, d_est, d_lb, nullptr)))
        return rc;
    RBQ_CUDA(cudaMemcpy(accu, d_accu, slots * 4, cudaMemcpyDeviceToHost));
    RBQ_CUDA(cudaMemcpy(ip, d_ip, slots * 4, cudaMemcpyDeviceToHost));
    RBQ_CUDA(cudaMemcpy(est, d_est, slots * 4, cudaMemcpyDeviceToHost));
    RBQ_CUDA(cudaMemcpy(lb, d_lb, slots * 4, cudaMemcpyDeviceToHost));
    return RBQ_OK;
}

}  // extern "C"
