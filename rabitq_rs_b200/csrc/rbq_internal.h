// rbq_internal.h -- shared host/device declarations of librbq (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/rbq.h"

struct rbq_index;

namespace rbq {

constexpr int kBatch = 32;  // FASTSCAN_BATCH_SIZE (reference src/simd.rs:768)
// suspend-time hint of mbarrier.try_wait (ns): the hardware may park the waiting thread until the phase completes or the hint
// expires, so a waiting warp re-issues the instruction far less often.  Without a hint a spinning warp (the MMA issuers of the
// tail kernel spend most of their time waiting) costs ~3 issue slots every ~60 clk: 19 % of that kernel's instructions.
constexpr uint32_t kMbarSuspendHintNs = 0x989680u;
constexpr int kMaxShards = 32;  // list_owner is one byte per list; the device merge walks one sorted list per lane

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
#define RBQ_CUDA(call)                                                                           \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return ::rbq::fail(RBQ_CUDA_ERROR, std::string("CUDA error: ") + cudaGetErrorString(e_) + \
                                                   " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

// ---- host image of an index (what the RBQ1 v3 stream holds, lists concatenated) -------------
struct HostIndex {
    uint32_t dim = 0, D = 0;
    int metric = 0, rot_type = 1, ex_bits = 0;
    std::vector<uint8_t> rot_bytes;  // FHT: 4*D/8 flip bytes; Matrix: D*D f32 row-major
    size_t nlist = 0;
    uint64_t nvec_total = 0;         // vectors in the whole index (all shards)
    int shard_rank = 0, shard_count = 1;
    std::vector<float> centroids;    // nlist*D (rotated space)
    std::vector<uint32_t> list_n_all;  // size of every list (owned or not)
    std::vector<uint8_t> list_owner;   // shard that owns every list (empty: single shard)
    std::vector<uint32_t> list_n;      // size of every list on this shard (0 if not owned)
    std::vector<uint32_t> blk_off;     // nlist+1, in 32-vector blocks, owned lists only
    std::vector<uint64_t> vec_off;     // nlist+1
    std::vector<uint8_t> blocks;       // owned blocks, each 4D+384 bytes
    std::vector<uint64_t> ids;
    std::vector<uint8_t> ex;           // owned vectors * (D*ex_bits/8)
    std::vector<float> f_add_ex, f_rescale_ex, delta, vl;
    size_t block_stride() const { return (size_t)D * 4 + 384; }
    size_t ex_stride() const { return ex_bits > 0 ? (size_t)D * ex_bits / 8 : 0; }
};

// format.cc
uint32_t crc32_ieee(uint32_t crc, const uint8_t* p, size_t n);
int parse_rbq1(const uint8_t* p, size_t n, int shard_rank, int shard_count, HostIndex& out);
void write_rbq1(const HostIndex& ix, std::vector<uint8_t>& out);
void assign_shards(const std::vector<uint64_t>& list_bytes, int shard_count, std::vector<int>& owner);
int shard_layout(HostIndex& ix);  // list_owner / list_n / blk_off / vec_off from list_n_all and the shard coordinates

// ---- device view passed by value to kernels ---------------------------------------------------
struct DevIndex {
    int dim, D, metric, ex_bits, rot_type;
    int trunc;        // FHT window (largest power of two <= dim)
    float fac;        // 1/sqrt(trunc)
    uint32_t nlist;
    uint32_t block_stride;  // 4D+384
    uint32_t max_list_n;    // longest inverted list on this shard (vectors)
    uint32_t ex_stride;     // D*ex_bits/8
    const uint8_t* flip;    // 4*D/8
    const float* matrix_t;  // Matrix rotator, TRANSPOSED (k-major) for coalesced reads
    const float* centroids; // nlist*D
    const float* cent_q4;   // the centroids again with every row regrouped for 16-byte loads by the 8 "AVX lanes" of the exact
                            // re-score: float4 (k4, j) at [(k4*8 + j)*4, +4) = elements 8*(4*k4 + e) + j, e = 0..3 (D % 32 == 0; else nullptr)
    const void* cent_split; // nlist*3D bf16: [hi | lo | hi] split of the centroids (tensor-core coarse stage)
    const float* cent_n2;   // |c|^2 per centroid
    float cmax_norm;        // max |c|
    // strided sample of the centroids (coarse filter mode: a per-query score threshold is estimated from the sample, the
    // full GEMM's epilogue then keeps only the centroids that beat it)
    const void* samp_split; // samp_n*3D bf16, rows gathered from cent_split
    const float* samp_n2;   // |c|^2 of the sampled centroids
    uint32_t samp_n;        // 0: filter mode unavailable (small nlist)
    const uint32_t* list_n; // nlist
    const uint8_t* list_owner;  // nlist: owning shard of every list (nullptr: this handle owns all lists)
    int shard_rank;
    const uint32_t* blk_off;
    const uint64_t* vec_off;
    const uint8_t* blocks;
    const uint64_t* ids;
    const uint8_t* ex;
    const uint8_t* exl;     // lane-major ex-codes (one byte per code, 8 rows of exl_lane bytes per vector; resolve.cu)
    uint32_t exl_lane;      // bytes per row: D/8 rounded up to 16
    uint32_t exl_stride;    // 8 * exl_lane
    uint32_t exl_copy_coalesced;  // refine staging: 1 = a candidate's rows are copied in address order by its lane group (scan_common.cuh)
    const float* f_add_ex;
    const float* f_rescale_ex;
};

struct QueryScalars {  // one per query, written by the prep kernel
    float delta, sum_vl, k1x, kbx, qnorm, sum_q, bscale, pad;
};
struct Probe {  // one per (query, probe rank), written by the probe kernel (32 bytes)
    uint32_t cid;
    float g_add, g_error, dot_qc;
    uint32_t nv;        // vectors of the list on this shard (0: empty or owned elsewhere)
    uint32_t blk_off;   // first 32-vector block of the list
    uint64_t vec_off;   // first vector of the list
};
// TailWs::counters: slots 0..7 belong to the tail / replay kernels, then one work cursor per head-stage launch of a call
constexpr uint32_t kTailCounters = 12, kHeadCursors = 32;
constexpr uint32_t kOvfCtas = 160;  // CTAs of the overflow tier (>= one per SM), each with 4096 x 24 B of record scratch
struct DevStats {
    unsigned long long blocks, candidates, refined, admitted;
    // list-major tail stage (scan_tail.cu)
    unsigned long long tail_blocks;      // (query, block) evaluations done by the tail kernel
    unsigned long long tail_pairs;       // (query, list) pairs handed to the tail kernel
    unsigned long long survivors;        // tail candidates with lower bound < the query's head threshold
    unsigned long long overflow_queries; // queries whose survivor buffer overflowed (re-walked sequentially)
    unsigned long long fallback_queries; // queries the head pass could not take (list longer than the dense buffer, heap not full after it)
};
// A tail candidate that may still enter the top-k: replayed in (rank, pos) order by the replay pass.
struct Survivor {
    uint32_t rank;   // probe rank of its list
    uint32_t pos;    // position inside the list (visit order)
    float lower;     // lower bound (after the non-finite fallback)
    float x;         // ex_bits > 0: ip_x0_qr (feeds the ex-code distance); ex_bits == 0: the estimate
};
struct TailItem {  // work item of the tail kernel: a chunk of the (query, rank) pairs probing one list
    uint32_t cid, pair_begin, pair_count, pad;
};
// kScanFallback: the queries listed in TailWs::fb_list (bit 31 clear: the whole probe sequence from scratch; bit 31 set:
// resume from the head state at tail_start and walk the rest sequentially).
enum ScanMode { kScanFull = 0, kScanHead = 1, kScanReplay = 2, kScanFallback = 3, kScanFallbackResume = 4 };
// kScanFallback walks TailWs::fb_list (queries the head pass could not take: from scratch; known after the head pass, so the
// launch runs on a side stream beside the tail and replay kernels); kScanFallbackResume walks TailWs::fb2_list (queries the
// replay tiers handed back: resume from the head state), after the replay.
struct TailWs {  // device workspace of the head/tail/replay pipeline (per query tile)
    uint32_t* tail_start;  // [nq] first probe rank left to the tail stage (== nprobe: none)
    float* tau;            // [nq] k-th distance after the head stage (INF if the heap is not full)
    uint32_t* surv_cnt;    // [nq]
    Survivor* surv;        // [nq * surv_cap]
    uint32_t surv_cap;     // slots per query (buffer stride)
    uint32_t sort_cap;     // survivors the lazy replay sorts in shared memory; queries with sort_cap < n <= surv_cap take the overflow tier
    uint32_t* ovf_list;    // [nq] queries of the overflow tier (counters[7] = how many, counters[8] = its work cursor)
    void* ovf_recs;        // [kOvfCtas * 4096] XrRec: per-CTA record scratch of the overflow tier
    uint32_t* list_cnt;    // [nlist] pairs per list      } zeroed together with surv_cnt and the counters
    uint32_t* list_fill;   // [nlist] scatter cursors     }
    uint32_t* list_off;    // [nlist + 1]
    void* plan_tot;        // [1024] per-CTA totals of the work planner (scan_tail.cu)
    uint32_t* pairs;       // [nq * nprobe] pair ids (q * nprobe + rank) grouped by list
    TailItem* items;       // [max_items]
    uint32_t* counters;    // [0] n_items, [1] item cursor
    uint32_t max_items;
    uint32_t pairs_per_item;
    // head stage (resolve.cu): dense (lower bound, ip | estimate) of every vector of a query's first owned list
    float2* head_buf;      // [head_rows * head_cap]: row (q - first query of the head sub-chunk)
    uint32_t head_cap;     // slots per query (multiple of 32); longer lists send the query to the fallback path
    uint32_t head_rows;    // queries per head sub-chunk
    uint32_t* fb_list;     // [nq] queries left to the sequential fallback by the head pass (counters[2] = how many, counters[6] = cursor)
    uint32_t* fb2_list;    // [nq] queries handed back by the replay tiers, resume entries (counters[9] = how many, counters[10] = cursor)
    uint32_t* qlist;       // [nq] phased search: queries whose head pass runs on this shard, compacted; qcount = how many
    uint32_t* qcount;
};
constexpr uint32_t kFbResume = 0x80000000u;
// counters[]: [0] tail items, [1] tail item cursor, [2] fallback queries, [3] head-resolve cursor, [4] refine cursor,
// [5] replay cursor, [6] fallback cursor, [7] overflow-tier queries, [8] overflow-tier cursor, [9] resume-fallback queries,
// [10] resume-fallback cursor

// arguments of the tail kernels (scan_tail.cu: PRMT lookups; tail_tc.cu: one-hot GEMM on the tensor cores)
struct TailArgs {
    const uint8_t* lut;
    const QueryScalars* qs;
    const Probe* probes;
    uint32_t nprobe;
    const float* tau;
    const uint32_t* pairs;
    const TailItem* items;
    uint32_t* counters;  // [0] number of items, [1] next item
    Survivor* surv;
    uint32_t* surv_cnt;
    uint32_t surv_cap;
    const unsigned long long* filter;
    unsigned long long filter_nbits;
    DevStats* stats;
    uint32_t seg_blocks;  // blocks of a list staged at a time
    uint32_t has_ex;
};
bool tail_tc_supported(const DevIndex& ix);
int launch_tail_tc(const DevIndex& ix, const TailArgs& a, cudaStream_t st);

// kernels (each .cu exposes a launcher)
int launch_query_prep(const DevIndex& ix, const float* d_queries, size_t nq, float* d_rot, uint8_t* d_lut,
                      QueryScalars* d_qs, cudaStream_t st,
                      void* d_split = nullptr, float* d_n2 = nullptr, bool* split_done = nullptr);  // optional: also writes the coarse GEMM's bf16 operand
int launch_coarse_exact(const DevIndex& ix, const float* d_rot, size_t nq, float* d_scores, cudaStream_t st);
int launch_probe_select(const DevIndex& ix, const float* d_rot, const float* d_scores, size_t nq, size_t nprobe,
                        Probe* d_probes, cudaStream_t st);
// mode kScanFull: the whole probe sequence per query.  kScanHead: stop after the first list that leaves the
// heap full (writes tw->tail_start / tw->tau).  kScanReplay: continue from the head state with the tail
// kernel's survivors.  tw may be null for kScanFull.
int launch_scan(const DevIndex& ix, const float* d_rot, const uint8_t* d_lut, const QueryScalars* d_qs,
                const Probe* d_probes, size_t nq, size_t nprobe, size_t top_k, const uint64_t* d_filter,
                size_t filter_nbits, uint64_t* d_ids, float* d_scores, uint32_t* d_counts, DevStats* d_stats,
                unsigned int* d_work_counter, int mode, const TailWs* tw, cudaStream_t st);
// scan_tail.cu: group the tail (query, rank) pairs by list, then FastScan list-major; survivors -> tw.
int launch_tail(const DevIndex& ix, const uint8_t* d_lut, const QueryScalars* d_qs, const Probe* d_probes, size_t nq,
                size_t nprobe, const uint64_t* d_filter, size_t filter_nbits, DevStats* d_stats, const TailWs& tw,
                cudaStream_t st, uint64_t* launches, cudaEvent_t ev_begin = nullptr, cudaEvent_t ev_end = nullptr);
// resolve.cu: the list-major pipeline around the tail kernel.
//   head scan    FastScan of every query's first owned list -> tw.head_buf (dense)
//   head resolve the reference's sequential prune/refine/top-k over that list -> heap state, tw.tau, tw.tail_start
//   replay       survivors in reference order against the live threshold (distances precomputed)
// resolve.cu: phased multi-GPU search helpers.  export: Probe rows [q_begin, q_begin+q_count) -> records; import: records of all
// queries -> Probe with this shard's list geometry, head_owner[q] = this shard owns the query's nearest probed list
int launch_probe_export(const Probe* d_probes, size_t q_begin, size_t q_count, size_t nprobe, rbq_probe_rec* d_out, cudaStream_t st);
int launch_probe_import(const DevIndex& ix, const rbq_probe_rec* d_in, size_t nq, size_t nprobe, Probe* d_probes, uint8_t* d_head_owner,
                        cudaStream_t st);
// fetch.cu: fetch_embedding (position of an id; reconstruction of one stored vector)
int launch_find_id(const DevIndex& ix, size_t nvec, uint64_t id, unsigned long long* d_pos, cudaStream_t st);
int launch_fetch_embedding(const DevIndex& ix, uint32_t cid, uint32_t local, float delta, float vl, float* d_out, cudaStream_t st);
// exact_merge.cu: global replay of the one-call sharded search (bit-identical multi-GPU results)
struct XrRec {  // a candidate on its way to the query's home rank
    unsigned long long key;  // visit order: head-state entry i -> i; survivor -> (probe rank + 1) << 32 | position
    float lower;             // lower bound (head-state entries: -inf = admitted already)
    float dist;              // refined distance (1-bit index: the estimate)
    unsigned long long id;
};
int xr_launch_count_scan(const TailWs& tw, const uint8_t* d_head_owner, const uint32_t* d_head_cnt, size_t nq, uint32_t per, uint32_t world,
                         uint32_t* d_cnt, uint32_t* d_off, uint32_t* d_cta_tot, unsigned long long* d_send_tot, cudaStream_t st);
int xr_launch_records(const DevIndex& ix, const float* d_rot, const QueryScalars* d_qs, const Probe* d_probes, size_t nq, size_t nprobe, size_t top_k,
                      const TailWs& tw, const uint8_t* d_head_owner, const uint64_t* d_head_ids, const float* d_head_sc, const uint32_t* d_head_cnt,
                      const uint32_t* d_cnt, const uint32_t* d_off, XrRec* d_recs, uint32_t* d_cursor, cudaStream_t st);
int xr_launch_plan(const uint32_t* d_all_cnt, size_t nq, size_t q_begin, size_t q_count, int world, uint32_t* d_qoff, unsigned long long* d_recv_tot,
                   uint32_t* d_flagged, cudaStream_t st);
int xr_launch_replay(const XrRec* d_recv, const unsigned long long* d_rbase, const uint32_t* d_all_cnt, const uint32_t* d_qoff, const uint32_t* d_flagged,
                     size_t nq, size_t q_begin, size_t q_count, int world, size_t top_k, int metric, uint64_t* d_out_ids, float* d_out_sc,
                     uint32_t* d_out_cn, unsigned long long* d_inexact, cudaStream_t st);
int xr_launch_override(const void* d_gath, size_t chunk, size_t per, size_t nq, size_t top_k, uint64_t* d_ids, float* d_sc, uint32_t* d_cn, cudaStream_t st);
// resolve.cu: builds DevIndex::exl from the packed ex-codes already on the device (no-op for 1-bit indexes)
int prepare_ex_lanes(rbq_index* h);
int launch_head(const DevIndex& ix, const float* d_rot, const uint8_t* d_lut, const QueryScalars* d_qs, const Probe* d_probes,
                size_t nq, size_t nprobe, size_t top_k, const uint64_t* d_filter, size_t filter_nbits, uint64_t* d_ids,
                float* d_scores, uint32_t* d_counts, DevStats* d_stats, const TailWs& tw, cudaStream_t st, uint64_t* launches,
                size_t q_begin, size_t q_count, int* launch_index, const uint8_t* d_head_owner = nullptr);
int launch_refine_replay(const DevIndex& ix, const float* d_rot, const QueryScalars* d_qs, const Probe* d_probes, size_t nq,
                         size_t nprobe, size_t top_k, uint64_t* d_ids, float* d_scores, uint32_t* d_counts, DevStats* d_stats,
                         const TailWs& tw, cudaStream_t st, uint64_t* launches);
void tail_debug_set_survivor_cap(uint32_t cap);  // 0 = default
int launch_fill_u32(uint32_t* d_p, size_t n, uint32_t v, cudaStream_t st);
int launch_ex_dot_debug(const DevIndex& ix, const float* d_rot, const unsigned long long* d_gv, int n, float* d_out, const TailWs& tw, cudaStream_t st);
size_t tail_ws_bytes(const DevIndex& ix, size_t nq, size_t nprobe, size_t top_k);
void tail_ws_carve(const DevIndex& ix, size_t nq, size_t nprobe, size_t top_k, char* base, TailWs& tw);
int launch_scan_debug(const DevIndex& ix, const uint8_t* d_lut, const QueryScalars* d_qs, uint32_t cluster,
                      float g_add, float g_error, uint32_t* d_accu, float* d_ip, float* d_est, float* d_lb,
                      cudaStream_t st);
int launch_merge(int metric, int nshards, size_t nq, size_t top_k, const uint64_t* in_ids, const float* in_scores,
                 const uint32_t* in_counts, uint64_t* out_ids, float* out_scores, uint32_t* out_counts,
                 cudaStream_t st, size_t ids_stride = 0, size_t sc_stride = 0, size_t cn_stride = 0);
int launch_probe_select_tc(const DevIndex& ix, const float* d_rot, float* d_scores, const QueryScalars* d_qs, size_t nq,
                           size_t nprobe, float eps_g, Probe* d_probes, unsigned int* d_fallbacks, cudaStream_t st,
                           bool need_ip = true);
int launch_split_bf16(const float* d_x, size_t rows, int D, int centroid_side, void* d_out, float* d_n2, cudaStream_t st);
int launch_centroid_q4(const float* d_in, size_t rows, int D, float* d_out, cudaStream_t st);  // coarse.cu: DevIndex::cent_q4
int launch_split_bf16_pad(const float* d_x, size_t rows, int in_dim, int D, int centroid_side, void* d_out, float* d_n2, cudaStream_t st);
int launch_coarse_tc(const DevIndex& ix, const void* d_qsplit, const float* d_qn2, size_t nq, float* d_scores, cudaStream_t st, int terms = 3);
// coarse_tc.cu: the persistent tcgen05 GEMM behind every dense contraction of the engine.  A: rows x (3D bf16, pitch 3D), B: cols x
// (3D bf16); terms = 3 uses the whole split (fp32-class dot products), terms = 1 only the leading hi x hi block (bf16-class).
enum GemmMode { kGemmScores = 0, kGemmFilter = 1, kGemmArgmin = 2 };
struct CandRec {  // a centroid that passed the filter: its approximate score and id
    float score;
    uint32_t cid;
};
struct GemmEpi {
    int nq = 0, ncols = 0, metric = 0;
    int shifted = 0;                 // L2 scores without the row's |a|^2: fma(-2, a.b, |b|^2) -- the filter mode's score domain (ordering per
                                     // row is unchanged; one FFMA per score in the epilogue)
    const float* qn2 = nullptr;      // |a|^2 per row (L2, unshifted)
    const float* cn2 = nullptr;      // |b|^2 per column (L2)
    float* scores = nullptr;         // kGemmScores: [nq][ncols]
    const float* thr = nullptr;      // kGemmFilter: per-row threshold (L2: keep score <= thr; IP: keep score >= thr)
    CandRec* cand = nullptr;         // kGemmFilter: [nq][cap]
    uint32_t* cand_cnt = nullptr;    //              [nq], zeroed by the caller; may exceed cap (overflow)
    uint32_t cap = 0;
    unsigned long long* best = nullptr;  // kGemmArgmin: [nq], (order key of max(score, 0)) << 32 | column, initialised to ~0
};
int launch_coarse_gemm(int mode, const void* d_a, size_t rows, const void* d_b, size_t cols, int D, int terms, const GemmEpi& epi,
                       cudaStream_t st);
// coarse.cu, filter mode: threshold from the sample scores, selection from the candidate lists, exact fallback
struct FilterWs {
    float* samp_scores;   // [cq][samp_n]
    float* thr;           // [cq]
    CandRec* cand;        // [cq][cap]
    uint32_t* cand_cnt;   // [cq]  } zeroed together per chunk
    uint32_t* fb_count;   // [2]   } [0] queries sent to the exact fallback, [1] its work cursor
    uint32_t* fb_list;    // [cq]
    float* fb_scratch;    // [fb_ctas][nlist]
    uint32_t cap, fb_ctas;
};
int launch_sample_threshold(const float* d_samp_scores, size_t nq, uint32_t samp_n, uint32_t rank, int metric, float* d_thr, cudaStream_t st);
int launch_probe_select_cand(const DevIndex& ix, const float* d_rot, const QueryScalars* d_qs, size_t nq, size_t nprobe, float eps_g,
                             const FilterWs& fw, Probe* d_probes, unsigned int* d_fallbacks, cudaStream_t st, bool need_ip);
uint32_t filter_sample_rank(uint32_t nlist, uint32_t samp_n, size_t nprobe, int terms);
uint32_t filter_cand_cap(uint32_t nlist, uint32_t samp_n, size_t nprobe, int terms);
int prepare_coarse_sample(rbq_index* h);
int launch_gather_rows(const void* d_src, size_t row_bytes, const uint32_t* d_idx, size_t n, void* d_dst, cudaStream_t st);
int prepare_coarse_tc(rbq_index* h);  // builds cent_split / cent_n2 / cmax_norm from dev.centroids (api.cu)
size_t probe_select_max_nprobe();
size_t scan_max_topk();

// build.cu: quantise n vectors (already grouped by list) on the device.
struct BuildOut {  // device arrays, one entry per vector in list-concatenated order
    uint8_t* bin_rows;  // n * D/8, MSB-first row-major sign codes
    uint8_t* ex;        // n * ex_stride
    float *f_add, *f_rescale, *f_error, *f_add_ex, *f_rescale_ex, *delta, *vl;
    float* rnorm = nullptr;  // |residual| (optional: stored by the brute-force index file only)
};
int launch_build_quantize(const DevIndex& ix, const float* d_rot, const uint32_t* d_list_of, size_t n,
                          const float* d_cents, float t_const, const double* d_t_per_vec, BuildOut out,
                          cudaStream_t st, const unsigned long long* d_dst_pos = nullptr);
int launch_rotate_only(const DevIndex& ix, const float* d_in, size_t n, float* d_out, cudaStream_t st, const uint32_t* d_src = nullptr);
double best_rescale_factor_host(const float* o_abs, size_t dim, int ex_bits);
float const_scaling_factor_host(size_t D, int ex_bits, uint64_t seed);

}  // namespace rbq

// the opaque handle
struct rbq_index {
    rbq::HostIndex host;  // metadata + (small) host-side arrays; bulk arrays are released after upload
    rbq::DevIndex dev{};
    int device = 0;
    std::vector<void*> allocations;  // device allocations owned by the handle
    // workspace (grown on demand, guarded by mu)
    mutable std::mutex mu;
    mutable void* ws = nullptr;
    mutable size_t ws_bytes = 0;
    mutable rbq::DevStats* d_stats = nullptr;     // followed by the scan kernel's work counter
    unsigned int* work_counter() const { return reinterpret_cast<unsigned int*>(d_stats + 1); }
    unsigned int* fallback_counter() const { return work_counter() + 1; }
    mutable rbq_search_stats last_stats{};
    mutable cudaEvent_t ev[10] = {};  // stage boundaries 0..6, begin/end of the tail FastScan kernel alone, end of the tail stage (phased search)
    mutable cudaEvent_t ev_chunk[32][5] = {};  // profiled calls: front-end / head boundaries of every chunk of a tile (created on first use)
    mutable int dist_phase = 0;       // phased search: 3 after rbq_dist_tail (its stage times are read back lazily by rbq_last_search_stats)
    mutable cudaStream_t copy_stream = nullptr;  // H2D of host queries, overlapped with the front end (api.cu, HostFeed)
    mutable cudaStream_t compute_stream = nullptr;  // the host entry points' compute stream
    mutable cudaEvent_t feed_ev[16] = {};
    bool profiling = false;
    int scan_mode = 0;         // 0: auto, 1: sequential per-query walk, 2: list-major head/tail/replay
    int coarse_mode = -1;      // -1: auto (2 when the centroid table and nprobe allow it, else 1), 0: exact FP32 all-pairs,
                               // 1: dense tensor-core scores + exact re-score, 2: tensor-core scores filtered in the GEMM epilogue
    int exact_merge = 0;       // one-call sharded search: 1 = global replay at the query's home rank (bit-identical to one GPU)
    mutable unsigned long long last_inexact = 0;  // queries of the last exact-merge call that kept the phased answer
    int coarse_terms = 0;      // bf16 split terms multiplied by the coarse GEMM: 3 (fp32-class scores), 1 (bf16-class, wider re-score band), 0 auto
    float coarse_eps = 4.8828125e-4f;  // 2^-11: assumed bound on |gemm(q.c) - q.c| / (|q||c|) with 3 terms
    void* comm = nullptr;                   // ncclComm_t of the one-call sharded search (rbq_comm_init); NCCL is dlopen'ed
    mutable void* dist_ws = nullptr;        // its exchange buffers
    mutable size_t dist_ws_bytes = 0;
    mutable void* xr_ws = nullptr;          // exact merge: head-state snapshot, counts, offsets, home results (sized by nq, top_k)
    mutable size_t xr_ws_bytes = 0;
    mutable void* xr_recs = nullptr;        // exact merge: packed records to send | received records (sized per call)
    mutable size_t xr_recs_bytes = 0;
    mutable unsigned long long* xr_host = nullptr;  // pinned: per-peer record counts read back once per call
    mutable unsigned long long* xr_inexact = nullptr;  // device counter behind last_stats.inexact_queries (inside xr_ws)
    mutable cudaStream_t side_stream = nullptr;  // the head pass' sequential fallback runs here, beside the tail and replay kernels
    mutable cudaEvent_t side_fork = nullptr, side_join = nullptr;
    // host entry: the front end + head pass of consecutive feed chunks alternate over these streams (slot 0 = the compute stream), so
    // that the partial waves of one chunk's kernels are filled by the next chunk's (api.cu, search_device)
    mutable cudaStream_t slot_stream[3] = {};
    mutable cudaEvent_t slot_fork = nullptr, slot_join[3] = {};
    mutable cudaEvent_t busy_ev = nullptr;  // recorded after the last kernel of every call: the next call's stream waits on it
};
