/* rbq.h -- C ABI of the B200-native IVF+RaBitQ search engine (librbq.so).
 *
 * Drop-in boundary for the search path of lqhl/rabitq-rs v0.7.0.  Every entry point names the
 * reference interface it replaces (paths relative to the reference repo).  Plain pointers and
 * sizes only; no exceptions cross the boundary; all functions return an rbq_status.
 *
 * Threading: a handle may be shared between host threads; calls that launch work on one handle
 * are serialised by an internal mutex (the reference's `&self` search is re-entrant; here the
 * device workspace is per handle).
 */
#ifndef RBQ_H_
#define RBQ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors RabitqError (src/lib.rs:39-57). */
typedef enum rbq_status {
    RBQ_OK = 0,
    RBQ_DIMENSION_MISMATCH = 1,  /* RabitqError::DimensionMismatch{expected,got} */
    RBQ_INVALID_CONFIG = 2,      /* RabitqError::InvalidConfig(&str) */
    RBQ_EMPTY_INDEX = 3,         /* RabitqError::EmptyIndex */
    RBQ_IO = 4,                  /* RabitqError::Io */
    RBQ_INVALID_PERSISTENCE = 5, /* RabitqError::InvalidPersistence(&str) */
    RBQ_CUDA_ERROR = 6           /* no reference counterpart: device/driver failure */
} rbq_status;

/* Metric (src/lib.rs:31-37) and RotatorType (src/rotation.rs:8-15) tags = the on-disk tags. */
enum { RBQ_METRIC_L2 = 0, RBQ_METRIC_INNER_PRODUCT = 1 };
enum { RBQ_ROTATOR_MATRIX = 0, RBQ_ROTATOR_FHT_KAC = 1 };

typedef struct rbq_index rbq_index; /* IvfRabitqIndex (src/ivf.rs:935-946), device resident */

/* Thread-local message of the last failing call on this thread (Display of RabitqError). */
const char* rbq_last_error(void);

/* ---- persistence: IvfRabitqIndex::load_from_path / load_from_reader (src/ivf.rs:1477-1702) ----
 * Parses the "RBQ1" v3 stream (same validation and error strings), uploads it to `device`.
 * shard_rank/shard_count: inverted lists are assigned size-balanced to `shard_count` shards and only
 * the lists of `shard_rank` are kept on this device (centroids and rotator replicated).  Use 0/1 for
 * a complete index. */
int rbq_index_load(const char* path, int device, int shard_rank, int shard_count, rbq_index** out);
int rbq_index_load_mem(const uint8_t* bytes, size_t len, int device, int shard_rank, int shard_count,
                       rbq_index** out);
/* IvfRabitqIndex::save_to_path / save_to_writer (src/ivf.rs:1310-1474); complete (unsharded) handles only.
 * rbq_index_save_mem: pass out==NULL to query the size. */
int rbq_index_save(const rbq_index* ix, const char* path);
int rbq_index_save_mem(const rbq_index* ix, uint8_t* out, size_t cap, size_t* written);
/* The same stream with only the lists flagged in keep_list[cluster_count] (the other lists are written empty, their
 * centroids stay): probe selection on the result equals probe selection on the full index, so queries whose probed lists
 * are all kept get the full index's answers -- parity checks against a CPU implementation on a 100M-vector index. */
int rbq_index_save_lists_mem(const rbq_index* ix, const uint8_t* keep_list, size_t nlist, uint8_t* out, size_t cap, size_t* written);
void rbq_index_free(rbq_index* ix);

/* ---- build: IvfRabitqIndex::train_with_clusters (src/ivf.rs:1025-1103) on the GPU ----
 * data: n x dim row-major, centroids: nlist x dim, assignments: n cluster ids.  total_bits 1..9.
 * rotator_type/seed as the reference; the rotator state is drawn from a seeded splitmix64 stream (the
 * reference's ChaCha12 stream is not reproduced; the state is stored in the index file either way).
 * faster_config != 0 uses a constant rescale factor (RabitqConfig::faster, src/quantizer.rs:33-45).
 * rotator_state (optional, may be NULL): explicit flip bytes (4*padded/8) or matrix (padded^2 f32). */
int rbq_index_build(const float* data, size_t n, size_t dim, const float* centroids, size_t nlist,
                    const uint32_t* assignments, int total_bits, int metric, int rotator_type,
                    uint64_t seed, int faster_config, const uint8_t* rotator_state, int device,
                    rbq_index** out);

/* ---- streaming build on device-resident data (the same train_with_clusters, for data sets too large to pass as one host
 * array: BASELINE config 5 is 100M x 128).  The caller clusters first (rbq_kmeans_* below, or its own k-means) and announces
 * how many vectors every list will receive (list_sizes, host, nlist entries, WHOLE data set); chunks of vectors that already
 * live on the device are then added in ascending id order (vector i of a chunk gets id id_base + i; inside a list, ids keep
 * their order of arrival, like the forward scan of src/ivf.rs:1141-1149).  RabitqConfig::faster semantics (constant rescale
 * factor).  shard_rank/shard_count as rbq_index_load: only the lists of this shard are kept, the others' vectors are skipped.
 * max_chunk bounds the vectors per rbq_builder_add_device call.  rbq_builder_finish checks the announced sizes, releases the
 * build scratch and turns the builder into a searchable index (it consumes the builder, also on failure). */
typedef struct rbq_builder rbq_builder;
int rbq_builder_create(size_t dim, const float* centroids, size_t nlist, const uint32_t* list_sizes, int total_bits, int metric,
                       int rotator_type, uint64_t seed, const uint8_t* rotator_state, int device, int shard_rank, int shard_count,
                       size_t max_chunk, rbq_builder** out);
int rbq_builder_add_device(rbq_builder* b, const float* d_data, const uint32_t* d_assign, size_t m, uint64_t id_base, void* stream);
int rbq_builder_finish(rbq_builder* b, rbq_index** out);
void rbq_builder_free(rbq_builder* b);

/* ---- k-means for `train` on the device (src/kmeans.rs:71-186, 291-326, 439-643): training subset of at most
 * max_points_per_centroid * k points (0 = the reference's 256), Forgy initialisation, niter Lloyd iterations; the assignment
 * step is the engine's tcgen05 GEMM with the arg-min fused into its epilogue, the centroid update a deterministic segmented
 * mean.  d_data: n x dim (device), d_centroids: k x dim (device, out).  rbq_kmeans_assign_device = assign_full_dataset
 * (nearest centroid by |x|^2 + |c|^2 - 2 x.c clamped at 0, ties to the lower cluster id). */
int rbq_kmeans_device(const float* d_data, size_t n, size_t dim, size_t k, int niter, uint64_t seed, size_t max_points_per_centroid,
                      float* d_centroids, int device, void* stream);
int rbq_kmeans_assign_device(const float* d_data, size_t n, size_t dim, const float* d_centroids, size_t k, uint32_t* d_assign,
                             int device, void* stream);

/* ---- accessors: len / cluster_count (src/ivf.rs:1218-1230) and index fields ---- */
size_t rbq_index_len(const rbq_index* ix);           /* all vectors of the index (all shards) */
size_t rbq_index_local_len(const rbq_index* ix);     /* vectors resident in this shard */
size_t rbq_index_dim(const rbq_index* ix);
size_t rbq_index_padded_dim(const rbq_index* ix);
size_t rbq_index_cluster_count(const rbq_index* ix);
int rbq_index_metric(const rbq_index* ix);
int rbq_index_ex_bits(const rbq_index* ix);
int rbq_index_rotator_type(const rbq_index* ix);
int rbq_index_device(const rbq_index* ix);

/* ---- IvfRabitqIndex::fetch_embedding (src/ivf.rs:1247-1307): the vector reconstructed from its stored codes
 * (centroid + delta*code + vl in rotated space, then the rotator's inverse).  out: dim floats (host).  *found = 0
 * when no stored vector has this id (the reference returns None) -- on a list shard, when the id lives elsewhere. */
int rbq_fetch_embedding(const rbq_index* ix, uint64_t vector_id, float* out, int* found);

/* ---- search: IvfRabitqIndex::batch_search / search / search_filtered (src/ivf.rs:1705-1752) ----
 * queries: nq x dim row-major HOST memory.  Outputs (host): ids[nq*top_k], scores[nq*top_k] (L2:
 * estimated squared distance ascending; IP: score descending = -distance), counts[nq] = results per
 * query (<= top_k).  Unused slots: id = UINT64_MAX, score = 0.
 * nprobe is clamped to [1, cluster_count]; top_k == 0 yields counts == 0 (src/ivf.rs:1791-1794).
 * filter_bits (optional): dense bitset over u32 ids standing in for RoaringBitmap::contains
 * (src/ivf.rs:2017-2022); bit i of word i/64. */
int rbq_search_batch(const rbq_index* ix, const float* queries, size_t nq, size_t dim, size_t top_k,
                     size_t nprobe, uint64_t* ids, float* scores, uint32_t* counts);
int rbq_search_batch_filtered(const rbq_index* ix, const float* queries, size_t nq, size_t dim,
                              size_t top_k, size_t nprobe, const uint64_t* filter_bits,
                              size_t filter_nbits, uint64_t* ids, float* scores, uint32_t* counts);
/* Same, all buffers DEVICE pointers on the index's device; enqueued on `stream` (cudaStream_t cast
 * to void*, NULL = default stream); returns after enqueue (no host sync). d_dist (optional) receives
 * the raw distances (used by the multi-GPU merge). */
int rbq_search_batch_device(const rbq_index* ix, const float* d_queries, size_t nq, size_t dim,
                            size_t top_k, size_t nprobe, const uint64_t* d_filter_bits,
                            size_t filter_nbits, uint64_t* d_ids, float* d_scores, uint32_t* d_counts,
                            void* stream);

/* ---- multi-GPU merge: k-way merge of per-shard top-k lists (no reference counterpart; the
 * reference is single-process).  in_*: [nshards][nq][top_k] device arrays gathered from all shards
 * (e.g. ncclAllGather); out_*: [nq][top_k].  Comparator = the reference's result order. */
int rbq_merge_topk_device(const rbq_index* ix, int nshards, size_t nq, size_t top_k,
                          const uint64_t* in_ids, const float* in_scores, const uint32_t* in_counts,
                          uint64_t* out_ids, float* out_scores, uint32_t* out_counts, void* stream);

/* ---- multi-GPU search in three phases (list shards, two small exchanges) ----
 * No reference counterpart (src/ivf.rs is single-process).  All buffers are DEVICE pointers on the handle's
 * device, everything is enqueued on `stream`, nothing synchronises with the host.  The three calls of one batch must
 * be made in order on the same handle with the same (nq, top_k, nprobe) and no other search in between: the handle's
 * workspace carries the rotated queries, LUTs and probe lists from phase to phase.
 *  1. rbq_dist_front: rotate + LUT for ALL nq queries (every shard needs them), coarse scores + probe selection for
 *     the slice [q_begin, q_begin + q_count) only; writes that slice's rows of `d_probes` (nq x nprobe records).
 *     The caller all-gathers the rows (each rank contributes its slice).
 *  2. rbq_dist_head: attaches this shard's list geometry to the gathered probe lists and runs the head pass for the
 *     queries whose NEAREST probed list this shard owns; d_tau[q] = the k-th distance after it (+inf for the other
 *     queries).  The caller all-reduces d_tau with MIN: a valid upper bound of every shard's final k-th distance.
 *  3. rbq_dist_tail: all remaining (query, owned list) pairs pruned with the reduced d_tau, ordered replay; leaves the
 *     shard's local top-k in d_ids / d_scores / d_counts (the buffers given to rbq_dist_head) for
 *     rbq_merge_topk_device. */
typedef struct rbq_probe_rec {
    uint32_t cid;                   /* cluster id */
    float g_add, g_error, dot_qc;   /* per-(query, list) constants (src/ivf.rs:1850-1857) */
} rbq_probe_rec;
int rbq_dist_front(const rbq_index* ix, const float* d_queries, size_t nq, size_t dim, size_t top_k, size_t nprobe,
                   size_t q_begin, size_t q_count, rbq_probe_rec* d_probes, void* stream);
int rbq_dist_head(const rbq_index* ix, size_t nq, size_t top_k, size_t nprobe, const rbq_probe_rec* d_probes,
                  float* d_tau, uint64_t* d_ids, float* d_scores, uint32_t* d_counts, void* stream);
int rbq_dist_tail(const rbq_index* ix, size_t nq, size_t top_k, size_t nprobe, const float* d_tau, uint64_t* d_ids,
                  float* d_scores, uint32_t* d_counts, void* stream);

/* ---- multi-GPU search as ONE call (librbq owns the NCCL communicator; NCCL is dlopen'ed when these are first used) ----
 * What a Rust caller of batch_search (src/ivf.rs:1743) uses to spread one index over the GPUs of a box: one process (or
 * thread) per GPU, each with its shard handle (rbq_index_load / rbq_builder_* with shard_rank, shard_count).
 *  rbq_comm_unique_id: rank 0 creates the 128-byte rendezvous id and ships it to the other ranks (any transport).
 *  rbq_comm_init:      collective; rank / world must equal the handle's shard coordinates.
 *  rbq_search_batch_sharded_device: the three phases above with their exchanges (all-gather of the probe slices, MIN
 *      all-reduce of the thresholds, all-gather of the packed local top-k) and the merge, all enqueued on `stream`; every
 *      rank passes the whole batch (device) and receives the merged result (device).  Same answers as the three-call form.
 *  rbq_search_batch_sharded: host buffers; every rank passes the same batch, uploads only its 1/world slice over its own
 *      host link (the slices are all-gathered over NVLink) and receives the merged result.  Synchronous. */
#define RBQ_COMM_ID_BYTES 128
int rbq_comm_unique_id(uint8_t* id_out);
int rbq_comm_init(rbq_index* ix, const uint8_t* id, int rank, int world);
int rbq_comm_destroy(rbq_index* ix);
int rbq_search_batch_sharded_device(const rbq_index* ix, const float* d_queries, size_t nq, size_t dim, size_t top_k, size_t nprobe,
                                    uint64_t* d_ids, float* d_scores, uint32_t* d_counts, void* stream);
int rbq_search_batch_sharded(const rbq_index* ix, const float* queries, size_t nq, size_t dim, size_t top_k, size_t nprobe, uint64_t* ids,
                             float* scores, uint32_t* counts);
/* Exact merge (off by default).  The phased search lets every shard replay its own survivors against min(tau, local k-th
 * distance); a shard may then admit a bound-violating candidate the single sequence of src/ivf.rs:2013-2127 skipped (about 4
 * ids in 10 000 at GIST-1M).  With exact merge on, the one-call sharded search additionally refines every survivor, ships
 * {visit key, lower bound, distance, id} records -- and the head pass' heap from the shard that ran it -- to the query's
 * HOME rank (grouped ncclSend/ncclRecv), which replays the union in the reference's visit order against the live threshold:
 * the single-GPU answer, bit for bit.  The home results are all-gathered over the phased answer.  The call synchronises
 * with the host once (record counts).  Queries the exact path cannot take keep the phased answer and are counted in
 * rbq_search_stats::inexact_queries.  Must be set identically on every rank. */
int rbq_set_exact_merge(rbq_index* ix, int on);

/* The same merge over ONE all-gathered buffer: shard s occupies bytes [s*chunk_bytes, (s+1)*chunk_bytes) holding
 * ids[nq*top_k] (u64) | scores[nq*top_k] (f32) | counts[nq] (u32); chunk_bytes is a multiple of 8.  One collective
 * instead of three. */
int rbq_merge_topk_packed_device(const rbq_index* ix, int nshards, size_t nq, size_t top_k, const void* packed,
                                 size_t chunk_bytes, uint64_t* out_ids, float* out_scores, uint32_t* out_counts,
                                 void* stream);

/* Host-only: the shard every inverted list of an RBQ1 stream is assigned to for `shard_count` shards
 * (the same deterministic size-balanced map rbq_index_load uses; no GPU needed).  owner[i] in
 * [0, shard_count); list_sizes (optional) receives the vector count of every list. */
int rbq_shard_assignment(const uint8_t* bytes, size_t len, int shard_count, int32_t* owner, uint32_t* list_sizes,
                         size_t cap_lists, size_t* nlist_out);

/* ---- BruteForceRabitqIndex (src/brute_force.rs:203-650): no clustering, zero centroid, exhaustive scan ----
 * rbq_bf_train        = train (:214-285; RabitqConfig::faster only on the device; MatrixRotator needs rotator_state)
 * rbq_bf_load[_mem]   = load_from_path / load_from_reader (:388-523), "RBF1" v1, same validation and messages
 * rbq_bf_save[_mem]   = save_to_path / save_to_writer (:298-385); out == NULL queries the size
 * rbq_bf_search_batch = search / search_filtered (:526-650) for a batch of queries (host buffers): ids = positions in the
 *                       training slice, L2: distance ascending, IP: score (= -distance) descending; the reference's
 *                       sequential scalar float order is kept, so scores are bit-identical to it. */
typedef struct rbq_bf_index rbq_bf_index;
int rbq_bf_train(const float* data, size_t n, size_t dim, int total_bits, int metric, int rotator_type, uint64_t seed,
                 int faster_config, const uint8_t* rotator_state, int device, rbq_bf_index** out);
int rbq_bf_load(const char* path, int device, rbq_bf_index** out);
int rbq_bf_load_mem(const uint8_t* bytes, size_t len, int device, rbq_bf_index** out);
int rbq_bf_save(rbq_bf_index* ix, const char* path);
int rbq_bf_save_mem(rbq_bf_index* ix, uint8_t* out, size_t cap, size_t* written);
void rbq_bf_free(rbq_bf_index* ix);
size_t rbq_bf_len(const rbq_bf_index* ix);
size_t rbq_bf_dim(const rbq_bf_index* ix);
size_t rbq_bf_padded_dim(const rbq_bf_index* ix);
int rbq_bf_search_batch(rbq_bf_index* ix, const float* queries, size_t nq, size_t dim, size_t top_k, const uint64_t* filter_bits,
                        size_t filter_nbits, uint64_t* ids, float* scores, uint32_t* counts);

/* ---- diagnostics (SearchDiagnostics, src/ivf.rs:151-155, plus roofline accounting) ---- */
typedef struct rbq_search_stats {
    uint64_t queries;           /* queries in the last search call on this handle */
    uint64_t blocks_scanned;    /* 32-vector FastScan blocks streamed */
    uint64_t bytes_scanned;     /* blocks_scanned * (4*padded_dim + 384): algorithmic scan bytes */
    uint64_t candidates;        /* vectors whose lower bound was evaluated */
    uint64_t refined;           /* ex-code dot products evaluated on device (>= reference's extended_evaluations) */
    uint64_t admitted;          /* candidates that passed lb < distk in reference order (== estimated + non-finite) */
    uint64_t kernel_launches;   /* CUDA kernels launched by the call */
    uint64_t coarse_fallbacks;  /* queries whose tensor-core candidate set overflowed and were scored exactly */
    float ms_prep, ms_coarse, ms_select, ms_scan; /* CUDA-event times of the last *profiled* call */
    /* list-major scan (large batches): head pass / tail kernel / replay pass */
    uint64_t tail_blocks;       /* (query, block) FastScan evaluations done by the list-major tail kernel */
    uint64_t tail_bytes;        /* tail_blocks * (4*padded_dim + 384): the tail kernel's algorithmic bytes */
    uint64_t tail_pairs;        /* (query, list) pairs handled by the tail kernel */
    uint64_t survivors;         /* tail vectors with lower bound < the head threshold (replayed in reference order) */
    uint64_t overflow_queries;  /* queries whose survivor buffer overflowed (their tail was re-walked sequentially) */
    float ms_scan_head, ms_scan_tail, ms_scan_replay; /* split of ms_scan (sequential mode: all in head) */
    float ms_tail_kernel;       /* the tail FastScan kernel alone (events around that one launch; 0 in sequential mode) */
    uint32_t coarse_mode_used;  /* coarse stage the call ran: 0 exact, 1 dense tensor-core scores, 2 filtered in the GEMM epilogue */
    uint32_t front_chunk;       /* queries per front-end chunk */
    uint64_t fallback_queries;  /* queries the list-major head pass handed to the sequential kernel (nearest list longer than the
                                   dense head buffer, or fewer than top_k candidates in it) */
    uint64_t inexact_queries;   /* exact-merge sharded search: home queries of this rank that kept the phased answer (survivor
                                   overflow on some shard, head pass that did not fill the heap, > 4096 candidate records) */
    uint64_t exchanged_records; /* exact-merge sharded search: candidate records this rank received for its home queries */
    uint32_t coarse_terms_used; /* bf16 terms the coarse GEMM multiplied (3 or 1) */
    uint32_t reserved_;
} rbq_search_stats;
int rbq_last_search_stats(const rbq_index* ix, rbq_search_stats* out);
/* When on, search calls time each stage with CUDA events (adds host syncs; off by default). */
int rbq_set_profiling(rbq_index* ix, int on);
/* Scan-stage schedule: 0 = auto (list-major for large batches), 1 = sequential (one warp walks one query's
 * probe sequence, the reference's loop order), 2 = list-major (head pass until the heap is full, then all
 * remaining (query, list) pairs grouped by list, then an ordered replay of the surviving candidates).
 * Both reproduce search_cluster_v2_batched's decisions exactly (src/ivf.rs:2013-2127). */
int rbq_set_scan_mode(rbq_index* ix, int mode);
/* Coarse-stage implementation (replaces the centroid loop of src/ivf.rs:1782-1835):
 *  0 = exact FP32 scoring of every centroid (CUDA cores);
 *  1 = tensor-core (tcgen05) GEMM writing the dense query x centroid score matrix, then an exact FP32 re-score of the
 *      near-threshold centroids;
 *  2 = the same GEMM with the selection fused into its epilogue: a per-query score threshold is estimated from a strided
 *      centroid sample, the epilogue appends only the centroids that beat it to a short per-query candidate list, and the
 *      selection kernel works on that list (queries whose list is provably incomplete take an exact fallback).  No
 *      nq x nlist matrix is ever written; needs cluster_count >= 2048 and nprobe well below cluster_count;
 * -1 = auto (default): 2 when available, else 1.
 * All of them produce the reference's probe list bit for bit. */
int rbq_set_coarse_mode(rbq_index* ix, int mode);
/* bf16 terms of the coarse GEMM: 3 (operands split hi+lo, fp32-class scores, a handful of re-scored centroids), 1 (a third
 * of the tensor work, bf16-class scores, a wider band of centroids re-scored exactly) or 0 = auto (default): 1 when the GEMM
 * is the larger part of the front end (padded_dim >= 512), else 3.  Exact either way. */
int rbq_set_coarse_terms(rbq_index* ix, int terms);

/* ---- stage probes (parity tests call each device stage in isolation; host buffers) ----
 * rotated[nq*D], lut[nq*4D], scalars[nq*8] = delta,sum_vl,k1x,kbx,qnorm,sum_q,binary_scale,0
 * (FhtKacRotator::rotate_into src/rotation.rs:350-401, QueryLut::new src/ivf.rs:798-845,
 * QueryPrecomputed::new src/ivf.rs:862-878). */
int rbq_debug_query_prep(const rbq_index* ix, const float* queries, size_t nq, size_t dim,
                         float* rotated, uint8_t* lut, float* scalars);
/* probe_cids[nq*nprobe], probe_consts[nq*nprobe*3] = g_add,g_error,dot_qc in visit order
 * (src/ivf.rs:1782-1857). */
int rbq_debug_probe(const rbq_index* ix, const float* queries, size_t nq, size_t dim, size_t nprobe,
                    uint32_t* probe_cids, float* probe_consts);
/* FastScan over one list for one query: accu[nb*32] (u16 semantics), ip/est/lb[nb*32]
 * (simd::accumulate_batch_avx2 src/simd.rs:972, compute_batch_distances_u16 :1932). */
int rbq_debug_scan_list(const rbq_index* ix, const float* query, size_t dim, size_t cluster,
                        uint32_t* accu, float* ip, float* est, float* lb, size_t cap_vectors);

/* Stage probes through the PRODUCT kernels of the list-major schedule (host buffers; nq*cap output arrays).
 * which = 0: head_scan_kernel -- for every query the dense rows of its nearest non-empty probed list: out_a = lower bound (after
 *   the non-finite fallback, src/ivf.rs:2031-2042), out_b = ip_x0_qr (ex_bits > 0) or the estimate, out_n[q] = list length;
 * which = 1: the tail FastScan kernel (tcgen05 one-hot GEMM) over ALL probed lists with the pruning threshold at +inf: one
 *   record per vector (out_r = probe rank, out_p = position, out_a, out_b as above) in arbitrary order, out_n[q] records. */
int rbq_debug_stage(const rbq_index* ix, int which, const float* queries, size_t nq, size_t dim, size_t nprobe, size_t cap,
                    float* out_a, float* out_b, uint32_t* out_r, uint32_t* out_p, uint32_t* out_n);
/* K10 (ip_packed_ex2_f32 / ip_packed_ex6_f32 / generic, src/simd.rs:1722-1825) through the product's refine path: the ex-code
 * dot of the first n vectors of `cluster` against the rotated query. */
int rbq_debug_ex_dot(const rbq_index* ix, const float* query, size_t dim, size_t cluster, size_t n, float* out);

/* Test knob (process-wide): capacity of the per-query survivor buffer of the list-major scan; 0 restores the
 * default (derived from top_k).  A query that overflows it has its tail re-walked sequentially (still exact). */
int rbq_debug_set_survivor_cap(uint32_t cap);

#ifdef __cplusplus
}
#endif
#endif /* RBQ_H_ */
