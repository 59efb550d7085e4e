"""Host-side mirror of the reference's public search API over the C ABI.

`IvfRabitqIndex` follows the pyo3 class of the reference (src/python_bindings.rs:338-720:
constructor(dimension, metric), fit, fit_with_clusters, query, batch_query, save, load, __len__,
cluster_count) and adds the typed entry points the Rust API has (`search`, `batch_search`,
`search_filtered` with `SearchParams`, src/ivf.rs:1705-1752) returning u64 ids (the pyo3 class
squeezes ids through f32 and loses precision above 2^24).

Everything here is plumbing around librbq.so; there is no Python or CPU implementation of the
search path in this package.
"""
import ctypes as C
import enum
from dataclasses import dataclass

import numpy as np

from . import _ffi


class Metric(enum.IntEnum):  # reference src/lib.rs:31-37
    L2 = 0
    InnerProduct = 1


class RotatorType(enum.IntEnum):  # reference src/rotation.rs:8-15
    MatrixRotator = 0
    FhtKacRotator = 1


@dataclass(frozen=True)
class SearchParams:  # reference src/ivf.rs:22-26
    top_k: int
    nprobe: int


class RabitqError(Exception):  # reference src/lib.rs:39-57
    code = -1

    def __init__(self, msg, code=None):
        super().__init__(msg)
        if code is not None:
            self.code = code


class DimensionMismatch(RabitqError):
    code = _ffi.DIMENSION_MISMATCH


class InvalidConfig(RabitqError):
    code = _ffi.INVALID_CONFIG


class EmptyIndex(RabitqError):
    code = _ffi.EMPTY_INDEX


class IoError(RabitqError):
    code = _ffi.IO


class InvalidPersistence(RabitqError):
    code = _ffi.INVALID_PERSISTENCE


class CudaError(RabitqError):
    code = _ffi.CUDA_ERROR


_ERRORS = {c.code: c for c in (DimensionMismatch, InvalidConfig, EmptyIndex, IoError, InvalidPersistence, CudaError)}


def _check(rc):
    if rc != 0:
        raise _ERRORS.get(rc, RabitqError)(_ffi.last_error(), rc)


def _metric_from_str(metric):
    if isinstance(metric, (Metric, int)):
        return Metric(int(metric))
    if metric in ("euclidean", "l2"):
        return Metric.L2
    if metric in ("angular", "ip", "inner_product"):
        return Metric.InnerProduct
    raise ValueError(f"Invalid metric: {metric}. Use 'euclidean' or 'angular'")


def _rotator_from_str(rt):
    if isinstance(rt, (RotatorType, int)):
        return RotatorType(int(rt))
    if rt in ("fht", "random"):
        return RotatorType.FhtKacRotator
    if rt in ("matrix", "identity"):
        return RotatorType.MatrixRotator
    raise ValueError(f"Invalid rotator_type: {rt}. Use 'fht', 'random', 'matrix', or 'identity'")


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class IvfRabitqIndex:
    """Device-resident IVF+RaBitQ index (one GPU, or one shard of a list-sharded index)."""

    def __init__(self, dimension=None, metric="euclidean", device=0):
        self._h = None
        self.dimension = dimension
        self.metric = _metric_from_str(metric)
        self.device = int(device)

    # ---- lifetime -------------------------------------------------------------------------
    def _adopt(self, handle):
        self.close()
        self._h = handle
        L = _ffi.lib()
        self.dimension = int(L.rbq_index_dim(handle))
        self.metric = Metric(L.rbq_index_metric(handle))

    def close(self):
        if self._h is not None:
            _ffi.lib().rbq_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _need(self):
        if self._h is None:
            raise RuntimeError("Index not built yet. Call fit() first.")
        return self._h

    # ---- persistence (load_from_path / save_to_path) -----------------------------------------
    @classmethod
    def load_from_path(cls, path, device=0, shard_rank=0, shard_count=1):
        self = cls(device=device)
        self.load(path, shard_rank, shard_count)
        return self

    @classmethod
    def load_from_bytes(cls, blob, device=0, shard_rank=0, shard_count=1):
        self = cls(device=device)
        buf = np.frombuffer(bytes(blob), np.uint8)
        h = C.c_void_p()
        _check(_ffi.lib().rbq_index_load_mem(_ptr(buf), buf.size, self.device, shard_rank, shard_count, C.byref(h)))
        self._adopt(h)
        return self

    def load(self, path, shard_rank=0, shard_count=1):
        h = C.c_void_p()
        _check(_ffi.lib().rbq_index_load(str(path).encode(), self.device, shard_rank, shard_count, C.byref(h)))
        self._adopt(h)

    def save(self, path):
        _check(_ffi.lib().rbq_index_save(self._need(), str(path).encode()))

    def save_to_bytes(self):
        n = C.c_size_t()
        _check(_ffi.lib().rbq_index_save_mem(self._need(), None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        _check(_ffi.lib().rbq_index_save_mem(self._need(), _ptr(buf), buf.size, C.byref(n)))
        return buf.tobytes()

    def save_lists_to_bytes(self, keep):
        """RBQ1 stream holding only the lists flagged in keep[cluster_count] (the others empty): see rbq_index_save_lists_mem."""
        keep = np.ascontiguousarray(keep, np.uint8)
        n = C.c_size_t()
        _check(_ffi.lib().rbq_index_save_lists_mem(self._need(), _ptr(keep), keep.size, None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        _check(_ffi.lib().rbq_index_save_lists_mem(self._need(), _ptr(keep), keep.size, _ptr(buf), buf.size, C.byref(n)))
        return buf.tobytes()

    # ---- build (train_with_clusters / train) ----------------------------------------------------
    def fit_with_clusters(self, data, centroids, assignments, total_bits=7, rotator_type="random", seed=42,
                          faster_config=True, rotator_state=None):
        data = np.ascontiguousarray(data, np.float32)
        centroids = np.ascontiguousarray(centroids, np.float32)
        assignments = np.ascontiguousarray(assignments, np.uint32)
        if data.ndim != 2 or centroids.ndim != 2:
            raise ValueError("Data and centroids must be 2D arrays")
        if self.dimension is not None and (data.shape[1] != self.dimension or centroids.shape[1] != self.dimension):
            raise ValueError(f"Data/centroids dimension must match expected {self.dimension}")
        if data.shape[0] != assignments.shape[0]:
            raise ValueError("Data and assignments must have same length")
        rt = _rotator_from_str(rotator_type)
        rs = None if rotator_state is None else np.ascontiguousarray(rotator_state, np.uint8)
        h = C.c_void_p()
        _check(_ffi.lib().rbq_index_build(_ptr(data), data.shape[0], data.shape[1], _ptr(centroids), centroids.shape[0],
                                          _ptr(assignments), int(total_bits), int(self.metric), int(rt), int(seed),
                                          int(bool(faster_config)), _ptr(rs), self.device, C.byref(h)))
        self._adopt(h)

    def fit(self, data, nlist, total_bits=7, rotator_type="random", seed=42, faster_config=True, kmeans_iters=10):
        """k-means (torch, on the GPU) followed by fit_with_clusters.  The reference's k-means
        (src/kmeans.rs) is outside the drop-in scope; any clustering yields a valid index."""
        from .kmeans import kmeans_gpu

        data = np.ascontiguousarray(data, np.float32)
        if data.ndim != 2:
            raise ValueError("Data must be 2D array (N x D)")
        cents, assign = kmeans_gpu(data, nlist, iters=kmeans_iters, seed=seed, device=self.device)
        self.fit_with_clusters(data, cents, assign, total_bits, rotator_type, seed, faster_config)

    # ---- accessors ----------------------------------------------------------------------------
    def __len__(self):
        return int(_ffi.lib().rbq_index_len(self._need()))

    def local_len(self):
        return int(_ffi.lib().rbq_index_local_len(self._need()))

    def cluster_count(self):
        return int(_ffi.lib().rbq_index_cluster_count(self._need()))

    @property
    def padded_dim(self):
        return int(_ffi.lib().rbq_index_padded_dim(self._need()))

    @property
    def ex_bits(self):
        return int(_ffi.lib().rbq_index_ex_bits(self._need()))

    @property
    def handle(self):
        return self._need()

    def __repr__(self):
        return (f"IvfRabitqIndex(dimension={self.dimension}, metric={self.metric.name}, built={self._h is not None}, "
                f"clusters={self.cluster_count() if self._h is not None else 0})")

    # ---- search ---------------------------------------------------------------------------------
    def batch_search(self, queries, params, filter_bits=None):
        """IvfRabitqIndex::batch_search: returns (ids[nq,k] u64, scores[nq,k] f32, counts[nq] u32)."""
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq, dim = q.shape
        k = int(params.top_k)
        ids = np.full((nq, max(k, 1)), np.iinfo(np.uint64).max, np.uint64)
        scores = np.zeros((nq, max(k, 1)), np.float32)
        counts = np.zeros(nq, np.uint32)
        L = _ffi.lib()
        if filter_bits is None:
            rc = L.rbq_search_batch(self._need(), _ptr(q), nq, dim, k, int(params.nprobe), _ptr(ids), _ptr(scores),
                                    _ptr(counts))
        else:
            fb = np.ascontiguousarray(filter_bits, np.uint64)
            nbits = fb.size * 64
            if fb.size == 0:  # an empty filter is still a filter (admits nothing): the ABI tells "absent" by a NULL pointer
                fb = np.zeros(1, np.uint64)
            rc = L.rbq_search_batch_filtered(self._need(), _ptr(q), nq, dim, k, int(params.nprobe), _ptr(fb),
                                             nbits, _ptr(ids), _ptr(scores), _ptr(counts))
        _check(rc)
        return ids[:, :k], scores[:, :k], counts

    def search(self, query, params):
        """IvfRabitqIndex::search: list of (id, score)."""
        ids, scores, counts = self.batch_search(np.asarray(query, np.float32)[None, :], params)
        return [(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]

    def search_filtered(self, query, params, allowed_ids):
        """IvfRabitqIndex::search_filtered; `allowed_ids` plays the RoaringBitmap (u32 ids)."""
        fb = ids_to_bitset(allowed_ids)
        ids, scores, counts = self.batch_search(np.asarray(query, np.float32)[None, :], params, fb)
        return [(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]

    def query(self, query, k, nprobe=1):
        """pyo3 `query`: (k, 2) float32 array [id, score] (src/python_bindings.rs:536-583)."""
        q = np.asarray(query, np.float32)
        if q.ndim != 1 or (self.dimension is not None and q.shape[0] != self.dimension):
            raise ValueError(f"Query dimension {q.shape[-1]} does not match expected {self.dimension}")
        return self.batch_query(q[None, :], k, nprobe)[0]

    def batch_query(self, queries, k, nprobe=1):
        """pyo3 `batch_query`: list of (n_i, 2) float32 arrays (src/python_bindings.rs:593-665)."""
        q = np.asarray(queries, np.float32)
        if q.ndim != 2:
            raise ValueError("Queries must be 2D array (N x D)")
        if self.dimension is not None and q.shape[1] != self.dimension:
            raise ValueError(f"Query dimension {q.shape[1]} does not match expected {self.dimension}")
        ids, scores, counts = self.batch_search(q, SearchParams(k, nprobe))
        out = []
        for i in range(q.shape[0]):
            n = int(counts[i])
            out.append(np.stack([ids[i, :n].astype(np.float32), scores[i, :n]], axis=1))
        return out

    # device-resident entry (torch tensors on this index's device; no host copies, no sync)
    def batch_search_device(self, queries, top_k, nprobe, out_ids, out_scores, out_counts, filter_bits=None, stream=None):
        import torch

        assert queries.is_cuda and queries.dtype == torch.float32 and queries.is_contiguous()
        nq, dim = queries.shape
        st = C.c_void_p(stream) if stream else C.c_void_p(torch.cuda.current_stream(queries.device).cuda_stream)
        fb, fn = (None, 0) if filter_bits is None else (C.c_void_p(filter_bits.data_ptr()), filter_bits.numel() * 64)
        _check(_ffi.lib().rbq_search_batch_device(self._need(), C.c_void_p(queries.data_ptr()), nq, dim, int(top_k),
                                                  int(nprobe), fb, fn, C.c_void_p(out_ids.data_ptr()),
                                                  C.c_void_p(out_scores.data_ptr()), C.c_void_p(out_counts.data_ptr()), st))

    # ---- multi-GPU search in three phases (include/rbq.h: rbq_dist_front / _head / _tail) ----------------
    def _stream(self, t, stream):
        import torch

        return C.c_void_p(stream) if stream else C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    def dist_front(self, queries, top_k, nprobe, q_begin, q_count, probes, stream=None):
        """Phase 1: rotate + LUT for all queries, probe lists for the slice [q_begin, q_begin+q_count) -> rows of `probes`
        (int32 CUDA tensor [>= nq, nprobe, 4] = rbq_probe_rec)."""
        nq, dim = queries.shape
        assert queries.is_cuda and queries.is_contiguous() and probes.is_cuda and probes.is_contiguous() and probes.numel() >= nq * nprobe * 4
        _check(_ffi.lib().rbq_dist_front(self._need(), C.c_void_p(queries.data_ptr()), nq, dim, int(top_k), int(nprobe), int(q_begin),
                                         int(q_count), C.c_void_p(probes.data_ptr()), self._stream(queries, stream)))

    def dist_head(self, nq, top_k, nprobe, probes, tau, out_ids, out_scores, out_counts, stream=None):
        """Phase 2: head pass for the queries whose nearest probed list this shard owns; tau[q] = k-th distance (+inf elsewhere)."""
        _check(_ffi.lib().rbq_dist_head(self._need(), int(nq), int(top_k), int(nprobe), C.c_void_p(probes.data_ptr()),
                                        C.c_void_p(tau.data_ptr()), C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_scores.data_ptr()),
                                        C.c_void_p(out_counts.data_ptr()), self._stream(tau, stream)))

    def dist_tail(self, nq, top_k, nprobe, tau, out_ids, out_scores, out_counts, stream=None):
        """Phase 3: remaining (query, owned list) pairs pruned with the MIN-reduced tau; leaves the shard's local top-k."""
        _check(_ffi.lib().rbq_dist_tail(self._need(), int(nq), int(top_k), int(nprobe), C.c_void_p(tau.data_ptr()),
                                        C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_scores.data_ptr()),
                                        C.c_void_p(out_counts.data_ptr()), self._stream(tau, stream)))

    def merge_topk_device(self, nshards, nq, top_k, g_ids, g_scores, g_counts, out_ids, out_scores, out_counts, stream=None):
        _check(_ffi.lib().rbq_merge_topk_device(self._need(), int(nshards), int(nq), int(top_k), C.c_void_p(g_ids.data_ptr()),
                                                C.c_void_p(g_scores.data_ptr()), C.c_void_p(g_counts.data_ptr()),
                                                C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_scores.data_ptr()),
                                                C.c_void_p(out_counts.data_ptr()), self._stream(g_ids, stream)))

    def merge_topk_packed_device(self, nshards, nq, top_k, packed, chunk_bytes, out_ids, out_scores, out_counts, stream=None):
        _check(_ffi.lib().rbq_merge_topk_packed_device(self._need(), int(nshards), int(nq), int(top_k), C.c_void_p(packed.data_ptr()),
                                                       int(chunk_bytes), C.c_void_p(out_ids.data_ptr()), C.c_void_p(out_scores.data_ptr()),
                                                       C.c_void_p(out_counts.data_ptr()), self._stream(packed, stream)))

    def fetch_embedding(self, vector_id):
        """IvfRabitqIndex::fetch_embedding (src/ivf.rs:1247-1307): the reconstructed vector, or None if the id is unknown."""
        out = np.empty(int(_ffi.lib().rbq_index_dim(self._need())), np.float32)
        found = C.c_int(0)
        _check(_ffi.lib().rbq_fetch_embedding(self._need(), int(vector_id), _ptr(out), C.byref(found)))
        return out if found.value else None

    def stats(self):
        s = _ffi.SearchStats()
        _check(_ffi.lib().rbq_last_search_stats(self._need(), C.byref(s)))
        return {f: getattr(s, f) for f, _ in _ffi.SearchStats._fields_}

    def set_profiling(self, on):
        _check(_ffi.lib().rbq_set_profiling(self._need(), int(bool(on))))

    def set_scan_mode(self, mode):
        """0 auto, 1 sequential per-query walk, 2 list-major head/tail/replay (both exact)."""
        _check(_ffi.lib().rbq_set_scan_mode(self._need(), int(mode)))

    def set_coarse_mode(self, mode):
        """-1 auto, 0 exact FP32, 1 dense tensor-core scores, 2 scores filtered in the GEMM epilogue (all exact)."""
        _check(_ffi.lib().rbq_set_coarse_mode(self._need(), int(mode)))

    def set_coarse_terms(self, terms):
        _check(_ffi.lib().rbq_set_coarse_terms(self._need(), int(terms)))

    def set_exact_merge(self, on):
        """One-call sharded search: replay every query's candidates globally at its home rank (the single-GPU answer, bit for bit)."""
        _check(_ffi.lib().rbq_set_exact_merge(self._need(), int(bool(on))))

    # ---- stage probes (tests) ---------------------------------------------------------------------
    def debug_query_prep(self, queries):
        q = np.ascontiguousarray(queries, np.float32)
        nq, dim = q.shape
        D = self.padded_dim
        rot = np.empty((nq, D), np.float32)
        lut = np.empty((nq, 4 * D), np.uint8)
        sc = np.empty((nq, 8), np.float32)
        _check(_ffi.lib().rbq_debug_query_prep(self._need(), _ptr(q), nq, dim, _ptr(rot), _ptr(lut), _ptr(sc)))
        return rot, lut, sc

    def debug_probe(self, queries, nprobe):
        q = np.ascontiguousarray(queries, np.float32)
        nq, dim = q.shape
        npb = min(max(int(nprobe), 1), self.cluster_count())
        cids = np.empty((nq, npb), np.uint32)
        consts = np.empty((nq, npb, 3), np.float32)
        _check(_ffi.lib().rbq_debug_probe(self._need(), _ptr(q), nq, dim, npb, _ptr(cids), _ptr(consts)))
        return cids, consts

    def debug_stage(self, which, queries, nprobe, cap):
        """Product-kernel stage probe (rbq_debug_stage): which=0 head scan rows, which=1 tail survivors with threshold +inf.
        Returns (a, b, rank, pos, n) with [nq, cap] arrays (rank/pos only meaningful for which=1)."""
        q = np.ascontiguousarray(queries, np.float32)
        nq, dim = q.shape
        a = np.zeros((nq, cap), np.float32)
        b = np.zeros((nq, cap), np.float32)
        r = np.zeros((nq, cap), np.uint32)
        p = np.zeros((nq, cap), np.uint32)
        n = np.zeros(nq, np.uint32)
        _check(_ffi.lib().rbq_debug_stage(self._need(), int(which), _ptr(q), nq, dim, int(nprobe), int(cap), _ptr(a), _ptr(b), _ptr(r), _ptr(p), _ptr(n)))
        return a, b, r, p, n

    def debug_ex_dot(self, query, cluster, n):
        q = np.ascontiguousarray(query, np.float32)
        out = np.zeros(max(int(n), 1), np.float32)
        _check(_ffi.lib().rbq_debug_ex_dot(self._need(), _ptr(q), q.size, int(cluster), int(n), _ptr(out)))
        return out[:n]

    def debug_scan_list(self, query, cluster, list_len):
        q = np.ascontiguousarray(query, np.float32)
        slots = (list_len + 31) // 32 * 32
        accu = np.zeros(max(slots, 1), np.uint32)
        ip, est, lb = (np.zeros(max(slots, 1), np.float32) for _ in range(3))
        _check(_ffi.lib().rbq_debug_scan_list(self._need(), _ptr(q), q.size, int(cluster), _ptr(accu), _ptr(ip),
                                              _ptr(est), _ptr(lb), max(slots, 1)))
        return accu[:slots], ip[:slots], est[:slots], lb[:slots]


@dataclass(frozen=True)
class BruteForceSearchParams:  # reference src/brute_force.rs:21-31
    top_k: int


class BruteForceRabitqIndex:
    """Device-resident BruteForceRabitqIndex (src/brute_force.rs:203-650): train / save / load / search / search_filtered."""

    def __init__(self, device=0):
        self._h = None
        self.device = int(device)

    def close(self):
        if self._h is not None:
            _ffi.lib().rbq_bf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _need(self):
        if self._h is None:
            raise RuntimeError("Index not built yet. Call train() first.")
        return self._h

    @classmethod
    def train(cls, data, total_bits, metric="euclidean", rotator_type="random", seed=42, use_faster_config=True, device=0, rotator_state=None):
        data = np.ascontiguousarray(data, np.float32)
        if data.ndim != 2:
            raise ValueError("Data must be 2D array (N x D)")
        self = cls(device)
        rs = None if rotator_state is None else np.ascontiguousarray(rotator_state, np.uint8)
        h = C.c_void_p()
        _check(_ffi.lib().rbq_bf_train(_ptr(data), data.shape[0], data.shape[1], int(total_bits), int(_metric_from_str(metric)),
                                       int(_rotator_from_str(rotator_type)), int(seed), int(bool(use_faster_config)), _ptr(rs), self.device,
                                       C.byref(h)))
        self._h = h
        return self

    @classmethod
    def load_from_bytes(cls, blob, device=0):
        self = cls(device)
        buf = np.frombuffer(bytes(blob), np.uint8)
        h = C.c_void_p()
        _check(_ffi.lib().rbq_bf_load_mem(_ptr(buf), buf.size, self.device, C.byref(h)))
        self._h = h
        return self

    @classmethod
    def load_from_path(cls, path, device=0):
        self = cls(device)
        h = C.c_void_p()
        _check(_ffi.lib().rbq_bf_load(str(path).encode(), self.device, C.byref(h)))
        self._h = h
        return self

    def save_to_path(self, path):
        _check(_ffi.lib().rbq_bf_save(self._need(), str(path).encode()))

    def save_to_bytes(self):
        n = C.c_size_t()
        _check(_ffi.lib().rbq_bf_save_mem(self._need(), None, 0, C.byref(n)))
        buf = np.empty(n.value, np.uint8)
        _check(_ffi.lib().rbq_bf_save_mem(self._need(), _ptr(buf), buf.size, C.byref(n)))
        return buf.tobytes()

    def __len__(self):
        return int(_ffi.lib().rbq_bf_len(self._need()))

    def is_empty(self):
        return len(self) == 0

    def batch_search(self, queries, params, filter_bits=None):
        q = np.ascontiguousarray(queries, np.float32)
        if q.ndim == 1:
            q = q[None, :]
        nq, dim = q.shape
        k = int(params.top_k)
        ids = np.full((nq, max(k, 1)), np.iinfo(np.uint64).max, np.uint64)
        scores = np.zeros((nq, max(k, 1)), np.float32)
        counts = np.zeros(nq, np.uint32)
        fb, nbits = None, 0
        if filter_bits is not None:
            fb = np.ascontiguousarray(filter_bits, np.uint64)
            nbits = fb.size * 64
            if fb.size == 0:
                fb = np.zeros(1, np.uint64)
        _check(_ffi.lib().rbq_bf_search_batch(self._need(), _ptr(q), nq, dim, k, _ptr(fb), nbits, _ptr(ids), _ptr(scores), _ptr(counts)))
        return ids[:, :k], scores[:, :k], counts

    def search(self, query, params):
        ids, scores, counts = self.batch_search(np.asarray(query, np.float32)[None, :], params)
        return [(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]

    def search_filtered(self, query, params, allowed_ids):
        ids, scores, counts = self.batch_search(np.asarray(query, np.float32)[None, :], params, ids_to_bitset(allowed_ids))
        return [(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]


class IndexBuilder:
    """Streaming train_with_clusters on device-resident data (include/rbq.h: rbq_builder_*): announce the list sizes, add
    chunks of (vectors, assignments) CUDA tensors in ascending id order, finish() -> IvfRabitqIndex.  With shard_count > 1
    only the lists of shard_rank are kept (the shard rbq_index_load would keep from the complete file)."""

    def __init__(self, dim, centroids, list_sizes, total_bits=7, metric="euclidean", rotator_type="random", seed=42, device=0,
                 shard_rank=0, shard_count=1, max_chunk=1 << 20, rotator_state=None):
        cents = np.ascontiguousarray(centroids, np.float32)
        sizes = np.ascontiguousarray(list_sizes, np.uint32)
        if cents.ndim != 2 or cents.shape[1] != dim or sizes.shape != (cents.shape[0],):
            raise ValueError("centroids must be [nlist, dim] and list_sizes [nlist]")
        rs = None if rotator_state is None else np.ascontiguousarray(rotator_state, np.uint8)
        self.device, self.metric = int(device), _metric_from_str(metric)
        self._b = C.c_void_p()
        _check(_ffi.lib().rbq_builder_create(int(dim), _ptr(cents), cents.shape[0], _ptr(sizes), int(total_bits), int(self.metric),
                                             int(_rotator_from_str(rotator_type)), int(seed), _ptr(rs), self.device, int(shard_rank),
                                             int(shard_count), int(max_chunk), C.byref(self._b)))

    def add(self, x, assign, id_base, stream=None):
        """x: [m, dim] float32 CUDA tensor, assign: [m] int32/uint32 CUDA tensor of list ids; vector i gets id id_base + i."""
        import torch

        assert x.is_cuda and x.is_contiguous() and x.dtype == torch.float32 and assign.is_cuda and assign.is_contiguous() and assign.numel() == x.shape[0]
        st = C.c_void_p(stream) if stream else C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _check(_ffi.lib().rbq_builder_add_device(self._b, C.c_void_p(x.data_ptr()), C.c_void_p(assign.data_ptr()), x.shape[0], int(id_base), st))

    def finish(self):
        h = C.c_void_p()
        b, self._b = self._b, None
        _check(_ffi.lib().rbq_builder_finish(b, C.byref(h)))
        ix = IvfRabitqIndex(device=self.device)
        ix._adopt(h)
        return ix

    def __del__(self):
        try:
            if getattr(self, "_b", None):
                _ffi.lib().rbq_builder_free(self._b)
                self._b = None
        except Exception:
            pass


def shard_assignment(blob, shard_count):
    """(owner[nlist] int32, list_sizes[nlist] uint32): the deterministic size-balanced list -> shard map
    used by load(..., shard_rank, shard_count).  Host-only (no GPU needed)."""
    buf = np.frombuffer(bytes(blob), np.uint8)
    n = C.c_size_t()
    _check(_ffi.lib().rbq_shard_assignment(_ptr(buf), buf.size, int(shard_count), None, None, 0, C.byref(n)))
    owner = np.empty(n.value, np.int32)
    sizes = np.empty(n.value, np.uint32)
    _check(_ffi.lib().rbq_shard_assignment(_ptr(buf), buf.size, int(shard_count), _ptr(owner), _ptr(sizes), n.value, C.byref(n)))
    return owner, sizes


def ids_to_bitset(allowed_ids, nbits=None):
    """Dense u64 bitset over u32 ids (the C ABI's stand-in for RoaringBitmap)."""
    a = np.asarray(list(allowed_ids) if not isinstance(allowed_ids, np.ndarray) else allowed_ids, np.uint64)
    n = int(nbits if nbits is not None else (int(a.max()) + 1 if a.size else 0))
    words = np.zeros((n + 63) // 64, np.uint64)
    if a.size:
        np.bitwise_or.at(words, (a // 64).astype(np.int64), np.uint64(1) << (a % np.uint64(64)))
    return words
