"""ctypes front-end for the CPU oracle (oracle/oracle.cc).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference leg of bench.py.  The product package
(rabitq_rs_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")


def _cpu_stamp():
    """The library is built -march=native: one built on another host (it travels with the repo snapshot) must be rebuilt."""
    try:
        import hashlib

        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
        return hashlib.sha1(flags.encode()).hexdigest()
    except OSError:
        return "unknown"


def build(force=False):
    src = os.path.join(_HERE, "oracle.cc")
    stamp_path = os.path.join(_HERE, ".build_stamp")
    stamp = _cpu_stamp()
    try:
        have = open(stamp_path).read().strip()
    except OSError:
        have = ""
    stale = not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src) or \
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "Makefile"))
    if force or stale or have != stamp:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], env={**os.environ, "CXX": ""})
        with open(stamp_path, "w") as f:
            f.write(stamp)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        L = _lib
        L.orc_last_error.restype = C.c_char_p
        L.orc_dot.restype = C.c_float
        L.orc_l2sqr.restype = C.c_float
        L.orc_ip_ex.restype = C.c_float
        L.orc_floor_log2.restype = C.c_size_t
        L.orc_padded_dim.restype = C.c_size_t
        L.orc_crc32.restype = C.c_uint32
        L.orc_best_rescale_factor.restype = C.c_double
        L.orc_const_scaling_factor.restype = C.c_float
        L.orc_heap_topk.restype = C.c_size_t
        L.orc_index_new.restype = C.c_void_p
        for n in ("len", "dim", "padded_dim", "nlist", "list_len"):
            getattr(L, "orc_index_" + n).restype = C.c_size_t
        L.orc_index_save.restype = C.c_size_t
        for n in ("list_blocks", "list_ids", "list_centroid", "list_ex", "list_f_add_ex", "list_f_rescale_ex"):
            getattr(L, "orc_index_" + n).restype = C.c_void_p
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _sz(x):
    return C.c_size_t(int(x))


def set_mode(fused_dist=1, ex_lanes=8):
    lib().orc_set_mode(int(fused_dist), int(ex_lanes))


def set_simd(on):
    lib().orc_set_simd(int(on))


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    """OpenMP threads of the batch search (one query per thread).  Call with os.cpu_count() when the process was started
    by a launcher that exports OMP_NUM_THREADS=1 (torchrun)."""
    lib().orc_set_num_threads(int(n))


def simd_level():
    return lib().orc_simd_level()


def dot(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot(_p(a), _p(b), _sz(a.size)))


def l2sqr(a, b):
    a, b = _f32(a), _f32(b)
    return float(lib().orc_l2sqr(_p(a), _p(b), _sz(a.size)))


def floor_log2(x):
    return int(lib().orc_floor_log2(_sz(x)))


def padded_dim(rotator_type, dim):
    return int(lib().orc_padded_dim(int(rotator_type), _sz(dim)))


def fht(d):
    d = _f32(d).copy()
    lib().orc_fht(_p(d), _sz(d.size))
    return d


def rotate_fht(flip, dim, x):
    x = _f32(x)
    flip = np.ascontiguousarray(flip, dtype=np.uint8)
    out = np.empty(padded_dim(1, dim), np.float32)
    lib().orc_rotate_fht(_p(flip), _sz(dim), _p(x), _p(out))
    return out


def pack_lut_f32(q):
    q = _f32(q)
    out = np.empty(q.size * 4, np.float32)
    lib().orc_pack_lut_f32(_p(q), _sz(q.size), _p(out))
    return out


def build_lut(rq):
    rq = _f32(rq)
    out = np.empty(rq.size * 4, np.uint8)
    d, s = C.c_float(), C.c_float()
    lib().orc_build_lut(_p(rq), _sz(rq.size), _p(out), C.byref(d), C.byref(s))
    return out, d.value, s.value


def pack_binary_code(bits):
    bits = np.ascontiguousarray(bits, np.uint8)
    out = np.empty((bits.size + 7) // 8, np.uint8)
    lib().orc_pack_binary_code(_p(bits), _sz(bits.size), _p(out))
    return out


def pack_codes(codes, nvec, dim_bytes):
    codes = np.ascontiguousarray(codes, np.uint8)
    nb = (nvec + 31) // 32
    out = np.zeros(nb * 32 * dim_bytes, np.uint8)
    lib().orc_pack_codes(_p(codes), _sz(nvec), _sz(dim_bytes), _p(out))
    return out


def unpack_single_vector(packed, vec, dim_bytes):
    packed = np.ascontiguousarray(packed, np.uint8)
    out = np.empty(dim_bytes * 8, np.uint8)
    lib().orc_unpack_single_vector(_p(packed), int(vec), _sz(dim_bytes), _p(out))
    return out


def accumulate_block(codes, lut, D, fast=False):
    codes = np.ascontiguousarray(codes, np.uint8)
    lut = np.ascontiguousarray(lut).view(np.uint8)
    out = np.empty(32, np.uint16)
    lib().orc_accumulate_block(_p(codes), _p(lut), _sz(D), _p(out), int(fast))
    return out


def batch_distances(accu, delta, sum_vl, f_add, f_rescale, f_error, g_add, g_error, k1x):
    accu = np.ascontiguousarray(accu, np.uint16)
    fa, fr, fe = _f32(f_add), _f32(f_rescale), _f32(f_error)
    ip, est, lb = (np.empty(32, np.float32) for _ in range(3))
    lib().orc_batch_distances(_p(accu), C.c_float(delta), C.c_float(sum_vl), _p(fa), _p(fr), _p(fe),
                              C.c_float(g_add), C.c_float(g_error), C.c_float(k1x), _p(ip), _p(est), _p(lb))
    return ip, est, lb


def ex_bytes(D, bits):
    return D * bits // 8 if bits > 0 else 0


def pack_ex(code, bits):
    code = np.ascontiguousarray(code, np.uint16)
    out = np.zeros((code.size * bits + 7) // 8, np.uint8)
    lib().orc_pack_ex(_p(code), _sz(code.size), int(bits), _p(out))
    return out


def unpack_ex(packed, dim, bits):
    packed = np.ascontiguousarray(packed, np.uint8)
    out = np.empty(dim, np.uint16)
    lib().orc_unpack_ex(_p(packed), _sz(dim), int(bits), _p(out))
    return out


def ip_ex(q, packed, bits, fast=False):
    q = _f32(q)
    packed = np.ascontiguousarray(packed, np.uint8)
    return float(lib().orc_ip_ex(_p(q), _p(packed), _sz(q.size), int(bits), int(fast)))


def crc32(b):
    a = np.frombuffer(bytes(b), np.uint8)
    return int(lib().orc_crc32(_p(a), _sz(a.size)))


def best_rescale_factor(o_abs, ex_bits):
    o = _f32(o_abs)
    return float(lib().orc_best_rescale_factor(_p(o), _sz(o.size), int(ex_bits)))


def const_scaling_factor(D, ex_bits, seed):
    return float(lib().orc_const_scaling_factor(_sz(D), int(ex_bits), C.c_uint64(seed)))


def quantize(data, centroid, ex_bits, metric, t_const=-1.0):
    data, centroid = _f32(data), _f32(centroid)
    D = data.size
    b = np.empty(D // 8, np.uint8)
    e = np.zeros(max(ex_bytes(D, ex_bits), 1), np.uint8)
    f = np.empty(7, np.float32)
    lib().orc_quantize(_p(data), _p(centroid), _sz(D), int(ex_bits), int(metric), C.c_float(t_const),
                       _p(b), _p(e), _p(f))
    return b, e[:ex_bytes(D, ex_bits)], f


def heap_topk(dist, ids, k):
    dist = _f32(dist)
    ids = np.ascontiguousarray(ids, np.uint64)
    od = np.empty(k, np.float32)
    oi = np.empty(k, np.uint64)
    n = lib().orc_heap_topk(_p(dist), _p(ids), _sz(dist.size), _sz(k), _p(od), _p(oi))
    return od[:n], oi[:n]


def kmeans(x, k, iters=10, seed=1):
    x = _f32(x)
    n, dim = x.shape
    cents = np.empty((k, dim), np.float32)
    assign = np.empty(n, np.uint32)
    lib().orc_kmeans(_p(x), _sz(n), _sz(dim), _sz(k), int(iters), C.c_uint64(seed), _p(cents), _p(assign))
    return cents, assign


def make_flip_bytes(dim, seed):
    """Rotator state for an FhtKac index: 4*padded/8 random bytes (the reference draws them
    from StdRng=ChaCha12, src/rotation.rs:256-261; the file carries them, so any source works)."""
    D = padded_dim(1, dim)
    return np.random.default_rng(seed).integers(0, 256, 4 * D // 8, dtype=np.uint8)


def make_matrix_bytes(dim, seed):
    """Orthonormal matrix (QR) for a MatrixRotator index, row-major f32."""
    g = np.random.default_rng(seed).standard_normal((dim, dim))
    q, _ = np.linalg.qr(g)
    return np.ascontiguousarray(q.astype(np.float32)).view(np.uint8).reshape(-1)


class OracleError(Exception):
    def __init__(self, code, msg):
        super().__init__(f"[{code}] {msg}")
        self.code = code
        self.msg = msg


class Index:
    """Restated IvfRabitqIndex (src/ivf.rs:935-946)."""

    def __init__(self):
        self.h = C.c_void_p(lib().orc_index_new())

    def __del__(self):
        if getattr(self, "h", None) and _lib is not None:
            _lib.orc_index_free(self.h)
            self.h = None

    @classmethod
    def train_with_clusters(cls, data, centroids, assignments, total_bits, metric, rotator_type=1,
                            seed=42, faster_config=False, rotator_bytes=None):
        data, centroids = _f32(data), _f32(centroids)
        assignments = np.ascontiguousarray(assignments, np.uint32)
        n, dim = data.shape
        if rotator_bytes is None:
            rotator_bytes = make_flip_bytes(dim, seed) if rotator_type == 1 else make_matrix_bytes(dim, seed)
        rotator_bytes = np.ascontiguousarray(rotator_bytes, np.uint8)
        t_const = -1.0
        if faster_config and total_bits > 1:
            t_const = const_scaling_factor(padded_dim(rotator_type, dim), total_bits - 1, seed)
        self = cls()
        rc = lib().orc_index_train_with_clusters(self.h, _p(data), _sz(n), _sz(dim), _p(centroids),
                                                 _sz(centroids.shape[0]), _p(assignments), int(total_bits),
                                                 int(metric), int(rotator_type), _p(rotator_bytes),
                                                 C.c_float(t_const))
        if rc != 0:
            raise OracleError(rc, "invalid configuration")
        return self

    @classmethod
    def train(cls, data, nlist, total_bits, metric, rotator_type=1, seed=42, faster_config=False, iters=8):
        cents, assign = kmeans(data, nlist, iters=iters, seed=seed)
        return cls.train_with_clusters(data, cents, assign, total_bits, metric, rotator_type, seed, faster_config)

    def save_bytes(self):
        n = lib().orc_index_save(self.h, None, _sz(0))
        buf = np.empty(n, np.uint8)
        lib().orc_index_save(self.h, _p(buf), _sz(n))
        return buf.tobytes()

    @classmethod
    def load_bytes(cls, b):
        a = np.frombuffer(bytes(b), np.uint8)
        self = cls()
        rc = lib().orc_index_load(self.h, _p(a), _sz(a.size))
        if rc != 0:
            raise OracleError(rc, lib().orc_last_error().decode())
        return self

    def __len__(self):
        return int(lib().orc_index_len(self.h))

    dim = property(lambda s: int(lib().orc_index_dim(s.h)))
    padded_dim = property(lambda s: int(lib().orc_index_padded_dim(s.h)))
    nlist = property(lambda s: int(lib().orc_index_nlist(s.h)))
    metric = property(lambda s: int(lib().orc_index_metric(s.h)))
    ex_bits = property(lambda s: int(lib().orc_index_ex_bits(s.h)))

    def list_len(self, c):
        return int(lib().orc_index_list_len(self.h, _sz(c)))

    def _view(self, fn, c, n, dtype):
        ptr = getattr(lib(), fn)(self.h, _sz(c))
        if n == 0 or not ptr:
            return np.empty(0, dtype)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (n * np.dtype(dtype).itemsize,)).view(dtype).copy()

    def list_blocks(self, c):
        nb = (self.list_len(c) + 31) // 32
        return self._view("orc_index_list_blocks", c, nb * (4 * self.padded_dim + 384), np.uint8)

    def list_ids(self, c):
        return self._view("orc_index_list_ids", c, self.list_len(c), np.uint64)

    def list_centroid(self, c):
        return self._view("orc_index_list_centroid", c, self.padded_dim, np.float32)

    def list_ex(self, c):
        return self._view("orc_index_list_ex", c, self.list_len(c) * ex_bytes(self.padded_dim, self.ex_bits), np.uint8)

    def list_f_add_ex(self, c):
        return self._view("orc_index_list_f_add_ex", c, self.list_len(c), np.float32)

    def list_f_rescale_ex(self, c):
        return self._view("orc_index_list_f_rescale_ex", c, self.list_len(c), np.float32)

    def rotate(self, x):
        x = _f32(x)
        out = np.empty(self.padded_dim, np.float32)
        lib().orc_index_rotate(self.h, _p(x), _p(out))
        return out

    def inverse_rotate(self, x):
        x = _f32(x)
        out = np.empty(self.dim, np.float32)
        lib().orc_index_inverse_rotate(self.h, _p(x), _p(out))
        return out

    def fetch_embedding(self, vector_id):
        """IvfRabitqIndex::fetch_embedding (src/ivf.rs:1247-1307): the reconstructed vector or None."""
        out = np.empty(self.dim, np.float32)
        lib().orc_fetch_embedding.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        return out if lib().orc_fetch_embedding(self.h, int(vector_id), _p(out)) else None

    def search_batch(self, queries, top_k, nprobe, filter_bits=None, naive=False, want_diag=False):
        """Returns (ids[nq,k] u64, scores[nq,k] f32, counts[nq] u32[, diag[nq,4]])."""
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, dim = q.shape
        k = max(int(top_k), 1)
        ids = np.zeros((nq, k), np.uint64)
        scores = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        diag = np.zeros((nq, 4), np.uint64) if want_diag else None
        fb, fn = None, 0
        if filter_bits is not None:
            fb = np.ascontiguousarray(filter_bits, np.uint64)
            fn = fb.size * 64
        rc = lib().orc_search_batch(self.h, _p(q), _sz(nq), _sz(dim), _sz(top_k), _sz(nprobe), _p(fb), _sz(fn),
                                    _p(ids), _p(scores), _p(counts), _p(diag), int(naive))
        if rc != 0:
            raise OracleError(rc, {1: "dimension mismatch", 3: "index is empty"}.get(rc, "error"))
        return (ids, scores, counts, diag) if want_diag else (ids, scores, counts)

    def search_dump(self, query, top_k, nprobe):
        """Single query with every intermediate stage (for stage-wise CUDA parity tests)."""
        q = _f32(query)
        D, nl = self.padded_dim, self.nlist
        npb = min(max(nprobe, 1), nl)
        ids = np.zeros(max(top_k, 1), np.uint64)
        scores = np.zeros(max(top_k, 1), np.float32)
        cnt = C.c_uint32()
        rot = np.empty(D, np.float32)
        lut = np.empty(4 * D, np.uint8)
        sc = np.empty(6, np.float32)
        probe = np.zeros(npb, np.uint32)
        pf = np.zeros((npb, 3), np.float32)
        rc = lib().orc_search_dump(self.h, _p(q), _sz(top_k), _sz(nprobe), _p(ids), _p(scores), C.byref(cnt),
                                   _p(rot), _p(lut), _p(sc), _p(probe), _p(pf))
        if rc != 0:
            raise OracleError(rc, "search failed")
        return dict(ids=ids[:cnt.value], scores=scores[:cnt.value], rotated=rot, lut=lut, delta=sc[0],
                    sum_vl=sc[1], k1x=sc[2], kbx=sc[3], qnorm=sc[4], sum_q=sc[5], probe=probe, probe_f=pf)


class BruteForceIndex:
    """Oracle restatement of BruteForceRabitqIndex (src/brute_force.rs): train / save ("RBF1" v1) / load / search."""

    def __init__(self):
        L = lib()
        L.orc_bf_new.restype = C.c_void_p
        L.orc_bf_len.restype = C.c_size_t
        L.orc_bf_padded_dim.restype = C.c_size_t
        L.orc_bf_save.restype = C.c_size_t
        self.h = C.c_void_p(L.orc_bf_new())

    def __del__(self):
        try:
            lib().orc_bf_free(self.h)
        except Exception:
            pass

    def __len__(self):
        return int(lib().orc_bf_len(self.h))

    @property
    def padded_dim(self):
        return int(lib().orc_bf_padded_dim(self.h))

    @classmethod
    def train(cls, data, total_bits, metric, rotator_type=1, seed=42, faster_config=False, rotator_bytes=None):
        data = _f32(data)
        n, dim = data.shape
        if rotator_bytes is None:
            rotator_bytes = make_flip_bytes(dim, seed) if rotator_type == 1 else make_matrix_bytes(dim, seed)
        rotator_bytes = np.ascontiguousarray(rotator_bytes, np.uint8)
        t_const = -1.0
        if faster_config and total_bits > 1:
            t_const = const_scaling_factor(padded_dim(rotator_type, dim), total_bits - 1, seed)
        self = cls()
        rc = lib().orc_bf_train(self.h, _p(data), _sz(n), _sz(dim), int(total_bits), int(metric), int(rotator_type), _p(rotator_bytes),
                                C.c_float(t_const))
        if rc != 0:
            raise OracleError(rc, "invalid configuration")
        return self

    def save_bytes(self):
        n = lib().orc_bf_save(self.h, None, _sz(0))
        buf = np.empty(n, np.uint8)
        lib().orc_bf_save(self.h, _p(buf), _sz(n))
        return buf.tobytes()

    @classmethod
    def load_bytes(cls, blob):
        self = cls()
        a = np.frombuffer(bytes(blob), np.uint8)
        rc = lib().orc_bf_load(self.h, _p(a), _sz(a.size))
        if rc != 0:
            raise OracleError(rc, lib().orc_last_error().decode())
        return self

    def search_batch(self, queries, top_k, filter_bits=None):
        q = _f32(queries)
        if q.ndim == 1:
            q = q[None, :]
        nq, dim = q.shape
        k = max(int(top_k), 1)
        ids = np.zeros((nq, k), np.uint64)
        scores = np.zeros((nq, k), np.float32)
        counts = np.zeros(nq, np.uint32)
        fb, fn = None, 0
        if filter_bits is not None:
            fb = np.ascontiguousarray(filter_bits, np.uint64)
            fn = fb.size * 64
        rc = lib().orc_bf_search_batch(self.h, _p(q), _sz(nq), _sz(dim), _sz(top_k), _p(fb), _sz(fn), _p(ids), _p(scores), _p(counts))
        if rc != 0:
            raise OracleError(rc, "search failed")
        return ids[:, :top_k], scores[:, :top_k], counts
