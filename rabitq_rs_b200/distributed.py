"""Multi-GPU plumbing (torch.distributed): one process per GPU, inverted lists sharded across ranks.

Data path per query batch: every rank runs rotate/LUT/coarse/probe-select on the whole batch (centroids
are replicated), scans only the probed lists it owns, and produces a local top-k; the local results are
all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and merged with the reference's result
order.  The reference itself is single-process (src/ivf.rs has no communication).

`ShardedSearcher` is the scalable form of that data path (three phases, two small exchanges):
  1. every rank rotates all queries (its scan needs every LUT) but selects probe lists for its own SLICE of the
     batch only; the slices (16 B per probe) are all-gathered;
  2. the head pass of a query -- the sequential loop over its nearest list that fills the heap -- runs only on the
     shard that owns that list; the resulting thresholds tau (4 B per query) are all-reduced with MIN;
  3. every shard prunes its remaining (query, owned list) pairs with tau, keeps a local top-k, and the local
     results are all-gathered and merged on the device.
Per-rank work then shrinks with the number of shards in every stage but the query rotation."""
import numpy as np


def broadcast_index(blob, queries, nprobe, extra, rank, device=None):
    """Rank 0 holds (blob bytes, queries float32 [nq, dim], nprobe, extra ints); everyone returns them.
    One serialized RBQ1 stream is the unit of distribution: each rank then loads its own shard of it."""
    import torch
    import torch.distributed as dist

    dev = device if device is not None else torch.device("cpu")
    meta = torch.zeros(4 + 4, dtype=torch.int64, device=dev)
    if rank == 0:
        ex = list(extra)[:4] + [0] * (4 - len(list(extra)[:4]))
        meta[:] = torch.tensor([len(blob), queries.shape[0], queries.shape[1], int(nprobe)] + [int(x) for x in ex], dtype=torch.int64)
    dist.broadcast(meta, 0)
    nbytes, nq, dim, nprobe = (int(x) for x in meta[:4])
    tb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    tq = torch.empty((nq, dim), dtype=torch.float32, device=dev)
    if rank == 0:
        tb.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
        tq.copy_(torch.from_numpy(np.ascontiguousarray(queries, np.float32)))
    dist.broadcast(tb, 0)
    dist.broadcast(tq, 0)
    if rank != 0:
        blob = tb.cpu().numpy().tobytes()
        queries = tq.cpu().numpy()
    return blob, queries, nprobe, [int(x) for x in meta[4:]]


def merge_topk_host(ids, scores, counts, metric):
    """CPU statement of rbq_merge_topk_device (used by the gloo tests): ids/scores [nshards, nq, k], counts
    [nshards, nq] -> merged (ids, scores, counts).  Order: L2 ascending, IP descending; ties -> lower shard."""
    ns, nq, k = ids.shape
    out_i = np.full((nq, k), np.iinfo(np.uint64).max, np.uint64)
    out_s = np.zeros((nq, k), np.float32)
    out_c = np.zeros(nq, np.uint32)
    for q in range(nq):
        cand = []
        for s in range(ns):
            for j in range(int(counts[s, q])):
                key = scores[s, q, j] if metric == 0 else -scores[s, q, j]
                cand.append((key, s, j))
        cand.sort()
        n = min(k, len(cand))
        for t in range(n):
            _, s, j = cand[t]
            out_i[q, t] = ids[s, q, j]
            out_s[q, t] = scores[s, q, j]
        out_c[q] = n
    return out_i, out_s, out_c


def query_slices(nq, world, align=128):
    """Contiguous query slices of equal padded length (a multiple of `align`, the coarse GEMM's row tile):
    returns (per, [(begin, count)] * world); rank r writes rows [r*per, r*per+count) of the padded probe buffer."""
    per = (((nq + world - 1) // world) + align - 1) // align * align
    return per, [(min(r * per, nq), max(0, min((r + 1) * per, nq) - min(r * per, nq))) for r in range(world)]


class ShardedSearcher:
    """Phased multi-GPU search over one list shard per rank (rabitq_rs_b200.IvfRabitqIndex loaded with
    shard_rank/shard_count).  Buffers are allocated once per (nq, top_k, nprobe)."""

    def __init__(self, ix, rank, world, group=None):
        self.ix, self.rank, self.world, self.group = ix, rank, world, group
        self._key = None
        self.native = False

    def init_comm(self):
        """Give librbq its own NCCL communicator (rbq_comm_init): search()/search_host() then run as ONE C call per batch
        (rbq_search_batch_sharded[_device]: kernels and collectives enqueued back to back on one stream) instead of three
        C calls interleaved with torch.distributed collectives.  The 128-byte rendezvous id travels over torch.distributed."""
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _ffi
        from .index import _check

        dev = self.ix_device() if dist.get_backend(self.group) == "nccl" else torch.device("cpu")
        idt = torch.zeros(128, dtype=torch.uint8, device=dev)
        if self.rank == 0:
            buf = (C.c_uint8 * 128)()
            _check(_ffi.lib().rbq_comm_unique_id(buf))
            idt.copy_(torch.tensor(list(buf), dtype=torch.uint8))
        dist.broadcast(idt, 0, group=self.group)
        raw = (C.c_uint8 * 128)(*idt.cpu().tolist())
        _check(_ffi.lib().rbq_comm_init(self.ix.handle, raw, self.rank, self.world))
        self.native = True
        return self

    def _buffers(self, nq, k, nprobe, dev):
        import torch

        key = (nq, k, nprobe)
        if self._key != key:
            self.per, self.slices = query_slices(nq, self.world)
            self.probes = torch.zeros((self.per * self.world, nprobe, 4), dtype=torch.int32, device=dev)
            self.tau = torch.empty(nq, dtype=torch.float32, device=dev)
            # the shard's local top-k lives in one packed chunk (ids | scores | counts) so that ONE all-gather moves it
            self.chunk = (nq * k * 12 + nq * 4 + 15) // 16 * 16
            self.gathered = torch.zeros(self.world * self.chunk, dtype=torch.uint8, device=dev)
            loc = self.gathered[self.rank * self.chunk:(self.rank + 1) * self.chunk]
            self.local = loc
            self.l_ids = loc[:nq * k * 8].view(torch.int64).view(nq, k)
            self.l_sc = loc[nq * k * 8:nq * k * 12].view(torch.float32).view(nq, k)
            self.l_cn = loc[nq * k * 12:nq * k * 12 + nq * 4].view(torch.int32)
            self.m_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
            self.m_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
            self.m_cn = torch.empty(nq, dtype=torch.int32, device=dev)
            self._key = key
        return self

    def search(self, dq, k, nprobe):
        """dq: [nq, dim] float32 CUDA tensor (the whole batch, on every rank) -> merged (ids, scores, counts) tensors."""
        import torch.distributed as dist

        nq = dq.shape[0]
        # the C side clamps nprobe to [1, cluster_count] and strides the probe rows by the clamped value: size and slice
        # the exchange buffers with the same number
        nprobe = min(max(int(nprobe), 1), self.ix.cluster_count())
        b = self._buffers(nq, k, nprobe, dq.device)
        if self.native:
            import ctypes as C

            import torch

            from . import _ffi
            from .index import _check

            _check(_ffi.lib().rbq_search_batch_sharded_device(self.ix.handle, C.c_void_p(dq.data_ptr()), nq, dq.shape[1], int(k), nprobe,
                                                              C.c_void_p(b.m_ids.data_ptr()), C.c_void_p(b.m_sc.data_ptr()),
                                                              C.c_void_p(b.m_cn.data_ptr()),
                                                              C.c_void_p(torch.cuda.current_stream(dq.device).cuda_stream)))
            return b.m_ids, b.m_sc, b.m_cn
        q0, qc = b.slices[self.rank]
        self.ix.dist_front(dq, k, nprobe, q0, qc, b.probes)
        mine = b.probes[self.rank * b.per:(self.rank + 1) * b.per]
        dist.all_gather_into_tensor(b.probes.view(-1), mine.reshape(-1), group=self.group)
        self.ix.dist_head(nq, k, nprobe, b.probes, b.tau, b.l_ids, b.l_sc, b.l_cn)
        dist.all_reduce(b.tau, op=dist.ReduceOp.MIN, group=self.group)
        self.ix.dist_tail(nq, k, nprobe, b.tau, b.l_ids, b.l_sc, b.l_cn)
        dist.all_gather_into_tensor(b.gathered, b.local, group=self.group)  # in place: every rank owns one chunk
        self.ix.merge_topk_packed_device(self.world, nq, k, b.gathered, b.chunk, b.m_ids, b.m_sc, b.m_cn)
        return b.m_ids, b.m_sc, b.m_cn

    def search_host(self, hq, k, nprobe):
        """hq: [nq, dim] float32 PINNED host tensor holding the whole batch (every rank sees the same buffer, or at least
        its own slice of it).  Each rank uploads only its 1/world slice over PCIe and the slices are all-gathered over
        NVLink -- instead of every rank pulling the whole batch through its host link -- then the phased search runs."""
        import torch
        import torch.distributed as dist

        nq, dim = hq.shape
        if self.native:  # one synchronous C call: H2D of this rank's slice, NVLink all-gather, phased search, D2H of the merged result
            import ctypes as C

            from . import _ffi
            from .index import _check

            if getattr(self, "_hout", None) is None or self._hout[0].shape != (nq, k):
                self._hout = (torch.empty((nq, k), dtype=torch.int64).pin_memory(), torch.empty((nq, k), dtype=torch.float32).pin_memory(),
                              torch.empty(nq, dtype=torch.int32).pin_memory())
            a, b_, c = self._hout
            _check(_ffi.lib().rbq_search_batch_sharded(self.ix.handle, C.c_void_p(hq.data_ptr()), nq, dim, int(k), int(nprobe),
                                                       C.c_void_p(a.data_ptr()), C.c_void_p(b_.data_ptr()), C.c_void_p(c.data_ptr())))
            return a, b_, c
        per = (nq + self.world - 1) // self.world
        if getattr(self, "_dq", None) is None or self._dq.shape != (per * self.world, dim):
            self._dq = torch.empty((per * self.world, dim), dtype=torch.float32, device=self.ix_device())
        lo, hi = min(self.rank * per, nq), min((self.rank + 1) * per, nq)
        mine = self._dq[self.rank * per:(self.rank + 1) * per]
        if hi > lo:
            mine[:hi - lo].copy_(hq[lo:hi], non_blocking=True)
        dist.all_gather_into_tensor(self._dq.view(-1), mine.reshape(-1), group=self.group)
        return self.search(self._dq[:nq], k, nprobe)

    def ix_device(self):
        import torch

        return torch.device("cuda", self.ix.device)
