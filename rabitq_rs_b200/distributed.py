"""Multi-GPU plumbing (torch.distributed): one process per GPU, inverted lists sharded across ranks.

Data path per query batch: every rank runs rotate/LUT/coarse/probe-select on the whole batch (centroids
are replicated), scans only the probed lists it owns, and produces a local top-k; the local results are
all-gathered (NCCL over NVLink on GPUs, gloo in the CPU tests) and merged with the reference's result
order.  The reference itself is single-process (src/ivf.rs has no communication)."""
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
