#!/usr/bin/env python
"""bench.py -- batched IVF+RaBitQ search throughput on synthetic data of the BASELINE.json shapes.

Metric (BASELINE.json): batched QPS at recall@10 >= 0.95 (GIST-1M-shape synthetic), plus the scan's roofline.
A "step" = one pass of the hot path (rotate+LUT -> coarse -> probe select -> scan/refine/top-k) over one
batch of queries.

  value : whole-job QPS with queries and outputs resident in HBM (CUDA-event timed)
  e2e   : the same through the reference-facing C ABI (rbq_search_batch) with pinned HOST buffers,
          H2D/D2H copies inside the timed region
  roofline      : dominant kernel (tail_tc_kernel, tensor i8) + the whole scan stage on the HBM roof
                  (algorithmic bytes = blocks scanned x (4D+384))
  cpu_baseline  : the CPU oracle (restated reference, "port") on the box's host cores, bounded sample

`--impl reference` times the CPU oracle alone (the Rust reference cannot be built here: no cargo).
Launch: python bench.py [--gpus N --steps K --warmup W]; for N>1 via torch.distributed.run (one rank per GPU):
inverted lists are sharded size-balanced across ranks and the batch runs the phased search
(rabitq_rs_b200.distributed.ShardedSearcher): probe selection for a query slice per rank (all-gather), head pass
on the shard that owns a query's nearest list (MIN all-reduce of the thresholds), tail + replay on every shard's
lists, local top-k all-gathered over NCCL and merged on the device.

Workloads: gist1m (default; BASELINE config 3, the configuration the metric is quoted on), gist1m_b3, sift1m
(config 2), quick (config 1), emb10m (config 4), deep10m / deep100m (config 5 and a stepping stone).  The last
three are STREAMED: the data never exists as one array -- every rank regenerates seeded chunks on its GPU, k-means
and the quantiser run on the device (rbq_kmeans_*, rbq_builder_*), each rank keeps only its list shard.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: n, dim, nlist, total_bits, metric, nq, top_k
    "gist1m": dict(n=1_000_000, dim=960, nlist=4096, total_bits=7, metric=0, nq=10_000, top_k=10),
    "gist1m_b3": dict(n=1_000_000, dim=960, nlist=4096, total_bits=3, metric=0, nq=10_000, top_k=10),
    "sift1m": dict(n=1_000_000, dim=128, nlist=4096, total_bits=7, metric=0, nq=10_000, top_k=10),
    "quick": dict(n=10_000, dim=128, nlist=256, total_bits=7, metric=0, nq=1_000, top_k=10),
    "small": dict(n=100_000, dim=960, nlist=512, total_bits=7, metric=0, nq=2_000, top_k=10),
    # streamed (device-resident build, list shards built in place)
    "deep1m": dict(n=1_000_000, dim=128, nlist=4096, total_bits=7, metric=0, nq=20_000, top_k=10, streamed=True, n_centers=1024),
    "deep10m": dict(n=10_000_000, dim=128, nlist=16384, total_bits=7, metric=0, nq=50_000, top_k=10, streamed=True, n_centers=4096),
    "deep100m": dict(n=100_000_000, dim=128, nlist=65536, total_bits=7, metric=0, nq=100_000, top_k=10, streamed=True, n_centers=16384),
    "emb10m": dict(n=10_000_000, dim=768, nlist=16384, total_bits=5, metric=1, nq=10_000, top_k=10, streamed=True, n_centers=4096),
}
# Frozen generator (SURVEY.md 8d "tune once, then freeze"): points live near a `latent`-dimensional
# random subspace (real descriptors have low intrinsic dimension), drawn from a Gaussian mixture there.
GEN = dict(latent=32, n_centers=256, center_scale=1.0, sigma=0.6, noise=0.05, seed_base=1234, seed_query=5678)
NPROBE_GRID = (1, 2, 4, 8, 16, 32, 64, 128, 256, 512)
CHUNK = 1 << 20          # vectors per streamed chunk
GT_QUERIES = 2000        # streamed workloads: queries with exact ground truth (recall / nprobe choice)


def gen_points(n, dim, seed, device, gen=GEN):
    """Seeded synthetic data, generated on the GPU (plumbing) and returned as host float32."""
    import torch

    g = torch.Generator(device=device).manual_seed(gen["seed_base"])  # shared structure
    r = gen["latent"]
    centers = torch.randn(gen["n_centers"], r, generator=g, device=device) * gen["center_scale"]
    proj = torch.randn(r, dim, generator=g, device=device) / (r ** 0.5)
    gp = torch.Generator(device=device).manual_seed(seed)
    out = torch.empty((n, dim), dtype=torch.float32)
    for s in range(0, n, 1 << 18):
        m = min(1 << 18, n - s)
        u = torch.randint(0, gen["n_centers"], (m,), generator=gp, device=device)
        z = centers[u] + gen["sigma"] * torch.randn(m, r, generator=gp, device=device)
        x = z @ proj + gen["noise"] * torch.randn(m, dim, generator=gp, device=device)
        out[s:s + m] = x.cpu()
    return out.numpy()


class ChunkGen:
    """Streamed workloads: chunk i of the base set (or of the query set) is a pure function of (seed, i), so every rank can
    regenerate any chunk on its own GPU.  Same mixture family as gen_points, with more centres for the larger sets."""

    def __init__(self, wl, device):
        import torch

        self.wl, self.dev, self.r = wl, device, GEN["latent"]
        g = torch.Generator(device=device).manual_seed(GEN["seed_base"])
        self.centers = torch.randn(wl["n_centers"], self.r, generator=g, device=device) * GEN["center_scale"]
        self.proj = torch.randn(self.r, wl["dim"], generator=g, device=device) / (self.r ** 0.5)

    def chunk(self, seed, i, m):
        import torch

        gp = torch.Generator(device=self.dev).manual_seed((seed * 1_000_003 + i) & 0x7fffffffffff)
        u = torch.randint(0, self.wl["n_centers"], (m,), generator=gp, device=self.dev)
        z = self.centers[u] + GEN["sigma"] * torch.randn(m, self.r, generator=gp, device=self.dev)
        x = z @ self.proj + GEN["noise"] * torch.randn(m, self.wl["dim"], generator=gp, device=self.dev)
        if self.wl["metric"] == 1:
            x = x / x.norm(dim=1, keepdim=True)
        return x.contiguous()

    def base_chunks(self):
        n = self.wl["n"]
        return [(i, s, min(CHUNK, n - s)) for i, s in enumerate(range(0, n, CHUNK))]

    def base(self, i, m):
        return self.chunk(GEN["seed_base"] + 1, i, m)

    def queries(self):
        import torch

        nq = self.wl["nq"]
        return torch.cat([self.chunk(GEN["seed_query"], i, min(CHUNK, nq - s)) for i, s in enumerate(range(0, nq, CHUNK))])


def ground_truth(base, queries, k, device, metric):
    import torch

    q = torch.from_numpy(queries).to(device)
    best_d = torch.full((q.shape[0], k), float("inf"), device=device)
    best_i = torch.zeros((q.shape[0], k), dtype=torch.int64, device=device)
    qn = (q * q).sum(1, keepdim=True)
    for s in range(0, base.shape[0], 1 << 17):
        x = torch.from_numpy(base[s:s + (1 << 17)]).to(device)
        d = -(q @ x.T) if metric == 1 else qn - 2.0 * (q @ x.T) + (x * x).sum(1)[None, :]
        cd = torch.cat([best_d, d], 1)
        ci = torch.cat([best_i, torch.arange(s, s + x.shape[0], device=device)[None, :].expand(q.shape[0], -1)], 1)
        best_d, idx = cd.topk(k, dim=1, largest=False)
        best_i = ci.gather(1, idx)
    return best_i.cpu().numpy()


def recall_at_k(ids, gt):
    k = gt.shape[1]
    return float(np.mean([len(set(ids[i, :k].tolist()) & set(gt[i].tolist())) / k for i in range(gt.shape[0])]))


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons sampled through NVML while the timed region runs (the same counters
    as the nvidia-smi clocks line of B200_PROFILING.md, polled every ~2 ms so short regions are covered)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.sm, self.reasons, self.stop_flag, self.max_mhz, self.window = gpu_index, [], set(), False, None, None

    def run(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            while not self.stop_flag:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.sm.append((time.time(), float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), r))
                time.sleep(0.002)
        except Exception as e:  # NVML missing: report no samples rather than failing the bench
            self.err = repr(e)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=2.0)
        rows = self.sm
        if self.window:  # keep the samples taken inside the timed region (fall back to all if none landed in it)
            inside = [r for r in rows if self.window[0] <= r[0] <= self.window[1]]
            rows = inside or rows
        import pynvml as nv

        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted({n for _, _, r in rows for n, bit in names.items() if r & bit})
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(rows)}


def build_index(wl, device_index, log):
    """Host-array workloads: data on the host, k-means + quantiser on the GPU (rbq_kmeans_*, rbq_index_build)."""
    import torch
    import rabitq_rs_b200 as rbq
    from rabitq_rs_b200.kmeans import kmeans_gpu

    dev = torch.device("cuda", device_index)
    t0 = time.time()
    base = gen_points(wl["n"], wl["dim"], GEN["seed_base"] + 1, dev)
    queries = gen_points(wl["nq"], wl["dim"], GEN["seed_query"], dev)
    log(f"data {base.shape} generated in {time.time() - t0:.1f}s")
    t0 = time.time()
    cents, assign = kmeans_gpu(base, wl["nlist"], iters=8, seed=42, device=device_index)
    log(f"k-means nlist={wl['nlist']} in {time.time() - t0:.1f}s (deterministic: rbq_kmeans_device)")
    t0 = time.time()
    ix = rbq.IvfRabitqIndex(wl["dim"], wl["metric"], device=device_index)
    ix.fit_with_clusters(base, cents, assign, wl["total_bits"], "fht", seed=42, faster_config=True)
    log(f"index built on GPU in {time.time() - t0:.1f}s ({len(ix)} vectors, padded_dim {ix.padded_dim})")
    t0 = time.time()
    gt = ground_truth(base, queries, wl["top_k"], dev, wl["metric"])
    log(f"ground truth in {time.time() - t0:.1f}s")
    return ix, base, queries, gt


def build_streamed(wl, rank, world, dev, log):
    """Streamed workloads: returns (index shard of this rank, queries [nq, dim] CUDA tensor, gt ids [GT_QUERIES, k] numpy).
    Pass 1 (chunks round-robin over the ranks): nearest centroid of every vector + exact ground truth of the first
    GT_QUERIES queries; pass 2 (every rank, all chunks): quantise the vectors of the lists this rank owns."""
    import torch
    import torch.distributed as dist
    import rabitq_rs_b200 as rbq
    from rabitq_rs_b200.kmeans import assign_device, kmeans_device

    n, dim, nlist, k = wl["n"], wl["dim"], wl["nlist"], wl["top_k"]
    gen = ChunkGen(wl, dev)
    chunks = gen.base_chunks()
    t0 = time.time()
    nt = min(n, 64 * nlist)  # k-means training subset: the leading chunks (the stream is i.i.d.)
    xt = torch.cat([gen.base(i, m) for i, s, m in chunks if s < nt])[:nt].contiguous()
    cents = kmeans_device(xt, nlist, iters=6, seed=42, max_points_per_centroid=64)
    del xt
    torch.cuda.synchronize(dev)
    log(f"k-means on {nt} points, nlist={nlist}: {time.time() - t0:.1f}s")
    t0 = time.time()
    queries = gen.queries()
    ngt = min(GT_QUERIES, wl["nq"])
    qg = queries[:ngt]
    qn = (qg * qg).sum(1, keepdim=True)
    best_d = torch.full((ngt, k), float("inf"), device=dev)
    best_i = torch.zeros((ngt, k), dtype=torch.int64, device=dev)
    assign = torch.zeros(n, dtype=torch.int32, device=dev)
    for i, s, m in chunks:
        if i % world != rank:
            continue
        x = gen.base(i, m)
        assign_device(x, cents, out=assign[s:s + m])
        d = -(qg @ x.T) if wl["metric"] == 1 else qn - 2.0 * (qg @ x.T) + (x * x).sum(1)[None, :]
        cd, ci = d.topk(k, dim=1, largest=False)
        cd = torch.cat([best_d, cd], 1)
        ci = torch.cat([best_i, ci + s], 1)
        best_d, idx = cd.topk(k, dim=1, largest=False)
        best_i = ci.gather(1, idx)
    if world > 1:
        dist.all_reduce(assign)  # every rank filled its own chunks, zeros elsewhere
        gd = torch.empty((world, ngt, k), device=dev)
        gi = torch.empty((world, ngt, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gd.view(-1), best_d.view(-1))
        dist.all_gather_into_tensor(gi.view(-1), best_i.contiguous().view(-1))
        cd, ci = gd.permute(1, 0, 2).reshape(ngt, -1), gi.permute(1, 0, 2).reshape(ngt, -1)
        best_d, idx = cd.topk(k, dim=1, largest=False)
        best_i = ci.gather(1, idx)
    sizes = torch.bincount(assign, minlength=nlist).cpu().numpy().astype(np.uint32)
    torch.cuda.synchronize(dev)
    log(f"pass 1 (assignment + ground truth of {ngt} queries): {time.time() - t0:.1f}s; list sizes min/mean/max = "
        f"{sizes.min()}/{sizes.mean():.0f}/{sizes.max()}")
    t0 = time.time()
    b = rbq.IndexBuilder(dim, cents.cpu().numpy(), sizes, wl["total_bits"], wl["metric"], "fht", seed=42, device=dev.index,
                         shard_rank=rank, shard_count=world, max_chunk=CHUNK)
    for i, s, m in chunks:
        b.add(gen.base(i, m), assign[s:s + m], s)
    ix = b.finish()
    del assign
    torch.cuda.synchronize(dev)
    log(f"pass 2 (rotate + quantise + pack, shard {rank}/{world}: {ix.local_len()} of {len(ix)} vectors): {time.time() - t0:.1f}s")
    return ix, queries, best_i.cpu().numpy()


def oracle_baseline(blob, queries, top_k, nprobe, seconds=12.0, log=lambda s: None, max_sample=None, oix=None):
    """CPU oracle (restated reference) over a bounded query sample, all host threads (OpenMP over
    queries == batch_search's rayon par_iter).  Returns (qps, cores, sample, results, oracle index); pass the returned index
    back as `oix` to time further steps without parsing the RBQ1 bytes again (the load is never inside the timed region)."""
    from oracle import oracle as orc

    orc.set_num_threads(os.cpu_count() or 1)  # launchers (torchrun) export OMP_NUM_THREADS=1: use the host's cores anyway
    if oix is None:
        t0 = time.time()
        oix = orc.Index.load_bytes(blob)
        log(f"oracle loaded the same RBQ1 bytes in {time.time() - t0:.1f}s ({orc.num_threads()} threads, SIMD {orc.simd_level()})")
    nthreads = orc.num_threads()
    nmax = queries.shape[0] if max_sample is None else min(max_sample, queries.shape[0])
    probe = min(4 * nthreads, nmax)
    t0 = time.time()
    oix.search_batch(queries[:probe], top_k, nprobe)
    per_q = (time.time() - t0) / probe
    sample = int(min(nmax, max(probe, seconds / max(per_q, 1e-9))))
    sample = min(nmax, max(nthreads, sample // nthreads * nthreads))
    t0 = time.time()
    res = oix.search_batch(queries[:sample], top_k, nprobe)
    dt = time.time() - t0
    return sample / dt, nthreads, sample, res, oix


def covered_subset_blob(ix, queries_np, nprobe):
    """RBQ1 stream with only the lists the given queries probe (the CPU oracle then answers exactly those queries as on the
    full index): how the oracle is fed on workloads whose full index does not fit a host-side copy comfortably."""
    cids, _ = ix.debug_probe(queries_np, nprobe)
    keep = np.zeros(ix.cluster_count(), np.uint8)
    keep[np.unique(cids)] = 1
    return ix.save_lists_to_bytes(keep), int(keep.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("RBQ_BENCH_WORKLOAD", "gist1m"), choices=sorted(WORKLOADS))
    ap.add_argument("--recall", type=float, default=0.95)
    ap.add_argument("--nprobe", type=int, default=0, help="skip the recall sweep and use this nprobe")
    ap.add_argument("--sweep", action="store_true", help="time every nprobe of the grid (recall@10 vs QPS table in the JSON line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scan-mode", type=int, default=int(os.environ.get("RBQ_SCAN_MODE", "0")), help="0 auto, 1 sequential, 2 list-major")
    ap.add_argument("--coarse-mode", type=int, default=int(os.environ.get("RBQ_COARSE_MODE", "-1")), help="-1 auto, 0 exact, 1 dense TC, 2 filtered TC")
    ap.add_argument("--coarse-terms", type=int, default=int(os.environ.get("RBQ_COARSE_TERMS", "0")), help="bf16 terms of the coarse GEMM (3, 1, 0 = auto)")
    ap.add_argument("--save-ids", default="", help="write the (merged) result ids of the timed configuration to this .npy")
    ap.add_argument("--check-ids", default="", help="compare the (merged) result ids with this .npy (e.g. of a 1-GPU run of the same workload)")
    ap.add_argument("--oracle-sample", type=int, default=1024, help="queries checked against the CPU oracle on the same index bytes")
    ap.add_argument("--no-exact-merge", action="store_true", help="multi-GPU: skip the extra exact-merge measurement")
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: everything else that libraries print there (NCCL's version banner, ...) is sent to
    # stderr by pointing fd 1 at fd 2 for the duration of the run; the result goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    wl = dict(WORKLOADS[args.workload])
    streamed = bool(wl.get("streamed"))
    log = (lambda s: print(f"[bench r{rank}] {s}", file=sys.stderr, flush=True)) if (args.verbose or rank == 0) else (lambda s: None)

    if args.impl == "reference" and rank != 0:
        return  # rank 0 alone runs the CPU arm
    if args.impl == "reference":
        world = 1  # the CPU arm is one process whatever the launcher started
    multi = world > 1
    if multi:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries the one JSON line only
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    import rabitq_rs_b200 as rbq

    nq, k = wl["nq"], wl["top_k"]
    ix_full = base = None
    blob = None
    if streamed:
        # every rank builds its list shard in place; recall is measured on the first GT_QUERIES queries
        ix, dq, gt = build_streamed(wl, rank, world, dev, log)
        queries = None
        D = ix.padded_dim
    else:
        # Rank 0 generates the data, builds the index once and picks nprobe; the other ranks receive the
        # serialized RBQ1 stream and the queries (one index file, every rank loads its shard of it).
        if rank == 0:
            ix_full, base, queries, gt = build_index(wl, local, log)
            D = ix_full.padded_dim
        if multi:
            meta = torch.zeros(2, dtype=torch.int64, device=dev)
            if rank == 0:
                blob = ix_full.save_to_bytes()
                meta[:] = torch.tensor([len(blob), D], dtype=torch.int64)
            dist.broadcast(meta, 0)
            nbytes, D = int(meta[0]), int(meta[1])
            tb = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            tq = torch.empty((nq, wl["dim"]), dtype=torch.float32, device=dev)
            if rank == 0:
                tb.copy_(torch.frombuffer(bytearray(blob), dtype=torch.uint8))
                tq.copy_(torch.from_numpy(queries))
            dist.broadcast(tb, 0)
            dist.broadcast(tq, 0)
            if rank != 0:
                blob = tb.cpu().numpy().tobytes()
                queries = tq.cpu().numpy()
            del tb, tq
            ix = rbq.IvfRabitqIndex.load_from_bytes(blob, device=local, shard_rank=rank, shard_count=world)
            log(f"shard {rank}/{world}: {ix.local_len()} of {len(ix)} vectors")
        else:
            ix = ix_full
        dq = torch.from_numpy(queries).to(dev)

    import ctypes as C
    from rabitq_rs_b200 import _ffi

    ix.set_scan_mode(args.scan_mode)
    ix.set_coarse_mode(args.coarse_mode)
    ix.set_coarse_terms(args.coarse_terms)
    d_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    d_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_cn = torch.empty(nq, dtype=torch.int32, device=dev)
    searcher = None
    if multi:
        from rabitq_rs_b200.distributed import ShardedSearcher

        searcher = ShardedSearcher(ix, rank, world)
        if os.environ.get("RBQ_BENCH_PY_COLLECTIVES") is None:  # A/B knob: the three-call form with torch.distributed collectives
            searcher.init_comm()                                # default: librbq's own NCCL communicator, one C call per batch
    stream = torch.cuda.current_stream(dev)

    def search_device(npb):
        """One batch through the device-resident entry; returns the (merged) result tensors."""
        if multi:
            return searcher.search(dq, k, npb)
        ix.batch_search_device(dq, k, npb, d_ids, d_sc, d_cn)
        return d_ids, d_sc, d_cn

    def barrier():
        if multi:
            dist.barrier()
        torch.cuda.synchronize(dev)

    ngt = gt.shape[0] if (streamed or rank == 0) else 0

    def measure_recall(npb):
        r_ids, _, _ = search_device(npb)
        torch.cuda.synchronize(dev)
        if not (streamed or rank == 0):
            return 0.0
        return recall_at_k(r_ids[:ngt].cpu().numpy().astype(np.uint64), gt)

    def bcast_float(v):
        if not multi:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.broadcast(t, 0)
        return float(t.item())

    # ---- nprobe: the smallest of the grid reaching the recall target (every rank runs the same searches) ----
    table = []
    if args.nprobe:
        nprobe = args.nprobe
    else:
        nprobe = None
        for npb in NPROBE_GRID:
            if npb > wl["nlist"]:
                break
            r = bcast_float(measure_recall(npb))
            table.append((npb, round(r, 4)))
            if r >= args.recall:
                nprobe = npb
                break
        nprobe = nprobe or table[-1][0]
        log(f"recall@{k} sweep: {table}")
    recall = bcast_float(measure_recall(nprobe))

    config = {"workload": f"{args.workload}: {wl['n']}x{wl['dim']} synthetic (latent-{GEN['latent']} Gaussian mixture), nlist={wl['nlist']}, "
                          f"total_bits={wl['total_bits']}, {'L2' if wl['metric'] == 0 else 'IP'}, FhtKacRotator, {nq}-query batch, top-{k}",
              "nprobe": nprobe, "recall_at_10": round(recall, 4), "recall_target": args.recall, "padded_dim": D,
              "recall_queries": ngt if streamed else nq,
              "l2_policy": "256 MiB L2 flush between timed steps; scanned blocks+ex-codes also exceed the 126 MB L2",
              "parallelism": f"lists sharded over {world} GPU(s), centroids replicated" if world > 1 else "1 GPU",
              "generator": dict(GEN, n_centers=wl.get("n_centers", GEN["n_centers"]), streamed=streamed)}
    metric_name = "batched QPS at recall@10>=0.95 (GIST-1M-shape synthetic)" if args.workload.startswith("gist1m") else \
        f"batched QPS at recall@10>=0.95 ({args.workload} synthetic)"

    if args.impl == "reference":
        # the CPU arm: the oracle on the same index bytes (streamed workloads: the lists its query sample probes), all host cores
        if streamed:
            qs = dq[:min(nq, 4096)].cpu().numpy()
            blob, kept = covered_subset_blob(ix, qs, nprobe)
            log(f"oracle index = the {kept} lists probed by the first {qs.shape[0]} queries ({len(blob) / 1e9:.2f} GB)")
        else:
            qs = queries
            blob = ix_full.save_to_bytes()
        config["parallelism"] = "CPU only (host cores)"
        qps_runs, info, oix_ref = [], None, None
        for s in range(args.warmup + args.steps):
            qps, cores, sample, _, oix_ref = oracle_baseline(blob, qs, k, nprobe, seconds=6.0, log=log if s == 0 else (lambda x: None), oix=oix_ref)
            info = (cores, sample)
            if s >= args.warmup:
                qps_runs.append(qps)
        v = float(np.mean(qps_runs))
        out = {"impl": "reference", "metric": metric_name, "value": v, "unit": "queries/s",
               "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * info[1] / v,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 LUT sums + f32", "data": "synthetic",
               "config": config,
               "cpu_baseline": {"value": v, "unit": "queries/s", "cores": info[0], "kind": "port",
                                "sample": f"{info[1]} of the {nq} queries per step; CPU oracle (C++ restatement of rabitq-rs search, -O3 -march=native, "
                                          f"AVX-512/AVX2 FastScan; the Rust crate cannot be built: no cargo)"},
               "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(out)
        return

    # ---- our arm: device-resident timing --------------------------------------------------------
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(npb, steps, warmup, profile):
        """(ms total max over ranks, accumulated stats dict) of `steps` batches at nprobe npb."""
        ix.set_profiling(profile)
        for _ in range(warmup):
            flush.fill_(1)
            search_device(npb)
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        acc = {}
        barrier()
        for s in range(steps):
            flush.fill_(s & 0xFF)
            ev[s][0].record(stream)
            search_device(npb)
            ev[s][1].record(stream)
            ev[s][1].synchronize()
            if profile:
                for kk, vv in ix.stats().items():
                    acc[kk] = acc.get(kk, 0) + vv
        barrier()
        ms = float(np.sum([a.elapsed_time(b) for a, b in ev]))
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if multi:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ix.set_profiling(False)
        return float(t.item()), acc

    sampler = ClockSampler(local)
    sampler.start()
    prof_range = os.environ.get("RBQ_CUDA_PROFILER") == "1"  # ncu --profile-from-start off: capture the timed steps only
    # warm-up, then the timed region (profiled: CUDA events around every stage on the launching stream)
    ix.set_profiling(True)
    for _ in range(args.warmup):
        flush.fill_(1)
        search_device(nprobe)
    barrier()
    if prof_range:
        torch.cuda.profiler.start()
    wall0 = time.time()
    ms_total, st = timed(nprobe, args.steps, 0, True)
    wall = time.time() - wall0
    if prof_range:
        torch.cuda.profiler.stop()
    sampler.window = (wall0, wall0 + wall)
    clocks = sampler.finish()
    value = nq * args.steps / (ms_total / 1000.0)
    steps = args.steps
    launches = st["kernel_launches"] + (steps if multi else 0)  # + the merge kernel of the phased search
    final_ids = search_device(nprobe)[0].cpu().numpy().astype(np.uint64)

    # multi-GPU: per-rank stage times (max over ranks) and scanned bytes (sum over ranks)
    stage_keys = ["ms_prep", "ms_coarse", "ms_select", "ms_scan", "ms_scan_head", "ms_scan_tail", "ms_scan_replay", "ms_tail_kernel"]
    sum_keys = ["bytes_scanned", "tail_bytes", "tail_pairs", "survivors", "overflow_queries", "fallback_queries", "candidates", "refined", "admitted",
                "coarse_fallbacks", "blocks_scanned", "tail_blocks"]
    if multi:
        tmax = torch.tensor([st[kk] for kk in stage_keys], dtype=torch.float64, device=dev)
        tsum = torch.tensor([float(st[kk]) for kk in sum_keys], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        for kk, vv in zip(stage_keys, tmax.tolist()):
            st[kk] = vv
        for kk, vv in zip(sum_keys, tsum.tolist()):
            st[kk] = vv

    # ---- multi-GPU: the exact-merge mode of the one-call sharded search (global replay at the home rank), timed beside the default
    exact_info = x_ids = None
    if multi and searcher.native and not args.no_exact_merge:
        ix.set_exact_merge(True)
        nx = max(3, steps // 2)
        ms_x, _ = timed(nprobe, nx, 2, False)
        x_ids = search_device(nprobe)[0].cpu().numpy().astype(np.uint64)
        sx = ix.stats()
        ix.set_exact_merge(False)
        tx = torch.tensor([float(sx["inexact_queries"]), float(sx["exchanged_records"])], dtype=torch.float64, device=dev)
        dist.all_reduce(tx, op=dist.ReduceOp.SUM)
        exact_info = {"ms_per_step": ms_x / nx, "qps": nq * nx / (ms_x / 1000.0), "inexact_queries": int(tx[0].item()),
                      "exchanged_records_per_step": int(tx[1].item()),
                      "note": "rbq_set_exact_merge(1): survivors refined eagerly, records sent to the query's home rank, one global replay"}
        if rank == 0 and ix_full is not None:  # rank 0 still holds the complete index: the single-GPU answer of the same batch
            ix_full.batch_search_device(dq, k, nprobe, d_ids, d_sc, d_cn)
            torch.cuda.synchronize(dev)
            one = d_ids.cpu().numpy().astype(np.uint64)
            exact_info["exact_ids_equal_single_gpu"] = float(np.mean(np.sort(one, 1) == np.sort(x_ids, 1)))
            exact_info["phased_ids_equal_single_gpu"] = float(np.mean(np.sort(one, 1) == np.sort(final_ids, 1)))
            exact_info["exact_queries_identical"] = int(np.sum(np.all(np.sort(one, 1) == np.sort(x_ids, 1), axis=1)))
        if rank == 0 and args.check_ids and os.path.exists(args.check_ids):
            ref_ids = np.load(args.check_ids)
            exact_info["exact_ids_equal_reference_run"] = float(np.mean(np.sort(ref_ids, 1) == np.sort(x_ids, 1)))
        log(f"exact merge: {exact_info}")

    # ---- optional recall-vs-QPS sweep (BASELINE config 3: "recall@10 vs QPS sweep") ----------------
    sweep = None
    if args.sweep:
        sweep = []
        for npb in NPROBE_GRID:
            if npb > wl["nlist"]:
                break
            r = bcast_float(measure_recall(npb))
            ns_ = max(3, steps // 4)
            ms_s, _ = timed(npb, ns_, 1, False)
            sweep.append({"nprobe": npb, "recall_at_10": round(r, 4), "qps": nq * ns_ / (ms_s / 1000.0)})
            log(f"sweep nprobe {npb}: recall {r:.4f}, {sweep[-1]['qps']:.0f} QPS")

    # ---- e2e through the C ABI with pinned host buffers ------------------------------------------
    hq = (dq.cpu() if queries is None else torch.from_numpy(queries)).pin_memory()
    h_ids = torch.empty((nq, k), dtype=torch.int64).pin_memory()
    h_sc = torch.empty((nq, k), dtype=torch.float32).pin_memory()
    h_cn = torch.empty(nq, dtype=torch.int32).pin_memory()
    L = _ffi.lib()

    def step_host():
        if multi:  # host buffers in, merged result out: H2D of the batch, phased search, D2H of the merged top-k
            a, b, c = searcher.search_host(hq, k, nprobe)  # each rank uploads 1/world of the batch, NVLink all-gather
            h_ids.copy_(a, non_blocking=True)
            h_sc.copy_(b, non_blocking=True)
            h_cn.copy_(c, non_blocking=True)
            torch.cuda.synchronize(dev)
            return
        rc = L.rbq_search_batch(ix.handle, C.c_void_p(hq.data_ptr()), nq, wl["dim"], k, nprobe, C.c_void_p(h_ids.data_ptr()),
                                C.c_void_p(h_sc.data_ptr()), C.c_void_p(h_cn.data_ptr()))
        assert rc == 0, _ffi.last_error()

    for _ in range(2):
        step_host()
    barrier()
    e2e_ms = 0.0
    for s in range(steps):
        flush.fill_(s & 0xFF)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        step_host()
        e2e_ms += (time.perf_counter() - t0) * 1000.0  # the call is synchronous: host wall clock == device + copies
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if multi:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = nq * steps / (float(t.item()) / 1000.0)
    host_ids = h_ids.numpy().astype(np.uint64)

    # ---- result identity across configurations (not timed) ---------------------------------------
    same_ids = None
    if rank == 0:
        if args.save_ids:
            np.save(args.save_ids, final_ids)
        if args.check_ids and os.path.exists(args.check_ids):
            ref_ids = np.load(args.check_ids)
            same_ids = float(np.mean(np.sort(ref_ids, 1) == np.sort(final_ids, 1)))
            log(f"ids equal to {args.check_ids}: {same_ids:.6f}")

    if rank != 0:
        dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bf16_peak = float(peaks.get("bf16_tflops", 1640.0))
    # ---- roofline --------------------------------------------------------------------------------
    # Dominant kernel of the list-major schedule: tail_tc_kernel = FastScan as an exact one-hot u8 GEMM on tcgen05 (kind::i8).
    # It is bound by the tensor pipe / its own operand production, NOT by HBM: the list-major grouping reads a list once per
    # <= 64 (query, list) pairs, so its DRAM traffic is ~1/25 of the algorithmic bytes.  Algorithmic ops = 2 x (vectors x pairs
    # evaluated) x 16 x (D/4) one-hot MACs; peak = 2 x the measured dense bf16 rate (i8 runs at twice bf16 on this part).
    # The number tracked against north_star's "scan >= 70 % of HBM peak" is the whole scan stage by algorithmic bytes (scan_stage).
    scan_bytes = st["bytes_scanned"] + st["tail_bytes"]
    scan_ms = st["ms_scan"]
    list_major = st["tail_bytes"] > 0
    stage_achieved = (scan_bytes / 1e9) / (scan_ms / 1000.0) if scan_ms > 0 else None
    traffic = ncu_pipe = None
    try:  # one `ncu --set full` capture of the dominant kernel (profiles/traffic.json): DRAM bytes and tensor-pipe activity
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(args.workload, {}).get("tail_tc_kernel" if list_major else "scan_kernel")
        if ent and ent.get("nprobe") == nprobe and world == 1:
            traffic = ent.get("dram_bytes_per_launch")
            ncu_pipe = ent.get("imma_pipe_active_frac")
    except Exception:
        pass
    if list_major:
        tail_ops = 2.0 * (st["tail_blocks"] * 32.0) * (16.0 * D / 4.0)  # vectors x pairs x K (one-hot columns)
        k_ms = st["ms_tail_kernel"]
        k_ach = tail_ops / 1e12 / (k_ms / 1000.0) if k_ms > 0 else None
        roofline = {"bound": "tensor", "kernel": "tail_tc_kernel (list-major FastScan: exact one-hot u8 GEMM on tcgen05 kind::i8 + distances + prune)",
                    "achieved": k_ach, "peak": 2.0 * bf16_peak, "unit": "TOP/s (i8)", "frac": (k_ach / (2.0 * bf16_peak)) if k_ach else None,
                    "peak_source": "2 x MEASURED_PEAKS.json bf16_tflops (i8 tensor rate = 2 x bf16; burst figure)" if "bf16_tflops" in peaks else "fallback 2 x 1640 TFLOP/s",
                    "ops_per_launch": tail_ops / steps, "ms_per_launch": k_ms / steps, "traffic": traffic,
                    "imma_pipe_active_frac_ncu": ncu_pipe,
                    "algorithmic_hbm": {"bytes_per_launch": st["tail_bytes"] / steps,
                                        "achieved_gbs": (st["tail_bytes"] / 1e9) / (k_ms / 1000.0) if k_ms > 0 else None,
                                        "note": "algorithmic bytes / kernel time exceeds the HBM peak because lists are re-used from L2/shared "
                                                "memory across the <= 64 pairs of a work item: HBM is not this kernel's roof"}}
    else:
        roofline = {"bound": "hbm", "kernel": "scan_kernel (FastScan accumulate + prune + refine + top-k)", "achieved": stage_achieved,
                    "peak": hbm_peak, "unit": "GB/s", "frac": (stage_achieved / hbm_peak) if stage_achieved else None,
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "bytes_per_launch": scan_bytes / steps, "ms_per_launch": scan_ms / steps, "traffic": traffic}
    roofline["scan_stage"] = {"bound": "hbm", "achieved": stage_achieved, "peak": hbm_peak, "unit": "GB/s",
                              "frac": (stage_achieved / hbm_peak) if stage_achieved else None,
                              "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                              "bytes_per_step": scan_bytes / steps, "ms_per_step": scan_ms / steps,
                              "kernels": "head_scan + resolve_head + tail_* + resolve_lazy/replay (+ fallback scan); "
                                         + ("per-rank time = max over ranks, bytes = sum over ranks" if multi else "1 GPU")}
    terms_used = int(st.get("coarse_terms_used", 0)) // max(steps, 1) or args.coarse_terms
    coarse_flops = 2.0 * nq * wl["nlist"] * D * terms_used
    out = {"metric": metric_name, "value": value, "unit": "queries/s", "n_gpus": world,
           "steps": steps, "warmup": args.warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "u8 LUT sums + f32", "data": "synthetic", "config": config,
           "e2e": {"value": e2e_value, "unit": "queries/s", "h2d_bytes_per_step": int(nq * wl["dim"] * 4),
                   "d2h_bytes_per_step": int(nq * k * 12 + nq * 4)},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
           "stage_ms_per_step": {n_: float(st[kk]) / steps for n_, kk in zip(("prep", "coarse", "select", "scan"), stage_keys[:4])},
           "coarse": {"mode": int(st.get("coarse_mode_used", 0)) // max(steps, 1), "terms": terms_used,
                      "front_chunk": int(st.get("front_chunk", 0)) // max(steps, 1),
                      "gemm_tflops_executed": coarse_flops / 1e12 / (st["ms_coarse"] / steps / 1000.0) if st["ms_coarse"] > 0 and not multi else None},
           "scan_split": {"schedule": "list-major (head/tail/replay)" if list_major else "sequential",
                          "ms_head": st["ms_scan_head"] / steps, "ms_tail": st["ms_scan_tail"] / steps, "ms_replay": st["ms_scan_replay"] / steps,
                          "ms_tail_kernel": st["ms_tail_kernel"] / steps,
                          "bytes_head": st["bytes_scanned"] / steps, "bytes_tail": st["tail_bytes"] / steps,
                          "tail_pairs": st["tail_pairs"] / steps, "survivors": st["survivors"] / steps,
                          "overflow_queries": st["overflow_queries"] / steps, "fallback_queries": st["fallback_queries"] / steps},
           "per_query": {"vectors_scanned": st["candidates"] / steps / nq, "refined": st["refined"] / steps / nq,
                         "admitted": st["admitted"] / steps / nq, "coarse_fallbacks": st["coarse_fallbacks"] / steps / nq},
           "host_ids_equal_device_ids": float(np.mean(host_ids == final_ids)),
           "wall_s_timed_region": wall}
    if table:
        out["config"]["recall_sweep"] = table
    if sweep:
        out["recall_qps_sweep"] = sweep
    if same_ids is not None:
        out["config"]["ids_equal_reference_run"] = round(same_ids, 6)
    if exact_info is not None:
        out["exact_merge"] = exact_info
    # ---- CPU baseline + parity of the timed configuration against the oracle ---------------------
    out["cpu_baseline"] = None
    if not args.no_cpu_baseline:
        ns = min(args.oracle_sample, nq)
        src = None
        if not streamed and not multi:  # the whole batch is available to the oracle: sample sized by time
            src = (blob or ix_full.save_to_bytes(), queries, wl["nlist"], None)
        elif not streamed:              # rank 0 still holds the complete index it built
            src = (blob, queries, wl["nlist"], ns)
        elif not multi:                 # streamed, one GPU: the lists the sample probes
            qs = dq[:ns].cpu().numpy()
            b_, kept = covered_subset_blob(ix, qs, nprobe)
            src = (b_, qs, kept, ns)
        # sharded streamed build: no rank holds the complete index (parity there = --check-ids against the 1-GPU run)
        if src is not None:
            qps, cores, sample, res, _ = oracle_baseline(src[0], src[1], k, nprobe, log=log, max_sample=src[3])
            same = float(np.mean(res[0][:, :k] == final_ids[:sample, :k]))
            same_x = ""
            if x_ids is not None:  # position-wise, and as per-query id sets (the order inside runs of bit-equal distances is class D1)
                sets_eq = float(np.mean(np.sort(res[0][:, :k], 1) == np.sort(x_ids[:sample, :k], 1)))
                same_x = f" (exact-merge mode: {float(np.mean(res[0][:, :k] == x_ids[:sample, :k])):.4f} position-wise, {sets_eq:.4f} as id sets)"
            out["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                   "sample": f"first {sample} of the {nq} queries, same index bytes"
                                             + (f" (the {src[2]} lists those queries probe)" if src[2] != wl["nlist"] else "")
                                             + f", nprobe={nprobe}; ids identical to the GPU result: {same:.4f}{same_x}"}
    emit(out)
    if multi:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
