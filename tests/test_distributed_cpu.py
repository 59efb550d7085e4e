"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the index distribution protocol, the
deterministic list -> shard map every rank derives independently, and the merge order.  The per-shard
search itself is emulated with the CPU oracle over each shard's lists (test infrastructure)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rabitq_rs_b200 as rbq
    from rabitq_rs_b200.distributed import broadcast_index, merge_topk_host
    from oracle import oracle as orc

    blob = queries = None
    if rank == 0:
        from helpers import oracle_index

        data, oix, blob = oracle_index(3000, 64, 32, 7, 0, kind="clustered")
        queries = data[:40].copy()
    blob, queries, nprobe, extra = broadcast_index(blob, queries, 8, [123], rank)
    assert nprobe == 8 and extra[0] == 123 and queries.shape == (40, 64)
    owner, sizes = rbq.shard_assignment(blob, world)
    # every rank derives the same map from the same bytes
    import torch

    t = torch.from_numpy(owner.astype(np.int64))
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    assert set(owner.tolist()) == set(range(world)) and int(sizes.sum()) == 3000
    load = [int(sizes[owner == r].sum()) for r in range(world)]
    assert max(load) - min(load) <= 0.1 * 3000 / world, load
    # emulate this rank's shard with the oracle: search the full index but keep only owned lists' vectors
    full = orc.Index.load_bytes(blob)
    mine = np.zeros(3000, bool)
    for c in np.flatnonzero(owner == rank):
        mine[full.list_ids(int(c)).astype(np.int64)] = True
    bits = rbq.ids_to_bitset(np.flatnonzero(mine), 3000)
    ids, sc, cnt = full.search_batch(queries, 10, nprobe, filter_bits=bits)
    k = 10
    gi = [torch.empty((40, k), dtype=torch.int64) for _ in range(world)]
    gs = [torch.empty((40, k), dtype=torch.float32) for _ in range(world)]
    gc = [torch.empty(40, dtype=torch.int32) for _ in range(world)]
    dist.all_gather(gi, torch.from_numpy(ids.astype(np.int64)))
    dist.all_gather(gs, torch.from_numpy(sc))
    dist.all_gather(gc, torch.from_numpy(cnt.astype(np.int32)))
    mi, ms, mc = merge_topk_host(np.stack([g.numpy() for g in gi]).astype(np.uint64), np.stack([g.numpy() for g in gs]),
                                 np.stack([g.numpy() for g in gc]), 0)
    ref = full.search_batch(queries, 10, nprobe)
    agree = np.mean([len(set(mi[i].tolist()) & set(ref[0][i].tolist())) / k for i in range(40)])
    assert agree >= 0.99, agree
    assert np.all(np.diff(ms, axis=1) >= 0)          # merged order = ascending distance
    assert np.all(ms <= ref[1] + 1e-6)               # never worse than the single-sequence answer
    np.save(os.path.join(tmp, f"ok{rank}.npy"), np.array([agree]))
    dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    from conftest import build_librbq

    build_librbq()
    from oracle import oracle as orc

    orc.build()
    port = 29500 + (os.getpid() % 400)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}.npy") for r in range(2))


def test_shard_assignment_is_a_balanced_partition(oracle):
    import rabitq_rs_b200 as rbq
    from helpers import oracle_index

    _, oix, blob = oracle_index(3000, 64, 32, 7, 0, kind="clustered")
    for world in (1, 2, 3, 8):
        owner, sizes = rbq.shard_assignment(blob, world)
        assert owner.shape == (32,) and owner.min() >= 0 and owner.max() < world
        assert [oix.list_len(c) for c in range(32)] == sizes.tolist()
        load = np.array([sizes[owner == r].sum() for r in range(world)])
        assert load.sum() == 3000 and load.max() - load.min() <= max(sizes.max(), 1)
    with pytest.raises(rbq.InvalidPersistence):
        rbq.shard_assignment(b"nope" + blob[4:], 2)


def _phased_worker(rank, world, port, tmp):
    """The exchange protocol of ShardedSearcher over gloo: probe slices all-gathered in place, tau MIN-reduced with +inf
    from the shards that do not own a query's nearest list, local top-k all-gathered and merged.  The per-shard stages
    are emulated with the CPU oracle (test infrastructure)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import rabitq_rs_b200 as rbq
    from rabitq_rs_b200.distributed import merge_topk_host, query_slices
    from helpers import oracle_index

    data, full, blob = oracle_index(3000, 64, 32, 7, 0, kind="clustered")  # same seed -> same bytes on every rank
    nq, k, nprobe = 70, 10, 8
    queries = data[:nq].copy()
    owner, sizes = rbq.shard_assignment(blob, world)
    per, slices = query_slices(nq, world, align=16)
    assert per % 16 == 0 and sum(c for _, c in slices) == nq and all(b == min(r * per, nq) for r, (b, c) in enumerate(slices))
    # phase 1: probe lists of my slice -> my rows of the padded buffer; in-place all-gather
    probes = torch.full((per * world, nprobe), -1, dtype=torch.int64)
    q0, qc = slices[rank]
    for i in range(q0, q0 + qc):
        probes[i] = torch.from_numpy(full.search_dump(queries[i], k, nprobe)["probe"].astype(np.int64))
    dist.all_gather_into_tensor(probes.view(-1), probes[rank * per:(rank + 1) * per].reshape(-1).clone())
    assert int(probes[:nq].min()) >= 0, "a slice is missing after the all-gather"
    ref_probe = np.stack([full.search_dump(queries[i], k, nprobe)["probe"] for i in range(nq)]).astype(np.int64)
    assert np.array_equal(probes[:nq].numpy(), ref_probe)
    # phase 2: the head pass runs where the nearest list lives: exactly one shard per query contributes a finite tau
    head_owner = owner[ref_probe[:, 0]] == rank
    mine = np.zeros(3000, bool)
    for c in np.flatnonzero(owner == rank):
        mine[full.list_ids(int(c)).astype(np.int64)] = True
    bits = rbq.ids_to_bitset(np.flatnonzero(mine), 3000)
    tau = torch.full((nq,), float("inf"))
    for i in np.flatnonzero(head_owner):  # k-th distance over the nearest list alone (nprobe = 1)
        ids1, sc1, cnt1 = full.search_batch(queries[i:i + 1], k, 1, filter_bits=bits)
        if cnt1[0] == k:
            tau[i] = float(sc1[0, k - 1])
    n_owner = torch.from_numpy(head_owner.astype(np.int64))
    dist.all_reduce(n_owner)
    assert torch.all(n_owner == 1)
    dist.all_reduce(tau, op=dist.ReduceOp.MIN)
    # phase 3: local top-k over my lists (the oracle's own thresholds are looser than tau: a superset), then gather + merge
    ids, sc, cnt = full.search_batch(queries, k, nprobe, filter_bits=bits)
    gi, gs, gc = (torch.empty((world, nq, k), dtype=torch.int64), torch.empty((world, nq, k), dtype=torch.float32),
                  torch.empty((world, nq), dtype=torch.int32))
    dist.all_gather_into_tensor(gi.view(-1), torch.from_numpy(ids.astype(np.int64)).view(-1))
    dist.all_gather_into_tensor(gs.view(-1), torch.from_numpy(sc).view(-1))
    dist.all_gather_into_tensor(gc.view(-1), torch.from_numpy(cnt.astype(np.int32)).view(-1))
    mi, ms, mc = merge_topk_host(gi.numpy().astype(np.uint64), gs.numpy(), gc.numpy(), 0)
    ref = full.search_batch(queries, k, nprobe)
    # tau bounds the final k-th distance of every query whose head list held k vectors
    finite = np.isfinite(tau.numpy())
    assert finite.mean() > 0.9 and np.all(ms[finite, k - 1] <= tau.numpy()[finite] + 1e-6)
    agree = np.mean([len(set(mi[i].tolist()) & set(ref[0][i].tolist())) / k for i in range(nq)])
    assert agree >= 0.99, agree
    np.save(os.path.join(tmp, f"phased{rank}.npy"), np.array([agree]))
    dist.destroy_process_group()


def test_phased_protocol_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    from conftest import build_librbq

    build_librbq()
    from oracle import oracle as orc

    orc.build()
    port = 29900 + (os.getpid() % 90)
    mp.spawn(_phased_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert all(os.path.exists(tmp_path / f"phased{r}.npy") for r in range(2))


def test_query_slices_cover_the_batch_exactly():
    """Slices are contiguous, disjoint, aligned to the GEMM row tile, and cover [0, nq) for any world size."""
    sys.path.insert(0, ROOT)
    from rabitq_rs_b200.distributed import query_slices

    for nq in (1, 7, 128, 129, 1000, 10000, 32768):
        for world in (1, 2, 3, 4, 8):
            per, slices = query_slices(nq, world)
            assert per % 128 == 0 and per * world >= nq and len(slices) == world
            covered = 0
            for r, (b, c) in enumerate(slices):
                assert b == min(r * per, nq) and 0 <= c <= per and b + c <= nq
                covered += c
            assert covered == nq
