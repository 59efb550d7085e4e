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
