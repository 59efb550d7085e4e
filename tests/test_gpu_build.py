"""GPU index builder (rbq_index_build == IvfRabitqIndex::train_with_clusters on the device) against
the oracle's restated quantiser: same rotator state and rescale constant in, byte-identical RBQ1 v3
stream out."""
import numpy as np
import pytest

from helpers import assert_results_match, clustered_data

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rbq():
    import torch

    assert torch.cuda.is_available()
    import rabitq_rs_b200 as r

    return r


CASES = [  # (n, dim, nlist, total_bits, metric, rotator, faster)
    (3000, 128, 32, 7, 0, 1, True),
    (3000, 128, 32, 3, 1, 1, True),
    (2000, 960, 16, 7, 0, 1, True),
    (2000, 960, 16, 3, 0, 1, True),
    (2000, 100, 16, 1, 0, 1, True),
    (1500, 768, 16, 5, 1, 1, True),    # total_bits=5: generic ex packing (parity unpinned in the reference)
    (800, 64, 8, 7, 0, 1, False),      # precise mode: per-vector best_rescale_factor
    (800, 48, 8, 3, 0, 0, True),       # MatrixRotator
    (1000, 1536, 8, 3, 0, 1, True),
]


@pytest.mark.parametrize("case", CASES)
def test_build_is_byte_identical_to_oracle(rbq, oracle, case):
    n, dim, nlist, bits, metric, rot, faster = case
    data = clustered_data(n, dim, max(nlist // 2, 1), 77, normalize=(metric == 1))
    cents, assign = oracle.kmeans(data, nlist, 5, 3)
    assign[assign == nlist - 1] = 0  # leave one list empty, make another ragged
    state = oracle.make_flip_bytes(dim, 9) if rot == 1 else oracle.make_matrix_bytes(dim, 9)
    oix = oracle.Index.train_with_clusters(data, cents, assign, bits, metric, rot, seed=42, faster_config=faster,
                                           rotator_bytes=state)
    gix = rbq.IvfRabitqIndex(dim, metric)
    gix.fit_with_clusters(data, cents, assign, bits, rot, seed=42, faster_config=faster, rotator_state=state)
    assert len(gix) == n and gix.cluster_count() == nlist and gix.padded_dim == oix.padded_dim
    gb, ob = gix.save_to_bytes(), oix.save_bytes()
    if gb != ob:
        a, b = np.frombuffer(gb, np.uint8), np.frombuffer(ob, np.uint8)
        assert a.size == b.size, "stream sizes differ"
        bad = np.flatnonzero(a != b)
        raise AssertionError(f"{bad.size} bytes differ, first at offset {bad[0]} of {a.size}")
    q = data[:50]
    assert_results_match(gix.batch_search(q, rbq.SearchParams(10, 8)), oix.search_batch(q, 10, 8))


def test_build_validation_errors(rbq):
    ix = rbq.IvfRabitqIndex(16)
    data = np.zeros((10, 16), np.float32)
    cents = np.zeros((2, 16), np.float32)
    ok = np.zeros(10, np.uint32)
    with pytest.raises(rbq.InvalidConfig, match="total_bits must be between 1 and 16"):
        ix.fit_with_clusters(data, cents, ok, total_bits=0)
    with pytest.raises(rbq.InvalidConfig, match="assignments reference invalid cluster ids"):
        ix.fit_with_clusters(data, cents, ok + 5)
    with pytest.raises(rbq.InvalidConfig, match="nlist cannot exceed number of vectors"):
        ix.fit_with_clusters(data[:1], cents, ok[:1])
    with pytest.raises(ValueError):
        ix.fit_with_clusters(data, cents, ok[:3])


def test_fit_end_to_end_recall(rbq):
    """fit (GPU k-means + GPU quantiser) -> search: recall@10 against exact search on clustered data."""
    import torch

    data = clustered_data(20000, 96, 64, 5)
    ix = rbq.IvfRabitqIndex(96, "euclidean")
    ix.fit(data, 128, total_bits=7, seed=1)
    assert len(ix) == 20000 and ix.cluster_count() == 128
    q = clustered_data(200, 96, 64, 5)[:200] + 0.01
    x = torch.from_numpy(data).cuda()
    d = torch.cdist(torch.from_numpy(q).cuda(), x)
    gt = d.topk(10, largest=False).indices.cpu().numpy()
    ids, sc, cnt = ix.batch_search(q, rbq.SearchParams(10, 32))
    recall = np.mean([len(set(ids[i].tolist()) & set(gt[i].tolist())) / 10 for i in range(200)])
    assert recall > 0.9, recall
    # save -> load keeps results (tests.rs:393-431)
    again = rbq.IvfRabitqIndex.load_from_bytes(ix.save_to_bytes())
    b = again.batch_search(q, rbq.SearchParams(10, 32))
    assert np.array_equal(ids, b[0]) and np.array_equal(sc, b[1])


# ---- streaming builder on device-resident chunks (rbq_builder_*) ------------------------------------------------------
@pytest.mark.parametrize("case", [(5000, 128, 40, 7, 0, 1), (3000, 96, 24, 3, 1, 1), (2500, 960, 12, 7, 0, 1), (2000, 64, 16, 1, 0, 0)])
def test_streaming_builder_is_byte_identical(rbq, oracle, case):
    """Chunks added through rbq_builder_add_device give the same RBQ1 bytes as the oracle's quantiser (and therefore as
    rbq_index_build), whatever the chunking."""
    import torch

    n, dim, nlist, bits, metric, rot = case
    data = clustered_data(n, dim, max(nlist // 2, 1), 78, normalize=(metric == 1))
    cents, assign = oracle.kmeans(data, nlist, 5, 3)
    assign[assign == nlist - 1] = 0  # an empty list
    state = oracle.make_flip_bytes(dim, 9) if rot == 1 else oracle.make_matrix_bytes(dim, 9)
    oix = oracle.Index.train_with_clusters(data, cents, assign, bits, metric, rot, seed=42, faster_config=True, rotator_bytes=state)
    ob = oix.save_bytes()
    x = torch.from_numpy(data).cuda()
    a = torch.from_numpy(assign.astype(np.int32)).cuda()
    sizes = np.bincount(assign, minlength=nlist).astype(np.uint32)
    for chunk in (n, 777):
        b = rbq.IndexBuilder(dim, cents, sizes, bits, metric, rot, seed=42, max_chunk=chunk, rotator_state=state)
        for s in range(0, n, chunk):
            b.add(x[s:s + chunk].contiguous(), a[s:s + chunk].contiguous(), s)
        gix = b.finish()
        assert len(gix) == n and gix.cluster_count() == nlist
        gb = gix.save_to_bytes()
        if gb != ob:
            u, v = np.frombuffer(gb, np.uint8), np.frombuffer(ob, np.uint8)
            assert u.size == v.size, "stream sizes differ"
            bad = np.flatnonzero(u != v)
            raise AssertionError(f"chunk {chunk}: {bad.size} bytes differ, first at offset {bad[0]} of {u.size}")
    # a shard built in place == the same shard loaded from the complete stream
    q = data[:64]
    for rank in range(3):
        b = rbq.IndexBuilder(dim, cents, sizes, bits, metric, rot, seed=42, max_chunk=1024, rotator_state=state, shard_rank=rank, shard_count=3)
        for s in range(0, n, 1024):
            b.add(x[s:s + 1024].contiguous(), a[s:s + 1024].contiguous(), s)
        built = b.finish()
        loaded = rbq.IvfRabitqIndex.load_from_bytes(ob, shard_rank=rank, shard_count=3)
        assert built.local_len() == loaded.local_len() and len(built) == n
        r1, r2 = built.batch_search(q, rbq.SearchParams(10, 8)), loaded.batch_search(q, rbq.SearchParams(10, 8))
        assert all(np.array_equal(u, v) for u, v in zip(r1, r2))
    # wrong announced sizes are reported at finish
    bad_sizes = sizes.copy()
    bad_sizes[0] += 1
    b = rbq.IndexBuilder(dim, cents, bad_sizes, bits, metric, rot, seed=42, max_chunk=n, rotator_state=state)
    b.add(x, a, 0)
    with pytest.raises(rbq.InvalidConfig, match="announced list sizes"):
        b.finish()


def test_subset_save_keeps_answers_of_covered_queries(rbq, oracle):
    """rbq_index_save_lists_mem: an RBQ1 stream with only the probed lists of a query sample gives the oracle the full
    index's answers for those queries (how parity is checked on indexes too large to hand to the CPU whole)."""
    from oracle import oracle as orc

    data = clustered_data(8000, 64, 32, 5)
    ix = rbq.IvfRabitqIndex(64, "euclidean")
    ix.fit(data, 64, total_bits=7, seed=1, kmeans_iters=4)
    q = data[:40] + 0.01
    nprobe = 6
    cids, _ = ix.debug_probe(q, nprobe)
    keep = np.zeros(64, np.uint8)
    keep[np.unique(cids)] = 1
    assert 0 < keep.sum() < 64
    sub = orc.Index.load_bytes(ix.save_lists_to_bytes(keep))
    full = orc.Index.load_bytes(ix.save_to_bytes())
    got = ix.batch_search(q, rbq.SearchParams(10, nprobe))
    assert assert_results_match(got, sub.search_batch(q, 10, nprobe)) == 40
    assert assert_results_match(got, full.search_batch(q, 10, nprobe)) == 40


# ---- k-means on the device (rbq_kmeans_device / rbq_kmeans_assign_device) ---------------------------------------------
@pytest.mark.parametrize("geom", [(20000, 96, 300), (6000, 100, 64), (50000, 128, 2048)])
def test_kmeans_assignment_and_update_match_reference_steps(rbq, geom):
    import torch
    from oracle import kmeans_ref
    from rabitq_rs_b200.kmeans import assign_device, kmeans_device

    n, dim, k = geom
    data = clustered_data(n, dim, max(k // 3, 2), 9)
    x = torch.from_numpy(data).cuda()
    c0 = kmeans_device(x, k, iters=0, seed=7)            # Forgy initialisation only
    assert torch.equal(c0, kmeans_device(x, k, iters=0, seed=7))
    init = c0.cpu().numpy()
    assert all((data == init[j]).all(1).any() for j in range(0, k, max(1, k // 50))), "initial centroids must be data points"
    # assignment == the reference's arg-min, up to near-ties of the (fp32-class) distances
    a_g = assign_device(x, c0).cpu().numpy().view(np.uint32)
    a_r, d_r = kmeans_ref.assign(data, init)
    d_all = kmeans_ref.distances(data, init)
    diff = np.flatnonzero(a_g != a_r)
    assert diff.size <= n // 500, f"{diff.size} assignments differ"
    for i in diff:
        assert d_all[i, a_g[i]] - d_r[i] <= 1e-4 * max(d_r[i], 1e-6), "assignment differs away from a tie"
    # one Lloyd step (all points in the training subset) == the reference's update on the device's own assignment
    c1 = kmeans_device(x, k, iters=1, seed=7, max_points_per_centroid=n).cpu().numpy()
    assert torch.equal(torch.from_numpy(c1), kmeans_device(x, k, iters=1, seed=7, max_points_per_centroid=n).cpu()), "not deterministic"
    ref1 = kmeans_ref.update(data, a_g, d_all[np.arange(n), a_g], k, init)
    err = np.abs(c1 - ref1).max(1)
    moved = np.flatnonzero(err > 1e-4 * (1.0 + np.abs(ref1).max(1)))
    # clusters whose membership contains a near-tie point, or empty clusters re-seeded among equal distances, may differ
    assert moved.size <= max(2, k // 100), f"{moved.size} centroids differ from the reference update"
    # several iterations reduce the objective
    def objective(c):
        return float(kmeans_ref.assign(data, c)[1].sum())
    assert objective(kmeans_device(x, k, iters=5, seed=7).cpu().numpy()) < 0.9 * objective(init)
