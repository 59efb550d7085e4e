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
