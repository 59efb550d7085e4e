"""BruteForceRabitqIndex on the device (rbq_bf_*) against the oracle's restatement of
src/brute_force.rs: byte-identical RBF1 streams, bit-identical scores, the reference's own test properties."""
import numpy as np
import pytest

from conftest import rust_like_uniform

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rbq():
    import torch

    assert torch.cuda.is_available()
    import rabitq_rs_b200 as r

    return r


def _canon(ids, sc):
    out = ids.copy()
    i = 0
    while i < len(ids):
        j = i + 1
        while j < len(ids) and sc[j] == sc[i]:
            j += 1
        out[i:j] = np.sort(ids[i:j])
        i = j
    return out


CASES = [  # (n, dim, total_bits, metric, rotator)
    (900, 128, 7, 0, 1),
    (900, 96, 3, 1, 1),
    (700, 960, 7, 0, 1),
    (600, 64, 1, 0, 1),
    (500, 768, 5, 1, 1),   # generic ex packing
    (500, 48, 3, 0, 0),    # MatrixRotator
]


@pytest.mark.parametrize("case", CASES)
def test_bruteforce_matches_oracle(rbq, oracle, case):
    n, dim, bits, metric, rot = case
    data = rust_like_uniform(n, dim, 21)
    if metric == 1:
        data /= np.linalg.norm(data, axis=1, keepdims=True)
    state = oracle.make_flip_bytes(dim, 9) if rot == 1 else oracle.make_matrix_bytes(dim, 9)
    oix = oracle.BruteForceIndex.train(data, bits, metric, rot, seed=42, faster_config=True, rotator_bytes=state)
    gix = rbq.BruteForceRabitqIndex.train(data, bits, metric, rot, seed=42, use_faster_config=True, rotator_state=state)
    assert len(gix) == n
    ob, gb = oix.save_bytes(), gix.save_to_bytes()
    if ob != gb:
        a, b = np.frombuffer(gb, np.uint8), np.frombuffer(ob, np.uint8)
        assert a.size == b.size, "stream sizes differ"
        bad = np.flatnonzero(a != b)
        raise AssertionError(f"{bad.size} bytes differ, first at offset {bad[0]} of {a.size}")
    q = data[:40] + 0.01
    for k in (1, 10, 100):
        gi, gs, gc = gix.batch_search(q, rbq.BruteForceSearchParams(k))
        oi, osc, oc = oix.search_batch(q, k)
        assert np.array_equal(gc, oc)
        for i in range(q.shape[0]):
            m = int(oc[i])
            assert np.array_equal(gs[i, :m].view(np.uint32), osc[i, :m].view(np.uint32)), "scores not bit-identical"
            assert np.array_equal(_canon(gi[i, :m], gs[i, :m]), _canon(oi[i, :m], osc[i, :m]))
    # load of the oracle's stream gives the same answers (src/tests.rs brute_force_persistence_roundtrip)
    again = rbq.BruteForceRabitqIndex.load_from_bytes(ob)
    a, b = again.batch_search(q, rbq.BruteForceSearchParams(5)), gix.batch_search(q, rbq.BruteForceSearchParams(5))
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert again.save_to_bytes() == ob


def test_bruteforce_reference_properties(rbq):
    """src/tests.rs:912-1107: identical vectors are recovered at rank 1; a filter restricts the ids; empty filter -> nothing."""
    data = rust_like_uniform(400, 32, 3)
    ix = rbq.BruteForceRabitqIndex.train(data, 7, "euclidean", "random", seed=7)
    ids, sc, cnt = ix.batch_search(data[:64], rbq.BruteForceSearchParams(1))
    assert (ids[:, 0] == np.arange(64)).all()
    allowed = [3, 17, 99, 250]
    res = ix.search_filtered(data[0], rbq.BruteForceSearchParams(5), allowed)
    assert len(res) == 4 and set(i for i, _ in res) == set(allowed)
    assert ix.search_filtered(data[0], rbq.BruteForceSearchParams(5), []) == []
    assert ix.search(data[0], rbq.BruteForceSearchParams(0)) == []
    with pytest.raises(rbq.DimensionMismatch):
        ix.search(data[0, :16], rbq.BruteForceSearchParams(3))
    with pytest.raises(rbq.InvalidPersistence, match="unrecognized file header"):
        rbq.BruteForceRabitqIndex.load_from_bytes(b"XXXX" + bytes(64))
    blob = bytearray(ix.save_to_bytes())
    blob[-1] ^= 0xFF
    with pytest.raises(rbq.InvalidPersistence, match="checksum mismatch"):
        rbq.BruteForceRabitqIndex.load_from_bytes(bytes(blob))
