"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical index
bytes and queries.  Bar (BASELINE.json north_star): integer accumulations and probed-list
selection bit-exact; estimated distances within 1e-5 relative (asserted bit-exact wherever the
kernel reproduces the reference's float order); top-k ids identical except documented near-ties."""
import numpy as np
import pytest

from helpers import assert_results_match, oracle_index

pytestmark = pytest.mark.gpu

TOL = 1e-5  # north_star tolerance for estimated distances (relative)


@pytest.fixture(scope="module")
def rbq():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import rabitq_rs_b200 as r

    return r


def _load(rbq, blob, **kw):
    return rbq.IvfRabitqIndex.load_from_bytes(blob, **kw)


def _queries(data, nq, seed):
    rng = np.random.default_rng(seed)
    base = data[rng.integers(0, data.shape[0], nq)]
    return (base + 0.05 * rng.standard_normal(base.shape)).astype(np.float32)


GEOMS = [  # (n, dim, nlist, bits, metric, rotator)
    (600, 128, 8, 7, 0, 1),   # power-of-two FHT
    (600, 960, 8, 3, 0, 1),   # GIST geometry: trunc 512, window start 448
    (600, 768, 8, 7, 1, 1),   # trunc 512, start 256, inner product
    (600, 100, 8, 7, 0, 1),   # padded 128, trunc 64
    (400, 24, 8, 1, 1, 1),    # padded 64, 1-bit
    (400, 32, 8, 3, 0, 0),    # MatrixRotator
    (600, 1536, 4, 3, 0, 1),  # > 1024 dims: wide (32-bit) reduction with u16 wrap semantics
]


@pytest.mark.parametrize("geom", GEOMS)
def test_query_prep_bit_exact(rbq, oracle, geom):
    n, dim, nlist, bits, metric, rot = geom
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind="uniform11")
    gix = _load(rbq, blob)
    q = _queries(data, 9, 1)
    q[0] = 0.0  # degenerate: delta == 0 -> all-zero LUT
    rot_g, lut_g, sc_g = gix.debug_query_prep(q)
    for i in range(q.shape[0]):
        d = oix.search_dump(q[i], 5, 4)
        assert np.array_equal(rot_g[i].view(np.uint32), d["rotated"].view(np.uint32)), "rotation not bit-exact"
        assert np.array_equal(lut_g[i], d["lut"]), "LUT bytes differ"
        exp = np.array([d["delta"], d["sum_vl"], d["k1x"], d["kbx"], d["qnorm"], d["sum_q"]], np.float32)
        assert np.array_equal(sc_g[i, :6].view(np.uint32), exp.view(np.uint32)), "query scalars not bit-exact"


@pytest.mark.parametrize("geom", GEOMS)
def test_probe_selection_bit_exact(rbq, oracle, geom):
    n, dim, nlist, bits, metric, rot = geom
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind="uniform11")
    gix = _load(rbq, blob)
    q = _queries(data, 7, 2)
    for nprobe in (1, 3, nlist):
        cids, consts = gix.debug_probe(q, nprobe)
        for i in range(q.shape[0]):
            d = oix.search_dump(q[i], 5, nprobe)
            assert np.array_equal(cids[i], d["probe"]), "probed lists / order differ"
            assert np.array_equal(consts[i].view(np.uint32), d["probe_f"].view(np.uint32)), "g_add/g_error/dot_qc not bit-exact"


def test_probe_selection_ties_break_on_cluster_id(rbq, oracle):
    # duplicate centroids -> equal scores; the reference orders ties by cluster id (ivf.rs:1808-1823)
    from oracle import oracle as orc

    rng = np.random.default_rng(5)
    data = rng.standard_normal((400, 64)).astype(np.float32)
    cents = np.repeat(rng.standard_normal((4, 64)).astype(np.float32), 4, axis=0)  # 16 lists, 4 distinct centroids
    assign = (np.arange(400) % 16).astype(np.uint32)
    for metric in (0, 1):
        oix = orc.Index.train_with_clusters(data, cents, assign, 3, metric)
        gix = _load(rbq, oix.save_bytes())
        for nprobe in (2, 5, 9, 16):
            cids, _ = gix.debug_probe(data[:6], nprobe)
            for i in range(6):
                assert np.array_equal(cids[i], oix.search_dump(data[i], 3, nprobe)["probe"])


@pytest.mark.parametrize("geom", GEOMS)
def test_fastscan_accumulate_and_distances(rbq, oracle, geom):
    n, dim, nlist, bits, metric, rot = geom
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind="uniform11")
    gix = _load(rbq, blob)
    q = _queries(data, 1, 3)[0]
    d = oix.search_dump(q, 5, nlist)
    D = oix.padded_dim
    for c in (0, nlist - 1):
        nv = oix.list_len(c)
        accu, ip, est, lb = gix.debug_scan_list(q, c, nv)
        blocks = oix.list_blocks(c).reshape(-1, 4 * D + 384)
        rank = int(np.where(d["probe"] == c)[0][0])
        g_add, g_err = d["probe_f"][rank, 0], d["probe_f"][rank, 1]
        for b in range(blocks.shape[0]):
            ea = oracle.accumulate_block(blocks[b, :4 * D], d["lut"], D)
            assert np.array_equal(accu[b * 32:(b + 1) * 32].astype(np.uint16), ea), "integer accumulation differs"
            fac = blocks[b, 4 * D:].view(np.float32)
            eip, eest, elb = oracle.batch_distances(ea, d["delta"], d["sum_vl"], fac[:32], fac[32:64], fac[64:], g_add, g_err, d["k1x"])
            for got, exp in ((ip, eip), (est, eest), (lb, elb)):
                assert np.array_equal(got[b * 32:(b + 1) * 32].view(np.uint32), exp.view(np.uint32))


SEARCH_CASES = [  # (n, dim, nlist, bits, metric, rot, kind, k, nprobe)
    (10000, 128, 256, 7, 0, 1, "uniform01", 10, 32),   # BASELINE config 1 (README quickstart)
    (10000, 128, 256, 3, 0, 1, "uniform01", 10, 32),
    (10000, 128, 256, 1, 0, 1, "uniform01", 10, 32),
    (4000, 96, 64, 7, 1, 1, "clustered", 10, 16),      # inner product, normalised, non-pow2 FHT
    (3000, 960, 32, 3, 0, 1, "clustered", 10, 8),      # GIST geometry
    (3000, 960, 32, 7, 0, 1, "clustered", 100, 8),
    (2000, 32, 16, 7, 0, 0, "uniform11", 5, 16),       # MatrixRotator
    (2000, 1280, 16, 3, 0, 1, "clustered", 10, 4),     # wide path
    (3000, 256, 32, 7, 0, 1, "clustered", 10, 8),      # 32-byte code rows (paired refine, two rows per copy stripe)
    (3000, 512, 32, 7, 1, 1, "clustered", 10, 8),      # 64-byte code rows
    (3000, 640, 32, 3, 0, 1, "clustered", 20, 8),      # 80-byte code rows (irregular stripes)
    (3000, 768, 32, 5, 1, 1, "clustered", 10, 8),      # BASELINE config 4 geometry (768-d, inner product, total_bits 5:
                                                       # the reference panics on this width -- parity vs our oracle only)
]


@pytest.mark.parametrize("case", SEARCH_CASES)
def test_search_end_to_end(rbq, oracle, case):
    n, dim, nlist, bits, metric, rot, kind, k, nprobe = case
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind=kind)
    gix = _load(rbq, blob)
    assert len(gix) == n and gix.cluster_count() == nlist
    q = np.concatenate([data[:32], _queries(data, 224, 4)])
    exp = oix.search_batch(q, k, nprobe, want_diag=True)
    got = gix.batch_search(q, rbq.SearchParams(k, nprobe))
    exact = assert_results_match(got, exp[:3], TOL, f"case {case}")
    assert exact == q.shape[0], f"only {exact}/{q.shape[0]} queries bit-identical"
    st = gix.stats()
    want_blocks = int(exp[3][:, 3].sum())                                        # same lists, same blocks (auto schedule;
    assert st["blocks_scanned"] + st["tail_blocks"] >= want_blocks              #  an overflowed tail is walked twice)
    assert st["blocks_scanned"] + st["tail_blocks"] == want_blocks or st["overflow_queries"] > 0
    assert st["admitted"] == int(exp[3][:, 0].sum()) or bits > 1     # estimated (+ non-finite) in reference order
    assert st["refined"] >= int(exp[3][:, 2].sum()) or bits == 1     # refined set is a superset


def test_search_edge_cases(rbq, oracle):
    data, oix, blob = oracle_index(2000, 64, 16, 7, 0, kind="uniform11")
    gix = _load(rbq, blob)
    q = data[:5]
    P = rbq.SearchParams
    ids, sc, cnt = gix.batch_search(q, P(0, 4))
    assert np.all(cnt == 0)                                           # top_k == 0 (ivf.rs:1792)
    assert_results_match(gix.batch_search(q, P(5, 0)), oix.search_batch(q, 5, 1))            # nprobe clamps up
    assert_results_match(gix.batch_search(q, P(5, 10 ** 6)), oix.search_batch(q, 5, 16))     # nprobe clamps down
    assert_results_match(gix.batch_search(q, P(300, 2)), oix.search_batch(q, 300, 2))        # fewer than k results
    with pytest.raises(rbq.DimensionMismatch) as e:
        gix.batch_search(np.zeros((2, 65), np.float32), P(5, 4))
    assert "expected 64, got 65" in str(e.value)
    with pytest.raises(rbq.InvalidConfig):
        gix.batch_search(q, P(5000, 4))
    assert gix.search(q[0], P(3, 4))[0][0] == int(oix.search_batch(q[:1], 3, 4)[0][0, 0])
    r = gix.batch_query(q, 4, 8)                                      # pyo3 surface
    assert len(r) == 5 and r[0].shape == (4, 2) and r[0].dtype == np.float32
    one = gix.query(q[1], 4, 8)
    assert np.array_equal(one, r[1])
    nanq = q.copy()
    nanq[0, 3] = np.nan                                               # NaN query -> no finite distance -> empty
    ids, sc, cnt = gix.batch_search(nanq, P(5, 4))
    assert cnt[0] == 0 and cnt[1] == oix.search_batch(q[1:2], 5, 4)[2][0]


def test_empty_index_and_empty_lists(rbq, oracle):
    import struct
    import zlib

    # header-only stream: zero clusters, zero vectors -> loads, search reports EmptyIndex
    body = struct.pack("<IIBBBBQQQ", 16, 64, 0, 1, 6, 7, 0, 0, 32) + bytes(32)
    blob = b"RBQ1" + struct.pack("<I", 3) + body + struct.pack("<I", zlib.crc32(body))
    gix = _load(rbq, blob)
    assert len(gix) == 0
    with pytest.raises(rbq.EmptyIndex):
        gix.batch_search(np.zeros((1, 16), np.float32), rbq.SearchParams(5, 4))
    # lists with zero vectors and ragged tails (n not a multiple of 32)
    from oracle import oracle as orc

    rng = np.random.default_rng(3)
    data = rng.standard_normal((333, 48)).astype(np.float32)
    cents = rng.standard_normal((12, 48)).astype(np.float32)
    assign = rng.integers(0, 9, 333).astype(np.uint32)  # lists 9..11 stay empty
    oix = orc.Index.train_with_clusters(data, cents, assign, 7, 0)
    gix = _load(rbq, oix.save_bytes())
    q = data[:40]
    assert_results_match(gix.batch_search(q, rbq.SearchParams(7, 12)), oix.search_batch(q, 7, 12))


def test_filtered_search(rbq, oracle):
    data, oix, blob = oracle_index(3000, 64, 16, 7, 0, kind="uniform11")
    gix = _load(rbq, blob)
    q = _queries(data, 64, 9)
    allow = np.arange(0, 3000, 3)
    bits = rbq.ids_to_bitset(allow, 3000)
    got = gix.batch_search(q, rbq.SearchParams(10, 16), filter_bits=bits)
    exp = oix.search_batch(q, 10, 16, filter_bits=bits)
    assert assert_results_match(got, exp) == 64
    assert np.all(got[0][got[0] != np.iinfo(np.uint64).max] % 3 == 0)
    got = gix.batch_search(q, rbq.SearchParams(10, 16), filter_bits=np.zeros_like(bits))
    assert np.all(got[2] == 0)                                        # empty filter -> empty results
    res = gix.search_filtered(q[0], rbq.SearchParams(5, 16), allow)
    assert [r[0] for r in res] == exp[0][0, :5].tolist()


def test_persistence_roundtrip_through_device(rbq, oracle):
    for bits, metric in ((7, 1), (3, 0), (1, 0)):
        data, oix, blob = oracle_index(500, 24, 32, bits, metric, seed=7412, faster=False, kind="uniform11")
        gix = _load(rbq, blob)
        assert gix.save_to_bytes() == blob                            # byte-identical RBQ1 v3 stream
        again = _load(rbq, gix.save_to_bytes())
        a = gix.batch_search(data[:5], rbq.SearchParams(5, 12))
        b = again.batch_search(data[:5], rbq.SearchParams(5, 12))
        assert all(np.array_equal(x, y) for x, y in zip(a, b))       # tests.rs:393-431


def test_sharded_lists_and_device_merge(rbq, oracle):
    import ctypes as C

    import torch

    from rabitq_rs_b200 import _ffi

    data, oix, blob = oracle_index(6000, 64, 64, 7, 0, kind="clustered")
    full = _load(rbq, blob)
    nsh, k, nprobe = 4, 10, 16
    shards = [_load(rbq, blob, shard_rank=r, shard_count=nsh) for r in range(nsh)]
    assert sum(s.local_len() for s in shards) == 6000 and all(len(s) == 6000 for s in shards)
    sizes = [s.local_len() for s in shards]
    assert max(sizes) - min(sizes) < 0.1 * 6000 / nsh                 # size-balanced
    q = _queries(data, 200, 11)
    dq = torch.from_numpy(q).cuda()
    ids = torch.empty((nsh, 200, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((nsh, 200, k), dtype=torch.float32, device="cuda")
    cn = torch.empty((nsh, 200), dtype=torch.int32, device="cuda")
    for r, s in enumerate(shards):
        s.batch_search_device(dq, k, nprobe, ids[r], sc[r], cn[r])
    oi = torch.empty((200, k), dtype=torch.int64, device="cuda")
    os_ = torch.empty((200, k), dtype=torch.float32, device="cuda")
    oc = torch.empty(200, dtype=torch.int32, device="cuda")
    rc = _ffi.lib().rbq_merge_topk_device(full.handle, nsh, 200, k, C.c_void_p(ids.data_ptr()), C.c_void_p(sc.data_ptr()),
                                          C.c_void_p(cn.data_ptr()), C.c_void_p(oi.data_ptr()), C.c_void_p(os_.data_ptr()),
                                          C.c_void_p(oc.data_ptr()), None)
    assert rc == 0
    torch.cuda.synchronize()
    exp = full.batch_search(q, rbq.SearchParams(k, nprobe))
    gid = oi.cpu().numpy().astype(np.uint64)
    # per-shard thresholds can admit bound-violating candidates the single sequence skipped (class D2):
    # results must agree on (almost) every id and never be worse than the single-GPU answer
    agree = np.mean([len(set(gid[i]) & set(exp[0][i])) / k for i in range(200)])
    assert agree >= 0.995
    assert np.all(os_.cpu().numpy() <= exp[1] + 1e-6)
    assert np.array_equal(oc.cpu().numpy().astype(np.uint32), exp[2])


def test_device_resident_entry_matches_host_entry(rbq, oracle):
    import torch

    data, oix, blob = oracle_index(10000, 128, 256, 7, 0, kind="uniform01")
    gix = _load(rbq, blob)
    q = _queries(data, 300, 12)
    host = gix.batch_search(q, rbq.SearchParams(10, 32))
    dq = torch.from_numpy(q).cuda()
    ids = torch.empty((300, 10), dtype=torch.int64, device="cuda")
    sc = torch.empty((300, 10), dtype=torch.float32, device="cuda")
    cn = torch.empty(300, dtype=torch.int32, device="cuda")
    gix.batch_search_device(dq, 10, 32, ids, sc, cn)
    torch.cuda.synchronize()
    assert np.array_equal(ids.cpu().numpy().astype(np.uint64), host[0])
    assert np.array_equal(sc.cpu().numpy(), host[1])


@pytest.mark.gpu
@pytest.mark.parametrize("slots,chunks,first", [(1, 4, 0), (2, 4, 0), (3, 5, 0), (4, 8, 0), (2, 3, 384), (4, 16, 128), (4, 3, -1)])
@pytest.mark.parametrize("geom", [(20000, 128, 256, 7, 0), (12000, 960, 64, 7, 0), (12000, 96, 128, 3, 1)])
def test_host_feed_chunks_on_slot_streams(rbq, oracle, geom, slots, chunks, first, monkeypatch):
    """Host entry: the feed chunks' front end + head pass alternate over `slots` streams (own front-end scratch, own part of
    the dense head buffer).  Whatever the chunking, the answer is the device entry's (one chunk, one stream) and the oracle's,
    in both scan schedules."""
    import torch

    n, dim, nlist, bits, metric = geom
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, kind="clustered")
    gix = _load(rbq, blob)
    nq = 2600
    q = _queries(data, nq, 33)
    dq = torch.from_numpy(q).cuda()
    ids = torch.empty((nq, 10), dtype=torch.int64, device="cuda")
    sc = torch.empty((nq, 10), dtype=torch.float32, device="cuda")
    cn = torch.empty(nq, dtype=torch.int32, device="cuda")
    exp = oix.search_batch(q[:400], 10, 12)
    monkeypatch.setenv("RBQ_FEED_SLOTS", str(slots))
    monkeypatch.setenv("RBQ_FEED_CHUNKS", str(chunks))
    if first > 0:
        monkeypatch.setenv("RBQ_FEED_FIRST", str(first))
    if first < 0:  # optional tapered end of the tile (the last chunks halve)
        monkeypatch.setenv("RBQ_FEED_TAPER", "1")
    for mode in (2, 1, 0):
        gix.set_scan_mode(mode)
        gix.batch_search_device(dq, 10, 12, ids, sc, cn)
        torch.cuda.synchronize()
        for rep in range(2):  # twice: the second call reuses the slot streams and the workspace
            host = gix.batch_search(q, rbq.SearchParams(10, 12))
            cnt = cn.cpu().numpy().astype(np.uint32)
            assert np.array_equal(cnt, np.asarray(host[2]).astype(np.uint32)), (mode, rep)
            live = np.arange(10)[None, :] < cnt[:, None]
            assert np.array_equal(ids.cpu().numpy().astype(np.uint64)[live], host[0][live]), (mode, rep)
            assert np.array_equal(sc.cpu().numpy()[live], host[1][live]), (mode, rep)
        assert assert_results_match(tuple(a[:400] for a in host), exp, TOL, "slots") == 400


COARSE_CASES = [  # (n, dim, nlist, metric, rotator, kind)
    (20000, 128, 512, 0, 1, "clustered"),
    (20000, 960, 300, 0, 1, "clustered"),     # nlist not a multiple of the GEMM tile, K' = 2880
    (20000, 768, 256, 1, 1, "clustered"),     # inner product
    (6000, 96, 64, 0, 1, "uniform11"),        # padded 128
    (3000, 48, 40, 1, 0, "uniform11"),        # MatrixRotator, K' = 144 (not a multiple of 64: TMA zero fill)
]


@pytest.mark.parametrize("case", COARSE_CASES)
def test_tensor_core_coarse_matches_exact_coarse(rbq, oracle, case):
    """Mode 1 (tcgen05 GEMM candidates + exact re-score) must give the reference's probe list bit for bit,
    exactly like mode 0 (exact FP32 scoring of every centroid) and the oracle."""
    n, dim, nlist, metric, rot, kind = case
    data, oix, blob = oracle_index(n, dim, nlist, 1, metric, rotator=rot, kind=kind)
    gix = _load(rbq, blob)
    q = np.concatenate([data[:100], _queries(data, 412, 21)])
    for nprobe in (1, 16, 64):
        nprobe = min(nprobe, nlist)
        gix.set_coarse_mode(0)
        c0, f0 = gix.debug_probe(q, nprobe)
        gix.set_coarse_mode(1)
        c1, f1 = gix.debug_probe(q, nprobe)
        assert np.array_equal(c0, c1), "tensor-core candidate path changed the probe list"
        assert np.array_equal(f0.view(np.uint32), f1.view(np.uint32))
        for i in range(0, q.shape[0], 37):
            assert np.array_equal(c1[i], oix.search_dump(q[i], 1, nprobe)["probe"])
    gix.set_coarse_mode(1)
    got = gix.batch_search(q, rbq.SearchParams(10, min(16, nlist)))
    st = gix.stats()
    assert st["coarse_fallbacks"] <= q.shape[0] // 20, st       # the margin is tight enough to avoid the slow path
    gix.set_coarse_mode(0)
    ref = gix.batch_search(q, rbq.SearchParams(10, min(16, nlist)))
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))


# ---- list-major scan (head pass -> tail kernel over pairs grouped by list -> ordered replay) ----------------
@pytest.mark.parametrize("case", SEARCH_CASES)
def test_list_major_scan_end_to_end(rbq, oracle, case):
    """Scan mode 2 must reproduce the reference's sequential decisions exactly: same ids, same score bits,
    the same number of admitted candidates, and every (query, block) evaluated exactly once."""
    n, dim, nlist, bits, metric, rot, kind, k, nprobe = case
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind=kind)
    gix = _load(rbq, blob)
    gix.set_scan_mode(2)
    q = np.concatenate([data[:32], _queries(data, 480, 4)])
    exp = oix.search_batch(q, k, nprobe, want_diag=True)
    got = gix.batch_search(q, rbq.SearchParams(k, nprobe))
    exact = assert_results_match(got, exp[:3], TOL, f"list-major case {case}")
    assert exact == q.shape[0], f"only {exact}/{q.shape[0]} queries bit-identical"
    st = gix.stats()
    want_blocks = int(exp[3][:, 3].sum())
    assert st["blocks_scanned"] + st["tail_blocks"] >= want_blocks  # an overflowed tail is walked twice
    if kind == "clustered":  # prunable data: the survivor buffer must hold
        assert st["overflow_queries"] == 0 and st["blocks_scanned"] + st["tail_blocks"] == want_blocks
    assert st["tail_blocks"] > 0 or nprobe == 1
    assert st["admitted"] == int(exp[3][:, 0].sum()) or bits > 1
    seq = _load(rbq, blob)
    seq.set_scan_mode(1)
    ref = seq.batch_search(q, rbq.SearchParams(k, nprobe))
    assert all(np.array_equal(x, y) for x, y in zip(_canon_all(got), _canon_all(ref)))


def _canon_all(res):
    from helpers import _canon

    ids, sc, cnt = res
    out = ids.copy()
    for i in range(len(cnt)):
        out[i, :cnt[i]] = _canon(ids[i, :cnt[i]], sc[i, :cnt[i]])
    return out, sc, cnt


@pytest.mark.parametrize("geom", [(6000, 128, 64, 7, 0), (4000, 960, 32, 7, 0), (5000, 96, 48, 3, 1), (5000, 64, 48, 1, 0)])
@pytest.mark.parametrize("cap", [1, 8, 48])
def test_list_major_survivor_overflow_falls_back_exactly(rbq, oracle, geom, cap):
    """Small survivor caps push queries off the lazy replay: with `cap` sorted entries the buffer holds 4 * cap, so a query
    with cap < n <= 4 * cap survivors takes the overflow tier (one CTA: eager refinement, sort, replay) and one with more is
    re-walked by the sequential kernel.  Every tier must reproduce the oracle, and the admitted count (the reference's
    decisions) must not depend on the tier."""
    from rabitq_rs_b200 import _ffi

    n, dim, nlist, bits, metric = geom
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, kind="clustered")
    gix = _load(rbq, blob)
    gix.set_scan_mode(2)
    q = _queries(data, 300, 21)
    exp = oix.search_batch(q, 20, 24)
    got0 = gix.batch_search(q, rbq.SearchParams(20, 24))
    st0 = gix.stats()
    assert _ffi.lib().rbq_debug_set_survivor_cap(cap) == 0
    try:
        got = gix.batch_search(q, rbq.SearchParams(20, 24))
        st = gix.stats()
    finally:
        _ffi.lib().rbq_debug_set_survivor_cap(0)
    assert assert_results_match(got, exp, TOL, "overflow") == 300
    assert assert_results_match(got0, exp, TOL, "default caps") == 300
    assert st["overflow_queries"] > 0
    assert st["admitted"] == st0["admitted"], (st["admitted"], st0["admitted"])


def test_list_major_filtered_edge_and_sharded(rbq, oracle):
    data, oix, blob = oracle_index(3000, 64, 16, 7, 0, kind="uniform11")
    gix = _load(rbq, blob)
    gix.set_scan_mode(2)
    q = _queries(data, 200, 9)
    bits = rbq.ids_to_bitset(np.arange(0, 3000, 3), 3000)
    got = gix.batch_search(q, rbq.SearchParams(10, 16), filter_bits=bits)
    exp = oix.search_batch(q, 10, 16, filter_bits=bits)
    assert assert_results_match(got, exp, TOL, "filtered") == 200
    # heap never fills (top_k > vectors probed): everything stays in the head pass
    assert_results_match(gix.batch_search(q, rbq.SearchParams(1000, 3)), oix.search_batch(q, 1000, 3), TOL, "k>n")
    # single probe, and nprobe == nlist
    assert_results_match(gix.batch_search(q, rbq.SearchParams(5, 1)), oix.search_batch(q, 5, 1), TOL, "nprobe1")
    assert_results_match(gix.batch_search(q, rbq.SearchParams(5, 16)), oix.search_batch(q, 5, 16), TOL, "all lists")
    # a shard owns a subset of the lists: pairs of foreign lists are skipped
    sh = _load(rbq, blob, shard_rank=1, shard_count=3)
    sh.set_scan_mode(2)
    a = sh.batch_search(q, rbq.SearchParams(10, 16))
    sh.set_scan_mode(1)
    b = sh.batch_search(q, rbq.SearchParams(10, 16))
    assert all(np.array_equal(x, y) for x, y in zip(_canon_all(a), _canon_all(b)))
    # ragged / empty lists
    from oracle import oracle as orc

    rng = np.random.default_rng(3)
    d2 = rng.standard_normal((777, 48)).astype(np.float32)
    cents = rng.standard_normal((12, 48)).astype(np.float32)
    assign = rng.integers(0, 9, 777).astype(np.uint32)
    o2 = orc.Index.train_with_clusters(d2, cents, assign, 7, 0)
    g2 = _load(rbq, o2.save_bytes())
    g2.set_scan_mode(2)
    assert assert_results_match(g2.batch_search(d2[:100], rbq.SearchParams(7, 12)), o2.search_batch(d2[:100], 7, 12)) == 100


def test_full_size_properties_sift1m_shape(rbq, oracle):
    """BASELINE config 2 at full size (1M x 128, nlist 4096, total_bits 7, L2, 10k-query batch): too large for
    the oracle to build in a CPU test, so parity is checked through size-independent properties -- the list-major
    (tensor-core) schedule and the sequential per-query schedule are two independent implementations of the
    reference loop and must agree on every query; results are sorted, idempotent, within the index, and the
    host entry equals the device-resident entry; a sample of queries is checked against the oracle on the same
    RBQ1 bytes."""
    import torch
    from oracle import oracle as orc
    from rabitq_rs_b200.kmeans import kmeans_gpu

    n, dim, nlist, nq, k, nprobe = 1_000_000, 128, 4096, 10_000, 10, 32
    g = torch.Generator(device="cuda").manual_seed(7)
    centers = torch.randn(1024, dim, generator=g, device="cuda")
    base = (centers[torch.randint(0, 1024, (n,), generator=g, device="cuda")] + 0.35 * torch.randn(n, dim, generator=g, device="cuda")).cpu().numpy()
    q = (centers[torch.randint(0, 1024, (nq,), generator=g, device="cuda")] + 0.35 * torch.randn(nq, dim, generator=g, device="cuda")).cpu().numpy()
    cents, assign = kmeans_gpu(base, nlist, iters=4, seed=42, device=0)
    ix = rbq.IvfRabitqIndex(dim, 0, device=0)
    ix.fit_with_clusters(base, cents, assign, 7, "fht", seed=42, faster_config=True)
    assert len(ix) == n and ix.cluster_count() == nlist
    p = rbq.SearchParams(k, nprobe)
    ix.set_scan_mode(2)
    a = ix.batch_search(q, p)
    st = ix.stats()
    # a few queries of this tightly clustered data overflow their survivor buffer and are re-walked sequentially
    assert st["tail_pairs"] > 0 and st["overflow_queries"] <= nq // 100, st
    b = ix.batch_search(q, p)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)), "not idempotent"
    ix.set_scan_mode(1)
    c = ix.batch_search(q, p)
    assert all(np.array_equal(x, y) for x, y in zip(_canon_all(a), _canon_all(c))), "list-major != sequential schedule"
    ids, sc, cnt = a
    assert (cnt == k).all() and (ids < n).all()
    assert (np.diff(sc, axis=1) >= 0).all(), "L2 results must be sorted by ascending distance"
    assert all(len(set(r.tolist())) == k for r in ids[:512]), "duplicate ids in a result"
    # device-resident entry == host entry
    dq = torch.from_numpy(q).cuda()
    d_ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    d_sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    d_cn = torch.empty(nq, dtype=torch.int32, device="cuda")
    ix.set_scan_mode(0)
    ix.batch_search_device(dq, k, nprobe, d_ids, d_sc, d_cn)
    torch.cuda.synchronize()
    assert np.array_equal(d_ids.cpu().numpy().astype(np.uint64), ids) and np.array_equal(d_sc.cpu().numpy(), sc)
    # a sample against the oracle on the same bytes
    oix = orc.Index.load_bytes(ix.save_to_bytes())
    exp = oix.search_batch(q[:256], k, nprobe)
    assert assert_results_match((ids[:256], sc[:256], cnt[:256]), exp, TOL, "sift1m sample") == 256


@pytest.mark.parametrize("metric,bits,nprobe", [(0, 7, 16), (1, 3, 16), (0, 1, 16), (0, 7, 1), (1, 3, 2), (0, 7, 64)])
def test_phased_sharded_search_matches_single_shard(rbq, oracle, metric, bits, nprobe):
    """The three-phase multi-GPU search (rbq_dist_front / _head / _tail) with 3 list shards emulated on one device:
    the exchanges (all-gather of the probe slices, MIN all-reduce of tau, all-gather of the local top-k) are done
    with torch ops.  The merged result must equal the single-shard search except where a lower bound is violated
    (class D2: the shards prune with a different threshold sequence), and recall@10 must not drop."""
    import torch
    from rabitq_rs_b200.distributed import query_slices

    data, oix, blob = oracle_index(20000, 128, 64, bits, metric, kind="clustered")
    q = _queries(data, 2000, 17)
    nq, k, world = q.shape[0], 10, 3
    full = _load(rbq, blob)
    want = full.batch_search(q, rbq.SearchParams(k, nprobe))
    shards = [_load(rbq, blob, shard_rank=r, shard_count=world) for r in range(world)]
    assert sum(s.local_len() for s in shards) == len(full)
    dq = torch.from_numpy(q).cuda()
    per, slices = query_slices(nq, world)
    probes = torch.zeros((per * world, nprobe, 4), dtype=torch.int32, device="cuda")
    for r, s in enumerate(shards):  # phase 1 (+ "all-gather": the slices land in one buffer)
        s.dist_front(dq, k, nprobe, slices[r][0], slices[r][1], probes)
    l_ids = torch.empty((world, nq, k), dtype=torch.int64, device="cuda")
    l_sc = torch.empty((world, nq, k), dtype=torch.float32, device="cuda")
    l_cn = torch.empty((world, nq), dtype=torch.int32, device="cuda")
    taus = torch.empty((world, nq), dtype=torch.float32, device="cuda")
    for r, s in enumerate(shards):  # phase 2
        s.dist_head(nq, k, nprobe, probes, taus[r], l_ids[r], l_sc[r], l_cn[r])
    torch.cuda.synchronize()
    owners = torch.isfinite(taus).sum(0)
    assert int(owners.max()) <= 1, "two shards claimed the same query's head pass"
    tau = taus.min(0).values.contiguous()  # MIN all-reduce
    for r, s in enumerate(shards):  # phase 3
        s.dist_tail(nq, k, nprobe, tau, l_ids[r], l_sc[r], l_cn[r])
    m_ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    m_sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    m_cn = torch.empty(nq, dtype=torch.int32, device="cuda")
    shards[0].merge_topk_device(world, nq, k, l_ids, l_sc, l_cn, m_ids, m_sc, m_cn)
    torch.cuda.synchronize()
    got = (m_ids.cpu().numpy().astype(np.uint64), m_sc.cpu().numpy(), m_cn.cpu().numpy().astype(np.uint32))
    assert np.array_equal(got[2], want[2])
    # the one-collective form: every shard's ids | scores | counts in one packed chunk
    chunk = (nq * k * 12 + nq * 4 + 15) // 16 * 16
    packed = torch.zeros(world * chunk, dtype=torch.uint8, device="cuda")
    for r in range(world):
        c = packed[r * chunk:(r + 1) * chunk]
        c[:nq * k * 8].view(torch.int64).copy_(l_ids[r].reshape(-1))
        c[nq * k * 8:nq * k * 12].view(torch.float32).copy_(l_sc[r].reshape(-1))
        c[nq * k * 12:nq * k * 12 + nq * 4].view(torch.int32).copy_(l_cn[r])
    p_ids, p_sc, p_cn = torch.empty_like(m_ids), torch.empty_like(m_sc), torch.empty_like(m_cn)
    shards[0].merge_topk_packed_device(world, nq, k, packed, chunk, p_ids, p_sc, p_cn)
    torch.cuda.synchronize()
    assert torch.equal(p_ids, m_ids) and torch.equal(p_sc, m_sc) and torch.equal(p_cn, m_cn)
    # Every shard's threshold sequence is min(tau, local k-th) >= the single sequence's threshold at the same point, so a shard
    # admits a superset of what the single search admits: the merged result can differ only by bound-violating candidates the
    # single sequence skipped (class D2) and is never worse.
    agree = np.mean([len(set(got[0][i, :k].tolist()) & set(want[0][i, :k].tolist())) / k for i in range(nq)])
    assert agree >= 0.995, agree
    if metric == 0:
        assert np.all(got[1] <= want[1] + 1e-6 * np.abs(want[1]) + 1e-6)
    else:
        assert np.all(got[1] >= want[1] - 1e-6 * np.abs(want[1]) - 1e-6)
    d = (data @ q.T) if metric == 1 else -((data * data).sum(1)[:, None] - 2.0 * (data @ q.T))
    gt = np.argsort(-d, axis=0)[:k].T
    rec = lambda ids: float(np.mean([len(set(ids[i, :k].tolist()) & set(gt[i].tolist())) / k for i in range(nq)]))
    assert rec(got[0]) >= rec(want[0]) - 1e-3, (rec(got[0]), rec(want[0]))  # a superset of admissions can only help recall


def _canon(ids, sc):
    from helpers import _canon as c

    return c(ids, sc)


@pytest.mark.parametrize("geom", [(500, 64, 4, 7, 0, 0), (500, 128, 8, 7, 0, 1), (400, 96, 4, 3, 1, 1), (300, 960, 4, 7, 0, 1),
                                  (300, 40, 4, 1, 0, 1), (300, 768, 4, 5, 1, 1)])
def test_fetch_embedding_bit_exact(rbq, oracle, geom):
    """rbq_fetch_embedding == the oracle's fetch_embedding bit for bit (every rotator / ex-code layout), None for an
    unknown id, and only the owning shard finds an id."""
    n, dim, nlist, bits, metric, rot = geom
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind="uniform11")
    gix = _load(rbq, blob)
    for i in list(range(0, n, 7)) + [n - 1]:
        want = oix.fetch_embedding(i)
        got = gix.fetch_embedding(i)
        assert got is not None and np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"id {i}"
    assert gix.fetch_embedding(n + 10) is None and oix.fetch_embedding(n + 10) is None
    shards = [_load(rbq, blob, shard_rank=r, shard_count=2) for r in range(2)]
    for i in (0, n // 2, n - 1):
        found = [s.fetch_embedding(i) for s in shards]
        assert sum(f is not None for f in found) == 1
        assert np.array_equal(next(f for f in found if f is not None), oix.fetch_embedding(i))


@pytest.mark.parametrize("name", ["l2_b7_fht128", "ip_b3_fht96", "l2_b1_matrix32"])
def test_committed_fixtures(rbq, name):
    """The CUDA path against the committed golden fixtures (frozen RBQ1 bytes + the oracle's answers; see
    tests/golden/make_fixtures.py): identical ids and score bits, both scan schedules, byte-identical re-serialisation."""
    from test_oracle_golden import load_golden

    blob, q, ids, scores, counts, (k, nprobe, metric) = load_golden(name)
    gix = _load(rbq, blob)
    assert gix.save_to_bytes() == blob
    for mode in (1, 2):
        gix.set_scan_mode(mode)
        got = gix.batch_search(q, rbq.SearchParams(k, nprobe))
        assert assert_results_match(got, (ids, scores, counts), TOL, f"golden {name} mode {mode}") == q.shape[0]


# ---- coarse filter mode: probe selection fused into the GEMM epilogue (no nq x nlist score matrix) ----------------
def _random_cluster_index(orc, n, dim, nlist, metric, seed, bits=1, sort_centroids=False, dup=False):
    rng = np.random.default_rng(seed)
    cents = rng.standard_normal((nlist, dim)).astype(np.float32)
    if dup:  # runs of identical centroids: equal scores, the reference orders them by cluster id
        cents[1::2] = cents[0::2][: len(cents[1::2])]
    if sort_centroids:  # adversarial for a strided sample: the table is ordered along one coordinate
        cents = cents[np.argsort(cents[:, 0])]
    data = (cents[rng.integers(0, nlist, n)] + 0.3 * rng.standard_normal((n, dim))).astype(np.float32)
    if metric == 1:
        data /= np.linalg.norm(data, axis=1, keepdims=True)
    d2 = (data * data).sum(1)[:, None] - 2.0 * data @ cents.T + (cents * cents).sum(1)[None, :]
    assign = d2.argmin(1).astype(np.uint32)
    return data, orc.Index.train_with_clusters(data, cents, assign, bits, metric)


FILTER_CASES = [  # (n, dim, nlist, metric, sorted centroid table, duplicated centroids)
    (30000, 64, 2560, 0, False, False),
    (30000, 128, 4096, 1, False, False),
    (20000, 96, 2100, 0, True, False),    # nlist not a multiple of the GEMM tile; sorted table
    (20000, 64, 2048, 0, False, True),    # ties between centroids
]


@pytest.mark.parametrize("case", FILTER_CASES)
def test_coarse_filter_mode_matches_exact_coarse(rbq, oracle, case):
    """Mode 2 (threshold from a centroid sample, candidates appended by the GEMM epilogue, selection on the lists,
    exact fallback) must give the reference's probe list bit for bit, like mode 0 and the oracle."""
    from oracle import oracle as orc

    n, dim, nlist, metric, srt, dup = case
    data, oix = _random_cluster_index(orc, n, dim, nlist, metric, 11, sort_centroids=srt, dup=dup)
    gix = _load(rbq, oix.save_bytes())
    q = np.concatenate([data[:200], _queries(data, 824, 31)])
    q[5] = 0.0
    for nprobe in (1, 8, 40, 200):
        gix.set_coarse_mode(0)
        c0, f0 = gix.debug_probe(q, nprobe)
        for terms in (3, 1):
            gix.set_coarse_mode(2)
            gix.set_coarse_terms(terms)
            c2, f2 = gix.debug_probe(q, nprobe)
            st = gix.stats()
            if nprobe <= 40:  # a larger nprobe may make the filter ineligible (then the dense path runs)
                assert st["coarse_mode_used"] == 2, (nprobe, terms, st)
            assert np.array_equal(c0, c2), f"filter mode changed the probe list (nprobe {nprobe}, terms {terms})"
            assert np.array_equal(f0.view(np.uint32), f2.view(np.uint32))
        gix.set_coarse_terms(3)
        for i in range(0, q.shape[0], 97):
            assert np.array_equal(c0[i], oix.search_dump(q[i], 1, nprobe)["probe"])
    # end to end in filter mode, few fallbacks
    gix.set_coarse_mode(2)
    got = gix.batch_search(q, rbq.SearchParams(10, 16))
    st = gix.stats()
    assert st["coarse_mode_used"] == 2 and st["coarse_fallbacks"] <= q.shape[0] // 20, st
    gix.set_coarse_mode(0)
    ref = gix.batch_search(q, rbq.SearchParams(10, 16))
    assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    # nprobe too large for the filter to pay: auto falls back to the dense tensor-core path
    gix.set_coarse_mode(-1)
    gix.batch_search(q[:64], rbq.SearchParams(10, nlist // 2))
    assert gix.stats()["coarse_mode_used"] == 1


def test_empty_filter_admits_nothing(rbq, oracle):
    """RoaringBitmap::new() as the filter: no results (reference src/tests.rs filtered_search_with_empty_filter)."""
    data, oix, blob = oracle_index(2000, 64, 16, 7, 0, kind="uniform11")
    gix = _load(rbq, blob)
    assert gix.search_filtered(data[3], rbq.SearchParams(10, 8), []) == []
    ids, sc, cnt = gix.batch_search(data[:300], rbq.SearchParams(10, 8), np.zeros(0, np.uint64))
    assert (cnt == 0).all()


def test_one_call_sharded_search_world_of_one(rbq, oracle):
    """rbq_search_batch_sharded[_device] (librbq's own NCCL communicator, the three phases + exchanges in one call) on a
    one-rank communicator: every code path of the multi-GPU call, same answers as the plain search and the oracle."""
    import ctypes as C

    import torch
    from rabitq_rs_b200 import _ffi

    data, oix, blob = oracle_index(6000, 96, 48, 7, 0, kind="clustered")
    gix = _load(rbq, blob)
    L = _ffi.lib()
    uid = (C.c_uint8 * 128)()
    assert L.rbq_comm_unique_id(uid) == 0, _ffi.last_error()
    assert L.rbq_comm_init(gix.handle, uid, 0, 1) == 0, _ffi.last_error()
    assert L.rbq_comm_init(gix.handle, uid, 0, 1) != 0          # second communicator on the same handle is refused
    q = _queries(data, 700, 5)
    nq, k, nprobe = q.shape[0], 10, 12
    dq = torch.from_numpy(q).cuda()
    ids = torch.empty((nq, k), dtype=torch.int64, device="cuda")
    sc = torch.empty((nq, k), dtype=torch.float32, device="cuda")
    cn = torch.empty(nq, dtype=torch.int32, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        rc = L.rbq_search_batch_sharded_device(gix.handle, C.c_void_p(dq.data_ptr()), nq, q.shape[1], k, nprobe, C.c_void_p(ids.data_ptr()),
                                               C.c_void_p(sc.data_ptr()), C.c_void_p(cn.data_ptr()), st)
        assert rc == 0, _ffi.last_error()
    torch.cuda.synchronize()
    got = (ids.cpu().numpy().astype(np.uint64), sc.cpu().numpy(), cn.cpu().numpy().astype(np.uint32))
    assert assert_results_match(got, oix.search_batch(q, k, nprobe)) == nq
    h_ids = np.zeros((nq, k), np.uint64)
    h_sc = np.zeros((nq, k), np.float32)
    h_cn = np.zeros(nq, np.uint32)
    rc = L.rbq_search_batch_sharded(gix.handle, q.ctypes.data_as(C.c_void_p), nq, q.shape[1], k, nprobe, h_ids.ctypes.data_as(C.c_void_p),
                                    h_sc.ctypes.data_as(C.c_void_p), h_cn.ctypes.data_as(C.c_void_p))
    assert rc == 0, _ffi.last_error()
    assert np.array_equal(h_ids, got[0]) and np.array_equal(h_sc, got[1]) and np.array_equal(h_cn, got[2])
    # exact merge on the one-rank communicator: eager refinement, record exchange (send/recv to self), global replay, override
    plain = gix.batch_search(q, rbq.SearchParams(k, nprobe))
    gix.set_exact_merge(True)
    for kk, npb in ((k, nprobe), (40, 7), (3, 48), (k, 1), (k, 2)):
        ids2 = torch.empty((nq, kk), dtype=torch.int64, device="cuda")
        sc2 = torch.empty((nq, kk), dtype=torch.float32, device="cuda")
        cn2 = torch.empty(nq, dtype=torch.int32, device="cuda")
        rc = L.rbq_search_batch_sharded_device(gix.handle, C.c_void_p(dq.data_ptr()), nq, q.shape[1], kk, npb, C.c_void_p(ids2.data_ptr()),
                                               C.c_void_p(sc2.data_ptr()), C.c_void_p(cn2.data_ptr()), st)
        assert rc == 0, _ffi.last_error()
        torch.cuda.synchronize()
        got2 = (ids2.cpu().numpy().astype(np.uint64), sc2.cpu().numpy(), cn2.cpu().numpy().astype(np.uint32))
        assert assert_results_match(got2, oix.search_batch(q, kk, npb)) == nq
        stx = gix.stats()
        assert stx["exchanged_records"] > 0 and stx["inexact_queries"] < nq // 4, stx
    gix.set_exact_merge(False)
    assert np.array_equal(plain[0], got[0])
    assert L.rbq_comm_destroy(gix.handle) == 0
    rc = L.rbq_search_batch_sharded_device(gix.handle, C.c_void_p(dq.data_ptr()), nq, q.shape[1], k, nprobe, C.c_void_p(ids.data_ptr()),
                                           C.c_void_p(sc.data_ptr()), C.c_void_p(cn.data_ptr()), st)
    assert rc == 2 and "rbq_comm_init" in _ffi.last_error()


# ---- stage probes through the PRODUCT kernels of the list-major schedule ---------------------------------------------
def _oracle_list_rows(oracle, oix, d, rank, metric, has_ex):
    """(lower bound after the non-finite fallback, ip | estimate) of every vector of the rank-th probed list (oracle)."""
    D = oix.padded_dim
    c = int(d["probe"][rank])
    nv = oix.list_len(c)
    blocks = oix.list_blocks(c).reshape(-1, 4 * D + 384)
    g_add, g_err, dot_qc = (d["probe_f"][rank, i] for i in range(3))
    lows, xs = [], []
    for b in range(blocks.shape[0]):
        ea = oracle.accumulate_block(blocks[b, :4 * D], d["lut"], D)
        fac = blocks[b, 4 * D:].view(np.float32)
        ip, est, lb = oracle.batch_distances(ea, d["delta"], d["sum_vl"], fac[:32], fac[32:64], fac[64:], g_add, g_err, d["k1x"])
        lb = lb.copy()
        bad = ~np.isfinite(lb)
        lb[bad] = np.float32(0.0) if metric == 0 else np.float32(-(np.float32(dot_qc) + np.float32(d["qnorm"])))
        lows.append(lb)
        xs.append(ip if has_ex else est)
    if not lows:
        return np.zeros(0, np.float32), np.zeros(0, np.float32)
    return np.concatenate(lows)[:nv], np.concatenate(xs)[:nv]


@pytest.mark.parametrize("geom", GEOMS)
def test_product_head_and_tail_kernels_bit_exact(rbq, oracle, geom):
    """The kernels a list-major search launches -- head_scan_kernel (PRMT lookups) and the tail FastScan kernel (tcgen05 one-hot
    GEMM) -- dumped directly: their lower bounds and ip / estimate values (hence the integer LUT sums behind them) equal the
    oracle's for every vector of every probed list."""
    n, dim, nlist, bits, metric, rot = geom
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind="uniform11")
    gix = _load(rbq, blob)
    q = _queries(data, 6, 13)
    nprobe = min(4, nlist)
    cap = 1024
    ha, hb, _, _, hn = gix.debug_stage(0, q, nprobe, cap)
    ta, tb, tr, tp, tn = gix.debug_stage(1, q, nprobe, cap)
    for i in range(q.shape[0]):
        d = oix.search_dump(q[i], 5, nprobe)
        rows = [_oracle_list_rows(oracle, oix, d, r, metric, bits > 1) for r in range(nprobe)]
        first = next(r for r in range(nprobe) if len(rows[r][0]))
        lo, x = rows[first]
        assert hn[i] == len(lo)
        assert np.array_equal(ha[i, :hn[i]].view(np.uint32), lo.view(np.uint32)), "head scan lower bounds differ"
        assert np.array_equal(hb[i, :hn[i]].view(np.uint32), x.view(np.uint32)), "head scan ip/estimate differs"
        assert tn[i] == sum(len(r[0]) for r in rows), "tail kernel must report every vector when the threshold is +inf"
        order = np.lexsort((tp[i, :tn[i]], tr[i, :tn[i]]))
        exp_lo = np.concatenate([r[0] for r in rows])
        exp_x = np.concatenate([r[1] for r in rows])
        exp_rank = np.concatenate([np.full(len(r[0]), j, np.uint32) for j, r in enumerate(rows)])
        assert np.array_equal(tr[i, :tn[i]][order], exp_rank)
        assert np.array_equal(ta[i, :tn[i]][order].view(np.uint32), exp_lo.view(np.uint32)), "tail kernel lower bounds differ"
        assert np.array_equal(tb[i, :tn[i]][order].view(np.uint32), exp_x.view(np.uint32)), "tail kernel ip/estimate differs"


@pytest.mark.parametrize("geom", [(600, 128, 8, 7, 0, 1), (600, 960, 8, 3, 0, 1), (600, 768, 8, 7, 1, 1), (600, 768, 8, 5, 1, 1), (400, 32, 8, 3, 0, 0),
                                  (600, 1536, 4, 7, 0, 1),
                                  # one geometry per shape of the staging copy (code rows of 32, 48, 64, 80 bytes: paired form; 128, 256: eight-lane)
                                  (500, 256, 4, 7, 0, 1), (500, 384, 4, 3, 0, 1), (500, 512, 4, 7, 1, 1), (500, 640, 4, 7, 0, 1), (400, 1024, 4, 7, 0, 1),
                                  (300, 2048, 4, 3, 0, 1)])
def test_product_ex_dot_bit_exact(rbq, oracle, geom):
    """K10 through the product's refine path (lane-major rows, 8 FMA chains, AVX2-order horizontal sum) == the oracle's
    ip_packed_ex (AVX2 lane order), bit for bit."""
    n, dim, nlist, bits, metric, rot = geom
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind="uniform11")
    gix = _load(rbq, blob)
    ex_bits, D = bits - 1, oix.padded_dim
    q = _queries(data, 1, 17)[0]
    rq = oix.rotate(q)
    from oracle.oracle import ex_bytes

    for c in (0, nlist - 1):
        nv = oix.list_len(c)
        got = gix.debug_ex_dot(q, c, nv)
        ex = oix.list_ex(c).reshape(nv, ex_bytes(D, ex_bits))
        exp = np.array([oracle.ip_ex(rq, ex[v], ex_bits, fast=True) for v in range(nv)], np.float32)
        assert np.array_equal(got.view(np.uint32), exp.view(np.uint32)), "ex-code dot differs"


@pytest.mark.parametrize("bits", [7, 3])
def test_full_size_gist1m_shape(rbq, oracle, bits):
    """BASELINE config 3 at full size (1M x 960, nlist 4096, total_bits 7 and 3, 10k-query batch): the list-major schedule and
    the sequential schedule agree on every query, the device entry equals the host entry, and 1 024 queries are checked
    against the oracle on the same index bytes."""
    import torch
    from oracle import oracle as orc
    from rabitq_rs_b200.kmeans import kmeans_gpu

    n, dim, nlist, nq, k, nprobe = 1_000_000, 960, 4096, 10_000, 10, 16
    g = torch.Generator(device="cuda").manual_seed(11)
    centers = torch.randn(256, 32, generator=g, device="cuda")
    proj = torch.randn(32, dim, generator=g, device="cuda") / 32 ** 0.5

    def draw(m):
        out = torch.empty((m, dim), dtype=torch.float32)
        for s in range(0, m, 1 << 18):
            mm = min(1 << 18, m - s)
            z = centers[torch.randint(0, 256, (mm,), generator=g, device="cuda")] + 0.6 * torch.randn(mm, 32, generator=g, device="cuda")
            out[s:s + mm] = (z @ proj + 0.05 * torch.randn(mm, dim, generator=g, device="cuda")).cpu()
        return out.numpy()

    base, q = draw(n), draw(nq)
    cents, assign = kmeans_gpu(base, nlist, iters=4, seed=42, device=0)
    ix = rbq.IvfRabitqIndex(dim, 0, device=0)
    ix.fit_with_clusters(base, cents, assign, bits, "fht", seed=42, faster_config=True)
    del base
    p = rbq.SearchParams(k, nprobe)
    ix.set_scan_mode(2)
    a = ix.batch_search(q, p)
    st = ix.stats()
    assert st["tail_pairs"] > 0 and st["overflow_queries"] <= nq // 100, st
    ix.set_scan_mode(1)
    c = ix.batch_search(q, p)
    assert all(np.array_equal(x, y) for x, y in zip(_canon_all(a), _canon_all(c))), "list-major != sequential schedule"
    ids, sc, cnt = a
    assert (cnt == k).all() and (ids < n).all() and (np.diff(sc, axis=1) >= 0).all()
    ix.set_scan_mode(0)
    oix = orc.Index.load_bytes(ix.save_to_bytes())
    exp = oix.search_batch(q[:1024], k, nprobe)
    assert assert_results_match((ids[:1024], sc[:1024], cnt[:1024]), exp, TOL, f"gist1m bits {bits}") == 1024
