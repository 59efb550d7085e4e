"""CPU-side checks of bench.py's helpers (no GPU): the CPU-baseline leg, the recall metric and the workload table."""
import numpy as np

from helpers import oracle_index

import bench


def test_oracle_baseline_reuses_a_loaded_index():
    """The reference arm parses the RBQ1 bytes once and times every step on the same oracle index."""
    data, _, blob = oracle_index(3000, 64, 32, 7, 0, kind="uniform01")
    q = data[:128]
    qps1, cores1, n1, res1, oix = bench.oracle_baseline(blob, q, 10, 8, seconds=0.1)
    qps2, cores2, n2, res2, oix2 = bench.oracle_baseline(None, q, 10, 8, seconds=0.1, oix=oix)
    assert oix2 is oix and cores1 == cores2 >= 1 and qps1 > 0 and qps2 > 0
    m = min(n1, n2)
    assert m >= 1 and np.array_equal(res1[0][:m], res2[0][:m]) and np.array_equal(res1[1][:m], res2[1][:m])


def test_recall_at_k():
    gt = np.array([[1, 2, 3, 4], [5, 6, 7, 8]])
    ids = np.array([[4, 3, 9, 1, 77], [0, 0, 0, 0, 0]], np.uint64)
    assert bench.recall_at_k(ids, gt) == (3 / 4 + 0) / 2


def test_workloads_match_baseline_configs():
    """BASELINE.json configs 2-5 as bench workloads (shape, nlist, bits, metric, batch)."""
    w = bench.WORKLOADS
    assert (w["sift1m"]["n"], w["sift1m"]["dim"], w["sift1m"]["nlist"], w["sift1m"]["total_bits"], w["sift1m"]["nq"]) == (1_000_000, 128, 4096, 7, 10_000)
    assert (w["gist1m"]["n"], w["gist1m"]["dim"], w["gist1m"]["nlist"], w["gist1m"]["total_bits"]) == (1_000_000, 960, 4096, 7)
    assert w["gist1m_b3"]["total_bits"] == 3
    assert (w["emb10m"]["n"], w["emb10m"]["dim"], w["emb10m"]["metric"], w["emb10m"]["total_bits"]) == (10_000_000, 768, 1, 5)
    assert (w["deep100m"]["n"], w["deep100m"]["dim"], w["deep100m"]["nlist"], w["deep100m"]["nq"]) == (100_000_000, 128, 65536, 100_000)
