import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import rabitq_rs_b200 as rbq
from helpers import oracle_index
import test_gpu_parity as T
for case in T.SEARCH_CASES:
    n, dim, nlist, bits, metric, rot, kind, k, nprobe = case
    data, oix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind=kind)
    gix = rbq.IvfRabitqIndex.load_from_bytes(blob)
    q = np.concatenate([data[:32], T._queries(data, 480, 4)])
    for mode in (1, 2):
        gix.set_scan_mode(mode)
        gix.batch_search(q, rbq.SearchParams(k, nprobe))
        st = gix.stats()
        nq = len(q)
        print(case, "mode", mode, {kk: round(st[kk] / nq, 1) for kk in ("blocks_scanned", "tail_blocks", "refined", "admitted", "survivors", "overflow_queries", "tail_pairs")})
