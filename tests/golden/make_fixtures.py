#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

The reference (lqhl/rabitq-rs) is a Rust crate that cannot be built in this image (no cargo/rustc), so these are NOT outputs of
the reference itself: the byte-level known answers the reference's own tests pin live in tests/test_oracle_golden.py.  These
fixtures freeze what the CPU oracle (oracle/oracle.cc, the restated reference) returns for three small seeded indexes -- the
RBQ1 bytes it writes and the search results it computes -- so that a later change to the oracle (which DEFINES end-to-end
parity for the CUDA path) cannot drift unnoticed, and so that the CUDA path can be checked against committed vectors.
Run from the repo root:  python tests/golden/make_fixtures.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {  # name: (n, dim, nlist, total_bits, metric, rotator, kind, top_k, nprobe)
    "l2_b7_fht128": (600, 128, 8, 7, 0, 1, "uniform01", 10, 4),
    "ip_b3_fht96": (500, 96, 8, 3, 1, 1, "clustered", 10, 4),
    "l2_b1_matrix32": (400, 32, 8, 1, 0, 0, "uniform11", 5, 8),
}


def main():
    from helpers import oracle_index

    here = os.path.dirname(os.path.abspath(__file__))
    for name, (n, dim, nlist, bits, metric, rot, kind, k, nprobe) in CASES.items():
        data, ix, blob = oracle_index(n, dim, nlist, bits, metric, rotator=rot, kind=kind)
        rng = np.random.default_rng(2024)
        q = (data[rng.integers(0, n, 24)] + 0.05 * rng.standard_normal((24, dim))).astype(np.float32)
        ids, scores, counts = ix.search_batch(q, k, nprobe)
        np.savez_compressed(os.path.join(here, name + ".npz"), blob=np.frombuffer(blob, np.uint8), queries=q, ids=ids, scores=scores,
                            counts=counts, params=np.array([k, nprobe, metric], np.int64),
                            blob_sha256=np.frombuffer(hashlib.sha256(blob).digest(), np.uint8))
        print(name, len(blob), "bytes", hashlib.sha256(blob).hexdigest()[:16])


if __name__ == "__main__":
    main()
