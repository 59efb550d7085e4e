#!/usr/bin/env python
"""Host-path enqueue-vs-total trace (RBQ_TRACE) at the bench's gist1m configuration for a few slot/chunk settings. Run under gpurun."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from rabitq_rs_b200 import _ffi  # noqa: E402

wl = dict(bench.WORKLOADS["gist1m"])
log = lambda s: print("[trace]", s, file=sys.stderr, flush=True)
ix, base, queries, gt = bench.build_index(wl, 0, log)
nq, k, dim = wl["nq"], wl["top_k"], wl["dim"]
hq = torch.from_numpy(queries).pin_memory()
h_ids = torch.empty((nq, k), dtype=torch.int64).pin_memory()
h_sc = torch.empty((nq, k), dtype=torch.float32).pin_memory()
h_cn = torch.empty(nq, dtype=torch.int32).pin_memory()
L = _ffi.lib()


def host_call():
    rc = L.rbq_search_batch(ix.handle, C.c_void_p(hq.data_ptr()), nq, dim, k, 16, C.c_void_p(h_ids.data_ptr()), C.c_void_p(h_sc.data_ptr()),
                            C.c_void_p(h_cn.data_ptr()))
    assert rc == 0, _ffi.last_error()


for _ in range(3):
    host_call()
os.environ["RBQ_TRACE"] = "1"
for cfg in sys.argv[1:] or ["RBQ_FEED_SLOTS=4"]:
    kv = [e.split("=") for e in cfg.split(",")]
    for a, b in kv:
        os.environ[a] = b
    log(f"--- {cfg}")
    for _ in range(4):
        host_call()
    for a, _ in kv:
        del os.environ[a]
