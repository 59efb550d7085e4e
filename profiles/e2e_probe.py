#!/usr/bin/env python
"""Host-path (rbq_search_batch, pinned host buffers) timing probe at the bench's gist1m configuration: raw PCIe copy times,
the device-resident step, and the end-to-end step for several feed-chunk counts (RBQ_FEED_CHUNKS).  Run under gpurun."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from rabitq_rs_b200 import _ffi  # noqa: E402

wl = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "gist1m"])
nprobe = int(sys.argv[2]) if len(sys.argv) > 2 else 16
log = lambda s: print("[probe]", s, file=sys.stderr, flush=True)
ix, base, queries, gt = bench.build_index(wl, 0, log)
nq, k, dim = wl["nq"], wl["top_k"], wl["dim"]
dev = torch.device("cuda", 0)
hq = torch.from_numpy(queries).pin_memory()
dq = torch.empty((nq, dim), dtype=torch.float32, device=dev)
h_ids = torch.empty((nq, k), dtype=torch.int64).pin_memory()
h_sc = torch.empty((nq, k), dtype=torch.float32).pin_memory()
h_cn = torch.empty(nq, dtype=torch.int32).pin_memory()
d_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
d_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
d_cn = torch.empty(nq, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
L = _ffi.lib()
out = {}


def wall(fn, n=10):
    ts = []
    for i in range(n + 2):
        flush.fill_(i & 255)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.mean(ts)), float(np.min(ts))


out["h2d_ms"] = wall(lambda: dq.copy_(hq, non_blocking=True))
out["d2h_ms"] = wall(lambda: (h_ids.copy_(d_ids, non_blocking=True), h_sc.copy_(d_sc, non_blocking=True), h_cn.copy_(d_cn, non_blocking=True)))
dq.copy_(hq)
out["device_ms"] = wall(lambda: ix.batch_search_device(dq, k, nprobe, d_ids, d_sc, d_cn))


def host_call():
    rc = L.rbq_search_batch(ix.handle, C.c_void_p(hq.data_ptr()), nq, dim, k, nprobe, C.c_void_p(h_ids.data_ptr()), C.c_void_p(h_sc.data_ptr()),
                            C.c_void_p(h_cn.data_ptr()))
    assert rc == 0, _ffi.last_error()


for ch in (1, 2, 3, 4, 5, 6, 8):
    os.environ["RBQ_FEED_CHUNKS"] = str(ch)
    out[f"e2e_ms_chunks{ch}"] = wall(host_call)
    log(f"chunks {ch}: {out[f'e2e_ms_chunks{ch}']}")
if os.environ.get("RBQ_TRACE"):
    for ch in (1, 4, 8):
        os.environ["RBQ_FEED_CHUNKS"] = str(ch)
        log(f"--- trace, chunks {ch}")
        for _ in range(3):
            host_call()
ix.set_profiling(True)
for ch in (1, 2, 4, 8):
    os.environ["RBQ_FEED_CHUNKS"] = str(ch)
    host_call()
    host_call()
    st = ix.stats()
    out[f"stages_chunks{ch}"] = {kk: round(float(vv), 4) for kk, vv in st.items() if kk.startswith("ms_")}
    log(f"stages, chunks {ch}: {out[f'stages_chunks{ch}']}")
ix.set_profiling(False)
os.environ.pop("RBQ_FEED_CHUNKS", None)
out["e2e_ms_default"] = wall(host_call)
for extra in sys.argv[3:]:  # KEY=VALUE[,KEY=VALUE...] environment knobs, each combination timed (chunk count: the default unless given)
    kv = [e.split("=") for e in extra.split(",")]
    for kk, vv in kv:
        os.environ[kk] = vv
    out[f"e2e_ms_{extra}"] = wall(host_call)
    log(f"{extra}: {out[f'e2e_ms_{extra}']}")
    for kk, _ in kv:
        del os.environ[kk]
print(json.dumps(out))
