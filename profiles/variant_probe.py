#!/usr/bin/env python
"""Device-resident step at a bench workload under several environment-knob settings (one process, knobs read per launch):
per-stage times (CUDA events inside librbq) and the whole step, plus a check that every setting returns the same ids.
Usage: variant_probe.py <workload> <nprobe> KEY=V[,KEY=V...] ...   ('-' = no knob).  Run under gpurun."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402

wl = dict(bench.WORKLOADS[sys.argv[1]])
nprobe = int(sys.argv[2])
log = lambda s: print("[variant]", s, file=sys.stderr, flush=True)
ix, base, queries, gt = bench.build_index(wl, 0, log)
nq, k = wl["nq"], wl["top_k"]
dev = torch.device("cuda", 0)
dq = torch.from_numpy(queries).to(dev)
d_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
d_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
d_cn = torch.empty(nq, dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out, ref = {}, None
for cfg in sys.argv[3:] or ["-"]:
    kv = [] if cfg == "-" else [e.split("=") for e in cfg.split(",")]
    for a, b in kv:
        os.environ[a] = b
    res = {}
    for prof in (True, False):
        ix.set_profiling(prof)
        acc, ms = {}, []
        for s in range(4 + 20):
            flush.fill_(s & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ix.batch_search_device(dq, k, nprobe, d_ids, d_sc, d_cn)
            e1.record()
            e1.synchronize()
            if s >= 4:
                ms.append(e0.elapsed_time(e1))
                if prof:
                    for kk, vv in ix.stats().items():
                        if kk.startswith("ms_"):
                            acc[kk] = acc.get(kk, 0.0) + vv / 20.0
        if prof:
            res.update({kk: round(vv, 4) for kk, vv in acc.items()})
        res["step_ms_profiled" if prof else "step_ms"] = round(float(np.mean(ms)), 4)
    ids = d_ids.cpu().numpy()
    if ref is None:
        ref = ids.copy()
    res["ids_equal_first"] = bool(np.array_equal(ids, ref))
    out[cfg] = res
    log(f"{cfg}: {res}")
    for a, _ in kv:
        del os.environ[a]
print(json.dumps(out))
