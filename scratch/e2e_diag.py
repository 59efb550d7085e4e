import os, sys, time, json, subprocess
import torch
x = torch.empty(10000, 960).pin_memory(); d = torch.empty(10000, 960, device='cuda')
for _ in range(3): d.copy_(x, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"H2D 38.4 MB pinned: {dt*1e3:.3f} ms = {38.4e6/dt/1e9:.1f} GB/s")
y = torch.empty(10000, 10, dtype=torch.int64).pin_memory(); dy = torch.empty(10000, 10, dtype=torch.int64, device='cuda')
t0 = time.perf_counter()
for _ in range(20): y.copy_(dy, non_blocking=True); torch.cuda.synchronize()
print(f"D2H 0.8 MB + sync: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
