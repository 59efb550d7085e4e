"""Coarse clustering for `fit` (plumbing around the build path, torch on the GPU).

The reference's k-means (src/kmeans.rs: Faiss-style Lloyd, <=256 sampled points per centroid, 30
iterations, sgemm assignment) is outside the drop-in scope: any clustering produces a valid index
and search parity never depends on it, because both engines read the same index file."""
import numpy as np


def kmeans_gpu(data, k, iters=10, seed=42, device=0, max_points_per_centroid=256, chunk=1 << 16):
    import torch

    dev = torch.device("cuda", device)
    n, dim = data.shape
    g = torch.Generator(device="cpu").manual_seed(int(seed) & 0x7FFFFFFF)
    x_all = torch.from_numpy(np.ascontiguousarray(data, np.float32))
    n_train = min(n, k * max_points_per_centroid)
    sel = torch.randperm(n, generator=g)[:n_train] if n_train < n else torch.arange(n)
    xt = x_all[sel].to(dev)
    cents = xt[torch.randperm(n_train, generator=g)[:k].to(dev)].clone()

    def assign(x, c):
        out = torch.empty(x.shape[0], dtype=torch.int64, device=dev)
        cn = (c * c).sum(1)
        for s in range(0, x.shape[0], chunk):
            xc = x[s:s + chunk]
            out[s:s + chunk] = (cn[None, :] - 2.0 * (xc @ c.T)).argmin(1)
        return out

    for _ in range(iters):
        a = assign(xt, cents)
        sums = torch.zeros_like(cents).index_add_(0, a, xt)
        cnt = torch.bincount(a, minlength=k).to(torch.float32)
        new = sums / cnt.clamp(min=1.0)[:, None]
        empty = cnt == 0
        if empty.any():  # re-seed empty clusters on random training points
            idx = torch.randint(0, n_train, (int(empty.sum()),), generator=g).to(dev)
            new[empty] = xt[idx]
        cents = new
    final = torch.empty(n, dtype=torch.int64)
    for s in range(0, n, 1 << 20):
        final[s:s + (1 << 20)] = assign(x_all[s:s + (1 << 20)].to(dev), cents).cpu()
    return cents.cpu().numpy().astype(np.float32), final.numpy().astype(np.uint32)
