"""k-means for `fit` on the GPU: thin wrappers over librbq's rbq_kmeans_device / rbq_kmeans_assign_device
(csrc/kmeans.cu -- the reference's src/kmeans.rs pipeline with the assignment on the engine's tcgen05 GEMM, arg-min fused
into the epilogue, and a deterministic centroid update).  torch only holds the device buffers."""
import ctypes as C

import numpy as np

from . import _ffi
from .index import _check


def _stream(dev):
    import torch

    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def kmeans_device(x, k, iters=10, seed=42, max_points_per_centroid=256):
    """x: [n, dim] float32 CUDA tensor -> centroids [k, dim] CUDA tensor (Lloyd on a training subset, like the reference)."""
    import torch

    assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()
    n, dim = x.shape
    cents = torch.empty((k, dim), dtype=torch.float32, device=x.device)
    _check(_ffi.lib().rbq_kmeans_device(C.c_void_p(x.data_ptr()), n, dim, int(k), int(iters), int(seed), int(max_points_per_centroid),
                                        C.c_void_p(cents.data_ptr()), x.device.index or 0, _stream(x.device)))
    return cents


def assign_device(x, cents, out=None):
    """Nearest centroid of every row of x (CUDA tensors) -> int32 CUDA tensor (values are u32 cluster ids)."""
    import torch

    assert x.is_cuda and cents.is_cuda and x.is_contiguous() and cents.is_contiguous()
    n, dim = x.shape
    out = out if out is not None else torch.empty(n, dtype=torch.int32, device=x.device)
    _check(_ffi.lib().rbq_kmeans_assign_device(C.c_void_p(x.data_ptr()), n, dim, C.c_void_p(cents.data_ptr()), cents.shape[0],
                                               C.c_void_p(out.data_ptr()), x.device.index or 0, _stream(x.device)))
    return out


def kmeans_gpu(base, k, iters=10, seed=42, device=0, max_points_per_centroid=256):
    """base: [n, dim] float32 numpy array (host).  Returns (centroids [k, dim] float32, assignments [n] uint32), both numpy.
    The data is uploaded in slices; only the training subset and one slice are resident at a time."""
    import torch

    dev = torch.device("cuda", device)
    base = np.ascontiguousarray(base, np.float32)
    n, dim = base.shape
    nt = min(n, k * max_points_per_centroid)
    with torch.cuda.device(dev):
        if nt < n:  # the training subset is drawn on the host side of the wrapper so that the whole set never has to be resident
            rng = np.random.default_rng(seed)
            sel = np.sort(rng.choice(n, nt, replace=False))
            xt = torch.from_numpy(base[sel]).to(dev)
        else:
            xt = torch.from_numpy(base).to(dev)
        cents = kmeans_device(xt, k, iters, seed, max_points_per_centroid=max(max_points_per_centroid, (nt + k - 1) // k))
        assign = np.empty(n, np.uint32)
        step = max(1, (1 << 30) // (dim * 4))
        for s in range(0, n, step):
            xs = xt[s:s + step] if nt == n else torch.from_numpy(base[s:s + step]).to(dev)
            assign[s:s + step] = assign_device(xs, cents).cpu().numpy().view(np.uint32)
        out = cents.cpu().numpy()
    return out, assign
