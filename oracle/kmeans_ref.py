"""numpy restatement of the reference's k-means steps (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

Follows /root/reference/src/kmeans.rs:
  assign          -- assign_points_for_update / compute_chunk_assignments_only (:439-547, :645-700): distance =
                     |x|^2 + |c|^2 - 2 x.c, clamped at 0, strict '<' so the first (lowest) cluster wins ties
  update          -- update_centroids (:564-602): centroid = sum * (1 / count); an empty cluster is re-seeded from the
                     candidate pool of points farthest from their centroids, taken in descending distance order
The reference's random streams (StdRng = ChaCha12) and its rayon summation order are not reproducible; parity tests
therefore compare one Lloyd step from a given set of centroids, with a near-tie allowance on the assignment.
"""
import numpy as np


def distances(x, cents):
    x = np.asarray(x, np.float64)
    c = np.asarray(cents, np.float64)
    d = (x * x).sum(1)[:, None] + (c * c).sum(1)[None, :] - 2.0 * x @ c.T
    return np.maximum(d, 0.0)


def assign(x, cents):
    d = distances(x, cents)
    a = d.argmin(1)  # first minimum: ties go to the lower cluster id (src/kmeans.rs:505-516)
    return a.astype(np.uint32), d[np.arange(len(a)), a]


def update(x, assignment, best_dist, k, cents_prev):
    x = np.asarray(x, np.float64)
    out = np.array(cents_prev, np.float64, copy=True)
    far = np.argsort(-best_dist, kind="stable")  # candidate pool, farthest first
    nxt = 0
    for c in range(k):
        m = assignment == c
        n = int(m.sum())
        if n:
            out[c] = x[m].sum(0) * (1.0 / n)
        else:
            out[c] = x[far[nxt]]
            nxt += 1
    return out.astype(np.float32)
