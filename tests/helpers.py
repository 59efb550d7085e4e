"""Shared builders for the parity tests: indexes are manufactured by the CPU oracle (test
infrastructure), serialised to RBQ1 v3 bytes, and the same bytes are fed to the CUDA engine."""
import functools

import numpy as np


def clustered_data(n, dim, n_centers, seed, sigma=0.35, normalize=False):
    """Seeded Gaussian mixture: centre[u] + sigma * N(0, I)."""
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((n_centers, dim)).astype(np.float32)
    x = centers[rng.integers(0, n_centers, n)] + sigma * rng.standard_normal((n, dim)).astype(np.float32)
    x = x.astype(np.float32)
    if normalize:
        x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


@functools.lru_cache(maxsize=32)
def oracle_index(n, dim, nlist, total_bits, metric, seed=42, rotator=1, faster=True, kind="uniform01"):
    from oracle import oracle as orc

    rng = np.random.default_rng(seed)
    if kind == "uniform01":  # examples/readme_quickstart.rs: rng.gen::<f32>() in [0,1)
        data = rng.random((n, dim), dtype=np.float32)
    elif kind == "uniform11":  # src/tests.rs:16-18
        data = (rng.random((n, dim), dtype=np.float32) * 2 - 1).astype(np.float32)
    else:
        data = clustered_data(n, dim, max(nlist // 4, 1), seed, normalize=(metric == 1))
    ix = orc.Index.train(data, nlist, total_bits, metric, rotator, seed, faster, iters=6)
    return data, ix, ix.save_bytes()


def _canon(ids, sc):
    """Order inside runs of bit-equal scores is implementation-defined in the reference (HeapEntry
    orders on distance only, src/ivf.rs:927-931): canonicalise such runs by id before comparing."""
    order = np.lexsort((ids, sc.view(np.uint32) if sc.dtype == np.float32 else sc))
    # lexsort on raw bits would mis-order negatives; scores in one result are sorted already, so only
    # permute inside equal-score runs
    out = ids.copy()
    i = 0
    n = len(ids)
    while i < n:
        j = i + 1
        while j < n and sc[j] == sc[i]:
            j += 1
        if j - i > 1:
            out[i:j] = np.sort(ids[i:j])
        i = j
    del order
    return out


def assert_results_match(got, exp, tol=1e-5, context="", verbose=True):
    """ids identical and scores bit-identical (after canonicalising exact-tie runs).  Anything else must
    fall under the documented near-tie rule (class D1): scores agree within `tol`
    relative and id sets differ only among candidates tied within `tol` of the k-th score.
    Returns the number of queries that matched bit-for-bit."""
    gid, gsc, gcn = got
    eid, esc, ecn = exp
    assert np.array_equal(gcn, ecn), f"{context}: result counts differ"
    exact = 0
    for q in range(len(ecn)):
        n = int(ecn[q])
        same_scores = np.array_equal(gsc[q, :n].view(np.uint32), esc[q, :n].view(np.uint32))
        if same_scores and np.array_equal(_canon(gid[q, :n], gsc[q, :n]), _canon(eid[q, :n], esc[q, :n])):
            exact += 1
            continue
        if verbose:
            bad = [i for i in range(n) if gid[q, i] != eid[q, i] or gsc[q, i] != esc[q, i]]
            print(f"{context}: query {q} differs at ranks {bad[:8]}: got {[(int(gid[q, i]), float(gsc[q, i])) for i in bad[:4]]} "
                  f"exp {[(int(eid[q, i]), float(esc[q, i])) for i in bad[:4]]} kth={float(esc[q, n - 1])!r}")
        rel = np.abs(gsc[q, :n] - esc[q, :n]) / np.maximum(np.abs(esc[q, :n]), 1e-12)
        assert rel.max() <= tol, f"{context}: query {q} scores differ by {rel.max()}"
        diff = set(gid[q, :n].tolist()) ^ set(eid[q, :n].tolist())
        if diff:
            kth = esc[q, n - 1]
            for i in range(n):
                if eid[q, i] in diff:
                    assert abs(esc[q, i] - kth) <= tol * max(abs(kth), 1e-12), f"{context}: query {q} id sets differ away from the boundary"
    return exact
