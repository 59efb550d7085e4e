import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rust_like_uniform(n, dim, seed):
    """Uniform [-1, 1) vectors, the distribution src/tests.rs:16-18 uses (the RNG itself,
    StdRng/ChaCha12, is not reproducible here; no reference test pins values derived from it)."""
    import numpy as np

    return (np.random.default_rng(seed).random((n, dim), dtype=np.float32) * 2.0 - 1.0).astype(np.float32)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.lib()
    return orc
