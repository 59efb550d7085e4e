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


def build_librbq():
    """Build librbq.so in-tree (loaded by path: importing the package needs the library to exist)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("rbq_build", os.path.join(ROOT, "rabitq_rs_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()
