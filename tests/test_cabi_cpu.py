"""CPU-side checks of the product's host code: the C-ABI library loads, exports every symbol
include/rbq.h declares, and its RBQ1 parser reports the reference's errors -- no GPU needed
(parsing fails before any CUDA call)."""
import ctypes as C
import os
import re
import struct
import zlib

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def rbq():
    from conftest import build_librbq

    build_librbq()
    import rabitq_rs_b200 as r

    return r


def test_library_exports_every_declared_symbol(rbq):
    hdr = open(os.path.join(ROOT, "include", "rbq.h")).read()
    declared = sorted(set(re.findall(r"\b(rbq_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(os.path.join(ROOT, "rabitq_rs_b200", "librbq.so"))
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"librbq.so does not export: {missing}"


def test_no_cpu_fallback_in_product(rbq):
    """The product package must not import or reference the oracle."""
    pkg = os.path.join(ROOT, "rabitq_rs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower().replace("oracle/", "").replace("the oracle", "") or f in ("build.py",), f
                assert "liboracle" not in src and "import oracle" not in src and "from oracle" not in src, f


def _expect(rbq, blob, exc, msg=None):
    with pytest.raises(exc) as e:
        rbq.IvfRabitqIndex.load_from_bytes(blob)
    if msg:
        assert msg in str(e.value)
    return e.value


def test_parser_error_strings(rbq, oracle):
    from helpers import oracle_index

    _, _, blob = oracle_index(96, 12, 16, 3, 1, seed=0xC0FFEE, faster=False, kind="uniform11")
    _expect(rbq, b"XXXX" + blob[4:], rbq.InvalidPersistence, "unrecognized file header")
    _expect(rbq, blob[:4] + struct.pack("<I", 2) + blob[8:], rbq.InvalidPersistence, "unsupported index format version")
    _expect(rbq, blob[:100], rbq.IoError)
    bad = bytearray(blob)
    bad[len(bad) - 5] ^= 0xAA  # src/tests.rs:433-468
    _expect(rbq, bytes(bad), rbq.InvalidPersistence)
    bad = bytearray(blob)  # src/tests.rs:470-517
    n = struct.unpack_from("<Q", bad, 20)[0]
    struct.pack_into("<Q", bad, 20, n + 1)
    struct.pack_into("<I", bad, len(bad) - 4, zlib.crc32(bytes(bad[8:-4])))
    e = _expect(rbq, bytes(bad), rbq.InvalidPersistence, "vector count metadata mismatch")
    assert e.code == 5
    bad = bytearray(blob)
    bad[16] = 7
    _expect(rbq, bytes(bad), rbq.InvalidPersistence, "unknown metric tag")
    bad = bytearray(blob)
    bad[19] = 9
    _expect(rbq, bytes(bad), rbq.InvalidPersistence, "total_bits does not match ex_bits")
    with pytest.raises(rbq.IoError):
        rbq.IvfRabitqIndex.load_from_path("/nonexistent/index.bin")


def test_crc_matches_zlib(rbq, oracle):
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 1000, 4097):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.crc32(b) == zlib.crc32(b)


def test_python_surface_mirrors_reference(rbq):
    ix = rbq.IvfRabitqIndex(128, "angular")
    assert ix.metric == rbq.Metric.InnerProduct
    with pytest.raises(ValueError):
        rbq.IvfRabitqIndex(8, "manhattan")
    with pytest.raises(RuntimeError):
        len(ix)
    with pytest.raises(RuntimeError):
        ix.query(np.zeros(128, np.float32), 5)
    for name in ("fit", "fit_with_clusters", "query", "batch_query", "save", "load", "cluster_count",
                 "search", "batch_search", "search_filtered"):
        assert callable(getattr(ix, name))
    bits = rbq.ids_to_bitset([0, 3, 64, 130])
    assert bits.tolist() == [9, 1, 4]


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of rbq_search_stats / rbq_probe_rec must have the size and field offsets the C header gives
    them (a silent mismatch would scramble the diagnostics and the multi-GPU probe records)."""
    import ctypes as C
    import subprocess

    from rabitq_rs_b200 import _ffi

    fields = [f for f, _ in _ffi.SearchStats._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rbq.h"\nint main(void) {\n'
                   '  printf("%zu %zu\\n", sizeof(rbq_search_stats), sizeof(rbq_probe_rec));\n'
                   + "".join(f'  printf("%zu\\n", offsetof(rbq_search_stats, {f}));\n' for f in fields)
                   + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == C.sizeof(_ffi.SearchStats) and int(out[1]) == 16
    for f, off in zip(fields, out[2:]):
        assert getattr(_ffi.SearchStats, f).offset == int(off), f
