"""rabitq_rs_b200 -- B200-native batched IVF+RaBitQ search behind the rabitq-rs search API.

The product is librbq.so (hand-written sm_100a CUDA behind a C ABI, include/rbq.h); this package is
the thin host-side mirror of the reference's interface.  Importing it loads the CUDA library and
raises if it is missing -- there is no CPU fallback.
"""
from . import _ffi
from .index import (BruteForceRabitqIndex, BruteForceSearchParams, CudaError, DimensionMismatch, EmptyIndex, InvalidConfig, InvalidPersistence, IoError,
                    IndexBuilder, IvfRabitqIndex, Metric, RabitqError, RotatorType, SearchParams, ids_to_bitset, shard_assignment)

_ffi.lib()  # fail loudly at import time if the CUDA library is not built

__all__ = ["IvfRabitqIndex", "IndexBuilder", "BruteForceRabitqIndex", "BruteForceSearchParams", "SearchParams", "Metric", "RotatorType", "RabitqError", "DimensionMismatch",
           "InvalidConfig", "EmptyIndex", "IoError", "InvalidPersistence", "CudaError", "ids_to_bitset", "shard_assignment"]
