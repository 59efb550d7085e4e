"""ctypes binding of librbq.so (include/rbq.h).  There is no CPU fallback: if the CUDA library is
missing or fails to load, importing the product raises."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librbq.so")

OK, DIMENSION_MISMATCH, INVALID_CONFIG, EMPTY_INDEX, IO, INVALID_PERSISTENCE, CUDA_ERROR = range(7)


class SearchStats(C.Structure):
    _fields_ = [("queries", C.c_uint64), ("blocks_scanned", C.c_uint64), ("bytes_scanned", C.c_uint64),
                ("candidates", C.c_uint64), ("refined", C.c_uint64), ("admitted", C.c_uint64),
                ("kernel_launches", C.c_uint64), ("coarse_fallbacks", C.c_uint64), ("ms_prep", C.c_float), ("ms_coarse", C.c_float),
                ("ms_select", C.c_float), ("ms_scan", C.c_float),
                ("tail_blocks", C.c_uint64), ("tail_bytes", C.c_uint64), ("tail_pairs", C.c_uint64),
                ("survivors", C.c_uint64), ("overflow_queries", C.c_uint64),
                ("ms_scan_head", C.c_float), ("ms_scan_tail", C.c_float), ("ms_scan_replay", C.c_float), ("ms_tail_kernel", C.c_float), ("coarse_mode_used", C.c_uint32), ("front_chunk", C.c_uint32), ("fallback_queries", C.c_uint64),
                ("inexact_queries", C.c_uint64), ("exchanged_records", C.c_uint64),
                ("coarse_terms_used", C.c_uint32), ("reserved_", C.c_uint32)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python rabitq_rs_b200/build.py` "
                          "(rabitq_rs_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    L.rbq_last_error.restype = C.c_char_p
    L.rbq_index_load.argtypes = [C.c_char_p, i32, i32, i32, C.POINTER(vp)]
    L.rbq_index_load_mem.argtypes = [vp, sz, i32, i32, i32, C.POINTER(vp)]
    L.rbq_index_save.argtypes = [vp, C.c_char_p]
    L.rbq_index_save_mem.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.rbq_index_save_lists_mem.argtypes = [vp, vp, sz, vp, sz, C.POINTER(sz)]
    L.rbq_index_free.argtypes = [vp]
    L.rbq_index_free.restype = None
    if hasattr(L, "rbq_index_build"):
        L.rbq_index_build.argtypes = [vp, sz, sz, vp, sz, vp, i32, i32, i32, u64, i32, vp, i32, C.POINTER(vp)]
    L.rbq_builder_create.argtypes = [sz, vp, sz, vp, i32, i32, i32, u64, vp, i32, i32, i32, sz, C.POINTER(vp)]
    L.rbq_builder_add_device.argtypes = [vp, vp, vp, sz, u64, vp]
    L.rbq_builder_finish.argtypes = [vp, C.POINTER(vp)]
    L.rbq_builder_free.argtypes = [vp]
    L.rbq_builder_free.restype = None
    L.rbq_kmeans_device.argtypes = [vp, sz, sz, sz, i32, u64, sz, vp, i32, vp]
    L.rbq_kmeans_assign_device.argtypes = [vp, sz, sz, vp, sz, vp, i32, vp]
    for name in ("len", "local_len", "dim", "padded_dim", "cluster_count"):
        f = getattr(L, "rbq_index_" + name)
        f.argtypes, f.restype = [vp], sz
    for name in ("metric", "ex_bits", "rotator_type", "device"):
        f = getattr(L, "rbq_index_" + name)
        f.argtypes, f.restype = [vp], i32
    L.rbq_search_batch.argtypes = [vp, vp, sz, sz, sz, sz, vp, vp, vp]
    L.rbq_search_batch_filtered.argtypes = [vp, vp, sz, sz, sz, sz, vp, sz, vp, vp, vp]
    L.rbq_search_batch_device.argtypes = [vp, vp, sz, sz, sz, sz, vp, sz, vp, vp, vp, vp]
    L.rbq_merge_topk_device.argtypes = [vp, i32, sz, sz, vp, vp, vp, vp, vp, vp, vp]
    L.rbq_merge_topk_packed_device.argtypes = [vp, i32, sz, sz, vp, sz, vp, vp, vp, vp]
    L.rbq_fetch_embedding.argtypes = [vp, u64, vp, C.POINTER(i32)]
    L.rbq_dist_front.argtypes = [vp, vp, sz, sz, sz, sz, sz, sz, vp, vp]
    L.rbq_dist_head.argtypes = [vp, sz, sz, sz, vp, vp, vp, vp, vp, vp]
    L.rbq_dist_tail.argtypes = [vp, sz, sz, sz, vp, vp, vp, vp, vp]
    L.rbq_bf_train.argtypes = [vp, sz, sz, i32, i32, i32, u64, i32, vp, i32, C.POINTER(vp)]
    L.rbq_bf_load.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
    L.rbq_bf_load_mem.argtypes = [vp, sz, i32, C.POINTER(vp)]
    L.rbq_bf_save.argtypes = [vp, C.c_char_p]
    L.rbq_bf_save_mem.argtypes = [vp, vp, sz, C.POINTER(sz)]
    L.rbq_bf_free.argtypes = [vp]
    L.rbq_bf_free.restype = None
    for name in ("len", "dim", "padded_dim"):
        f = getattr(L, "rbq_bf_" + name)
        f.argtypes, f.restype = [vp], sz
    L.rbq_bf_search_batch.argtypes = [vp, vp, sz, sz, sz, vp, sz, vp, vp, vp]
    L.rbq_comm_unique_id.argtypes = [vp]
    L.rbq_comm_init.argtypes = [vp, vp, i32, i32]
    L.rbq_comm_destroy.argtypes = [vp]
    L.rbq_search_batch_sharded_device.argtypes = [vp, vp, sz, sz, sz, sz, vp, vp, vp, vp]
    L.rbq_search_batch_sharded.argtypes = [vp, vp, sz, sz, sz, sz, vp, vp, vp]
    L.rbq_set_exact_merge.argtypes = [vp, i32]
    L.rbq_shard_assignment.argtypes = [vp, sz, i32, vp, vp, sz, C.POINTER(sz)]
    L.rbq_last_search_stats.argtypes = [vp, C.POINTER(SearchStats)]
    L.rbq_set_profiling.argtypes = [vp, i32]
    L.rbq_set_coarse_mode.argtypes = [vp, i32]
    L.rbq_set_coarse_terms.argtypes = [vp, i32]
    L.rbq_set_scan_mode.argtypes = [vp, i32]
    L.rbq_debug_set_survivor_cap.argtypes = [C.c_uint32]
    L.rbq_debug_query_prep.argtypes = [vp, vp, sz, sz, vp, vp, vp]
    L.rbq_debug_probe.argtypes = [vp, vp, sz, sz, sz, vp, vp]
    L.rbq_debug_scan_list.argtypes = [vp, vp, sz, sz, vp, vp, vp, vp, sz]
    L.rbq_debug_stage.argtypes = [vp, i32, vp, sz, sz, sz, sz, vp, vp, vp, vp, vp]
    L.rbq_debug_ex_dot.argtypes = [vp, vp, sz, sz, sz, vp]
    _lib = L
    return L


def last_error():
    return lib().rbq_last_error().decode("utf-8", "replace")
