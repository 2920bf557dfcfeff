"""ctypes binding of include/b200_flat.h (the C-ABI library c99_vectordb_b200/_b200flat.so).

There is no fallback: if the library is missing it is built with nvcc; if that fails, or if a
compute entry point reports an error (no CUDA device, wrong architecture …), a RuntimeError is
raised.  Nothing here imports the oracle.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_b200flat.so"

METRIC_IP, METRIC_L2 = 0, 1
STORE_F32, STORE_BF16 = 0, 1
SCAN_AUTO, SCAN_BULK, SCAN_LDG = 0, 1, 2

class MemoInfo(C.Structure):
    """b200_memo_info (include/b200_flat.h)."""
    _fields_ = [("kind", C.c_int32), ("d", C.c_int32), ("metric", C.c_int32), ("from_hnsw", C.c_int32),
                ("ntotal", C.c_int64), ("rows_offset", C.c_int64), ("ids_offset", C.c_int64)]


# every symbol include/b200_flat.h declares: (name, restype, argtypes)
_fp = C.POINTER(C.c_float)
_ip = C.POINTER(C.c_int64)
_h = C.c_void_p
SIGNATURES = [
    ("b200_abi_version", C.c_int, []),
    ("b200_last_error", C.c_char_p, []),
    ("b200_device_count", C.c_int, [C.POINTER(C.c_int)]),
    ("b200_index_create", C.c_int, [C.POINTER(_h), C.c_int, C.c_int, C.c_int, C.c_int]),
    ("b200_index_destroy", C.c_int, [_h]),
    ("b200_index_reset", C.c_int, [_h]),
    ("b200_index_reserve", C.c_int, [_h, C.c_int64]),
    ("b200_index_set_option", C.c_int, [_h, C.c_char_p, C.c_int64]),
    ("b200_index_get_option", C.c_int, [_h, C.c_char_p, _ip]),
    ("b200_index_add", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    ("b200_index_add_dev", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_void_p, C.c_int]),
    ("b200_index_add_file", C.c_int, [_h, C.c_char_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]),
    ("b200_index_write_file", C.c_int, [_h, C.c_char_p, C.c_int64, C.c_int64]),
    ("b200_memo_probe", C.c_int, [C.c_char_p, C.POINTER(MemoInfo)]),
    ("b200_memo_write_headers", C.c_int, [C.c_char_p, C.POINTER(MemoInfo), _ip, _ip]),
    ("b200_index_save", C.c_int, [_h, C.c_char_p, C.c_int]),
    ("b200_index_load", C.c_int, [C.POINTER(_h), C.c_char_p, C.c_int, C.c_int, C.POINTER(MemoInfo)]),
    ("b200_index_add_synthetic", C.c_int, [_h, C.c_int64, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int64]),
    ("b200_index_search", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    ("b200_index_search_dev", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("b200_index_search_masked", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("b200_index_search_masked_dev", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("b200_index_search_ids_allowed", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    ("b200_ipc_alloc", C.c_int, [C.POINTER(C.c_void_p), C.c_size_t, C.c_char_p]),
    ("b200_ipc_open", C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    ("b200_ipc_close", C.c_int, [C.c_void_p]),
    ("b200_ipc_free", C.c_int, [C.c_void_p]),
    ("b200_exchange_slot_bytes", C.c_size_t, []),
    ("b200_index_set_exchange", C.c_int, [_h, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("b200_index_search_exchange_dev", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("b200_index_exchange_status", C.c_int, [_h]),
    ("b200_index_read_phase_stamps", C.c_int, [_h, C.c_void_p, C.c_int64, _ip]),
    ("b200_index_launch_count", C.c_int64, [_h]),
    ("b200_index_sync", C.c_int, [_h]),
    ("b200_index_ntotal", C.c_int64, [_h]),
    ("b200_index_d", C.c_int, [_h]),
    ("b200_index_metric", C.c_int, [_h]),
    ("b200_index_store", C.c_int, [_h]),
    ("b200_index_has_ids", C.c_int, [_h]),
    ("b200_index_get_ids", C.c_int, [_h, C.c_void_p]),
    ("b200_index_get_rows", C.c_int, [_h, C.c_int64, C.c_int64, C.c_void_p]),
    ("b200_index_rows_dev", C.c_int, [_h, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    ("b200_normalize_rows", C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_int]),
    ("b200_merge_topk_dev", C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("b200_index_search_exchange", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    ("b200_index_search_shard_dev", C.c_int, [_h, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("b200_merge_certify_dev", C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                          C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("b200_hash_embed", C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    ("b200_index_add_texts", C.c_int, [_h, C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, _ip]),
    ("b200_py_hash_seed0", C.c_int64, [C.c_char_p, C.c_int64]),
    ("b200_synth_rows_dev", C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_int64, C.c_int, C.c_void_p]),
]

_lib = None


def load() -> C.CDLL:
    """Load (building first if absent) the native library; raise loudly if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        from . import build as _build

        _build.build_cuda()
    if not LIB_PATH.exists():
        raise RuntimeError(f"native library {LIB_PATH} is missing and could not be built (no CPU fallback exists)")
    L = C.CDLL(str(LIB_PATH))
    for name, restype, argtypes in SIGNATURES:
        fn = getattr(L, name)  # AttributeError here == header/library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = L
    return L


def last_error() -> str:
    return load().b200_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    """Status -> exception.  faiss raises RuntimeError from C++ FaissException [upstream]."""
    if rc != 0:
        raise RuntimeError(last_error() or f"b200 native call failed with status {rc}")


def device_count() -> int:
    n = C.c_int(0)
    check(load().b200_device_count(C.byref(n)))
    return n.value


def device_count_or_zero() -> int:
    """Number of CUDA devices, 0 when the driver is absent (used by tests to tell a GPU box from a build box)."""
    try:
        return device_count()
    except RuntimeError:
        return 0
