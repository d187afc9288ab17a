"""ctypes binding of the C ABI in include/cvr_b200.h (libcvr_b200.so, built in-tree).

There is deliberately no fallback: if the CUDA library is missing, loading fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libcvr_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class CvrCsr(C.Structure):  # cvr_csr_t
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int64), ("nnz", C.c_int64),
                ("val", C.c_void_p), ("col", C.c_void_p),
                ("row_delim32", C.c_void_p), ("row_delim64", C.c_void_p)]


class CvrArrays(C.Structure):  # cvr_arrays_t
    _fields_ = [("vals", C.c_void_p), ("cols", C.c_void_p), ("record", C.c_void_p),
                ("nnz_rows", C.c_void_p), ("final_2", C.c_void_p), ("split", C.c_void_p)]


class CvrInfo(C.Structure):  # cvr_info_t
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int64), ("nnz", C.c_int64),
                ("n_chunks", C.c_int32), ("device", C.c_int32),
                ("n_records", C.c_int64), ("record_ints", C.c_int64),
                ("algorithmic_bytes", C.c_int64),
                ("convert_seconds", C.c_double), ("create_seconds", C.c_double),
                ("kernel_launches", C.c_int64), ("device_bytes", C.c_int64),
                ("convert_kernel_seconds", C.c_double), ("row_lists_seconds", C.c_double)]


class CvrPublish(C.Structure):  # cvr_publish_t
    _fields_ = [("n_dst", C.c_int32), ("self", C.c_int32), ("mode", C.c_int32), ("row_offset", C.c_int64),
                ("needs", C.c_void_p),
                ("chunk_any", C.c_void_p), ("clear_next", C.c_void_p), ("dst", C.c_void_p * 8),
                ("multicast", C.c_void_p)]


class CvrHostCsr(C.Structure):  # cvr_host_csr_t
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int64), ("nnz", C.c_int64),
                ("nnz_file", C.c_int64), ("val", c_double_p), ("col", c_int32_p),
                ("row_delim32", c_int32_p), ("row_delim64", c_int64_p)]


class CvrShardedInfo(C.Structure):  # cvr_sharded_info_t
    _fields_ = [("n_parts", C.c_int32), ("exchange", C.c_int32),
                ("n_rows", C.c_int64), ("n_cols", C.c_int64), ("nnz", C.c_int64),
                ("device", C.c_int32 * 8), ("row_begin", C.c_int64 * 8), ("row_end", C.c_int64 * 8),
                ("part_nnz", C.c_int64 * 8), ("part_chunks", C.c_int32 * 8), ("peer_bytes_per_iter", C.c_int64 * 8),
                ("create_seconds", C.c_double), ("convert_seconds", C.c_double), ("kernel_launches", C.c_int64)]


SHARD_PEER, SHARD_NCCL, SHARD_DENSE = 0, 1, 2
MM_REF_LAST_DELIM = 1
MM_KEEP_LAST_LINE = 2

# every symbol include/cvr_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "cvr_abi_version": (C.c_int, []),
    "cvr_last_error": (C.c_char_p, []),
    "cvr_device_init": (C.c_int, [C.c_int]),
    "cvr_read_matrix_market": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(CvrHostCsr)]),
    "cvr_free_host_csr": (None, [C.POINTER(CvrHostCsr)]),
    "cvr_record_ints": (C.c_int64, [C.c_int64, C.c_int32]),
    "cvr_auto_chunks": (C.c_int, [C.c_int64, C.c_int, c_int32_p]),
    "cvr_create": (C.c_int, [C.POINTER(CvrCsr), C.c_int32, C.c_int, C.POINTER(C.c_void_p)]),
    "cvr_create_from_device": (C.c_int, [C.POINTER(CvrCsr), C.c_int32, C.c_int, C.POINTER(C.c_void_p)]),
    "cvr_spmv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, c_double_p]),
    "cvr_spmv_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvr_spmv_publish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(CvrPublish), C.POINTER(C.c_void_p),
                                  C.c_int32, C.c_int32, C.c_uint32, C.c_int32, C.c_void_p]),
    "cvr_check_async_error": (C.c_int, [C.c_void_p]),
    "cvr_column_footprint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvr_chunk_needs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cvr_peer_alloc": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]),
    "cvr_peer_open": (C.c_int, [C.c_int, C.c_char_p, C.POINTER(C.c_void_p)]),
    "cvr_peer_close": (C.c_int, [C.c_int, C.c_void_p]),
    "cvr_peer_free": (C.c_int, [C.c_int, C.c_void_p]),
    "cvr_peer_barrier": (C.c_int, [C.c_int, C.POINTER(C.c_void_p), C.c_int32, C.c_int32, C.c_uint32, C.c_void_p]),
    "cvr_create_sharded": (C.c_int, [C.POINTER(CvrCsr), C.c_int32, C.POINTER(C.c_int), C.c_int, C.c_int,
                                    C.POINTER(C.c_void_p)]),
    "cvr_sharded_spmv": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, c_double_p]),
    "cvr_sharded_get_info": (C.c_int, [C.c_void_p, C.POINTER(CvrShardedInfo)]),
    "cvr_sharded_part": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]),
    "cvr_sharded_destroy": (None, [C.c_void_p]),
    "cvr_verify_csr": (C.c_int, [C.POINTER(CvrCsr), C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int,
                                c_int64_p, c_double_p, c_int64_p]),
    "cvr_export": (C.c_int, [C.c_void_p, C.POINTER(CvrArrays)]),
    "cvr_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cvr_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]),
    "cvr_get_info": (C.c_int, [C.c_void_p, C.POINTER(CvrInfo)]),
    "cvr_device_vectors": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "cvr_kernel_variant": (C.c_char_p, [C.c_void_p]),
    "cvr_device_arrays": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "cvr_set_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "cvr_get_kernel_timing": (C.c_int, [C.c_void_p, c_double_p, c_int64_p]),
    "cvr_destroy": (None, [C.c_void_p]),
}

_lib = None


class CvrError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libcvr_b200 error {code}: {message}")
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library is the only implementation of this path "
                "(no CPU fallback). Build it with `python -m cvr_b200.build`.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError = ABI drift, fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.cvr_abi_version() != 3:
            raise RuntimeError("libcvr_b200 ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise CvrError(rc, load().cvr_last_error().decode(errors="replace"))
