"""cvr_b200 -- B200-native CVR SpMV (CSR->CVR conversion + CVR SpMV) behind a C ABI.

Python here is plumbing: ctypes over include/cvr_b200.h.  The product is libcvr_b200.so
(hand-written sm_100a CUDA) and the `spmv.cvr` command line, both built by cvr_b200.build.
"""
from ._lib import CvrError, load as load_library  # noqa: F401
from .csr import CsrMatrix, read_matrix, write_mtx  # noqa: F401
from .matrix import CvrMatrix, DeviceCsr, ShardedCvr, pre_processing, spmv_compute_kernel, verify_csr  # noqa: F401

__all__ = ["CvrError", "load_library", "CsrMatrix", "read_matrix", "write_mtx", "CvrMatrix",
           "DeviceCsr", "ShardedCvr", "pre_processing", "spmv_compute_kernel", "verify_csr"]
