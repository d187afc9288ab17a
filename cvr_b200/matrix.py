"""Host-side mirror of the reference's function-level interface for the hot path.

The reference exposes three functions to its main():
  readMatrix(...)            spmv.cpp:311   -> cvr_b200.read_matrix
  pre_processing(...)        spmv.cpp:565   -> cvr_b200.pre_processing  / CvrMatrix(...)
  spmv_compute_kernel(...)   spmv.cpp:1016  -> cvr_b200.spmv_compute_kernel / CvrMatrix.spmv

Everything here is a thin wrapper over the C ABI (include/cvr_b200.h); the arithmetic runs in
the CUDA kernels of libcvr_b200.so.  Without that library, or without a GPU, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .csr import CsrMatrix


def _ptr(a) -> int:
    """Raw address of a numpy array, a torch tensor, or an int."""
    if a is None:
        return 0
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return int(a.data_ptr())
    raise TypeError(type(a))


class DeviceCsr:
    """A CSR whose arrays live on a CUDA device as torch tensors (val f64, col i32,
    row_delim i32 or i64), same conventions as CsrMatrix.  Produced by cvr_b200.gen."""

    def __init__(self, n_rows, n_cols, val, col, row_delim, nnz_true=None):
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.val, self.col, self.row_delim = val, col, row_delim
        self.nnz = int(val.shape[0])
        self.nnz_true = int(nnz_true) if nnz_true is not None else self.nnz

    @property
    def device(self):
        return self.val.device

    def to_host(self) -> CsrMatrix:
        return CsrMatrix(self.n_rows, self.n_cols, self.val.cpu().numpy(), self.col.cpu().numpy(),
                         self.row_delim.cpu().numpy(), self.nnz_true)


class CvrMatrix:
    """A matrix converted to CVR and resident on one GPU.  Construction = pre_processing."""

    def __init__(self, csr, n_chunks: int = 0, device: int = 0):
        lib = _lib.load()
        self._lib = lib
        self._h = C.c_void_p()
        desc = _csr_desc(csr)
        if isinstance(csr, DeviceCsr):
            if csr.device.type != "cuda":
                raise ValueError("DeviceCsr must live on a CUDA device (no CPU path)")
            device = csr.device.index if csr.device.index is not None else device
            import torch
            torch.cuda.synchronize(device)  # the generator's stream is not ours
            rc = lib.cvr_create_from_device(C.byref(desc), int(n_chunks), int(device), C.byref(self._h))
        else:
            rc = lib.cvr_create(C.byref(desc), int(n_chunks), int(device), C.byref(self._h))
        _lib.check(rc)
        self.n_rows, self.n_cols, self.nnz = csr.n_rows, csr.n_cols, csr.nnz
        self.nnz_true = getattr(csr, "nnz_true", csr.nnz)

    # -- serialisation
    def save(self, path: str) -> None:
        _lib.check(self._lib.cvr_save(self._h, os.fsencode(path)))

    @classmethod
    def load(cls, path: str, device: int = 0) -> "CvrMatrix":
        """A matrix converted earlier (CvrMatrix.save): no CSR, no conversion."""
        self = cls.__new__(cls)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.cvr_load(os.fsencode(path), int(device), C.byref(self._h)))
        i = self.info
        self.n_rows, self.n_cols, self.nnz, self.nnz_true = i["n_rows"], i["n_cols"], i["nnz"], i["nnz"]
        return self

    # -- lifetime
    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cvr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- queries
    @property
    def info(self) -> dict:
        i = _lib.CvrInfo()
        _lib.check(self._lib.cvr_get_info(self._h, C.byref(i)))
        return {name: getattr(i, name) for name, _ in _lib.CvrInfo._fields_}

    @property
    def n_chunks(self) -> int:
        return self.info["n_chunks"]

    def device_vectors(self):
        x, y = C.c_void_p(), C.c_void_p()
        _lib.check(self._lib.cvr_device_vectors(self._h, C.byref(x), C.byref(y)))
        return x.value, y.value

    @property
    def kernel_name(self) -> str:
        """The sweep geometry picked for this matrix ("tile7x5r", "tile11x5"; CVR_SPMV_KERNEL overrides)."""
        return self._lib.cvr_kernel_variant(self._h).decode()

    def column_footprint(self, used_dev, stream: int = 0) -> None:
        """used_dev[c] = 1 (uint8, n_cols+1 entries, pre-zeroed) for every column id this matrix touches."""
        _lib.check(self._lib.cvr_column_footprint(self._h, _ptr(used_dev), int(stream)))

    def check_async_error(self) -> None:
        """Raises CvrError(CVR_ERR_STATE) if a peer flag barrier of this handle timed out."""
        _lib.check(self._lib.cvr_check_async_error(self._h))

    def device_arrays(self):
        """Raw device addresses (vals, cols, record) of the converted matrix, for measurement tools."""
        v, c, r = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(self._lib.cvr_device_arrays(self._h, C.byref(v), C.byref(c), C.byref(r)))
        return v.value, c.value, r.value

    # -- the hot path
    def spmv(self, x, iters: int = 1):
        """spmv_compute_kernel with host vectors: returns (y[n_rows+1], seconds per iteration).
        x has n_cols+1 entries (index 0 is the phantom column)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape[0] != self.n_cols + 1:
            raise ValueError(f"x must have n_cols+1 = {self.n_cols + 1} entries")
        y = np.empty(self.n_rows + 1, dtype=np.float64)
        secs = C.c_double()
        _lib.check(self._lib.cvr_spmv(self._h, x.ctypes.data, y.ctypes.data, int(iters), C.byref(secs)))
        return y, secs.value

    def spmv_into(self, x_host, y_host, iters: int = 1) -> float:
        """Same call on caller-owned (e.g. pinned) host buffers given as raw addresses/arrays."""
        secs = C.c_double()
        _lib.check(self._lib.cvr_spmv(self._h, _ptr(x_host), _ptr(y_host), int(iters), C.byref(secs)))
        return secs.value

    def spmv_device(self, x_dev, y_dev, stream: int = 0) -> None:
        """Enqueue one SpMV on device vectors (torch tensors or raw pointers) on `stream`."""
        _lib.check(self._lib.cvr_spmv_device(self._h, _ptr(x_dev), _ptr(y_dev), int(stream)))

    def spmv_publish(self, x_dev, y_dev, pub, flag_arrays, rank, n_ranks, epoch, y_is_clear, stream: int = 0) -> None:
        """One iteration of the row-sharded iterated SpMV: the sweep publishes finished rows into the
        next iteration's x vectors (own + peer-mapped), an epilogue kernel publishes the accumulated
        rows and runs the flag barrier.  See cvr_b200.dist.PeerPublisher."""
        _lib.check(self._lib.cvr_spmv_publish(self._h, _ptr(x_dev), _ptr(y_dev), C.byref(pub), flag_arrays,
                                              int(rank), int(n_ranks), int(epoch), int(y_is_clear), int(stream)))

    # -- measurement aid
    def set_kernel_timing(self, enabled: bool) -> None:
        _lib.check(self._lib.cvr_set_kernel_timing(self._h, int(enabled)))

    def kernel_timing(self):
        """(summed SpMV-kernel seconds, launches) since timing was enabled / last read."""
        secs, n = C.c_double(), C.c_int64()
        _lib.check(self._lib.cvr_get_kernel_timing(self._h, C.byref(secs), C.byref(n)))
        return secs.value, n.value

    # -- the bit-exact gate
    def export(self) -> dict:
        """CVR structure arrays in the reference's layout and sizes (cvr_export)."""
        info = self.info
        T = info["n_chunks"]
        out = {
            "n_chunks": T,
            "vals": np.empty(self.nnz, dtype=np.float64),
            "cols": np.empty(self.nnz, dtype=np.int32),
            "record": np.empty(info["record_ints"], dtype=np.int32),
            "nnz_rows": np.empty(4 * T, dtype=np.int32),
            "final_2": np.full(16 * T, -777777, dtype=np.int32),
            "split": np.empty(2 * T, dtype=np.int32),
        }
        arr = _lib.CvrArrays(*(out[k].ctypes.data for k in ("vals", "cols", "record", "nnz_rows", "final_2", "split")))
        _lib.check(self._lib.cvr_export(self._h, C.byref(arr)))
        return out


class ShardedCvr:
    """The matrix row-sharded over several GPUs by ONE process (cvr_create_sharded): thin wrapper over the
    C++ multi-device host of libcvr_b200.  `devices` may repeat a device (several shards on one GPU)."""

    def __init__(self, csr: CsrMatrix, devices, n_chunks: int = 0, exchange: str = "peer", dense: bool = False):
        lib = _lib.load()
        self._lib = lib
        self._h = C.c_void_p()
        desc = _csr_desc(csr)
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        flags = (_lib.SHARD_NCCL if exchange == "nccl" else _lib.SHARD_PEER) | (_lib.SHARD_DENSE if dense else 0)
        _lib.check(lib.cvr_create_sharded(C.byref(desc), int(n_chunks), devs, len(devices), flags, C.byref(self._h)))
        self.n_rows, self.n_cols = csr.n_rows, csr.n_cols

    @property
    def info(self) -> dict:
        i = _lib.CvrShardedInfo()
        _lib.check(self._lib.cvr_sharded_get_info(self._h, C.byref(i)))
        n = i.n_parts
        out = {name: getattr(i, name) for name, _ in _lib.CvrShardedInfo._fields_}
        for k in ("device", "row_begin", "row_end", "part_nnz", "part_chunks", "peer_bytes_per_iter"):
            out[k] = list(out[k])[:n]
        return out

    def spmv(self, x, iters: int = 1, feed_y_to_x: bool = False):
        """(y[n_rows+1], seconds per iteration): `iters` SpMVs with the same x, or `iters` iterations of
        x <- A x with one exchange each when feed_y_to_x."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.shape[0] != self.n_cols + 1:
            raise ValueError(f"x must have n_cols+1 = {self.n_cols + 1} entries")
        y = np.empty(self.n_rows + 1, dtype=np.float64)
        secs = C.c_double()
        _lib.check(self._lib.cvr_sharded_spmv(self._h, x.ctypes.data, y.ctypes.data, int(iters), int(feed_y_to_x),
                                              C.byref(secs)))
        return y, secs.value

    def part_export(self, part: int) -> dict:
        """CVR structure arrays of one shard in the reference layout (the bit-exact gate per shard)."""
        h = C.c_void_p()
        _lib.check(self._lib.cvr_sharded_part(self._h, int(part), C.byref(h)))
        view = CvrMatrix.__new__(CvrMatrix)
        view._lib, view._h = self._lib, h
        i = view.info
        view.n_rows, view.n_cols, view.nnz, view.nnz_true = i["n_rows"], i["n_cols"], i["nnz"], i["nnz"]
        try:
            return view.export()
        finally:
            view._h = C.c_void_p()  # owned by the sharded handle

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.cvr_sharded_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def _csr_desc(csr):
    desc = _lib.CvrCsr()
    desc.n_rows, desc.n_cols, desc.nnz = csr.n_rows, csr.n_cols, csr.nnz
    desc.val, desc.col = _ptr(csr.val), _ptr(csr.col)
    rd = csr.row_delim
    wide = (rd.dtype == np.int64) if isinstance(rd, np.ndarray) else ("int64" in str(rd.dtype))
    desc.row_delim32 = 0 if wide else _ptr(rd)
    desc.row_delim64 = _ptr(rd) if wide else 0
    return desc


def verify_csr(csr: DeviceCsr, x_dev, y_dev, rel_tol: float = 1e-12, check_row0: bool = True) -> dict:
    """The self-check on the device (cvr_verify_csr): y against the CSR product, row by row,
    |dy| <= rel_tol * sum|a x|.  Returns {"rows_failing", "max_rel", "first_bad_row"}."""
    lib = _lib.load()
    import torch
    device = csr.device.index if csr.device.index is not None else 0
    torch.cuda.synchronize(device)
    bad, first, rel = C.c_int64(), C.c_int64(), C.c_double()
    desc = _csr_desc(csr)
    _lib.check(lib.cvr_verify_csr(C.byref(desc), int(device), _ptr(x_dev), _ptr(y_dev), float(rel_tol),
                                  int(check_row0), C.byref(bad), C.byref(rel), C.byref(first)))
    return {"rows_failing": bad.value, "max_rel": rel.value, "first_bad_row": first.value}


def pre_processing(csr, n_threads: int = 0, device: int = 0) -> CvrMatrix:
    """CSR -> CVR on the GPU (spmv.cpp:565).  n_threads = number of chunks; 0 = auto."""
    return CvrMatrix(csr, n_threads, device)


def spmv_compute_kernel(cvr: CvrMatrix, x, n_times: int = 1):
    """y = A x, n_times iterations (spmv.cpp:1016).  Returns (y, seconds per iteration)."""
    return cvr.spmv(x, n_times)
