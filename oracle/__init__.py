"""TEST INFRASTRUCTURE ONLY -- numpy-facing loader for the CPU checkers.

Two checkers live behind the same functions:

* ``impl="port"``: ``oracle/cvr_oracle.c``, the scalar C restatement of the reference
  path (``/root/reference/spmv.cpp``), built into ``oracle/_build/libcvr_oracle.so``.
* ``impl="ref"``: the UNMODIFIED reference compiled from where it lies into
  ``oracle/_ref/libcvr_ref.so`` (``oracle/ref_harness.cpp`` + ``oracle/Makefile``).  It only
  exists when the build container produced it and the host CPU has AVX-512F.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package.  The product (``cvr_b200/``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "_build", "libcvr_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libcvr_ref.so")
REF_CLI = os.path.join(HERE, "_ref", "spmv.cvr.ref")
REFERENCE_ROOT = "/root/reference"

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(ref: bool | None = None) -> None:
    """Compile the port (always) and the reference harness (when /root/reference exists)."""
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.run(["make", "-s", "-C", HERE, "port"], check=True, env=env)
    if ref is None:
        ref = os.path.exists(os.path.join(REFERENCE_ROOT, "spmv.cpp"))
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True, env=env)


def host_has_avx512f() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512f" in f.read()
    except OSError:
        return False


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        if not os.path.exists(PORT_SO):
            build(ref=False)
        lib = C.CDLL(PORT_SO)
        lib.cvr_oracle_record_ints.restype = C.c_longlong
        lib.cvr_oracle_record_ints.argtypes = [C.c_int, C.c_int]
        lib.cvr_oracle_record_offset.restype = C.c_longlong
        lib.cvr_oracle_record_offset.argtypes = [C.c_int, C.c_int]
        lib.cvr_oracle_convert.restype = C.c_int
        lib.cvr_oracle_convert.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, _ip,
                                           _dp, _ip, _ip, _ip, _ip, _ip, C.c_int]
        lib.cvr_oracle_spmv.restype = C.c_int
        lib.cvr_oracle_spmv.argtypes = [C.c_int, C.c_int, _dp, _ip, _ip, _ip, _ip, _ip, _dp, _dp]
        lib.cvr_oracle_csr_spmv.restype = None
        lib.cvr_oracle_csr_spmv.argtypes = [C.c_int, _dp, _ip, _ip, _dp, _dp, _dp]
        lib.cvr_oracle_read_mtx.restype = C.c_int
        lib.cvr_oracle_read_mtx.argtypes = [C.c_char_p, C.c_int, C.POINTER(_dp), C.POINTER(_ip),
                                            C.POINTER(_ip), _ip, _ip, _ip, _ip]
        lib.cvr_oracle_free.restype = None
        lib.cvr_oracle_free.argtypes = [C.c_void_p]
        _port = lib
    return _port


def ref_available() -> bool:
    return os.path.exists(REF_SO) and host_has_avx512f()


def ref_lib():
    """The unmodified reference, function level.  None when it cannot run on this host."""
    global _ref
    if _ref is None:
        if not ref_available():
            return None
        lib = C.CDLL(REF_SO)
        lib.cvr_ref_read_matrix.restype = C.c_int
        lib.cvr_ref_read_matrix.argtypes = [C.c_char_p, C.POINTER(_dp), C.POINTER(_ip),
                                            C.POINTER(_ip), _ip, _ip, _ip]
        lib.cvr_ref_free.restype = None
        lib.cvr_ref_free.argtypes = [C.c_void_p]
        lib.cvr_ref_record_ints.restype = C.c_longlong
        lib.cvr_ref_record_ints.argtypes = [C.c_int, C.c_int]
        lib.cvr_ref_preprocess.restype = C.c_double
        lib.cvr_ref_preprocess.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, _ip,
                                           _dp, _ip, _ip, _ip, _ip, _ip]
        lib.cvr_ref_spmv.restype = C.c_double
        lib.cvr_ref_spmv.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _ip, _ip, _ip, _ip, _ip,
                                     _dp, _dp, C.c_int]
        _ref = lib
    return _ref


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def aligned(n: int, dtype, fill=None) -> np.ndarray:
    """64-byte aligned array (the reference stores with aligned AVX-512 instructions)."""
    dtype = np.dtype(dtype)
    raw = np.empty(n * dtype.itemsize + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    out = raw[off:off + n * dtype.itemsize].view(dtype)
    if fill is not None:
        out[...] = fill
    return out


SENTINEL = -777777


class Csr:
    """1-based CSR exactly as readMatrix builds it (spmv.cpp:489-526): ``row_delim`` has
    n_rows+2 entries, row 0 is a phantom empty row, ``col`` holds 1..n_cols, nnz % 16 == 0."""

    def __init__(self, n_rows, n_cols, val, col, row_delim, nnz_file=None):
        self.n_rows = int(n_rows)
        self.n_cols = int(n_cols)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        self.row_delim = np.ascontiguousarray(row_delim, dtype=np.int32)
        self.nnz = int(self.val.shape[0])
        self.nnz_file = int(nnz_file) if nnz_file is not None else self.nnz
        assert self.nnz % 16 == 0 and self.row_delim.shape[0] == self.n_rows + 2


def record_offset(chunk: int, first_row: int) -> int:
    return (2 * (32 * chunk + first_row)) // 16 * 16


def convert(csr: Csr, n_chunks: int, impl: str = "port", fill_missing_tail: bool = False) -> dict:
    """CSR -> CVR.  Returns the reference-shaped arrays; unwritten ints hold SENTINEL."""
    T = int(n_chunks)
    n_rec_ints = int(2 * (csr.n_rows + 240 + 32 * T))
    out = {
        "n_chunks": T,
        "vals": aligned(csr.nnz, np.float64, 0.0),
        "cols": aligned(csr.nnz, np.int32, 0),
        "record": aligned(n_rec_ints, np.int32, SENTINEL),
        "nnz_rows": aligned(4 * T, np.int32, SENTINEL),
        "final_2": aligned(16 * T, np.int32, SENTINEL),
        "split": aligned(2 * T, np.int32, 0),
    }
    val = aligned(csr.nnz, np.float64)
    val[...] = csr.val
    col = aligned(csr.nnz, np.int32)
    col[...] = csr.col
    rd = aligned(csr.n_rows + 2 + 16, np.int32, csr.row_delim[-1])
    rd[:csr.n_rows + 2] = csr.row_delim
    if impl == "port":
        rc = port_lib().cvr_oracle_convert(T, csr.nnz, csr.n_rows, _d(val), _i(col), _i(rd),
                                           _d(out["vals"]), _i(out["cols"]), _i(out["record"]),
                                           _i(out["nnz_rows"]), _i(out["final_2"]),
                                           _i(out["split"]), int(fill_missing_tail))
        if rc != 0:
            raise ValueError(f"cvr_oracle_convert rejected n_chunks={T} nnz={csr.nnz}")
        out["seconds"] = None
    elif impl == "ref":
        lib = ref_lib()
        if lib is None:
            raise RuntimeError("reference build (oracle/_ref) unavailable on this host")
        if T > csr.nnz // 16:
            raise ValueError("n_chunks > nnz/16")
        out["seconds"] = lib.cvr_ref_preprocess(T, csr.nnz, csr.n_rows, _d(val), _i(col), _i(rd),
                                                _d(out["vals"]), _i(out["cols"]),
                                                _i(out["record"]), _i(out["nnz_rows"]),
                                                _i(out["final_2"]), _i(out["split"]))
    else:
        raise ValueError(impl)
    return out


def chunk_records(cvr: dict, t: int) -> np.ndarray:
    """(n,2) view of chunk t's record pairs up to and including its eight pos=-1 entries."""
    r0 = int(cvr["nnz_rows"][4 * t + 2])
    off = record_offset(t, r0)
    rec = cvr["record"][off:]
    n = 0
    while rec[2 * n] != -1:
        n += 1
    return rec[:2 * (n + 8)].reshape(-1, 2)


def structure_equal(a: dict, b: dict, compare_tail: bool = True) -> list:
    """Bit-exact comparison of the CVR contract: vals, cols, per chunk nnz_rows,
    split, tail[8] and the record list through its eight terminators.  Returns mismatches."""
    bad = []
    if a["n_chunks"] != b["n_chunks"]:
        return ["n_chunks"]
    if not np.array_equal(a["vals"].view(np.uint64), b["vals"].view(np.uint64)):
        bad.append("vals")
    if not np.array_equal(a["cols"], b["cols"]):
        bad.append("cols")
    if not np.array_equal(a["nnz_rows"], b["nnz_rows"]):
        bad.append("nnz_rows")
        return bad
    if not np.array_equal(a["split"], b["split"]):
        bad.append("split")
    for t in range(a["n_chunks"]):
        ra, rb = chunk_records(a, t), chunk_records(b, t)
        if ra.shape != rb.shape or not np.array_equal(ra, rb):
            bad.append(f"record[{t}]")
        if compare_tail and not np.array_equal(a["final_2"][16 * t:16 * t + 8],
                                              b["final_2"][16 * t:16 * t + 8]):
            bad.append(f"tail[{t}]")
        if len(bad) > 8:
            break
    return bad


def spmv(cvr: dict, n_rows: int, x: np.ndarray, impl: str = "port", iters: int = 1):
    """y (n_rows+1 entries, index 0 = phantom row) from the CVR arrays."""
    T = cvr["n_chunks"]
    x = np.ascontiguousarray(x, dtype=np.float64)
    xa = aligned(x.shape[0] + 16, np.float64, 0.0)
    xa[:x.shape[0]] = x
    y = aligned(n_rows + 2, np.float64, 0.0)
    if impl == "port":
        port_lib().cvr_oracle_spmv(T, n_rows, _d(cvr["vals"]), _i(cvr["cols"]), _i(cvr["record"]),
                                   _i(cvr["nnz_rows"]), _i(cvr["final_2"]), _i(cvr["split"]),
                                   _d(xa), _d(y))
        return y[:n_rows + 1].copy(), None
    lib = ref_lib()
    if lib is None:
        raise RuntimeError("reference build (oracle/_ref) unavailable on this host")
    nnz = int(cvr["vals"].shape[0])
    # the reference prefetch reads cols[] a step past the chunk end (spmv.cpp:1170): pad
    cols = aligned(nnz + 32, np.int32, 0)
    cols[:nnz] = cvr["cols"]
    secs = lib.cvr_ref_spmv(T, nnz, n_rows, _d(cvr["vals"]), _i(cols), _i(cvr["record"]),
                            _i(cvr["nnz_rows"]), _i(cvr["final_2"]), _i(cvr["split"]),
                            _d(xa), _d(y), int(iters))
    return y[:n_rows + 1].copy(), secs


def csr_spmv(csr: Csr, x: np.ndarray):
    """(y, sum|a*x|) for rows 0..n_rows by the reference's scalar check loop."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.zeros(csr.n_rows + 1)
    mag = np.zeros(csr.n_rows + 1)
    port_lib().cvr_oracle_csr_spmv(csr.n_rows, _d(csr.val), _i(csr.col), _i(csr.row_delim),
                                   _d(x), _d(y), _d(mag))
    return y, mag


def read_mtx(path: str, impl: str = "port", ref_last_delim: bool = False) -> Csr:
    val, col, rd = _dp(), _ip(), _ip()
    nnzp, nnzf, nr, nc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    if impl == "port":
        lib = port_lib()
        rc = lib.cvr_oracle_read_mtx(path.encode(), int(ref_last_delim), C.byref(val), C.byref(col),
                                     C.byref(rd), C.byref(nnzp), C.byref(nnzf), C.byref(nr),
                                     C.byref(nc))
        if rc != 0:
            raise IOError(f"cvr_oracle_read_mtx({path}) -> {rc}")
        free = lib.cvr_oracle_free
        nnz_file = nnzf.value
    else:
        lib = ref_lib()
        if lib is None:
            raise RuntimeError("reference build (oracle/_ref) unavailable on this host")
        lib.cvr_ref_read_matrix(path.encode(), C.byref(val), C.byref(col), C.byref(rd),
                                C.byref(nnzp), C.byref(nr), C.byref(nc))
        free = lib.cvr_ref_free
        nnz_file = None
    n = nnzp.value
    out = Csr(nr.value, nc.value,
              np.ctypeslib.as_array(val, (n,)).copy(),
              np.ctypeslib.as_array(col, (n,)).copy(),
              np.ctypeslib.as_array(rd, (nr.value + 2,)).copy(), nnz_file)
    for p in (val, col, rd):
        free(C.cast(p, C.c_void_p))
    return out
