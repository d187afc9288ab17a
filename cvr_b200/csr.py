"""Host-side CSR in the reference's conventions, and the ingest entry point.

`read_matrix` mirrors readMatrix (/root/reference/spmv.cpp:311): Matrix Market file ->
1-based CSR padded to a multiple of 16.  It calls the library's C++ reader through the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib


class CsrMatrix:
    """1-based CSR as readMatrix builds it (spmv.cpp:489-526): `row_delim` has n_rows+2 entries,
    row 0 is an empty phantom row, `col` holds 1..n_cols and nnz % 16 == 0."""

    def __init__(self, n_rows, n_cols, val, col, row_delim, nnz_true=None):
        self.n_rows = int(n_rows)
        self.n_cols = int(n_cols)
        self.val = np.ascontiguousarray(val, dtype=np.float64)
        self.col = np.ascontiguousarray(col, dtype=np.int32)
        rd = np.asarray(row_delim)
        wide = self.val.shape[0] > 0x7FFFFFFF
        self.row_delim = np.ascontiguousarray(rd, dtype=np.int64 if wide else np.int32)
        self.nnz = int(self.val.shape[0])
        self.nnz_true = int(nnz_true) if nnz_true is not None else self.nnz
        if self.nnz % 16 != 0 or self.nnz < 16:
            raise ValueError("nnz must be a positive multiple of 16 (spmv.cpp:457)")
        if self.row_delim.shape[0] != self.n_rows + 2 or self.col.shape[0] != self.nnz:
            raise ValueError("inconsistent CSR array lengths")

    @staticmethod
    def from_coo(rows, cols, vals, n_rows, n_cols) -> "CsrMatrix":
        """Build from 1-based coordinates the way readMatrix does: values rounded to float32
        (spmv.cpp:65), zero-valued copies of the LAST given entry pad nnz to a multiple of 16
        (:474-482), stable (row, col) order (:485)."""
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        vals = np.asarray(vals, dtype=np.float32).astype(np.float64)
        n = rows.shape[0]
        if n == 0:
            raise ValueError("no entries")
        npad = n if n % 16 == 0 else (n + 16) // 16 * 16
        if npad > n:
            rows = np.concatenate([rows, np.full(npad - n, rows[-1])])
            cols = np.concatenate([cols, np.full(npad - n, cols[-1])])
            vals = np.concatenate([vals, np.zeros(npad - n)])
        order = np.lexsort((cols, rows))  # stable
        rows, cols, vals = rows[order], cols[order], vals[order]
        counts = np.bincount(rows, minlength=n_rows + 1)[: n_rows + 1]
        rd = np.zeros(n_rows + 2, dtype=np.int64)
        rd[1:] = np.cumsum(counts)
        return CsrMatrix(n_rows, n_cols, vals, cols, rd, nnz_true=n)

    def row_of_entry(self) -> np.ndarray:
        return np.repeat(np.arange(self.n_rows + 1), np.diff(self.row_delim.astype(np.int64)))


def write_mtx(path: str, n_rows: int, n_cols: int, rows, cols, vals, column_major: bool = True) -> None:
    """Matrix Market `coordinate real general` file from 1-based coordinates, written
    column-major like the SuiteSparse collection and ending with a newline (the reference
    drops an unterminated last line, spmv.cpp:411).  Values are printed with 9 significant
    digits, enough to round-trip the float32 the reference parses (spmv.cpp:432)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float64)
    if column_major:
        order = np.lexsort((rows, cols))
        rows, cols, vals = rows[order], cols[order], vals[order]
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write(f"{n_rows} {n_cols} {rows.shape[0]}\n")
        np.savetxt(f, np.column_stack([rows, cols, vals]), fmt="%d %d %.9g")


def read_matrix(path: str, ref_last_delim: bool = False, keep_last_line: bool = False) -> CsrMatrix:
    """readMatrix (spmv.cpp:311) through the C ABI (`cvr_read_matrix_market`)."""
    lib = _lib.load()
    out = _lib.CvrHostCsr()
    flags = (_lib.MM_REF_LAST_DELIM if ref_last_delim else 0) | (_lib.MM_KEEP_LAST_LINE if keep_last_line else 0)
    _lib.check(lib.cvr_read_matrix_market(os.fsencode(path), flags, C.byref(out)))
    try:
        n = out.nnz
        val = np.ctypeslib.as_array(out.val, (n,)).copy()
        col = np.ctypeslib.as_array(out.col, (n,)).copy()
        if out.row_delim32:
            rd = np.ctypeslib.as_array(out.row_delim32, (out.n_rows + 2,)).copy()
        else:
            rd = np.ctypeslib.as_array(out.row_delim64, (out.n_rows + 2,)).copy()
        return CsrMatrix(out.n_rows, out.n_cols, val, col, rd, nnz_true=out.nnz_file)
    finally:
        lib.cvr_free_host_csr(C.byref(out))
