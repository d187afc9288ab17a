"""CPU suite: the C-ABI library loads and exports exactly what include/cvr_b200.h declares;
argument validation and the no-GPU failure mode (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cvr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvr_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(native_lib):
    from cvr_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(native_lib, n), f"{n} declared in include/cvr_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header drifted apart"
    assert native_lib.cvr_abi_version() == 3


def test_sm100a_code_is_in_the_library(native_lib):
    from cvr_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_record_ints_matches_reference_allocation(native_lib):
    # spmv.cpp:1806: 2 * (numRows + 240 + Nthrds * 32)
    assert native_lib.cvr_record_ints(916428, 68) == 2 * (916428 + 240 + 68 * 32)


def _csr_desc(n_rows=4, n_cols=4, nnz=16):
    from cvr_b200 import _lib
    val = np.ones(nnz)
    col = np.ones(nnz, dtype=np.int32)
    rd = np.array([0, 0, 4, 8, 12, 16][: n_rows + 2], dtype=np.int32)
    d = _lib.CvrCsr(n_rows, n_cols, nnz, val.ctypes.data, col.ctypes.data, rd.ctypes.data, 0)
    return d, (val, col, rd)


def test_argument_validation_happens_before_cuda(native_lib):
    h = C.c_void_p()
    d, keep = _csr_desc(nnz=16)
    d.nnz = 17
    assert native_lib.cvr_create(C.byref(d), 1, 0, C.byref(h)) == -1
    assert b"multiple of 16" in native_lib.cvr_last_error()
    d, keep = _csr_desc()
    assert native_lib.cvr_create(C.byref(d), 2, 0, C.byref(h)) == -1  # n_chunks > nnz/16
    d.row_delim64 = d.row_delim32
    assert native_lib.cvr_create(C.byref(d), 1, 0, C.byref(h)) == -1  # both delimiter widths set
    assert native_lib.cvr_create(None, 1, 0, C.byref(h)) == -1
    assert native_lib.cvr_spmv(None, None, None, 1, None) == -1


def test_no_gpu_means_loud_failure_not_fallback(native_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import cvr_b200
    h = C.c_void_p()
    d, keep = _csr_desc()
    assert native_lib.cvr_create(C.byref(d), 1, 0, C.byref(h)) == -2  # CVR_ERR_CUDA
    assert b"no CPU fallback" in native_lib.cvr_last_error()
    assert native_lib.cvr_device_init(0) == -2
    with pytest.raises(cvr_b200.CvrError):
        cvr_b200.CvrMatrix(cvr_b200.CsrMatrix(4, 4, *keep), 1)
    cli = os.path.join(ROOT, "cvr_b200", "bin", "spmv.cvr")
    p = subprocess.run([cli, os.path.join(ROOT, "tests", "golden", "kat12.mtx"), "1", "1"],
                       capture_output=True, text=True)
    assert p.returncode == 1 and "no CPU fallback" in p.stderr


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "cvr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(base, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "cvr_oracle" not in text, f
