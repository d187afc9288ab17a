"""CPU suite, build container only: the oracle port against the UNMODIFIED reference compiled
from /root/reference (oracle/_ref).  Skipped where that build is absent or the host lacks
AVX-512F; tests/golden/ carries the same evidence as committed fixtures."""
import numpy as np
import pytest

import oracle
from helpers import assert_y_close, to_oracle_csr

pytestmark = pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref not built / no AVX-512F")


def _cases():
    from cvr_b200 import gen
    return {
        "rand": lambda: gen.random_sparse(4000, 3000, 30000, seed=21, empty_frac=0.3),
        "long": lambda: gen.random_sparse(1500, 1500, 4000, seed=22, long_rows=5, long_len=1200),
        "web": lambda: gen.powerlaw_web(20000, 100000, seed=23),
        "fem": lambda: gen.fem27(10, 9, 8),
        "rmat": lambda: gen.rmat(11, 16, seed=24),
        "road": lambda: gen.road(30000, seed=25),
    }


@pytest.mark.parametrize("name", list(_cases()))
def test_port_equals_reference_conversion(name):
    csr = to_oracle_csr(_cases()[name]())
    for T in (1, 2, 3, 5, 8, 16, 61, 256, 1000):
        if T > csr.nnz // 16:
            continue
        a = oracle.convert(csr, T, "port")
        b = oracle.convert(csr, T, "ref")
        assert oracle.structure_equal(a, b) == [], f"{name} T={T}"


@pytest.mark.parametrize("name", ["rand", "fem", "road"])
def test_reference_kernel_agrees_at_host_thread_counts(name):
    """Where the reference kernel is valid (few chunks) all three agree."""
    csr = to_oracle_csr(_cases()[name]())
    x = np.random.default_rng(1).uniform(-1, 1, csr.n_cols + 1)
    for T in (1, 4, 8):
        cvr = oracle.convert(csr, T, "ref")
        y_ref, secs = oracle.spmv(cvr, csr.n_rows, x, "ref")
        y_port, _ = oracle.spmv(cvr, csr.n_rows, x, "port")
        assert secs is not None and secs >= 0
        assert_y_close(y_port, csr, x, f"port {name} T={T}")
        np.testing.assert_allclose(y_ref[1:csr.n_rows], y_port[1:csr.n_rows], rtol=0,
                                   atol=1e-13 * np.abs(y_port).max())


def test_reference_ingest_equals_port(tmp_path):
    from cvr_b200 import gen, write_mtx
    d = gen.random_sparse(300, 200, 1500, seed=31).to_host()
    rows = d.row_of_entry()[: d.nnz]
    keep = np.ones(d.nnz, bool)
    keep[np.flatnonzero(d.val == 0.0)] = False
    p = str(tmp_path / "m.mtx")
    write_mtx(p, d.n_rows, d.n_cols, rows[keep], d.col[keep], d.val[keep])
    a = oracle.read_mtx(p, "ref")
    b = oracle.read_mtx(p, "port", ref_last_delim=True)
    np.testing.assert_array_equal(a.col, b.col)
    np.testing.assert_array_equal(a.val, b.val)
    np.testing.assert_array_equal(a.row_delim, b.row_delim)
