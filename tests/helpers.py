"""Shared helpers for the parity tests (test infrastructure)."""
import glob
import os

import numpy as np

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-12  # BASELINE.json: per-row |dy| <= 1e-12 * sum_j |a_ij x_j|


def golden_cases():
    return sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_*.npz"))
                  if not p.endswith("ref_ingest.npz"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    csr = oracle.Csr(int(z["n_rows"]), int(z["n_cols"]), z["csr_val"], z["csr_col"], z["csr_rd"])
    return z, csr


def golden_structure(z, T):
    """The reference's output for chunk count T in the dict shape oracle.convert returns,
    with the record lists re-laid at the reference offsets."""
    nnz_rows = z[f"T{T}_nnz_rows"]
    n_rows = int(z["n_rows"])
    record = np.full(2 * (n_rows + 240 + 32 * T), oracle.SENTINEL, dtype=np.int32)
    recs, lens = z[f"T{T}_records"], z[f"T{T}_record_lens"]
    at = 0
    for t in range(T):
        off = oracle.record_offset(t, int(nnz_rows[4 * t + 2]))
        n = int(lens[t])
        record[off:off + 2 * n] = recs[at:at + n].reshape(-1)
        at += n
    final_2 = np.full(16 * T, oracle.SENTINEL, dtype=np.int32)
    final_2.reshape(T, 16)[:, :8] = z[f"T{T}_tail"]
    return {"n_chunks": T, "vals": z[f"T{T}_vals"], "cols": z[f"T{T}_cols"], "record": record,
            "nnz_rows": nnz_rows, "final_2": final_2, "split": z[f"T{T}_split"]}


def tail_written(golden, t):
    return golden["final_2"][16 * t] != oracle.SENTINEL


def assert_structure_equal(got, want, what=""):
    """Bit-exact CVR contract.  Chunks whose tail the reference never wrote (sentinel) are
    compared without the tail."""
    T = want["n_chunks"]
    assert got["n_chunks"] == T
    np.testing.assert_array_equal(got["nnz_rows"], want["nnz_rows"], err_msg=f"{what} nnz_rows")
    np.testing.assert_array_equal(got["split"], want["split"], err_msg=f"{what} split")
    assert np.array_equal(np.asarray(got["vals"]).view(np.uint64), np.asarray(want["vals"]).view(np.uint64)), f"{what} vals"
    np.testing.assert_array_equal(got["cols"], want["cols"], err_msg=f"{what} cols")
    for t in range(T):
        np.testing.assert_array_equal(oracle.chunk_records(got, t), oracle.chunk_records(want, t),
                                      err_msg=f"{what} record list of chunk {t}")
        if tail_written(want, t):
            np.testing.assert_array_equal(got["final_2"][16 * t:16 * t + 8], want["final_2"][16 * t:16 * t + 8],
                                          err_msg=f"{what} tail of chunk {t}")


def assert_y_close(y, csr, x, what=""):
    """Per-row |y - y_csr| <= 1e-12 * sum|a x| against the reference's scalar CSR loop."""
    yc, mag = oracle.csr_spmv(csr, x)
    err = np.abs(np.asarray(y) - yc)
    bad = np.flatnonzero(err[1:] > REL_TOL * mag[1:]) + 1
    assert bad.size == 0, f"{what}: {bad.size} rows out of tolerance, first {bad[:5]}, err {err[bad[:5]]}, mag {mag[bad[:5]]}"
    assert y[0] == 0.0, f"{what}: phantom row 0 must stay 0"


def to_oracle_csr(m):
    """cvr_b200.CsrMatrix / DeviceCsr -> oracle.Csr"""
    if hasattr(m, "to_host"):
        m = m.to_host()
    return oracle.Csr(m.n_rows, m.n_cols, m.val, m.col, m.row_delim.astype(np.int32), m.nnz_true)
