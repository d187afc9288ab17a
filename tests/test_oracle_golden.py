"""CPU suite: the oracle port against the reference's own outputs (tests/golden, produced by
tests/golden/make_golden.py from the unmodified reference) and against the known-answer test
of SURVEY.md Appendix C."""
import os

import numpy as np
import pytest

import oracle
from helpers import (GOLDEN, assert_structure_equal, assert_y_close, golden_cases, golden_structure,
                     load_golden)


def test_kat12_known_answer():
    """arrays the reference prints for kat12.mtx at T=1 and T=2."""
    m = oracle.read_mtx(os.path.join(GOLDEN, "kat12.mtx"), "port", ref_last_delim=True)
    assert m.row_delim.tolist() == [0, 0, 3, 4, 4, 6, 11, 12, 13, 17, 19, 20, 26, 31]
    assert m.col.tolist() == [2, 4, 9, 7, 1, 6, 2, 4, 7, 9, 12, 7, 10, 1, 4, 6, 11, 4, 9, 7, 1, 3, 6, 8, 10,
                              11, 1, 2, 4, 6, 9, 11]
    m = oracle.read_mtx(os.path.join(GOLDEN, "kat12.mtx"), "port")
    c = oracle.convert(m, 1)
    assert c["cols"].tolist() == [2, 7, 4, 1, 2, 7, 10, 1, 4, 7, 9, 6, 4, 1, 1, 4, 9, 7, 3, 8, 12, 11, 2, 6, 4,
                                  9, 6, 10, 6, 9, 11, 11]
    assert c["vals"].tolist() == [102, 207, 904, 401, 502, 607, 710, 801, 104, 1007, 909, 406, 504, 1101,
                                  1201, 804, 109, 507, 1103, 1108, 512, 1111, 1202, 806, 1204, 509, 1106,
                                  1110, 1206, 1209, 1211, 811]
    assert c["nnz_rows"].tolist() == [0, 32, 1, 12]
    assert c["split"].tolist() == [0, 14]
    assert c["final_2"][:8].tolist() == [1, 10, 9, 4, 5, 11, 12, 8]
    assert oracle.chunk_records(c, 0).tolist() == [
        [2, 3], [9, 2], [13, 6], [14, 7], [17, 1], [18, 2], [19, 3], [24, 0], [28, 4], [29, 5],
        [-1, 6], [-1, 4], [-1, 5], [-1, 5], [-1, 6], [-1, 6], [-1, 6], [-1, 7]]
    c2 = oracle.convert(m, 2)
    assert c2["nnz_rows"].tolist() == [0, 16, 1, 8, 16, 32, 8, 12]
    assert c2["split"].tolist() == [0, -1, 0, -1]
    assert c2["final_2"][:8].tolist() == [1, 2, 3, 4, 5, 6, 7, 8]
    assert c2["final_2"][16:24].tolist() == [8, 9, 10, 11, 12, 0, 0, 0]
    assert oracle.record_offset(1, 8) == 80
    assert oracle.chunk_records(c2, 1).tolist() == [
        [5, 5], [6, 6], [7, 7], [8, 0], [10, 2],
        [-1, 4], [-1, 1], [-1, 4], [-1, 3], [-1, 4], [-1, 3], [-1, 3], [-1, 4]]
    want_y = [0, 315, 207, 0, 807, 2534, 607, 710, 3222, 1813, 1007, 6639, 7233]
    for cc in (c, c2):
        y, _ = oracle.spmv(cc, 12, np.ones(13))
        assert y.tolist() == want_y


@pytest.mark.parametrize("name", golden_cases())
def test_port_conversion_matches_reference_fixture(name):
    z, csr = load_golden(name)
    for T in z["chunk_counts"]:
        T = int(T)
        got = oracle.convert(csr, T, "port")
        assert_structure_equal(got, golden_structure(z, T), f"{name} T={T}")


@pytest.mark.parametrize("name", golden_cases())
def test_port_spmv_matches_csr_and_reference_y(name):
    z, csr = load_golden(name)
    x = z["x"]
    for T in z["chunk_counts"]:
        T = int(T)
        cvr = oracle.convert(csr, T, "port", fill_missing_tail=True)
        y, _ = oracle.spmv(cvr, csr.n_rows, x)
        assert_y_close(y, csr, x, f"{name} T={T}")
        if bool(z[f"T{T}_ref_kernel_ok"]):
            # same per-lane FMA order as the reference; only the cross-chunk atomics may reorder
            np.testing.assert_allclose(y, z[f"T{T}_y"], rtol=0, atol=1e-13 * np.abs(z[f"T{T}_y"]).max())


def test_missing_tail_corner():
    """8 equal rows in one chunk: the reference never stores final_2 (SENTINEL in the fixture);
    the port leaves it alone by default and fills the intended rows on request."""
    z, csr = load_golden("tiny_equal_rows")
    assert (z["T1_tail"] == oracle.SENTINEL).all()
    c = oracle.convert(csr, 1, "port")
    assert (c["final_2"][:8] == oracle.SENTINEL).all()
    c = oracle.convert(csr, 1, "port", fill_missing_tail=True)
    assert c["final_2"][:8].tolist() == [1, 2, 3, 4, 5, 6, 7, 8]
    x = z["x"]
    y, _ = oracle.spmv(c, csr.n_rows, x)
    assert_y_close(y, csr, x)


INGEST = ["kat12", "pattern_symmetric", "no_trailing_newline", "unsorted_dups"]


@pytest.mark.parametrize("key", INGEST)
def test_port_ingest_matches_reference_readmatrix(key):
    z = np.load(os.path.join(GOLDEN, "ref_ingest.npz"))
    m = oracle.read_mtx(os.path.join(GOLDEN, key + ".mtx"), "port", ref_last_delim=True)
    assert [m.n_rows, m.n_cols, m.nnz] == z[f"{key}_shape"].tolist()
    np.testing.assert_array_equal(m.col, z[f"{key}_col"])
    np.testing.assert_array_equal(m.val, z[f"{key}_val"])
    np.testing.assert_array_equal(m.row_delim, z[f"{key}_rd"])


@pytest.mark.parametrize("key", INGEST)
def test_product_reader_matches_reference_readmatrix(key, native_lib):
    """cvr_read_matrix_market (host C++ in libcvr_b200.so, no GPU needed) vs readMatrix."""
    import cvr_b200
    z = np.load(os.path.join(GOLDEN, "ref_ingest.npz"))
    m = cvr_b200.read_matrix(os.path.join(GOLDEN, key + ".mtx"), ref_last_delim=True)
    assert [m.n_rows, m.n_cols, m.nnz] == z[f"{key}_shape"].tolist()
    np.testing.assert_array_equal(m.col, z[f"{key}_col"])
    np.testing.assert_array_equal(m.val, z[f"{key}_val"])
    np.testing.assert_array_equal(m.row_delim, z[f"{key}_rd"])
    # default mode: the corrected last delimiter, everything else identical
    m2 = cvr_b200.read_matrix(os.path.join(GOLDEN, key + ".mtx"))
    assert m2.row_delim[-1] == m2.nnz
    np.testing.assert_array_equal(m2.col, m.col)


def test_reader_keep_last_line_and_errors(native_lib, tmp_path):
    import cvr_b200
    m = cvr_b200.read_matrix(os.path.join(GOLDEN, "no_trailing_newline.mtx"))
    assert m.nnz_true == 2  # the reference drops the unterminated third entry (spmv.cpp:411)
    m = cvr_b200.read_matrix(os.path.join(GOLDEN, "no_trailing_newline.mtx"), keep_last_line=True)
    assert m.nnz_true == 3
    with pytest.raises(cvr_b200.CvrError):
        cvr_b200.read_matrix(str(tmp_path / "missing.mtx"))
    p = tmp_path / "dense.mtx"
    p.write_text("%%MatrixMarket matrix array real general\n2 2\n1\n2\n3\n4\n")
    with pytest.raises(cvr_b200.CvrError, match="dense"):
        cvr_b200.read_matrix(str(p))


def test_invalid_chunk_counts():
    _, csr = load_golden("fem6")
    with pytest.raises(ValueError):
        oracle.convert(csr, csr.nnz // 16 + 1, "port")


@pytest.mark.parametrize("kind", ["real_general", "pattern_symmetric"])
def test_parallel_ingest_of_a_large_file_matches_the_oracle(kind, native_lib, tmp_path):
    """Files above 1 MB are tokenised by several host threads (cvr_mm_reader.cpp); the result must not
    depend on where the pieces were cut -- compare with the reference's readMatrix when it is available,
    with the port otherwise."""
    import cvr_b200
    rng = np.random.default_rng(17)
    n, m = 50000, 160000
    r = rng.integers(1, n + 1, m)
    c = rng.integers(1, n + 1, m)
    p = str(tmp_path / f"{kind}.mtx")
    with open(p, "w") as f:
        if kind == "real_general":
            v = rng.uniform(-1, 1, m).astype(np.float32)
            f.write("%%MatrixMarket matrix coordinate real general\n% comment\n" + f"{n} {n} {m}\n")
            np.savetxt(f, np.column_stack([r, c, v]), fmt="%d %d %.9g")
        else:
            f.write("%%MatrixMarket matrix coordinate pattern symmetric\n" + f"{n} {n} {m}\n")
            np.savetxt(f, np.column_stack([r, c]), fmt="%d %d")
    assert os.path.getsize(p) > (1 << 20)
    got = cvr_b200.read_matrix(p, ref_last_delim=True)
    want = oracle.read_mtx(p, "ref") if oracle.ref_available() else oracle.read_mtx(p, "port", ref_last_delim=True)
    assert [got.n_rows, got.n_cols, got.nnz] == [want.n_rows, want.n_cols, want.nnz]
    np.testing.assert_array_equal(got.col, want.col)
    np.testing.assert_array_equal(got.val, want.val)
    np.testing.assert_array_equal(got.row_delim, want.row_delim)
