"""CPU suite: property tests on tiny random matrices -- the shapes where lane scheduling has its corner
cases (fewer than 8 rows, empty rows at chunk starts, rows longer than a chunk, every legal chunk count).
Port vs the unmodified reference where it is available, and the port's own invariants everywhere."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle
from helpers import assert_y_close


def build_csr(row_lengths, seed):
    rng = np.random.default_rng(seed)
    n_rows = len(row_lengths)
    n_cols = max(8, n_rows)
    rows, cols = [], []
    for r, k in enumerate(row_lengths, start=1):
        k = min(k, n_cols)
        c = np.sort(rng.choice(np.arange(1, n_cols + 1), size=k, replace=False))
        rows += [r] * k
        cols += c.tolist()
    if not rows:
        rows, cols = [n_rows, n_rows], [1, 2]
    n = len(rows)
    vals = rng.integers(-50, 50, n).astype(np.float64) + 0.5
    pad = (-n) % 16
    rows += [rows[-1]] * pad
    cols += [cols[-1]] * pad
    vals = np.concatenate([vals, np.zeros(pad)])
    order = np.lexsort((np.arange(len(rows)), cols, rows))  # (row, col), stable
    rows = np.asarray(rows)[order]
    cols = np.asarray(cols)[order]
    vals = vals[order]
    rd = np.zeros(n_rows + 2, dtype=np.int64)
    rd[1:] = np.cumsum(np.bincount(rows, minlength=n_rows + 1)[: n_rows + 1])
    return oracle.Csr(n_rows, n_cols, vals, cols, rd, nnz_file=n)


lengths = st.lists(st.one_of(st.integers(0, 3), st.integers(0, 40), st.just(0)), min_size=1, max_size=40)


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(row_lengths=lengths, seed=st.integers(0, 10_000), data=st.data())
def test_port_invariants_and_reference_agreement(row_lengths, seed, data):
    if sum(row_lengths) == 0:
        row_lengths = row_lengths + [3]
    if row_lengths[-1] < 2:  # the reference mis-sizes a last row of one entry
        row_lengths = row_lengths[:-1] + [2]
    csr = build_csr(row_lengths, seed)
    T = data.draw(st.integers(1, csr.nnz // 16))
    cvr = oracle.convert(csr, T, "port", fill_missing_tail=True)
    # invariants: the CVR arrays are a permutation of the CSR arrays inside every chunk
    nr = cvr["nnz_rows"].reshape(T, 4)
    assert nr[0, 0] == 0 and nr[-1, 1] == csr.nnz and np.all(nr[1:, 0] == nr[:-1, 1])
    for s, e, r0, r1 in nr:
        assert (e - s) % 16 == 0 and r0 <= r1
        a = np.lexsort((csr.val[s:e], csr.col[s:e]))
        b = np.lexsort((cvr["vals"][s:e], cvr["cols"][s:e]))
        np.testing.assert_array_equal(csr.col[s:e][a], cvr["cols"][s:e][b])
        np.testing.assert_array_equal(csr.val[s:e][a], cvr["vals"][s:e][b])
    x = np.random.default_rng(seed + 1).uniform(-1, 1, csr.n_cols + 1)
    y, _ = oracle.spmv(cvr, csr.n_rows, x)
    assert_y_close(y, csr, x, f"lengths={row_lengths} T={T}")
    if oracle.ref_available():
        ref = oracle.convert(csr, T, "ref")
        assert oracle.structure_equal(oracle.convert(csr, T, "port"), ref, compare_tail=False) == [], \
            f"lengths={row_lengths} T={T}"
        for t in range(T):  # tails agree wherever the reference wrote one
            if ref["final_2"][16 * t] != oracle.SENTINEL:
                np.testing.assert_array_equal(ref["final_2"][16 * t:16 * t + 8], cvr["final_2"][16 * t:16 * t + 8])
