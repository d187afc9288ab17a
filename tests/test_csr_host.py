"""CPU suite: host-side CSR helpers (the conventions of readMatrix) and the writer/reader round trip."""
import numpy as np
import pytest

import oracle


def test_from_coo_follows_readmatrix_conventions(native_lib):
    import cvr_b200
    rows = np.array([3, 1, 3, 2, 1])
    cols = np.array([2, 4, 1, 2, 1])
    vals = np.array([0.1, 0.2, 0.3, 0.4, 0.5])
    m = cvr_b200.CsrMatrix.from_coo(rows, cols, vals, 3, 4)
    assert m.nnz == 16 and m.nnz_true == 5                      # padded to a multiple of 16 (spmv.cpp:457)
    assert m.row_delim.tolist() == [0, 0, 13, 14, 16]           # pads are copies of the LAST given entry (1,1)
    assert m.col[:13].tolist() == [1] * 12 + [4]                # (1,1) x12 (1 real + 11 zero pads), then (1,4)
    assert np.count_nonzero(m.val) == 5
    assert m.val.dtype == np.float64 and np.all(m.val == m.val.astype(np.float32))  # float32-rounded (spmv.cpp:65)
    with pytest.raises(ValueError):
        cvr_b200.CsrMatrix(3, 4, m.val[:15], m.col[:15], m.row_delim)


def test_write_mtx_then_read_matrix_round_trip(native_lib, tmp_path):
    import cvr_b200
    from cvr_b200 import gen
    d = gen.random_sparse(400, 300, 2500, seed=33, empty_frac=0.2).to_host()
    rows = d.row_of_entry()
    p = str(tmp_path / "rt.mtx")
    cvr_b200.write_mtx(p, d.n_rows, d.n_cols, rows, d.col, d.val)   # padding zeros included as explicit entries
    back = cvr_b200.read_matrix(p)
    assert [back.n_rows, back.n_cols, back.nnz] == [d.n_rows, d.n_cols, d.nnz]
    np.testing.assert_array_equal(back.col, d.col)
    np.testing.assert_array_equal(back.val, d.val)
    np.testing.assert_array_equal(back.row_delim, d.row_delim)
    # and the oracle's reader agrees with the product's on the same file
    port = oracle.read_mtx(p, "port")
    np.testing.assert_array_equal(port.col, back.col)
    np.testing.assert_array_equal(port.row_delim, back.row_delim)


def test_rebalance_cuts_from_measured_times():
    """shard.rebalance_cuts: equal times keep the nnz partition, a slow part hands rows to its neighbours,
    cut points stay monotone with fixed ends, and full damping to zero is the identity."""
    from cvr_b200 import gen, shard
    full = gen.rmat(12, 8, seed=3, row_normalise=True).to_host()
    rd = np.asarray(full.row_delim).astype(np.int64)
    for parts in (2, 4, 8):
        cuts = [int(c) for c in shard.partition_rows_by_nnz(full.row_delim, parts)]
        same = shard.rebalance_cuts(full.row_delim, cuts, [1.0] * parts)
        assert all(abs(a - b) <= 1 for a, b in zip(same, cuts))  # (integer against floating-point targets)
        slow_last = [1.0] * (parts - 1) + [2.0]
        c2 = shard.rebalance_cuts(full.row_delim, cuts, slow_last)
        assert c2[0] == 1 and c2[-1] == full.n_rows + 1 and all(a <= b for a, b in zip(c2, c2[1:]))
        nnz_before = rd[cuts[-1]] - rd[cuts[-2]]
        nnz_after = rd[c2[-1]] - rd[c2[-2]]
        assert nnz_after < nnz_before  # the slow part shrinks ...
        assert rd[c2[1]] - rd[c2[0]] > rd[cuts[1]] - rd[cuts[0]]  # ... and the others grow
        # charged cost is equalised: (seconds per model weight of the OLD part) x weight of the new part
        undamped = shard.rebalance_cuts(full.row_delim, cuts, slow_last, damping=0.0)
        assert all(abs(a - b) <= 1 for a, b in zip(undamped, cuts))
    # with a row weight the model weight of a part is nnz + w per non-empty row
    cuts = [int(c) for c in shard.partition_rows_by_nnz(full.row_delim, 4, row_weight=3.0)]
    same = shard.rebalance_cuts(full.row_delim, cuts, [1.0] * 4, row_weight=3.0)
    assert all(abs(a - b) <= 1 for a, b in zip(same, cuts))


def test_rmat_shard_with_given_cuts_and_row_histogram():
    """gen.rmat_shard (config 5's per-shard generator): the shards of its own nnz-balanced cuts tile the matrix,
    passing those cuts back reproduces them exactly, other cuts (e.g. from shard.rebalance_cuts) give shards of
    exactly those row ranges with the same total, and the row histogram has the delimiter shape."""
    import torch
    from cvr_b200 import gen, shard
    scale, world = 10, 3
    first = [gen.rmat_shard(scale, 8, r, world, "cpu", seed=5, return_counts=True) for r in range(world)]
    cuts = first[0][1]
    assert all(f[1] == cuts for f in first) and cuts[0] == 1 and cuts[-1] == (1 << scale) + 1
    rd = first[0][3]
    assert rd.shape[0] == (1 << scale) + 2 and int(rd[0]) == 0 and int(rd[1]) == 0 and int(rd[-1]) == 8 << scale
    total = sum(f[2] for f in first)
    again = [gen.rmat_shard(scale, 8, r, world, "cpu", seed=5, cuts=cuts) for r in range(world)]
    for a, b in zip(first, again):
        assert torch.equal(a[0].col, b[0].col) and torch.equal(a[0].val, b[0].val) and torch.equal(a[0].row_delim, b[0].row_delim)
    moved = shard.rebalance_cuts(rd, cuts, [1.0, 1.0, 2.0])
    assert moved != cuts
    other = [gen.rmat_shard(scale, 8, r, world, "cpu", seed=5, cuts=moved) for r in range(world)]
    assert sum(o[2] for o in other) == total
    assert [o[0].n_rows for o in other] == [max(moved[r + 1] - moved[r], 1) for r in range(world)]
    with pytest.raises(ValueError):
        gen.rmat_shard(scale, 8, 0, world, "cpu", seed=5, cuts=[1, 5, 9])
