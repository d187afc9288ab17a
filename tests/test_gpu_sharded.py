"""GPU suite: the C++ multi-device host behind the C ABI (cvr_create_sharded / cvr_sharded_spmv).

Runs on ONE GPU by listing device 0 twice or three times (several shards on one device): the partition, the
per-shard conversion (bit-exact against the oracle port run on the same sub-CSR), the fused peer-store
exchange (sparse and dense), its flag barrier and the collective fallback are all exercised without a second
GPU; with >= 2 GPUs the same cases run on distinct devices (NVLink peer stores, real NCCL)."""
import numpy as np
import pytest
import torch

import oracle
from helpers import REL_TOL, assert_structure_equal, to_oracle_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cvr(native_lib):
    import cvr_b200
    return cvr_b200


def device_lists():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    out = [[0, 0], [0, 0, 0]]
    if n >= 2:
        out += [[0, 1]]
    if n >= 4:
        out += [[0, 1, 2, 3]]
    return out


def matrices(gen):
    return {
        "rmat": gen.rmat(13, 16, seed=81, row_normalise=True),
        "fem": gen.fem27(16, 16, 24),
        "web": gen.powerlaw_web(40000, 220000, seed=82),
        "long": gen.random_sparse(3000, 3000, 9000, seed=83, long_rows=5, long_len=2000),
    }


def step_check(csr, x_prev, x_next, what):
    """one iteration x_next = A x_prev, row by row within 1e-12 * sum|a x|"""
    yc, mag = oracle.csr_spmv(csr, x_prev)
    err = np.abs(x_next - yc)
    bad = np.flatnonzero(err[1:] > REL_TOL * mag[1:]) + 1
    assert bad.size == 0, f"{what}: {bad.size} rows out of tolerance, first {bad[:5]}, err {err[bad[:5]]}, mag {mag[bad[:5]]}"


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_single_spmv_and_shard_structure(cvr, devices):
    from cvr_b200 import gen, shard
    for name, d in matrices(gen).items():
        h = d.to_host()
        csr = to_oracle_csr(h)
        x = np.random.default_rng(5).uniform(-1, 1, csr.n_cols + 1)
        x[0] = 0.0
        with cvr.ShardedCvr(h, devices, n_chunks=37) as s:
            info = s.info
            cuts = shard.partition_rows_by_nnz(h.row_delim, len(devices))
            assert info["row_begin"] == [int(c) for c in cuts[:-1]] and info["row_end"] == [int(c) for c in cuts[1:]]
            # every shard converts to exactly what the reference conversion gives on that sub-CSR
            for g in range(len(devices)):
                sub = shard.shard_csr(h, int(cuts[g]), int(cuts[g + 1]))
                want = oracle.convert(to_oracle_csr(sub), 37, "port", fill_missing_tail=True)
                assert_structure_equal(s.part_export(g), want, f"{name} shard {g} of {devices}")
            y, secs = s.spmv(x, iters=2)
            assert secs > 0
            step_check(csr, x, y, f"{name} {devices} single")
            assert y[0] == 0.0


@pytest.mark.parametrize("exchange,dense", [("peer", False), ("peer", True), ("nccl", False)],
                         ids=["peer_sparse", "peer_dense", "collective"])
@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_iterated_spmv_step_by_step(cvr, devices, exchange, dense):
    """x <- A x: the k-iteration result must equal ONE oracle step applied to the (k-1)-iteration result, for
    k = 1..4 (covers both x buffers, the first-iteration clearing and the switch that stops re-publishing
    empty rows), and all exchanges must agree."""
    from cvr_b200 import gen
    for name, d in matrices(gen).items():
        h = d.to_host()
        csr = to_oracle_csr(h)
        x0 = np.random.default_rng(6).uniform(-1, 1, csr.n_cols + 1)
        x0[0] = 0.0
        with cvr.ShardedCvr(h, devices, n_chunks=0, exchange=exchange, dense=dense) as s:
            prev = x0
            for k in range(1, 5):
                xk, _ = s.spmv(x0, iters=k, feed_y_to_x=True)
                step_check(csr, prev, xk, f"{name} {devices} {exchange} dense={dense} iteration {k}")
                prev = xk
            # and the plain loop afterwards still works on the same handle
            y, _ = s.spmv(x0, iters=1)
            step_check(csr, x0, y, f"{name} {devices} single after iterated")


def test_sharded_accepts_the_reference_last_delimiter_and_rejects_garbage(cvr):
    from cvr_b200 import gen
    h = gen.random_sparse(2000, 2000, 16000, seed=84, empty_frac=0.2).to_host()
    csr = to_oracle_csr(h)
    quirk = cvr.CsrMatrix(h.n_rows, h.n_cols, h.val, h.col, np.where(h.row_delim == h.nnz, h.nnz - 1, h.row_delim),
                          h.nnz_true)
    x = np.random.default_rng(7).uniform(-1, 1, csr.n_cols + 1)
    x[0] = 0.0
    with cvr.ShardedCvr(quirk, [0, 0]) as s:
        y, _ = s.spmv(x)
        step_check(csr, x, y, "quirk sharded")
    with pytest.raises(cvr.CvrError):
        cvr.ShardedCvr(h, [0, 99])
    with pytest.raises(cvr.CvrError):
        cvr.ShardedCvr(h, [])
    rect = gen.random_sparse(500, 700, 4000, seed=85).to_host()
    with cvr.ShardedCvr(rect, [0, 0]) as s:
        with pytest.raises(cvr.CvrError):
            s.spmv(np.ones(701), iters=2, feed_y_to_x=True)
        y, _ = s.spmv(np.ones(701))
        step_check(to_oracle_csr(rect), np.ones(701), y, "rectangular sharded")


@pytest.mark.parametrize("devices", device_lists(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_sharded_rebalance_from_measured_sweep_times(cvr, devices, monkeypatch):
    """CVR_SHARD_REBALANCE: cvr_create_sharded times every part's sweep inside the real iteration and re-cuts the rows
    (the C++ twin of shard.rebalance_cuts).  Whatever the measured times are -- shards sharing one GPU time each other's
    kernels -- the result must be a valid partition (monotone cuts over all rows, every row owned once), every shard
    must convert to what the reference gives on its sub-CSR, and the iterated SpMV must stay right step by step."""
    from cvr_b200 import gen, shard
    monkeypatch.setenv("CVR_SHARD_REBALANCE", "2")
    d = gen.rmat(13, 16, seed=86, row_normalise=True)
    h = d.to_host()
    csr = to_oracle_csr(h)
    x0 = np.random.default_rng(8).uniform(-1, 1, csr.n_cols + 1)
    x0[0] = 0.0
    with cvr.ShardedCvr(h, devices, n_chunks=29) as s:
        info = s.info
        lo, hi = info["row_begin"], info["row_end"]
        assert lo[0] == 1 and hi[-1] == h.n_rows + 1 and all(a <= b for a, b in zip(lo, hi))
        assert all(hi[g] == lo[g + 1] for g in range(len(devices) - 1))
        for g in range(len(devices)):
            sub = shard.shard_csr(h, lo[g], hi[g])
            want = oracle.convert(to_oracle_csr(sub), 29, "port", fill_missing_tail=True)
            assert_structure_equal(s.part_export(g), want, f"re-cut shard {g} of {devices}")
        prev = x0
        for k in range(1, 4):
            xk, _ = s.spmv(x0, iters=k, feed_y_to_x=True)
            step_check(csr, prev, xk, f"re-cut {devices} iteration {k}")
            prev = xk
