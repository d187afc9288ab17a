"""CPU suite: the multi-GPU host logic (row partition by nnz, shard CSRs, the y -> x all-gather)
with world_size 2 and 3 over gloo.  The local SpMV is the oracle port here; on the GPU box the same
RowShardExchange drives CvrMatrix.spmv_device over NCCL (bench.py --gpus N, test_gpu_multi)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make(kind):
    from cvr_b200 import gen
    if kind == "rmat":
        return gen.rmat(10, 8, seed=61, row_normalise=True)
    if kind == "fem":
        return gen.fem27(8, 8, 12)
    return gen.random_sparse(900, 900, 7000, seed=62, long_rows=2, long_len=700)


def _worker(rank, world, port, kind, iters, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from cvr_b200 import shard
    from cvr_b200.dist import RowShardExchange, iterate
    from helpers import to_oracle_csr

    full = _make(kind).to_host()
    cuts = shard.partition_rows_by_nnz(full.row_delim, world)
    mine = shard.shard_csr(full, int(cuts[rank]), int(cuts[rank + 1]))
    csr = to_oracle_csr(mine)
    cvr = oracle.convert(csr, min(4, csr.nnz // 16), "port", fill_missing_tail=True)

    def local_spmv(x, y):
        yy, _ = oracle.spmv(cvr, csr.n_rows, x.numpy())
        y.copy_(torch.from_numpy(yy))

    ex = RowShardExchange(cuts, rank, world, "cpu")
    x = torch.from_numpy(np.random.default_rng(9).uniform(-1, 1, full.n_cols + 1))
    x[0] = 0.0
    y = torch.zeros(csr.n_rows + 1, dtype=torch.float64)
    iterate(local_spmv, ex, x, y, iters)
    if rank == 0:
        torch.save({"x": x, "cuts": torch.tensor(cuts), "recv": ex.bytes_received_per_rank()}, out)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,kind", [(2, "rmat"), (2, "fem"), (3, "long")])
def test_iterated_spmv_on_row_shards_matches_single_process(world, kind, tmp_path):
    import oracle
    from helpers import to_oracle_csr
    iters = 4
    out = str(tmp_path / "x.pt")
    mp.spawn(_worker, args=(world, _free_port(), kind, iters, out), nprocs=world, join=True)
    got = torch.load(out)
    full = _make(kind).to_host()
    csr = to_oracle_csr(full)
    x = np.random.default_rng(9).uniform(-1, 1, full.n_cols + 1)
    x[0] = 0.0
    scale = 0.0
    for _ in range(iters):
        y, mag = oracle.csr_spmv(csr, x)
        scale = max(scale, float(mag.max()))
        x = y.copy()
        x[0] = 0.0
    np.testing.assert_allclose(got["x"].numpy(), x, rtol=0, atol=1e-12 * max(scale, 1e-300) * iters)
    cuts = got["cuts"].numpy()
    assert cuts[0] == 1 and cuts[-1] == full.n_rows + 1 and np.all(np.diff(cuts) >= 0)
    assert got["recv"] == 8 * (full.n_rows - (cuts[1] - cuts[0]))


def test_partition_balances_nnz_and_never_splits_a_row():
    from cvr_b200 import shard
    full = _make("long").to_host()
    for parts in (2, 4, 8):
        cuts = shard.partition_rows_by_nnz(full.row_delim, parts)
        rd = full.row_delim.astype(np.int64)
        sizes = np.diff(rd[cuts])
        assert sizes.sum() == full.nnz
        longest_row = int(np.diff(rd).max())
        assert sizes.max() - sizes.min() <= 2 * longest_row + 16
        total = 0
        for g in range(parts):
            s = shard.shard_csr(full, int(cuts[g]), int(cuts[g + 1]))
            assert s.nnz % 16 == 0 and s.row_delim[-1] == s.nnz
            total += s.nnz_true
        assert total == full.nnz_true or total == full.nnz


def test_torch_and_numpy_sharding_agree():
    from cvr_b200 import shard
    d = _make("rmat")
    h = d.to_host()
    c1 = shard.partition_rows_by_nnz(h.row_delim, 4)
    c2 = shard.partition_rows_by_nnz_torch(d.row_delim, 4)
    assert c1.tolist() == c2
    for g in range(4):
        a = shard.shard_csr(h, c2[g], c2[g + 1])
        b = shard.shard_device_csr(d, c2[g], c2[g + 1]).to_host()
        np.testing.assert_array_equal(a.col, b.col)
        np.testing.assert_array_equal(a.val, b.val)
        np.testing.assert_array_equal(a.row_delim, b.row_delim)
