"""GPU suite, needs >= 2 GPUs on the box (skipped otherwise): row-sharded iterated SpMV over NCCL
through the same RowShardExchange the gloo tests cover, local SpMV = the CUDA path."""
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

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, iters, out, fused=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import cvr_b200
    from cvr_b200 import gen, shard
    from cvr_b200.dist import RowShardExchange, iterate

    full = gen.rmat(14, 16, device=dev, seed=71, row_normalise=True)
    cuts = shard.partition_rows_by_nnz_torch(full.row_delim, world)
    mine = shard.shard_device_csr(full, cuts[rank], cuts[rank + 1])
    m = cvr_b200.CvrMatrix(mine, 0, rank)
    ex = RowShardExchange(cuts, rank, world, dev)
    x = torch.from_numpy(np.random.default_rng(9).uniform(-1, 1, full.n_cols + 1)).to(dev)
    x[0] = 0.0
    y = torch.zeros(mine.n_rows + 1, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    if fused:
        from cvr_b200.dist import PeerPublisher
        pp = PeerPublisher(m, cuts, rank, world, rank)
        pp.set_x(x)
        for _ in range(iters):
            pp.step(y, stream)
        torch.cuda.synchronize()
        x = pp.full_x()
        pp.close()
    else:
        iterate(lambda xx, yy: m.spmv_device(xx, yy, stream), ex, x, y, iters)
    torch.cuda.synchronize()
    if rank == 0:
        torch.save(x.cpu(), out)
    m.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("fused", [False, True], ids=["nccl_allgather", "peer_publish"])
def test_two_gpu_iterated_spmv_matches_oracle(tmp_path, native_lib, fused):
    import oracle
    from cvr_b200 import gen
    from helpers import to_oracle_csr
    iters = 3
    out = str(tmp_path / "x.pt")
    mp.spawn(_worker, args=(2, _free_port(), iters, out, fused), nprocs=2, join=True)
    got = torch.load(out).numpy()
    full = gen.rmat(14, 16, device="cuda:0", seed=71, row_normalise=True)
    csr = to_oracle_csr(full)
    x = np.random.default_rng(9).uniform(-1, 1, full.n_cols + 1)
    x[0] = 0.0
    scale = 0.0
    for _ in range(iters):
        y, mag = oracle.csr_spmv(csr, x)
        scale = max(scale, float(mag.max()))
        x = y.copy()
        x[0] = 0.0
    np.testing.assert_allclose(got, x, rtol=0, atol=1e-12 * scale * iters)
