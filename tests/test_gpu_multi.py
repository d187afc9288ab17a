"""GPU suite, needs >= 2 GPUs on the box (skipped otherwise): row-sharded iterated SpMV with ONE PROCESS PER
GPU over torch.distributed (the launch model of bench.py --gpus N): the NCCL all-gather through
RowShardExchange (the path the gloo tests cover on CPU) and the exchange fused into the sweep kernel over
IPC-mapped peer memory (PeerPublisher), sparse and dense.  Every iteration is checked on its own: the assembled
x after iteration k must equal ONE scalar CSR step (oracle) applied to the assembled x after iteration k-1, row
by row within 1e-12 * sum|a x|.  (tests/test_gpu_sharded.py covers the single-process C++ host, also on 1 GPU.)"""
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
ITERS = 4


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import cvr_b200
    from cvr_b200 import gen, shard
    from cvr_b200.dist import PeerPublisher, RowShardExchange

    full = gen.rmat(14, 16, device=dev, seed=71, row_normalise=True)
    cuts = shard.partition_rows_by_nnz_torch(full.row_delim, world)
    mine = shard.shard_device_csr(full, cuts[rank], cuts[rank + 1])
    m = cvr_b200.CvrMatrix(mine, 0, rank)
    ex = RowShardExchange(cuts, rank, world, dev)
    x = torch.from_numpy(np.random.default_rng(9).uniform(-1, 1, full.n_cols + 1)).to(dev)
    x[0] = 0.0
    y = torch.zeros(mine.n_rows + 1, dtype=torch.float64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    iterates = []
    if mode == "nccl":
        for _ in range(ITERS):
            m.spmv_device(x, y, stream)
            ex(y, x)
            torch.cuda.synchronize()
            iterates.append(x.cpu().clone())
    else:
        try:
            pp = PeerPublisher(m, cuts, rank, world, rank, sparse=(mode == "peer_sparse"),
                               multicast=(mode == "multicast"))
        except RuntimeError as e:  # multicast requested, not available on this box: every rank raises alike
            if rank == 0:
                torch.save({"skip": str(e)}, out)
            m.close()
            dist.destroy_process_group()
            return
        pp.set_x(x)
        for _ in range(ITERS):
            pp.step(y, stream)
            full_x = pp.full_x()
            if pp.multicast:  # dense by construction: every GPU's own buffer holds the whole vector, bit for bit
                assert torch.equal(pp.x_tensor()[1:], full_x[1:]), f"rank {rank}: local multicast copy differs"
            iterates.append(full_x.cpu())
            m.check_async_error()
        pp.close()
    if rank == 0:
        torch.save(iterates, out)
    m.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["nccl", "peer_sparse", "peer_dense", "multicast"])
def test_two_gpu_iterated_spmv_step_by_step(tmp_path, native_lib, mode):
    import oracle
    from cvr_b200 import gen
    from helpers import REL_TOL, to_oracle_csr
    out = str(tmp_path / "x.pt")
    mp.spawn(_worker, args=(2, _free_port(), out, mode), nprocs=2, join=True)
    loaded = torch.load(out)
    if isinstance(loaded, dict):
        pytest.skip(loaded["skip"])
    iterates = [t.numpy() for t in loaded]
    full = gen.rmat(14, 16, device="cuda:0", seed=71, row_normalise=True)
    csr = to_oracle_csr(full)
    prev = np.random.default_rng(9).uniform(-1, 1, full.n_cols + 1)
    prev[0] = 0.0
    for k, got in enumerate(iterates, start=1):
        want, mag = oracle.csr_spmv(csr, prev)
        err = np.abs(got - want)
        bad = np.flatnonzero(err[1:] > REL_TOL * mag[1:]) + 1
        assert bad.size == 0, f"{mode} iteration {k}: {bad.size} rows out of tolerance, first {bad[:5]}"
        prev = got.copy()
        prev[0] = 0.0
