"""Row partitioning by nnz balance for the multi-GPU path.

Rows are independent, so the matrix is cut into G contiguous row ranges whose nnz are as
equal as the row boundaries allow: the cut for GPU g is the row found by the same kind of
search the reference uses for its per-thread slices (spmv.cpp:631-650), snapped to a row
start so that no row straddles two GPUs.  Each GPU converts its re-based CSR shard to its own
CVR, keeps a replicated x, and writes only its y range; iterated SpMV needs one all-gather
(y shards -> x) per iteration and nothing else.
"""
from __future__ import annotations

import numpy as np

from .csr import CsrMatrix


def partition_rows_by_nnz(row_delim, n_parts: int, row_weight: float = 0.0) -> np.ndarray:
    """Cut points c[0..G] over the 1-based rows: part g owns rows c[g] .. c[g+1]-1, c[0] = 1,
    c[G] = n_rows + 1.  row_delim is the n_rows+2 array of the 1-based CSR.
    row_weight = 0 balances nnz (the north star's rule); row_weight = w balances nnz + w per non-empty
    row, the measured cost model of the SpMV sweep on skewed matrices (a row costs about 3 nonzeros:
    record, y store, clearing -- DESIGN.md section 4)."""
    rd = np.asarray(row_delim).astype(np.int64)
    n_rows = rd.shape[0] - 2
    if row_weight:
        extra = np.zeros_like(rd)
        extra[1:] = np.cumsum((np.diff(rd) > 0) * float(row_weight)).astype(np.int64)
        rd = rd + extra
    nnz = int(rd[-1])
    cuts = np.empty(n_parts + 1, dtype=np.int64)
    cuts[0], cuts[n_parts] = 1, n_rows + 1
    for g in range(1, n_parts):
        target = (nnz * g) // n_parts
        # first row whose start is >= target (a row start, so rows never straddle parts)
        r = int(np.searchsorted(rd[1:n_rows + 2], target, side="left")) + 1
        cuts[g] = min(max(r, cuts[g - 1]), n_rows + 1)
    return cuts


def shard_csr(csr: CsrMatrix, row_begin: int, row_end: int) -> CsrMatrix:
    """Rows [row_begin, row_end) of `csr` as a CSR of its own: local rows 1..n, global columns,
    delimiters re-based, nnz padded to a multiple of 16 with zero-valued copies of the shard's
    last entry (the reference's padding rule, spmv.cpp:474-482, applied per shard).  An empty
    range yields a one-row shard holding 16 explicit zeros at column 1 (conversion needs
    nnz >= 16)."""
    rd = csr.row_delim.astype(np.int64)
    a, b = int(rd[row_begin]), int(rd[row_end])
    n_local = max(row_end - row_begin, 1)
    val, col = csr.val[a:b], csr.col[a:b]
    n = b - a
    if n == 0:
        val = np.zeros(16)
        col = np.ones(16, dtype=np.int32)
        new_rd = np.zeros(n_local + 2, dtype=np.int64)
        new_rd[2:] = 16
        return CsrMatrix(n_local, csr.n_cols, val, col, new_rd, nnz_true=0)
    npad = n if n % 16 == 0 else (n + 16) // 16 * 16
    new_rd = np.zeros(n_local + 2, dtype=np.int64)
    new_rd[1:] = rd[row_begin:row_end + 1] - a
    if npad > n:
        val = np.concatenate([val, np.zeros(npad - n)])
        col = np.concatenate([col, np.full(npad - n, col[-1], dtype=np.int32)])
        # the padding extends the last NON-EMPTY local row, like padding extends the last entry's row
        last = int(np.max(np.flatnonzero(np.diff(new_rd) > 0)))
        new_rd[last + 1:] += npad - n
    return CsrMatrix(n_local, csr.n_cols, val, col, new_rd, nnz_true=n)


def gather_layout(cuts) -> tuple[np.ndarray, np.ndarray]:
    """(offset, count) of every part's y rows inside the global x (x[0] is the phantom)."""
    cuts = np.asarray(cuts, dtype=np.int64)
    return cuts[:-1].copy(), (cuts[1:] - cuts[:-1]).copy()


def partition_rows_by_nnz_torch(row_delim, n_parts: int, row_weight: float = 0.0):
    """partition_rows_by_nnz on a torch tensor (any device); returns a python list of ints."""
    import torch
    rd = row_delim.to(torch.int64)
    n_rows = rd.shape[0] - 2
    if row_weight:
        extra = torch.zeros_like(rd)
        extra[1:] = torch.cumsum(((rd[1:] - rd[:-1]) > 0).to(torch.float64) * float(row_weight), 0).to(torch.int64)
        rd = rd + extra
    nnz = int(rd[-1])
    cuts = [1]
    for g in range(1, n_parts):
        target = torch.tensor([(nnz * g) // n_parts], dtype=torch.int64, device=rd.device)
        r = int(torch.searchsorted(rd[1:n_rows + 2], target, right=False)) + 1
        cuts.append(min(max(r, cuts[-1]), n_rows + 1))
    cuts.append(n_rows + 1)
    return cuts


def rebalance_cuts(row_delim, cuts, seconds, row_weight: float = 0.0, damping: float = 1.0):
    """New cut points from MEASURED per-part sweep times (works on numpy arrays and torch tensors).

    nnz (+ row_weight per non-empty row) is only a model of what a shard costs: on skewed matrices the parts
    with many short rows sweep up to 1.5x longer than the hub-row parts of equal nnz, and the slowest part sets
    the iteration time.  Given the seconds each part of `cuts` took, every row is charged its model weight
    scaled by (seconds of its part / model weight of its part), and the rows are re-cut into parts of equal
    CHARGED cost -- the model keeps its shape inside a part, the measurement fixes the level between parts.
    damping < 1 moves only that fraction of the way (a re-cut changes what the parts exchange, so the times
    it was derived from are not exactly the times it produces).  Returns a python list like
    partition_rows_by_nnz_torch."""
    import torch
    rd = torch.as_tensor(row_delim).to(torch.int64)
    n_rows = rd.shape[0] - 2
    n_parts = len(cuts) - 1
    if len(seconds) != n_parts:
        raise ValueError("one time per part")
    w = (rd[2:] - rd[1:-1]).to(torch.float64)          # model weight of rows 1..n_rows
    if row_weight:
        w = w + (w > 0).to(torch.float64) * float(row_weight)
    cost = torch.empty_like(w)
    mean = sum(float(t) for t in seconds) / n_parts
    for g in range(n_parts):
        a, b = int(cuts[g]) - 1, int(cuts[g + 1]) - 1  # 0-based slice of part g in w
        if b <= a:
            continue
        wg = float(w[a:b].sum())
        level = mean + damping * (float(seconds[g]) - mean)
        cost[a:b] = w[a:b] * (level / wg if wg > 0 else 0.0)
    cum = torch.cumsum(cost, 0)                        # cost of rows 1..r
    total = float(cum[-1]) if n_rows > 0 else 0.0
    out = [1]
    for g in range(1, n_parts):
        target = torch.tensor([total * g / n_parts], dtype=torch.float64, device=cum.device)
        # rows 1..r cost >= target for the first time at r: part g starts at the row after
        r = int(torch.searchsorted(cum, target, right=False)) + 2
        out.append(min(max(r, out[-1]), n_rows + 1))
    out.append(n_rows + 1)
    return out


def shard_device_csr(d, row_begin: int, row_end: int):
    """shard_csr for a DeviceCsr (torch tensors stay on their device)."""
    import torch
    from .matrix import DeviceCsr
    rd = d.row_delim.to(torch.int64)
    a, b = int(rd[row_begin]), int(rd[row_end])
    n_local = max(row_end - row_begin, 1)
    dev = d.val.device
    n = b - a
    if n == 0:
        new_rd = torch.zeros(n_local + 2, dtype=torch.int64, device=dev)
        new_rd[2:] = 16
        return DeviceCsr(n_local, d.n_cols, torch.zeros(16, dtype=torch.float64, device=dev),
                         torch.ones(16, dtype=torch.int32, device=dev), new_rd.to(torch.int32), nnz_true=0)
    val, col = d.val[a:b], d.col[a:b]
    npad = n if n % 16 == 0 else (n + 16) // 16 * 16
    new_rd = torch.zeros(n_local + 2, dtype=torch.int64, device=dev)
    new_rd[1:] = rd[row_begin:row_end + 1] - a
    if npad > n:
        val = torch.cat([val, torch.zeros(npad - n, dtype=torch.float64, device=dev)])
        col = torch.cat([col, col[-1:].expand(npad - n)])
        last = int(torch.nonzero(new_rd[1:] - new_rd[:-1] > 0).max())
        new_rd[last + 1:] += npad - n
    # padding zeros of the parent matrix that fall into this shard are ordinary explicit zeros here
    true = n if row_end <= d.n_rows else n - (d.nnz - d.nnz_true)
    if npad <= 0x7FFFFFFF:
        new_rd = new_rd.to(torch.int32)
    return DeviceCsr(n_local, d.n_cols, val.contiguous(), col.contiguous(), new_rd.contiguous(), nnz_true=true)
