"""Synthetic matrices of the shapes BASELINE.json names, generated with torch
so that the same code runs on the CPU (tests, small) and on the GPU (bench, full size).

All generators return a DeviceCsr in the reference's conventions: 1-based rows/columns,
row 0 an empty phantom row, values float32 U(-1,1) widened to double (the reference parses
"%f" into a float, spmv.cpp:65/:432), duplicates removed, at least 2 entries in the last row
(SURVEY.md 8a-R1 item 7), nnz padded to a multiple of 16 with zero-valued copies of the last
entry (spmv.cpp:474-482).  torch is plumbing here: this is input synthesis, not the hot path.
"""
from __future__ import annotations

import torch

from .matrix import DeviceCsr

MASTER_SEED = 20261017


def _gen(device, seed):
    g = torch.Generator(device=device)
    g.manual_seed(MASTER_SEED + int(seed))
    return g


def _finish(keys: torch.Tensor, n_rows: int, n_cols: int, g, row_normalise: bool = False) -> DeviceCsr:
    """keys = row * (n_cols + 1) + col, 1-based, any order, duplicates allowed."""
    dev = keys.device
    stride = n_cols + 1
    # at least two entries in the last row
    extra = torch.tensor([n_rows * stride + max(1, n_cols - 1), n_rows * stride + n_cols],
                         dtype=torch.int64, device=dev)
    keys = torch.unique(torch.cat([keys, extra]))  # sorted, de-duplicated
    n = int(keys.shape[0])
    rows = keys // stride
    cols = (keys - rows * stride).to(torch.int32)
    del keys
    vals = (torch.rand(n, generator=g, device=dev, dtype=torch.float32) * 2.0 - 1.0).to(torch.float64)
    counts = torch.bincount(rows, minlength=n_rows + 1)[: n_rows + 1]
    if row_normalise:
        # ||A||_inf <= 1 so iterated SpMV stays finite
        mag = torch.zeros(n_rows + 1, dtype=torch.float64, device=dev).index_add_(0, rows, vals.abs())
        vals = (vals / mag[rows].clamp_min(1e-300)).to(torch.float32).to(torch.float64)
    del rows
    npad = n if n % 16 == 0 else (n + 16) // 16 * 16
    if npad > n:  # copies of the last entry, value 0: they extend the last row
        cols = torch.cat([cols, cols[-1:].expand(npad - n)])
        vals = torch.cat([vals, torch.zeros(npad - n, dtype=torch.float64, device=dev)])
        counts[n_rows] += npad - n
    rd = torch.zeros(n_rows + 2, dtype=torch.int64, device=dev)
    rd[1:] = torch.cumsum(counts, 0)
    if npad <= 0x7FFFFFFF:
        rd = rd.to(torch.int32)
    return DeviceCsr(n_rows, n_cols, vals.contiguous(), cols.contiguous(), rd.contiguous(), nnz_true=n)


def fem27(nx: int, ny: int, nz: int, device="cpu", seed: int = 2) -> DeviceCsr:
    """27-point stencil on an nx*ny*nz grid: regular rows, banded (config 2: 100^3)."""
    dev = torch.device(device)
    g = _gen(dev, seed)
    n = nx * ny * nz
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    ix, iy, iz = idx % nx, (idx // nx) % ny, idx // (nx * ny)
    parts = []
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                ok = ((ix + dx >= 0) & (ix + dx < nx) & (iy + dy >= 0) & (iy + dy < ny) &
                      (iz + dz >= 0) & (iz + dz < nz))
                r = idx[ok]
                c = r + dx + dy * nx + dz * nx * ny
                parts.append((r + 1) * (n + 1) + (c + 1))
    return _finish(torch.cat(parts), n, n, g)


def rmat(scale: int, edge_factor: int = 16, device="cpu", seed: int = 3,
         abcd=(0.57, 0.19, 0.19, 0.05), row_normalise: bool = False,
         batch: int = 1 << 26) -> DeviceCsr:
    """R-MAT, 2^scale vertices, edge_factor * 2^scale edges before de-duplication, no vertex
    permutation: the long rows sit at the low ids (configs 3 and 5)."""
    dev = torch.device(device)
    g = _gen(dev, seed)
    n = 1 << scale
    total = edge_factor << scale
    a, b, c, _ = abcd
    keys = []
    done = 0
    while done < total:
        m = min(batch, total - done)
        r = torch.zeros(m, dtype=torch.int64, device=dev)
        col = torch.zeros(m, dtype=torch.int64, device=dev)
        for _level in range(scale):
            u = torch.rand(m, generator=g, device=dev, dtype=torch.float32)
            r = (r << 1) | (u >= a + b).to(torch.int64)
            col = (col << 1) | (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64)
        keys.append(torch.unique((r + 1) * (n + 1) + (col + 1)))
        done += m
    return _finish(torch.cat(keys), n, n, g, row_normalise)


def powerlaw_web(n: int = 916428, nnz_target: int = 5105039, device="cpu", seed: int = 1) -> DeviceCsr:
    """web-Google-shaped: out-degree ~ min(Zipf(2.1), 5000) rescaled to nnz_target with ~19 %
    empty rows; columns 70 % uniform, 30 % Zipf(1.6) through a fixed permutation (config 1)."""
    dev = torch.device(device)
    g = _gen(dev, seed)
    u = torch.rand(n, generator=g, device=dev, dtype=torch.float64).clamp_min(1e-12)
    deg = torch.clamp(torch.floor(u.pow(-1.0 / 1.1)), max=5000.0)  # P(d >= k) ~ k^-1.1
    empty = torch.rand(n, generator=g, device=dev) < 0.19
    deg[empty] = 0.0
    want = deg * (nnz_target * 1.03 / deg.sum())  # a little extra: de-duplication removes some
    deg = torch.floor(want + torch.rand(n, generator=g, device=dev, dtype=torch.float64)).to(torch.int64)
    deg[empty] = 0
    deg.clamp_(max=n)
    rows = torch.repeat_interleave(torch.arange(1, n + 1, device=dev, dtype=torch.int64), deg)
    m = int(rows.shape[0])
    perm = torch.randperm(n, generator=g, device=dev)
    uz = torch.rand(m, generator=g, device=dev, dtype=torch.float64).clamp_min(1e-12)
    zipf = torch.clamp(torch.floor(uz.pow(-1.0 / 0.6)), max=float(n)).to(torch.int64) - 1  # Zipf(1.6) rank
    hot = perm[zipf.clamp_(0, n - 1)]
    uni = torch.randint(0, n, (m,), generator=g, device=dev, dtype=torch.int64)
    pick = torch.rand(m, generator=g, device=dev) < 0.30
    cols = torch.where(pick, hot, uni) + 1
    return _finish(rows * (n + 1) + cols, n, n, g)


def road(n: int = 24_000_000, device="cpu", seed: int = 4, max_offset: int = 1000) -> DeviceCsr:
    """Road-network-like: degree P(1,2,3,4) = (.15,.40,.35,.10) (2.4 per row), columns within
    +-max_offset of the row (config 4)."""
    dev = torch.device(device)
    g = _gen(dev, seed)
    u = torch.rand(n, generator=g, device=dev)
    deg = 1 + (u >= 0.15).to(torch.int64) + (u >= 0.55).to(torch.int64) + (u >= 0.90).to(torch.int64)
    rows = torch.repeat_interleave(torch.arange(1, n + 1, device=dev, dtype=torch.int64), deg)
    m = int(rows.shape[0])
    off = torch.randint(-max_offset, max_offset + 1, (m,), generator=g, device=dev, dtype=torch.int64)
    cols = (rows + off).clamp_(1, n)
    return _finish(rows * (n + 1) + cols, n, n, g)


def random_sparse(n_rows: int, n_cols: int, nnz: int, device="cpu", seed: int = 0,
                  long_rows: int = 0, long_len: int = 0, empty_frac: float = 0.0) -> DeviceCsr:
    """Small test matrices: uniform random entries, optionally a few very long rows and a
    fraction of forced-empty rows."""
    dev = torch.device(device)
    g = _gen(dev, seed)
    r = torch.randint(1, n_rows + 1, (nnz,), generator=g, device=dev, dtype=torch.int64)
    c = torch.randint(1, n_cols + 1, (nnz,), generator=g, device=dev, dtype=torch.int64)
    if empty_frac > 0:
        dead = torch.rand(n_rows + 1, generator=g, device=dev) < empty_frac
        keep = ~dead[r]
        r, c = r[keep], c[keep]
    if long_rows > 0:
        lr = torch.randint(1, n_rows + 1, (long_rows,), generator=g, device=dev, dtype=torch.int64)
        rr = lr.repeat_interleave(long_len)
        cc = torch.randint(1, n_cols + 1, (rr.shape[0],), generator=g, device=dev, dtype=torch.int64)
        r, c = torch.cat([r, rr]), torch.cat([c, cc])
    return _finish(r * (n_cols + 1) + c, n_rows, n_cols, g)


def rmat_shard(scale: int, edge_factor: int, rank: int, world: int, device="cpu", seed: int = 3,
               abcd=(0.57, 0.19, 0.19, 0.05), row_normalise: bool = True, batch: int = 1 << 26,
               row_weight: float = 0.0, cuts=None, return_counts: bool = False):
    """Row shard `rank` of `world` of an R-MAT matrix too large to build on one GPU (config 5:
    scale 28, 4.3e9 edges).  Every rank draws the same edge stream twice from per-batch seeds
    (pass 1: row histogram -> nnz-balanced cuts, the rule of cvr_b200.shard; pass 2: keep only the
    rows it owns), so no edge ever leaves the GPU that generated it and no rank holds more than its
    own shard.  Returns (DeviceCsr with LOCAL rows 1..n_local and GLOBAL columns, cuts list,
    entries in this shard before padding).  With `cuts` given (e.g. from shard.rebalance_cuts) pass 1 is
    skipped and exactly those row ranges are built; return_counts appends the row histogram of pass 1 as a
    delimiter-like int64 array of n + 2 entries (rd[r + 1] - rd[r] = edges drawn for row r, duplicates
    included), the row weights shard.rebalance_cuts works from."""
    dev = torch.device(device)
    n = 1 << scale
    total = edge_factor << scale
    a, b, c, _ = abcd
    n_batches = (total + batch - 1) // batch

    def draw(i):
        g = torch.Generator(device=dev)
        g.manual_seed(MASTER_SEED + 7919 * int(seed) + i)
        m = min(batch, total - i * batch)
        r = torch.zeros(m, dtype=torch.int64, device=dev)
        col = torch.zeros(m, dtype=torch.int64, device=dev)
        for _level in range(scale):
            u = torch.rand(m, generator=g, device=dev, dtype=torch.float32)
            r = (r << 1) | (u >= a + b).to(torch.int64)
            col = (col << 1) | (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64)
        return r + 1, col + 1

    row_delim_like = None
    if cuts is None:
        counts = torch.zeros(n + 2, dtype=torch.int64, device=dev)
        for i in range(n_batches):
            r, _ = draw(i)
            counts += torch.bincount(r, minlength=n + 2)
            del r
        if return_counts:
            row_delim_like = torch.zeros(n + 2, dtype=torch.int64, device=dev)
            row_delim_like[2:] = torch.cumsum(counts[1:n + 1], 0)
        if row_weight:  # balance nnz + row_weight per non-empty row (shard.partition_rows_by_nnz)
            counts = counts + ((counts > 0).to(torch.float64) * float(row_weight)).to(torch.int64)
        ends = torch.cumsum(counts, 0)  # ends[r] = (weighted) edges in rows <= r
        del counts
        weighted_total = int(ends[-1])
        cuts = [1]
        for gidx in range(1, world):
            target = torch.tensor([(weighted_total * gidx) // world], dtype=torch.int64, device=dev)
            # first row whose start (= ends[row-1]) is >= target
            row = int(torch.searchsorted(ends, target, right=False)) + 1
            cuts.append(min(max(row, cuts[-1]), n + 1))
        cuts.append(n + 1)
        del ends
    else:
        cuts = [int(c) for c in cuts]
        if len(cuts) != world + 1 or cuts[0] != 1 or cuts[-1] != n + 1:
            raise ValueError("cuts must run from 1 to n_rows + 1 with one part per rank")
    lo, hi = cuts[rank], cuts[rank + 1]
    kept = []
    for i in range(n_batches):
        r, col = draw(i)
        sel = (r >= lo) & (r < hi)
        kept.append(torch.unique((r[sel] - lo + 1) * (n + 1) + col[sel]))
        del r, col, sel
    n_local = max(hi - lo, 1)
    g = _gen(dev, 1000 + seed * 131 + rank)
    keys = torch.cat(kept) if kept else torch.zeros(0, dtype=torch.int64, device=dev)
    del kept
    d = _finish(keys, n_local, n, g, row_normalise)
    if return_counts:
        return d, cuts, d.nnz_true, row_delim_like
    return d, cuts, d.nnz_true
