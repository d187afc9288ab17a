"""Multi-GPU host logic: one process per GPU, row shards by nnz balance, one collective.

The matrix is cut into contiguous row ranges with (nearly) equal nnz (cvr_b200.shard); rank g
converts its re-based shard to its own CVR, keeps a full replicated x and writes only its y
range.  A single SpMV needs no communication.  Iterated SpMV (x <- y, square A) needs exactly
one exchange per iteration: an all-gather of the y shards into every rank's x.  The reference
has no counterpart (single process, shared memory) -- this is the north star's
multi-GPU extension.

torch.distributed is the plumbing (NCCL over NVLink on GPUs, gloo on CPU for the tests); the
local SpMV is whatever callable the caller passes (CvrMatrix.spmv_device on GPUs).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


class RowShardExchange:
    """y shards -> replicated x.  `cuts` are the 1-based row cut points of shard.partition_rows_by_nnz
    (cuts[g] .. cuts[g+1]-1 belong to rank g); x has n_rows+1 entries, x[0] is the phantom."""

    def __init__(self, cuts, rank: int, world: int, device, group=None):
        self.cuts = [int(c) for c in cuts]
        self.rank, self.world, self.group = rank, world, group
        self.counts = [self.cuts[g + 1] - self.cuts[g] for g in range(world)]
        self.n_local = self.counts[rank]
        self.n_rows = self.cuts[-1] - 1
        self.equal = len(set(self.counts)) == 1
        self.max_count = max(self.counts)
        if not self.equal:
            # all-gather needs equal contributions: pad every shard to the longest one and compact
            # afterwards (the all-gather-v idiom of SURVEY.md 8e); still one collective per iteration
            self.send = torch.zeros(self.max_count, dtype=torch.float64, device=device)
            self.stage = torch.zeros(world * self.max_count, dtype=torch.float64, device=device)

    def bytes_received_per_rank(self) -> int:
        return 8 * (self.n_rows - self.n_local)

    def __call__(self, y_local: torch.Tensor, x: torch.Tensor) -> None:
        """y_local: this rank's y (n_local+1 entries, [0] phantom); x: replicated vector to rebuild."""
        if self.world == 1:
            x[1:1 + self.n_rows].copy_(y_local[1:1 + self.n_local])
            return
        if self.equal:
            dist.all_gather_into_tensor(x[1:1 + self.n_rows], y_local[1:1 + self.n_local], group=self.group)
            return
        self.send[:self.n_local].copy_(y_local[1:1 + self.n_local])
        dist.all_gather_into_tensor(self.stage, self.send, group=self.group)
        for g in range(self.world):
            x[self.cuts[g]:self.cuts[g + 1]].copy_(self.stage[g * self.max_count:g * self.max_count + self.counts[g]])


def iterate(local_spmv, exchange: RowShardExchange, x: torch.Tensor, y_local: torch.Tensor, iters: int) -> None:
    """`iters` iterations of x <- A x on row shards: local SpMV into y_local, then the exchange."""
    for _ in range(iters):
        local_spmv(x, y_local)
        exchange(y_local, x)


class PeerPublisher:
    """Fused exchange for the iterated, row-sharded SpMV: instead of an all-gather after the
    kernel, every finished y row is stored by the SpMV kernel itself into the x vector that each
    GPU reads in the NEXT iteration (own buffer + peer-mapped buffers over NVLink/NVSwitch), the
    few accumulated rows follow from a tiny kernel, and iterations are separated by a flag
    barrier over peer memory.  x is double-buffered: iteration k reads X[k % 2] and publishes
    into X[(k + 1) % 2] everywhere.  One process per GPU; torch.distributed only carries the
    64-byte IPC handles at set-up time.
    """

    def __init__(self, matrix, cuts, rank: int, world: int, device_index: int, group=None, sparse: bool = True,
                 multicast=None):
        """multicast: True / False / None (= CVR_MULTICAST, default off: measured slower than the footprint-sparse
        peer stores on R-MAT-24).
        With NVSwitch multicast the two x buffers are torch symmetric-memory tensors and every finished row is
        published with ONE store to the multicast address instead of one store per reading GPU (the row-heavy
        shards of a skewed matrix otherwise spend their load/store slots on up to 7 copies of every row); the
        exchange is then dense and y is a buffer of its own."""
        import ctypes as C
        from . import _lib
        if world > 8:
            raise ValueError("at most 8 peers (CVR_MAX_PEERS)")
        self.m, self.rank, self.world, self.dev = matrix, rank, world, device_index
        self.cuts = [int(c) for c in cuts]
        self.n_rows = self.cuts[-1] - 1
        self.n_local = self.cuts[rank + 1] - self.cuts[rank]
        self._lib = _lib.load()
        self._C, self._libmod = C, _lib
        nbytes = 8 * (self.n_rows + 1)

        def alloc(n):
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(self._lib.cvr_peer_alloc(device_index, n, C.byref(ptr), handle))
            return ptr.value, handle.raw

        self.multicast = False
        self.multicast_error = None
        if multicast is None:
            multicast = os.environ.get("CVR_MULTICAST", "0") not in ("", "0", "off", "no")
        if world > 1 and multicast:
            self._setup_multicast(group, required=True)
        if self.multicast:
            self._init_multicast_descriptors(alloc, group)
            return

        self._own = []      # (ptr, handle) of X[0], X[1], flags
        for n in (nbytes, nbytes, 4 * 64):
            self._own.append(alloc(n))
        handles = [None] * world
        dist.all_gather_object(handles, [h for _, h in self._own], group=group)
        self._opened = []
        self.ptrs = []      # ptrs[r] = [X0, X1, flags] of rank r as seen from this process
        for r in range(world):
            if r == rank:
                self.ptrs.append([p for p, _ in self._own])
                continue
            row = []
            for h in handles[r]:
                ptr = C.c_void_p()
                _lib.check(self._lib.cvr_peer_open(device_index, h, C.byref(ptr)))
                row.append(ptr.value)
                self._opened.append(ptr.value)
            self.ptrs.append(row)
        # who reads what: every rank marks the columns its shard touches, the marks are exchanged
        # once, and a row is then published only to the ranks whose mark is set (plus its owner)
        self.needs = None
        if sparse and world > 1:
            used = torch.zeros(self.n_rows + 1, dtype=torch.uint8, device=torch.device("cuda", device_index))
            _lib.check(self._lib.cvr_column_footprint(matrix._h, used.data_ptr(), 0))
            torch.cuda.synchronize(device_index)
            lo, hi = self.cuts[rank], self.cuts[rank + 1]
            needs = torch.zeros(max(hi - lo, 1) + 1, dtype=torch.uint8, device=used.device)  # [0] = phantom row
            # to rank q: my marks on q's rows; from rank q: q's marks on MY rows
            parts = [used[self.cuts[q]:self.cuts[q + 1]].contiguous() for q in range(world)]
            recv = [torch.empty(hi - lo, dtype=torch.uint8, device=used.device) for _ in range(world)]
            dist.all_to_all(recv, parts, group=group)
            for q in range(world):
                needs[1:1 + hi - lo] |= (recv[q] << q)
            del used, parts, recv
            needs[1:] &= 0xFF ^ (1 << rank)  # own rows are written in place (y aliases my slice of x)
            dry = os.environ.get("CVR_PUBLISH_DRY_RANKS", "")  # diagnosis only (results are WRONG): these ranks run the
            if dry and str(rank) in dry.split(","):            # publishing sweep but send nothing
                needs.zero_()
            self.needs = needs
            self.chunk_any = torch.zeros(matrix.n_chunks, dtype=torch.uint8, device=needs.device)
            _lib.check(self._lib.cvr_chunk_needs(matrix._h, needs.data_ptr(), self.chunk_any.data_ptr(), 0))
            torch.cuda.synchronize(device_index)
            self.needed_rows = [int(((needs[1:] >> q) & 1).sum()) for q in range(world)]
        self.pub = []
        for parity in (0, 1):
            p = _lib.CvrPublish()
            p.n_dst = world
            p.self = rank  # dst[rank] is this GPU's own buffer (y aliases its slice of it, mode bit 2)
            p.mode = int(os.environ.get("CVR_PUBLISH_MODE", "0"))
            p.mode |= (int(os.environ.get("CVR_PUSH_MIN_ROWS", "0")) // 16 & 0xFF) << 8   # experiment knobs
            p.mode |= (int(os.environ.get("CVR_SCATTER_FACTOR", "0")) & 0xFF) << 16
            p.row_offset = self.cuts[rank] - 1
            p.needs = self.needs.data_ptr() if self.needs is not None else None
            p.chunk_any = self.chunk_any.data_ptr() if self.needs is not None else None
            # the sweep writes y straight into my slice of the next x: no local copy of my own rows
            p.mode |= 4
            p.clear_next = self.ptrs[rank][1 - parity] + 8 * (self.cuts[rank] - 1)
            for r in range(world):
                p.dst[r] = self.ptrs[r][parity]
            self.pub.append(p)
        self._flags = (C.c_void_p * world)(*[self.ptrs[r][2] for r in range(world)])
        self.epoch = 0
        self.k = 0  # iterations done: X[k % 2] holds the current x
        self._reset_k = 0
        self._last_y = None
        dist.barrier(group=group)

    # ---- NVSwitch multicast flavour -------------------------------------------------------------------
    def _setup_multicast(self, group, required: bool) -> None:
        """Both x buffers as symmetric-memory tensors with a multicast mapping; every rank must reach the same
        verdict, so failures are agreed on with an all-reduce before anything depends on them."""
        dev = torch.device("cuda", self.dev)
        ok, err, bufs, hdls = 1, None, [], []
        def agreed(ok_here: int) -> bool:
            flag = torch.tensor([ok_here], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            return int(flag.item()) == 1

        try:  # stage 1, local: allocate (a rank that fails here must not leave the others in the rendezvous)
            import torch.distributed._symmetric_memory as symm
            g = group if group is not None else dist.group.WORLD
            bufs = [symm.empty(self.n_rows + 1, dtype=torch.float64, device=dev) for _ in range(2)]
        except Exception as e:  # noqa: BLE001 -- any failure means "use peer stores"
            ok, err = 0, f"{type(e).__name__}: {e}"
        if agreed(ok):
            try:  # stage 2, collective: exchange the handles, map the peers and the multicast object
                for t in bufs:
                    h = symm.rendezvous(t, g)
                    if not int(h.multicast_ptr):
                        raise RuntimeError("symmetric memory has no multicast mapping on this system")
                    t.zero_()
                    hdls.append(h)
            except Exception as e:  # noqa: BLE001
                ok, err = 0, f"{type(e).__name__}: {e}"
        else:
            ok = 0
        if agreed(ok):
            self.multicast, self._mc_bufs, self._mc_hdls = True, bufs, hdls
            return
        self.multicast_error = err or "a peer could not set up multicast"
        del bufs, hdls
        if required:
            raise RuntimeError(f"NVSwitch multicast requested but unavailable: {self.multicast_error}")

    def _init_multicast_descriptors(self, alloc, group) -> None:
        C, _lib = self._C, self._libmod
        rank, world = self.rank, self.world
        self._own = [alloc(4 * 64)]  # the flag array still travels as a CUDA IPC handle
        handles = [None] * world
        dist.all_gather_object(handles, self._own[0][1], group=group)
        self._opened, flags = [], []
        for r in range(world):
            if r == rank:
                flags.append(self._own[0][0])
                continue
            ptr = C.c_void_p()
            _lib.check(self._lib.cvr_peer_open(self.dev, handles[r], C.byref(ptr)))
            flags.append(ptr.value)
            self._opened.append(ptr.value)
        # ptrs[r] = [X0, X1, flags] of rank r as seen from this process (peer-mapped unicast addresses)
        self.ptrs = [[int(self._mc_hdls[0].buffer_ptrs[r]), int(self._mc_hdls[1].buffer_ptrs[r]), flags[r]]
                     for r in range(world)]
        self.needs = None
        self.y_local = torch.zeros(self.n_local + 1, dtype=torch.float64, device=torch.device("cuda", self.dev))
        self.pub = []
        for parity in (0, 1):
            p = _lib.CvrPublish()
            p.n_dst = world
            p.self = rank
            p.mode = int(os.environ.get("CVR_PUBLISH_MODE", "0")) & ~4  # y is a buffer of its own
            p.row_offset = self.cuts[rank] - 1
            p.needs = None
            p.chunk_any = None
            p.clear_next = None
            for r in range(world):
                p.dst[r] = self.ptrs[r][parity]
            p.multicast = int(self._mc_hdls[parity].multicast_ptr)
            self.pub.append(p)
        self._flags = (C.c_void_p * world)(*flags)
        self.epoch = 0
        self.k = 0
        self._reset_k = 0
        dist.barrier(group=group)

    def x_tensor(self, parity=None):
        """The local x buffer of the given parity (default: the one the next iteration reads) as a
        torch tensor view (n_rows + 1 doubles)."""
        from .matrix import _ptr  # noqa: F401
        parity = self.k % 2 if parity is None else parity
        return _as_tensor(self.ptrs[self.rank][parity], self.n_rows + 1, self.dev)

    def set_x(self, x: torch.Tensor) -> None:
        self._reset_k = self.k
        for p in self.pub:
            p.mode &= ~2
        self.x_tensor().copy_(x)
        # the first sweep after a reset runs its own clearing kernel on its y (= my slice of the other
        # buffer); nothing else to prepare
        torch.cuda.synchronize(self.dev)
        dist.barrier()

    def step(self, y_local, stream: int) -> None:
        """One iteration x <- A x.  y_local is unused (kept for call compatibility with the all-gather
        path): the sweep writes this rank's y directly into its slice of the next x buffer."""
        cur, nxt = self.k % 2, (self.k + 1) % 2
        pub = self.pub[nxt]
        if self.k == self._reset_k + 2:  # both x buffers now hold 0.0 at the never-written rows
            for p in self.pub:
                p.mode |= 2
        self.epoch += 1
        if self.multicast:
            y_ptr = self.y_local.data_ptr()
        else:
            y_ptr = self.ptrs[self.rank][nxt] + 8 * (self.cuts[self.rank] - 1)  # y[r] == x_next[lo - 1 + r]
        self.m.spmv_publish(self.ptrs[self.rank][cur], y_ptr, pub, self._flags, self.rank, self.world,
                            self.epoch, self.k > self._reset_k, stream)
        self.k += 1

    def full_x(self) -> torch.Tensor:
        """The complete current x on every rank.  With footprint-sparse publishing a rank's own
        buffer only holds the entries it reads (plus its own rows), so the full vector is assembled
        from the owners' slices -- one broadcast per rank, outside the iteration loop."""
        torch.cuda.synchronize(self.dev)
        self.m.check_async_error()  # a flag barrier that timed out means stale x: raise instead of returning it
        mine = self.x_tensor()
        out = torch.zeros_like(mine)
        for r in range(self.world):
            lo, hi = self.cuts[r], self.cuts[r + 1]
            piece = mine[lo:hi].clone() if r == self.rank else torch.empty(hi - lo, dtype=torch.float64, device=mine.device)
            dist.broadcast(piece, src=r)
            out[lo:hi] = piece
        return out

    def bytes_sent_per_iteration(self) -> int:
        """Bytes this rank stores into PEER memory per iteration."""
        if self.multicast:
            return 8 * self.n_local  # one copy leaves this GPU; the switch replicates it
        if self.needs is None:
            return 8 * self.n_local * (self.world - 1)
        return 8 * sum(n for q, n in enumerate(self.needed_rows) if q != self.rank)

    def close(self) -> None:
        torch.cuda.synchronize(self.dev)
        dist.barrier()
        for p in self._opened:
            self._lib.cvr_peer_close(self.dev, p)
        self._opened = []
        dist.barrier()
        for p, _ in self._own:
            self._lib.cvr_peer_free(self.dev, p)
        self._own = []
        if self.multicast:
            self._mc_hdls, self._mc_bufs = [], []  # symmetric memory is released with its tensors


class _RawCudaArray:
    """Minimal __cuda_array_interface__ wrapper so torch can view library-owned device memory."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def _as_tensor(ptr: int, n: int, device_index: int) -> torch.Tensor:
    return torch.as_tensor(_RawCudaArray(ptr, n), device=torch.device("cuda", device_index))
