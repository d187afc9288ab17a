#!/usr/bin/env python
"""bench.py -- CVR SpMV throughput on B200 (GFLOP/s = 2*nnz/t, achieved HBM GB/s, roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload fem|web|rmat24|road|rmat:S]
                    [--impl ours|reference] [--chunks T]

A "step" is one SpMV pass of the converted matrix: clear y + cvr_spmv_kernel (for N > 1
followed by the y -> x all-gather of the iterated SpMV).  The conversion (CSR -> CVR on the
device) happens once before the timed region, like the reference's pre_processing, and its
time is reported in `extra`.

Workload at N = 1 (default `fem`): BASELINE.json configs[1], the FEM-like banded matrix
(100^3 grid, 27-point stencil, 1,000,000 rows, 26.46 M nnz, 342 MB of algorithmic traffic per
SpMV -- larger than the 126 MB L2, so no flush is needed between steps).  For N > 1 the grid
grows to 100 x 100 x 100N (weak scaling, each rank owns one 100^3 slab = an nnz-balanced row
range) and every step ends with the all-gather that rebuilds the replicated x from the y
shards.  Other workloads are strong-scaled (same matrix, nnz-balanced row shards).

`--impl reference` times the reference's own CPU implementation of this path on the host cores
(the unmodified reference from oracle/_ref when it was built and the CPU has AVX-512F, else the
oracle C port), same matrix, same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "spmv_gflops"
UNIT = "GFLOP/s"

# dram__bytes_read.sum + dram__bytes_write.sum of ONE cvr_spmv_kernel launch, from the committed
# `ncu --set full` captures of this very command (NCU_TRAFFIC_SOURCE)
NCU_DRAM_TRAFFIC = {"fem": 336.50e6 + 4.35e6, "rmat24": 3952.51e6 + 85.03e6,
                    "web": 69.51e6 + 3.22e6, "road": 1094.03e6 + 167.53e6}
NCU_TRAFFIC_SOURCE = {w: "profiles/r01_final_tma_%s_ncu.csv" % w for w in NCU_DRAM_TRAFFIC}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- workloads
def make_workload(name: str, n_gpus: int, device, row_normalise: bool):
    """Returns (DeviceCsr of the WHOLE matrix, description, scaling)."""
    from cvr_b200 import gen
    if name == "fem":
        d = gen.fem27(100, 100, 100 * n_gpus, device=device)
        desc = f"fem27 100x100x{100 * n_gpus} grid (BASELINE configs[1]: FEM-like banded, ~27 nnz/row)"
        scaling = "weak"
    elif name == "web":
        d = gen.powerlaw_web(device=device)
        desc = "web-Google-shaped power law 916,428 rows (BASELINE configs[0])"
        scaling = "strong"
    elif name == "rmat24":
        d = gen.rmat(24, 16, device=device)
        desc = "R-MAT scale 24 edge factor 16 (BASELINE configs[2])"
        scaling = "strong"
    elif name.startswith("rmat:"):
        s = int(name.split(":")[1])
        d = gen.rmat(s, 16, device=device)
        desc = f"R-MAT scale {s} edge factor 16"
        scaling = "strong"
    elif name == "road":
        d = gen.road(24_000_000, device=device)
        desc = "road-network-like 24M rows ~2.4 nnz/row (BASELINE configs[3])"
        scaling = "strong"
    elif name == "tiny":  # smoke-sized
        d = gen.fem27(20, 20, 20 * n_gpus, device=device)
        desc = "fem27 20^3 (smoke)"
        scaling = "weak"
    else:
        raise SystemExit(f"unknown workload {name}")
    if row_normalise:
        import torch
        rd = d.row_delim.to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(d.n_rows + 1, device=d.val.device), rd[1:] - rd[:-1])
        mag = torch.zeros(d.n_rows + 1, dtype=torch.float64, device=d.val.device).index_add_(0, rows, d.val.abs())
        d.val = (d.val / mag[rows].clamp_min(1e-300)).to(torch.float32).to(torch.float64).contiguous()
    return d, desc, scaling


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (NVML)."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) \
                    if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thread:
            self._thread.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------- reference arm
def cpu_reference_time(d_host, iters: int, threads: int):
    """Seconds per SpMV of the reference CPU path (conversion excluded) on `threads` cores."""
    import numpy as np
    import oracle
    csr = oracle.Csr(d_host.n_rows, d_host.n_cols, d_host.val, d_host.col,
                     d_host.row_delim.astype(np.int32), d_host.nnz_true)
    x = np.ones(csr.n_cols + 1)
    if oracle.ref_available():
        kind = "reference"
        os.environ.setdefault("OMP_PROC_BIND", "true")
        t0 = time.perf_counter()
        cvr = oracle.convert(csr, threads, "ref")
        conv = cvr["seconds"] if cvr["seconds"] and cvr["seconds"] > 0 else time.perf_counter() - t0
        _, secs = oracle.spmv(cvr, csr.n_rows, x, "ref", iters=iters)
    else:
        kind = "port"
        t0 = time.perf_counter()
        cvr = oracle.convert(csr, threads, "port", fill_missing_tail=True)
        conv = time.perf_counter() - t0
        oracle.spmv(cvr, csr.n_rows, x, "port")
        t0 = time.perf_counter()
        for _ in range(iters):
            oracle.spmv(cvr, csr.n_rows, x, "port")
        secs = (time.perf_counter() - t0) / iters
    return secs, conv, kind


def run_reference(args, rank: int):
    if rank != 0:
        return
    import torch
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    d, desc, scaling = make_workload(args.workload, args.gpus, dev, row_normalise=True)
    h = d.to_host()
    del d
    threads = os.cpu_count() or 1
    if args.warmup > 0:
        cpu_reference_time(h, max(1, min(args.warmup, 3)), threads)
    secs, conv, kind = cpu_reference_time(h, args.steps, threads)
    gflops = 2.0 * h.nnz_true / secs / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs * 1e3, "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n_rows": h.n_rows, "nnz": h.nnz_true, "host_threads": threads},
        "cpu_baseline": {"value": gflops, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": f"whole matrix, {args.steps} SpMV iterations, reference timer "
                                   "(y zeroing excluded, OpenMP fork/join included)",
                         "convert_seconds": conv},
        "e2e": {"value": gflops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def run_ours(args, rank: int, world: int, local_rank: int):
    import numpy as np
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CVR path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # rank 0 must print exactly one JSON line on stdout: NCCL writes its version banner (any
        # NCCL_DEBUG level >= VERSION) and debug lines to stdout unless told otherwise
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    import cvr_b200
    from cvr_b200 import shard

    iterated = world > 1
    big_rmat = None
    if args.workload == "rmat28":
        big_rmat = 28
    elif args.workload.startswith("rmat:") and int(args.workload.split(":")[1]) >= 26:
        big_rmat = int(args.workload.split(":")[1])
    if big_rmat is not None:
        # config 5: too large to build on one GPU -- every rank generates only its own row shard
        from cvr_b200 import gen
        mine, cuts, nt = gen.rmat_shard(big_rmat, 16, rank, world, dev, row_normalise=True,
                                        row_weight=args.row_weight)
        tot = torch.tensor([nt], dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(tot)
        nnz_true_total = int(tot.item())
        n_rows_total = n_cols = 1 << big_rmat
        desc = f"R-MAT scale {big_rmat} edge factor 16, generated per row shard (BASELINE configs[4] at scale 28)"
        scaling = "strong"
        full = None
    else:
        full, desc, scaling = make_workload(args.workload, world, dev, row_normalise=True)
        nnz_true_total = full.nnz_true
        n_rows_total, n_cols = full.n_rows, full.n_cols
        cuts = (shard.partition_rows_by_nnz_torch(full.row_delim, world, args.row_weight) if world > 1
                else [1, n_rows_total + 1])
        mine = shard.shard_device_csr(full, cuts[rank], cuts[rank + 1]) if world > 1 else full
    keep_host = full.to_host() if (full is not None and rank == 0 and world == 1 and not args.no_cpu_baseline) else None
    keep_csr = None
    if full is not None and world == 1 and not args.no_cusparse and full.nnz < 600_000_000:
        keep_csr = (full.row_delim.to(torch.int64).clone(), full.col.to(torch.int64), full.val.clone())
    del full
    torch.cuda.synchronize()

    m = cvr_b200.CvrMatrix(mine, args.chunks, local_rank)
    info = m.info
    del mine
    torch.cuda.empty_cache()

    from cvr_b200.dist import RowShardExchange
    exchange = RowShardExchange(cuts, rank, world, dev)
    stream = torch.cuda.current_stream()
    g = torch.Generator(device=dev).manual_seed(99)
    x = torch.rand(n_cols + 1, generator=g, device=dev, dtype=torch.float64) - 0.5
    x[0] = 0.0
    y = torch.zeros(info["n_rows"] + 1, dtype=torch.float64, device=dev)

    publisher = None
    if iterated and args.exchange == "peer":
        from cvr_b200.dist import PeerPublisher
        publisher = PeerPublisher(m, cuts, rank, world, local_rank, sparse=not args.dense_exchange)
        publisher.set_x(x)

    def step():
        if publisher is not None:
            publisher.step(y, stream.cuda_stream)  # SpMV kernel publishes rows into every peer's next x
            return
        m.spmv_device(x, y, stream.cuda_stream)
        if iterated:
            exchange(y, x)  # y shards -> replicated x: the one collective of the iterated SpMV

    flush = None
    if info["algorithmic_bytes"] < 256e6:  # would sit in the 126 MB L2: flush between steps
        flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region A: K whole steps
    launches0 = m.info["kernel_launches"]
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    if True:
        if flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(stream)
            for _ in range(args.steps):
                step()
            e1.record(stream)
            barrier()
            total_ms = e0.elapsed_time(e1)
        else:
            pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(args.steps)]
            barrier()
            for a, b in pairs:
                flush.fill_(1)
                a.record(stream)
                step()
                b.record(stream)
            barrier()
            total_ms = sum(a.elapsed_time(b) for a, b in pairs)
    launches = m.info["kernel_launches"] - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    gflops = 2.0 * nnz_true_total / (ms_per_step * 1e-3) / 1e9

    # ---- region B: the SpMV kernel alone, one CUDA-event pair per launch
    m.set_kernel_timing(True)
    barrier()
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        if publisher is not None:
            publisher.step(y, stream.cuda_stream)  # the sweep kernel incl. its peer stores
        else:
            m.spmv_device(x, y, stream.cuda_stream)
    barrier()
    ksecs, klaunches = m.kernel_timing()
    m.set_kernel_timing(False)
    clocks.__exit__(None, None, None)
    kernel_s = ksecs / max(klaunches, 1)
    if world > 1 and os.environ.get("CVR_BENCH_DEBUG"):
        allk = [None] * world
        dist.all_gather_object(allk, (rank, kernel_s * 1e6, info["n_rows"], info["nnz"], info["n_records"]))
        if rank == 0:
            print("per-rank sweep kernel us / rows / nnz / records:", allk, file=sys.stderr)
    peak, peak_src = measured_peaks()
    achieved = info["algorithmic_bytes"] / kernel_s / 1e9

    # ---- end to end through the host-buffer C-ABI call (cvr_spmv): H2D x + SpMV + D2H y per step
    xh = torch.empty(n_cols + 1, dtype=torch.float64).pin_memory()
    yh = torch.empty(info["n_rows"] + 1, dtype=torch.float64).pin_memory()
    xh.copy_(x.cpu())
    e2e_steps = max(3, min(args.steps, 20))
    m.spmv_into(xh, yh, 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        m.spmv_into(xh, yh, 1)
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_gflops = 2.0 * nnz_true_total / float(e2e_s.item()) / 1e9

    # ---- context: cuSPARSE CSR SpMV (through torch.sparse) on the same matrix, same x -- the
    # library baseline a GPU user would reach for; not part of the headline
    cusparse = None
    if world == 1 and keep_csr is not None and not args.no_cusparse:
        try:
            crow, ccol, cval = keep_csr
            A = torch.sparse_csr_tensor(crow, ccol, cval, size=(info["n_rows"] + 1, n_cols + 1))
            for _ in range(3):
                yy = A @ x
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                if flush is not None:
                    flush.fill_(1)
                yy = A @ x
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            if flush is not None:  # subtract the flush, measured alone
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(args.steps):
                    flush.fill_(1)
                f1.record()
                torch.cuda.synchronize()
                ms -= f0.elapsed_time(f1) / args.steps
            cusparse = {"gflops": 2.0 * nnz_true_total / (ms * 1e-3) / 1e9, "ms_per_spmv": ms,
                        "what": "torch.sparse_csr (cuSPARSE) y = A @ x, fp64, same matrix and x"}
            del A, yy
        except Exception as ex:  # context only: never fail the bench on it
            cusparse = {"error": str(ex)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n_rows": n_rows_total, "nnz": nnz_true_total,
                       "chunks_per_gpu": info["n_chunks"], "iterated_x_from_y": iterated,
                       "shard_balance": ("nnz" if not args.row_weight else f"nnz + {args.row_weight:g} per non-empty row")
                       if world > 1 else None,
                       "l2": "inputs exceed L2 (no flush)" if flush is None else "L2 flushed between steps (384 MB write, untimed)",
                       "step": "clear accumulated rows + cvr_spmv_kernel (programmatic dependent launch)" + (
                           (" (publishes y rows into every peer's x over NVLink) + accumulated-rows publish + flag barrier"
                            if publisher is not None else " + NCCL all-gather y->x") if iterated else "")},
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_gflops, "unit": UNIT, "h2d_bytes_per_step": 8 * (n_cols + 1),
                    "d2h_bytes_per_step": 8 * (info["n_rows"] + 1), "steps": e2e_steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": NCU_DRAM_TRAFFIC.get(args.workload) if world == 1 else None,
                         "traffic_source": NCU_TRAFFIC_SOURCE.get(args.workload) if world == 1 else None,
                         "peak_source": peak_src,
                         "kernel": "cvr_spmv_kernel", "kernel_us": kernel_s * 1e6,
                         "algorithmic_bytes_per_launch": info["algorithmic_bytes"],
                         "kernel_share_of_step": kernel_s * 1e3 / ms_per_step,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "clocks": clocks.summary(),
            "extra": {"step_gbs": info["algorithmic_bytes"] * world / (ms_per_step * 1e-3) / 1e9 if scaling == "weak"
                      else None,
                      "peer_bytes_sent_per_step_rank0": publisher.bytes_sent_per_iteration() if publisher else None,
                      "cusparse_csr": cusparse,
                      "convert_seconds_device": info["convert_seconds"],
                      "create_seconds": info["create_seconds"], "n_records": info["n_records"]},
        }
        if keep_host is not None:
            big = keep_host.nnz > 100e6
            secs, conv, kind = cpu_reference_time(keep_host, 10 if big else 50, os.cpu_count() or 1)
            line["cpu_baseline"] = {"value": 2.0 * keep_host.nnz_true / secs / 1e9, "unit": UNIT,
                                    "cores": os.cpu_count() or 1, "kind": kind,
                                    "sample": f"whole matrix, {10 if big else 50} SpMV iterations on the host cores",
                                    "ms_per_spmv": secs * 1e3, "convert_seconds": conv}
        print(json.dumps(line))
    if publisher is not None:
        publisher.close()
    m.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="fem")
    ap.add_argument("--chunks", type=int, default=0, help="CVR chunks per GPU (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cusparse", action="store_true", help="skip the cuSPARSE CSR context measurement")
    ap.add_argument("--row-weight", type=float, default=0.0,
                    help="N > 1: balance shards by nnz + W per non-empty row instead of nnz alone (0 = nnz, the spec)")
    ap.add_argument("--dense-exchange", action="store_true",
                    help="peer exchange: publish every row to every GPU instead of only to the GPUs that read it")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: y->x exchange fused into the SpMV kernel over peer memory, or NCCL all-gather")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N > 1 with "
                         "`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...`")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
