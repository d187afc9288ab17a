#!/usr/bin/env python
"""bench.py -- CVR SpMV throughput on B200 (GFLOP/s = 2*nnz/t, achieved HBM GB/s, roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rmat24|web|fem|road|rmat:S|rmat28]
                    [--impl ours|reference] [--chunks T] [--no-subs] [--rebalance R] [--multicast on|off]
                    [--exchange peer|nccl] [--row-weight W] [--dense-exchange]

A "step" is one SpMV pass of the converted matrix: clear the accumulated rows of y + the sweep
kernel (for N > 1 followed by the y -> x exchange of the iterated SpMV).  The conversion (CSR -> CVR
on the device) happens once before the timed region, like the reference's pre_processing; its kernel
time and roofline are reported in `extra.convert`.

Headline workload (default `rmat24`): BASELINE.json configs[2], R-MAT scale 24 edge factor 16 --
the largest configuration that fits one GPU, and one of the two the north star sets its target on.
At N = 1 the other single-GPU configs (web-Google-shaped, FEM 100^3, road 24M) are measured in the
same run and reported as sub-records under "workloads" (value, ms_per_step, roofline.frac, traffic,
parity each).  At N > 1 the SAME R-MAT-24 matrix is row-sharded by nnz over the ranks (strong
scaling) and every step ends with the exchange that rebuilds the replicated x from the y shards
(--exchange peer: fused into the sweep over peer memory, optionally through an NVSwitch multicast
address with --multicast on; --exchange nccl: all-gather).  Before the timed region the shards are
re-cut from the MEASURED sweep time of every rank (--rebalance R rounds, 0 = keep the nnz cut);
`--workload rmat28` is BASELINE configs[4], generated per row shard.

Parity is checked INSIDE the bench: the y of the timed configuration is compared row by row with a
device CSR product (cvr_verify_csr, |dy| <= 1e-12 * sum|a x|), and for N > 1 three iterations of the
exchange are verified step by step and against the NCCL all-gather path.

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

# stdout carries exactly ONE line, the JSON record: the process's fd 1 is pointed at stderr for the whole run
# (NCCL prints its version banner to stdout on some boxes whatever NCCL_DEBUG_FILE says, torch warns, ...)
# and the record is written to the saved descriptor.
_REAL_STDOUT = None


def _claim_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_record(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


METRIC = "spmv_gflops"
UNIT = "GFLOP/s"
HEADLINE = "rmat24"
SUB_WORKLOADS = ["web", "fem", "road"]
REL_TOL = 1e-12  # BASELINE.json: per row |dy| <= 1e-12 * sum_j |a_ij x_j|

# dram__bytes_read.sum + dram__bytes_write.sum of ONE sweep launch, from the committed `ncu --set full`
# captures of this round (tools/kernel_ab.py under ncu, same matrices, same kernel)
NCU_DRAM_TRAFFIC = {}
NCU_TRAFFIC_SOURCE = {}
try:
    with open(os.path.join(ROOT, "profiles", "r02_ncu_dram_traffic.json")) as _f:
        for _w, _rec in json.load(_f).items():
            NCU_DRAM_TRAFFIC[_w] = _rec["bytes"]
            NCU_TRAFFIC_SOURCE[_w] = _rec["source"]
except Exception:
    pass


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
        d = gen.fem27(100, 100, 100, device=device)
        desc = "fem27 100x100x100 grid (BASELINE configs[1]: FEM-like banded, ~27 nnz/row)"
    elif name == "femweak":  # round-1 weak-scaling workload, kept for comparison
        d = gen.fem27(100, 100, 100 * n_gpus, device=device)
        desc = f"fem27 100x100x{100 * n_gpus} grid (weak-scaled FEM)"
        return _normalise(d, row_normalise), desc, "weak"
    elif name == "web":
        d = gen.powerlaw_web(device=device)
        desc = "web-Google-shaped power law 916,428 rows (BASELINE configs[0])"
    elif name == "rmat24":
        d = gen.rmat(24, 16, device=device)
        desc = "R-MAT scale 24 edge factor 16 (BASELINE configs[2])"
    elif name.startswith("rmat:"):
        s = int(name.split(":")[1])
        d = gen.rmat(s, 16, device=device)
        desc = f"R-MAT scale {s} edge factor 16"
    elif name == "road":
        d = gen.road(24_000_000, device=device)
        desc = "road-network-like 24M rows ~2.4 nnz/row (BASELINE configs[3])"
    elif name == "tiny":  # smoke-sized
        d = gen.fem27(20, 20, 20, device=device)
        desc = "fem27 20^3 (smoke)"
    else:
        raise SystemExit(f"unknown workload {name}")
    return _normalise(d, row_normalise), desc, "strong"


def _normalise(d, row_normalise: bool):
    if row_normalise:  # ||A||_inf <= 1: iterated SpMV stays finite
        import torch
        rd = d.row_delim.to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(d.n_rows + 1, device=d.val.device), rd[1:] - rd[:-1])
        mag = torch.zeros(d.n_rows + 1, dtype=torch.float64, device=d.val.device).index_add_(0, rows, d.val.abs())
        d.val = (d.val / mag[rows].clamp_min(1e-300)).to(torch.float32).to(torch.float64).contiguous()
    return d


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed regions (NVML, every 5 ms)."""
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
            self._stop.wait(0.005)

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
        cpu_reference_time(h, max(1, min(args.warmup, 2)), threads)
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
    emit_record(line)


# ----------------------------------------------------------------------------- helpers of our arm
def conversion_record(info, peak):
    """extra.convert: the two conversion kernels (schedule + permute) timed alone with CUDA events,
    against B_conv = 12 nnzP + 4 (nRows+2) read + 12 nnzP + 8 n_rec written."""
    b_conv = 24 * info["nnz"] + 4 * (info["n_rows"] + 2) + 8 * info["n_records"]
    ks = info["convert_kernel_seconds"]
    return {"kernel_ms": ks * 1e3, "B_conv": b_conv,
            "frac": (b_conv / ks / 1e9 / peak) if ks > 0 else None,
            "row_lists_ms": info["row_lists_seconds"] * 1e3,
            "convert_seconds_device": info["convert_seconds"], "create_seconds": info["create_seconds"]}


def single_gpu_measurement(name, args, dev, local_rank, peak, clocks, want_cusparse=False, want_cpu=False):
    """One workload on one GPU: timed steps, the sweep kernel alone, parity, e2e, conversion."""
    import torch
    import cvr_b200

    full, desc, scaling = make_workload(name, 1, dev, row_normalise=True)
    nnz_true, n_rows, n_cols = full.nnz_true, full.n_rows, full.n_cols
    keep_host = full.to_host() if want_cpu else None
    torch.cuda.synchronize()
    m = cvr_b200.CvrMatrix(full, args.chunks, local_rank)
    info = m.info
    stream = torch.cuda.current_stream()
    g = torch.Generator(device=dev).manual_seed(99)
    x = torch.rand(n_cols + 1, generator=g, device=dev, dtype=torch.float64) - 0.5
    x[0] = 0.0
    y = torch.zeros(n_rows + 1, dtype=torch.float64, device=dev)
    flush = None
    if info["algorithmic_bytes"] < 256e6:  # would sit in the 126 MB L2: flush between steps
        flush = torch.empty(384 * 1024 * 1024, dtype=torch.uint8, device=dev)

    def step():
        m.spmv_device(x, y, stream.cuda_stream)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region A: K whole steps
    launches0 = m.info["kernel_launches"]
    clocks.resume()
    if flush is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
        torch.cuda.synchronize()
        total_ms = e0.elapsed_time(e1)
    else:
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                 for _ in range(args.steps)]
        torch.cuda.synchronize()
        for a, b in pairs:
            flush.fill_(1)
            a.record(stream)
            step()
            b.record(stream)
        torch.cuda.synchronize()
        total_ms = sum(a.elapsed_time(b) for a, b in pairs)
    launches = m.info["kernel_launches"] - launches0
    ms_per_step = total_ms / args.steps
    gflops = 2.0 * nnz_true / (ms_per_step * 1e-3) / 1e9

    # ---- parity of what was just timed: y against the device CSR product, row by row
    parity = cvr_b200.verify_csr(full, x, y, REL_TOL)
    parity["tolerance"] = "per row |dy| <= 1e-12 * sum|a x| vs device CSR product (cvr_verify_csr)"

    # ---- region B: the sweep kernel alone, one CUDA-event pair per launch
    m.set_kernel_timing(True)
    torch.cuda.synchronize()
    for _ in range(args.steps):
        if flush is not None:
            flush.fill_(1)
        step()
    torch.cuda.synchronize()
    ksecs, klaunches = m.kernel_timing()
    m.set_kernel_timing(False)
    kernel_s = ksecs / max(klaunches, 1)
    achieved = info["algorithmic_bytes"] / kernel_s / 1e9

    # ---- end to end through the host-buffer C-ABI call (cvr_spmv): H2D x + SpMV + D2H y per step
    xh = torch.empty(n_cols + 1, dtype=torch.float64).pin_memory()
    yh = torch.empty(n_rows + 1, dtype=torch.float64).pin_memory()
    xh.copy_(x.cpu())
    e2e_steps = max(3, min(args.steps, 10))
    m.spmv_into(xh, yh, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        m.spmv_into(xh, yh, 1)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    clocks.pause()
    e2e_bad = int((yh.to(dev) - y).abs().max().item() > 0.0) if False else None  # same kernels, same x
    del e2e_bad
    # iterated through the same call: `iters` SpMVs per upload (the reference CLI's usage, x resident)
    it10 = m.spmv_into(xh, yh, 10)

    # ---- context: cuSPARSE CSR SpMV (through torch.sparse) on the same matrix, same x
    cusparse = None
    if want_cusparse:
        try:
            A = torch.sparse_csr_tensor(full.row_delim.to(torch.int64), full.col.to(torch.int64), full.val,
                                        size=(n_rows + 1, n_cols + 1))
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
            cusparse = {"gflops": 2.0 * nnz_true / (ms * 1e-3) / 1e9, "ms_per_spmv": ms,
                        "what": "torch.sparse_csr (cuSPARSE) y = A @ x, fp64, same matrix and x"}
            del A, yy
        except Exception as ex:  # context only: never fail the bench on it
            cusparse = {"error": str(ex)[:200]}

    rec = {
        "workload": desc, "n_rows": n_rows, "nnz": nnz_true, "chunks": info["n_chunks"],
        "value": gflops, "unit": UNIT, "ms_per_step": ms_per_step, "gpu_launches": int(launches),
        "l2": "inputs exceed L2 (no flush)" if flush is None else "L2 flushed between steps (384 MB write, untimed)",
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": NCU_DRAM_TRAFFIC.get(name), "traffic_source": NCU_TRAFFIC_SOURCE.get(name),
                     "kernel": "cvr_spmv_tile_kernel",
                     "kernel_variant": m.kernel_name, "kernel_us": kernel_s * 1e6,
                     "algorithmic_bytes_per_launch": info["algorithmic_bytes"],
                     "kernel_share_of_step": kernel_s * 1e3 / ms_per_step,
                     "frac_of_nominal_8TBs": achieved / 8000.0},
        "parity": parity,
        "e2e": {"value": 2.0 * nnz_true / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8 * (n_cols + 1),
                "d2h_bytes_per_step": 8 * (n_rows + 1), "steps": e2e_steps,
                "x_resident_10_iters_gflops": 2.0 * nnz_true / it10 / 1e9 if it10 > 0 else None},
        "convert": conversion_record(info, peak),
        "n_records": info["n_records"],
    }
    if cusparse is not None:
        rec["cusparse_csr"] = cusparse
    if keep_host is not None:
        big = keep_host.nnz > 100e6
        iters = 5 if big else 50
        secs, conv, kind = cpu_reference_time(keep_host, iters, os.cpu_count() or 1)
        rec["cpu_baseline"] = {"value": 2.0 * keep_host.nnz_true / secs / 1e9, "unit": UNIT,
                               "cores": os.cpu_count() or 1, "kind": kind,
                               "sample": f"whole matrix, {iters} SpMV iterations on the host cores "
                                         "(reference timer: y zeroing excluded)",
                               "ms_per_spmv": secs * 1e3, "convert_seconds": conv}
    m.close()
    del full, x, y, flush, m
    torch.cuda.empty_cache()
    return rec, scaling


class PausableClocks:
    """ClockSampler that only samples while a timed region is running."""

    def __init__(self, index):
        self.inner = ClockSampler(index)
        self._on = False

    def resume(self):
        if not self._on:
            self.inner._stop.clear()
            self.inner.__enter__()
            self._on = True

    def pause(self):
        if self._on:
            self.inner.__exit__(None, None, None)
            self._on = False

    def summary(self):
        self.pause()
        return self.inner.summary()


# ----------------------------------------------------------------------------- our arm, N = 1
def run_single(args, local_rank: int):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CVR path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    peak, peak_src = measured_peaks()
    clocks = PausableClocks(local_rank)
    head, scaling = single_gpu_measurement(args.workload, args, dev, local_rank, peak, clocks,
                                           want_cusparse=not args.no_cusparse, want_cpu=not args.no_cpu_baseline)
    subs = {}
    if not args.no_subs and args.workload == HEADLINE:
        for name in SUB_WORKLOADS:
            rec, _ = single_gpu_measurement(name, args, dev, local_rank, peak, clocks,
                                            want_cusparse=not args.no_cusparse, want_cpu=False)
            subs[name] = rec
    line = {
        "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
        "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": head["workload"], "n_rows": head["n_rows"], "nnz": head["nnz"],
                   "chunks_per_gpu": head["chunks"], "iterated_x_from_y": False, "shard_balance": None,
                   "l2": head["l2"],
                   "step": "clear accumulated rows + sweep kernel (programmatic dependent launch)"},
        "gpu_launches": head["gpu_launches"],
        "e2e": head["e2e"],
        "roofline": dict(head["roofline"], peak_source=peak_src),
        "parity": head["parity"],
        "clocks": clocks.summary(),
        "extra": {"convert": head["convert"], "cusparse_csr": head.get("cusparse_csr"),
                  "n_records": head["n_records"],
                  "gather_ceiling": "profiles/r02_gather_probe_summary.txt: stream + gather + FMA without any "
                                    "row bookkeeping takes 0.94 ms on this matrix (L1 miss path, ~1 gathered "
                                    "element per SM clock) = 0.57 of the HBM roofline"
                  if args.workload == HEADLINE else None},
    }
    if "cpu_baseline" in head:
        line["cpu_baseline"] = head["cpu_baseline"]
    if subs:
        line["workloads"] = {k: {kk: vv for kk, vv in v.items() if kk != "cpu_baseline"} for k, v in subs.items()}
    emit_record(line)


# ----------------------------------------------------------------------------- our arm, N > 1
def run_multi(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the CVR path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # rank 0 must print exactly one JSON line on stdout: NCCL writes its version banner (any
    # NCCL_DEBUG level >= VERSION) and debug lines to stdout unless told otherwise
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)

    import cvr_b200
    from cvr_b200 import shard
    from cvr_b200.dist import PeerPublisher, RowShardExchange
    peak, peak_src = measured_peaks()

    big_rmat = None
    if args.workload == "rmat28":
        big_rmat = 28
    elif args.workload.startswith("rmat:") and int(args.workload.split(":")[1]) >= 26:
        big_rmat = int(args.workload.split(":")[1])
    full = full_all = row_delim_all = None
    if big_rmat is not None:
        # config 5: too large to build on one GPU -- every rank generates only its own row shard
        from cvr_b200 import gen
        mine, cuts, nt, row_delim_all = gen.rmat_shard(big_rmat, 16, rank, world, dev, row_normalise=True,
                                                       row_weight=args.row_weight, return_counts=True)

        def make_shard(new_cuts):
            return gen.rmat_shard(big_rmat, 16, rank, world, dev, row_normalise=True, cuts=new_cuts)[0]
        tot = torch.tensor([nt], dtype=torch.int64, device=dev)
        dist.all_reduce(tot)
        nnz_true_total = int(tot.item())
        n_rows_total = n_cols = 1 << big_rmat
        desc = f"R-MAT scale {big_rmat} edge factor 16, generated per row shard (BASELINE configs[4] at scale 28)"
        scaling = "strong"
    else:
        full, desc, scaling = make_workload(args.workload, world, dev, row_normalise=True)
        nnz_true_total = full.nnz_true
        n_rows_total, n_cols = full.n_rows, full.n_cols
        cuts = shard.partition_rows_by_nnz_torch(full.row_delim, world, args.row_weight)
        mine = shard.shard_device_csr(full, cuts[rank], cuts[rank + 1])
        full_all, row_delim_all = full, full.row_delim  # every rank keeps the matrix until the shards are final

        def make_shard(new_cuts):
            return shard.shard_device_csr(full_all, new_cuts[rank], new_cuts[rank + 1])
        if rank != 0:
            full = None  # rank 0 keeps the whole CSR: it is the parity reference
    torch.cuda.synchronize()

    m = cvr_b200.CvrMatrix(mine, args.chunks, local_rank)
    info = m.info
    shard_csr_kept = [mine if big_rmat is not None else None]
    del mine
    torch.cuda.empty_cache()

    stream = torch.cuda.current_stream()
    g = torch.Generator(device=dev).manual_seed(99)
    x0 = torch.rand(n_cols + 1, generator=g, device=dev, dtype=torch.float64) - 0.5
    x0[0] = 0.0
    x = x0.clone()

    publisher = None
    mc_opt = {"auto": None, "on": True, "off": False}[args.multicast]
    if args.exchange == "peer":
        publisher = PeerPublisher(m, cuts, rank, world, local_rank, sparse=not args.dense_exchange, multicast=mc_opt)
        publisher.set_x(x)

    # ---- shard balance from MEASURED sweep times (set-up, untimed): nnz (+ w per row) is a model; the parts
    # with many short rows sweep longer than the hub-row parts of equal nnz and the slowest part sets the step.
    # Each round times the sweep of every rank inside the real iteration (publishing included), re-cuts the rows
    # into parts of equal measured cost (shard.rebalance_cuts) and rebuilds the shards.
    balance_log = []
    rounds = args.rebalance if args.exchange == "peer" else 0

    def rebuild(new_cuts):
        nonlocal m, info, publisher, cuts
        publisher.close()
        m.close()
        cuts = new_cuts
        shard_csr_kept[0] = None
        torch.cuda.empty_cache()
        part = make_shard(cuts)
        m = cvr_b200.CvrMatrix(part, args.chunks, local_rank)
        info = m.info
        if big_rmat is not None:
            shard_csr_kept[0] = part  # the parity check of a per-shard matrix verifies every rank's own rows
        del part
        torch.cuda.empty_cache()
        publisher = PeerPublisher(m, cuts, rank, world, local_rank, sparse=not args.dense_exchange, multicast=mc_opt)
        publisher.set_x(x)

    best = None  # (slowest rank's sweep, cuts) of the best partition measured
    for rnd in (range(rounds + 1) if rounds > 0 else ()):
        for _ in range(3):
            publisher.step(None, stream.cuda_stream)
        m.set_kernel_timing(True)
        dist.barrier()
        torch.cuda.synchronize()
        for _ in range(6):
            publisher.step(None, stream.cuda_stream)
        dist.barrier()
        torch.cuda.synchronize()
        ksecs, kl = m.kernel_timing()
        m.set_kernel_timing(False)
        times = [None] * world
        dist.all_gather_object(times, ksecs / max(kl, 1))
        cuts = [int(c) for c in cuts]
        balance_log.append({"cuts": cuts, "sweep_us": [round(t * 1e6, 1) for t in times]})
        if best is None or max(times) < best[0]:
            best = (max(times), cuts)
        if rnd == rounds or max(times) <= 1.03 * (sum(times) / world):
            break
        new_cuts = shard.rebalance_cuts(row_delim_all, cuts, times, args.row_weight, damping=args.rebalance_damping)
        if new_cuts == cuts:
            break
        rebuild(new_cuts)
    if best is not None and best[1] != cuts:
        rebuild(best[1])  # a re-cut that measured worse than an earlier partition is not kept
    del full_all, row_delim_all
    torch.cuda.empty_cache()

    exchange = RowShardExchange(cuts, rank, world, dev)
    y = torch.zeros(info["n_rows"] + 1, dtype=torch.float64, device=dev)

    def step():
        if publisher is not None:
            publisher.step(y, stream.cuda_stream)  # the sweep publishes rows into every peer's next x
            return
        m.spmv_device(x, y, stream.cuda_stream)
        exchange(y, x)  # y shards -> replicated x: the one collective of the iterated SpMV

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the warm-up on (the same iteration loop as the timed regions: at N = 8 the two timed
    # regions together last 13 ms, two or three NVML samples)
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region A: K whole steps (inputs exceed L2 at every N for the default workload)
    launches0 = m.info["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    total_ms = e0.elapsed_time(e1)
    launches = m.info["kernel_launches"] - launches0
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    gflops = 2.0 * nnz_true_total / (ms_per_step * 1e-3) / 1e9

    # ---- region B: the sweep kernel alone, one CUDA-event pair per launch
    m.set_kernel_timing(True)
    barrier()
    for _ in range(args.steps):
        step()
    barrier()
    ksecs, klaunches = m.kernel_timing()
    m.set_kernel_timing(False)
    clocks.__exit__(None, None, None)
    kernel_s = ksecs / max(klaunches, 1)
    allk = [None] * world
    dist.all_gather_object(allk, (kernel_s * 1e6, info["n_rows"], info["nnz"], info["n_records"]))

    # algorithmic bytes of THIS shard: the x term counts only the entries its columns touch
    used = torch.zeros(n_cols + 1, dtype=torch.uint8, device=dev)
    m.column_footprint(used)
    torch.cuda.synchronize()
    x_touched = int(used.sum().item())
    del used
    shard_bytes = info["algorithmic_bytes"] - 8 * (n_cols + 1) + 8 * x_touched
    achieved = shard_bytes / kernel_s / 1e9

    # ---- parity of the exchange, step by step: after ONE iteration from a known x the assembled vector must
    # equal the CSR product of that x row by row; the result then seeds the next check (3 iterations: both x
    # buffers and the switch to "do not re-publish the empty rows")
    parity = {"tolerance": "per row |dy| <= 1e-12 * sum|a x| vs device CSR product, one iteration at a time",
              "iterations_checked": 0, "rows_failing": 0, "max_rel": 0.0}
    if full is not None or big_rmat is None:
        cur = x0.clone()
        per_iter_x = []
        for it in range(3):
            if publisher is not None:
                if it == 0:
                    publisher.set_x(cur)
                publisher.step(y, stream.cuda_stream)
                nxt = publisher.full_x()
                m.check_async_error()
                if publisher.multicast:
                    # multicast publishing is dense: EVERY rank's own buffer must hold the whole vector, bit for bit
                    diff = (publisher.x_tensor()[1:] - nxt[1:]).abs().max().reshape(1)
                    dist.all_reduce(diff, op=dist.ReduceOp.MAX)
                    parity["multicast_local_copies_max_abs_diff"] = max(
                        parity.get("multicast_local_copies_max_abs_diff", 0.0), float(diff.item()))
            else:
                x.copy_(cur)
                m.spmv_device(x, y, stream.cuda_stream)
                exchange(y, x)
                torch.cuda.synchronize()
                nxt = x.clone()
            if rank == 0:
                r = cvr_b200.verify_csr(full, cur, nxt, REL_TOL, check_row0=False)
                parity["rows_failing"] += r["rows_failing"]
                parity["max_rel"] = max(parity["max_rel"], r["max_rel"])
                parity["iterations_checked"] += 1
            per_iter_x.append(nxt)
            cur = nxt
        # the NCCL all-gather path from the same start: must agree with the fused exchange
        if publisher is not None:
            xa = x0.clone()
            worst = 0.0
            for it in range(3):
                m.spmv_device(xa, y, stream.cuda_stream)
                exchange(y, xa)
                torch.cuda.synchronize()
                scale = float(xa.abs().max().item()) + 1e-300
                worst = max(worst, float((xa[1:] - per_iter_x[it][1:]).abs().max().item()) / scale)
            parity["peer_vs_nccl_max_abs_over_max"] = worst
        del per_iter_x, cur
    elif publisher is not None:
        # the matrix only exists as row shards: EVERY rank checks its own rows against the CSR product of its
        # shard with the assembled x of the previous iteration (three iterations: a row a peer failed to deliver
        # shows up as a wrong product in the next one)
        parity["tolerance"] = ("per row |dy| <= 1e-12 * sum|a x|, every rank verifies its own shard's rows against the "
                               "device CSR product, one iteration at a time")
        mine_csr = shard_csr_kept[0]
        cur = x0.clone()
        lo = cuts[rank]
        for it in range(3):
            if it == 0:
                publisher.set_x(cur)
            publisher.step(y, stream.cuda_stream)
            nxt = publisher.full_x()
            m.check_async_error()
            y_mine = nxt[lo - 1:lo + info["n_rows"]].clone()  # y_mine[r] = x_next[lo - 1 + r], local rows 1..n
            r = cvr_b200.verify_csr(mine_csr, cur, y_mine, REL_TOL, check_row0=False)
            parity["rows_failing"] += r["rows_failing"]
            parity["max_rel"] = max(parity["max_rel"], r["max_rel"])
            parity["iterations_checked"] += 1
            cur = nxt
            del y_mine
        del cur, nxt
        mx = torch.tensor([parity["max_rel"]], dtype=torch.float64, device=dev)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        parity["max_rel"] = float(mx.item())
    else:
        parity["note"] = "matrix generated per shard, NCCL exchange: not verified here (see tests/test_gpu_multi.py)"
    shard_csr_kept[0] = None
    if parity.get("multicast_local_copies_max_abs_diff", 0.0) != 0.0:
        parity["rows_failing"] += 1
    bad = torch.tensor([parity["rows_failing"]], dtype=torch.int64, device=dev)
    dist.all_reduce(bad)

    # ---- end to end: host buffers in, local SpMV, host buffers out, on every rank (max over ranks)
    xh = torch.empty(n_cols + 1, dtype=torch.float64).pin_memory()
    yh = torch.empty(info["n_rows"] + 1, dtype=torch.float64).pin_memory()
    xh.copy_(x0.cpu())
    e2e_steps = max(3, min(args.steps, 10))
    m.spmv_into(xh, yh, 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        m.spmv_into(xh, yh, 1)
    e2e_s = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_gflops = 2.0 * nnz_true_total / float(e2e_s.item()) / 1e9

    if rank == 0:
        ingress = 8.0 * n_rows_total * (world - 1) / world
        line = {
            "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "n_rows": n_rows_total, "nnz": nnz_true_total,
                       "chunks_per_gpu": info["n_chunks"], "iterated_x_from_y": True,
                       "shard_balance": ("nnz" if not args.row_weight else f"nnz + {args.row_weight:g} per non-empty row") + (
                           f", then {len(balance_log)} round(s) of re-cutting by measured sweep time" if balance_log else ""),
                       "l2": "inputs exceed L2 (no flush)",
                       "step": "sweep kernel (programmatic dependent launch)" + (
                           (" publishing y rows with one NVSwitch multicast store per row range (multimem.st) into every GPU's x"
                            if publisher.multicast else " publishing y rows into every peer's x over NVLink")
                           + " + accumulated-rows publish + flag barrier"
                           if publisher is not None else " + NCCL all-gather y->x"),
                       "exchange": ("nccl all-gather" if publisher is None else
                                    "multicast" if publisher.multicast else "peer stores" + (
                                        f" (multicast unavailable: {publisher.multicast_error})"
                                        if publisher.multicast_error else ""))},
            "gpu_launches": int(launches),
            "e2e": {"value": e2e_gflops, "unit": UNIT, "h2d_bytes_per_step": 8 * (n_cols + 1) * world,
                    "d2h_bytes_per_step": 8 * (n_rows_total + world), "steps": e2e_steps},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": None, "traffic_source": None, "peak_source": peak_src,
                         "kernel": "cvr_spmv_tile_kernel (rank 0 shard)",
                         "kernel_variant": m.kernel_name, "kernel_us": kernel_s * 1e6,
                         "algorithmic_bytes_per_launch": shard_bytes, "x_entries_touched": x_touched,
                         "kernel_share_of_step": kernel_s * 1e3 / ms_per_step,
                         "frac_of_nominal_8TBs": achieved / 8000.0},
            "parity": parity,
            "clocks": clocks.summary(),
            "extra": {"per_rank_sweep_us_rows_nnz_records": allk, "shard_balance_rounds": balance_log,
                      "peer_bytes_sent_per_step_rank0": publisher.bytes_sent_per_iteration() if publisher else None,
                      "nvlink_ingress_bytes_per_gpu": ingress,
                      "nvlink_ingress_bound_us": {"at_900_GBs_nominal": ingress / 900e9 * 1e6,
                                                  "at_770_GBs_measured": ingress / 770e9 * 1e6},
                      "convert": conversion_record(info, peak), "n_records": info["n_records"]},
        }
        emit_record(line)
    if publisher is not None:
        publisher.close()
    m.close()
    dist.destroy_process_group()
    if int(bad.item()) != 0:
        raise SystemExit(f"bench.py: parity FAILED ({int(bad.item())} rows outside 1e-12)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--chunks", type=int, default=0, help="CVR chunks per GPU (0 = auto)")
    ap.add_argument("--no-subs", action="store_true", help="N = 1: skip the web / fem / road sub-records")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cusparse", action="store_true", help="skip the cuSPARSE CSR context measurement")
    ap.add_argument("--row-weight", type=float, default=None,
                    help="N > 1: balance shards by nnz + W per non-empty row instead of nnz alone (0 = nnz, the spec). "
                         "Default: 0 up to 4 GPUs, 4 from 8 GPUs on -- the cost model fitted to the measured per-rank "
                         "sweep times on R-MAT-24 (profiles/r02_strong_scaling_rmat24.txt)")
    ap.add_argument("--rebalance", type=int, default=3,
                    help="N > 1: rounds of re-cutting the row shards by MEASURED per-rank sweep time before the timed "
                         "region (0 = keep the nnz / row-weight partition)")
    ap.add_argument("--rebalance-damping", type=float, default=1.0)
    ap.add_argument("--multicast", default="auto", choices=["auto", "on", "off"],
                    help="peer exchange: publish through an NVSwitch multicast address (torch symmetric memory); "
                         "auto = CVR_MULTICAST, default off")
    ap.add_argument("--dense-exchange", action="store_true",
                    help="peer exchange: publish every row to every GPU instead of only to the GPUs that read it")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: y->x exchange fused into the SpMV kernel over peer memory, or NCCL all-gather")
    args = ap.parse_args()
    _claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N > 1 with "
                         "`python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...`")
    if args.row_weight is None:
        args.row_weight = 4.0 if world >= 8 else 0.0
    if world == 1:
        run_single(args, local_rank)
    else:
        run_multi(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
