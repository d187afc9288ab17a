#!/usr/bin/env python
"""GPU comparison harness -- the B200 stand-in for the reference's solutions_for_comparison/run_comparison.sh
(:17-45: run every solution on the same matrix, grep the three result lines) and run_locality.sh (:39-60: cache
hit / miss counters per solution).

    python tools/compare.py --workload web|fem|road|rmat24|rmat:S  [--iters N]
    python tools/compare.py path/to/matrix.mtx                     [--iters N]
    python tools/compare.py --workload web --locality              (re-runs every solution under ncu)

Solutions (same matrix, same x, all fp64, all checked against the device CSR self-check):
    CVR-B200      this repository: CSR -> CVR on the device, cvr_spmv_tile_kernel
    cuSPARSE-CSR  torch.sparse_csr (cuSPARSE SpMV), the library a GPU user would reach for
    CSR-balanced  a CSR5-style nnz-balanced CSR kernel (tools/compare/csr_kernels.cu)
    CSR-vector    one warp per row
For each it prints the reference's three lines (spmv.cpp:1009, :1662, :1664 with the solution's name) and, with
--locality, the L2 sector hits / misses and DRAM bytes of one SpMV from ncu.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
CMP_LIB = os.path.join(HERE, "compare", "libcvr_compare.so")
SOLUTIONS = ["CVR-B200", "cuSPARSE-CSR", "CSR-balanced", "CSR-vector"]
KERNEL_REGEX = {"CVR-B200": "cvr_spmv_tile_kernel", "cuSPARSE-CSR": "csrmv|spmv|cusparse", "CSR-balanced": "csr_balanced_kernel",
                "CSR-vector": "csr_vector_kernel"}
NCU_METRICS = ("lts__t_sectors_lookup_hit.sum,lts__t_sectors_lookup_miss.sum,lts__t_sector_hit_rate.pct,"
               "dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum")


def build_compare_lib():
    src = os.path.join(HERE, "compare", "csr_kernels.cu")
    if os.path.exists(CMP_LIB) and os.path.getmtime(CMP_LIB) >= os.path.getmtime(src):
        return
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-shared", "-o", CMP_LIB, src])


def load_matrix(args, dev):
    import torch
    import cvr_b200
    if args.matrix:
        h = cvr_b200.read_matrix(args.matrix)
        d = cvr_b200.DeviceCsr(h.n_rows, h.n_cols, torch.from_numpy(h.val).to(dev), torch.from_numpy(h.col).to(dev),
                               torch.from_numpy(h.row_delim).to(dev), h.nnz_true)
        return d, args.matrix
    from bench import make_workload
    d, desc, _ = make_workload(args.workload, 1, dev, row_normalise=False)
    return d, args.workload


def run_solutions(args):
    import torch
    import cvr_b200
    dev = torch.device("cuda", 0)
    build_compare_lib()
    lib = C.CDLL(CMP_LIB)
    lib.csr_balanced_spmv.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.csr_vector_spmv.argtypes = [C.c_void_p] * 3 + [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    d, label = load_matrix(args, dev)
    n, nnz_true = d.n_rows, d.nnz_true
    if "int64" in str(d.row_delim.dtype):
        raise SystemExit("compare.py: the comparators take 32-bit row delimiters (nnz < 2^31)")
    x = torch.ones(d.n_cols + 1, dtype=torch.float64, device=dev)  # x = 1.0 like the reference (spmv.cpp:556-563)
    x[0] = 0.0
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev) if 12 * d.nnz < 256e6 and not args.no_flush else None
    results = {}
    only = args.only.split(",") if args.only else SOLUTIONS
    for name in SOLUTIONS:
        if name not in only:
            continue
        y = torch.zeros(n + 1, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        keep = None
        if name == "CVR-B200":
            keep = cvr_b200.CvrMatrix(d, args.chunks, 0)
            pre = keep.info["create_seconds"]

            def spmv():
                keep.spmv_device(x, y, stream)
        elif name == "cuSPARSE-CSR":
            A = torch.sparse_csr_tensor(d.row_delim.to(torch.int64), d.col.to(torch.int64), d.val, size=(n + 1, d.n_cols + 1))
            torch.cuda.synchronize()
            pre = time.perf_counter() - t0

            def spmv():
                y.copy_(A @ x)
        elif name == "CSR-balanced":
            pre = 0.0

            def spmv():
                lib.csr_balanced_spmv(d.row_delim.data_ptr(), d.val.data_ptr(), d.col.data_ptr(), n, d.nnz, x.data_ptr(),
                                      y.data_ptr(), stream)
        else:
            pre = 0.0

            def spmv():
                lib.csr_vector_spmv(d.row_delim.data_ptr(), d.val.data_ptr(), d.col.data_ptr(), n, x.data_ptr(),
                                    y.data_ptr(), stream)
        for _ in range(3):
            spmv()
        torch.cuda.synchronize()
        check = cvr_b200.verify_csr(d, x, y, 1e-12)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(args.iters):
            if flush is not None:
                flush.fill_(1)
            e0.record()
            spmv()
            e1.record()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        secs = total / args.iters * 1e-3
        gflops = 2.0 * nnz_true / secs / 1e9
        tag = f"[file: {label}] [solution: {name}]"
        print(f"The Pre-processing(CSR->{name})   Time of {name}   is {pre:.6g} seconds.   {tag}")
        print(f"The SpMV Execution Time of {name}    is {secs:.6g} seconds.   {tag}")
        print(f"         The Throughput of {name}    is {gflops:.6g} GFlops.    {tag} [2*nnz/t]")
        print("     Very Good! Your result is correct  " if check["rows_failing"] == 0
              else f"Warning: {check['rows_failing']} rows out of 1e-12 * sum|a x| (first {check['first_bad_row']})")
        results[name] = {"preprocess_seconds": pre, "spmv_seconds": secs, "gflops": gflops,
                         "rows_failing": check["rows_failing"], "max_rel": check["max_rel"]}
        if keep is not None:
            keep.close()
        del y
    print(json.dumps({"matrix": label, "n_rows": n, "nnz": nnz_true, "iters": args.iters,
                      "l2": "flushed between SpMVs" if flush is not None else "not flushed", "solutions": results}))
    return results


def run_locality(args):
    """run_locality.sh: one profiled SpMV per solution, cache counters side by side."""
    rows = []
    for name in SOLUTIONS:
        cmd = ["ncu", "--metrics", NCU_METRICS, "--clock-control", "none", "--csv", "-k", f"regex:{KERNEL_REGEX[name]}",
               "-s", "3", "-c", "1", sys.executable, os.path.abspath(__file__), "--only", name, "--iters", "1", "--no-flush"]
        cmd += [args.matrix] if args.matrix else ["--workload", args.workload]
        out = subprocess.run(cmd, capture_output=True, text=True).stdout
        vals = {}
        for line in out.splitlines():
            parts = [p.strip('"') for p in line.split('","')]
            if len(parts) >= 3 and parts[-3] and parts[-3][0].isalpha() and "__" in parts[-3]:
                try:
                    vals[parts[-3]] = (float(parts[-1].replace(",", "")), parts[-2])
                except ValueError:
                    pass
        rows.append((name, vals))
    print(f"{'solution':14s} {'kernel us':>10s} {'L2 hits':>14s} {'L2 misses':>14s} {'L2 hit %':>9s} {'L1 hit %':>9s} {'DRAM MB':>10s}")
    for name, v in rows:
        def g(k, scale=1.0):
            return v[k][0] * scale if k in v else float("nan")
        us = g("gpu__time_duration.sum") * ({"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(v.get("gpu__time_duration.sum", (0, "us"))[1], 1.0))
        def mb(k):
            if k not in v:
                return float("nan")
            return v[k][0] * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(v[k][1], 1e-6)
        print(f"{name:14s} {us:10.1f} {g('lts__t_sectors_lookup_hit.sum'):14.0f} {g('lts__t_sectors_lookup_miss.sum'):14.0f} "
              f"{g('lts__t_sector_hit_rate.pct'):9.1f} {g('l1tex__t_sector_hit_rate.pct'):9.1f} "
              f"{mb('dram__bytes_read.sum') + mb('dram__bytes_write.sum'):10.1f}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("matrix", nargs="?", help="Matrix Market file (else --workload)")
    ap.add_argument("--workload", default="web")
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-flush", action="store_true")
    ap.add_argument("--locality", action="store_true")
    args = ap.parse_args()
    if args.locality:
        run_locality(args)
    else:
        run_solutions(args)


if __name__ == "__main__":
    main()
