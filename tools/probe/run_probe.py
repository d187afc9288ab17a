#!/usr/bin/env python
"""Gather / stream ceilings of one B200 on the CVR column arrays of the BASELINE workloads
(measurement aid, see gather_probe.cu).  Prints one JSON object per (workload, variant, U, blocks).

    python tools/probe/run_probe.py [--workloads rmat24,web] [--out gpurun_out/probe.jsonl]
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libgather_probe.so")


def build():
    src = os.path.join(HERE, "gather_probe.cu")
    if os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return
    # ptxas -O1 keeps the program order (all loads of a tile, then the FMAs); at -O3 it interleaves
    # them to save registers, which serialises the gathers of this synthetic loop
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-Xcompiler", "-fPIC", "-Xptxas", "-O1", "-shared", "-o", LIB, src])


class _Raw:
    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="rmat24,web")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "probe.jsonl"))
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only-chunked", action="store_true")
    ap.add_argument("--hot", action="store_true", help="only the hot-x cache experiment (variant 7)")
    args = ap.parse_args()
    build()
    import torch
    import cvr_b200
    sys.path.insert(0, ROOT)
    from bench import make_workload
    lib = C.CDLL(LIB)
    lib.probe_run.restype = C.c_int
    lib.probe_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                              C.c_int, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_float)]
    dev = torch.device("cuda", 0)
    sm_mhz = 1965.0
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fout = open(args.out, "a")
    for name in args.workloads.split(","):
        d, desc, _ = make_workload(name, 1, dev, row_normalise=True)
        n_cols, nnz = d.n_cols, d.nnz
        m = cvr_b200.CvrMatrix(d, 0, 0)
        del d
        torch.cuda.empty_cache()
        vals, cols, _ = m.device_arrays()
        x = torch.rand(n_cols + 1, device=dev, dtype=torch.float64)
        out = torch.zeros(64, device=dev, dtype=torch.float64)
        flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev) if 12 * nnz < 256e6 else None

        def run(variant, U, blocks, mask=0, carve=-1):
            best = 1e30
            for _ in range(3 if flush is not None else 1):
                if flush is not None:
                    flush.fill_(1)
                ms = C.c_float()
                rc = lib.probe_run(variant, U, cols, vals, nnz, x.data_ptr(), out.data_ptr(), blocks, mask, carve,
                                   1 if flush is not None else 3, C.byref(ms))
                if rc != 0:
                    return None
                best = min(best, ms.value)
            rec = {"workload": name, "variant": variant, "U": U, "blocks_per_sm": blocks, "warps_per_sm": 4 * blocks,
                   "col_mask": mask, "carveout": carve, "us": best * 1e3,
                   "elem_per_clk_per_sm": nnz / (best * 1e-3) / (sm_mhz * 1e6) / sms,
                   "stream_gbs": (12 if variant >= 2 else 4) * nnz / (best * 1e-3) / 1e9}
            line = json.dumps(rec)
            print(line, flush=True)
            fout.write(line + "\n")
            fout.flush()
            return rec

        grid = [(0, 9, 8), (0, 9, 16)]
        for variant in (1, 2, 3, 4, 5):
            for U in (4, 9, 18):
                for blocks in (4, 6, 8, 12, 16):
                    if args.quick and (U == 4 or blocks in (4, 12)):
                        continue
                    if variant == 5 and U == 18 and blocks > 6:
                        continue  # 2 x 18 x (8+8) bytes of registers per thread
                    grid.append((variant, U, blocks))
        if args.hot:
            # relabel columns by popularity: rank 0 = most frequent
            ncol = n_cols + 1
            cview = torch.as_tensor(_Raw(cols, nnz, "<i4"), device=dev)
            counts = torch.bincount(cview.long(), minlength=ncol)
            order = torch.argsort(counts, descending=True)
            rank = torch.empty(ncol, dtype=torch.int32, device=dev)
            rank[order] = torch.arange(ncol, dtype=torch.int32, device=dev)
            cols_perm = rank[cview.long()].contiguous()
            csum = torch.cumsum(counts[order].double(), 0) / float(nnz)
            lib.probe_hot_run.restype = C.c_int
            lib.probe_hot_run.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_int, C.c_uint32, C.c_int, C.POINTER(C.c_float)]
            for hot_k in (0, 1024, 4096, 8192, 16384, 24576):
                for U, blocks, threads in ((9, 1, 512), (9, 1, 768), (9, 1, 1024), (4, 1, 1024), (9, 2, 384), (9, 2, 512)):
                    if hot_k * 8 * blocks > 200 * 1024:
                        continue
                    best = 1e30
                    for _ in range(3 if flush is not None else 1):
                        if flush is not None:
                            flush.fill_(1)
                        ms = C.c_float()
                        rc = lib.probe_hot_run(U, cols_perm.data_ptr(), vals, nnz, x.data_ptr(), out.data_ptr(), blocks,
                                               threads, hot_k, 1 if flush is not None else 3, C.byref(ms))
                        if rc == 0:
                            best = min(best, ms.value)
                    rec = {"workload": name, "variant": 7, "U": U, "blocks_per_sm": blocks, "threads": threads,
                           "warps_per_sm": blocks * threads // 32, "hot_k": hot_k,
                           "hot_hit_fraction": float(csum[hot_k - 1]) if hot_k else 0.0, "us": best * 1e3,
                           "elem_per_clk_per_sm": nnz / (best * 1e-3) / (sm_mhz * 1e6) / sms}
                    line = json.dumps(rec)
                    print(line, flush=True)
                    fout.write(line + "\n")
                    fout.flush()
            m.close()
            del x, out, flush, cols_perm
            torch.cuda.empty_cache()
            continue
        if args.only_chunked:
            grid = [(2, 9, 6), (2, 9, 8)]
        for variant, U, blocks in grid:
            run(variant, U, blocks)
        # the sweep's access pattern: every warp streams its own contiguous chunk
        for chunk_el in (4032, 1152):  # multiples of 32*4, 32*9, 32*18
            lib.probe_set_chunk(chunk_el)
            for U, blocks in ((9, 6), (9, 8), (18, 4)):
                r = run(6, U, blocks)
                if r:
                    r["chunk_el"] = chunk_el
        # x shrunk to an L2-resident window: separates "x misses L2" from "L1/L2 request rate"
        for mask in ((1 << 20) - 1, (1 << 14) - 1):
            for variant, U, blocks in ((2, 9, 8), (2, 18, 8), (4, 9, 6)):
                run(variant, U, blocks, mask)
        # maximum L1 (carveout 0 = prefer L1)
        for variant, U, blocks in ((2, 9, 8), (2, 18, 8)):
            run(variant, U, blocks, 0, 0)
        m.close()
        del x, out, flush
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
