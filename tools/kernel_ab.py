#!/usr/bin/env python
"""A/B of the SpMV kernel variants on the BASELINE workloads: one matrix build per workload, every
variant timed on it (CUDA events around the sweep kernel alone and around whole steps), y checked
against a device CSR product.  Development aid; prints one JSON object per (workload, variant).

    python tools/kernel_ab.py --workloads rmat24,web,road,fem --variants tma,ldg,window
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workloads", default="rmat24,web,road,fem")
    ap.add_argument("--variants", default="tile7x5r,tile11x5,tile7x6,tile7x6r")
    ap.add_argument("--env", default="CVR_SPMV_KERNEL")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--chunks", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "kernel_ab.jsonl"))
    args = ap.parse_args()
    import torch
    import cvr_b200
    from bench import make_workload, measured_peaks
    dev = torch.device("cuda", 0)
    peak, _ = measured_peaks()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    fout = open(args.out, "a")
    for name in args.workloads.split(","):
        d, desc, _ = make_workload(name, 1, dev, row_normalise=True)
        n = d.n_rows
        x = torch.rand(d.n_cols + 1, device=dev, dtype=torch.float64) - 0.5
        x[0] = 0.0
        rd = d.row_delim.to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(n + 1, device=dev), rd[1:] - rd[:-1])
        prod = d.val * x[d.col.long()]
        y_ref = torch.zeros(n + 1, dtype=torch.float64, device=dev).index_add_(0, rows, prod)
        mag = torch.zeros(n + 1, dtype=torch.float64, device=dev).index_add_(0, rows, prod.abs())
        del rows, prod, rd
        nnz_true = d.nnz_true
        for variant in args.variants.split(","):
            if variant == "auto":  # the per-matrix choice of the library
                os.environ.pop(args.env, None)
            else:
                os.environ[args.env] = variant
            m = cvr_b200.CvrMatrix(d, args.chunks, 0)
            info = m.info
            y = torch.empty(n + 1, dtype=torch.float64, device=dev)
            stream = torch.cuda.current_stream()
            flush = torch.empty(384 << 20, dtype=torch.uint8, device=dev) if info["algorithmic_bytes"] < 256e6 else None
            for _ in range(3):
                m.spmv_device(x, y, stream.cuda_stream)
            torch.cuda.synchronize()
            bad = int(((y - y_ref).abs() > 1e-12 * mag + 1e-300).sum())
            m.set_kernel_timing(True)
            step_ms = 0.0
            for _ in range(args.steps):
                if flush is not None:
                    flush.fill_(1)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                m.spmv_device(x, y, stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize()
                step_ms += e0.elapsed_time(e1)
            ksecs, kl = m.kernel_timing()
            m.set_kernel_timing(False)
            kus = ksecs / max(kl, 1) * 1e6
            rec = {"workload": name, "variant": variant, "kernel": m.kernel_name, "chunks": info["n_chunks"], "kernel_us": kus,
                   "step_us": step_ms / args.steps * 1e3, "gflops": 2.0 * nnz_true / (kus * 1e-6) / 1e9,
                   "frac": info["algorithmic_bytes"] / (kus * 1e-6) / 1e9 / peak, "rows_failing": bad}
            line = json.dumps(rec)
            print(line, flush=True)
            fout.write(line + "\n")
            fout.flush()
            m.close()
            del y, flush
        del d, x, y_ref, mag
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
