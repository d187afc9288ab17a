#!/usr/bin/env python
"""Conversion only (CSR -> CVR) on a BASELINE workload, for profiling the conversion kernels under ncu.
    python tools/convert_probe.py road [repeats]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import cvr_b200
    from bench import make_workload
    name = sys.argv[1] if len(sys.argv) > 1 else "road"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    d, desc, _ = make_workload(name, 1, torch.device("cuda", 0), row_normalise=False)
    torch.cuda.synchronize()
    for _ in range(reps):
        m = cvr_b200.CvrMatrix(d, 0, 0)
        i = m.info
        print(f"{name}: chunks {i['n_chunks']}, conversion kernels {i['convert_kernel_seconds'] * 1e3:.3f} ms, "
              f"row lists {i['row_lists_seconds'] * 1e3:.3f} ms, create {i['create_seconds'] * 1e3:.2f} ms", flush=True)
        m.close()


if __name__ == "__main__":
    main()
