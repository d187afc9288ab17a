#!/usr/bin/env python
"""BASELINE configs[0] the way the reference is run: a web-Google-shaped .mtx file through BOTH command
lines -- the unmodified reference (oracle/_ref/spmv.cvr.ref, host cores) and cvr_b200/bin/spmv.cvr (GPU) --
`<file> <numThreads> <numIterations>`, and the three greppable lines of each side by side.
Test/bench infrastructure (it runs the reference build); writes gpurun_out/cli_side_by_side.txt."""
import os
import re
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # tests/ -> repo root
sys.path.insert(0, ROOT)


def main():
    iters = sys.argv[1] if len(sys.argv) > 1 else "1000"
    import torch
    from cvr_b200 import gen, write_mtx
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    d = gen.powerlaw_web(device=dev).to_host()
    rows = d.row_of_entry()
    path = "/tmp/web_google_shaped.mtx"
    t0 = time.time()
    write_mtx(path, d.n_rows, d.n_cols, rows, d.col, d.val)  # padding zeros written as explicit entries
    out = [f"# {path}: {d.n_rows} rows, {d.nnz} entries, written in {time.time() - t0:.1f} s"]
    threads = str(os.cpu_count() or 1)
    env = dict(os.environ, OMP_PROC_BIND="true")
    for name, cmd in (("reference (host, %s threads)" % threads,
                       [os.path.join(ROOT, "oracle", "_ref", "spmv.cvr.ref"), path, threads, iters]),
                      ("cvr_b200 (GPU, auto chunks)", [os.path.join(ROOT, "cvr_b200", "bin", "spmv.cvr"), path, "0", iters])):
        if not os.path.exists(cmd[0]):
            out.append(f"## {name}: {cmd[0]} missing")
            continue
        t0 = time.time()
        p = subprocess.run(cmd, capture_output=True, text=True, env=env)
        out.append(f"## {name}: exit {p.returncode}, wall {time.time() - t0:.2f} s")
        for ln in p.stdout.splitlines():
            if re.search(r"Pre-processing|SpMV Execution|Throughput|Very Good|Warning|achieved|max \|y|ingest|conversion kernels", ln):
                out.append("   " + ln)
        if p.returncode not in (0,):
            out.append("   stderr: " + p.stderr[-300:])
    text = "\n".join(out)
    print(text)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "cli_side_by_side.txt"), "w") as f:
        f.write(text + "\n")


if __name__ == "__main__":
    main()
