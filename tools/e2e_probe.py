#!/usr/bin/env python
"""cvr_spmv end to end (pinned host x and y, copies inside) and first / second creation time for one workload;
run once per allocator setting:  CVR_NO_POOL=1 python tools/e2e_probe.py rmat24"""
import os
import sys
import time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cvr_b200  # noqa: E402
from bench import make_workload  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "rmat24"
dev = torch.device("cuda", 0)
d, desc, _ = make_workload(name, 1, dev, row_normalise=True)
torch.cuda.synchronize()
creates = []
for k in range(2):
    m = cvr_b200.CvrMatrix(d, 0, 0)
    creates.append(m.info["create_seconds"] * 1e3)
    if k == 0:
        m.close()
xh = (torch.rand(d.n_cols + 1, dtype=torch.float64) - 0.5).pin_memory()
yh = torch.empty(d.n_rows + 1, dtype=torch.float64).pin_memory()
for _ in range(3):
    m.spmv_into(xh, yh, 1)
ts = []
for _ in range(10):
    t0 = time.perf_counter()
    m.spmv_into(xh, yh, 1)
    ts.append(time.perf_counter() - t0)
ts.sort()
print(f"{name} pool={'off' if os.environ.get('CVR_NO_POOL') == '1' else 'on'}: create first {creates[0]:.2f} ms, second {creates[1]:.2f} ms; "
      f"e2e median {ts[5] * 1e3:.3f} ms, min {ts[0] * 1e3:.3f} ms, {2 * d.nnz_true / ts[5] / 1e9:.1f} GFLOP/s")
