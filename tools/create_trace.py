#!/usr/bin/env python
"""Where cvr_create spends its time: phase trace (CVR_CREATE_TRACE=1) for a HOST CSR and a DEVICE CSR of one workload,
three creations each (the first pays one-time costs)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["CVR_CREATE_TRACE"] = "1"
import torch  # noqa: E402
import cvr_b200  # noqa: E402
from bench import make_workload  # noqa: E402

for name in sys.argv[1:] or ["web"]:
    d, desc, _ = make_workload(name, 1, torch.device("cuda", 0), row_normalise=True)
    host = d.to_host()
    for kind, csr in (("host", host), ("device", d)):
        for k in range(3):
            print(f"== {name} {kind} CSR, creation {k}", file=sys.stderr, flush=True)
            m = cvr_b200.CvrMatrix(csr, 0, 0)
            print(f"   create_seconds {m.info['create_seconds'] * 1e3:.3f} ms", file=sys.stderr, flush=True)
            m.close()
