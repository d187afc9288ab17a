"""GPU suite, BASELINE.json's full sizes: configs[0] web-Google-shaped against the oracle port
(structure bit-exact at the automatic chunk count, y within 1e-12), configs[2] R-MAT-24 and configs[3] road
24M rows against the device self-check (cvr_verify_csr, itself pinned to the oracle loop in
test_gpu_parity.py) plus size-independent properties: linearity and y(x = 1) = row sums."""
import numpy as np
import pytest

import oracle
from helpers import assert_structure_equal, assert_y_close, to_oracle_csr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cvr(native_lib):
    import cvr_b200
    return cvr_b200


def test_web_google_shaped_full_size_vs_oracle_port(cvr):
    from cvr_b200 import gen
    d = gen.powerlaw_web(device="cuda")  # 916,428 rows, ~5.1 M nnz
    csr = to_oracle_csr(d)
    x = np.random.default_rng(21).uniform(-1, 1, csr.n_cols + 1)
    with cvr.CvrMatrix(d, 0) as m:
        T = m.n_chunks
        want = oracle.convert(csr, T, "port", fill_missing_tail=True)
        assert_structure_equal(m.export(), want, f"web full size T={T}")
        y, _ = m.spmv(x, iters=2)
        assert_y_close(y, csr, x, "web full size")
    with cvr.CvrMatrix(d, 16) as m:  # the reference's own scale: 16 host threads
        assert_structure_equal(m.export(), oracle.convert(csr, 16, "port", fill_missing_tail=True), "web T=16")


@pytest.mark.parametrize("name", ["rmat24", "road"])
def test_largest_single_gpu_configs_vs_device_self_check(cvr, name):
    import torch
    from cvr_b200 import gen
    d = gen.rmat(24, 16, device="cuda") if name == "rmat24" else gen.road(24_000_000, device="cuda")
    n = d.n_rows
    stream = torch.cuda.current_stream().cuda_stream
    with cvr.CvrMatrix(d, 0) as m:
        g = torch.Generator(device="cuda").manual_seed(3)
        x1 = torch.rand(n + 1, generator=g, device="cuda", dtype=torch.float64) - 0.5
        x2 = torch.rand(n + 1, generator=g, device="cuda", dtype=torch.float64) - 0.5
        x1[0] = x2[0] = 0.0
        ys = []
        for xv in (x1, x2, 2.0 * x1 - 3.0 * x2, torch.ones(n + 1, dtype=torch.float64, device="cuda")):
            y = torch.empty(n + 1, dtype=torch.float64, device="cuda")
            m.spmv_device(xv, y, stream)
            m.spmv_device(xv, y, stream)  # repeated: accumulated rows are re-cleared every sweep
            torch.cuda.synchronize()
            r = cvr.verify_csr(d, xv, y)
            assert r["rows_failing"] == 0, (name, r)
            ys.append(y)
        # linearity, row by row (bound scaled by the magnitudes that enter each row)
        rd = d.row_delim.to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(n + 1, device="cuda"), rd[1:] - rd[:-1])
        mag = torch.zeros(n + 1, dtype=torch.float64, device="cuda").index_add_(0, rows, d.val.abs())
        lin = (ys[2] - (2.0 * ys[0] - 3.0 * ys[1])).abs()
        assert bool((lin <= 1e-11 * mag + 1e-300).all())
        # x = 1 (the reference's own input, spmv.cpp:556-563): y = row sums
        sums = torch.zeros(n + 1, dtype=torch.float64, device="cuda").index_add_(0, rows, d.val)
        assert bool(((ys[3] - sums).abs() <= 1e-12 * mag + 1e-300).all())
        assert float(ys[3][0]) == 0.0
