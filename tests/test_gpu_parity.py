"""GPU suite (B200): the CUDA path through the C ABI against the oracle.

* structure arrays: bit-exact against the reference fixtures (tests/golden) and against the
  oracle port on generated matrices, at several chunk counts;
* y: per-row |dy| <= 1e-12 * sum|a x| against the reference's scalar CSR loop (oracle port);
* edge cases: steal-only chunks, empty rows, one chunk, the largest legal chunk count,
  a chunk that never stores its tail, non-square matrices;
* full-size properties: linearity and a checksum at a size the oracle would take long on.
"""
import os
import subprocess

import numpy as np
import pytest

import oracle
from helpers import (GOLDEN, assert_structure_equal, assert_y_close, golden_cases, golden_structure,
                     load_golden, to_oracle_csr)

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cvr(native_lib):
    import cvr_b200
    return cvr_b200


# every test of this module runs once per sweep geometry: the per-matrix choice (cvr_pick_sweep_variant)
# and each geometry forced through CVR_SPMV_KERNEL (read per launch)
@pytest.fixture(autouse=True, params=["auto", "tile7x5r", "tile11x5", "tile7x6", "tile7x6r"])
def sweep_geometry(request, monkeypatch):
    if request.param == "auto":
        monkeypatch.delenv("CVR_SPMV_KERNEL", raising=False)
    else:
        monkeypatch.setenv("CVR_SPMV_KERNEL", request.param)
    return request.param


def host_csr(cvr, c):
    return cvr.CsrMatrix(c.n_rows, c.n_cols, c.val, c.col, c.row_delim, c.nnz_file)


@pytest.mark.parametrize("name", golden_cases())
def test_structure_bit_exact_vs_reference_fixture(cvr, name):
    z, csr = load_golden(name)
    for T in z["chunk_counts"]:
        T = int(T)
        with cvr.CvrMatrix(host_csr(cvr, csr), T) as m:
            assert_structure_equal(m.export(), golden_structure(z, T), f"{name} T={T}")


@pytest.mark.parametrize("name", golden_cases())
def test_y_within_tolerance_on_fixtures(cvr, name):
    z, csr = load_golden(name)
    x = z["x"]
    for T in z["chunk_counts"]:
        with cvr.CvrMatrix(host_csr(cvr, csr), int(T)) as m:
            y, secs = m.spmv(x, iters=3)  # repeated: y is re-zeroed every iteration
            assert secs > 0
            assert_y_close(y, csr, x, f"{name} T={T}")


def test_kat12_through_the_abi(cvr):
    m = cvr.read_matrix(os.path.join(GOLDEN, "kat12.mtx"))
    want_y = [0, 315, 207, 0, 807, 2534, 607, 710, 3222, 1813, 1007, 6639, 7233]
    for T in (1, 2):
        with cvr.CvrMatrix(m, T) as a:
            y, _ = a.spmv(np.ones(13))
            assert y.tolist() == want_y
            e = a.export()
            if T == 1:
                assert e["split"].tolist() == [0, 14]
                assert e["final_2"][:8].tolist() == [1, 10, 9, 4, 5, 11, 12, 8]


GENERATED = {
    "rand": (lambda g: g.random_sparse(20000, 15000, 150000, seed=41, empty_frac=0.25), [1, 7, 64, 1000, 5000]),
    "long": (lambda g: g.random_sparse(3000, 3000, 9000, seed=42, long_rows=6, long_len=2500), [1, 16, 300, 1400]),
    "web": (lambda g: g.powerlaw_web(60000, 330000, seed=43), [8, 592, 4736]),
    "fem": (lambda g: g.fem27(20, 20, 20), [3, 148, 2000]),
    "rmat": (lambda g: g.rmat(14, 16, seed=44), [5, 1184, 9000]),
    "road": (lambda g: g.road(200000, seed=45), [2, 500, 8000]),
}


@pytest.mark.parametrize("name", list(GENERATED))
def test_generated_matrices_vs_oracle_port(cvr, name):
    from cvr_b200 import gen
    make, Ts = GENERATED[name]
    d = make(gen)
    csr = to_oracle_csr(d)
    rng = np.random.default_rng(7)
    x = rng.uniform(-1, 1, csr.n_cols + 1)
    for T in Ts:
        T = min(T, csr.nnz // 16)
        want = oracle.convert(csr, T, "port", fill_missing_tail=True)
        with cvr.CvrMatrix(d.to_host(), T) as m:
            assert_structure_equal(m.export(), want, f"{name} T={T}")
            y, _ = m.spmv(x)
            assert_y_close(y, csr, x, f"{name} T={T}")
            y1, _ = m.spmv(np.ones(csr.n_cols + 1))
            assert_y_close(y1, csr, np.ones(csr.n_cols + 1), f"{name} T={T} x=1")


def test_warp_per_chunk_scheduler_still_bit_exact(cvr, monkeypatch):
    """CVR_SCHEDULE=warp selects the one-warp-per-chunk lane scheduler (round 1); the default schedules a chunk
    with 8 threads.  Both must emit the reference's structure bit for bit."""
    from cvr_b200 import gen
    monkeypatch.setenv("CVR_SCHEDULE", "warp")
    for name in ("long", "road", "rmat"):
        make, Ts = GENERATED[name]
        d = make(gen)
        csr = to_oracle_csr(d)
        for T in Ts:
            T = min(T, csr.nnz // 16)
            with cvr.CvrMatrix(d.to_host(), T) as m:
                assert m.info["kernel_launches"] == 5  # no row bitmap kernel
                assert_structure_equal(m.export(), oracle.convert(csr, T, "port", fill_missing_tail=True), f"{name} T={T}")


def test_maximum_chunk_count_and_single_chunk(cvr):
    from cvr_b200 import gen
    d = gen.random_sparse(500, 700, 4000, seed=46, empty_frac=0.1)
    csr = to_oracle_csr(d)
    x = np.random.default_rng(3).uniform(-1, 1, csr.n_cols + 1)
    for T in (1, csr.nnz // 16):  # every chunk holds exactly 16 elements at the maximum
        with cvr.CvrMatrix(d.to_host(), T) as m:
            assert_structure_equal(m.export(), oracle.convert(csr, T, "port", fill_missing_tail=True), f"T={T}")
            y, _ = m.spmv(x)
            assert_y_close(y, csr, x, f"T={T}")
    with pytest.raises(cvr.CvrError):
        cvr.CvrMatrix(d.to_host(), csr.nnz // 16 + 1)


def test_auto_chunks_and_device_csr_entry(cvr):
    import torch
    from cvr_b200 import gen
    d = gen.fem27(30, 30, 30, device="cuda")
    csr = to_oracle_csr(d)
    x = np.random.default_rng(4).uniform(-1, 1, csr.n_cols + 1)
    with cvr.CvrMatrix(d, 0) as m:  # n_chunks = 0: cvr_auto_chunks, CSR already on the device
        info = m.info
        assert info["n_chunks"] % torch.cuda.get_device_properties(0).multi_processor_count == 0
        assert info["kernel_launches"] == 6  # row bitmap, schedule, permute, mark, 2 x collect
        want = oracle.convert(csr, info["n_chunks"], "port", fill_missing_tail=True)
        assert_structure_equal(m.export(), want, "auto")
        y, _ = m.spmv(x)
        assert_y_close(y, csr, x, "auto")
        # device-vector entry on a torch stream
        xd = torch.from_numpy(x).cuda()
        yd = torch.full((csr.n_rows + 1,), 7.0, dtype=torch.float64, device="cuda")
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        m.spmv_device(xd, yd, s.cuda_stream)
        s.synchronize()
        assert_y_close(yd.cpu().numpy(), csr, x, "spmv_device")
        assert m.info["algorithmic_bytes"] == (12 * csr.nnz + 8 * info["n_records"] + 56 * info["n_chunks"]
                                               + 8 * (csr.n_cols + 1) + 8 * (csr.n_rows + 1))


def test_iteration_loop_replayed_from_a_cuda_graph(cvr, monkeypatch):
    """cvr_spmv with many iterations replays a captured graph of 20 iterations (the reference's loop,
    spmv.cpp:1024-1034): same y as the plain loop, every iteration re-clears the accumulated rows, and the
    launch count is what the plain loop would have issued."""
    from cvr_b200 import gen
    d = gen.powerlaw_web(20000, 110000, seed=52)
    csr = to_oracle_csr(d)
    x = np.random.default_rng(12).uniform(-1, 1, csr.n_cols + 1)
    with cvr.CvrMatrix(d.to_host(), 300) as m:
        l0 = m.info["kernel_launches"]
        y_graph, secs = m.spmv(x, iters=45)  # 2 graph launches + 5 plain iterations
        assert secs > 0 and m.info["kernel_launches"] - l0 == 90
        assert_y_close(y_graph, csr, x, "graph loop")
        monkeypatch.setenv("CVR_NO_GRAPH", "1")
        y_plain, _ = m.spmv(x, iters=45)
        assert_y_close(y_plain, csr, x, "plain loop")
        monkeypatch.delenv("CVR_NO_GRAPH")
        y_again, _ = m.spmv(2.0 * x, iters=21)  # the cached graph reads the handle's x buffer: new x, same graph
        assert_y_close(y_again, csr, 2.0 * x, "graph loop, second x")


def test_single_spmv_pipelined_in_slabs(cvr, monkeypatch):
    """cvr_spmv with iters = 1 sweeps the chunks in 8 slabs and copies the finished rows of y back behind each
    slab (large matrices only by default; forced here): same y, also when slabs cut through shared rows, long
    rows spanning several slabs and empty rows."""
    from cvr_b200 import gen
    monkeypatch.setenv("CVR_SLAB_MIN_BYTES", "0")
    for d in (gen.random_sparse(30000, 30000, 200000, seed=53, empty_frac=0.3, long_rows=3, long_len=20000),
              gen.rmat(14, 16, seed=54), gen.road(100000, seed=55)):
        csr = to_oracle_csr(d)
        x = np.random.default_rng(13).uniform(-1, 1, csr.n_cols + 1)
        for T in (512, 3000):
            with cvr.CvrMatrix(d.to_host(), T) as m:
                l0 = m.info["kernel_launches"]
                y, secs = m.spmv(x, iters=1)
                assert m.info["kernel_launches"] - l0 == 9 and secs > 0  # one clearing kernel + 8 slabs
                assert_y_close(y, csr, x, f"slabs T={T}")
                y2, _ = m.spmv(x, iters=2)  # the plain loop on the same handle
                assert_y_close(y2, csr, x, f"after slabs T={T}")


def test_save_load_round_trip(cvr, tmp_path):
    """cvr_save / cvr_load: the reloaded matrix exports the same structure bit for bit and gives the
    same y (conversion skipped)."""
    from cvr_b200 import gen
    d = gen.random_sparse(5000, 4000, 40000, seed=49, empty_frac=0.15, long_rows=2, long_len=1500)
    csr = to_oracle_csr(d)
    x = np.random.default_rng(6).uniform(-1, 1, csr.n_cols + 1)
    p = str(tmp_path / "m.cvr")
    with cvr.CvrMatrix(d.to_host(), 200) as m:
        a = m.export()
        ya, _ = m.spmv(x)
        m.save(p)
    with cvr.CvrMatrix.load(p) as m2:
        assert m2.info["convert_seconds"] == 0.0 and m2.n_chunks == 200
        assert_structure_equal(m2.export(), a, "reloaded")
        yb, _ = m2.spmv(x)
        assert_y_close(yb, csr, x, "reloaded")
        assert np.array_equal(ya[np.abs(ya) > 0] != 0, yb[np.abs(ya) > 0] != 0)
    assert not os.path.exists(p + ".tmp")  # written under a temporary name, renamed when complete
    good = open(p, "rb").read()
    # the loader does not trust the file: wrong magic, truncation, padding, counts and descriptors out of range
    import struct
    hdr = struct.Struct("<8s5q4i")
    fields = list(hdr.unpack_from(good))

    def rewritten(**kw):
        f = list(fields)
        names = ["magic", "n_rows", "n_cols", "nnz", "record_ints", "n_records", "n_chunks", "n_boundary", "n_empty", "chunk_bytes"]
        for k, v in kw.items():
            f[names.index(k)] = v
        return hdr.pack(*f) + good[hdr.size:]

    chunk0 = hdr.size + 12 * fields[3] + 4 * fields[4]  # first chunk descriptor: start (i64), len, first_row, ...
    bad_desc = bytearray(good)
    struct.pack_into("<i", bad_desc, chunk0 + 12, 10 ** 9)  # first_row far beyond n_rows
    for what, blob in (("magic", b"XXXX" + good[4:]), ("truncated", good[:-8]), ("padded", good + b"\0" * 16),
                       ("n_rows", rewritten(n_rows=-5)), ("n_chunks", rewritten(n_chunks=fields[6] + 1)),
                       ("n_empty", rewritten(n_empty=2 ** 30)), ("descriptor", bytes(bad_desc))):
        with open(p, "wb") as f:
            f.write(blob)
        with pytest.raises(cvr.CvrError):
            cvr.CvrMatrix.load(p)
    with open(p, "wb") as f:
        f.write(good)
    with cvr.CvrMatrix.load(p) as m3:
        assert m3.n_chunks == 200


def test_64bit_row_delimiters_entry(cvr):
    """row_delim64 is the entry for nnz >= 2^31 (config 5 shards); exercise it at a small size."""
    import torch
    from cvr_b200 import gen
    d = gen.rmat(12, 16, device="cuda", seed=48)
    csr = to_oracle_csr(d)
    d.row_delim = d.row_delim.to(torch.int64)
    x = np.random.default_rng(5).uniform(-1, 1, csr.n_cols + 1)
    with cvr.CvrMatrix(d, 300) as m:
        assert_structure_equal(m.export(), oracle.convert(csr, 300, "port", fill_missing_tail=True), "rd64")
        y, _ = m.spmv(x)
        assert_y_close(y, csr, x, "rd64")
    h = d.to_host()
    h64 = cvr.CsrMatrix(h.n_rows, h.n_cols, h.val, h.col, h.row_delim.astype(np.int32), h.nnz_true)
    h64.row_delim = h.row_delim.astype(np.int64)
    with cvr.CvrMatrix(h64, 7) as m:
        y, _ = m.spmv(x)
        assert_y_close(y, csr, x, "rd64 host")


def test_full_size_properties_linearity_and_checksum(cvr):
    """Config-2-sized input (1M rows, 26.5M nnz): too slow for the scalar port in a unit test,
    so check size-independent properties: A(ax+by) = aAx + bAy row by row within the bound, and
    sum(y) against a torch fp64 CSR reference computed on the device."""
    import torch
    from cvr_b200 import gen
    d = gen.fem27(100, 100, 100, device="cuda")
    n = d.n_rows
    with cvr.CvrMatrix(d, 0) as m:
        g = torch.Generator(device="cuda").manual_seed(1)
        x1 = torch.rand(n + 1, generator=g, device="cuda", dtype=torch.float64) - 0.5
        x2 = torch.rand(n + 1, generator=g, device="cuda", dtype=torch.float64) - 0.5
        ys = []
        for xv in (x1, x2, 2.0 * x1 - 3.0 * x2):
            y = torch.empty(n + 1, dtype=torch.float64, device="cuda")
            m.spmv_device(xv, y, torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            ys.append(y)
        rd = d.row_delim.to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(n + 1, device="cuda"), rd[1:] - rd[:-1])
        prod = d.val * x1[d.col.long()]
        y_ref = torch.zeros(n + 1, dtype=torch.float64, device="cuda").index_add_(0, rows, prod)
        mag = torch.zeros(n + 1, dtype=torch.float64, device="cuda").index_add_(0, rows, prod.abs())
        assert bool(((ys[0] - y_ref).abs() <= 1e-12 * mag + 1e-300).all())
        lin = (ys[2] - (2.0 * ys[0] - 3.0 * ys[1])).abs()
        assert bool((lin <= 1e-11 * (mag + 1.0)).all())
        assert abs(float(ys[0].sum() - y_ref.sum())) <= 1e-9 * float(mag.sum())


def test_reference_last_delimiter_quirk_is_repaired(cvr):
    """The reference reader leaves row_delim[k] = nnz-1 after the last non-empty row (spmv.cpp:522-526).
    Such a CSR -- also with trailing EMPTY rows, where the reference reads out of bounds -- must convert to
    exactly what the correct delimiters give, from host and from device memory, and must not hang."""
    import torch
    from test_oracle_property import build_csr
    for lens in ([3, 0, 5, 2, 7, 4, 0, 0], [9, 1, 30, 2], [5, 5, 5, 0, 6, 2, 0, 0, 0, 0], [40, 3, 0, 17, 2]):
        csr = build_csr(lens, seed=len(lens))
        rd_quirk = csr.row_delim.copy()
        rd_quirk[rd_quirk == csr.nnz] = csr.nnz - 1
        assert rd_quirk[-1] == csr.nnz - 1
        x = np.random.default_rng(11).uniform(-1, 1, csr.n_cols + 1)
        for T in (1, max(1, csr.nnz // 32)):
            want = oracle.convert(csr, T, "port", fill_missing_tail=True)
            host = cvr.CsrMatrix(csr.n_rows, csr.n_cols, csr.val, csr.col, rd_quirk, csr.nnz_file)
            with cvr.CvrMatrix(host, T) as m:
                assert_structure_equal(m.export(), want, f"quirk host lens={lens} T={T}")
                y, _ = m.spmv(x)
                assert_y_close(y, csr, x, f"quirk host lens={lens} T={T}")
            dev = cvr.DeviceCsr(csr.n_rows, csr.n_cols, torch.from_numpy(csr.val).cuda(),
                                torch.from_numpy(csr.col).cuda(), torch.from_numpy(rd_quirk).cuda(), csr.nnz_file)
            with cvr.CvrMatrix(dev, T) as m:
                assert_structure_equal(m.export(), want, f"quirk device lens={lens} T={T}")
    bad = csr.row_delim.copy()
    bad[-1] = csr.nnz - 5
    with pytest.raises(cvr.CvrError):
        cvr.CvrMatrix(cvr.CsrMatrix(csr.n_rows, csr.n_cols, csr.val, csr.col, bad, csr.nnz_file), 1)
    bad = csr.row_delim.copy()
    bad[2] = bad[3] + 1  # decreasing
    with pytest.raises(cvr.CvrError):
        cvr.CvrMatrix(cvr.CsrMatrix(csr.n_rows, csr.n_cols, csr.val, csr.col, bad, csr.nnz_file), 1)


def test_device_self_check_agrees_with_the_oracle_loop(cvr):
    """cvr_verify_csr (the device restatement of the reference's verdict, spmv.cpp:1843-1850/:1916-1938) against
    the oracle's scalar CSR loop: accepts the CVR result, rejects a planted error, reports its row."""
    import torch
    from cvr_b200 import gen
    d = gen.powerlaw_web(30000, 160000, device="cuda", seed=51)
    csr = to_oracle_csr(d)
    x = np.random.default_rng(8).uniform(-1, 1, csr.n_cols + 1)
    yc, mag = oracle.csr_spmv(csr, x)
    xd = torch.from_numpy(x).cuda()
    ok = cvr.verify_csr(d, xd, torch.from_numpy(yc).cuda())
    assert ok["rows_failing"] == 0 and ok["first_bad_row"] == -1 and ok["max_rel"] <= 1e-13
    with cvr.CvrMatrix(d, 0) as m:
        yd = torch.empty(csr.n_rows + 1, dtype=torch.float64, device="cuda")
        m.spmv_device(xd, yd)
        torch.cuda.synchronize()
        got = cvr.verify_csr(d, xd, yd)
        assert got["rows_failing"] == 0 and got["max_rel"] <= 1e-12
        assert_y_close(yd.cpu().numpy(), csr, x, "device y")
        row = int(np.flatnonzero(mag > 0)[7])
        yd[row] += 1e-9 * mag[row]
        bad = cvr.verify_csr(d, xd, yd)
        assert bad["rows_failing"] == 1 and bad["first_bad_row"] == row and bad["max_rel"] > 1e-10
        yd[0] = 1.0  # the phantom row must stay 0
        assert cvr.verify_csr(d, xd, yd)["rows_failing"] == 2
        assert cvr.verify_csr(d, xd, yd, check_row0=False)["rows_failing"] == 1


def test_cli_prints_the_reference_lines(cvr, tmp_path):
    from cvr_b200 import gen, write_mtx
    d = gen.powerlaw_web(5000, 30000, seed=47).to_host()
    rows = d.row_of_entry()
    keep = np.arange(d.nnz) < d.nnz  # padding zeros are written too: harmless explicit zeros
    p = str(tmp_path / "web.mtx")
    write_mtx(p, d.n_rows, d.n_cols, rows[keep], d.col[keep], d.val[keep])
    cli = os.path.join(ROOT, "cvr_b200", "bin", "spmv.cvr")
    out = subprocess.run([cli, p, "64", "10"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    for pat in ("Pre-processing", "SpMV Execution", "Throughput"):  # README.md:47-49 greps
        assert sum(pat in ln for ln in lines) == 1, pat
    assert any(ln.startswith("The Pre-processing(CSR->CVR)   Time of CVR   is ") and "[threads: 64]" in ln for ln in lines)
    assert any(ln.startswith("The SpMV Execution Time of CVR    is ") for ln in lines)
    assert any(ln.startswith("         The Throughput of CVR    is ") and "GFlops." in ln for ln in lines)
    assert "     Very Good! Your result is correct  " in lines


def test_tiny_random_matrices_property(cvr):
    """Corner cases of the lane scheduler on the device: tiny matrices (fewer than 8 rows, empty rows at
    chunk starts, rows longer than a chunk) at random legal chunk counts, CUDA vs oracle port."""
    from test_oracle_property import build_csr
    rng = np.random.default_rng(2026)
    for case in range(60):
        n_rows = int(rng.integers(1, 40))
        kind = rng.integers(0, 3, n_rows)
        lens = np.where(kind == 0, 0, np.where(kind == 1, rng.integers(0, 4, n_rows), rng.integers(0, 41, n_rows)))
        lens = lens.tolist()
        if sum(lens) == 0:
            lens.append(3)
        if lens[-1] < 2:
            lens[-1] = 2
        csr = build_csr(lens, seed=case)
        T = int(rng.integers(1, csr.nnz // 16 + 1))
        want = oracle.convert(csr, T, "port", fill_missing_tail=True)
        x = rng.uniform(-1, 1, csr.n_cols + 1)
        with cvr.CvrMatrix(host_csr(cvr, csr), T) as m:
            assert_structure_equal(m.export(), want, f"case {case} lens={lens} T={T}")
            y, _ = m.spmv(x)
            assert_y_close(y, csr, x, f"case {case} lens={lens} T={T}")


def test_chunk_queue_with_many_small_chunks(cvr, monkeypatch):
    """The sweep hands out chunks beyond a warp's first from an atomic ticket queue once a launch covers at least four
    chunks per resident warp (cvr_launch_spmv); the queue resets itself when the last warp leaves.  A matrix cut into
    16 000 chunks of ~20 elements crosses that threshold on a B200 (<= 3552 resident warps): y must be right on
    every one of several back-to-back launches, with the queue and with the static round robin, and the two must
    agree to rounding (only the order of the atomic additions into rows shared by chunks may differ)."""
    from cvr_b200 import gen
    d = gen.random_sparse(60000, 50000, 320000, seed=48, empty_frac=0.2, long_rows=3, long_len=9000)
    csr = to_oracle_csr(d)
    T = min(16000, csr.nnz // 16)
    x = np.random.default_rng(11).uniform(-1, 1, csr.n_cols + 1)
    ys = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CVR_DYNAMIC_CHUNKS", mode)
        with cvr.CvrMatrix(d.to_host(), T) as m:
            for rep in range(4):  # the queue counters must come back to zero after every launch
                y, _ = m.spmv(x, iters=1 if rep < 3 else 5)
                assert_y_close(y, csr, x, f"queue={mode} launch {rep}")
            ys[mode] = y
    assert_y_close(ys["1"], csr, x, "queue on, final")
    assert np.max(np.abs(ys["1"] - ys["0"])) <= 1e-12 * np.max(np.abs(ys["0"]))
