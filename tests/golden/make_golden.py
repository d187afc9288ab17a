"""Regenerate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and AVX-512F):
    python tests/golden/make_golden.py
It builds oracle/_ref (the reference compiled from where it lies), runs the reference's own
readMatrix / pre_processing / spmv_compute_kernel on small seeded inputs and stores inputs and
outputs as .npz.  The reference ships no golden vectors of its own, so
these files are what pins the oracle port and the CUDA path on hosts where the reference
cannot run.  Fixture format (per file): CSR arrays + for every chunk count T the reference's
vals / cols / nnz_rows / split / tail (8 per chunk) / record lists, and y where the
reference kernel itself is correct.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from cvr_b200 import gen  # noqa: E402


def csr_of(d):
    h = d.to_host()
    return oracle.Csr(h.n_rows, h.n_cols, h.val, h.col, h.row_delim, h.nnz_true)


def tiny_equal_rows():
    # 8 rows x 2 entries: one chunk, all lanes end together -> the reference never stores final_2
    rows = np.repeat(np.arange(1, 9), 2)
    cols = np.tile(np.array([1, 5]), 8) + np.repeat(np.arange(8), 2) % 3
    vals = np.arange(1, 17, dtype=np.float64)
    rd = np.concatenate([[0], np.arange(0, 17, 2)])
    return oracle.Csr(8, 8, vals, cols, rd)


CASES = {
    "rand_small": (lambda: csr_of(gen.random_sparse(600, 600, 3000, seed=11, empty_frac=0.2)), [1, 2, 3, 8, 32]),
    "long_rows": (lambda: csr_of(gen.random_sparse(300, 300, 800, seed=12, long_rows=3, long_len=400)), [1, 4, 16]),
    "fem6": (lambda: csr_of(gen.fem27(6, 6, 6)), [1, 5, 16]),
    "rmat9": (lambda: csr_of(gen.rmat(9, 8, seed=13)), [1, 7, 64]),
    "road3k": (lambda: csr_of(gen.road(3000, seed=14)), [1, 6, 40]),
    "tiny_equal_rows": (tiny_equal_rows, [1]),
}


def pack(csr, chunk_counts):
    out = {"n_rows": csr.n_rows, "n_cols": csr.n_cols, "csr_val": csr.val, "csr_col": csr.col,
           "csr_rd": csr.row_delim, "chunk_counts": np.array(chunk_counts)}
    x = np.random.default_rng(5).uniform(-1, 1, csr.n_cols + 1)
    out["x"] = x
    yc, mag = oracle.csr_spmv(csr, x)
    for T in chunk_counts:
        r = oracle.convert(csr, T, "ref")
        recs, lens = [], []
        for t in range(T):
            cr = oracle.chunk_records(r, t)
            recs.append(cr)
            lens.append(cr.shape[0])
        out[f"T{T}_vals"] = r["vals"].copy()
        out[f"T{T}_cols"] = r["cols"].copy()
        out[f"T{T}_nnz_rows"] = r["nnz_rows"].copy()
        out[f"T{T}_split"] = r["split"].copy()
        out[f"T{T}_tail"] = r["final_2"].reshape(T, 16)[:, :8].copy()
        out[f"T{T}_records"] = np.concatenate(recs)
        out[f"T{T}_record_lens"] = np.array(lens)
        y, _ = oracle.spmv(r, csr.n_rows, x, "ref")
        ok = bool(np.all(np.abs(y - yc)[1:] <= 1e-12 * mag[1:]))
        out[f"T{T}_ref_kernel_ok"] = ok
        if ok:
            out[f"T{T}_y"] = y
    return out


def main():
    oracle.build(ref=True)
    assert oracle.ref_lib() is not None, "reference build unavailable"
    for name, (make, Ts) in CASES.items():
        csr = make()
        np.savez_compressed(os.path.join(HERE, f"ref_{name}.npz"), **pack(csr, Ts))
        print(name, "nnz", csr.nnz, "T", Ts)
    # ingest fixtures: what the reference's readMatrix returns for quirky files
    files = {
        "kat12.mtx": None,  # committed by hand
        "pattern_symmetric.mtx": "%%MatrixMarket matrix coordinate pattern symmetric\n% a comment\n4 4 3\n2 1\n3 3\n4 2\n",
        "no_trailing_newline.mtx": "%%MatrixMarket matrix coordinate real general\n3 3 3\n1 1 1.5\n2 3 -2.25\n3 2 0.1",
        "unsorted_dups.mtx": "%%MatrixMarket matrix coordinate real general\n5 6 8\n5 6 1.0\n1 2 0.3333333\n5 1 2.0\n"
                             "3 3 7.0\n1 2 4.0\n5 6 9.0\n2 6 1e-3\n5 5 -1.0\n",
    }
    ingest = {}
    for fname, text in files.items():
        path = os.path.join(HERE, fname)
        if text is not None:
            with open(path, "w") as f:
                f.write(text)
        c = oracle.read_mtx(path, "ref")
        key = fname.replace(".mtx", "")
        ingest[f"{key}_shape"] = np.array([c.n_rows, c.n_cols, c.nnz])
        ingest[f"{key}_val"] = c.val
        ingest[f"{key}_col"] = c.col
        ingest[f"{key}_rd"] = c.row_delim
    np.savez_compressed(os.path.join(HERE, "ref_ingest.npz"), **ingest)
    print("ingest fixtures:", list(files))


if __name__ == "__main__":
    main()
