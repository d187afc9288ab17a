"""Build recipe for the native parts: libcvr_b200.so (CUDA kernels + C ABI) and the
`spmv.cvr` command line.  sm_100a only; artefacts stay in-tree (cvr_b200/lib, cvr_b200/bin)
so they travel to the GPU box with the snapshot."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
BIN_DIR = os.path.join(PKG, "bin")
LIB = os.path.join(LIB_DIR, "libcvr_b200.so")
CLI = os.path.join(BIN_DIR, "spmv.cvr")

CUDA_SOURCES = ["cvr_api.cu", "cvr_convert.cu", "cvr_spmv.cu", "cvr_check.cu", "cvr_sharded.cu", "cvr_mm_reader.cpp"]
CLI_SOURCES = ["cli/spmv_cvr_main.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-fopenmp",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _run(cmd: list[str], verbose: bool) -> str:
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports wrapper compilers that cannot link OpenMP
    env.pop("CXX", None)
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    if verbose or p.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError(f"build step failed ({p.returncode}): {' '.join(cmd)}")
    return p.stdout


def build_library(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, "cvr_internal.h"), os.path.join(ROOT, "include", "cvr_b200.h"),
                   os.path.abspath(__file__)]
    if not force and _newer(LIB, deps):
        return LIB
    os.makedirs(LIB_DIR, exist_ok=True)
    extra = os.environ.get("CVR_NVCC_EXTRA", "").split()  # experiments, e.g. -DCVR_MIN_BLOCKS=8
    log = _run([_nvcc(), *NVCC_FLAGS, *extra, "-shared", "-o", LIB, *srcs, "-lgomp", "-ldl"], verbose)
    with open(os.path.join(LIB_DIR, "ptxas.log"), "w") as f:
        f.write(log)
    return LIB


def build_cli(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CLI_SOURCES]
    deps = srcs + [os.path.join(ROOT, "include", "cvr_b200.h"), LIB]
    if not force and _newer(CLI, deps):
        return CLI
    os.makedirs(BIN_DIR, exist_ok=True)
    _run(["g++", "-O2", "-std=c++17", "-Wall", "-fopenmp", "-I", os.path.join(ROOT, "include"),
          "-o", CLI, *srcs, "-L", LIB_DIR, "-lcvr_b200", "-Wl,-rpath,$ORIGIN/../lib"], verbose)
    return CLI


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_library(force, verbose)
    build_cli(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
