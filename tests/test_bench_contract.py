"""CPU suite: the bench.py contract that can be checked without a GPU -- the reference arm prints exactly
one JSON line with the agreed keys, non-zero ranks of a multi-rank reference run exit 0 without work, and
our own arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, env=e, timeout=600)


def test_reference_arm_prints_one_json_line():
    p = run(["--impl", "reference", "--workload", "tiny", "--steps", "3", "--warmup", "1"])
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "spmv_gflops" and d["unit"] == "GFLOP/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 3
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_do_nothing():
    p = run(["--impl", "reference", "--workload", "tiny", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run(["--workload", "tiny", "--steps", "1"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
