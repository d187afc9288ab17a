#!/bin/bash
# last single-GPU pass of round 2 with the final library: parity suite, bench default + reference arm, launch list, CLI side by side
set -x
O=gpurun_out/r02g
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.json
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 200 $O/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cusparse > $O/bench_under_ncu.log 2>&1
python tests/cli_side_by_side.py 1000 > $O/cli_side_by_side.log 2>&1; tail -14 $O/cli_side_by_side.log
python tools/convert_probe.py 2>&1 | tail -5
