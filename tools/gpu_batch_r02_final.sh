#!/bin/bash
# Final single-GPU evidence batch of round 2 (run under gpurun): parity suite, A/B of the chunk size, bench default,
# reference arm, ncu launch list of the bench command, one `ncu --set full` capture of the sweep per workload,
# the conversion kernels' launch list, the CLI side by side with the reference, compute-sanitizer on two parity cases.
set -x
O=gpurun_out/r02h
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 400 $O/bench_default.json
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 300 $O/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cusparse > $O/bench_under_ncu.log 2>&1
for w in rmat24 web road fem; do ncu --set full --clock-control none --import-source on -k regex:cvr_spmv_tile -s 3 -c 1 -o $O/prof_final_$w -f python tools/kernel_ab.py --workloads $w --variants auto --steps 1 --out $O/ab_ncu.jsonl > $O/ncu_$w.log 2>&1; tail -1 $O/ncu_$w.log; done
python tests/cli_side_by_side.py 1000 > $O/cli_side_by_side.log 2>&1; tail -12 $O/cli_side_by_side.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "iterated and dev00 and peer_sparse" > $O/racecheck_sharded.log 2>&1; tail -4 $O/racecheck_sharded.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generated and auto" > $O/memcheck_parity.log 2>&1; tail -4 $O/memcheck_parity.log
