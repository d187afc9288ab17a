set -x
mkdir -p gpurun_out/r02
for w in web fem road rmat24; do python tools/compare.py --workload $w --iters 20 > gpurun_out/r02/compare_$w.txt 2>gpurun_out/r02/compare_$w.err; tail -1 gpurun_out/r02/compare_$w.txt | cut -c1-600; done
python tools/compare.py --workload web --locality > gpurun_out/r02/locality_web.txt 2>&1; cat gpurun_out/r02/locality_web.txt | tail -6
python tools/compare.py --workload rmat:22 --locality > gpurun_out/r02/locality_rmat22.txt 2>&1; cat gpurun_out/r02/locality_rmat22.txt | tail -6
for w in rmat24 web road fem; do ncu --set full --clock-control none --import-source on -k regex:cvr_spmv_tile -s 3 -c 1 -o gpurun_out/r02/prof_final_$w -f python tools/kernel_ab.py --workloads $w --variants auto --steps 1 --out gpurun_out/r02/ab_ncu.jsonl > gpurun_out/r02/ncu_$w.log 2>&1; tail -1 gpurun_out/r02/ncu_$w.log; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches_bench_default.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cusparse > gpurun_out/r02/bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/r02/bench_under_ncu.log
python tests/cli_side_by_side.py 1000 > gpurun_out/r02/cli_side_by_side.log 2>&1; tail -25 gpurun_out/r02/cli_side_by_side.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_sharded.py -m gpu -x -q -k "iterated and dev00 and peer_sparse" > gpurun_out/r02/racecheck_sharded.log 2>&1; tail -8 gpurun_out/r02/racecheck_sharded.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generated and auto" > gpurun_out/r02/memcheck_parity.log 2>&1; tail -8 gpurun_out/r02/memcheck_parity.log
