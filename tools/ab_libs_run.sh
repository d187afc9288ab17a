#!/bin/bash
# A/B of the library variants under tools/ab_libs on the GPU box: tools/ab_libs_run.sh "<workloads>" <env-var> <values> lib1 lib2 ...
wl=$1; envn=$2; vals=$3; shift 3
mkdir -p gpurun_out/ab
cp cvr_b200/lib/libcvr_b200.so /tmp/libcvr_keep.so
for lib in "$@"; do
  cp tools/ab_libs/$lib.so cvr_b200/lib/libcvr_b200.so
  python tools/kernel_ab.py --workloads $wl --env $envn --variants $vals --out gpurun_out/ab/$lib.jsonl 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('$lib', d['workload'], d['variant'], d['kernel'], 'chunks', d['chunks'], 'kernel_us %.1f' % d['kernel_us'], 'frac %.3f' % d['frac'], 'bad', d['rows_failing'])
"
done
cp /tmp/libcvr_keep.so cvr_b200/lib/libcvr_b200.so
