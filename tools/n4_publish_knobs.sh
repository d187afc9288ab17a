mkdir -p gpurun_out/r02d
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) bench.py --gpus 4 --steps 20 --warmup 5 --rebalance 0 "$@" > gpurun_out/r02d/n4_$tag.json 2> gpurun_out/r02d/n4_$tag.err; python -c "
import json; d=json.loads(open('gpurun_out/r02d/n4_$tag.json').read().strip().splitlines()[-1]); print('$tag', round(d['value'],1), round(d['ms_per_step']*1e3,1), [round(r[0]) for r in d['extra']['per_rank_sweep_us_rows_nnz_records']], d['parity']['rows_failing'])
"; }
run base
CVR_PUSH_MIN_ROWS=256 run push256
CVR_PUSH_MIN_ROWS=1024 run push1024
CVR_SCATTER_FACTOR=2 run scat2
CVR_SCATTER_FACTOR=16 run scat16
run nccl --exchange nccl
