summ() { python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$1 N=%d'%d['n_gpus'], 'gflops', round(d['value'],1), 'step_ms', round(d['ms_per_step'],4), 'kernel_us', round(r['kernel_us'],1), 'sent', d['extra']['peer_bytes_sent_per_step_rank0'])"; }
python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2
for w in fem rmat:22 road; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --workload $w 2>/tmp/err | grep -E '^\{' | tail -1 | summ "sparse $w"; grep -E "Error|error" /tmp/err | tail -2
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --workload road --dense-exchange 2>/tmp/err | grep -E '^\{' | tail -1 | summ "dense road"
