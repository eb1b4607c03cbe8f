cd $GRAFT_REPO_ROOT
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; tail -c 300 gpurun_out/bench_n8.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().split('\n')[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'], d['config'].get('streams_per_gpu'), d.get('clocks'))"
