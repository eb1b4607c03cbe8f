cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -c 800 gpurun_out/bench_n2.err; cut -c1-1800 gpurun_out/bench_n2.json
timeout 300 $TR bench.py --gpus 2 --workload gop32_1080p_yuv420 --steps 2 --warmup 3 > gpurun_out/bench_gop_n2.json 2> gpurun_out/bench_gop_n2.err; tail -c 800 gpurun_out/bench_gop_n2.err; cut -c1-600 gpurun_out/bench_gop_n2.json
timeout 200 $TR tools/gpu_gop_sharded.py 8 2>&1 | tail -6
