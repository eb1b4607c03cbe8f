cd $GRAFT_REPO_ROOT
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 600 gpurun_out/bench_n1.err; cut -c1-3000 gpurun_out/bench_n1.json
timeout 300 python bench.py --workload kodak24_batch --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_batch_n1.json 2> gpurun_out/bench_batch_n1.err; tail -c 600 gpurun_out/bench_batch_n1.err; cut -c1-1500 gpurun_out/bench_batch_n1.json
timeout 300 python bench.py --workload gop32_1080p_yuv420 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_gop_n1.json 2> gpurun_out/bench_gop_n1.err; tail -c 600 gpurun_out/bench_gop_n1.err; cut -c1-1500 gpurun_out/bench_gop_n1.json
