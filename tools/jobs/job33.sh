cd $GRAFT_REPO_ROOT
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err; cut -c1-300 gpurun_out/bench_n1.json
timeout 400 python bench.py --workload gop32_1080p_yuv420 --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_gop_n1.json 2> gpurun_out/bench_gop_n1.err; tail -c 300 gpurun_out/bench_gop_n1.err; cut -c1-200 gpurun_out/bench_gop_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -c 200 gpurun_out/r02_bench_under_ncu.log
