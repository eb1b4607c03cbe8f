cd $GRAFT_REPO_ROOT
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 400 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err; cut -c1-300 gpurun_out/bench_ref_n1.json
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python tools/gpu_sanitize.py small > gpurun_out/r02_memcheck.out 2>&1; tail -2 gpurun_out/r02_memcheck.out; tail -2 gpurun_out/r02_memcheck.log
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_kodim14.log python tools/gpu_sanitize.py kodim14 > gpurun_out/r02_memcheck_kodim14.out 2>&1; tail -2 gpurun_out/r02_memcheck_kodim14.out; tail -2 gpurun_out/r02_memcheck_kodim14.log
timeout 300 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_synccheck.log python tools/gpu_sanitize.py small > gpurun_out/r02_synccheck.out 2>&1; tail -1 gpurun_out/r02_synccheck.out; tail -2 gpurun_out/r02_synccheck.log
