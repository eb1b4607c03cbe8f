cd $GRAFT_REPO_ROOT
for i in 1 2; do
  for lib in libccdec.so libccdec_vote.so libccdec_prev.so; do
    echo -n "$lib: "; CCD_LIB=cool-chic_b200/csrc/$lib timeout 100 python tools/gpu_1080.py 2>&1 | sed -n 3,3p | cut -c1-60
  done
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -3 gpurun_out/pytest_gpu.txt
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python tools/gpu_sanitize.py small > gpurun_out/r02_memcheck.out 2>&1; tail -2 gpurun_out/r02_memcheck.out; tail -1 gpurun_out/r02_memcheck.log
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_kodim14.log python tools/gpu_sanitize.py kodim14 > gpurun_out/r02_memcheck_kodim14.out 2>&1; tail -1 gpurun_out/r02_memcheck_kodim14.out; tail -1 gpurun_out/r02_memcheck_kodim14.log
timeout 300 compute-sanitizer --tool synccheck --log-file gpurun_out/r02_synccheck.log python tools/gpu_sanitize.py small > gpurun_out/r02_synccheck.out 2>&1; tail -1 gpurun_out/r02_synccheck.out; tail -1 gpurun_out/r02_synccheck.log
