set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc; lscpu | grep 'Model name'
CCD_LIB=cool-chic_b200/csrc/libccdec_prof.so timeout 200 python tools/gpu_1080.py > gpurun_out/prof1080.txt 2>&1
CCD_LIB=cool-chic_b200/csrc/libccdec_prof.so timeout 100 python tools/gpu_prof.py > gpurun_out/prof_kodim.txt 2>&1
timeout 300 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck_small.log python tools/gpu_sanitize.py small > gpurun_out/memcheck_small.out 2>&1
timeout 400 compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck_small.log python tools/gpu_sanitize.py small > gpurun_out/racecheck_small.out 2>&1
timeout 600 compute-sanitizer --tool racecheck --log-file gpurun_out/racecheck_kodim14.log python tools/gpu_sanitize.py kodim14 > gpurun_out/racecheck_kodim14.out 2>&1
tail -3 gpurun_out/*.out gpurun_out/*.log
