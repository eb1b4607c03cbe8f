cd $GRAFT_REPO_ROOT
for i in 1 2; do
  for lib in libccdec.so libccdec_vote.so; do
    echo -n "$lib: "; CCD_LIB=cool-chic_b200/csrc/$lib timeout 100 python tools/gpu_1080.py 2>&1 | sed -n 3,3p | cut -c1-60
  done
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -4 gpurun_out/pytest_gpu.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_entropy -s 3 -c 1 -o gpurun_out/r02_entropy_spec -f python tools/gpu_ncu_target.py > gpurun_out/ncu_r02_entropy_spec.log 2>&1
tail -2 gpurun_out/ncu_r02_entropy_spec.log
