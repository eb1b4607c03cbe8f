cd $GRAFT_REPO_ROOT
for lib in libccdec.so libccdec_reldone.so; do
  echo "== $lib"; CCD_LIB=cool-chic_b200/csrc/$lib timeout 45 python tools/gpu_1080.py 2>&1 | sed -n 3,4p | cut -c1-100
  CCD_LIB=cool-chic_b200/csrc/$lib timeout 40 python tools/gpu_stress.py 6 | tail -1 | cut -c1-80
done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -4 gpurun_out/pytest_gpu.txt
