cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -15 gpurun_out/pytest_gpu.txt
for lib in libccdec.so libccdec_k2.so libccdec_prof.so; do
  echo "== $lib"; CCD_LIB=cool-chic_b200/csrc/$lib timeout 120 python tools/gpu_1080.py 2>&1 | sed -n 1,10p
done
echo "== kodim14"; timeout 60 python tools/gpu_stress.py 6 | tail -2
