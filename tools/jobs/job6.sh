cd $GRAFT_REPO_ROOT
for lib in libccdec.so libccdec_k3.so libccdec_k6.so libccdec_k8.so libccdec_prof.so; do
  echo "== $lib"; CCD_LIB=cool-chic_b200/csrc/$lib timeout 120 python tools/gpu_1080.py 2>&1 | sed -n 3,5p | cut -c1-330
done
echo "== kodim14"; timeout 60 python tools/gpu_stress.py 6 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -5 gpurun_out/pytest_gpu.txt
