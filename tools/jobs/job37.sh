cd $GRAFT_REPO_ROOT
for i in 1 2; do
  for lib in libccdec.so libccdec_s27.so libccdec_prev.so libccdec_blocks.so; do
    echo -n "$lib: "; CCD_LIB=cool-chic_b200/csrc/$lib timeout 100 python tools/gpu_1080.py 2>&1 | sed -n 3,3p | cut -c1-60
  done
done
for lib in libccdec.so libccdec_s27.so libccdec_prev.so; do echo -n "$lib kodim14: "; CCD_LIB=cool-chic_b200/csrc/$lib timeout 60 python tools/gpu_stress.py 6 | tail -1 | cut -c1-80; done
