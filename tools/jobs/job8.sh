cd $GRAFT_REPO_ROOT
for lib in libccdec.so libccdec_prof.so; do
  echo "== $lib"; CCD_LIB=cool-chic_b200/csrc/$lib timeout 45 python tools/gpu_1080.py 2>&1 | sed -n 3,6p | cut -c1-330
done
echo "== kodim14"; timeout 40 python tools/gpu_stress.py 6 | tail -2
