cd $GRAFT_REPO_ROOT
for lib in libccdec.so libccdec_v3.so libccdec_v3k2.so libccdec_prof.so; do
  echo "== $lib"; CCD_LIB=cool-chic_b200/csrc/$lib timeout 120 python tools/gpu_1080.py 2>&1 | sed -n 2,7p | cut -c1-330
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_entropy -s 2 -c 1 -o gpurun_out/r02_entropy_v6b -f python tools/gpu_1080.py > gpurun_out/ncu_v6b.log 2>&1
tail -2 gpurun_out/ncu_v6b.log
