cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "kodim14_through or synthetic_stream or many_streams or corrupt or gop_decode or optional_synth or device_range" > gpurun_out/pytest_spec.txt 2>&1; tail -4 gpurun_out/pytest_spec.txt
for i in 1 2; do
  for lib in libccdec.so libccdec_vote.so libccdec_blocks.so; do
    echo -n "$lib: "; CCD_LIB=cool-chic_b200/csrc/$lib timeout 100 python tools/gpu_1080.py 2>&1 | sed -n 3,3p | cut -c1-60
  done
done
CCD_LIB=cool-chic_b200/csrc/libccdec.so timeout 100 python tools/gpu_1080.py 2>&1 | grep "round trip"
for lib in libccdec.so libccdec_vote.so libccdec_blocks.so; do echo -n "$lib kodim14: "; CCD_LIB=cool-chic_b200/csrc/$lib timeout 60 python tools/gpu_stress.py 6 | tail -1 | cut -c1-80; done
CCD_LIB=cool-chic_b200/csrc/libccdec_prof.so timeout 100 python tools/gpu_1080.py 2>&1 | sed -n 3,4p
