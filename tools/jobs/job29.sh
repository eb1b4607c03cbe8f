cd $GRAFT_REPO_ROOT
CCD_LIB=cool-chic_b200/csrc/libccdec_prof.so timeout 100 python tools/gpu_prof.py 2>&1 | tail -2
