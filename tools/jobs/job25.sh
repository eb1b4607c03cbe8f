cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_entropy -s 3 -c 1 -o gpurun_out/r02_entropy_spec -f python tools/gpu_ncu_target.py > gpurun_out/ncu_r02_entropy_spec.log 2>&1
tail -2 gpurun_out/ncu_r02_entropy_spec.log
