cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_entropy -s 1 -c 1 -o gpurun_out/r02_entropy_v6 -f python tools/gpu_1080.py > gpurun_out/ncu_v6.log 2>&1
tail -3 gpurun_out/ncu_v6.log
ls -la gpurun_out/*.ncu-rep
