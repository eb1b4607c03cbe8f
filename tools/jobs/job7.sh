cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_entropy -s 2 -c 1 -o gpurun_out/r02_entropy_v7 -f python tools/gpu_1080.py > gpurun_out/ncu_v7.log 2>&1
tail -2 gpurun_out/ncu_v7.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; tail -8 gpurun_out/pytest_gpu.txt
