cd $GRAFT_REPO_ROOT
for v in 1 2; do echo "== CCD_TAIL_V=$v"; CCD_TAIL_V=$v timeout 60 python tools/gpu_ncu_target.py 2>&1 | tail -1 | cut -c1-200; done
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; tail -12 gpurun_out/pytest_gpu.txt
