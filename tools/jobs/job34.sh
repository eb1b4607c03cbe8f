cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tail_syn|k_ups_level_b" -s 12 -c 6 -o gpurun_out/r02_tail_v2 -f python tools/gpu_ncu_target.py > gpurun_out/ncu_r02_tail_v2.log 2>&1
tail -2 gpurun_out/ncu_r02_tail_v2.log
