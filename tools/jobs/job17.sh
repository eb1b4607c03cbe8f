cd $GRAFT_REPO_ROOT
for m in 0x3777 0x7777 0x1777 0x0777 0x3333; do echo "== mask $m"; timeout 60 python tools/gpu_1080.py mask=$m 2>&1 | sed -n 3,3p | cut -c1-60; done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02_bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/r02_bench_under_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_entropy|k_tail_syn|k_ups_level_b" -s 22 -c 7 -o gpurun_out/r02_frame -f python tools/gpu_ncu_target.py > gpurun_out/ncu_r02_frame.log 2>&1
tail -3 gpurun_out/ncu_r02_frame.log
