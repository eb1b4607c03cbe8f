cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "full_size or fused or 148_streams or optional or kodim14_through or synthetic_stream" 2>&1 | tail -3
timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1])
print('value',d['value'],'ms',d['ms_per_step']); print(d['roofline_synthesis']); print({k:d['many_streams'][k] for k in ('ms','entropy_ms','synthesis_ms','device_value')}); print({k:d['scale_workload_at_n1'][k] for k in ('ms','entropy_ms','synthesis_ms','device_value')})"
