cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --workload kodak24_batch --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'], d.get('roofline',{}).get('kernel_ms'))"
