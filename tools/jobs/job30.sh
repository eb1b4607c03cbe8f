cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_decode.py -m gpu -q -x -k "148_streams or many_streams or corrupt" 2>&1 | tail -3
for n in 0 1; do
  echo "== CCD_NARROW_CTA=$n"
  CCD_NARROW_CTA=$n timeout 300 python bench.py --workload kodak24_batch --steps 3 --warmup 3 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'], d.get('roofline',{}).get('kernel_ms'))"
done
