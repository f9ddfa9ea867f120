timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in "" "PM_GEMM_NO_ROWS32=1"; do
env $v timeout 300 python bench.py --no-cpu-baseline --no-extra --no-callers --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v step', round(d['ms_per_step'],4), d['timing']['value']['block_ms_min'], {k:(v.get('ms'), v.get('launches_per_step')) for k,v in d['kernels'].items() if 'conv1x1' in k})
"
done
