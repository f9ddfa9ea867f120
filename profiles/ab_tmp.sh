python profiles/readloss_time.py
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline --no-extra --no-callers --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('step', round(d['ms_per_step'],4), d['timing']['value']['block_ms_min'], 'core', round(d['core']['ms_per_step'],4), round(d['core']['frac_of_peak'],4), {k:(v.get('ms')) for k,v in d['kernels'].items() if 'readloss_fwd8' == k})
"
