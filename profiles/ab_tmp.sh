timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('step', round(d['ms_per_step'],4), 'core', round(d['core']['ms_per_step'],4), round(d['core']['frac_of_peak'],4), {k:(v.get('ms')) for k,v in d['kernels'].items() if 'readloss' in k}, d.get('callers'), {k:(v.get('ms_per_step'), v.get('ms_per_image')) for k,v in d.get('configs',{}).items()})
"
