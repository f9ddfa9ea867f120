timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for v in "" "PINMEM_B200_BN_REDUCE_PER_CHANNEL=1"; do
env $v timeout 200 python bench.py --no-extra --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'step', round(d['ms_per_step'],4), 'core', round(d['core']['ms_per_step'],4), {k:(v.get('ms'), v.get('frac')) for k,v in d['kernels'].items() if 'bn_' in k})
"
done
