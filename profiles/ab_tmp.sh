for v in "PINMEM_B200_READLOSS_GEN=4" "PINMEM_B200_READLOSS_GEN=3"; do
for lab in blocky iid; do
env $v timeout 200 python bench.py --no-extra --no-cpu-baseline --no-callers --labels $lab --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v $lab', 'step', round(d['ms_per_step'],4), 'core', round(d['core']['ms_per_step'],4), round(d['core']['frac_of_peak'],4), {k:(v.get('ms')) for k,v in d['kernels'].items() if 'readloss' in k})
"
done; done
