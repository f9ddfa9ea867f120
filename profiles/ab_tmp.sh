timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --workload cfg5_dr101v2_eval_b1 --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg5', d['value'], d['ms_per_step'], d.get('cuda_graph'), {k:v.get('ms') for k,v in d['kernels'].items()})"
PINMEM_B200_NO_INFER_BRANCH=1 python bench.py --workload cfg5_dr101v2_eval_b1 --steps 50 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg5 nobranch', d['value'], d['ms_per_step'])"
timeout 200 python bench.py --no-extra --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('step', round(d['ms_per_step'],4), 'core', round(d['core']['ms_per_step'],4), d['roofline']['frac'], d['roofline'].get('executed_frac'), {k:(v.get('ms'), v.get('launches_per_step')) for k,v in d['kernels'].items() if 'conv1x1' in k})
"
