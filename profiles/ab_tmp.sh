timeout 300 python -m pytest tests -m gpu -q -x -k "operand_path" 2>&1 | tail -12; timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
for i in 1 2; do
for v in "" "PINMEM_B200_NO_BNBWD_FUSION=1"; do
env $v timeout 200 python bench.py --no-extra --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', round(d['ms_per_step'],4), d['timing']['value']['block_ms_min'], d['timing']['value']['block_ms_max'], {k:(v.get('ms'), v.get('launches_per_step')) for k,v in d['kernels'].items() if 'conv1x1' in k or 'bn_bwd' in k})
"
done; done
