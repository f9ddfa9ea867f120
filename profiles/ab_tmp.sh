timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for i in 1 2; do
timeout 200 python bench.py --no-extra --no-cpu-baseline --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('step', round(d['ms_per_step'],4), d['timing']['value']['block_ms_min'], 'core', round(d['core']['ms_per_step'],4), 'graph kernels', d.get('cuda_graph'))
"
done
