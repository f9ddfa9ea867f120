timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for v in "" "PINMEM_B200_NO_READ_BRANCHES=1"; do
env $v timeout 300 python bench.py --no-cpu-baseline --no-extra --no-callers --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v step', round(d['ms_per_step'],4), d['timing']['value']['block_ms_min'], 'core', round(d['core']['ms_per_step'],4), d.get('cuda_graph'))
"
done
timeout 300 python bench.py --no-cpu-baseline --no-extra --no-callers --steps 50 --dtype bf16 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bf16 step', round(d['ms_per_step'],4), d['timing']['value']['block_ms_min'])
"
