timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
PINMEM_B200_NO_READ_BRANCHES=1 timeout 600 python -m pytest tests -m gpu -q -x -k "graphed or edges or sharded" 2>&1 | tail -2
for i in 1 2 3; do
for v in "" "PINMEM_B200_NO_READ_BRANCHES=1"; do
env $v timeout 300 python bench.py --no-cpu-baseline --no-extra --no-callers --steps 50 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v step', round(d['ms_per_step'],4), d['timing']['value']['block_ms_min'], d['timing']['value']['block_ms_max'])
"
done; done
