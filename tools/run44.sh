for v in 0 1; do
if [ $v = 1 ]; then export YB_DEBUG_NO_PRIORITY=1; fi
for m in v8n v8s; do python bench.py --steps 40 --warmup 5 --no-cpu-baseline --model $m 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m no_priority=$v', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
done
