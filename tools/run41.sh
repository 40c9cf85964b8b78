for i in 1 2; do python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>gpurun_out/b41.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu', d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('host_affinity'))"; done
tail -3 gpurun_out/b41.err
