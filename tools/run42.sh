mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -2
for cfg in "v8n 32" "v8s 32" "v8x 8"; do set -- $cfg; python bench.py --steps 40 --warmup 5 --no-cpu-baseline --model $1 --batch $2 > gpurun_out/b42_$1.json 2> gpurun_out/b42_$1.err; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/b42_$1.json').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_net_tflops'], d['roofline']['frac'])
except Exception as e: print('$1 ERR', open('gpurun_out/b42_$1.err').read()[-400:])
"; done
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_42.txt 2>&1; tail -1 gpurun_out/ops_v8n_42.txt
