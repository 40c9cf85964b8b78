mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x -k "nms or NMS or detections or predict" 2>&1 | tail -5
for cfg in "v8n 32" "v8s 32" "v8x 8"; do set -- $cfg; python bench.py --steps 40 --warmup 5 --no-cpu-baseline --model $1 --batch $2 > gpurun_out/b23_$1.json 2> gpurun_out/b23_$1.err; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/b23_$1.json').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_net_tflops'])
except Exception as e: print('$1 ERR', open('gpurun_out/b23_$1.err').read()[-400:])
"; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms -c 12 --csv --log-file gpurun_out/nms23.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
grep -E "nms" gpurun_out/nms23.csv | awk -F'","' '{print $5, $(NF-0)}' | head -12
