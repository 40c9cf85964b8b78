timeout 300 python -m pytest tests -q -m gpu -x -k "nms or NMS or detections or predict" 2>&1 | tail -2
python tools/exp_nms_time.py v8n 32 2>&1 | tail -1
python tools/exp_nms_time.py v8s 32 2>&1 | tail -1
for i in 1 2; do python bench.py --steps 40 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v8n', d['value'], d['ms_per_step'], d['e2e']['value'])"; done
