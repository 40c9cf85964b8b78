mkdir -p gpurun_out
python tools/exp_nms_time.py v8n 32 2>&1 | tail -8
YB_DEBUG_NMS_GENERAL=1 python tools/exp_nms_time.py v8n 32 2>&1 | tail -1
python tools/exp_nms_time.py v8s 32 2>&1 | tail -7
for v in 0 1; do
if [ $v = 1 ]; then export YB_DEBUG_NMS_GENERAL=1; fi
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --model v8n --batch 32 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v8n general=$v', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
unset YB_DEBUG_NMS_GENERAL
python tools/exp_fixed_cost.py 2>&1 | tail -6
YB_DEBUG_NO_PDL=1 python tools/exp_fixed_cost.py 2>&1 | tail -6
