mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for cfg in "v8n 32" "v8s 32" "v8x 8" "v11n 32" "v8s-seg 16"; do set -- $cfg; python bench.py --steps 40 --warmup 5 --no-cpu-baseline --model $1 --batch $2 > gpurun_out/b26_$1.json 2> gpurun_out/b26_$1.err; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/b26_$1.json').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_net_tflops'], d['roofline']['frac'])
except Exception as e: print('$1 ERR', open('gpurun_out/b26_$1.err').read()[-400:])
"; done
python tools/profile_ops.py v8x 8 > gpurun_out/ops_v8x_26.txt 2>&1; tail -1 gpurun_out/ops_v8x_26.txt
YB_IN_DTYPE=u8 ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 1 -c 1 -f -o gpurun_out/stem26_u8 python tools/ncu_target.py v8n 32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 1 -c 1 -f -o gpurun_out/stem26_f16 python tools/ncu_target.py v8n 32 > /dev/null 2>&1
ls gpurun_out/*.ncu-rep
