mkdir -p gpurun_out
for cfg in "v8n 32" "v8s 32" "v8x 8" "v11n 32" "v11s 32" "v8s-seg 16"; do set -- $cfg; python bench.py --steps 30 --warmup 5 --no-cpu-baseline --model $1 --batch $2 > gpurun_out/b17_$1.json 2> gpurun_out/b17_$1.err; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/b17_$1.json').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_net_tflops'])
except Exception as e: print('$1 ERR', open('gpurun_out/b17_$1.err').read()[-400:])
"; done
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -s 62 -c 62 --csv --log-file gpurun_out/conv_traffic.csv python tools/ncu_target.py v8n 32 > /dev/null 2>&1
tail -2 gpurun_out/conv_traffic.csv
