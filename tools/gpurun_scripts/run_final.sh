# Round-end evidence: tests, bench lines of every model, per-op tables, ncu launch list / traffic / full captures.
mkdir -p gpurun_out/final
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for cfg in "v8n 32" "v8s 32" "v8x 8" "v11n 32" "v11s 32" "v8s-seg 16"; do set -- $cfg
  python bench.py --steps 40 --warmup 5 --no-cpu-baseline --model $1 --batch $2 > gpurun_out/final/bench_$1.json 2> gpurun_out/final/bench_$1.err
  python -c "
import json
try:
    d=json.loads(open('gpurun_out/final/bench_$1.json').read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_net_tflops'], d['roofline']['frac'])
except Exception as e: print('$1 ERR', open('gpurun_out/final/bench_$1.err').read()[-300:])
"; done
python bench.py > gpurun_out/final/bench_default.json 2> gpurun_out/final/bench_default.err; tail -c 600 gpurun_out/final/bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_reference.json 2>/dev/null; tail -c 400 gpurun_out/final/bench_reference.json
python tools/profile_ops.py v8n 32 > gpurun_out/final/ops_v8n_b32.txt 2>&1
python tools/profile_ops.py v8s 32 > gpurun_out/final/ops_v8s_b32.txt 2>&1
python tools/profile_ops.py v8x 8 > gpurun_out/final/ops_v8x_b8.txt 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -s 59 -c 59 --csv --log-file gpurun_out/final/conv_traffic.csv python tools/ncu_target.py v8n 32 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/final/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 59 -c 6 -f -o gpurun_out/final/conv_tc python tools/ncu_target.py v8n 32 > /dev/null 2>&1
ls -la gpurun_out/final | tail -20
