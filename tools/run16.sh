timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
python tools/exp_timeline.py 44 1 2>&1 | tail -28
python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v8n', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['kernel_ms_per_step'])"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --model v8x --batch 8 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v8x', d['value'], d['ms_per_step'], d['roofline']['whole_net_tflops'])"
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --model v8s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v8s', d['value'], d['ms_per_step'], d['roofline']['whole_net_tflops'])"
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_16.txt 2>&1
python tools/profile_ops.py v8x 8 > gpurun_out/ops_v8x_16.txt 2>&1
