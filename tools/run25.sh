mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for i in 1 2 3; do
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --model v8n --batch 32 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('v8n run$i', d['value'], d['ms_per_step'], d['e2e']['value'])"
done
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --model v8n --batch 32 > gpurun_out/b25_v8n.json 2> gpurun_out/b25_v8n.err; tail -c 300 gpurun_out/b25_v8n.err
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_25.txt 2>&1
python tools/profile_ops.py v8n 64 > gpurun_out/ops_v8n_25_b64.txt 2>&1
python tools/profile_ops.py v8n 32 u8 2>&1 | grep -E "model.0 " | cut -c1-120
python tools/ops_marginal.py gpurun_out/ops_v8n_25.txt gpurun_out/ops_v8n_25_b64.txt > gpurun_out/marginal_v8n.txt; tail -3 gpurun_out/marginal_v8n.txt
