mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/pytest10.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -c 500 gpurun_out/bench10.json; tail -3 gpurun_out/bench10.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --model v8x --batch 8 > gpurun_out/bench10x.json 2> gpurun_out/bench10x.err; tail -c 300 gpurun_out/bench10x.json
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --model v8s --batch 32 > gpurun_out/bench10s.json 2> gpurun_out/bench10s.err; tail -c 300 gpurun_out/bench10s.json
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_10.txt 2>&1
python tools/profile_ops.py v8x 8 > gpurun_out/ops_v8x_10.txt 2>&1
python tools/profile_ops.py v8s 32 > gpurun_out/ops_v8s_10.txt 2>&1
