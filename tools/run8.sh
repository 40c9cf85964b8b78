mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/pytest8.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; tail -c 900 gpurun_out/bench8.json; tail -3 gpurun_out/bench8.err
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_8.txt 2>&1
python tools/profile_ops.py v8x 8 > gpurun_out/ops_v8x_8.txt 2>&1
python tools/profile_ops.py v8s 32 > gpurun_out/ops_v8s_8.txt 2>&1
