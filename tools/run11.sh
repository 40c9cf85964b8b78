mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/pytest11.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; tail -c 300 gpurun_out/bench11.json; tail -3 gpurun_out/bench11.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --model v8x --batch 8 > gpurun_out/bench11x.json 2> gpurun_out/bench11x.err
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --model v8s --batch 32 > gpurun_out/bench11s.json 2> gpurun_out/bench11s.err
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_11.txt 2>&1
