mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/pytest9.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -c 700 gpurun_out/bench9.json; tail -3 gpurun_out/bench9.err
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_9.txt 2>&1
python tools/profile_ops.py v8x 8 > gpurun_out/ops_v8x_9.txt 2>&1
python tools/profile_ops.py v8s 32 > gpurun_out/ops_v8s_9.txt 2>&1
for idx in 0 1; do
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $((62+idx)) -c 1 -f -o gpurun_out/ncu9_v8n_tc$idx python tools/ncu_target.py v8n 32 > gpurun_out/ncu9_log_$idx.txt 2>&1
done
