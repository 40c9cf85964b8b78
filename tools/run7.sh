mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/pytest7.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -c 900 gpurun_out/bench7.json; tail -3 gpurun_out/bench7.err
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_7.txt 2>&1
python tools/profile_ops.py v8x 8 > gpurun_out/ops_v8x_7.txt 2>&1
python tools/profile_ops.py v8s 32 > gpurun_out/ops_v8s_7.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $((62+44)) -c 1 -f -o gpurun_out/ncu7_v8n_tc44 python tools/ncu_target.py v8n 32 > gpurun_out/ncu7_log_44.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:nms -c 4 --csv --log-file gpurun_out/nms7.csv python tools/ncu_target.py v8n 32 > /dev/null 2>&1
