mkdir -p gpurun_out
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_18.txt 2>&1
python tools/profile_ops.py v8s 32 > gpurun_out/ops_v8s_18.txt 2>&1
python tools/profile_ops.py v8x 8 > gpurun_out/ops_v8x_18.txt 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel -s 59 -c 59 --csv --log-file gpurun_out/conv_traffic.csv python tools/ncu_target.py v8n 32 > /dev/null 2>&1
tail -2 gpurun_out/conv_traffic.csv
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches18.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b18_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_tc -s 1 -c 1 -f -o gpurun_out/stem18 python tools/ncu_target.py v8n 32 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 59 -c 5 -f -o gpurun_out/conv18 python tools/ncu_target.py v8n 32 > /dev/null 2>&1
ls -la gpurun_out | tail -8
