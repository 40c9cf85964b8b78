mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -12 | tee gpurun_out/pytest6.log
python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err; tail -c 900 gpurun_out/bench6.json; tail -3 gpurun_out/bench6.err
python tools/profile_ops.py v8n 32 > gpurun_out/ops_v8n_6.txt 2>&1
python tools/ncu_target.py v8n 32 list 2>&1 | tail -2 > gpurun_out/ncu_order.txt
for idx in 2 44 30 47; do
  ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s $((62+idx)) -c 1 -f -o gpurun_out/ncu_v8n_tc$idx python tools/ncu_target.py v8n 32 > gpurun_out/ncu_log_$idx.txt 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:nms_kernel -c 1 -f -o gpurun_out/ncu_v8n_nms python tools/ncu_target.py v8n 32 > gpurun_out/ncu_log_nms.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:stem_kernel -s 1 -c 1 -f -o gpurun_out/ncu_v8n_stem python tools/ncu_target.py v8n 32 > gpurun_out/ncu_log_stem.txt 2>&1
ls -la gpurun_out/*.ncu-rep
